"""Batched solves (BASELINE config 3: many small programs with PSD + second-order + linear cones):
every program of a CONEXB200_BatchMaximize call must reproduce what the CPU oracle — and the
single-program device path — compute for that program alone: same solved flag, iteration count
within +-1, objectives within 1e-7 relative (or the oracle's own one-ulp sensitivity where larger),
y within 1e-6.
"""
import numpy as np
import pytest

from harness import Batch, add_cones, oracle, small_multicone_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    import devlib
    return oracle(), devlib.product()


SHAPES = {
    "tiny": dict(m=5, psd_blocks=1, psd_order=4, soc_cones=1, soc_order=3, lp_rows=6),
    "mixed": dict(m=8, psd_blocks=2, psd_order=6, soc_cones=2, soc_order=4, lp_rows=8),
    "c3": dict(m=40, psd_blocks=3, psd_order=20, soc_cones=2, soc_order=10, lp_rows=40),
    "psd_only": dict(m=6, psd_blocks=2, psd_order=7, soc_cones=0, soc_order=1, lp_rows=0),
    "lp_soc": dict(m=6, psd_blocks=0, psd_order=1, soc_cones=1, soc_order=5, lp_rows=9),
}


def oracle_reference(O, cones, b, cfg_kw):
    P = O.program()
    add_cones(P, cones)
    solved, y = P.maximize(b, O.default_config(**cfg_kw))
    log = P.iteration_log()
    # one-ulp perturbations of b: the oracle's own rounding sensitivity of the objectives
    spread = dict(by=0.0, cx=0.0)
    for j in (-2, -1, 1, 2):
        Pp = O.program()
        add_cones(Pp, cones)
        Pp.maximize(b * (1 + j * 2.0 ** -52), O.default_config(**cfg_kw))
        lp = Pp.iteration_log()
        if len(lp) == len(log):
            for key in spread:
                spread[key] = max(spread[key], abs(lp[-1][key] - log[-1][key]))
    return solved, y, log, spread


@pytest.mark.parametrize("shape,count", [("tiny", 7), ("mixed", 16), ("psd_only", 5), ("lp_soc", 5), ("c3", 6)])
def test_batch_matches_oracle_per_program(libs, shape, count):
    O, D = libs
    kw = SHAPES[shape]
    cfg_kw = dict(prepare_dual_variables=1)
    problems = [small_multicone_problem(1000 + p, **kw) for p in range(count)]
    programs = []
    for cones, _ in problems:
        P = D.program()
        add_cones(P, cones)
        programs.append(P)
    batch = Batch(D, programs)
    b = np.stack([pb for _, pb in problems])
    solved, y = batch.maximize(b, D.default_config(**cfg_kw))
    its, by, cx, k = batch.results()
    for p, (cones, pb) in enumerate(problems):
        so, yo, lo, spread = oracle_reference(O, cones, pb, cfg_kw)
        assert solved[p] == so, p
        assert abs(int(its[p]) - len(lo)) <= 1, (p, its[p], len(lo))
        if int(its[p]) == len(lo):
            for name, got in (("by", by[p]), ("cx", cx[p])):
                tol = max(1e-7 * max(1.0, abs(lo[-1][name])), 10 * spread[name])
                assert abs(got - lo[-1][name]) <= tol, (p, name, got, lo[-1][name], spread[name])
        assert np.abs(y[p] - yo).max() <= 1e-6 * max(1.0, np.abs(yo).max()), p
    assert batch.milliseconds() > 0 and len(batch.step_milliseconds()) == its.max()


def test_batch_matches_single_program_device_path(libs):
    """The same programs through CONEX_Maximize one at a time (large-block kernels for the PSD cones)
    and through the batch (small-block kernels): same answers."""
    _, D = libs
    kw = SHAPES["mixed"]
    problems = [small_multicone_problem(2000 + p, **kw) for p in range(4)]
    programs, singles = [], []
    for cones, pb in problems:
        P = D.program()
        add_cones(P, cones)
        programs.append(P)
        Q = D.program()
        add_cones(Q, cones)
        s, y = Q.maximize(pb)
        singles.append((s, y, Q.iteration_log()))
    batch = Batch(D, programs)
    solved, y = batch.maximize(np.stack([pb for _, pb in problems]))
    its, by, cx, _ = batch.results()
    for p, (s1, y1, log) in enumerate(singles):
        assert solved[p] == s1 and abs(int(its[p]) - len(log)) <= 1
        assert np.abs(y[p] - y1).max() <= 1e-6 * max(1.0, np.abs(y1).max())
        assert abs(by[p] - log[-1]["by"]) <= 1e-7 * max(1.0, abs(by[p]))


def test_programs_finish_at_different_iterations(libs):
    """Masking: programs that terminate early must not be touched by later launches."""
    O, D = libs
    kw = SHAPES["tiny"]
    problems = [small_multicone_problem(3000 + p, **kw) for p in range(12)]
    # different objective scales -> different iteration counts
    for p in range(12):
        problems[p] = (problems[p][0], problems[p][1] * (1.0 + 3.0 * (p % 4)))
    programs = []
    for cones, _ in problems:
        P = D.program()
        add_cones(P, cones)
        programs.append(P)
    batch = Batch(D, programs)
    solved, y = batch.maximize(np.stack([pb for _, pb in problems]))
    its, by, cx, _ = batch.results()
    ref_its = []
    for p, (cones, pb) in enumerate(problems):
        P = O.program()
        add_cones(P, cones)
        so, yo = P.maximize(pb)
        ref_its.append(len(P.iteration_log()))
        assert solved[p] == so
        assert np.abs(y[p] - yo).max() <= 1e-6 * max(1.0, np.abs(yo).max())
    assert np.abs(its - np.array(ref_its)).max() <= 1


def test_batch_dual_variables(libs):
    O, D = libs
    kw = SHAPES["mixed"]
    problems = [small_multicone_problem(4000 + p, **kw) for p in range(3)]
    programs = []
    for cones, _ in problems:
        P = D.program()
        add_cones(P, cones)
        programs.append(P)
    batch = Batch(D, programs)
    cfg_kw = dict(prepare_dual_variables=1)
    solved, y = batch.maximize(np.stack([pb for _, pb in problems]), D.default_config(**cfg_kw))
    for p, (cones, pb) in enumerate(problems):
        P = O.program()
        add_cones(P, cones)
        P.maximize(pb, O.default_config(**cfg_kw))
        for ci, (kind, A, c) in enumerate(cones):
            if kind == "soc":
                continue  # the reference's SOC dual variable is an unset dummy (workspace_soc.h:44)
            xo = P.dual_variable(ci)
            xd = batch.dual_variable(p, ci, xo.size).reshape(xo.shape, order="F")
            assert np.abs(xo - xd).max() <= 1e-5 * max(1.0, np.abs(xo).max()), (p, ci)


def test_batch_rejects_mismatched_structure(libs):
    _, D = libs
    a = D.program()
    add_cones(a, small_multicone_problem(1, **SHAPES["tiny"])[0])
    b = D.program()
    add_cones(b, small_multicone_problem(2, **SHAPES["mixed"])[0])
    import ctypes as C
    L = D.lib
    L.CONEXB200_CreateBatch.restype = C.c_void_p
    L.CONEXB200_CreateBatch.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    handles = (C.c_void_p * 2)(a.h, b.h)
    assert not L.CONEXB200_CreateBatch(handles, 2)
