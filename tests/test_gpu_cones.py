"""LP cone, second-order cone, equality constraints and the regularised LDL^T on the device, through
the C ABI, against the CPU oracle on identical seeded programs — the device twins of
tests/test_oracle_cones.py (reference: conex/test/test_socp.cc, equality_constraints_test.cc,
kkt_solver_options_test.cc, test_sdp.cc:13-59).

Tolerances (BASELINE.json): objectives 1e-7 relative, iteration counts +-1, Newton-system entries
1e-10 relative.
"""
import ctypes as C

import numpy as np
import pytest

from harness import dptr, oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    import devlib
    return oracle(), devlib.product()


def compare_solves(build, b, cfg_kw, libs, ytol=1e-6):
    """Builds the same program in both libraries, solves, compares trajectory end points.

    Objectives must agree to 1e-7 relative (BASELINE.json) — or to the oracle's own rounding
    sensitivity where that is larger: the dual objective cx = (2<w,c> + AQc.y - k<c,Qc>)/k is a
    difference of O(k) terms, and for second-order cones the oracle's cx moves by up to 1e-6
    relative when its input is perturbed by one ulp (b * (1 + 2^-52)); that self-difference (x10)
    is then the yardstick, exactly like the trajectory horizon used for the PSD path."""
    out = []
    for L in libs:
        P = L.program()
        build(P)
        solved, y = P.maximize(b, L.default_config(**cfg_kw))
        out.append((solved, y, P.iteration_log(), P))
    (so, yo, lo, Po), (sd, yd, ld, Pd) = out
    perturbed = []
    for j in (-3, -2, -1, 1, 2, 3):
        Pp = libs[0].program()
        build(Pp)
        Pp.maximize(np.asarray(b, dtype=np.float64) * (1 + j * 2.0 ** -52), libs[0].default_config(**cfg_kw))
        lp = Pp.iteration_log()
        if len(lp) == len(lo):
            perturbed.append(lp[-1])
    assert so == sd
    assert abs(len(lo) - len(ld)) <= 1, (len(lo), len(ld))
    for key in ("by", "cx"):
        self_diff = max([abs(lo[-1][key] - q[key]) for q in perturbed] + [0.0])
        tol = max(1e-7 * max(1.0, abs(lo[-1][key])), 10 * self_diff)
        assert abs(lo[-1][key] - ld[-1][key]) <= tol, (key, lo[-1][key], ld[-1][key], self_diff)
    assert np.abs(yo - yd).max() <= ytol * max(1.0, np.abs(yo).max())
    return out


def arrow_lmi(Wsqrt):
    n = Wsqrt.shape[0]
    mats = []
    for i in range(n):
        A = np.zeros((n + 1, n + 1))
        A[1:, 0] = Wsqrt[:, i]
        A[0, 1:] = Wsqrt[:, i]
        mats.append(A)
    return mats, np.eye(n + 1)


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("incremental", [False, True])
def test_soc_matches_oracle_and_arrow_lmi(libs, seed, incremental):
    # test_socp.cc:17-100
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    n = 3
    Wsqrt = rng.uniform(-1, 1, size=(n, n))
    As = np.zeros((n + 1, n))
    As[1:, :] = Wsqrt
    cs = np.zeros(n + 1)
    cs[0] = 1
    mats, Cm = arrow_lmi(Wsqrt)
    for i in range(-2, 2):
        b = np.full(n, float(i)) + rng.uniform(-1, 1, size=n) * .02
        res = compare_solves(lambda P: P.add_soc(As, cs, incremental=incremental), b,
                             dict(inv_sqrt_mu_max=10000), libs)
        y_soc = res[1][1]
        Pd = libs[1].program()
        Pd.add_dense_lmi(mats, Cm)
        s2, y_lmi = Pd.maximize(b, libs[1].default_config(inv_sqrt_mu_max=10000))
        assert s2 == 1
        assert np.linalg.norm(y_soc - y_lmi) < 1e-4


@pytest.mark.parametrize("i", range(3))
def test_lp_with_dual_recovery(libs, i):
    # kkt_solver_options_test.cc:26-70
    rng = np.random.Generator(np.random.PCG64(70 + i))
    nv, nc = 5, 6 + 2 * i
    A = rng.uniform(-1, 1, size=(nc, nv))
    Cv = np.abs(rng.uniform(-1, 1, size=nc))
    x0 = np.abs(rng.uniform(-1, 1, size=nc))
    x0 *= 0.01 / np.linalg.norm(x0)
    b = A.T @ x0
    kw = dict(prepare_dual_variables=1, inv_sqrt_mu_max=5e5, divergence_upper_bound=1000,
              dinf_upper_bound=1.35, final_centering_tolerance=1)
    res = compare_solves(lambda P: P.add_linear(A, Cv), b, kw, libs)
    xo = res[0][3].dual_variable(0).ravel()
    xd = res[1][3].dual_variable(0).ravel()
    assert np.abs(xo - xd).max() <= 1e-6 * max(1.0, np.abs(xo).max())
    slack = Cv - A @ res[1][1]
    assert slack.min() >= -1e-12 and xd.min() >= -1e-12
    assert np.linalg.norm(A.T @ xd - b) <= 1e-9 * max(np.linalg.norm(b), 1e-3)


def test_mixed_sdp_lp_slack_is_ones(libs):
    # test_sdp.cc:13-59: 2x2 LMI + upper/lower bound on y_1; optimal slack = ones(2, 2) +- 1e-6
    A = [np.array([[-1.0, 0], [0, 0]]), np.array([[0, -1.0], [-1.0, 0]]), np.array([[0, 0], [0, -1.0]])]
    ub = np.zeros((1, 3)); ub[0, 1] = 1.0
    lb = np.zeros((1, 3)); lb[0, 1] = -1.0

    def build(P):
        assert P.L.lib.CONEX_SetNumberOfVariables(P.h, 3) == 0
        P.m = 3
        P.add_linear(ub, [1.0])
        P.add_linear(lb, [-1.0])
        P.add_dense_lmi(A, np.zeros((2, 2)))
    res = compare_solves(build, [-1.0, 0.0, -1.0], dict(max_iterations=30), libs)
    y = res[1][1]
    S = -(y[0] * A[0] + y[1] * A[1] + y[2] * A[2])
    assert np.linalg.norm(S - np.ones((2, 2))) < 1e-6


@pytest.mark.parametrize("seed", range(4))
def test_equality_basic(libs, seed):
    # equality_constraints_test.cc:13-56
    rng = np.random.Generator(np.random.PCG64(seed))
    nv, neq, nin = 3, 1, 4
    A = rng.uniform(-1, 1, size=(nin, nv))
    slack = np.ones(nin)
    dual = np.ones(nin)
    slack[:nin // 2] = 0
    dual[nin // 2:] = 0
    ystar = rng.uniform(-1, 1, size=nv)
    Cv = slack + A @ ystar
    eq = rng.uniform(-1, 1, size=(neq, nv))

    def build(P):
        if P.m == 0:
            assert P.L.lib.CONEX_SetNumberOfVariables(P.h, nv) == 0
            P.m = nv
        P.add_equality(eq, eq @ ystar)
        P.add_linear(A, Cv)
        assert P.kkt_size() == nv + neq
    res = compare_solves(build, A.T @ dual, {}, libs, ytol=1e-5)
    y = res[1][1]
    assert np.linalg.norm(eq @ y - eq @ ystar) < 1e-5
    assert np.linalg.norm(y - ystar) < 5e-5


@pytest.mark.parametrize("separate", [False, True])
@pytest.mark.parametrize("seed", range(2))
def test_equality_many(libs, separate, seed):
    # equality_constraints_test.cc:58-126
    rng = np.random.Generator(np.random.PCG64(40 + seed))
    nv = 10
    nin, neq = nv + 10, nv - 2
    A = rng.uniform(-1, 1, size=(nin, nv))
    mm = nin // 2
    slack = np.ones(nin)
    dual = np.ones(nin)
    slack[:mm] = 1e-7
    dual[mm:] = 1e-7
    ystar = rng.uniform(-1, 1, size=nv)
    Cv = slack + A @ ystar
    eq = np.zeros((neq, nv))
    Bi = np.array([[1.0, 2.0, 3.0]])
    for i in range(neq):
        eq[i, [0, i + 1, nv - 1]] = Bi[0]

    def build(P):
        P.add_linear(A, Cv)
        if separate:
            for i in range(neq):
                P.add_equality(Bi, [eq[i] @ ystar], variables=[0, i + 1, nv - 1])
        else:
            P.add_equality(eq, eq @ ystar)
    cost = A.T @ dual
    kw = dict(final_centering_steps=10, initial_centering_steps_coldstart=0, max_iterations=40,
              divergence_upper_bound=.5)
    # late iterates are ill conditioned (slack 1e-7): compare the solutions, not every digit of by/cx
    out = []
    for L in libs:
        P = L.program()
        build(P)
        solved, y = P.maximize(cost, L.default_config(**kw))
        out.append((solved, y, P.iteration_log()))
    (so, yo, lo), (sd, yd, ld) = out
    assert so == sd
    assert abs(len(lo) - len(ld)) <= 1
    assert np.abs(yo - yd).max() < 1e-5
    assert (Cv - A @ yd).min() > -1e-8
    assert np.linalg.norm(eq @ yd - eq @ ystar) < 5e-7
    assert cost @ yd + 1e-4 >= cost @ ystar


def test_linear_inequalities_entry_point(libs):
    # interfaces/conex.cc:190-215
    rng = np.random.Generator(np.random.PCG64(9))
    nv = 4
    A = rng.uniform(-1, 1, size=(6, nv))
    ystar = rng.uniform(-1, 1, size=nv)
    lb = A @ ystar - 1.0
    ub = A @ ystar + 1.0
    lb[0] = ub[0] = (A @ ystar)[0]
    lb[1] = -1e9
    ub[2] = 1e9
    b = rng.uniform(-1, 1, size=nv)

    def build(P):
        assert P.L.lib.CONEX_SetNumberOfVariables(P.h, nv) == 0
        P.m = nv
        assert P.add_linear_inequalities(A, lb, ub) == -1
        assert P.kkt_size() == nv + 1
    res = compare_solves(build, b, dict(inv_sqrt_mu_max=1e4), libs, ytol=1e-5)
    r = A @ res[1][1]
    assert abs(r[0] - ub[0]) < 1e-6


def test_newton_system_with_equalities_matches_oracle(libs):
    # H entries 1e-10 relative, including the multiplier block [H A'; A 0]
    rng = np.random.Generator(np.random.PCG64(3))
    nv = 6
    A = rng.uniform(-1, 1, size=(9, nv))
    eq = rng.uniform(-1, 1, size=(2, nv))
    sys = []
    for L in libs:
        P = L.program(nv)
        P.add_linear(A, np.ones(9))
        P.add_equality(eq, np.array([0.1, -0.2]))
        sys.append(P.newton_system(coldstart=True))
    (Ho, AWo, AQo, sco), (Hd, AWd, AQd, scd) = sys
    assert np.abs(Ho - Hd).max() <= 1e-10 * np.abs(Ho).max()
    assert np.abs(AWo - AWd).max() <= 1e-12 and np.abs(AQo - AQd).max() <= 1e-12
    assert np.abs(sco - scd).max() <= 1e-12 * max(1.0, np.abs(sco).max())


@pytest.mark.parametrize("N", [1, 3, 64, 129, 300])
def test_device_ldlt_kernels(N):
    """cxb_sym_permute_lower + cxb_ldlt_lower + cxb_ldlt_solve against numpy on a quasi-definite
    KKT-shaped matrix; multi-block sizes exercise the panel solve and the DMMA trailing update."""
    import devlib as dev
    import torch
    L = dev.product().lib
    vp = C.c_void_p
    L.cxb_ldlt_worksize.restype = C.c_size_t
    L.cxb_ldlt_worksize.argtypes = [C.c_int]
    L.cxb_sym_permute_lower.argtypes = [vp, C.c_int, vp, C.c_long, vp, vp, C.c_long]
    L.cxb_ldlt_lower.argtypes = [vp, C.c_int, vp, C.c_long, vp, vp, vp]
    L.cxb_ldlt_solve.argtypes = [vp, C.c_int, vp, C.c_long, vp, vp, vp, vp]
    L.CONEXB200_RldltPivotOrder.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    rng = np.random.Generator(np.random.PCG64(N))
    k = N // 3
    h = N - k
    R = rng.uniform(-1, 1, size=(h, h))
    M = np.zeros((N, N))
    M[:h, :h] = R @ R.T + np.eye(h)
    if k:
        Aeq = rng.uniform(-1, 1, size=(k, h))
        M[h:, :h] = Aeq
        M[:h, h:] = Aeq.T
    diag = np.ascontiguousarray(np.diag(M).copy())
    perm = (C.c_int * N)()
    L.CONEXB200_RldltPivotOrder(N, dptr(diag), perm)
    perm = np.array(list(perm), dtype=np.int32)
    dK = dev.to_dev(np.tril(M))
    dperm = torch.from_numpy(perm).cuda()
    dKp = dev.dzeros(N * N)
    assert L.cxb_sym_permute_lower(None, N, dev.ptr(dK), N, dev.ptr(dperm), dev.ptr(dKp), N) == 0
    Kp = dev.from_dev(dKp, N, N)
    assert np.array_equal(np.tril(Kp), np.tril(M[np.ix_(perm, perm)]))
    signs = dev.dzeros(N)
    work = dev.dzeros(L.cxb_ldlt_worksize(N))
    info = dev.izeros(2)
    assert L.cxb_ldlt_lower(None, N, dev.ptr(dKp), N, dev.ptr(signs), dev.ptr(work), dev.ptr(info)) == 0
    Lf = np.tril(dev.from_dev(dKp, N, N))
    S = np.diag(dev.from_dev(signs))
    assert info.cpu().tolist() == [0, 0]
    PMP = M[np.ix_(perm, perm)]
    assert np.abs(Lf @ S @ Lf.T - PMP).max() < 1e-9 * max(1.0, np.abs(M).max())
    assert (np.diag(S)[:h] == 1).all() and (np.diag(S)[h:] == -1).all()
    x = rng.uniform(-1, 1, size=N)
    rhs = dev.to_dev(M @ x)
    tmp = dev.dzeros(N)
    assert L.cxb_ldlt_solve(None, N, dev.ptr(dKp), N, dev.ptr(signs), dev.ptr(dperm), dev.ptr(rhs), dev.ptr(tmp)) == 0
    assert np.abs(dev.from_dev(rhs) - x).max() < 1e-7


def test_device_ldlt_regularises_zero_pivots():
    # RLDLT.h:378-389
    import devlib as dev
    L = dev.product().lib
    vp = C.c_void_p
    L.cxb_ldlt_lower.argtypes = [vp, C.c_int, vp, C.c_long, vp, vp, vp]
    L.cxb_ldlt_worksize.restype = C.c_size_t
    M = np.diag([2.0, 0.0, -1e-12])
    dK = dev.to_dev(M)
    signs, work, info = dev.dzeros(3), dev.dzeros(L.cxb_ldlt_worksize(3)), dev.izeros(2)
    assert L.cxb_ldlt_lower(None, 3, dev.ptr(dK), 3, dev.ptr(signs), dev.ptr(work), dev.ptr(info)) == 0
    Lf = dev.from_dev(dK, 3, 3)
    d = np.diag(Lf) ** 2 * dev.from_dev(signs)
    assert np.allclose(d, [2.0, 1e-9, -1e-9], rtol=1e-12, atol=0)
    assert info.cpu().tolist() == [0, 1]
