"""The incremental real LMI (reference HermitianPsdConstraint<Real>, conex/hermitian_psd.cc) on the
device: CONEX_NewLinearMatrixInequality + CONEX_UpdateLinearOperator / CONEX_UpdateAffineTerm.
Pins: interfaces/python/test/run_tests.py:299-321 (y = (-1, -1)), conex/test/hermitian_psd_test.cc:24-63
(agreement with the dense-LMI path to 1e-11 under tight centering), and parity with the oracle when
both draw their Lanczos start vectors from the same libc rand() stream.
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


def srand(seed):
    C.CDLL(None).srand(seed)


def random_instance(rank, dim, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = []
    for _ in range(dim):
        R = rng.uniform(-1, 1, size=(rank, rank))
        mats.append(R + R.T)
    return mats, np.eye(rank)


def test_known_answer(libs):
    _, D = libs
    A0 = np.zeros((3, 3)); A0[1, 0] = A0[0, 1] = -1.0
    A1 = np.zeros((3, 3)); A1[2, 1] = A1[1, 2] = -1.0
    P = D.program(2)
    P.add_hermitian_lmi([A0, A1], np.diag([1.0, 2.0, 1.0]))
    solved, y = P.maximize([-1.0, -1.0])
    assert solved == 1 and np.linalg.norm(y + 1.0) < 1e-6


@pytest.mark.parametrize("rank,dim", [(3, 2), (8, 5), (20, 4), (70, 6)])
def test_agrees_with_dense_lmi_under_tight_centering(libs, rank, dim):
    _, D = libs
    mats, Cm = random_instance(rank, dim, rank * 10 + dim)
    cfg = D.default_config(inv_sqrt_mu_max=float(np.sqrt(1.0 / 1e-4)), final_centering_tolerance=1e-8,
                           prepare_dual_variables=1)
    P1 = D.program(dim)
    P1.add_hermitian_lmi(mats, Cm)
    b = P1.feasible_objective()
    s1, y1 = P1.maximize(b, cfg)
    X1 = P1.dual_variable(0)
    P2 = D.program()
    P2.add_dense_lmi(mats, Cm)
    s2, y2 = P2.maximize(b, cfg)
    X2 = P2.dual_variable(0)
    assert s1 == 1 and s2 == 1
    assert np.linalg.norm(y1 - y2) < 1e-11 * max(1.0, np.linalg.norm(y2))
    assert np.linalg.norm(X1 - X2) < 1e-11 * max(1.0, np.linalg.norm(X2))


@pytest.mark.parametrize("rank,dim", [(3, 2), (12, 6), (70, 9)])
def test_matches_oracle_with_shared_rand_stream(libs, rank, dim):
    """Both libraries draw Random(n, 1) from libc rand(); seeded alike they follow the same trajectory."""
    O, D = libs
    mats, Cm = random_instance(rank, dim, 500 + rank + dim)
    logs, ys = [], []
    for L in (O, D):
        P = L.program(dim)
        P.add_hermitian_lmi(mats, Cm)
        b = P.feasible_objective()
        srand(11)
        solved, y = P.maximize(b, L.default_config(prepare_dual_variables=1))
        assert solved == 1
        logs.append(P.iteration_log())
        ys.append(y)
    lo, ld = logs
    # Same trajectory step by step — until one of the un-reorthogonalised Lanczos estimates jumps on one side (a
    # rounding-level difference can make one d_inf estimate an outlier: the oracle itself shows d_inf = 1.19 at step 9
    # of the 70 x 9 instance where the device estimates 0.61, and with the polled triangular solves the device meets
    # such an outlier at step 6 and needs 20 instead of 16 iterations, profiles/r02_e_hermitian_n70_lanczos_outlier.txt).
    # From there on only the solution is compared.
    diverged = None
    for i, (a, b_) in enumerate(zip(lo, ld)):
        if abs(a["inv_sqrt_mu"] - b_["inv_sqrt_mu"]) > 1e-6 * a["inv_sqrt_mu"] or abs(a["d_inf"] - b_["d_inf"]) > 1e-3:
            diverged = i
            break
    if diverged is None:
        assert abs(len(lo) - len(ld)) <= 1
    else:
        assert diverged >= 3 and len(ld) <= 25
    assert abs(lo[-1]["by"] - ld[-1]["by"]) <= 1e-7 * max(1.0, abs(lo[-1]["by"]))
    if diverged is None:
        assert abs(lo[-1]["cx"] - ld[-1]["cx"]) <= 1e-7 * max(1.0, abs(lo[-1]["cx"]))
    for a, b_ in zip(lo[:3], ld[:3]):  # early iterates: same mu, same step norms
        assert abs(a["inv_sqrt_mu"] - b_["inv_sqrt_mu"]) <= 1e-8 * a["inv_sqrt_mu"]
        assert abs(a["d_inf"] - b_["d_inf"]) <= 1e-7 * max(1.0, a["d_inf"])
    assert np.abs(ys[0] - ys[1]).max() <= 1e-6 * max(1.0, np.abs(ys[0]).max())


def test_update_after_solve_is_picked_up(libs):
    O, D = libs
    mats, Cm = random_instance(6, 3, 42)
    out = []
    for L in (O, D):
        P = L.program(3)
        cid = P.add_hermitian_lmi(mats, Cm)
        b = P.feasible_objective()
        srand(3)
        P.maximize(b)
        assert L.lib.CONEX_UpdateAffineTerm(P.h, cid, 2.5, 1, 1, 0) == 0
        assert L.lib.CONEX_UpdateLinearOperator(P.h, cid, 0.75, 2, 4, 1, 0) == 0
        srand(3)
        solved, y = P.maximize(b)
        out.append((solved, y))
    assert out[0][0] == out[1][0] == 1
    assert np.abs(out[0][1] - out[1][1]).max() < 1e-6


def test_argument_validation(libs):
    _, D = libs
    P = D.program(2)
    cid = C.c_int(-1)
    L = D.lib
    assert L.CONEX_NewLinearMatrixInequality(P.h, 0, 1, C.byref(cid)) == 1      # order >= 1
    assert L.CONEX_NewLinearMatrixInequality(P.h, 3, 3, C.byref(cid)) == 1      # dim in {1,2,4,8}
    assert L.CONEX_NewLinearMatrixInequality(P.h, 3, 2, C.byref(cid)) == 1      # complex: not on the device path
    assert L.CONEX_NewLinearMatrixInequality(P.h, 3, 1, C.byref(cid)) == 0 and cid.value == 0
    assert L.CONEX_UpdateLinearOperator(P.h, 0, 1.0, 0, 3, 0, 0) == 1           # row out of bounds
    assert L.CONEX_UpdateLinearOperator(P.h, 0, 1.0, 0, 1, 0, 1) == 1           # imaginary part of a real LMI
    assert L.CONEX_UpdateLinearOperator(P.h, 1, 1.0, 0, 1, 0, 0) == 1           # invalid constraint
    assert L.CONEX_UpdateAffineTerm(P.h, 0, 1.0, 1, 5, 0) == 1


@pytest.mark.parametrize("n", [1, 5, 64, 130])
def test_taylor_expm_and_hermitian_lanczos_kernels(libs, n):
    O, D = libs
    import devlib as dev
    vp = C.c_void_p
    L = D.lib
    L.cxb_taylor_expm.argtypes = [vp, C.c_int, vp, vp, vp]
    L.cxb_lanczos_two_sided_ex.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_double]
    O.lib.ORACLE_HermitianLanczos.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                              C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
    rng = np.random.Generator(np.random.PCG64(n))
    X = rng.uniform(-1, 1, size=(n, n)) * 0.2
    dX, dout, dwork = dev.to_dev(X), dev.dzeros(n * n), dev.dzeros(n * n)
    assert L.cxb_taylor_expm(None, n, dev.ptr(dX), dev.ptr(dout), dev.ptr(dwork)) == 0
    Y = np.eye(n) + X / 4 + X @ X / 32
    ref = np.linalg.matrix_power(Y, 4)
    assert np.abs(dev.from_dev(dout, n, n) - ref).max() < 1e-12 * np.abs(ref).max()
    if n == 1:
        return
    R = rng.uniform(-1, 1, size=(n, n))
    W = R @ R.T / n + np.eye(n)
    S = rng.uniform(-1, 1, size=(n, n)); S = (S + S.T) / np.sqrt(n)
    WS = W @ S
    r = rng.uniform(-1, 1, size=n)
    num_iter = n // 2 + 1
    ritz = np.zeros(num_iter + 1)
    k = O.lib.ORACLE_HermitianLanczos(n, dptr(np.asfortranarray(WS)), dptr(np.asfortranarray(W)), dptr(r),
                                      num_iter, dptr(ritz))
    alpha, beta, count = dev.dzeros(num_iter + 1), dev.dzeros(num_iter + 1), dev.izeros(2)
    work = dev.dzeros(L.cxb_lanczos_worksize(n))
    assert L.cxb_lanczos_two_sided_ex(None, n, dev.ptr(dev.to_dev(WS)), dev.ptr(dev.to_dev(W)), dev.ptr(dev.to_dev(r)),
                                      None, num_iter, dev.ptr(alpha), dev.ptr(beta), dev.ptr(count), dev.ptr(work),
                                      1e-5) == 0
    cnt = int(count.cpu()[0])
    a = dev.from_dev(alpha)[:cnt + 1]
    b = dev.from_dev(beta)[:cnt]
    T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
    ev = np.linalg.eigvalsh(T)
    assert cnt + 1 == k
    assert abs(ev[0] - ritz[:k].min()) < 1e-8 * max(1.0, abs(ev[0]))
    assert abs(ev[-1] - ritz[:k].max()) < 1e-8 * max(1.0, abs(ev[-1]))


def maxcut_instance(n, seed):
    from harness import maxcut_lmi
    return maxcut_lmi(n, seed)


@pytest.mark.parametrize("kind,size", [("maxcut", 24), ("maxcut", 70), ("lovasz", 20)])
def test_entry_sparse_operators_match_oracle(libs, kind, size):
    """MaxCut / Lovasz-theta operators given entry by entry run the entry-sparse device kernels
    (no dense A_i anywhere); the oracle runs the reference's dense HermitianPsdConstraint arithmetic."""
    from harness import lovasz_theta_lmi, maxcut_lmi
    O, D = libs
    mats, Cm, b = maxcut_lmi(size, 2) if kind == "maxcut" else lovasz_theta_lmi(size, 3 * size, 4)
    out = []
    for L in (O, D):
        P = L.program(len(mats))
        cid = P.add_hermitian_lmi(mats, Cm)
        srand(5)
        solved, y = P.maximize(b, L.default_config(prepare_dual_variables=1))
        out.append((solved, y, P.iteration_log(), P.dual_variable(cid)))
        if L.kind == "b200":
            assert L.lib.CONEXB200_ConstraintIsEntrySparse(P.h, cid) == 1
    (so, yo, lo, Xo), (sd, yd, ld, Xd) = out
    assert so == sd == 1
    assert abs(len(lo) - len(ld)) <= 1
    assert abs(lo[-1]["by"] - ld[-1]["by"]) <= 1e-7 * max(1.0, abs(lo[-1]["by"]))
    assert np.abs(yo - yd).max() <= 1e-6 * max(1.0, np.abs(yo).max())
    assert np.abs(Xo - Xd).max() <= 1e-5 * max(1.0, np.abs(Xo).max())
    if kind == "maxcut":
        assert np.abs(np.diag(Xd) - 1.0).max() < 1e-6  # the MaxCut relaxation has unit diagonal


def test_dense_operator_through_incremental_api_stays_dense(libs):
    _, D = libs
    mats, Cm = random_instance(10, 4, 3)
    P = D.program(4)
    cid = P.add_hermitian_lmi(mats, Cm)
    P.maximize(P.feasible_objective())
    assert D.lib.CONEXB200_ConstraintIsEntrySparse(P.h, cid) == 0


@pytest.mark.parametrize("n,m,per", [(7, 5, 2), (30, 40, 3), (64, 10, 12)])
def test_sparse_schur_and_slack_kernels(libs, n, m, per):
    """cxb_sparse_lmi_schur / cxb_sparse_lmi_slack against numpy with dense copies of the matrices."""
    import devlib as dev
    import torch
    _, D = libs
    L = D.lib
    vp = C.c_void_p
    L.cxb_sparse_lmi_schur.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_long]
    L.cxb_sparse_lmi_slack.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_double, vp]
    rng = np.random.Generator(np.random.PCG64(n * m))
    mats = []
    offsets, rows, cols, vals = [0], [], [], []
    by_pos = {}
    for i in range(m):
        A = np.zeros((n, n))
        for _ in range(per):
            r, c = sorted(rng.integers(0, n, size=2), reverse=True)
            A[r, c] = A[c, r] = rng.uniform(-1, 1)
        mats.append(A)
        for c in range(n):
            for r in range(n):
                if A[r, c] != 0:
                    rows.append(r); cols.append(c); vals.append(A[r, c])
                    by_pos.setdefault(c * n + r, []).append((i, A[r, c]))
        offsets.append(len(rows))
    R = rng.uniform(-1, 1, size=(n, n))
    W = R @ R.T / n + np.eye(n)
    Cm = rng.uniform(-1, 1, size=(n, n)); Cm = Cm + Cm.T
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    dvals = torch.tensor(vals, dtype=torch.float64, device="cuda")
    ldh = (m + 3) & ~1
    dH, work = dev.dzeros(ldh * (m + 2)), dev.dzeros(2 * n * n)
    doff, drows, dcols = i32(offsets), i32(rows), i32(cols)
    assert L.cxb_sparse_lmi_schur(None, n, m, dev.ptr(doff), dev.ptr(drows), dev.ptr(dcols), dev.ptr(dvals),
                                  dev.ptr(dev.to_dev(Cm)), dev.ptr(dev.to_dev(W)), dev.ptr(work), dev.ptr(dH), ldh) == 0
    Haug = dev.from_dev(dH, ldh, m + 2)
    WCW = W @ Cm @ W
    for i in range(m):
        for j in range(i + 1):
            ref = np.trace(mats[i] @ W @ mats[j] @ W)
            assert abs(Haug[i, j] - ref) <= 1e-12 * max(1.0, abs(ref))
        assert abs(Haug[m, i] - np.sum(mats[i] * WCW)) <= 1e-11 * max(1.0, np.abs(WCW).max())
        assert abs(Haug[m + 1, i] - np.sum(mats[i] * W)) <= 1e-12 * max(1.0, np.abs(W).max())
    assert abs(Haug[m, m] - np.sum(Cm * WCW)) <= 1e-11 * abs(np.sum(Cm * WCW))
    assert abs(Haug[m + 1, m] - np.sum(Cm * W)) <= 1e-11 * max(1.0, abs(np.sum(Cm * W)))
    # slack
    keys = sorted(by_pos)
    pos_ptr, pos_var, pos_val = [0], [], []
    for k in keys:
        for v, a in by_pos[k]:
            pos_var.append(v); pos_val.append(a)
        pos_ptr.append(len(pos_var))
    y = rng.uniform(-1, 1, size=m)
    out = dev.dzeros(n * n)
    dpos = torch.tensor(keys, dtype=torch.int64, device="cuda")
    dpv = torch.tensor(pos_val, dtype=torch.float64, device="cuda")
    dptr_, dvar = i32(pos_ptr), i32(pos_var)
    assert L.cxb_sparse_lmi_slack(None, n, len(keys), dev.ptr(dptr_), dev.ptr(dpos), dev.ptr(dvar),
                                  dev.ptr(dpv), dev.ptr(dev.to_dev(Cm)), dev.ptr(dev.to_dev(y)), 0.7, dev.ptr(out)) == 0
    ref = sum(y[i] * mats[i] for i in range(m)) - 0.7 * Cm
    assert np.abs(dev.from_dev(out, n, n) - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
