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
    assert len(lo) == len(ld)
    for key in ("by", "cx"):
        assert abs(lo[-1][key] - ld[-1][key]) <= 1e-7 * max(1.0, abs(lo[-1][key]))
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
