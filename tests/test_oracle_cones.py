"""Pins the oracle's LP / second-order / equality cones and its RLDLT restatement on the
reference's own known-answer and cross-formulation tests (CPU only):

* conex/test/test_socp.cc:17-100        — SOC constraint == arrow-matrix LMI, |y1 - y2| <= 1e-4
* conex/test/equality_constraints_test.cc — Basic / ManyConstraints / ManySeparateConstraints
* conex/test/assembly_test.cc:162-168   — LDL^T reproduces the matrix (vs Eigen::LDLT, 1e-9)
* conex/test/kkt_solver_options_test.cc:26-70 — LP primal/dual residuals with dual recovery
"""
import ctypes as C

import numpy as np
import pytest

from harness import random_dense_lmi, dptr, oracle

c_int_p = C.POINTER(C.c_int)


def arrow_lmi(Wsqrt):
    n = Wsqrt.shape[0]
    mats = []
    for i in range(n):
        A = np.zeros((n + 1, n + 1))
        A[1:, 0] = Wsqrt[:, i]
        A[0, 1:] = Wsqrt[:, i]
        mats.append(A)
    return mats, np.eye(n + 1)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("incremental", [False, True])
def test_soc_equals_arrow_lmi(seed, incremental):
    # test_socp.cc:17-100
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    n = 3
    Wsqrt = rng.uniform(-1, 1, size=(n, n))
    As = np.zeros((n + 1, n))
    As[1:, :] = Wsqrt
    cs = np.zeros(n + 1)
    cs[0] = 1
    mats, Cm = arrow_lmi(Wsqrt)
    cfg = O.default_config(inv_sqrt_mu_max=10000)
    for i in range(-2, 2):
        b = np.full(n, float(i)) + rng.uniform(-1, 1, size=n) * .02
        P1 = O.program()
        P1.add_soc(As, cs, incremental=incremental)
        s1, y1 = P1.maximize(b, cfg)
        P2 = O.program()
        P2.add_dense_lmi(mats, Cm)
        s2, y2 = P2.maximize(b, cfg)
        assert s1 == s2 == 1
        assert np.linalg.norm(y1 - y2) < 1e-4
        # the optimum of max b'y s.t. |Wsqrt y| <= 1 is y = Wsqrt^{-1} u, u = Wsqrt^{-T} b / |.|
        u = np.linalg.solve(Wsqrt.T, b)
        ystar = np.linalg.solve(Wsqrt, u / np.linalg.norm(u))
        assert np.linalg.norm(y1 - ystar) < 2e-3 * max(1.0, np.linalg.norm(ystar))


@pytest.mark.parametrize("seed", range(8))
def test_equality_basic(seed):
    # equality_constraints_test.cc:13-56. The reference checks one Eigen::Random instance at 1e-5;
    # that instance cannot be regenerated here, and over PCG64 instances the distance to the optimum
    # at the default mu = 1e-6 ranges over 1e-6 .. 3e-5, hence the 5e-5 on |y - y*|.
    O = oracle()
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
    P = O.program(nv)
    P.add_equality(eq, eq @ ystar)
    P.add_linear(A, Cv)
    assert P.kkt_size() == nv + neq
    solved, y = P.maximize(A.T @ dual)
    assert solved == 1
    assert np.linalg.norm(eq @ y - eq @ ystar) < 1e-5
    assert np.linalg.norm(y - ystar) < 5e-5


@pytest.mark.parametrize("separate", [False, True])
@pytest.mark.parametrize("seed", range(5))
def test_equality_many(separate, seed):
    # equality_constraints_test.cc:58-126
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(40 + seed))
    nv = 10
    nin, neq = nv + 10, nv - 2
    A = rng.uniform(-1, 1, size=(nin, nv))
    m = nin // 2
    slack = np.ones(nin)
    dual = np.ones(nin)
    slack[:m] = 1e-7
    dual[m:] = 1e-7
    ystar = rng.uniform(-1, 1, size=nv)
    Cv = slack + A @ ystar
    P = O.program(nv)
    P.add_linear(A, Cv)
    eq = np.zeros((neq, nv))
    Bi = np.array([[1.0, 2.0, 3.0]])
    for i in range(neq):
        vars_ = [0, i + 1, nv - 1]
        eq[i, vars_] = Bi[0]
        if separate:
            P.add_equality(Bi, [eq[i] @ ystar], variables=vars_)
    if not separate:
        P.add_equality(eq, eq @ ystar)
    cost = A.T @ dual
    cfg = O.default_config(final_centering_steps=10, initial_centering_steps_coldstart=0,
                           max_iterations=40, divergence_upper_bound=.5)
    solved, y = P.maximize(cost, cfg)
    assert (Cv - A @ y).min() > -1e-8
    assert np.linalg.norm(eq @ y - eq @ ystar) < 5e-7
    assert cost @ y + 1e-4 >= cost @ ystar


def test_linear_inequalities_preprocessing():
    # interfaces/conex.cc:190-215: lb == ub rows -> equalities, the rest -> scaled LP rows
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(9))
    nv = 4
    A = rng.uniform(-1, 1, size=(6, nv))
    ystar = rng.uniform(-1, 1, size=nv)
    lb = A @ ystar - 1.0
    ub = A @ ystar + 1.0
    lb[0] = ub[0] = (A @ ystar)[0]   # equality
    lb[1] = -1e9                      # one-sided
    ub[2] = 1e9
    P = O.program(nv)
    assert P.add_linear_inequalities(A, lb, ub) == -1
    assert P.kkt_size() == nv + 1
    b = rng.uniform(-1, 1, size=nv)
    solved, y = P.maximize(b, O.default_config(inv_sqrt_mu_max=1e4))
    assert solved == 1
    r = A @ y
    assert abs(r[0] - ub[0]) < 1e-6
    assert (r <= ub + 1e-6).all() and (r >= lb - 1e-6).all()


@pytest.mark.parametrize("n", [1, 2, 7, 30])
def test_rldlt_reconstructs(n):
    # assembly_test.cc:162-168 (1e-9) — indefinite KKT-like matrix [H A'; A 0]
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(n))
    k = max(n // 3, 0) if n > 1 else 0
    h = n - k
    R = rng.uniform(-1, 1, size=(h, h))
    M = np.zeros((n, n))
    M[:h, :h] = R @ R.T + np.eye(h)
    if k:
        Aeq = rng.uniform(-1, 1, size=(k, h))
        M[h:, :h] = Aeq
        M[:h, h:] = Aeq.T
    LD = np.asfortranarray(np.tril(M))
    tr = (C.c_int * n)()
    ok = O.lib.ORACLE_LdltLower(n, dptr(LD), tr)
    assert ok == 1
    tr = list(tr)
    L = np.tril(LD, -1) + np.eye(n)
    D = np.diag(np.diag(LD))
    perm = list(range(n))
    for i, t in enumerate(tr):
        perm[i], perm[t] = perm[t], perm[i]
    PMP = M[np.ix_(perm, perm)]
    assert np.abs(L @ D @ L.T - PMP).max() < 1e-9 * max(1.0, np.abs(M).max())
    # pivots are taken in decreasing |original diagonal| order (RLDLT.h:330-334)
    d0 = np.abs(np.diag(M))[perm]
    assert (np.diff(d0) <= 1e-15).all()
    x = rng.uniform(-1, 1, size=n)
    rhs = M @ x
    O.lib.ORACLE_SolveLdlt(n, dptr(LD), (C.c_int * n)(*tr), dptr(rhs))
    assert np.abs(rhs - x).max() < 1e-8


def test_rldlt_regularises_zero_pivot():
    # RLDLT.h:378-389: |pivot| <= 1e-9 is replaced by +-1e-9 and reported
    O = oracle()
    M = np.asfortranarray(np.array([[2.0, 0, 0], [0, 0.0, 0], [0, 0, -1e-12]]))
    tr = (C.c_int * 3)()
    ok = O.lib.ORACLE_LdltLower(3, dptr(M), tr)
    assert ok == 0
    assert sorted(np.diag(M).tolist()) == [-1e-9, 1e-9, 2.0]


@pytest.mark.parametrize("mode,refine", [(0, 0), (0, 3), (1, 0)])
@pytest.mark.parametrize("i", range(3))
def test_kkt_solver_options(i, mode, refine):
    """kkt_solver_options_test.cc:70-88 (`UseLLT`, `UseIterativeRefinement` with 3 correction solves —
    kkt_solver.cc:174-178,248-261 —, `LP UseLDLT`): the LP properties of :26-66 hold under every option.
    (LLT and LDLT modes take the same branch in the reference: Cholesky unless multipliers exist,
    kkt_solver.cc:180-193. The QR mode is outside the hot path.)"""
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(170 + i))
    nv, nc = 5, 6 + 2 * i
    A = rng.uniform(-1, 1, size=(nc, nv))
    Cv = np.abs(rng.uniform(-1, 1, size=nc))
    x0 = np.abs(rng.uniform(-1, 1, size=nc))
    x0 *= 0.01 / np.linalg.norm(x0)
    b = A.T @ x0
    cfg = O.default_config(prepare_dual_variables=1, inv_sqrt_mu_max=5e5, divergence_upper_bound=1000,
                           dinf_upper_bound=1.35, final_centering_tolerance=1, kkt_solver=mode,
                           iterative_refinement_iterations=refine)
    P = O.program()
    P.add_linear(A, Cv)
    solved, y = P.maximize(b, cfg)
    x = P.dual_variable(0).ravel()
    slack = Cv - A @ y
    assert solved == 1
    assert np.linalg.norm(A.T @ x - b) <= 1e-11 * np.linalg.norm(b)
    assert slack.min() >= -1e-12 and x.min() >= -1e-12 and slack @ x >= -1e-12
    assert slack @ x <= (1.0 / 5e5 ** 2 + 1e-6) * nc


def test_iterative_refinement_changes_nothing_on_a_well_conditioned_lmi():
    O = oracle()
    mats, Cm = random_dense_lmi(12, 7, 9)
    ys = []
    for refine in (0, 2):
        P = O.program()
        P.add_dense_lmi(mats, Cm)
        solved, y = P.maximize(P.feasible_objective(), O.default_config(iterative_refinement_iterations=refine))
        assert solved == 1
        ys.append(y)
    assert np.abs(ys[0] - ys[1]).max() < 1e-9


@pytest.mark.parametrize("i", range(4))
def test_lp_dual_recovery(i):
    # kkt_solver_options_test.cc:26-70 (LLT mode)
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(70 + i))
    nv, nc = 5, 6 + 2 * i
    A = rng.uniform(-1, 1, size=(nc, nv))
    Cv = np.abs(rng.uniform(-1, 1, size=nc))
    x0 = np.abs(rng.uniform(-1, 1, size=nc))
    x0 *= 0.01 / np.linalg.norm(x0)
    b = A.T @ x0
    cfg = O.default_config(prepare_dual_variables=1, inv_sqrt_mu_max=5e5, divergence_upper_bound=1000,
                           dinf_upper_bound=1.35, final_centering_tolerance=1)
    P = O.program()
    P.add_linear(A, Cv)
    solved, y = P.maximize(b, cfg)
    x = P.dual_variable(0).ravel()
    eps = 1e-12
    slack = Cv - A @ y
    assert np.linalg.norm(A.T @ x - b) <= 1e-9 * max(np.linalg.norm(b), 1e-3)
    assert slack.min() >= -eps and x.min() >= -eps
    mu = 1.0 / cfg.inv_sqrt_mu_max ** 2
    assert slack @ x <= (mu + np.sqrt(eps)) * nc


@pytest.mark.parametrize("n", [1, 5, 40, 300])
def test_product_pivot_order_matches_rldlt(n):
    """Host logic of the product (no GPU needed): the pivot order it derives from the diagonal alone
    is the one the RLDLT restatement produces by actually factoring (RLDLT.h:328-356)."""
    import devlib
    L = devlib.product().lib
    L.CONEXB200_RldltPivotOrder.argtypes = [C.c_int, C.POINTER(C.c_double), c_int_p]
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(n))
    k = n // 3
    M = np.zeros((n, n))
    R = rng.uniform(-1, 1, size=(n - k, n - k))
    M[:n - k, :n - k] = R @ R.T + np.eye(n - k)
    if k:
        M[n - k:, :n - k] = rng.uniform(-1, 1, size=(k, n - k))
        # a few exact ties and negative entries on the diagonal
        M[0, 0] = M[1, 1]
        M[n - 1, n - 1] = -M[2, 2]
    diag = np.ascontiguousarray(np.diag(M).copy())
    perm = (C.c_int * n)()
    L.CONEXB200_RldltPivotOrder(n, dptr(diag), perm)
    LD = np.asfortranarray(np.tril(M))
    tr = (C.c_int * n)()
    O.lib.ORACLE_LdltLower(n, dptr(LD), tr)
    ref = list(range(n))
    for i, t in enumerate(tr):
        ref[i], ref[t] = ref[t], ref[i]
    assert list(perm) == ref


def libc_srand(seed):
    C.CDLL(None).srand(seed)


def test_hermitian_lmi_known_answer():
    # interfaces/python/test/run_tests.py:299-321 through the incremental API: y = (-1, -1) +- 1e-6
    O = oracle()
    A0 = np.zeros((3, 3)); A0[1, 0] = A0[0, 1] = -1.0
    A1 = np.zeros((3, 3)); A1[2, 1] = A1[1, 2] = -1.0
    P = O.program(2)
    P.add_hermitian_lmi([A0, A1], np.diag([1.0, 2.0, 1.0]))
    libc_srand(1)
    solved, y = P.maximize([-1.0, -1.0])
    assert solved == 1 and np.linalg.norm(y + 1.0) < 1e-6


@pytest.mark.parametrize("rank,dim", [(3, 2), (8, 5), (20, 4)])
def test_hermitian_agrees_with_dense_lmi_under_tight_centering(rank, dim):
    # conex/test/hermitian_psd_test.cc:24-63: the two LMI code paths agree to 1e-11
    O = oracle()
    rng = np.random.Generator(np.random.PCG64(rank * 10 + dim))
    mats = []
    for _ in range(dim):
        R = rng.uniform(-1, 1, size=(rank, rank))
        mats.append(R + R.T)
    Cm = np.eye(rank)
    cfg = O.default_config(inv_sqrt_mu_max=float(np.sqrt(1.0 / 1e-4)), final_centering_tolerance=1e-8,
                           prepare_dual_variables=1)
    P1 = O.program(dim)
    P1.add_hermitian_lmi(mats, Cm)
    b = P1.feasible_objective()
    libc_srand(1)
    s1, y1 = P1.maximize(b, cfg)
    X1 = P1.dual_variable(0)
    P2 = O.program()
    P2.add_dense_lmi(mats, Cm)
    s2, y2 = P2.maximize(b, cfg)
    X2 = P2.dual_variable(0)
    assert s1 == 1 and s2 == 1
    assert np.linalg.norm(y1 - y2) < 1e-11
    assert np.linalg.norm(X1 - X2) < 1e-11


@pytest.mark.parametrize("n", [1, 4, 9])
def test_taylor_exponential_and_hermitian_lanczos(n):
    O = oracle()
    L = O.lib
    L.ORACLE_TaylorExpm.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ORACLE_HermitianLanczos.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
    rng = np.random.Generator(np.random.PCG64(n))
    X = np.asfortranarray(rng.uniform(-1, 1, size=(n, n)) * 0.2)
    out = np.zeros((n, n), order="F")
    L.ORACLE_TaylorExpm(n, dptr(X), dptr(out))
    Y = np.eye(n) + X / 4 + X @ X / 32
    assert np.abs(out - np.linalg.matrix_power(Y, 4)).max() < 1e-14
    # exact spectrum after n steps for a symmetric positive W-weighted problem (cf. the 4x4 golden)
    R = rng.uniform(-1, 1, size=(n, n))
    W = np.asfortranarray(R @ R.T + np.eye(n))
    S = rng.uniform(-1, 1, size=(n, n)); S = S + S.T
    WS = np.asfortranarray(W @ S)
    r = rng.uniform(-1, 1, size=n)
    ritz = np.zeros(n + 1)
    k = L.ORACLE_HermitianLanczos(n, dptr(WS), dptr(W), dptr(r), n, dptr(ritz))
    ev = np.sort(np.linalg.eigvals(WS).real)
    assert k == n and np.abs(np.sort(ritz[:k]) - ev).max() < 1e-8 * max(1.0, np.abs(ev).max())


def test_known_bad_instance_equality_constraint_failing_ldlt():
    """conex/test/solver_failures.cc:12-46 (`EqualityConstraintFailingLDLT`): x0 = x1, x0 + x1 <= 1,
    maximise x0 + x1. The KKT matrix [[1, 1, 1], [1, 1, -1], [1, -1, 0]] has a singular leading block —
    the reference lists it as a case its factorisation fails on; with the +-1e-9 pivot regularisation
    restated from RLDLT.h:378-389 the program solves to (1/2, 1/2)."""
    O = oracle()
    P = O.program(2)
    P.add_equality(np.array([[1.0, -1.0]]), np.array([0.0]), [0, 1])
    P.add_linear(np.array([[1.0, 1.0]]), np.array([1.0]))
    solved, y = P.maximize(np.array([1.0, 1.0]), O.default_config())
    assert solved == 1
    assert np.abs(y - 0.5).max() < 1e-5 and abs(y[0] - y[1]) < 1e-9


def test_taylor_exponential_reference_known_answer():
    """conex/test/exponential_map_test.cc:31-48: ExponentialMap of A * 0.001, A the 4 x 4 matrix of the
    Pade test, against the true exponential at 1e-7."""
    import scipy.linalg as sla
    L = oracle().lib
    L.ORACLE_TaylorExpm.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    A = np.array([[3, 1, 0, 1], [1, 3, 1, 0], [0, 1, 4, 1], [1, 0, 1, 5]], dtype=np.float64) * .001
    X = np.asfortranarray(A)
    out = np.zeros((4, 4), order="F")
    L.ORACLE_TaylorExpm(4, dptr(X), dptr(out))
    assert np.abs(out - sla.expm(A)).max() < 1e-7
