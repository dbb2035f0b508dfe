"""The reference's own Python unit tests (interfaces/python/test/run_tests.py:62-388), ported to pytest
and run through `conex_b200.Conex` — the mirror of the reference's `Conex` class — twice: bound to the
CPU oracle's library (no GPU: this pins the oracle AND executes every line of the wrapper class), and
bound to the product on the device. Only the seeds differ from the reference (np.random.randn without a
seed there). The complex / quaternion / octonion variants of tests 9-11 are outside the device path
(hyper_complex_dim > 1 fails loudly): they are checked to fail, not to solve.

The file sorts last on purpose: it exercises the Python wrapper, not new kernels."""
import ctypes as C

import numpy as np
import pytest

from harness import build_oracle


def program_factory(kind):
    import conex_b200
    if kind == "oracle":
        lib = conex_b200.bind_conex_abi(C.CDLL(build_oracle()))
        return lambda m: conex_b200.Conex(m, library=lib)
    return lambda m: conex_b200.Conex(m)


@pytest.fixture(params=["oracle", pytest.param("device", marks=pytest.mark.gpu)])
def Conex(request):
    return program_factory(request.param)


def randsym(rng, d):
    A = rng.standard_normal((d, d))
    return 0.5 * (A + A.T)


def check_errors(err, eps=1e-5):  # run_tests.py:40-43
    return err.Ax_minus_b < eps and abs(err.x_dot_s) < eps


def lp_instance():  # run_tests.py:48-60 (`randominstance`, deterministic there too)
    A = np.ones((3, 2))
    A[0, 1] = 3
    A[1, 0] = 4
    c = np.ones(3)
    return A, A.T @ c, c


def test1_lmi(Conex):  # run_tests.py:114-153
    rng = np.random.default_rng(1)
    m, n = 3, 4
    prog = Conex(m)
    Amat = np.stack([randsym(rng, n) for _ in range(m)], axis=2)
    Amat[:, :, m - 1] = 0
    Amat[0, 0, m - 1] = 1
    prog.AddDenseLinearMatrixInequality(Amat, np.eye(n))
    b = np.array([np.trace(Amat[:, :, i]) for i in range(m)])      # A' * eye(n)
    sol = prog.Maximize(b)
    assert sol.status
    s, err = prog.ComputeErrors(sol.y, prog.GetDualVariables(), b)
    assert check_errors(err)
    assert min(err.min_eig_S) > -1e-6 and min(err.min_eig_X) > -1e-6


def test2_random_instance_with_lp_and_lmi_blocks(Conex):  # run_tests.py:62-89
    rng = np.random.default_rng(2)
    A1, b, c1 = lp_instance()
    m, n = A1.shape[1], 4
    prog = Conex(m)
    prog.AddLinearInequality(A1, c1)
    prog.AddLinearInequality(A1.copy(), c1.copy())
    Amat = np.stack([randsym(rng, n) for _ in range(m)], axis=2)
    Amat[:, :, m - 1] = 0
    Amat[0, 0, m - 1] = 1
    prog.AddDenseLinearMatrixInequality(Amat, np.eye(n))
    sol = prog.Maximize(b)
    assert sol.status
    s, err = prog.ComputeErrors(sol.y, prog.GetDualVariables(), b)
    assert check_errors(err)


def test3_dual_fails_slater(Conex):  # run_tests.py:210-228
    prog = Conex(2)
    prog.AddLinearInequality(np.eye(2), np.ones(2))
    b = np.array([1.0, 0.0])
    sol = prog.Maximize(b)
    assert sol.status
    s, err = prog.ComputeErrors(sol.y, prog.GetDualVariables(), b)
    assert check_errors(err)


def test4_dual_infeasible(Conex):  # run_tests.py:191-208
    m = 2
    prog = Conex(m)
    prog.AddLinearInequality(np.vstack([np.eye(m), np.eye(m)]), np.ones(2 * m))
    assert prog.Maximize(np.array([0.0, -1.0])).status == 0


def test5_primal_infeasible(Conex):  # run_tests.py:230-247
    m = 2
    prog = Conex(m)
    prog.AddLinearInequality(np.vstack([np.eye(m), -np.eye(m)]), -np.ones(2 * m))
    assert prog.Maximize(np.ones(m)).status == 0


def test6_sparse_instance(Conex):  # run_tests.py:91-112
    rng = np.random.default_rng(6)
    prog = Conex(3)
    n = 4
    prog.AddSparseLinearMatrixInequality(np.stack([randsym(rng, n) for _ in range(2)], axis=2), np.eye(n), np.arange(0, 2))
    prog.AddSparseLinearMatrixInequality(np.stack([randsym(rng, n) for _ in range(2)], axis=2), np.eye(n), np.arange(1, 3))
    sol = prog.Maximize(np.ones(3))
    assert sol.status == 1
    s, err = prog.ComputeErrors(sol.y, prog.GetDualVariables(), np.ones(3))
    assert check_errors(err, 1e-4)


def test7_mu_is_non_increasing(Conex):  # run_tests.py:249-281
    m = 2
    prog = Conex(m)
    prog.AddLinearInequality(np.vstack([np.eye(m), np.eye(m)]), -np.ones(2 * m))
    config = prog.DefaultConfiguration()
    config.max_iterations = 6
    prog.Maximize(np.ones(m), config)
    assert prog.GetIterationNumberStats(-1).iteration_number + 1 <= config.max_iterations
    stats = prog.GetIterationStats()
    assert all(stats[i].mu <= stats[i - 1].mu for i in range(1, len(stats)))


def test8_hermitian_lmi_interface(Conex):  # run_tests.py:283-297
    prog = Conex(2)
    prog.NewLinearMatrixInequality(2, 1)
    with pytest.raises(NameError):
        prog.NewLinearMatrixInequality(-2, 2)       # invalid order


def test9_hermitian_lmi_known_answer(Conex):  # run_tests.py:299-321, real algebra
    order, num_vars = 3, 2
    prog = Conex(num_vars)
    con = prog.NewLinearMatrixInequality(order, 1)
    # 0 <= [[1, x, 0], [x, 2, y], [0, y, 1]]
    for i in range(num_vars):
        prog.UpdateLinearOperator(con, -1.0, i, i + 1, i, 0)
    for i in range(order):
        prog.UpdateAffineTerm(con, 2 if i == 1 else 1, i, i, 0)
    sol = prog.Maximize(-np.ones(num_vars))
    assert sol.status and np.linalg.norm(sol.y + np.ones(num_vars)) < 1e-6


def add_random_lmi(prog, rng, num_vars, order):  # run_tests.py:6-21, hyper_complex_dim = 1
    con = prog.NewLinearMatrixInequality(order, 1)
    b = np.zeros(num_vars)
    for v in range(num_vars):
        M = randsym(rng, order)
        for r in range(order):
            for c in range(r + 1):
                prog.UpdateLinearOperator(con, M[r, c], v, r, c, 0)
        b[v] = np.trace(M)
    for i in range(order):
        prog.UpdateAffineTerm(con, 1.0, i, i, 0)
    return b


def test10_random_hermitian_lmi(Conex):  # run_tests.py:323-332, real algebra
    rng = np.random.default_rng(10)
    prog = Conex(10)
    b = add_random_lmi(prog, rng, 10, 20)
    assert prog.Maximize(b).status


def test12_random_socp(Conex):  # run_tests.py:23-34, 348-356
    rng = np.random.default_rng(12)
    order, num_vars = 20, 10
    prog = Conex(num_vars)
    con = prog.NewLorentzConeConstraint(order)
    b = np.zeros(num_vars)
    for v in range(num_vars):
        col = rng.standard_normal(order + 1)
        for r in range(order + 1):
            prog.UpdateLinearOperator(con, col[r], v, r, 0, 0)
        b[v] = col[0]
    prog.UpdateAffineTerm(con, 1.0, 0, 0, 0)
    assert prog.Maximize(b).status


# ---- interfaces/matlab/test/run_conex_tests.m (LPTests, SDPTests, SparseTests) -----------------------
MATLAB_A = np.stack([np.eye(3), np.array([[1.0, 1, 0], [1, 1, 0], [0, 0, 0]])], axis=2)


def test_matlab_lp_tests(Conex):  # run_conex_tests.m:11-44
    rng = np.random.default_rng(3)
    A = rng.uniform(0, 1, size=(4, 2))
    prog = Conex(2)
    prog.AddLinearInequality(A, np.ones(4))
    b = A.T @ np.ones(4)
    assert prog.Maximize(b).status
    prog = Conex(2)                                   # 1 >= y1, 1 >= y2, -2 >= -y1 ... : infeasible
    prog.AddLinearInequality(np.array([[1.0, 0], [0, 1], [-1, 0], [0, -1]]), np.array([1.0, 1, -2, 1]))
    assert prog.Maximize(b).status == 0


def test_matlab_sdp_tests(Conex):  # run_conex_tests.m:46-68
    prog = Conex(2)
    prog.AddDenseLinearMatrixInequality(MATLAB_A, np.eye(3))
    sol = prog.Maximize(np.array([1.0, 1.0]))
    assert sol.status
    slack = np.eye(3) - sol.y[0] * MATLAB_A[:, :, 0] - sol.y[1] * MATLAB_A[:, :, 1]
    assert np.linalg.eigvalsh(slack).min() >= -1e-9


def test_matlab_sparse_tests(Conex):  # run_conex_tests.m:71-103
    prog = Conex(3)
    prog.AddSparseLinearMatrixInequality(MATLAB_A, np.eye(3), [0, 1])
    prog.AddSparseLinearMatrixInequality(MATLAB_A, np.eye(3), [1, 2])
    b = np.ones(3)
    sol = prog.Maximize(b)
    assert sol.status
    x = prog.GetDualVariables()
    for variables in ([0, 1], [1, 2]):
        slack = np.eye(3) - sum(sol.y[v] * MATLAB_A[:, :, k] for k, v in enumerate(variables))
        assert np.linalg.eigvalsh(slack).min() >= -1e-9
    Ax = np.zeros(3)
    Ax[0:2] += np.tensordot(MATLAB_A, x[0], axes=([0, 1], [0, 1]))
    Ax[1:3] += np.tensordot(MATLAB_A, x[1], axes=([0, 1], [0, 1]))
    assert np.linalg.norm(Ax - b) < 1e-8            # the reference asserts 1e-12 on its own run


def test_linear_inequalities_bookkeeping_of_dual_variables(Conex):
    """CONEX_AddLinearInequalities adds zero, one or two constraints (interfaces/conex.cc:190-215: an LP cone with
    one row per finite bound, an equality block for lb == ub). The wrapper must size every dual-variable buffer
    from what the library reports, and later constraint indices must not shift."""
    m = 3
    prog = Conex(m)
    A = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [1.0, 1.0, 1.0]])
    lb = np.array([-1.0, -2.0, -1e20, 0.5])   # row 0, 1: both bounds; row 2: upper only; row 3: equality
    ub = np.array([1.0, 2.0, 3.0, 0.5])
    prog.AddLinearInequalities(A, lb, ub)
    assert prog.num_constraints == 2            # LP cone with 5 rows + equality block with 1 row
    prog.AddLinearInequality(np.eye(m), 10 * np.ones(m))
    assert prog.num_constraints == 3
    sol = prog.Maximize(np.array([1.0, 1.0, 0.0]))
    assert sol.status
    x = prog.GetDualVariables()
    assert [xi.shape for xi in x][0] == (5, 1) and x[2].shape == (m, 1)
    assert abs(sol.y.sum() - 0.5) < 1e-7 and sol.y[0] <= 1 + 1e-6 and sol.y[1] <= 2 + 1e-6


@pytest.mark.gpu
def test_known_bad_instance_equality_constraint_failing_ldlt_on_the_device():
    """conex/test/solver_failures.cc:12-46 (see tests/test_oracle_cones.py): the regularised LDL^T of the
    device meets a singular leading block; same answer as the oracle."""
    import devlib
    from harness import oracle
    res = []
    for L in (oracle(), devlib.product()):
        P = L.program(2)
        P.add_equality(np.array([[1.0, -1.0]]), np.array([0.0]), [0, 1])
        P.add_linear(np.array([[1.0, 1.0]]), np.array([1.0]))
        res.append(P.maximize(np.array([1.0, 1.0]), L.default_config()))
    (so, yo), (sd, yd) = res
    assert so == sd == 1
    assert np.abs(yd - yo).max() < 1e-5 and abs(yd[0] - yd[1]) < 1e-8


@pytest.mark.gpu
def test_warmstart_from_converged_iterate_returns_the_same_point_on_the_device():
    """conex/test/test_warmstart.cc:47-79 (see tests/test_oracle_golden.py): LMI + LP, full solve, then a
    two-iteration warm start from the device-resident arena."""
    import devlib
    from harness import random_dense_lmi
    L = devlib.product()
    mats, Cm = random_dense_lmi(15, 13, 3)
    rng = np.random.default_rng(4)
    A, c = rng.uniform(-1, 1, size=(15, 13)), np.ones(15)
    P = L.program(13)
    P.add_dense_lmi(mats, Cm)
    P.add_linear(A, c)
    b = P.feasible_objective()
    kw = dict(final_centering_steps=3, final_centering_tolerance=.01)
    solved, y = P.maximize(b, L.default_config(**kw))
    solved_w, y_warm = P.maximize(b, L.default_config(initialization_mode=1, max_iterations=2, **kw))
    assert solved == 1 and solved_w == 1
    assert np.linalg.norm(y - y_warm) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["lp", "lmi"])
def test_iterative_refinement_on_the_device_matches_the_oracle(kind):
    """kkt_solver_options_test.cc:70-75 (`UseIterativeRefinement`, 3 correction solves per solve,
    kkt_solver.cc:248-261): same y as the oracle's restatement, and the LP properties hold."""
    import devlib
    from harness import oracle, random_dense_lmi
    rng = np.random.Generator(np.random.PCG64(171))
    res = []
    for L in (oracle(), devlib.product()):
        P = L.program()
        if kind == "lp":
            A = np.random.Generator(np.random.PCG64(171)).uniform(-1, 1, size=(8, 5))
            Cv = np.abs(np.random.Generator(np.random.PCG64(172)).uniform(-1, 1, size=8))
            P.add_linear(A, Cv)
            x0 = np.abs(np.random.Generator(np.random.PCG64(173)).uniform(-1, 1, size=8))
            b = A.T @ (x0 * 0.01 / np.linalg.norm(x0))
            cfg = L.default_config(prepare_dual_variables=1, inv_sqrt_mu_max=5e5, divergence_upper_bound=1000,
                                   dinf_upper_bound=1.35, final_centering_tolerance=1,
                                   iterative_refinement_iterations=3)
        else:
            mats, Cm = random_dense_lmi(12, 7, 9)
            P.add_dense_lmi(mats, Cm)
            b = P.feasible_objective()
            cfg = L.default_config(prepare_dual_variables=1, iterative_refinement_iterations=3)
        solved, y = P.maximize(b, cfg)
        res.append((solved, y, P))
    (so, yo, Po), (sd, yd, Pd) = res
    assert so == sd == 1
    assert np.abs(yo - yd).max() <= 1e-6 * max(1.0, np.abs(yo).max())
    if kind == "lp":
        x = Pd.dual_variable(0).ravel()
        assert np.linalg.norm(A.T @ x - b) <= 1e-9 * np.linalg.norm(b) and x.min() >= -1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kkt_kind", [1, 2])
def test_variables_specified_out_of_order_on_the_device(kkt_kind):
    """conex/test/assembly_test.cc:196-219 (see tests/test_oracle_golden.py) through the dense and the
    multifrontal KKT solver of the device."""
    import devlib
    from harness import oracle, random_sym
    rng = np.random.default_rng(8)
    n, m = 5, 4
    cones = []
    for variables in ([1, 0, 3], [1, 0, 2]):
        cones.append(([random_sym(rng, n) for _ in variables], np.eye(n), variables))
    res = []
    for L, kind in ((oracle(), None), (devlib.product(), kkt_kind)):
        P = L.program(m)
        if kind is not None:
            L.lib.CONEXB200_SetKKTSolverKind.argtypes = [C.c_void_p, C.c_int]
            L.lib.CONEXB200_SetKKTSolverKind.restype = None
            L.lib.CONEXB200_SetKKTSolverKind(P.h, kind)
        for mats, Cm, variables in cones:
            P.add_dense_lmi(mats, Cm, variables)
        H = P.newton_system(coldstart=True)[0]
        solved, y = P.maximize(P.feasible_objective(), L.default_config())
        res.append((np.tril(H), solved, y))
    (Ho, so, yo), (Hd, sd, yd) = res
    assert so == sd == 1
    assert np.abs(Ho - Hd).max() <= 1e-10 * np.abs(Ho).max()
    assert np.abs(yo - yd).max() <= 1e-6 * max(1.0, np.abs(yo).max())
