"""Chordal-sparse KKT systems.

CPU part: the symbolic step of the multifrontal solver (conex_b200/csrc/host/supernodal_kkt_solver.cc,
AnalyzeCliques — pure host code reached through CONEXB200_SupernodalAnalysis) on the clique patterns of
the reference's own tests (block_triangular_operations_test.cc:120-124,143-145,164-166,
supernodal_solver_test.cc:131-186, clique_ordering_test.cc) and on random patterns: the result must be
a valid elimination structure, and a numpy execution of the multifrontal factorisation / solves that it
prescribes must reproduce the dense Cholesky solution (the reference checks its block Cholesky against
Eigen::LLT the same way, block_triangular_operations_test.cc:102-118).

GPU part (-m gpu): programs whose cones act on overlapping subsets of the variables, solved through
CONEX_* with the multifrontal device solver, against the oracle (dense) and the device's dense solver."""
import ctypes as C

import numpy as np
import pytest

from harness import PRODUCT_SO, block_arrow_program, oracle, random_sym

REFERENCE_PATTERNS = [
    [[0, 1, 2], [2]],
    [[0, 1, 2, 4, 7], [3, 4], [5, 6, 7]],
    [[0, 1, 5], [1, 2, 5], [3, 4, 5]],
    [[0, 1, 2], [1, 2, 3], [3, 4, 2]],
    [[0, 1], [2, 4], [3, 4], [5, 6, 7], [7, 8, 9, 10]],
    [[0, 1, 2, 3], [3, 4, 5]],
    [[0, 1, 2, 3]],
    [[0, 1, 2, 3], [3, 4], [4, 5, 6]],
    [[0, 1, 2, 5], [3, 4, 5], [5, 6]],
    [[0, 1, 2, 4], [3, 4], [5, 6, 7]],
    [[0, 1, 2, 5], [1, 4, 2, 5], [3, 4, 5]],
    # not chordal as given (a 4-cycle of pairwise overlaps): needs fill
    [[0, 1], [1, 2], [2, 3], [3, 0]],
    [[0, 1, 2], [2, 3, 4], [4, 5, 0], [1, 3, 5]],
    # clique_ordering_test.cc:76-127: PerfectEliminationOrderFound (two forests with unused indices),
    # SmallSize, ...Diagonal (five singletons), FillIn (the 4-cycle again, as the reference orders it),
    # Nonmaximal (nested cliques)
    [[1, 2, 3, 5], [3, 4, 5], [4, 5, 6, 7], [8, 9], [1, 11]],
    [[0, 2, 3, 5], [3, 4, 5], [4, 5, 6, 7], [0, 11]],
    [[0, 1]],
    [[0, 1], [1, 2]],
    [[1], [2], [3], [4], [5]],
    [[0, 1], [1, 2], [0, 3], [2, 3]],
    [[0, 1], [0, 1, 2], [0, 1, 2, 3, 4]],
]


def analysis(N, cliques):
    L = C.CDLL(PRODUCT_SO)
    ip = C.POINTER(C.c_int)
    L.CONEXB200_SupernodalAnalysis.argtypes = [C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, ip, ip, C.c_int,
                                               C.POINTER(C.c_double)]
    ptr = np.zeros(len(cliques) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(c) for c in cliques])
    flat = np.array([v for c in cliques for v in c], dtype=np.int32)
    position, node_of = np.zeros(N, np.int32), np.zeros(N, np.int32)
    node_ptr, node_vars = np.zeros(N + 1, np.int32), np.zeros(N, np.int32)
    cap = N * N + 1
    sep_ptr, sep_vars = np.zeros(N + 1, np.int32), np.zeros(cap, np.int32)
    flops = np.zeros(2)
    as_ip = lambda a: a.ctypes.data_as(ip)  # noqa: E731
    nodes = L.CONEXB200_SupernodalAnalysis(N, len(cliques), as_ip(ptr), as_ip(flat), as_ip(position),
                                           as_ip(node_of), as_ip(node_ptr), as_ip(node_vars), as_ip(sep_ptr),
                                           as_ip(sep_vars), cap, flops.ctypes.data_as(C.POINTER(C.c_double)))
    assert nodes >= 1
    supers = [node_vars[node_ptr[k]:node_ptr[k + 1]].tolist() for k in range(nodes)]
    seps = [sep_vars[sep_ptr[k]:sep_ptr[k + 1]].tolist() for k in range(nodes)]
    return position, node_of, supers, seps, flops


def check_structure(N, cliques, position, node_of, supers, seps):
    assert sorted(position.tolist()) == list(range(N))                       # a permutation
    assert sorted(v for s in supers for v in s) == list(range(N))            # eliminated exactly once
    pos = 0
    for k, (s, p) in enumerate(zip(supers, seps)):
        assert len(s) >= 1
        for v in s:                                                           # nodes in elimination order
            assert position[v] == pos and node_of[v] == k
            pos += 1
        assert all(position[v] > position[s[-1]] for v in p)                  # separators come later
        assert [position[v] for v in p] == sorted(position[v] for v in p)
    front = [set(s) | set(p) for s, p in zip(supers, seps)]
    for c in cliques:                                                         # every coupled pair has a home
        for u in c:
            for v in c:
                first = u if position[u] <= position[v] else v
                assert {u, v} <= front[node_of[first]]
    for k, p in enumerate(seps):                                              # closed under elimination
        for u in p:
            for v in p:
                first = u if position[u] <= position[v] else v
                assert {u, v} <= front[node_of[first]], (k, u, v)


def pattern_matrix(N, cliques, rng):
    H = np.zeros((N, N))
    for c in cliques:
        R = rng.standard_normal((len(c), len(c) + 2))
        H[np.ix_(c, c)] += R @ R.T
    H += 0.5 * np.eye(N)
    return H


def multifrontal_solve(H, b, position, supers, seps):
    """The factorisation and solves exactly as SupernodalKKTSolver runs them, in numpy."""
    fronts = []
    for s, p in zip(supers, seps):
        rows = s + p
        fronts.append((rows, H[np.ix_(rows, s)].copy()))                      # assembled front (all columns of s)
    where = {}
    for k, (rows, _) in enumerate(fronts):
        for ci, v in enumerate(supers[k]):
            for ri, u in enumerate(rows):
                where[(u, v)] = (k, ri, ci)
    for k, (rows, F) in enumerate(fronts):
        sk = len(supers[k])
        L11 = np.linalg.cholesky(np.tril(F[:sk]) + np.tril(F[:sk], -1).T)
        F[:sk] = L11
        if len(rows) > sk:
            F[sk:] = np.linalg.solve(L11, F[sk:].T).T
            U = F[sk:] @ F[sk:].T
            p = seps[k]
            for a, u in enumerate(p):
                for bb, v in enumerate(p):
                    if position[u] < position[v]:
                        continue                                              # lower triangle: u eliminated after v
                    kk, ri, ci = where[(u, v)]
                    assert kk > k
                    fronts[kk][1][ri, ci] -= U[a, bb]
    x = b.astype(float).copy()
    for k, (rows, F) in enumerate(fronts):
        s, p = supers[k], seps[k]
        x[s] = np.linalg.solve(np.tril(F[:len(s)]), x[s])
        if p:
            x[p] -= F[len(s):] @ x[s]
    for k in reversed(range(len(fronts))):
        rows, F = fronts[k]
        s, p = supers[k], seps[k]
        if p:
            x[s] -= F[len(s):].T @ x[p]
        x[s] = np.linalg.solve(np.tril(F[:len(s)]).T, x[s])
    return x


@pytest.mark.parametrize("idx", range(len(REFERENCE_PATTERNS)))
def test_reference_clique_patterns(idx):
    cliques = REFERENCE_PATTERNS[idx]
    N = max(max(c) for c in cliques) + 1
    position, node_of, supers, seps, flops = analysis(N, cliques)
    check_structure(N, cliques, position, node_of, supers, seps)
    rng = np.random.default_rng(idx)
    # indices that no clique mentions (the reference's patterns skip some) get a diagonal entry only
    used = set(v for c in cliques for v in c)
    H = pattern_matrix(N, cliques + [[v] for v in range(N) if v not in used], rng)
    b = rng.standard_normal(N)
    x = multifrontal_solve(H, b, position, supers, seps)
    assert np.abs(x - np.linalg.solve(H, b)).max() < 1e-11 * max(1.0, np.abs(x).max())


@pytest.mark.parametrize("seed", range(12))
def test_random_clique_patterns(seed):
    rng = np.random.default_rng(100 + seed)
    N = int(rng.integers(5, 60))
    cliques = []
    for _ in range(int(rng.integers(1, 14))):
        size = int(rng.integers(1, min(N, 9) + 1))
        cliques.append(sorted(rng.choice(N, size=size, replace=False).tolist()))
    covered = set(v for c in cliques for v in c)
    cliques += [[v] for v in range(N) if v not in covered]                    # every variable constrained
    position, node_of, supers, seps, flops = analysis(N, cliques)
    check_structure(N, cliques, position, node_of, supers, seps)
    H = pattern_matrix(N, cliques, rng)
    b = rng.standard_normal(N)
    x = multifrontal_solve(H, b, position, supers, seps)
    assert np.abs(x - np.linalg.solve(H, b)).max() < 1e-10 * max(1.0, np.abs(x).max())


def test_chain_and_arrow_structures_save_flops():
    # banded chain: 40 cliques of 12 variables overlapping by 4 -> 40 small supernodes
    chain = [list(range(8 * k, 8 * k + 12)) for k in range(40)]
    N = 8 * 39 + 12
    position, node_of, supers, seps, flops = analysis(N, chain)
    check_structure(N, chain, position, node_of, supers, seps)
    assert len(supers) == 40 and max(len(p) for p in seps) == 4
    assert flops[0] < 0.01 * flops[1]
    # block arrow: 6 blocks of 50 private variables, all coupled to the same 10 shared variables
    arrow = [list(range(50 * k, 50 * k + 50)) + list(range(300, 310)) for k in range(6)]
    position, node_of, supers, seps, flops = analysis(310, arrow)
    check_structure(310, arrow, position, node_of, supers, seps)
    assert sorted(len(s) for s in supers) == [50] * 5 + [60]
    assert flops[0] < 0.1 * flops[1]
    # one clique on everything: a single dense supernode
    position, node_of, supers, seps, flops = analysis(7, [list(range(7)), [1, 2]])
    assert len(supers) == 1 and seps == [[]]


# ---- device -------------------------------------------------------------------------------------------

def chain_program(links, width, overlap, order, seed):
    rng = np.random.default_rng(seed)
    step = width - overlap
    m = step * (links - 1) + width
    cones = []
    for k in range(links):
        variables = list(range(k * step, k * step + width))
        mats = [random_sym(rng, order) for _ in variables]
        cones.append((mats, np.eye(order), variables))
    return m, cones


def lp_chain_program(links=50, width=5, rows=10, seed=1):
    """The reference's own sparse test (conex/test/test_lp.cc:133-214, `LP Sparse`): 50 LP blocks of 10
    rows on 5 variables each, consecutive blocks sharing one variable (201 variables), c_i = 1 + 0.01 i.
    The C ABI has no LP constraint on a variable subset (the reference test uses the C++ AddConstraint),
    so every block is given as the equivalent LMI with diagonal matrices."""
    rng = np.random.default_rng(seed)
    cones, start = [], 0
    for i in range(links):
        variables = list(range(start, start + width))
        start = variables[-1]
        A = rng.uniform(-1, 1, size=(rows, width))
        mats = [np.diag(A[:, j]) for j in range(width)]
        cones.append((mats, (1.0 + 0.01 * i) * np.eye(rows), variables))
    return start + 1, cones


def solve_with(L, m, cones, kind=None):
    P = L.program(m)
    if kind is not None:
        L.lib.CONEXB200_SetKKTSolverKind.argtypes = [C.c_void_p, C.c_int]
        L.lib.CONEXB200_SetKKTSolverKind.restype = None
        L.lib.CONEXB200_SetKKTSolverKind(P.h, kind)
    for mats, Cm, variables in cones:
        P.add_dense_lmi(mats, Cm, variables)
    b = P.feasible_objective()
    solved, y = P.maximize(b, L.default_config(prepare_dual_variables=1))
    nodes = L.lib.CONEXB200_GetNumberOfSupernodes(P.h) if kind is not None else 1
    return P, solved, y, b, nodes


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["arrow", "chain", "chain_long", "disconnected", "lp_chain", "cycle_fill_in"])
def test_sparse_programs_match_oracle_and_dense_solver(shape):
    import devlib
    dev, ora = devlib.product(), oracle()
    if shape == "arrow":
        m, cones = block_arrow_program(blocks=4, private=6, shared=3, order=7, seed=1)
    elif shape == "chain":
        m, cones = chain_program(links=5, width=6, overlap=2, order=6, seed=2)
    elif shape == "chain_long":
        m, cones = chain_program(links=30, width=5, overlap=2, order=5, seed=3)
    elif shape == "lp_chain":
        m, cones = lp_chain_program()
        assert m == 201
    elif shape == "cycle_fill_in":
        # test_lp.cc:232-309 (`LP SparseWithFillIn`): cliques {0,1}, {1,2}, {2,3}, {0,3} — a 4-cycle, which
        # is not chordal: the symbolic step has to add fill. LMI cones of order 4 on those pairs.
        rng = np.random.default_rng(5)
        m, cones = 4, []
        for variables in ([0, 1], [1, 2], [2, 3], [0, 3]):
            cones.append(([random_sym(rng, 4) for _ in variables], np.eye(4), variables))
    else:
        m, cones = chain_program(links=3, width=4, overlap=0, order=5, seed=4)
    Po, so, yo, bo, _ = solve_with(ora, m, cones)
    Pd, sd, yd, bd, nd = solve_with(dev, m, cones, kind=1)
    Ps, ss, ys, bs, ns = solve_with(dev, m, cones, kind=2)
    assert nd == 1 and ns > 1
    assert so == sd == ss == 1
    assert np.allclose(bo, bs, rtol=1e-12, atol=1e-14)
    lo, ld, ls = Po.iteration_log(), Pd.iteration_log(), Ps.iteration_log()
    assert abs(len(lo) - len(ls)) <= 1 and abs(len(ld) - len(ls)) <= 1
    for ref in (lo, ld):
        assert abs(ref[-1]["by"] - ls[-1]["by"]) <= 1e-7 * max(1.0, abs(ref[-1]["by"]))
    assert np.abs(yo - ys).max() <= 1e-6 * max(1.0, np.abs(yo).max())
    assert np.abs(yd - ys).max() <= 1e-6 * max(1.0, np.abs(yd).max())
    if shape == "lp_chain":
        # the reference's checks on its sparse solve (test_lp.cc:186-202): slack >= -eps, A'x = b
        assert ns == 50
        Ax = np.zeros(m)
        for i, (mats, Cm, variables) in enumerate(cones):
            slack = Cm - sum(ys[v] * A for v, A in zip(variables, mats))
            assert np.diag(slack).min() >= -1e-8
            X = Ps.dual_variable(i)
            for v, A in zip(variables, mats):
                Ax[v] += np.sum(A * X)
        assert np.linalg.norm(Ax - bs) < 1e-8 * max(1.0, np.linalg.norm(bs))
    # the assembled Newton systems agree entry by entry (dense export of the fronts)
    Hd = Pd.newton_system(coldstart=True)[0]
    Hs = Ps.newton_system(coldstart=True)[0]
    Ho = Po.newton_system(coldstart=True)[0]
    scale = np.sqrt(np.outer(np.diag(Ho), np.diag(Ho)))
    assert (np.abs(np.tril(Hs) - np.tril(Ho)) / scale).max() < 1e-10
    assert (np.abs(np.tril(Hs) - np.tril(Hd)) / scale).max() < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["chain", "arrow", "two_blocks_on_one_clique"])
def test_sparse_programs_with_equality_constraints_take_ldlt_fronts(shape):
    """Equality constraints (conex/test/equality_constraints_test.cc: Basic / Many / ManySeparate shapes) placed on the
    cliques of chordal-sparse programs: the multipliers are extra unknowns of their clique, the fronts are factored
    by the regularised LDL^T with pivoting inside every supernode (reference BlockLDLTInPlace,
    block_triangular_operations.cc:315-349, driven by kkt_solver.cc:180-193) instead of falling back to one dense
    supernode. Oracle vs dense device solver vs multifrontal device solver."""
    import devlib
    dev, ora = devlib.product(), oracle()
    rng = np.random.default_rng(17)
    if shape == "arrow":
        m, cones = block_arrow_program(blocks=4, private=6, shared=3, order=7, seed=1)
        eq_sets = [[0, 1, 2], [24, 25, 26, 7]]          # inside cone 0; shared variables + one of cone 1
    else:
        m, cones = chain_program(links=5, width=6, overlap=2, order=6, seed=2)
        eq_sets = [[4, 5, 6]] if shape == "chain" else [[8, 9, 10], [9, 11, 12, 13]]   # one / two blocks on clique 2
    y0 = 0.02 * rng.standard_normal(m)
    equalities = []
    for variables in eq_sets:
        rows = 2 if len(variables) > 3 else 1
        A = rng.uniform(-1, 1, size=(rows, len(variables)))
        equalities.append((A, A @ y0[variables], variables))
    out = []
    for L, kind in ((ora, None), (dev, 1), (dev, 2)):
        P = L.program(m)
        if kind is not None:
            L.lib.CONEXB200_SetKKTSolverKind.argtypes = [C.c_void_p, C.c_int]
            L.lib.CONEXB200_SetKKTSolverKind.restype = None
            L.lib.CONEXB200_SetKKTSolverKind(P.h, kind)
        for mats, Cm, variables in cones:
            P.add_dense_lmi(mats, Cm, variables)
        for A, beq, variables in equalities:
            P.add_equality(A, beq, variables)
        b = P.feasible_objective()
        solved, y = P.maximize(b, L.default_config())
        nodes = L.lib.CONEXB200_GetNumberOfSupernodes(P.h) if kind is not None else 1
        out.append((solved, y, P.iteration_log(), nodes))
    (so, yo, lo, _), (sd, yd, ld, nd), (ss, ys, ls, ns) = out
    assert nd == 1 and ns > 1
    assert so == sd == ss == 1
    assert abs(len(lo) - len(ls)) <= 1 and abs(len(ld) - len(ls)) <= 1
    for ref in (lo, ld):
        assert abs(ref[-1]["by"] - ls[-1]["by"]) <= 1e-7 * max(1.0, abs(ref[-1]["by"]))
    assert np.abs(yo - ys).max() <= 1e-6 * max(1.0, np.abs(yo).max())
    assert np.abs(yd - ys).max() <= 1e-6 * max(1.0, np.abs(yd).max())
    for A, beq, variables in equalities:
        assert np.abs(A @ ys[variables] - beq).max() <= 1e-8


@pytest.mark.gpu
def test_iterative_refinement_on_a_sparse_program_takes_the_dense_solver():
    """A configuration that is valid in the reference (kkt_solver.cc:248-261, kkt_solver_options_test.cc) must not
    turn into "not solved" because the clique structure made the library pick the multifrontal solver: with
    iterative_refinement_iterations > 0 and the default solver kind the program is solved by the dense solver."""
    import devlib
    dev, ora = devlib.product(), oracle()
    # m = 330 >= 256: the multifrontal solver pays; order 14: 105 free entries per cone >= its 90 variables
    m, cones = block_arrow_program(blocks=4, private=80, shared=10, order=14, seed=3)
    out = []
    for L, kind in ((ora, None), (dev, 0)):
        P = L.program(m)
        if kind is not None:
            L.lib.CONEXB200_SetKKTSolverKind.argtypes = [C.c_void_p, C.c_int]
            L.lib.CONEXB200_SetKKTSolverKind.restype = None
            L.lib.CONEXB200_SetKKTSolverKind(P.h, kind)
        for mats, Cm, variables in cones:
            P.add_dense_lmi(mats, Cm, variables)
        b = P.feasible_objective()
        Q = None
        if kind is not None:   # without refinement the same program does take the multifrontal solver
            Q = L.program(m)
            for mats, Cm, variables in cones:
                Q.add_dense_lmi(mats, Cm, variables)
            assert Q.maximize(b, L.default_config())[0] == 1
            assert L.lib.CONEXB200_GetNumberOfSupernodes(Q.h) > 1
        solved, y = P.maximize(b, L.default_config(iterative_refinement_iterations=1))
        out.append((solved, y, P.iteration_log()))
        if kind is not None:
            assert L.lib.CONEXB200_GetNumberOfSupernodes(P.h) == 1
    (so, yo, lo), (sd, yd, ld) = out
    assert so == sd == 1 and abs(len(lo) - len(ld)) <= 1
    assert abs(lo[-1]["by"] - ld[-1]["by"]) <= 1e-7 * max(1.0, abs(lo[-1]["by"]))
    assert np.abs(yo - yd).max() <= 1e-6 * max(1.0, np.abs(yo).max())


@pytest.mark.gpu
def test_supernodal_factor_and_solve_against_lapack():
    """A bigger block-arrow system (supernodes of 300-340 unknowns, beyond one 128-column block)
    through the solver's own factor / solve calls, driven by LMI cones whose G is random SPD."""
    import devlib
    dev = devlib.product()
    m, cones = block_arrow_program(blocks=3, private=300, shared=40, order=30, seed=7)
    P, solved, y, b, nodes = solve_with(dev, m, cones, kind=2)
    Pd, solved_d, yd, bd, _ = solve_with(dev, m, cones, kind=1)
    assert nodes == 3 and solved == solved_d == 1
    assert abs(P.iteration_log()[-1]["by"] - Pd.iteration_log()[-1]["by"]) <= 1e-7 * abs(Pd.iteration_log()[-1]["by"])
    assert np.abs(y - yd).max() <= 1e-6 * max(1.0, np.abs(yd).max())


def test_analysis_stays_fast_when_many_cones_share_variables():
    """400 blocks of 6 private variables, all coupled to the same 5 shared ones: every shared variable
    sits in 400 cliques (beyond the all-pairs threshold of the clique-graph weights)."""
    import time
    blocks, private, shared = 400, 6, 5
    N = blocks * private + shared
    cliques = [list(range(k * private, (k + 1) * private)) + list(range(N - shared, N)) for k in range(blocks)]
    t0 = time.perf_counter()
    position, node_of, supers, seps, flops = analysis(N, cliques)
    assert time.perf_counter() - t0 < 5.0
    check_structure(N, cliques, position, node_of, supers, seps)
    assert len(supers) == blocks and max(len(p) for p in seps) == shared
    assert sorted(len(s) for s in supers)[-1] == private + shared             # the root also eliminates the shared ones
    rng = np.random.default_rng(0)
    H = pattern_matrix(N, cliques, rng)
    b = rng.standard_normal(N)
    x = multifrontal_solve(H, b, position, supers, seps)
    assert np.abs(x - np.linalg.solve(H, b)).max() < 1e-9 * max(1.0, np.abs(x).max())
