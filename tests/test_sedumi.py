"""SeDuMi-format front end (conex_b200/sedumi.py, mirror of interfaces/matlab/conex.m and its util/
helpers) on CPU: the front end is written over the program interface of the reference's Python class, so
here it drives the ORACLE through a four-method adapter; the product class `conex_b200.Conex` offers the
same four methods (the GPU test at the bottom runs it on the device)."""
import numpy as np
import pytest

from harness import oracle, random_sym


class OracleProgram:
    """AddDense/AddSparseLinearMatrixInequality, Maximize, GetDualVariables over the oracle's CONEX_* ABI."""

    class Solution:
        pass

    def __init__(self, m):
        self.P = oracle().program(m)
        self.count = 0

    def AddDenseLinearMatrixInequality(self, A, c):
        self.P.add_dense_lmi([A[:, :, i] for i in range(A.shape[2])], np.asarray(c))
        self.count += 1

    def AddSparseLinearMatrixInequality(self, A, c, variables):
        self.P.add_dense_lmi([A[:, :, i] for i in range(A.shape[2])], np.asarray(c), list(variables))
        self.count += 1

    def Maximize(self, b):
        from conex_b200 import sedumi
        cfg = sedumi.reference_options(oracle().default_config())
        sol = OracleProgram.Solution()
        sol.status, sol.y = self.P.maximize(np.asarray(b), cfg)
        return sol

    def GetDualVariables(self):
        return [self.P.dual_variable(i) for i in range(self.count)]


def sedumi_problem(sizes, m, seed, couple=None):
    """Random strictly feasible pair: A rows = vec of symmetric blocks, x0 = identity blocks, b = A x0,
    c = A'y0 + vec(I-ish slack). couple[i] = blocks that row i touches (default: all)."""
    rng = np.random.default_rng(seed)
    N = sum(n * n for n in sizes)
    A = np.zeros((m, N))
    for i in range(m):
        s = 0
        for k, n in enumerate(sizes):
            if couple is None or k in couple[i]:
                A[i, s:s + n * n] = random_sym(rng, n).ravel(order="F")
            s += n * n
    x0 = np.concatenate([np.eye(n).ravel(order="F") for n in sizes])
    y0 = rng.uniform(-1, 1, size=m) * 0.1
    c = A.T @ y0 + np.concatenate([np.eye(n).ravel(order="F") for n in sizes])
    return A, A @ x0, c, {"s": list(sizes)}


def check_optimality(A, b, c, K, x, y, tol=1e-5):
    assert np.linalg.norm(A @ x - b) <= tol * max(1.0, np.linalg.norm(b))          # primal feasibility
    s, off = c - A.T @ y, 0
    for n in K["s"]:
        S = s[off:off + n * n].reshape(n, n, order="F")
        X = x[off:off + n * n].reshape(n, n, order="F")
        assert np.linalg.eigvalsh(0.5 * (S + S.T)).min() >= -tol                    # dual feasibility
        assert np.linalg.eigvalsh(0.5 * (X + X.T)).min() >= -tol
        off += n * n
    assert abs(c @ x - b @ y) <= 1e-4 * max(1.0, abs(b @ y))                        # duality gap


def test_single_block_is_the_dense_lmi_call():
    from conex_b200 import sedumi
    A, b, c, K = sedumi_problem([6], 4, 1)
    x, y, info = sedumi.solve(A, b, c, K, new_program=OracleProgram, errors=True)
    assert info["pinf"] == info["dinf"] == 0 and info["blocks"] == [6]
    check_optimality(A, b, c, K, x, y)
    # the same program given directly through the ABI
    P = oracle().program(4)
    P.add_dense_lmi([A[i].reshape(6, 6, order="F") for i in range(4)], c.reshape(6, 6, order="F"))
    solved, y_direct = P.maximize(b, sedumi.reference_options(oracle().default_config()))
    assert solved == 1 and np.abs(y - y_direct).max() < 1e-12
    assert info["errors"][0] < 1e-4


def test_several_blocks_become_lmis_on_variable_subsets():
    from conex_b200 import sedumi
    # chain coupling: row i touches blocks i and i + 1 only
    sizes, m = [4, 3, 5, 4], 8
    couple = [{i % 4, (i + 1) % 4} if i < 6 else {i % 4} for i in range(m)]
    A, b, c, K = sedumi_problem(sizes, m, 2, couple)
    x, y, info = sedumi.solve(A, b, c, K, new_program=OracleProgram)
    assert info["pinf"] == 0 and info["blocks"] == sizes
    check_optimality(A, b, c, K, x, y)
    cons = sedumi.extract_constraints(sedumi.symmetrize(A, sizes), sedumi.symmetrize(c, sizes)[0], sizes)
    assert [sorted(con["variables"].tolist()) for con in cons] == \
        [sorted(i for i in range(m) if k in couple[i]) for k in range(4)]


def test_zero_rows_are_removed_and_mapped_back():
    from conex_b200 import sedumi
    A, b, c, K = sedumi_problem([5], 3, 3)
    A2 = np.vstack([A[:1], np.zeros((1, A.shape[1])), A[1:]])
    b2 = np.concatenate([b[:1], [0.0], b[1:]])
    x, y, _ = sedumi.solve(A, b, c, K, new_program=OracleProgram)
    x2, y2, _ = sedumi.solve(A2, b2, c, K, new_program=OracleProgram)
    assert y2[1] == 0.0 and np.abs(np.delete(y2, 1) - y).max() < 1e-12 and np.abs(x2 - x).max() < 1e-12
    Ak, bk, T = sedumi.clean_linear(A2, b2)
    assert Ak.shape[0] == 3 and T.shape == (4, 3) and np.array_equal(T @ np.arange(1.0, 4.0), [1, 0, 2, 3])


def test_block_diagonal_structure_inside_a_block_is_split():
    from conex_b200 import sedumi
    # one 7 x 7 block whose data only couples indices {0, 2, 5} and {1, 3, 4, 6}
    rng = np.random.default_rng(4)
    groups, n, m = [[0, 2, 5], [1, 3, 4, 6]], 7, 5
    A = np.zeros((m, n * n))
    for i in range(m):
        M = np.zeros((n, n))
        for g in groups:
            M[np.ix_(g, g)] = random_sym(rng, len(g))
        A[i] = M.ravel(order="F")
    x0 = np.eye(n).ravel(order="F")
    c = A.T @ (rng.uniform(-1, 1, size=m) * 0.1) + np.eye(n).ravel(order="F")
    K = {"s": [n]}
    b = A @ x0
    x, y, info = sedumi.solve(A, b, c, K, new_program=OracleProgram, blkdiag=True)
    assert sorted(info["blocks"]) == [3, 4]
    check_optimality(A, b, c, K, x, y)
    X = x.reshape(n, n, order="F")
    assert np.all(X[np.ix_(groups[0], groups[1])] == 0)                      # scattered back with zeros
    x1, y1, info1 = sedumi.solve(A, b, c, K, new_program=OracleProgram, blkdiag=False)
    assert info1["blocks"] == [7]
    assert np.abs(y - y1).max() < 1e-5 and abs(b @ y - b @ y1) < 1e-7 * max(1.0, abs(b @ y1))


def test_extract_constraints_reference_fixture():
    """interfaces/matlab/test/test_extract_constraints.m: K.s = [2; 2], the 3 x 8 matrix given there
    (entries below 0.5 zeroed): per block, the variables that touch it and their matrices."""
    from conex_b200 import sedumi
    A = np.array([[1, 2, 2, 1, 0, 0, 0, 0],
                  [0, 0, 0, 0, 2, 1, 1, 2],
                  [1, 3, 3, 1, 2, -3, -3, 2]], dtype=np.float64)
    A[A < .5] = 0
    c = np.random.default_rng(0).standard_normal(8)
    cons = sedumi.extract_constraints(A, c, [2, 2])
    assert cons[0]["variables"].tolist() == [0, 2] and cons[1]["variables"].tolist() == [1, 2]
    assert np.array_equal(cons[0]["mats"], [[[1, 2], [2, 1]], [[1, 3], [3, 1]]])
    assert np.array_equal(cons[1]["mats"], [[[2, 1], [1, 2]], [[2, 0], [0, 2]]])
    for k, con in enumerate(cons):
        assert np.array_equal(con["affine"], c[4 * k:4 * k + 4].reshape(2, 2, order="F"))


def test_unsupported_cones_are_rejected_like_the_reference():
    from conex_b200 import sedumi
    A, b, c, K = sedumi_problem([3], 2, 5)
    for field in ("l", "q", "r", "f"):
        with pytest.raises(sedumi.SedumiError):
            sedumi.solve(A, b, c, dict(K, **{field: 2}), new_program=OracleProgram)
    sedumi.solve(A, b, c, dict(K, l=0, q=[]), new_program=OracleProgram)      # empty fields are fine
    with pytest.raises(sedumi.SedumiError):
        sedumi.solve(A[:, :-1], b, c[:-1], K, new_program=OracleProgram)


@pytest.mark.gpu
def test_front_end_on_the_device_matches_the_oracle():
    import conex_b200
    from conex_b200 import sedumi
    sizes, m = [4, 3, 5, 4], 8
    couple = [{i % 4, (i + 1) % 4} if i < 6 else {i % 4} for i in range(m)]
    A, b, c, K = sedumi_problem(sizes, m, 2, couple)
    xo, yo, _ = sedumi.solve(A, b, c, K, new_program=OracleProgram)
    xd, yd, info = sedumi.solve(A, b, c, K, new_program=conex_b200.Conex)
    assert info["pinf"] == 0
    check_optimality(A, b, c, K, xd, yd)
    assert np.abs(yd - yo).max() <= 1e-5 * max(1.0, np.abs(yo).max())
    assert abs(b @ yd - b @ yo) <= 1e-7 * max(1.0, abs(b @ yo))
