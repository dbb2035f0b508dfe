"""SeDuMi-format front end: `x, y, info = solve(A, b, c, K)`.

Mirror of the reference's MATLAB entry point `interfaces/matlab/conex.m:1-89` and its helpers
(`util/CleanLinear.m`, `util/ExtractConstraintMatrices.m`, `util/ConexPreprocess.m`, `util/blkdiagPrg.m`)
over the same program interface the reference's Python class offers (`AddDenseLinearMatrixInequality`,
`AddSparseLinearMatrixInequality`, `Maximize`, `GetDualVariables`; `conex_b200.Conex` here, any object
with those four methods in the tests).

SeDuMi's pair is   min c'x  s.t.  A x = b, x in K     /     max b'y  s.t.  c - A'y in K,
with x the concatenation of the column-major vectorised PSD blocks `K['s'] = [n_1, n_2, ...]`. conex solves
the right-hand problem: row i of A, cut into blocks, is the constraint matrix A_i of each block, c the
affine term. Like the reference, only PSD blocks are accepted (`conex.m:7-15` raises for K.l, K.q, K.r).

Several blocks (the reference's default `pars.blkdiag = length(K.s) > 1`): every block becomes an LMI on
the subset of the variables whose rows touch it (`ExtractConstraintMatrices.m:29-36`), i.e.
`AddSparseLinearMatrixInequality` — the chordal-sparse path of the KKT solver. Before that each block is
split along the connected components of its aggregate sparsity pattern |C| + sum_i |A_i|: a
block-diagonal slack is PSD iff its diagonal blocks are. (The reference's `BuildMask.m` goes further — a
facial-reduction style subspace closure; that part is not mirrored.)
"""
import numpy as _np


class SedumiError(ValueError):
    pass


def _nontrivial(K, field):
    v = K.get(field)
    if v is None:
        return False
    v = _np.atleast_1d(_np.asarray(v))
    return v.size > 0 and bool(_np.any(v > 0))


def clean_linear(A, b):
    """util/CleanLinear.m (default branch): drop the equations whose row of [A, b] is zero. Returns the
    kept rows and T (original rows x kept rows) mapping reduced dual variables back (removed ones get 0)."""
    A = _np.asarray(A, dtype=_np.float64)
    b = _np.asarray(b, dtype=_np.float64).ravel()
    if A.shape[0] != b.shape[0]:
        raise SedumiError("Number of rows of A and b do not match.")
    keep = _np.flatnonzero(_np.any(_np.hstack([A, b[:, None]]) != 0, axis=1))
    T = _np.zeros((A.shape[0], keep.size))
    T[keep, _np.arange(keep.size)] = 1.0
    return A[keep], b[keep], T


def _blocks(K):
    sizes = [int(n) for n in _np.atleast_1d(_np.asarray(K["s"])).ravel() if int(n) > 0]
    if not sizes:
        raise SedumiError("K.s is empty")
    return sizes


def symmetrize(rows, sizes):
    """coneBase.Symmetrize: every n x n block of every row replaced by its symmetric part."""
    rows = _np.array(rows, dtype=_np.float64, ndmin=2, copy=True)
    s = 0
    for n in sizes:
        blk = rows[:, s:s + n * n].reshape(rows.shape[0], n, n)
        rows[:, s:s + n * n] = (0.5 * (blk + blk.transpose(0, 2, 1))).reshape(rows.shape[0], n * n)
        s += n * n
    return rows


def split_blocks(A, c, sizes):
    """Splits every PSD block along the connected components of its aggregate sparsity pattern.
    Returns (A_r, c_r, sizes_r, index): column j of the reduced problem is column index[j] of the
    original one (blkdiagPrg.RecoverPrimal scatters x back through the same index list)."""
    cols, new_sizes = [], []
    s = 0
    for n in sizes:
        pattern = (c[s:s + n * n] != 0) | _np.any(A[:, s:s + n * n] != 0, axis=0)
        P = pattern.reshape(n, n)
        P = P | P.T | _np.eye(n, dtype=bool)
        seen = _np.zeros(n, dtype=bool)
        for start in range(n):
            if seen[start]:
                continue
            comp, stack = [], [start]
            seen[start] = True
            while stack:
                u = stack.pop()
                comp.append(u)
                for v in _np.flatnonzero(P[u] & ~seen):
                    seen[v] = True
                    stack.append(int(v))
            comp = sorted(comp)
            # column-major positions of the component's principal submatrix inside the block
            cols.extend(s + cj * n + ci for cj in comp for ci in comp)
            new_sizes.append(len(comp))
        s += n * n
    index = _np.array(cols, dtype=_np.int64)
    return A[:, index], c[index], new_sizes, index


def extract_constraints(A, c, sizes):
    """util/ExtractConstraintMatrices.m: per block the variables (rows of A) that touch it, their n x n
    matrices and the affine term."""
    out, s = [], 0
    for n in sizes:
        blk = A[:, s:s + n * n]
        variables = _np.flatnonzero(_np.any(blk != 0, axis=1))
        mats = blk[variables].reshape(variables.size, n, n).transpose(0, 2, 1)  # column-major vec -> matrix
        out.append(dict(n=n, variables=variables, mats=mats, affine=c[s:s + n * n].reshape(n, n).T))
        s += n * n
    return out


def reference_options(config):
    """The library defaults (ConexProgram.m:34-36 calls CONEX_SetDefaultOptions) with the overrides of
    interfaces/matlab/conex.m:49-55; `maximum_mu` is reset because the Python class's
    DefaultConfiguration() raises it."""
    config.maximum_mu = 1e4
    config.inv_sqrt_mu_max = 1000
    config.infeasibility_threshold = 1e3
    config.max_iterations = 25
    config.prepare_dual_variables = 1
    config.divergence_upper_bound = 1
    config.final_centering_steps = 5
    return config


def solve(A, b, c, K, new_program=None, blkdiag=None, errors=False):
    """Returns (x, y, info). `new_program(m)` builds the program object (default: conex_b200.Conex)."""
    for field in ("l", "q", "r", "f"):
        if _nontrivial(K, field):
            raise SedumiError("Cone not supported yet")  # conex.m:7-15 (K.f: EliminateFreeVars is not mirrored)
    sizes = _blocks(K)
    A_in = _np.asarray(A, dtype=_np.float64)
    b_in = _np.asarray(b, dtype=_np.float64).ravel()
    c_in = _np.asarray(c, dtype=_np.float64).ravel()
    if A_in.shape[1] != sum(n * n for n in sizes) or c_in.shape[0] != A_in.shape[1]:
        raise SedumiError("A, c and K.s do not match.")
    A1, b1, T = clean_linear(A_in, b_in)
    A1 = symmetrize(A1, sizes)
    c1 = symmetrize(c_in, sizes)[0]
    if blkdiag is None:
        blkdiag = len(sizes) > 1  # conex.m:20
    index = _np.arange(A1.shape[1])
    sizes_r = sizes
    T2 = _np.eye(A1.shape[0])
    if blkdiag:
        A1, c1, sizes_r, index = split_blocks(A1, c1, sizes)
        A1, b1, T2 = clean_linear(A1, b1)  # blkdiagPrg.m:33
    m = A1.shape[0]
    if new_program is None:
        from . import Conex
        new_program = Conex
    prog = new_program(m)
    cons = extract_constraints(A1, c1, sizes_r)
    if len(sizes_r) > 1:
        for con in cons:
            prog.AddSparseLinearMatrixInequality(_np.ascontiguousarray(con["mats"].transpose(1, 2, 0)), con["affine"],
                                                 [int(v) for v in con["variables"]])
    else:
        con = cons[0]
        n = con["n"]
        full = _np.zeros((n, n, m))
        full[:, :, con["variables"]] = con["mats"].transpose(1, 2, 0)
        prog.AddDenseLinearMatrixInequality(full, con["affine"])
    if hasattr(prog, "DefaultConfiguration"):
        sol = prog.Maximize(b1, reference_options(prog.DefaultConfiguration()))
    else:
        sol = prog.Maximize(b1)
    duals = prog.GetDualVariables()
    x_r = _np.concatenate([_np.asarray(X, dtype=_np.float64).ravel(order="F") for X in duals])
    x = _np.zeros(A_in.shape[1])
    x[index] = x_r                      # blkdiagPrg.RecoverPrimal
    y = T @ (T2 @ _np.asarray(sol.y, dtype=_np.float64).ravel())  # conex.m:72, blkdiagPrg.RecoverDual
    solved = bool(sol.status)
    info = dict(numerr=0, pinf=int(not solved), dinf=int(not solved), feasratio=1, blocks=list(sizes_r))
    if errors:
        gap = float(c_in @ x - b_in @ y)
        info["errors"] = [abs(gap), gap]
    return x, y, info
