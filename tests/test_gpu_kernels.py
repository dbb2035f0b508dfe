"""Kernel-level parity tests: every cxb_* entry point (include/conex_b200_device.h) against the CPU
oracle / numpy on identical seeded inputs. Integer/index outputs must match exactly; FP64 results
are compared at the tolerance BASELINE.json states for the Newton system (1e-10 relative,
reordered summation), tightened where the kernel allows.
"""
import numpy as np
import pytest

from harness import dptr, fmat, oracle, pack_matrices, random_dense_lmi, random_sym

pytestmark = pytest.mark.gpu


def rel_err(x, ref):
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.fixture(scope="module")
def dev():
    import devlib
    return devlib


def run_gemm(dev, ta, tb, M, N, K, alpha, beta, rng, lower=False, batch=1):
    import torch
    L = dev.product().lib
    As = [rng.standard_normal((K, M) if ta else (M, K)) for _ in range(batch)]
    Bs = [rng.standard_normal((N, K) if tb else (K, N)) for _ in range(batch)]
    Cs = [rng.standard_normal((M, N)) for _ in range(batch)]
    dA = torch.stack([dev.to_dev(a) for a in As]).contiguous()
    dB = torch.stack([dev.to_dev(b) for b in Bs]).contiguous()
    dC = torch.stack([dev.to_dev(c) for c in Cs]).contiguous()
    lda = As[0].shape[0]
    ldb = Bs[0].shape[0]
    rc = L.cxb_dgemm(None, int(ta), int(tb), M, N, K, alpha, dev.ptr(dA), lda, As[0].size, dev.ptr(dB),
                     ldb, Bs[0].size, beta, dev.ptr(dC), M, M * N, batch, int(lower))
    assert rc == 0
    torch.cuda.synchronize()
    out = dC.cpu().numpy()
    for i in range(batch):
        opA = As[i].T if ta else As[i]
        opB = Bs[i].T if tb else Bs[i]
        ref = alpha * opA @ opB + beta * Cs[i]
        got = out[i].T
        if lower:
            mask = np.tril(np.ones((M, N), dtype=bool))
            assert rel_err(got[mask], ref[mask]) < 1e-13
            assert np.array_equal(got[~mask], Cs[i][~mask])  # strictly-upper part untouched
        else:
            assert rel_err(got, ref) < 1e-13, (ta, tb, M, N, K)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(50, 50, 50), (64, 64, 16), (129, 257, 77), (33, 17, 5),
                                   (300, 200, 1000), (128, 128, 4), (1, 1, 1), (257, 131, 64)])
def test_dgemm_layouts_and_ragged_shapes(dev, ta, tb, shape):
    rng = np.random.default_rng(hash((ta, tb) + shape) % 2**32)
    M, N, K = shape
    run_gemm(dev, ta, tb, M, N, K, 1.0, 0.0, rng)
    run_gemm(dev, ta, tb, M, N, K, -0.5, 2.0, rng)


def test_dgemm_batched_and_lower(dev):
    rng = np.random.default_rng(7)
    run_gemm(dev, 0, 0, 50, 50, 50, 1.0, 0.0, rng, batch=5)
    run_gemm(dev, 1, 0, 100, 100, 2500, 1.0, 0.0, rng, lower=True)
    run_gemm(dev, 0, 1, 300, 300, 128, -1.0, 1.0, rng, lower=True)
    run_gemm(dev, 1, 0, 259, 258, 900, 1.0, 0.0, rng, lower=True)  # trapezoid, ragged tiles
    run_gemm(dev, 1, 0, 131, 130, 49, 1.0, 0.0, rng, lower=True)   # odd K -> 8-byte copies


@pytest.mark.parametrize("config", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_dgemm_every_tile_configuration(dev, config, ta, tb):
    """Every compiled tile configuration, split-K (fixed-order reduction) and the mirrored store."""
    import torch
    L = dev.product().lib
    rng = np.random.default_rng(100 * config + 2 * ta + tb)
    for (M, N, K, lower, splits, odd) in [(391, 263, 530, 0, 1, 0), (391, 391, 1100, 1, 3, 0),
                                         (200, 200, 2100, 1, 0, 0), (77, 130, 67, 0, 2, 1)]:
        A = rng.standard_normal((K, M) if ta else (M, K))
        B = rng.standard_normal((N, K) if tb else (K, N))
        C0 = rng.standard_normal((M, N))
        dA, dB, dC = dev.to_dev(A), dev.to_dev(B), dev.to_dev(C0)
        rc = L.cxb_dgemm_ex(None, config, splits, ta, tb, M, N, K, 0.5, dev.ptr(dA), A.shape[0], 0,
                            dev.ptr(dB), B.shape[0], 0, -1.0, dev.ptr(dC), M, 0, 1, lower, 0, 0)
        assert rc == 0
        ref = 0.5 * (A.T if ta else A) @ (B.T if tb else B) - C0
        got = dev.from_dev(dC)
        if lower:
            mask = np.tril(np.ones((M, N), dtype=bool))
            assert rel_err(got[mask], ref[mask]) < 1e-13
            assert np.array_equal(got[~mask], C0[~mask])
        else:
            assert rel_err(got, ref) < 1e-13
    # mirrored store: batched W * T_i with exactly symmetric output
    n, batch = 150, 3
    W = random_sym(rng, n)
    Ts = [W @ random_sym(rng, n) @ W for _ in range(batch)]   # symmetric products
    Winv_T = [np.linalg.solve(W, t) for t in Ts]               # so that W * (W^-1 T) = T symmetric
    dW = dev.to_dev(W)
    dT = torch.stack([dev.to_dev(x) for x in Winv_T]).contiguous()
    dC = dev.dzeros(batch, n, n)
    rc = L.cxb_dgemm_ex(None, config, 1, 0, 0, n, n, n, 1.0, dev.ptr(dW), n, 0, dev.ptr(dT), n, n * n, 0.0,
                        dev.ptr(dC), n, n * n, batch, 1, 1, 0)
    assert rc == 0
    torch.cuda.synchronize()
    out = dC.cpu().numpy()
    for i in range(batch):
        assert np.array_equal(out[i], out[i].T)
        assert rel_err(np.tril(out[i].T), np.tril(W @ Winv_T[i])) < 1e-12
    # row panel of a lower-triangular result: keep row + off >= col
    M, N, K, off = 130, 300, 257, 70
    A = rng.standard_normal((K, M))
    B = rng.standard_normal((K, N))
    C0 = rng.standard_normal((M, N))
    dA, dB, dC = dev.to_dev(A), dev.to_dev(B), dev.to_dev(C0)
    for splits in (1, 3):
        dC = dev.to_dev(C0)
        assert L.cxb_dgemm_ex(None, config, splits, 1, 0, M, N, K, 1.0, dev.ptr(dA), K, 0, dev.ptr(dB), K, 0,
                              0.0, dev.ptr(dC), M, 0, 1, 1, 0, off) == 0
        got = dev.from_dev(dC)
        mask = (np.arange(M)[:, None] + off) >= np.arange(N)[None, :]
        assert rel_err(got[mask], (A.T @ B)[mask]) < 1e-13
        assert np.array_equal(got[~mask], C0[~mask])
    # split-K is deterministic: two runs give identical bits
    M = N = 260
    K = 4096
    A = rng.standard_normal((K, M))
    B = rng.standard_normal((K, N))
    dA, dB = dev.to_dev(A), dev.to_dev(B)
    outs = []
    for _ in range(2):
        dC = dev.dzeros(N, M)
        assert L.cxb_dgemm_ex(None, config, 4, 1, 0, M, N, K, 1.0, dev.ptr(dA), K, 0, dev.ptr(dB), K, 0, 0.0,
                              dev.ptr(dC), M, 0, 1, 1, 0, 0) == 0
        outs.append(dev.from_dev(dC))
    assert np.array_equal(outs[0], outs[1])


def test_dgemm_zero_k_and_empty(dev):
    import torch
    L = dev.product().lib
    C0 = np.arange(12.0).reshape(3, 4)
    dC = dev.to_dev(C0)
    dA = dev.dzeros(4)
    assert L.cxb_dgemm(None, 0, 0, 3, 4, 0, 1.0, dev.ptr(dA), 3, 0, dev.ptr(dA), 1, 0, 2.0, dev.ptr(dC), 3,
                       0, 1, 0) == 0
    assert np.allclose(dev.from_dev(dC), 2.0 * C0)
    assert L.cxb_dgemm(None, 0, 0, 0, 4, 5, 1.0, dev.ptr(dA), 1, 0, dev.ptr(dA), 5, 0, 0.0, dev.ptr(dC), 1,
                       0, 1, 0) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("n,m,seed", [(10, 7, 1), (50, 100, 2), (33, 20, 3), (4, 1, 4), (16, 129, 5), (150, 9, 6),
                                      (64, 5, 7), (65, 4, 8)])
def test_schur_dense_lmi_matches_oracle(dev, n, m, seed):
    import ctypes as C
    """K1+K2: H_ij = tr(A_i W A_j W), AW, AQc, <w,c>, <c,Qc> within 1e-10 relative of the oracle."""
    import torch
    L = dev.product().lib
    O = oracle().lib
    mats, _ = random_dense_lmi(n, m, seed)
    rng = np.random.default_rng(seed + 100)
    Cm = random_sym(rng, n) + np.eye(n)
    R = rng.standard_normal((n, n))
    W = R @ R.T / n + 0.1 * np.eye(n)
    A = pack_matrices(mats)
    G = np.zeros((m, m), order="F")
    AW = np.zeros(m)
    AQc = np.zeros(m)
    sc = np.zeros(2)
    O.ORACLE_SchurDenseLMI(n, m, dptr(A), dptr(fmat(Cm)), dptr(fmat(W)), 0, dptr(G), dptr(AW), dptr(AQc),
                           dptr(sc))
    Aall = np.concatenate([A.ravel(), np.asfortranarray(Cm).ravel(order="F")])
    dA = torch.from_numpy(Aall).cuda()
    dW = dev.to_dev(W)
    dB = dev.dzeros((m + 2) * n * n)
    for panel in (1, 3, m + 1):
        dT = dev.dzeros(panel * n * n)
        ldh = (m + 3) & ~1
        dH = dev.dzeros(ldh * (m + 2))
        assert L.cxb_schur_dense_lmi(None, n, m, dev.ptr(dA), dev.ptr(dW), dev.ptr(dB), dev.ptr(dT), panel,
                                     dev.ptr(dH), ldh) == 0
        Haug = dev.from_dev(dH, ldh, m + 2)
        H = np.tril(Haug[:m, :m])
        scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
        assert (np.abs(H - np.tril(G)) / scale).max() < 1e-10
        assert rel_err(Haug[m, :m], AQc) < 1e-10
        assert rel_err(Haug[m + 1, :m], AW) < 1e-10
        assert abs(Haug[m + 1, m] - sc[0]) <= 1e-10 * abs(sc[0])
        assert abs(Haug[m, m] - sc[1]) <= 1e-10 * abs(sc[1])
    # symmetric form: packed L^T A_i L (W = L L^T) and a SYRK-shaped Gram — same Newton system
    vp = C.c_void_p
    L.cxb_packed_symmetric_size.restype = C.c_size_t
    L.cxb_packed_symmetric_size.argtypes = [C.c_int]
    L.cxb_schur_dense_lmi_sym.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_long]
    kp = L.cxb_packed_symmetric_size(n)
    dX = dev.dzeros((m + 2) * kp)
    dL = dev.dzeros(n * n)
    info = dev.izeros(2)
    for panel in (1, 3, m + 1):
        dT = dev.dzeros(panel * n * n)
        ldh = (m + 3) & ~1
        dH = dev.dzeros(ldh * (m + 2))
        assert L.cxb_schur_dense_lmi_sym(None, n, m, dev.ptr(dA), dev.ptr(dW), dev.ptr(dX), dev.ptr(dT), panel,
                                         dev.ptr(dL), dev.ptr(info), dev.ptr(dH), ldh) == 0
        assert int(info.cpu()[0]) == 0
        Haug = dev.from_dev(dH, ldh, m + 2)
        H = np.tril(Haug[:m, :m])
        scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
        assert (np.abs(H - np.tril(G)) / scale).max() < 1e-10
        assert rel_err(Haug[m, :m], AQc) < 1e-10
        assert rel_err(Haug[m + 1, :m], AW) < 1e-10
        assert abs(Haug[m + 1, m] - sc[0]) <= 1e-10 * abs(sc[0])
        assert abs(Haug[m, m] - sc[1]) <= 1e-10 * abs(sc[1])
    # streamed variant: one row panel of scaled matrices at a time, same Newton system
    for panel in (1, 4, m + 1):
        dT = dev.dzeros(panel * n * n)
        dBp = dev.dzeros((panel + 1) * n * n)
        ldh = (m + 3) & ~1
        dH = dev.dzeros(ldh * (m + 2))
        assert L.cxb_schur_dense_lmi_streamed(None, n, m, dev.ptr(dA), dev.ptr(dW), dev.ptr(dBp), dev.ptr(dT),
                                              panel, dev.ptr(dH), ldh) == 0
        Haug = dev.from_dev(dH, ldh, m + 2)
        scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
        assert (np.abs(np.tril(Haug[:m, :m]) - np.tril(G)) / scale).max() < 1e-10
        assert rel_err(Haug[m, :m], AQc) < 1e-10
        assert rel_err(Haug[m + 1, :m], AW) < 1e-10
        assert abs(Haug[m + 1, m] - sc[0]) <= 1e-10 * abs(sc[0])
        assert abs(Haug[m, m] - sc[1]) <= 1e-10 * abs(sc[1])


@pytest.mark.parametrize("m", [1, 5, 31, 33, 100, 128, 129, 300, 511, 513, 700, 1000, 2500])
def test_potrf_and_potrs(dev, m):
    """K3/K5 against numpy's Cholesky and solve; lower triangle only is read."""
    import torch
    L = dev.product().lib
    rng = np.random.default_rng(m)
    R = rng.standard_normal((m, m + 5))
    H = R @ R.T + m * np.eye(m)
    ld = m + 2
    Hp = np.zeros((ld, m))
    Hp[:m, :] = np.tril(H) + np.triu(np.full((m, m), np.nan), 1)  # upper part must never be read
    dH = dev.to_dev(Hp)
    info = dev.izeros(1)
    assert L.cxb_potrf_lower(None, m, dev.ptr(dH), ld, None, dev.ptr(info)) == 0
    torch.cuda.synchronize()
    assert int(info.cpu()[0]) == 0
    Ld = np.tril(dev.from_dev(dH, ld, m)[:m, :])
    Lref = np.linalg.cholesky(H)
    assert rel_err(Ld, Lref) < 1e-12
    for nrhs in (1, 3):
        B = rng.standard_normal((m, nrhs))
        dX = dev.to_dev(B)
        assert L.cxb_potrs_lower(None, m, dev.ptr(dH), ld, dev.ptr(dX), m, nrhs) == 0
        X = dev.from_dev(dX)
        assert rel_err(H @ X, B) < 1e-10


@pytest.mark.parametrize("m,block", [(1, 4), (97, 8), (700, 128), (1500, 256), (2200, 512), (1300, 100)])
def test_potrf_by_panels_single_rank(dev, m, block):
    """The multi-GPU Cholesky driver (host/distributed_cholesky.cc) with a world of one: the same
    schedule, panel kernels and per-block-column trailing updates, broadcasts being no-ops. The
    multi-rank runs are in tests/test_multi_gpu.py."""
    import ctypes as C

    import torch
    L = dev.product().lib
    rng = np.random.default_rng(m + block)
    R = rng.standard_normal((m, m + 5))
    H = R @ R.T / m + np.eye(m)
    ld = m + 2
    Hp = np.zeros((ld, m))
    Hp[:m, :] = np.tril(H) + np.triu(np.full((m, m), np.nan), 1)  # upper part must never be read
    dH = dev.to_dev(Hp)
    info = C.c_int(-1)
    assert L.CONEXB200_DistributedPotrf(m, dev.ptr(dH), ld, block, C.byref(info)) == 0
    torch.cuda.synchronize()
    assert info.value == 0
    Ld = np.tril(dev.from_dev(dH, ld, m)[:m, :])
    assert rel_err(Ld, np.linalg.cholesky(H)) < 1e-12
    # non-positive pivot in a later panel
    if m >= 97:
        H[m - 3, m - 3] = -1.0
        Hp[:m, :] = np.tril(H)
        dH = dev.to_dev(Hp)
        assert L.CONEXB200_DistributedPotrf(m, dev.ptr(dH), ld, block, C.byref(info)) == 0
        assert info.value != 0


def test_potrf_reports_non_positive_pivot(dev):
    import torch
    L = dev.product().lib
    m = 200
    rng = np.random.default_rng(0)
    R = rng.standard_normal((m, m))
    H = R @ R.T + np.eye(m)
    H[150, 150] = -1.0  # not positive definite: the failing pivot is at or before column 150
    dH = dev.to_dev(np.tril(H))
    info = dev.izeros(1)
    assert L.cxb_potrf_lower(None, m, dev.ptr(dH), m, None, dev.ptr(info)) == 0
    torch.cuda.synchronize()
    assert 1 <= int(info.cpu()[0]) <= 151


@pytest.mark.parametrize("nn,cols", [(2500, 101), (9, 3), (10000, 1030), (49, 7)])
def test_gemv_n(dev, nn, cols):
    L = dev.product().lib
    rng = np.random.default_rng(nn)
    A = rng.standard_normal((nn, cols))
    x = rng.standard_normal(cols)
    dA, dx, dy = dev.to_dev(A), dev.to_dev(x), dev.dzeros(nn)
    assert L.cxb_gemv_n(None, nn, cols, dev.ptr(dA), dev.ptr(dx), dev.ptr(dy)) == 0
    assert rel_err(dev.from_dev(dy), A @ x) < 1e-13


def lanczos_inputs(n, seed):
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((n, n))
    W = R @ R.T / n + 0.05 * np.eye(n)
    W = 0.5 * (W + W.T)
    S = random_sym(rng, n)
    return W, S, W @ S


@pytest.mark.parametrize("n,seed", [(4, 1), (25, 2), (64, 3), (7, 5), (2, 6), (3, 7)])
def test_two_sided_lanczos_matches_oracle(dev, n, seed):
    """K7: alpha/beta of the Jacobi matrix and the Ritz extremes vs. the oracle's recurrence."""
    import torch
    L = dev.product().lib
    O = oracle().lib
    W, S, WS = lanczos_inputs(n, seed)
    num_iter = max(n // 2, 1)
    idx = int(np.argmax(np.diag(WS)))
    r = S[:, idx].copy()
    ev = np.zeros(num_iter)
    k = O.ORACLE_ApproximateEigenvalues(n, dptr(fmat(WS)), dptr(fmat(W)), dptr(r), num_iter, dptr(ev))
    dWS, dW, dS = dev.to_dev(WS), dev.to_dev(W), dev.to_dev(S)
    red = dev.dzeros(8)
    assert L.cxb_ws_reductions(None, n, dev.ptr(dWS), dev.ptr(red)) == 0
    r4 = dev.from_dev(red)
    assert int(r4[2]) == idx
    assert abs(r4[0] - np.trace(WS)) < 1e-12 * max(1, abs(np.trace(WS)))
    assert abs(r4[1] - np.trace(WS @ WS)) < 1e-12 * abs(np.trace(WS @ WS))
    alpha, beta, cnt = dev.dzeros(num_iter + 1), dev.dzeros(num_iter + 1), dev.izeros(2)
    work = dev.dzeros(L.cxb_lanczos_worksize(n))
    idx_dev = red[2:3]
    assert L.cxb_lanczos_two_sided(None, n, dev.ptr(dWS), dev.ptr(dW), dev.ptr(dS), dev.ptr(idx_dev),
                                   num_iter, dev.ptr(alpha), dev.ptr(beta), dev.ptr(cnt),
                                   dev.ptr(work)) == 0
    torch.cuda.synchronize()
    c = int(cnt.cpu()[0])
    assert c + 1 == k
    a = alpha.cpu().numpy()[: c + 1]
    b = beta.cpu().numpy()[:c]
    T = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
    ritz = np.linalg.eigvalsh(T)
    tol = 1e-9 * max(1.0, np.abs(ev[:k]).max())
    assert abs(ritz[0] - ev[0]) < tol and abs(ritz[-1] - ev[k - 1]) < tol


def test_two_sided_lanczos_long_run_properties(dev):
    """n = 200, 100 steps. Without re-orthogonalisation the recurrence amplifies rounding (the
    oracle run with OpenBLAS and with plain loops already breaks down at different steps, see
    tests/test_oracle_golden.py), so a long run is pinned by (i) the first coefficients against a
    numpy evaluation of the same recurrence and (ii) the reference's own acceptance criterion for
    the Ritz extremes: within 1e-2 relative of the true spectrum bounds
    (conex/test/approximate_eigenvalues.cc:87-113)."""
    import torch
    L = dev.product().lib
    n, num_iter = 200, 100
    W, S, WS = lanczos_inputs(n, 4)
    idx = int(np.argmax(np.diag(WS)))
    r = S[:, idx].copy()
    alpha, beta, cnt = dev.dzeros(num_iter + 1), dev.dzeros(num_iter + 1), dev.izeros(2)
    work = dev.dzeros(L.cxb_lanczos_worksize(n))
    assert L.cxb_lanczos_two_sided(None, n, dev.ptr(dev.to_dev(WS)), dev.ptr(dev.to_dev(W)),
                                   dev.ptr(dev.to_dev(r)), None, num_iter, dev.ptr(alpha), dev.ptr(beta),
                                   dev.ptr(cnt), dev.ptr(work)) == 0
    torch.cuda.synchronize()
    c = int(cnt.cpu()[0])
    a, b = alpha.cpu().numpy()[: c + 1], beta.cpu().numpy()[:c]
    # numpy evaluation of approximate_eigenvalues.cc:178-239 for the first steps
    v1 = r.copy()
    v0 = W @ r
    s = np.sqrt(v0 @ v1)
    v0, v1 = v0 / s, v1 / s
    u0, u1 = WS @ v0, WS.T @ v1
    al = [v0 @ u1]
    be = []
    u0, u1 = u0 - al[0] * v0, u1 - al[0] * v1
    for j in range(1, 8):
        bj = np.sqrt(u0 @ u1)
        be.append(bj)
        p0, p1 = v0, v1
        v0, v1 = u0 / bj, u1 / bj
        u0, u1 = WS @ v0, WS.T @ v1
        al.append(v0 @ u1)
        u0, u1 = u0 - al[j] * v0 - bj * p0, u1 - al[j] * v1 - bj * p1
    assert c >= 8
    assert np.abs(a[:8] - np.array(al)).max() < 1e-9 * np.abs(al).max()
    assert np.abs(b[:7] - np.array(be)).max() < 1e-9 * np.abs(be).max()
    ritz = np.linalg.eigvalsh(np.diag(a) + np.diag(b, 1) + np.diag(b, -1))
    true = np.sort(np.linalg.eigvals(WS).real)
    assert abs(ritz[-1] / true[-1] - 1) < 1e-2 and abs(ritz[0] / true[0] - 1) < 1e-2


def test_lanczos_golden_4x4(dev):
    """Reference known answer (conex/test/approximate_eigenvalues.cc:16-42): n steps on the 4x4
    matrix from r0 = (1,2,0,4) reproduce the exact spectrum of W*A to 1e-12."""
    import torch
    L = dev.product().lib
    A = np.array([[3, 1, 0, 1], [1, 3, 1, 0], [0, 1, 4, 1], [1, 0, 1, 5]], float)
    A /= np.trace(A)
    rng = np.random.default_rng(0)
    R = rng.uniform(-1, 1, (4, 4))
    W = R @ R.T
    WS = W @ A
    r0 = np.array([1.0, 2, 0, 4])
    alpha, beta, cnt = dev.dzeros(5), dev.dzeros(5), dev.izeros(2)
    work = dev.dzeros(L.cxb_lanczos_worksize(4))
    assert L.cxb_lanczos_two_sided(None, 4, dev.ptr(dev.to_dev(WS)), dev.ptr(dev.to_dev(W)),
                                   dev.ptr(dev.to_dev(r0)), None, 4, dev.ptr(alpha), dev.ptr(beta),
                                   dev.ptr(cnt), dev.ptr(work)) == 0
    torch.cuda.synchronize()
    c = int(cnt.cpu()[0])
    assert c == 3
    a, b = alpha.cpu().numpy()[:4], beta.cpu().numpy()[:3]
    ritz = np.linalg.eigvalsh(np.diag(a) + np.diag(b, 1) + np.diag(b, -1))
    exact = np.sort(np.linalg.eigvals(WS).real)
    assert np.abs(ritz - exact).max() < 1e-12


@pytest.mark.parametrize("n", [4, 5, 31, 50, 129, 300])
def test_pade_expm_matches_oracle(dev, n):
    """K8 (Padé + pivoted LU): against the oracle's restatement of exponential_map_pade.cc."""
    import torch
    L = dev.product().lib
    O = oracle().lib
    rng = np.random.default_rng(n)
    if n == 4:
        X = np.array([[3, 1, 0, 1], [1, 3, 1, 0], [0, 1, 4, 1], [1, 0, 1, 5]], float)
        X /= np.trace(X)
    else:
        X = rng.standard_normal((n, n))
        X *= 1.2 / np.linalg.norm(X, 2)
    ref = np.zeros((n, n), order="F")
    O.ORACLE_PadeExpm(n, dptr(fmat(X)), dptr(ref))
    dX, dout = dev.to_dev(X), dev.dzeros(n * n)
    work, iwork, info = dev.dzeros(4 * n * n), dev.izeros(2 * n + 2), dev.izeros(1)
    assert L.cxb_pade_expm(None, n, dev.ptr(dX), dev.ptr(dout), dev.ptr(work), dev.ptr(iwork),
                           dev.ptr(info)) == 0
    got = dev.from_dev(dout, n, n)
    assert int(info.cpu()[0]) == 0
    assert rel_err(got, ref) < 1e-11
    if n == 4:
        import scipy.linalg
        assert np.abs(got - scipy.linalg.expm(X)).max() < 1e-7  # exponential_map_pade_test.cc:16-37


@pytest.mark.parametrize("n,nrhs", [(3, 2), (40, 40), (100, 7), (257, 257)])
def test_lu_solve_with_pivoting(dev, n, nrhs):
    L = dev.product().lib
    rng = np.random.default_rng(n + nrhs)
    A = rng.standard_normal((n, n))
    A[0, 0] = 0.0  # forces a row interchange in the first column
    B = rng.standard_normal((n, nrhs))
    dA, dB = dev.to_dev(A), dev.to_dev(B)
    ipiv, info = dev.izeros(2 * n), dev.izeros(1)
    assert L.cxb_lu_solve(None, n, dev.ptr(dA), n, nrhs, dev.ptr(dB), n, dev.ptr(ipiv), dev.ptr(info)) == 0
    X = dev.from_dev(dB)
    assert int(info.cpu()[0]) == 0
    assert rel_err(A @ X, B) < 1e-9
    assert rel_err(X, np.linalg.solve(A, B)) < 1e-8


@pytest.mark.parametrize("n,m,seed", [(6, 4, 1), (50, 20, 2), (100, 10, 3)])
def test_geodesic_update_matches_oracle_psd_step(dev, n, m, seed):
    """K6+K7+K8 chained exactly as PrepareStep/TakeStep do (psd_constraint.cc:45-90)."""
    import torch
    L = dev.product().lib
    O = oracle().lib
    mats, Cm = random_dense_lmi(n, m, seed)
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((n, n))
    W = R @ R.T / n + 0.5 * np.eye(n)
    W = 0.5 * (W + W.T)
    y = 0.1 * rng.standard_normal(m)
    k = 0.7
    A = pack_matrices(mats)
    Wref = fmat(W)
    out4 = np.zeros(4)
    WSref = np.zeros((n, n), order="F")
    O.ORACLE_PsdStep(n, m, dptr(A), dptr(fmat(Cm)), dptr(Wref), dptr(y), k, 1.0, 0, 1, dptr(out4),
                     dptr(WSref))
    S = sum(y[i] * mats[i] for i in range(m)) - k * Cm
    Aall = np.concatenate([A.ravel(), np.asfortranarray(Cm).ravel(order="F")])
    dA = torch.from_numpy(Aall).cuda()
    coef = dev.to_dev(np.concatenate([y, [-k]]))
    dS = dev.dzeros(n * n)
    assert L.cxb_gemv_n(None, n * n, m + 1, dev.ptr(dA), dev.ptr(coef), dev.ptr(dS)) == 0
    assert rel_err(dev.from_dev(dS, n, n), S) < 1e-13
    dW = dev.to_dev(W)
    dWS = dev.dzeros(n * n)
    assert L.cxb_dgemm(None, 0, 0, n, n, n, 1.0, dev.ptr(dW), n, 0, dev.ptr(dS), n, 0, 0.0, dev.ptr(dWS), n,
                       0, 1, 0) == 0
    assert rel_err(dev.from_dev(dWS, n, n), WSref) < 1e-12
    work, iwork, info = dev.dzeros(4 * n * n), dev.izeros(2 * n + 2), dev.izeros(1)
    assert L.cxb_geodesic_update(None, n, dev.ptr(dW), dev.ptr(dWS), 1.0, out4[2], dev.ptr(work),
                                 dev.ptr(iwork), dev.ptr(info)) == 0
    Wnew = dev.from_dev(dW, n, n)
    assert np.array_equal(Wnew, Wnew.T)
    assert rel_err(Wnew, np.asarray(Wref)) < 1e-10


def test_small_vector_helpers(dev):
    import torch
    L = dev.product().lib
    rng = np.random.default_rng(3)
    n = 1000
    x, y, z = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    dx, dy, dz, out = dev.to_dev(x), dev.to_dev(y), dev.to_dev(z), dev.dzeros(1)
    assert L.cxb_dot(None, n, dev.ptr(dx), dev.ptr(dy), dev.ptr(out)) == 0
    assert abs(dev.from_dev(out)[0] - x @ y) < 1e-11
    assert L.cxb_axpbypcz(None, n, 2.0, dev.ptr(dx), -1.0, dev.ptr(dy), 0.5, dev.ptr(dz)) == 0
    assert np.allclose(dev.from_dev(dy), 2 * x - y + 0.5 * z, rtol=0, atol=1e-14)
    idx = rng.permutation(n)[:100].astype(np.int32)
    didx = torch.from_numpy(idx).cuda()
    dst = dev.dzeros(n)
    src = dev.to_dev(rng.standard_normal(100))
    assert L.cxb_scatter_add_vec(None, 100, dev.ptr(src), dev.ptr(didx), dev.ptr(dst)) == 0
    ref = np.zeros(n)
    ref[idx] += dev.from_dev(src)
    assert np.array_equal(dev.from_dev(dst), ref)
    g = dev.dzeros(100)
    assert L.cxb_gather_vec(None, 100, dev.ptr(dst), dev.ptr(didx), dev.ptr(g)) == 0
    assert np.array_equal(dev.from_dev(g), ref[idx])
    # scatter-add of a lower triangle through a clique index list
    mc = 40
    G = np.tril(rng.standard_normal((mc, mc)))
    cl = rng.permutation(60)[:mc].astype(np.int32)
    dH = dev.dzeros(60 * 60)
    assert L.cxb_scatter_add_lower(None, mc, dev.ptr(dev.to_dev(G)), mc, dev.ptr(torch.from_numpy(cl).cuda()),
                                   dev.ptr(dH), 60) == 0
    Href = np.zeros((60, 60))
    for a in range(mc):
        for b in range(a + 1):
            Href[max(cl[a], cl[b]), min(cl[a], cl[b])] += G[a, b]
    assert np.array_equal(dev.from_dev(dH, 60, 60), Href)


@pytest.mark.parametrize("n", [1, 17, 64, 100, 193])
def test_triangular_gemm_and_symmetric_packing(dev, n):
    """The structured GEMMs behind the symmetric Schur form: A L with L lower triangular (zero k-tiles
    skipped) and L^T T stored as packed 64 x 64 lower tiles with sqrt(2)-scaled off-diagonal tiles —
    checked through cxb_schur_dense_lmi_sym with m = 1 (the packed buffer holds L^T A L, L^T C L, I)
    and cxb_pack_symmetric: <pack(X), pack(Y)> == tr(X Y) for symmetric X, Y."""
    import ctypes as C
    L = dev.product().lib
    vp = C.c_void_p
    L.cxb_packed_symmetric_size.restype = C.c_size_t
    L.cxb_packed_symmetric_size.argtypes = [C.c_int]
    L.cxb_pack_symmetric.argtypes = [vp, C.c_int, vp, vp]
    L.cxb_schur_dense_lmi_sym.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_long]
    rng = np.random.default_rng(n)
    X, Y = random_sym(rng, n), random_sym(rng, n)
    kp = L.cxb_packed_symmetric_size(n)
    T = (n + 63) // 64
    assert kp == T * (T + 1) // 2 * 4096
    px, py = dev.dzeros(kp), dev.dzeros(kp)
    assert L.cxb_pack_symmetric(None, n, dev.ptr(dev.to_dev(X)), dev.ptr(px)) == 0
    assert L.cxb_pack_symmetric(None, n, dev.ptr(dev.to_dev(Y)), dev.ptr(py)) == 0
    got = float((px * py).sum().cpu())
    assert abs(got - np.trace(X @ Y)) <= 1e-12 * max(1.0, abs(np.trace(X @ Y)), np.abs(X).sum())
    # scaled matrices: compare the packed L^T A L with numpy, tile by tile
    R = rng.standard_normal((n, n))
    W = R @ R.T / n + 0.5 * np.eye(n)
    Lw = np.linalg.cholesky(W)
    import torch
    Aall = np.concatenate([np.asfortranarray(X).ravel(order="F"), np.asfortranarray(Y).ravel(order="F")])
    dA = torch.from_numpy(Aall).cuda()
    dX, dT, dL, info = dev.dzeros(3 * kp), dev.dzeros(n * n), dev.dzeros(n * n), dev.izeros(2)
    dH = dev.dzeros(4 * 3)
    assert L.cxb_schur_dense_lmi_sym(None, n, 1, dev.ptr(dA), dev.ptr(dev.to_dev(W)), dev.ptr(dX), dev.ptr(dT), 1,
                                     dev.ptr(dL), dev.ptr(info), dev.ptr(dH), 4) == 0
    assert np.abs(np.tril(dev.from_dev(dL, n, n)) - Lw).max() < 1e-12 * np.abs(Lw).max()
    ref = dev.dzeros(kp)
    S = Lw.T @ X @ Lw
    assert L.cxb_pack_symmetric(None, n, dev.ptr(dev.to_dev(S)), dev.ptr(ref)) == 0
    gotS = dX[:kp].cpu().numpy()
    refS = ref.cpu().numpy()
    # diagonal tiles of the GEMM output carry both triangles of S; the reference packs a symmetric S
    assert np.abs(gotS - refS).max() < 1e-11 * max(1.0, np.abs(S).max())
    H = dev.from_dev(dH, 4, 3)
    assert abs(H[0, 0] - np.trace(X @ W @ X @ W)) <= 1e-10 * abs(np.trace(X @ W @ X @ W))
    assert abs(H[2, 0] - np.trace(X @ W)) <= 1e-10 * max(1.0, np.abs(X @ W).sum())


@pytest.mark.parametrize("M,N,K,batch,tri", [(64, 64, 16, 1, 0), (128, 192, 64, 2, 0), (200, 200, 208, 3, 1), (2000, 2000, 2000, 2, 1),
                                             (1000, 130, 48, 1, 0)])
def test_dgemm_with_tma_bulk_copies_is_bit_identical(dev, M, N, K, batch, tri):
    """The A/B arm of the DMMA GEMM whose operand tiles are staged by cp.async.bulk (TMA engine, mbarrier completion)
    into the same padded layout: same fragment loads and k order, so the result is bit-identical to gemm.cu's."""
    rng = np.random.Generator(np.random.PCG64(M + N + K))
    A = rng.uniform(-1, 1, size=(batch, K, M))      # batch of column-major M x K blocks
    B = rng.uniform(-1, 1, size=(batch, N, K))      # batch of column-major K x N blocks
    if tri:
        for b in range(batch):
            B[b] = np.triu(B[b])                     # B(k, c) = 0 for k < c  (stored [c][k])
    import torch
    L = dev.product().lib
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    out = []
    for bulk in (0, 1):
        dC = torch.zeros((batch, N, M), dtype=torch.float64, device="cuda")
        if bulk:
            rc = L.cxb_dgemm_bulk(None, M, N, K, dev.ptr(dA), M, M * K, dev.ptr(dB), K, K * N, dev.ptr(dC), M, M * N, batch, tri)
        else:
            rc = L.cxb_dgemm_ex(None, 2, 1, 0, 0, M, N, K, 1.0, dev.ptr(dA), M, M * K, dev.ptr(dB), K, K * N, 0.0, dev.ptr(dC),
                                M, M * N, batch, 0, 0, 0)
        assert rc == 0
        torch.cuda.synchronize()
        out.append(dC.cpu().numpy())
    ref = np.einsum("bkm,bnk->bnm", A, B)
    assert np.abs(out[0] - ref).max() <= 1e-12 * K
    assert np.array_equal(out[0], out[1])

