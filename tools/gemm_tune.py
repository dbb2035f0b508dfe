#!/usr/bin/env python
"""Times every DGEMM tile configuration of conex_b200/csrc/device/gemm.cu on the shapes of the
Newton step (run on a B200 through gpurun). Prints one line per (shape, config) with TFLOP/s and
the cuBLAS (torch.matmul) time of the same product for orientation. Not part of the product."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devlib  # noqa: E402

dev = devlib.product()
L = dev.lib
vp = C.c_void_p


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def stream():
    return vp(torch.cuda.current_stream().cuda_stream)


def gemm(cfg, splits, ta, tb, M, N, K, A, lda, sA, B, ldb, sB, Cm, ldc, sC, batch, lower, mirror, beta=0.0):
    rc = L.cxb_dgemm_ex(stream(), cfg, splits, ta, tb, M, N, K, 1.0, vp(A.data_ptr()), lda, sA,
                        vp(B.data_ptr()), ldb, sB, beta, vp(Cm.data_ptr()), ldc, sC, batch, lower, mirror, 0)
    assert rc == 0, rc


def main():
    torch.manual_seed(0)
    f64 = dict(dtype=torch.float64, device="cuda")
    configs = [1, 2, 3, 5]
    for n, batch in ((2000, 16), (1000, 64), (500, 128)):
        nn = n * n
        A = torch.randn(batch, n, n, **f64)
        W = torch.randn(n, n, **f64)
        T = torch.empty(batch, n, n, **f64)
        ref = timed(lambda: torch.matmul(W.T, A, out=T))  # same flops through cuBLAS
        print(f"[K1 NN batched n={n} batch={batch}] cuBLAS {2.0 * n ** 3 * batch / ref / 1e9:.2f} TF/s")
        for cfg in configs:
            t = timed(lambda: gemm(cfg, 1, 0, 0, n, n, n, A, n, nn, W, n, 0, T, n, nn, batch, 0, 0))
            print(f"  cfg {cfg} full   : {t:8.3f} ms  {2.0 * n ** 3 * batch / t / 1e9:6.2f} TF/s")
            t = timed(lambda: gemm(cfg, 1, 0, 0, n, n, n, W, n, 0, A, n, nn, T, n, nn, batch, 1, 1))
            print(f"  cfg {cfg} lower+mirror: {t:8.3f} ms  {2.0 * n ** 3 * batch / t / 1e9:6.2f} TF/s (dense-equivalent)")
        del A, T
    # Gram: H = B^T A, K = n^2
    for n, m in ((1000, 2000), (1000, 1000), (700, 4000)):
        nn = n * n
        Bm = torch.randn(m, nn, **f64)
        Am = torch.randn(m, nn, **f64)
        H = torch.empty(m, m, **f64)
        ref = timed(lambda: torch.matmul(Am, Bm.T, out=H), reps=2)
        print(f"[K2 Gram TN lower m={m} K={nn}] cuBLAS full square {2.0 * m * m * nn / ref / 1e9:.2f} TF/s ({ref:.1f} ms)")
        for cfg in configs:
            for splits in (1, 0, 2, 3):
                t = timed(lambda: gemm(cfg, splits, 1, 0, m, m, nn, Bm, nn, 0, Am, nn, 0, H, m, 0, 1, 1, 0), reps=2)
                print(f"  cfg {cfg} splits {splits}: {t:8.3f} ms  {1.0 * m * (m + 1) * nn / t / 1e9:6.2f} TF/s (lower flops)")
        del Bm, Am, H
    # Cholesky trailing update: C -= L21 L21^T, K = 128
    for m in (1872, 4000, 16000):
        Lp = torch.randn(128, m, **f64)
        Cm = torch.randn(m, m, **f64)
        for cfg in configs:
            t = timed(lambda: gemm(cfg, 1, 0, 1, m, m, 128, Lp, m, 0, Lp, m, 0, Cm, m, 0, 1, 1, 0, beta=1.0))
            print(f"[K3 SYRK NT lower m={m} K=128] cfg {cfg}: {t:8.3f} ms  {1.0 * m * (m + 1) * 128 / t / 1e9:6.2f} TF/s")
        del Lp, Cm


if __name__ == "__main__":
    main()
