#!/usr/bin/env python
"""Round 2: the wide-warp-tile DGEMM configurations (6, 7, 8 of gemm.cu) against the default (2) on the two shapes of
the C2 assembly. One JSON line per (shape, config). Not part of the product."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_tune import gemm, timed  # noqa: E402


def main():
    torch.manual_seed(0)
    f64 = dict(dtype=torch.float64, device="cuda")
    configs = [2, 6, 7, 8]
    n, batch = 2000, 33
    nn = n * n
    A = torch.randn(batch, n, n, **f64)
    W = torch.randn(n, n, **f64)
    T = torch.empty(batch, n, n, **f64)
    ref = timed(lambda: torch.matmul(W.T, A, out=T), reps=5)
    print(json.dumps({"shape": f"K1 NN batched n={n} batch={batch}", "cublas_TFLOPs": 2.0 * n ** 3 * batch / ref / 1e9}))
    base = None
    for cfg in configs:
        t = timed(lambda: gemm(cfg, 1, 0, 0, n, n, n, A, n, nn, W, n, 0, T, n, nn, batch, 0, 0), reps=5)
        out = T.clone()
        base = out if base is None else base
        print(json.dumps({"shape": "K1 NN", "config": cfg, "ms": t, "TFLOPs": 2.0 * n ** 3 * batch / t / 1e9,
                          "identical_to_config_2": bool(torch.equal(out, base))}))
    del A, T, base, out
    m, K = 2000, 1000000
    Bm = torch.randn(m, K, **f64)
    H = torch.empty(m, m, **f64)
    ref = timed(lambda: torch.matmul(Bm, Bm.T, out=H), reps=2)
    print(json.dumps({"shape": f"K2 Gram TN lower m={m} K={K}", "cublas_full_square_TFLOPs": 2.0 * m * m * K / ref / 1e9}))
    base = None
    for cfg in configs:
        for splits in (0, 1):
            t = timed(lambda: gemm(cfg, splits, 1, 0, m, m, K, Bm, K, 0, Bm, K, 0, H, m, 0, 1, 1, 0), reps=2)
            out = torch.tril(H).clone()
            base = out if base is None else base
            print(json.dumps({"shape": "K2 Gram lower", "config": cfg, "splits": splits, "ms": t,
                              "TFLOPs_lower": 1.0 * m * (m + 1) * K / t / 1e9,
                              "max_rel_diff_to_first": float(((out - base).abs().max() / base.abs().max()).item())}))


if __name__ == "__main__":
    main()
