#!/usr/bin/env python
"""A/B of the two operand-staging schemes of the DMMA GEMM on K1's first product, T_i = A_i L (n = 2000, a panel of 33
matrices, L lower triangular: tri = 1) and on the plain product (tri = 0): per-thread cp.async (LDGSTS, gemm.cu) against
the TMA engine's bulk copies (cp.async.bulk + mbarrier, gemm_bulk.cu). CUDA events, best / median of 10, JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import conex_b200.binding as dev  # noqa: E402

L = dev.product().lib
n, batch = 2000, 33
A = torch.rand((batch, n, n), dtype=torch.float64, device="cuda")
B = torch.tril(torch.rand((n, n), dtype=torch.float64, device="cuda")).T.contiguous()  # column-major lower triangular
C1 = torch.zeros((batch, n, n), dtype=torch.float64, device="cuda")
C2 = torch.zeros_like(C1)
s = torch.cuda.current_stream().cuda_stream


def timed(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


for tri in (1, 0):
    flops = 2.0 * batch * n ** 3 * (0.5 if tri else 1.0)
    ldg = lambda: L.cxb_dgemm_ex(dev.vp(s), 2, 1, 0, 0, n, n, n, 1.0, dev.ptr(A), n, n * n, dev.ptr(B), n, 0, 0.0, dev.ptr(C1), n, n * n, batch, 0, 0, 0)  # noqa: E731
    if tri:
        # the production path: cxb_dgemm with the tri mask goes through the structured entry point of K1
        pass
    bulk = lambda: L.cxb_dgemm_bulk(dev.vp(s), n, n, n, dev.ptr(A), n, n * n, dev.ptr(B), n, 0, dev.ptr(C2), n, n * n, batch, tri)  # noqa: E731
    for name, fn in (("cp.async (LDGSTS), full k range" if tri else "cp.async (LDGSTS)", ldg), ("cp.async.bulk (TMA engine)", bulk)):
        best, med = timed(fn)
        print(json.dumps(dict(kernel=name, shape=f"n = {n}, batch = {batch}, tri = {tri}", best_ms=best, median_ms=med,
                              executed_TFLOPs=(2.0 * batch * n ** 3 if "full" in name or not tri else flops) / (best * 1e-3) / 1e12)))
    torch.cuda.synchronize()
    print(json.dumps(dict(check=f"tri = {tri}", identical=bool(torch.equal(C1, C2)))))
