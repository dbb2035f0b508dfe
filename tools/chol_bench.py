#!/usr/bin/env python
"""Times cxb_potrf_lower / cxb_potrs_lower (K3 / K5) on SPD matrices of the BASELINE sizes."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devlib  # noqa: E402

L = devlib.product().lib
vp = C.c_void_p
for m in [int(x) for x in (sys.argv[1:] or ["2000", "10001", "20000"])]:
    ld = m + 2
    R = torch.randn(m, m // 4 + 8, dtype=torch.float64, device="cuda")
    H0 = torch.zeros(m, ld, dtype=torch.float64, device="cuda")  # column-major (ld x m) as (m, ld) rows
    H0[:, :m] = R @ R.T + m * torch.eye(m, dtype=torch.float64, device="cuda")
    del R
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    x0 = torch.randn(m, dtype=torch.float64, device="cuda")
    st = vp(torch.cuda.current_stream().cuda_stream)
    best_f, best_s = 1e30, 1e30
    for _ in range(3):
        H = H0.clone()
        x = x0.clone()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        assert L.cxb_potrf_lower(st, m, vp(H.data_ptr()), ld, None, vp(info.data_ptr())) == 0
        e[1].record()
        assert L.cxb_potrs_lower(st, m, vp(H.data_ptr()), ld, vp(x.data_ptr()), m, 1) == 0
        e[2].record()
        torch.cuda.synchronize()
        best_f = min(best_f, e[0].elapsed_time(e[1]))
        best_s = min(best_s, e[1].elapsed_time(e[2]))
    assert int(info.cpu()[0]) == 0
    res = (torch.tril(H0[:, :m]) @ x + torch.tril(H0[:, :m], -1).T @ x - x0).abs().max().item() / x0.abs().max().item()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    Hc = H0[:, :m].contiguous()
    torch.linalg.cholesky(Hc)
    t0.record(); torch.linalg.cholesky(Hc); t1.record(); torch.cuda.synchronize()
    print(f"m={m}: potrf {best_f:.3f} ms ({m ** 3 / 3 / best_f / 1e9:.2f} TF/s), potrs {best_s:.3f} ms "
          f"({2 * 8.0 * m * m / best_s / 1e6:.0f} GB/s of L reads), residual {res:.2e}; "
          f"cuSOLVER potrf (torch) {t0.elapsed_time(t1):.3f} ms")
