#!/usr/bin/env python
"""One cuBLAS DGEMM and one conex-b200 DGEMM of the same shape, for side-by-side ncu captures."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devlib  # noqa: E402

L = devlib.product().lib
vp = C.c_void_p
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
f64 = dict(dtype=torch.float64, device="cuda")
A = torch.randn(n, n, **f64)
B = torch.randn(n, n, **f64)
Cm = torch.empty(n, n, **f64)
for _ in range(2):
    torch.matmul(A, B, out=Cm)
    L.cxb_dgemm_ex(vp(torch.cuda.current_stream().cuda_stream), cfg, 1, 0, 0, n, n, n, 1.0, vp(A.data_ptr()), n,
                   0, vp(B.data_ptr()), n, 0, 0.0, vp(Cm.data_ptr()), n, 0, 1, 0, 0, 0)
torch.cuda.synchronize()
