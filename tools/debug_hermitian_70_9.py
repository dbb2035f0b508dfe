#!/usr/bin/env python
"""Iteration logs of the incremental (Hermitian) LMI n = 70, m = 9 instance of tests/test_gpu_hermitian.py on the
oracle and on the device under the three triangular-solve schemes (cxb_set_trsv_mode)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_hermitian as T  # noqa: E402
from harness import oracle  # noqa: E402
import devlib  # noqa: E402

O, D = oracle(), devlib.product()
D.lib.cxb_set_trsv_mode.argtypes = [C.c_int]
mats, Cm = T.random_instance(70, 9, 500 + 70 + 9)
runs = [("oracle", O, None)] + [(f"device trsv_mode={m}", D, m) for m in (0, 1, 2)]
for name, L, mode in runs:
    if mode is not None:
        D.lib.cxb_set_trsv_mode(mode)
    P = L.program(9)
    P.add_hermitian_lmi(mats, Cm)
    b = P.feasible_objective()
    T.srand(11)
    solved, y = P.maximize(b, L.default_config(prepare_dual_variables=1))
    log = P.iteration_log()
    print(name, "solved", solved, "iterations", len(log))
    for i, r in enumerate(log):
        print(f"  {i:2d} k={r['inv_sqrt_mu']:.15e} d_inf={r['d_inf']:.6e} d_2={r['d_2']:.6e} by={r['by']:.12e} step={r['step_size']:.3e}")
D.lib.cxb_set_trsv_mode(2)
