#!/usr/bin/env python
"""Parity of the two Schur-assembly forms of the device path with the oracle (reference algorithm as
written) on whole solves: iteration counts and relative differences of the final objectives.
  classic   (CONEXB200_SetAssemblyMode 1): H_ij = <W A_i W, A_j>, the reference's own formula
  symmetric (mode 3, default when it fits): H_ij = <L^T A_i L, L^T A_j L>, W = L L^T
Also lists the oracle's own symmetric variant (ORACLE_SetGramVariant 2) so that the effect of the
reformulation can be told apart from the device arithmetic. Usage: python tools/form_parity.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devlib  # noqa: E402
from harness import lovasz_theta_lmi, maxcut_lmi, oracle, random_dense_lmi  # noqa: E402


def run(L, mats, Cm, b, mode=None, variant=None):
    P = L.program()
    if mode is not None:
        L.lib.CONEXB200_SetAssemblyMode(P.h, mode)
    P.add_dense_lmi(mats, Cm)
    if variant is not None:
        L.lib.ORACLE_SetGramVariant(P.h, variant)
    bb = P.feasible_objective() if b is None else b
    solved, y = P.maximize(bb, L.default_config(prepare_dual_variables=1))
    lg = P.iteration_log()
    return solved, len(lg), lg[-1]["by"], lg[-1]["cx"], y


def main():
    O, D = oracle(), devlib.product()
    cases = {}
    for n in (40, 60, 100, 150):
        cases[f"maxcut n={n}"] = maxcut_lmi(n, 2)
    cases["lovasz n=30 e=80"] = lovasz_theta_lmi(30, 80, 4)
    cases["lovasz n=50 e=200"] = lovasz_theta_lmi(50, 200, 5)
    for (n, m, s) in ((50, 100, 1), (30, 20, 2), (80, 40, 3), (120, 10, 4)):
        mats, Cm = random_dense_lmi(n, m, s)
        cases[f"random n={n} m={m}"] = (mats, Cm, None)
    print(f"{'instance':22s} {'its ref/osym/cls/sym':>22s} {'by: osym':>10s} {'cls':>10s} {'sym':>10s} "
          f"{'cx: osym':>10s} {'cls':>10s} {'sym':>10s} {'|y-yref| cls':>13s} {'sym':>10s}")
    for name, (mats, Cm, b) in cases.items():
        ref = run(O, mats, Cm, b)
        osym = run(O, mats, Cm, b, variant=2)
        cls = run(D, mats, Cm, b, mode=1)
        sym = run(D, mats, Cm, b, mode=3)
        rel = lambda a, r: abs(a - r) / max(1.0, abs(r))
        ys = max(1.0, np.abs(ref[4]).max())
        print(f"{name:22s} {ref[1]:5d}/{osym[1]:3d}/{cls[1]:3d}/{sym[1]:3d}       "
              f"{rel(osym[2], ref[2]):10.1e} {rel(cls[2], ref[2]):10.1e} {rel(sym[2], ref[2]):10.1e} "
              f"{rel(osym[3], ref[3]):10.1e} {rel(cls[3], ref[3]):10.1e} {rel(sym[3], ref[3]):10.1e} "
              f"{np.abs(cls[4] - ref[4]).max() / ys:13.1e} {np.abs(sym[4] - ref[4]).max() / ys:10.1e}")


if __name__ == "__main__":
    main()
