#!/usr/bin/env python
"""Block-arrow program (cones on private variable blocks + a few shared variables) solved through
CONEX_AddSparseLMIConstraint / CONEX_Maximize with the dense KKT solver (kind 1) and the multifrontal one
(kind 2): per-phase device times of a Newton step, iteration counts and objectives side by side.
Usage: python tools/sparse_bench.py [blocks private shared order]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devlib  # noqa: E402
from test_supernodal import block_arrow_program  # noqa: E402


def main():
    blocks, private, shared, order = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (8, 1500, 100, 60)
    dev = devlib.product()
    L = dev.lib
    L.CONEXB200_SetKKTSolverKind.argtypes = [C.c_void_p, C.c_int]
    L.CONEXB200_SetKKTSolverKind.restype = None
    m, cones = block_arrow_program(blocks, private, shared, order, seed=11)
    for kind, name in ((1, "dense"), (2, "multifrontal")):
        P = dev.program(m)
        L.CONEXB200_SetKKTSolverKind(P.h, kind)
        for mats, Cm, variables in cones:
            P.add_dense_lmi(mats, Cm, variables)
        b = P.feasible_objective()
        for rep in range(2):  # second solve: warmed-up allocations
            solved, y = P.maximize(b, dev.default_config(prepare_dual_variables=1))
        its = P.status()["num_iterations"]
        assert its > 2, "the solve did not run (singular Schur complement: each cone needs at most n(n+1)/2 variables)"
        phases = []
        for i in range(its):
            ph = np.zeros(5)
            L.CONEXB200_GetIterationPhaseMilliseconds(P.h, i, ph.ctypes.data_as(C.POINTER(C.c_double)))
            phases.append(ph)
        ph = np.array(phases[2:]).mean(axis=0)
        log = P.iteration_log()
        print(json.dumps({
            "workload": f"block_arrow_{blocks}x(psd{order} on {private} private + {shared} shared variables)",
            "kkt_order": m, "kkt_solver": name, "supernodes": L.CONEXB200_GetNumberOfSupernodes(P.h),
            "solved": int(solved), "iterations": its, "by": log[-1]["by"],
            "newton_step_ms": float(ph.sum()),
            "phase_ms": dict(zip(["assemble", "factor", "mu", "solve", "update"], [round(float(v), 3) for v in ph])),
        }))


if __name__ == "__main__":
    main()
