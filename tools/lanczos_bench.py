#!/usr/bin/env python
"""Times cxb_lanczos_two_sided (K7) on a random W-weighted problem that does not break down:
ms per call and us per Lanczos step, on a non-default stream (CUDA-graph replay) and on the
default stream (direct launches). Usage: python tools/lanczos_bench.py [n ...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devlib as dev  # noqa: E402
import torch  # noqa: E402


def main():
    L = dev.product().lib
    sizes = [int(a) for a in sys.argv[1:]] or [500, 1000, 2000]
    for n in sizes:
        rng = np.random.default_rng(n)
        R = rng.standard_normal((n, n))
        W = R @ R.T / n + np.eye(n)
        S = rng.standard_normal((n, n))
        S = (S + S.T) / np.sqrt(n)
        dWS, dW, dr = dev.to_dev(W @ S), dev.to_dev(W), dev.to_dev(rng.standard_normal(n))
        it = n // 2
        alpha, beta, count = dev.dzeros(it + 2), dev.dzeros(it + 2), dev.izeros(2)
        work = dev.dzeros(L.cxb_lanczos_worksize(n))
        stream = torch.cuda.Stream()
        for name, sp in (("graph", C.c_void_p(stream.cuda_stream)), ("direct", None)):
            ts = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream if sp else torch.cuda.default_stream()):
                    e0.record()
                    rc = L.cxb_lanczos_two_sided(sp, n, dev.ptr(dWS), dev.ptr(dW), dev.ptr(dr), None, it,
                                                 dev.ptr(alpha), dev.ptr(beta), dev.ptr(count), dev.ptr(work))
                    e1.record()
                torch.cuda.synchronize()
                assert rc == 0
                ts.append(e0.elapsed_time(e1))
            steps = int(count.cpu()[0]) + 1
            print(f"n={n} {name}: {min(ts):.3f} ms per call, {steps} steps, {min(ts) * 1e3 / steps:.1f} us/step "
                  f"(first call {ts[0]:.3f} ms)")


if __name__ == "__main__":
    main()
