#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_supernodal.py tests/test_gpu_solve.py tests/test_gpu_cones.py tests/test_gpu_hermitian.py -m gpu -q -x > gpurun_out/v_t.log 2>&1
tail -2 gpurun_out/v_t.log
timeout 100 python bench.py --workload sparse > gpurun_out/v_bench_sparse.json 2> gpurun_out/v_bench_sparse.err
grep -o '"value": [0-9.]*, "unit": "ms"\|"solve_ms": [0-9.]*' gpurun_out/v_bench_sparse.json | head -4; tail -2 gpurun_out/v_bench_sparse.err
