#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_small_cones.py tests/test_gpu_batch.py tests/test_gpu_cones.py -m gpu -q > gpurun_out/s_t_small.log 2>&1
tail -6 gpurun_out/s_t_small.log
timeout 200 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/s_bench_c3.json 2> gpurun_out/s_bench_c3.err
grep -o '"value": [0-9.]*, "unit"\|"solve_ms": [0-9.]*\|"programs_per_s": [0-9.]*' gpurun_out/s_bench_c3.json | head -3; tail -2 gpurun_out/s_bench_c3.err
