#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/l_t_all.log 2>&1
tail -6 gpurun_out/l_t_all.log
timeout 200 python tools/prof_hbm_kernels.py time chol,geo > gpurun_out/l_time.jsonl 2>&1
cat gpurun_out/l_time.jsonl
timeout 120 python bench.py --workload c4s --no-cpu-baseline > gpurun_out/l_bench_c4s.json 2> gpurun_out/l_bench_c4s.err
grep -o '"value": [0-9.]*, "unit"\|"phase_ms": {[^}]*}' gpurun_out/l_bench_c4s.json | head -3; tail -2 gpurun_out/l_bench_c4s.err
