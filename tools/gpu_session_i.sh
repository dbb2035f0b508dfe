#!/bin/bash
# Session "i" (one B200): blocked diagonal Cholesky kernel + single-launch wavefront triangular solves.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf or potrs or ldlt" > gpurun_out/i_t_chol.log 2>&1
tail -3 gpurun_out/i_t_chol.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/i_t_all.log 2>&1
tail -3 gpurun_out/i_t_all.log
timeout 200 python tools/prof_hbm_kernels.py time chol > gpurun_out/i_chol_time.jsonl 2>&1
cat gpurun_out/i_chol_time.jsonl
timeout 120 python bench.py --workload c4s --no-cpu-baseline > gpurun_out/i_bench_c4s.json 2> gpurun_out/i_bench_c4s.err
tail -c 1200 gpurun_out/i_bench_c4s.json; tail -3 gpurun_out/i_bench_c4s.err
