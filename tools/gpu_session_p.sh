#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_supernodal.py tests/test_gpu_solve.py tests/test_gpu_cones.py -m gpu -q > gpurun_out/p_t_sparse.log 2>&1
tail -4 gpurun_out/p_t_sparse.log
timeout 300 python tools/sparse_bench.py > gpurun_out/p_sparse_bench.jsonl 2> gpurun_out/p_sparse_bench.err
cat gpurun_out/p_sparse_bench.jsonl; tail -3 gpurun_out/p_sparse_bench.err
timeout 300 python tools/sparse_bench.py 32 400 40 30 > gpurun_out/p_sparse_bench32.jsonl 2>> gpurun_out/p_sparse_bench.err
cat gpurun_out/p_sparse_bench32.jsonl
