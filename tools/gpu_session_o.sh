#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_supernodal.py -m gpu -q > gpurun_out/o_t_sparse.log 2>&1
tail -12 gpurun_out/o_t_sparse.log
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/o_t_all.log 2>&1
tail -3 gpurun_out/o_t_all.log
timeout 300 python tools/sparse_bench.py > gpurun_out/o_sparse_bench.jsonl 2> gpurun_out/o_sparse_bench.err
cat gpurun_out/o_sparse_bench.jsonl; tail -3 gpurun_out/o_sparse_bench.err
