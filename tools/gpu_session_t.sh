#!/bin/bash
# Two B200s: the multi-GPU parity tests at world 2 and the default bench exactly as the driver launches it.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/t_t_multi.log 2>&1
tail -3 gpurun_out/t_t_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t_bench_c2_2gpu.json 2> gpurun_out/t_bench_c2_2gpu.err
grep -o '"value": [0-9.]*, "unit": "ms", "n_gpus": [0-9]*\|"phase_ms": {[^}]*}' gpurun_out/t_bench_c2_2gpu.json | head -2; tail -2 gpurun_out/t_bench_c2_2gpu.err
