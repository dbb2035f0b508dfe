#!/bin/bash
# Round-1 session "h" on two B200s: the multi-GPU Cholesky (parity tests, then A/B benches against the
# replicated factorisation on a Cholesky-heavy dense workload and on the entry-sparse Lovasz theta).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "cholesky or random-8 or maxcut-16" > gpurun_out/h_t_multi.log 2>&1
tail -5 gpurun_out/h_t_multi.log
RUN="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$RUN --master-port 29511 bench.py --gpus 2 --workload c5 --size 200 --constraints 20000 --no-cpu-baseline > gpurun_out/h_c5s_dist.json 2> gpurun_out/h_c5s_dist.err
$RUN --master-port 29512 bench.py --gpus 2 --workload c5 --size 200 --constraints 20000 --no-cpu-baseline --replicated-cholesky > gpurun_out/h_c5s_repl.json 2> gpurun_out/h_c5s_repl.err
$RUN --master-port 29513 bench.py --gpus 2 --workload c4s --no-cpu-baseline > gpurun_out/h_c4s_dist.json 2> gpurun_out/h_c4s_dist.err
$RUN --master-port 29514 bench.py --gpus 2 --workload c4s --no-cpu-baseline --replicated-cholesky > gpurun_out/h_c4s_repl.json 2> gpurun_out/h_c4s_repl.err
for f in h_c5s_dist h_c5s_repl h_c4s_dist h_c4s_repl; do echo $f; tail -c 1500 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
