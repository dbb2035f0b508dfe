#!/bin/bash
# Session "n" on four B200s: BASELINE config 5 (n = 1000, m = 20000) sharded, with the Schur complement
# factored across the ranks and, for comparison, replicated.
mkdir -p gpurun_out
RUN="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$RUN --master-port 29521 bench.py --gpus 4 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c5_dist.json 2> gpurun_out/n_c5_dist.err
$RUN --master-port 29522 bench.py --gpus 4 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --replicated-cholesky > gpurun_out/n_c5_repl.json 2> gpurun_out/n_c5_repl.err
for f in n_c5_dist n_c5_repl; do echo $f; grep -o '"value": [0-9.]*, "unit": "ms", "n_gpus"\|"phase_ms": {[^}]*}' gpurun_out/$f.json | head -2; tail -2 gpurun_out/$f.err; done
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "cholesky and 4" > gpurun_out/n_t_multi4.log 2>&1
tail -2 gpurun_out/n_t_multi4.log
