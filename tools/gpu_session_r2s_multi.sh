#!/bin/bash
# Round 2, last multi-GPU session (gpurun --gpus 8): the batched workload after the small-cone work, programs partitioned
# over the ranks (bench.py --workload c3), and the batch tests of test_multi_gpu.py.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus $N --workload c3 --no-cpu-baseline > gpurun_out/r02_s_bench_c3_${N}gpu.json 2> gpurun_out/r02_s_bench_c3_${N}gpu.err
tail -2 gpurun_out/r02_s_bench_c3_${N}gpu.err | cut -c1-300
python -c "import json; raw=open('gpurun_out/r02_s_bench_c3_${N}gpu.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c3 x$N', d['value'], d.get('solve_ms'), d.get('programs_per_s'), d.get('gpu_launches'), d['config'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus $N --workload c3 --programs 32768 --no-cpu-baseline > gpurun_out/r02_s_bench_c3_32768_${N}gpu.json 2> gpurun_out/s_tmp.err
python -c "import json; raw=open('gpurun_out/r02_s_bench_c3_32768_${N}gpu.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c3 32768 x$N', d['value'], d.get('solve_ms'), d.get('programs_per_s'))"
