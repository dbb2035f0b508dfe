#!/bin/bash
# First GPU session of the next round (one B200, ~6 min of box time): what round 1 ran out of budget for.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_kickoff.sh'
# 1. the whole GPU suite WITHOUT -x, so that every device variant of tests/test_zz_reference_python_suite.py
#    (added after the last GPU minute of round 1 was spent; validated on the oracle only) reports;
# 2. compute-sanitizer over the kernels of round 1's last sitting (tools/gpu_sanitize.sh);
# 3. a full-size launch list of `bench.py` (one Newton step after the warm-up, final kernels).
# Multi-GPU refresh (8 GPUs, separate call, ~3 min of box time):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round2_kickoff.sh multi'
mkdir -p gpurun_out
if [ "$1" == "multi" ]; then
  N=$(nvidia-smi -L | wc -l)
  RUN="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  $RUN --master-port 29541 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2_bench_c2_${N}gpu.json 2> gpurun_out/r2_bench_c2_${N}gpu.err
  $RUN --master-port 29542 bench.py --gpus $N --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
  $RUN --master-port 29543 bench.py --gpus $N --workload c3 --no-cpu-baseline > gpurun_out/r2_bench_c3_${N}gpu.json 2> gpurun_out/r2_bench_c3_${N}gpu.err
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2_t_multi_${N}gpu.log 2>&1
  tail -3 gpurun_out/r2_t_multi_${N}gpu.log
  for f in c2 c5 c3; do grep -o '"value": [0-9.]*, "unit": "ms", "n_gpus": [0-9]*\|"phase_ms": {[^}]*}' gpurun_out/r2_bench_${f}_${N}gpu.json | head -2; done
  exit 0
fi
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2_t_all.log 2>&1
tail -8 gpurun_out/r2_t_all.log
bash tools/gpu_sanitize.sh r02
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --launch-skip 4400 -c 2200 \
  --log-file gpurun_out/r2_bench_n2000_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_bench_n2000_launches.csv > gpurun_out/r2_bench_n2000_launches.txt
head -14 gpurun_out/r2_bench_n2000_launches.txt
