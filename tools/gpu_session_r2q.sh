#!/bin/bash
# Round 2, session q: 64 x 64 tile with BK = 32 (configuration 9) against the default on the C2 step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "dgemm_every_tile_configuration and 9" 2>&1 | tail -3 | tee gpurun_out/r02_q_gemm_tests.txt
: > gpurun_out/r02_q_bench_c2_bk32.txt
for variant in "--gemm-config 9" ""; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra $variant > gpurun_out/q_tmp.json 2> gpurun_out/q_tmp.err
  python -c "import json; raw=open('gpurun_out/q_tmp.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c2 [$variant]', d['value'], d['phase_ms'], 'frac', d['roofline']['frac'])" | tee -a gpurun_out/r02_q_bench_c2_bk32.txt
done
tail -3 gpurun_out/q_tmp.err
