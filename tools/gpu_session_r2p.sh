#!/bin/bash
# Round 2, session p: Gram contractions on the 64 x 128 / 32 x 64 configuration (A/B on the C2 step), kernel tests of the
# new configurations, C3 after the revert of the two-bracket multi-section.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "dgemm" 2>&1 | tail -3 | tee gpurun_out/r02_p_gemm_tests.txt
: > gpurun_out/r02_p_bench_c2_gram_config.txt
for variant in "" "--gram-config -1" "--gram-config 7"; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra $variant > gpurun_out/p_tmp.json 2> gpurun_out/p_tmp.err
  python -c "import json; raw=open('gpurun_out/p_tmp.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c2 [$variant]', d['value'], d['phase_ms'], 'frac', d['roofline']['frac'])" | tee -a gpurun_out/r02_p_bench_c2_gram_config.txt
done
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/p_tmp.json 2> gpurun_out/p_tmp.err
python -c "import json; raw=open('gpurun_out/p_tmp.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c3', d['value'], d['solve_ms'], d['programs_per_s'])" | tee gpurun_out/r02_p_bench_c3.txt
