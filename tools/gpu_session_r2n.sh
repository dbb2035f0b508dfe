#!/bin/bash
# Round 2, session n: dual-bracket multi-section, warp-0 pivot search, 2 x 2 register tiles in the small products.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_small_cones.py tests/test_gpu_batch.py tests/test_gpu_cones.py tests/test_gpu_hermitian.py -m gpu -q > gpurun_out/r02_n_small_cone_tests.txt 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_n_small_cone_tests.txt | cut -c1-220 | tail -12
timeout 300 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k "c3" 2>&1 | tail -3 | tee -a gpurun_out/r02_n_small_cone_tests.txt
: > gpurun_out/r02_n_bench_c3_variants.txt
for variant in "" "--programs 512" "--programs 16384"; do
  timeout 300 python bench.py --workload c3 --no-cpu-baseline $variant > gpurun_out/n_tmp.json 2> gpurun_out/n_tmp.err
  python -c "import json; raw=open('gpurun_out/n_tmp.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c3 [$variant]', d['value'], d['solve_ms'], d['programs_per_s'], d['gpu_launches'])" | tee -a gpurun_out/r02_n_bench_c3_variants.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_c3_launches.csv \
  python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_n_c3_launches.err
python tools/launch_summary.py gpurun_out/n_c3_launches.csv > gpurun_out/r02_n_c3_launches_4096_programs.txt
rm -f gpurun_out/n_c3_launches.csv
head -16 gpurun_out/r02_n_c3_launches_4096_programs.txt | cut -c1-150
du -sh gpurun_out
