#!/bin/bash
# Round 2, session j: the batched small-cone engine after the fused multi-cone launches, the ballot-based multi-section
# and the smaller eigen-bound CTAs: parity tests, A/B of the switches at 4096 and at 512 programs, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_small_cones.py tests/test_gpu_batch.py tests/test_gpu_cones.py tests/test_gpu_hermitian.py -m gpu -q -x > gpurun_out/r02_j_small_cone_tests.txt 2>&1
tail -3 gpurun_out/r02_j_small_cone_tests.txt
timeout 300 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k "c3" 2>&1 | tail -3 | tee -a gpurun_out/r02_j_small_cone_tests.txt
: > gpurun_out/r02_j_bench_c3_variants.txt
for variant in "" "--small-fused-launches 0" "--small-cone-threads 64" "--small-cone-threads 32" "--small-cone-threads 64 --small-fused-launches 0" \
               "--programs 512" "--programs 512 --small-fused-launches 0" "--programs 512 --small-cone-threads 64" "--programs 512 --small-cone-threads 32"; do
  timeout 300 python bench.py --workload c3 --no-cpu-baseline $variant > gpurun_out/j_tmp.json 2> gpurun_out/j_tmp.err
  python -c "import json; raw=open('gpurun_out/j_tmp.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c3 [$variant]', d['value'], d['solve_ms'], d['programs_per_s'], d['gpu_launches'])" | tee -a gpurun_out/r02_j_bench_c3_variants.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_c3_launches.csv \
  python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_j_c3_launches.err
python tools/launch_summary.py gpurun_out/j_c3_launches.csv > gpurun_out/r02_j_c3_launches_4096_programs.txt
rm -f gpurun_out/j_c3_launches.csv
head -18 gpurun_out/r02_j_c3_launches_4096_programs.txt | cut -c1-160
du -sh gpurun_out
