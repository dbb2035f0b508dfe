#!/bin/bash
# Round 2, session g: tests after the LDLT-front / layout changes, C3 A/B, split-K policy A/B on C2 (time + DRAM traffic of K2).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_configs.py > gpurun_out/r02_g_gpu_tests_all.txt 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_g_gpu_tests_all.txt | cut -c1-300 | tail -12
for variant in "" "--small-team-mode 0" "--small-team-mode 1" "--small-psd-mma 1"; do
  timeout 300 python bench.py --workload c3 --no-cpu-baseline $variant > gpurun_out/g_tmp.json 2> gpurun_out/g_tmp.err
  python -c "import json; d=json.load(open('gpurun_out/g_tmp.json')); print('c3 [$variant]', d['value'], d['solve_ms'], d['programs_per_s'])" | tee -a gpurun_out/r02_g_bench_c3_variants.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_c3_launches.csv \
  python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_g_c3_launches.err
python tools/launch_summary.py gpurun_out/g_c3_launches.csv > gpurun_out/r02_g_c3_launches_4096_programs.txt
rm -f gpurun_out/g_c3_launches.csv
head -16 gpurun_out/r02_g_c3_launches_4096_programs.txt
for kt in 0 4096 1024; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra --gemm-max-ktiles $kt > gpurun_out/g_tmp.json 2> gpurun_out/g_tmp.err
  python -c "import json; d=json.load(open('gpurun_out/g_tmp.json')); print('c2 max_ktiles=$kt', d['value'], d['phase_ms'], d['roofline']['frac'])" | tee -a gpurun_out/r02_g_c2_split_policy.txt
done
timeout 400 ncu --set full --clock-control none -k regex:'DgemmKernel' --launch-skip 140 -c 3 \
  -o gpurun_out/g_c2_k2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra --gemm-max-ktiles 4096 > /dev/null 2> gpurun_out/r02_g_c2_ncu.err
python tools/ncu_summary.py gpurun_out/g_c2_k2.ncu-rep > gpurun_out/r02_g_c2_k2_split_policy_4096_ncu_full.txt
rm -f gpurun_out/g_c2_k2.ncu-rep
grep -E "kernel |grid |time |dram_read|dmma_pipe|l2_hit" gpurun_out/r02_g_c2_k2_split_policy_4096_ncu_full.txt | cut -c1-150
du -sh gpurun_out
