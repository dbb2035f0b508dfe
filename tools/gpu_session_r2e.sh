#!/bin/bash
# Round 2, session e (session d's outputs were lost: a 128 MB ncu report pushed gpurun_out over the 64 MiB limit).
# Keep every artefact small: digests are made on the box, big reports deleted.
mkdir -p gpurun_out
timeout 120 python tools/debug_hermitian_70_9.py > gpurun_out/r02_e_hermitian_70_9.txt 2>&1
grep -E "iterations" gpurun_out/r02_e_hermitian_70_9.txt
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -k "n400 or c3" > gpurun_out/r02_e_baseline_configs.txt 2>&1
grep -E "passed|failed|FAILED|oracle:|symmetric:|classic:|structured:|worst|^E  " gpurun_out/r02_e_baseline_configs.txt | cut -c1-300
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_configs.py > gpurun_out/r02_e_gpu_tests_all.txt 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_e_gpu_tests_all.txt | cut -c1-300 | tail -12
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02_e_bench_c3.json 2> gpurun_out/r02_e_bench_c3.err
python -c "import json; d=json.load(open('gpurun_out/r02_e_bench_c3.json')); print('c3 default', d['value'], d['solve_ms'], d['programs_per_s'])"
for variant in "--small-psd-mma 1" "--small-psd-mma 0" "--small-team-mode 0"; do
  timeout 300 python bench.py --workload c3 --no-cpu-baseline $variant > gpurun_out/e_tmp.json 2> gpurun_out/e_tmp.err
  python -c "import json; d=json.load(open('gpurun_out/e_tmp.json')); print('c3 $variant', d['value'], d['solve_ms'], d['programs_per_s'])" | tee -a gpurun_out/r02_e_bench_c3_variants.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_c3_launches.csv \
  python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_e_c3_launches.err
python tools/launch_summary.py gpurun_out/e_c3_launches.csv > gpurun_out/r02_e_c3_launches_4096_programs.txt
rm -f gpurun_out/e_c3_launches.csv
head -16 gpurun_out/r02_e_c3_launches_4096_programs.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'PsdSchurMma2Kernel|PsdFactorKernel' --launch-skip 12 -c 2 \
  -o gpurun_out/r02_e_c3_schur_mma16 python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_e_c3_ncu.err
python tools/ncu_summary.py gpurun_out/r02_e_c3_schur_mma16.ncu-rep > gpurun_out/r02_e_c3_schur_mma16_ncu_full.txt
timeout 400 ncu --set full --clock-control none -k regex:'DgemmKernel' --launch-skip 130 -c 12 \
  -o gpurun_out/e_c2_sym_assembly python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra > /dev/null 2> gpurun_out/r02_e_c2_ncu.err
python tools/ncu_summary.py gpurun_out/e_c2_sym_assembly.ncu-rep > gpurun_out/r02_e_c2_symmetric_assembly_ncu_full.txt
rm -f gpurun_out/e_c2_sym_assembly.ncu-rep
grep -E "kernel |grid |time |dram_read|dram_write|dmma_pipe|l2_hit" gpurun_out/r02_e_c2_symmetric_assembly_ncu_full.txt | cut -c1-160
du -sh gpurun_out
