#!/bin/bash
# Round 2, session d: the Hermitian n=70 instance under the three triangular-solve schemes, re-run of the tests that
# failed in session c, the 16-warp DMMA Schur kernel (bench + launch list + ncu --set full), ncu --set full of the
# symmetric-form assembly GEMMs of C2 (K1a tri=1, K1b tri=2 + pack, K2 packed Gram) for roofline.traffic.
mkdir -p gpurun_out
timeout 120 python tools/debug_hermitian_70_9.py > gpurun_out/r02_d_hermitian_70_9.txt 2>&1
grep -E "iterations" gpurun_out/r02_d_hermitian_70_9.txt
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -k "n400 or c4 or c3" > gpurun_out/r02_d_baseline_configs.txt 2>&1
grep -E "passed|failed|FAILED|oracle:|symmetric:|classic:|structured:|worst" gpurun_out/r02_d_baseline_configs.txt | cut -c1-300
timeout 300 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_configs.py > gpurun_out/r02_d_gpu_tests_all.txt 2>&1
tail -6 gpurun_out/r02_d_gpu_tests_all.txt
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02_d_bench_c3.json 2> gpurun_out/r02_d_bench_c3.err
python -c "import json; d=json.load(open('gpurun_out/r02_d_bench_c3.json')); print('c3', d['value'], d['solve_ms'], d['programs_per_s'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_d_c3_launches.csv \
  python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_d_c3_launches.err
python tools/launch_summary.py gpurun_out/r02_d_c3_launches.csv > gpurun_out/r02_d_c3_launches_4096_programs.txt
head -14 gpurun_out/r02_d_c3_launches_4096_programs.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'PsdSchurMmaKernel' --launch-skip 7 -c 2 \
  -o gpurun_out/r02_d_c3_schur_mma16 python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_d_c3_ncu.err
timeout 400 ncu --set full --clock-control none -k regex:'DgemmKernel' --launch-skip 100 -c 48 \
  -o gpurun_out/r02_d_c2_sym_assembly python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra > /dev/null 2> gpurun_out/r02_d_c2_ncu.err
ls -la gpurun_out/*.ncu-rep
