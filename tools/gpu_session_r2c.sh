#!/bin/bash
# Round 2, session c: remaining full-size parity tests, the whole GPU suite with the polled triangular solves and the
# DMMA small-cone kernel, kernel timings (K3 / K5), the launch list of the batched solver + ncu --set full of its
# Schur kernel, racecheck of the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s --durations=0 > gpurun_out/r02_c_baseline_configs.txt 2>&1
grep -E "passed|failed|FAILED|durations|s call" gpurun_out/r02_c_baseline_configs.txt | tail -16
timeout 300 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_configs.py > gpurun_out/r02_c_gpu_tests_all.txt 2>&1
tail -4 gpurun_out/r02_c_gpu_tests_all.txt
timeout 300 python tools/prof_hbm_kernels.py time chol > gpurun_out/r02_c_kernel_timings.jsonl 2> gpurun_out/r02_c_kernel_timings.err
cut -c1-220 gpurun_out/r02_c_kernel_timings.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c_c3_launches.csv \
  python bench.py --workload c3 --programs 1024 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_c_c3_launches.err
python tools/launch_summary.py gpurun_out/r02_c_c3_launches.csv > gpurun_out/r02_c_c3_launches_1024_programs.txt
head -14 gpurun_out/r02_c_c3_launches_1024_programs.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'PsdSchurMmaKernel|PrepareKernel' --launch-skip 12 -c 4 \
  -o gpurun_out/r02_c_c3_schur_mma python bench.py --workload c3 --programs 4096 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_c_c3_ncu.err
ls -la gpurun_out/*.ncu-rep
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_small_cones.py -m gpu -q -x -p no:cacheprovider \
  -k "(potrf_and_potrs and (129 or 513)) or (dmma and 8-3) or (indefinite and device-8-5)" > gpurun_out/r02_sanitizer_racecheck_new_kernels.txt 2>&1
echo "racecheck new kernels: rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_racecheck_new_kernels.txt | tail -2
