#!/bin/bash
# Round 2, session b: full-size parity tests, the whole GPU suite, racecheck of the small-cone / batch / multifrontal
# kernels (the selection that timed out in session a, narrowed), one default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -x -s --durations=0 > gpurun_out/r02_b_baseline_configs.txt 2>&1
tail -25 gpurun_out/r02_b_baseline_configs.txt
timeout 300 python -m pytest tests -m gpu -q --deselect tests/test_gpu_baseline_configs.py > gpurun_out/r02_b_gpu_tests_all.txt 2>&1
tail -5 gpurun_out/r02_b_gpu_tests_all.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_small_cones.py tests/test_gpu_batch.py tests/test_supernodal.py -m gpu -q -x -p no:cacheprovider \
  -k "(trajectory and (device-2-5-3 or device-2-9-4 or device-1-5-3 or device-0-7-3)) or (small_cholesky and device-17) or (per_program and tiny) or lapack" > gpurun_out/r02_sanitizer_racecheck_sparse_small_batch.txt 2>&1
echo "racecheck small: rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_racecheck_sparse_small_batch.txt | tail -2
timeout 400 python bench.py > gpurun_out/r02_b_bench_c2.json 2> gpurun_out/r02_b_bench_c2.err
tail -c 600 gpurun_out/r02_b_bench_c2.json
