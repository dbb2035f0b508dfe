#!/bin/bash
# Round-1 session "g" on one B200: the GPU parity suite, the headline bench, C4 dense, CUDA-event
# timings and ncu captures of the bandwidth-bound kernels (K5/K6/K7) and of the K3/K8 launch chains.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/g_t_potrf.log 2>&1
tail -3 gpurun_out/g_t_potrf.log
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/g_t_all.log 2>&1
tail -20 gpurun_out/g_t_all.log
python bench.py > gpurun_out/g_bench_c2.json 2> gpurun_out/g_bench_c2.err
python bench.py --workload c4 > gpurun_out/g_bench_c4.json 2> gpurun_out/g_bench_c4.err
python tools/prof_hbm_kernels.py time > gpurun_out/g_hbm_time.jsonl 2>&1
cat gpurun_out/g_hbm_time.jsonl
NCU="ncu --set full --clock-control none"
$NCU -k regex:GemvN --launch-skip 1 -c 1 -f -o gpurun_out/g_gemv python tools/prof_hbm_kernels.py once gemv > /dev/null 2>&1
$NCU -k regex:TrsvFwdStep --launch-skip 80 -c 2 -f -o gpurun_out/g_trsvfwd python tools/prof_hbm_kernels.py once chol > /dev/null 2>&1
$NCU -k regex:TrsvBwdStep --launch-skip 80 -c 2 -f -o gpurun_out/g_trsvbwd python tools/prof_hbm_kernels.py once chol > /dev/null 2>&1
$NCU -k regex:LanczosStep --launch-skip 1100 -c 2 -f -o gpurun_out/g_lanczos python tools/prof_hbm_kernels.py once lanczos > /dev/null 2>&1
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$LIST --log-file gpurun_out/g_chol_launches.csv python tools/prof_hbm_kernels.py once chol > /dev/null 2>&1
$LIST --log-file gpurun_out/g_geo_launches.csv python tools/prof_hbm_kernels.py once geo > /dev/null 2>&1
ls -la gpurun_out | tail -20
