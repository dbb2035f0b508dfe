#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/k_t_all.log 2>&1
tail -8 gpurun_out/k_t_all.log
timeout 200 python tools/prof_hbm_kernels.py time chol,geo > gpurun_out/k_time.jsonl 2>&1
cat gpurun_out/k_time.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_geo_launches.csv python tools/prof_hbm_kernels.py once geo > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/k_geo_launches.csv | head -12
