#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/m_t_all.log 2>&1
tail -3 gpurun_out/m_t_all.log
timeout 200 python tools/prof_hbm_kernels.py time chol > gpurun_out/m_time.jsonl 2>&1
grep "K5" gpurun_out/m_time.jsonl
NCU="ncu --set full --clock-control none"
$NCU -k regex:TrsvFwdWave --launch-skip 1 -c 1 -f -o gpurun_out/m_trsvfwd_wave python tools/prof_hbm_kernels.py once chol > /dev/null 2>&1
$NCU -k regex:TrsvBwdWave --launch-skip 1 -c 1 -f -o gpurun_out/m_trsvbwd_wave python tools/prof_hbm_kernels.py once chol > /dev/null 2>&1
ls -la gpurun_out | grep m_
