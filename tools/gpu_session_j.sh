#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/c1_trajectory.py > gpurun_out/j_c1_traj.txt 2>&1
cat gpurun_out/j_c1_traj.txt
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/j_t_all.log 2>&1
tail -15 gpurun_out/j_t_all.log
