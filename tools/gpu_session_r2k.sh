#!/bin/bash
# Round 2, session k: source-level ncu captures of the three fused small-cone kernels (steady state of a C3 solve).
mkdir -p gpurun_out
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/k_tmp.json 2> gpurun_out/k_tmp.err
python -c "import json; raw=open('gpurun_out/k_tmp.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('c3 default', d['value'], d['solve_ms'], d['programs_per_s'], d['gpu_launches'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'PrepareMultiKernel|TakeStepMultiKernel|EigenMultiKernel' --launch-skip 9 -c 3 \
  -o gpurun_out/r02_k_c3_team_kernels python bench.py --workload c3 --no-cpu-baseline > /dev/null 2> gpurun_out/r02_k_c3_ncu.err
python tools/ncu_summary.py gpurun_out/r02_k_c3_team_kernels.ncu-rep > gpurun_out/r02_k_c3_team_kernels_ncu_full.txt
grep -E "kernel |time |dram_read|issue_active|warps_active|occ_limit|stall_" gpurun_out/r02_k_c3_team_kernels_ncu_full.txt | cut -c1-150
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
