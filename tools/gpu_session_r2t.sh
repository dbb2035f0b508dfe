#!/bin/bash
# Round 2, session t: the ncu launch list of the default bench command (C2, shortened to 1 + 2 steps, no extras).
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/t_c2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-solve --no-extra > gpurun_out/t_tmp.json 2> gpurun_out/r02_t_c2_launches.err
python tools/launch_summary.py gpurun_out/t_c2_launches.csv > gpurun_out/r02_t_c2_launches_default_bench.txt
rm -f gpurun_out/t_c2_launches.csv
head -24 gpurun_out/r02_t_c2_launches_default_bench.txt | cut -c1-170
tail -2 gpurun_out/r02_t_c2_launches.err | cut -c1-200
du -sh gpurun_out
