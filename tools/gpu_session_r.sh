#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r_bench_c3.json 2> gpurun_out/r_bench_c3.err
grep -o '"value": [0-9.]*, "unit"\|"solve_ms": [0-9.]*\|"programs_per_s": [0-9.]*' gpurun_out/r_bench_c3.json | head; tail -2 gpurun_out/r_bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_c3_launches.csv python bench.py --workload c3 --programs 1024 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r_c3_launches.csv > gpurun_out/r_c3_launches.txt; head -24 gpurun_out/r_c3_launches.txt
