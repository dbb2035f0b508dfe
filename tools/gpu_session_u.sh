#!/bin/bash
# Final session of round 1 on one B200: the whole GPU suite, smoke(), the headline bench with its CPU
# baseline, the sparse and entry-sparse workloads, ncu captures of this session's new kernels.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/u_t_all.log 2>&1
tail -4 gpurun_out/u_t_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; tail -1 gpurun_out/u_smoke.log
timeout 200 python bench.py > gpurun_out/u_bench_c2.json 2> gpurun_out/u_bench_c2.err
grep -o '"value": [0-9.]*, "unit": "ms", "n_gpus": [0-9]*\|"phase_ms": {[^}]*}' gpurun_out/u_bench_c2.json | head -2; tail -2 gpurun_out/u_bench_c2.err
timeout 100 python bench.py --workload sparse > gpurun_out/u_bench_sparse.json 2> gpurun_out/u_bench_sparse.err
grep -o '"value": [0-9.]*, "unit": "ms", "n_gpus": [0-9]*\|"phase_ms": {[^}]*}' gpurun_out/u_bench_sparse.json | head -3; tail -2 gpurun_out/u_bench_sparse.err
timeout 60 python bench.py --workload c4s --no-cpu-baseline > gpurun_out/u_bench_c4s.json 2> gpurun_out/u_bench_c4s.err
grep -o '"value": [0-9.]*, "unit": "ms", "n_gpus": [0-9]*\|"phase_ms": {[^}]*}' gpurun_out/u_bench_c4s.json | head -2
NCU="ncu --set full --clock-control none"
timeout 60 $NCU -k regex:PotrfDiagBlocked --launch-skip 80 -c 1 -f -o gpurun_out/u_potrf_diag python tools/prof_hbm_kernels.py once chol > /dev/null 2>&1
timeout 60 $NCU -k regex:LuPanelSmem --launch-skip 64 -c 1 -f -o gpurun_out/u_lu_panel python tools/prof_hbm_kernels.py once geo > /dev/null 2>&1
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_bench_n1000_launches.csv python bench.py --size 1000 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | grep "^u_"
