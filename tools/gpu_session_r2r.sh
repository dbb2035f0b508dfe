#!/bin/bash
# Round 2, final single-GPU validation (second pass, after the small-cone work): the whole GPU suite (incl. the full-size parity tests), smoke(), the default bench
# line exactly as the driver runs it, the structured / sparse workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_r_gpu_tests_all.txt 2>&1
grep -E "passed|failed|FAILED|^E  |s call" gpurun_out/r02_r_gpu_tests_all.txt | cut -c1-200 | tail -14
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_r_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_r_bench_c2.json 2> gpurun_out/r02_r_bench_c2.err
python - <<PY
import json
raw = open("gpurun_out/r02_r_bench_c2.json").read()
d = json.loads([l for l in raw.splitlines() if l.startswith("{")][0])
print("c2", d["value"], d["phase_ms"], "frac", d["roofline"]["frac"], "solve", d.get("solve_ms"), d.get("solve_iterations"), d.get("solved"))
print("e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "setup", d.get("operator_setup"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:160], d["cpu_baseline"].get("model_check", {}).get("measured_over_predicted"))
print("c5", json.dumps(d.get("c5"))[:400])
print("c3", json.dumps(d.get("c3"))[:400])
PY
for wl in c2s c4s sparse c4; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r02_r_bench_$wl.json 2> gpurun_out/r_tmp.err
  python -c "import json; raw=open('gpurun_out/r02_r_bench_$wl.json').read(); d=json.loads([l for l in raw.splitlines() if l.startswith('{')][0]); print('$wl', d['value'], d.get('phase_ms'), d['e2e'].get('solve_ms'), d['e2e'].get('solve_iterations'))"
done
du -sh gpurun_out
# compute-sanitizer over the small-cone kernels changed in this part of the round (fused launches, ballot multi-section,
# compile-time-order DMMA Schur kernel with the shared leftover matrix, packed slack passes, warp-0 pivot search)
NEW='fused or dmma_schur or packed or both_thread_layouts or (trajectory and device)'
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/r02_r_sanitizer_${tool}_small_cones.txt
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_small_cones.py tests/test_gpu_batch.py -m gpu -q -x -p no:cacheprovider -k "$NEW or batch" > $log 2>&1
  echo "$tool: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -E ' passed| failed| error' $log | tail -1)"
done
du -sh gpurun_out
