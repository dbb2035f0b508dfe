#!/bin/bash
# compute-sanitizer passes over the kernels added in round 1's last sitting (not run yet: the round's
# GPU budget was spent on parity and measurement). memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards inside a CTA (the blocked diagonal Cholesky, the shared-memory LU
# panel, the wavefront solves' staging buffers); synccheck: divergent barriers.
# Usage (one B200): gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh'
mkdir -p gpurun_out
SEL='potrf or potrs or ldlt or pade or lu_solve or geodesic'
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests/test_gpu_kernels.py -q -x -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_supernodal.py tests/test_small_cones.py -m gpu -q -x > gpurun_out/sanitize_memcheck_sparse_small.log 2>&1
echo "memcheck (multifrontal + small cones): rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_sparse_small.log | tail -2
