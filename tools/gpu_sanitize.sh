#!/bin/bash
# compute-sanitizer passes over the device kernels (the reference's structural asserts,
# supernodal_assembler.cc:124-138, become memcheck / racecheck / synccheck runs here).
#   memcheck : out-of-bounds / misaligned accesses, whole kernel test file + multifrontal + batched solver
#   racecheck: shared-memory hazards inside a CTA (blocked diagonal Cholesky, shared-memory LU panel, the
#              wavefront solves' staging buffers, the small-cone team kernels)
#   synccheck: divergent / invalid barriers in the same kernels
# Usage (one B200): gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh [tag]'
# Logs: gpurun_out/<tag>_sanitizer_<tool>_<suite>.txt (copy the summaries to profiles/).
TAG=${1:-r02}
mkdir -p gpurun_out
SMALL='(potrf_and_potrs and (33 or 129 or 513 or 700)) or (by_panels and (97 or 700)) or non_positive_pivot or (pade and (31 or 129)) or (lu_solve and (40 or 257)) or geodesic or golden_4x4 or (lanczos_matches and (25 or 64)) or (triangular_gemm and (64 or 193)) or (schur_dense and (10-7 or 33-20))'
run() { # tool suite timeout pytest-args...
  local tool=$1 suite=$2 to=$3; shift 3
  local log=gpurun_out/${TAG}_sanitizer_${tool}_${suite}.txt
  timeout $to compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" -q -x -p no:cacheprovider > $log 2>&1
  echo "$tool/$suite: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -E ' passed| failed| error' $log | tail -1)"
}
run memcheck kernels 420 tests/test_gpu_kernels.py -m gpu
run memcheck sparse_small_batch 420 tests/test_supernodal.py tests/test_small_cones.py tests/test_gpu_batch.py -m gpu -k "not c3"
run racecheck kernels 420 tests/test_gpu_kernels.py -m gpu -k "$SMALL"
run racecheck sparse_small_batch 420 tests/test_supernodal.py tests/test_small_cones.py tests/test_gpu_batch.py -m gpu -k "arrow or lapack or trajectory or small_cholesky or tiny or mixed"
run synccheck kernels 300 tests/test_gpu_kernels.py -m gpu -k "$SMALL"
run synccheck sparse_small_batch 300 tests/test_supernodal.py tests/test_small_cones.py tests/test_gpu_batch.py -m gpu -k "arrow or lapack or trajectory or small_cholesky or tiny or mixed"
