// A CTA seen as a "team": the small-cone kernels (small_cone_math.cuh) are written as sequences of
// team-wide phases — parallel loops, reductions and single-thread sections — each ending in a
// barrier, so that one problem (one cone of one program of a batch) is handled by one CTA with all
// intermediate data in shared memory. The same phase structure can be executed serially by a host
// stand-in of this class, which is how tests/emul checks the index arithmetic of the math header on
// the CPU; the product only ever instantiates DeviceTeam.
#pragma once
#include "common.cuh"

namespace cxb {

struct DeviceTeam {
  int tid, nt;
  double* red;  // >= 40 doubles of shared memory

  __device__ DeviceTeam(double* scratch) : tid(threadIdx.x), nt(blockDim.x), red(scratch) {}

  __device__ __forceinline__ int size() const { return nt; }

  template <class F>
  __device__ __forceinline__ void par(int n, F f) {
    for (int i = tid; i < n; i += nt) f(i);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ double sum(int n, F f) {
    double v = 0;
    for (int i = tid; i < n; i += nt) v += f(i);
    return BlockSum(v, red);
  }
  template <class F>
  __device__ __forceinline__ double maxv(int n, F f) {
    double v = -1.7976931348623157e308;
    for (int i = tid; i < n; i += nt) v = fmax(v, f(i));
    return BlockMax(v);
  }
  template <class F>
  __device__ __forceinline__ double minv(int n, F f) {
    double v = -1.7976931348623157e308;
    for (int i = tid; i < n; i += nt) v = fmax(v, -f(i));
    return -BlockMax(v);
  }
  // f() evaluated by one thread, result given to all.
  template <class F>
  __device__ __forceinline__ double bcast(F f) {
    __syncthreads();
    if (tid == 0) red[34] = f();
    __syncthreads();
    const double v = red[34];
    __syncthreads();
    return v;
  }
  template <class F>
  __device__ __forceinline__ void single(F f) {
    __syncthreads();
    if (tid == 0) f();
    __syncthreads();
  }

 private:
  __device__ __forceinline__ double BlockMax(double v) {
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (nt + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double t = (lane < nwarps) ? red[lane] : -1.7976931348623157e308;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
      if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
  }
};

}  // namespace cxb
