// A CTA seen as a "team": the small-cone kernels (small_cone_math.cuh) are written as sequences of
// team-wide phases — parallel loops, reductions and single-thread sections — each ending in a
// barrier, so that one problem (one cone of one program of a batch) is handled by one CTA with all
// intermediate data in shared memory. The same phase structure can be executed serially by a host
// stand-in of this class, which is how tests/emul checks the index arithmetic of the math header on
// the CPU; the product only ever instantiates DeviceTeam.
#pragma once
#include "common.cuh"

namespace cxb {

// The team interface on ONE WARP: every phase boundary is a __syncwarp() and the reductions are shuffles, so a
// small problem whose phases are short (Lanczos on a 20 x 20 block: ~10 steps of two 20-entry matvecs and two dot
// products; a 40 x 40 Cholesky: 40 columns) pays a few cycles per phase instead of a 128-thread barrier, and four
// problems share a CTA. Reductions combine in a fixed order (shfl_down tree, result broadcast from lane 0).
struct WarpTeam {
  int tid;
  static constexpr unsigned kFull = 0xffffffffu;

  __device__ WarpTeam() : tid(threadIdx.x & 31) {}

  __device__ __forceinline__ int size() const { return 32; }

  template <class F>
  __device__ __forceinline__ void par(int n, F f) {
    for (int i = tid; i < n; i += 32) f(i);
    __syncwarp();
  }
  // f(i, c, row_value(i)) for every (i, c) in rows x cols: rows on lanes, columns in sequence. row_value must only read
  // data that f does not write.
  template <class R, class F>
  __device__ __forceinline__ void par2(int rows, int cols, R row_value, F f) {
    for (int i = tid; i < rows; i += 32) {
      const double rv = row_value(i);
      for (int c = 0; c < cols; c++) f(i, c, rv);
    }
    __syncwarp();
  }
  template <class F>
  __device__ __forceinline__ double sum(int n, F f) {
    double v = 0;
    for (int i = tid; i < n; i += 32) v += f(i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    v = __shfl_sync(kFull, v, 0);
    __syncwarp();
    return v;
  }
  template <class F>
  __device__ __forceinline__ double maxv(int n, F f) {
    double v = -1.7976931348623157e308;
    for (int i = tid; i < n; i += 32) v = fmax(v, f(i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    __syncwarp();
    return v;
  }
  template <class F>
  __device__ __forceinline__ double minv(int n, F f) {
    double v = 1.7976931348623157e308;
    for (int i = tid; i < n; i += 32) v = fmin(v, f(i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    __syncwarp();
    return v;
  }
  // index of the FIRST largest f(i), i in [0, n) (n >= 1), and that value
  template <class F>
  __device__ __forceinline__ int argmax_first(int n, F f, double* vmax) {
    double v = -1.7976931348623157e308;
    int idx = 0;  // stays valid when every f(i) is NaN
    for (int i = tid; i < n; i += 32) {
      const double x = f(i);
      if (x > v) {
        v = x;
        idx = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(kFull, v, o);
      const int oi = __shfl_xor_sync(kFull, idx, o);
      if (ov > v || (ov == v && oi < idx)) {
        v = ov;
        idx = oi;
      }
    }
    __syncwarp();
    *vmax = v;
    return idx;
  }
  // smallest i in [0, n) with pred(i), or n
  template <class F>
  __device__ __forceinline__ int first_true(int n, F pred) {
    int found = n;
    for (int base = 0; base < n; base += 32) {
      const int i = base + tid;
      const bool p = (i < n) && pred(i);
      const unsigned mask = __ballot_sync(kFull, p);
      if (mask) {
        found = base + __ffs(mask) - 1;
        break;
      }
    }
    __syncwarp();
    return found;
  }
  template <class F>
  __device__ __forceinline__ double bcast(F f) {
    __syncwarp();
    double v = 0;
    if (tid == 0) v = f();
    v = __shfl_sync(kFull, v, 0);
    __syncwarp();
    return v;
  }
  template <class F>
  __device__ __forceinline__ void single(F f) {
    __syncwarp();
    if (tid == 0) f();
    __syncwarp();
  }
  // a section that only has warp-sized parallelism (see DeviceTeam::warp0): here it is simply the team itself
  template <class F>
  __device__ __forceinline__ void warp0(F f) {
    f(*this);
  }
};


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
  // f(i, c, row_value(i)) for every (i, c) in rows x cols: rows on lanes, columns on warps (no index division, and a
  // per-row value is computed once per thread and row). row_value must only read data that f does not write.
  template <class R, class F>
  __device__ __forceinline__ void par2(int rows, int cols, R row_value, F f) {
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nt + 31) >> 5;
    for (int i = lane; i < rows; i += 32) {
      const double rv = row_value(i);
      for (int c = warp; c < cols; c += nwarps) f(i, c, rv);
    }
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
  // index of the FIRST largest f(i), i in [0, n) (n >= 1), and that value. Two barriers; at most 16 warps.
  template <class F>
  __device__ __forceinline__ int argmax_first(int n, F f, double* vmax) {
    if (n <= 32) {
      // all candidates sit on warp 0: the other warps only wait for its result
      __syncthreads();
      if ((tid >> 5) == 0) {
        WarpTeam w;
        double v;
        const int idx = w.argmax_first(n, f, &v);
        if (tid == 0) {
          red[0] = v;
          red[16] = (double)idx;
        }
      }
      __syncthreads();
      *vmax = red[0];
      return (int)red[16];
    }
    double v = -1.7976931348623157e308;
    int idx = 0;  // stays valid when every f(i) is NaN
    for (int i = tid; i < n; i += nt) {
      const double x = f(i);
      if (x > v) {
        v = x;
        idx = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) {
        v = ov;
        idx = oi;
      }
    }
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (nt + 31) >> 5;
    __syncthreads();
    if (lane == 0) {
      red[warp] = v;
      red[16 + warp] = (double)idx;
    }
    __syncthreads();
    v = red[0];
    idx = (int)red[16];
    for (int w = 1; w < nwarps; w++) {
      const double ov = red[w];
      const int oi = (int)red[16 + w];
      if (ov > v || (ov == v && oi < idx)) {
        v = ov;
        idx = oi;
      }
    }
    *vmax = v;
    return idx;
  }
  // smallest i in [0, n) with pred(i), or n
  template <class F>
  __device__ __forceinline__ int first_true(int n, F pred) {
    double v = -1.7976931348623157e308;
    for (int i = tid; i < n; i += nt)
      if (pred(i)) {
        v = -(double)i;
        break;  // this thread's later candidates are larger
      }
    v = BlockMax(v);
    return v < -1e300 ? n : (int)(-v);
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
  // A section whose phases never have more than a warp's worth of parallel work (the Lanczos recurrence on a small
  // block: 20-entry matvecs and dot products, then the multi-section of the Ritz values): warp 0 runs it as a
  // WarpTeam — __syncwarp and shuffles between its ~200 phases instead of block barriers — the other warps wait once.
  template <class F>
  __device__ __forceinline__ void warp0(F f) {
    __syncthreads();
    if ((tid >> 5) == 0) {
      WarpTeam w;
      f(w);
    }
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
