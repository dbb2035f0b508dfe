// Bandwidth-bound helpers of the Newton step: slack GEMV (K6), trace/Frobenius reductions (K7),
// symmetrisation and the small vector updates of the host loop. All reductions are deterministic
// (fixed partition + fixed-order final sum by the last block to arrive), so that W stays
// bit-identical across replicas.
#include "common.cuh"
#include "device_api.h"

namespace cxb {
namespace {

constexpr int kMaxPartials = 2048;
// Scratch for two-stage reductions; one set per concurrent reduction "slot".
__device__ double g_partials[4][kMaxPartials];
__device__ unsigned int g_tickets[4];

// Block-level partial `v` -> deterministic grid sum written to out[0] by the last block.
__device__ void GridSumFinalize(double v, int slot, double* out, double* scratch) {
  v = BlockSum(v, scratch);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    g_partials[slot][blockIdx.x] = v;
    __threadfence();
    const unsigned int t = atomicAdd(&g_tickets[slot], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += g_partials[slot][i];
    // fixed order: each thread sums a fixed strided subset, then a fixed tree
    s = BlockSum(s, scratch);
    if (threadIdx.x == 0) {
      out[0] = s;
      g_tickets[slot] = 0;
    }
  }
}

__global__ void SetIdentityKernel(int n, double* W) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)n * n;
  if (idx < total) W[idx] = (idx % n == idx / n) ? 1.0 : 0.0;
}

__global__ void DotKernel(long n, const double* __restrict__ x, const double* __restrict__ y,
                          double* out) {
  __shared__ double scratch[33];
  double s = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    s += x[i] * y[i];
  GridSumFinalize(s, 0, out, scratch);
}

__global__ void AxpbypczKernel(long n, double a, const double* __restrict__ x, double b, double* y,
                               double c, const double* __restrict__ z) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = a * x[i];
  if (b != 0.0) v += b * y[i];
  if (z != nullptr) v += c * z[i];
  y[i] = v;
}

__global__ void CopyStridedKernel(long n, const double* __restrict__ src, long incs, double* dst,
                                  long incd) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i * incd] = src[i * incs];
}

__global__ void ScatterAddLowerKernel(int mc, const double* __restrict__ G, long ldg,
                                      const int* __restrict__ idx, double* H, long ldh) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;  // row within the cone
  const int b = blockIdx.y;                             // col within the cone
  if (a >= mc || a < b) return;
  const int ga = idx[a], gb = idx[b];
  const int r = ga > gb ? ga : gb, c = ga > gb ? gb : ga;
  // distinct (a, b) pairs map to distinct (r, c) because idx has no repeats: no atomics needed.
  H[(long)c * ldh + r] += G[(long)b * ldg + a];
}

__global__ void FillKernel(long n, double v, double* dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

// dst[idx[i]] += src[i] (idx == nullptr: identity). idx has no repeats, so no atomics.
__global__ void ScatterAddVecKernel(int n, const double* __restrict__ src,
                                    const int* __restrict__ idx, double* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = idx ? idx[i] : i;
  dst[k] += src[i];
}

// out = K x for a symmetric K stored by its lower triangle (ld): thread i sums row i of the lower
// triangle (coalesced across the warp for every column) and, down column i, the mirrored part.
__global__ void __launch_bounds__(128) SymvLowerKernel(int N, const double* __restrict__ K, long ld,
                                                       const double* __restrict__ x, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = 0;
  for (int j = 0; j <= i; j++) s += K[(long)j * ld + i] * x[j];       // K(i, j), j <= i
  const double* col = K + (long)i * ld;
  for (int r = i + 1; r < N; r++) s += col[r] * x[r];                  // K(r, i) = K(i, r), r > i
  out[i] = s;
}

// dst[i] = src[idx[i]]
__global__ void GatherVecKernel(int n, const double* __restrict__ src, const int* __restrict__ idx,
                                double* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = src[idx ? idx[i] : i];
}

__global__ void AffineUpdateKernel(long total, double* W, const double* __restrict__ WSW, double w_e) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double w = W[i];
  if (w_e != 0.0) w *= (1.0 + w_e);
  W[i] = w + WSW[i];
}

// out[r] = sum_j A[r + j*nn] * coef[j]; two rows per thread when VEC == 2.
template <int VEC>
__global__ void __launch_bounds__(256) GemvNKernel(long nn, int cols, const double* __restrict__ A,
                                                   const double* __restrict__ coef, double* out) {
  constexpr int kChunk = 1024;
  __shared__ double sc[kChunk];
  const long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  double acc0 = 0, acc1 = 0;
  for (int j0 = 0; j0 < cols; j0 += kChunk) {
    const int jn = min(kChunk, cols - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < jn; j += blockDim.x) sc[j] = coef[j0 + j];
    __syncthreads();
    if (r < nn) {
      const double* p = A + (long)j0 * nn + r;
      int j = 0;
      if (VEC == 2) {
        for (; j + 4 <= jn; j += 4) {
          const double2 v0 = *reinterpret_cast<const double2*>(p + (long)(j + 0) * nn);
          const double2 v1 = *reinterpret_cast<const double2*>(p + (long)(j + 1) * nn);
          const double2 v2 = *reinterpret_cast<const double2*>(p + (long)(j + 2) * nn);
          const double2 v3 = *reinterpret_cast<const double2*>(p + (long)(j + 3) * nn);
          acc0 += v0.x * sc[j];
          acc1 += v0.y * sc[j];
          acc0 += v1.x * sc[j + 1];
          acc1 += v1.y * sc[j + 1];
          acc0 += v2.x * sc[j + 2];
          acc1 += v2.y * sc[j + 2];
          acc0 += v3.x * sc[j + 3];
          acc1 += v3.y * sc[j + 3];
        }
        for (; j < jn; j++) {
          const double2 v = *reinterpret_cast<const double2*>(p + (long)j * nn);
          acc0 += v.x * sc[j];
          acc1 += v.y * sc[j];
        }
      } else {
        for (; j + 4 <= jn; j += 4) {
          const double v0 = p[(long)(j + 0) * nn], v1 = p[(long)(j + 1) * nn];
          const double v2 = p[(long)(j + 2) * nn], v3 = p[(long)(j + 3) * nn];
          acc0 += v0 * sc[j];
          acc0 += v1 * sc[j + 1];
          acc0 += v2 * sc[j + 2];
          acc0 += v3 * sc[j + 3];
        }
        for (; j < jn; j++) acc0 += p[(long)j * nn] * sc[j];
      }
    }
  }
  if (r < nn) {
    out[r] = acc0;
    if (VEC == 2) out[r + 1] = acc1;
  }
}

// out[0] = tr(X), out[2] = argmax diag (first max), out[3] = max diag. Single block.
__global__ void DiagStatsKernel(int n, const double* __restrict__ X, double* out) {
  __shared__ double scratch[33];
  __shared__ double smax[32];
  __shared__ int sarg[32];
  double tr = 0, best = 0;
  int arg = -1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = X[(long)i * n + i];
    tr += d;
    if (arg < 0 || d > best) {
      best = d;
      arg = i;
    }
  }
  tr = BlockSum(tr, scratch);
  // argmax with ties -> smallest index (Eigen maxCoeff keeps the first maximum)
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) {
      best = ob;
      arg = oa;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    smax[warp] = best;
    sarg[warp] = arg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double b = smax[0];
    int a = sarg[0];
    for (int w = 1; w < nw; w++) {
      if (sarg[w] >= 0 && (a < 0 || smax[w] > b || (smax[w] == b && sarg[w] < a))) {
        b = smax[w];
        a = sarg[w];
      }
    }
    out[0] = tr;
    out[2] = (double)a;
    out[3] = b;
  }
}

// out[1] = sum_ij X_ij X_ji. Grid over 32x32 tile pairs (I >= J); off-diagonal pairs count twice.
__global__ void __launch_bounds__(256) TraceSquareKernel(int n, const double* __restrict__ X,
                                                         double* out) {
  __shared__ double ta[32][33];
  __shared__ double tb[32][33];
  __shared__ double scratch[33];
  const int nt = (n + 31) / 32;
  double s = 0;
  const long num_pairs = (long)nt * (nt + 1) / 2;
  for (long t = blockIdx.x; t < num_pairs; t += gridDim.x) {
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) I++;
    while ((long)I * (I + 1) / 2 > t) I--;
    const int J = (int)(t - (long)I * (I + 1) / 2);
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
      const int r = e & 31, c = e >> 5;
      const int gr = I * 32 + r, gc = J * 32 + c;
      ta[c][r] = (gr < n && gc < n) ? X[(long)gc * n + gr] : 0.0;  // tile (I,J): ta[c][r] = X[I r, J c]
      const int hr = J * 32 + r, hc = I * 32 + c;
      tb[c][r] = (hr < n && hc < n) ? X[(long)hc * n + hr] : 0.0;  // tile (J,I): tb[c][r] = X[J r, I c]
    }
    __syncthreads();
    double local = 0;
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
      const int r = e & 31, c = e >> 5;
      local += ta[c][r] * tb[r][c];  // X[I r, J c] * X[J c, I r]
    }
    s += (I == J) ? local : 2.0 * local;
  }
  GridSumFinalize(s, 1, out + 1, scratch);
}

// In place W <- (W + W^T)/2 (psd_constraint.cc:26-27). Grid over 32x32 tile pairs I >= J.
__global__ void __launch_bounds__(256) SymmetrizeKernel(int n, double* W) {
  __shared__ double ta[32][33];
  __shared__ double tb[32][33];
  const int nt = (n + 31) / 32;
  const long num_pairs = (long)nt * (nt + 1) / 2;
  for (long t = blockIdx.x; t < num_pairs; t += gridDim.x) {
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) I++;
    while ((long)I * (I + 1) / 2 > t) I--;
    const int J = (int)(t - (long)I * (I + 1) / 2);
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
      const int r = e & 31, c = e >> 5;
      const int gr = I * 32 + r, gc = J * 32 + c;
      ta[c][r] = (gr < n && gc < n) ? W[(long)gc * n + gr] : 0.0;
      const int hr = J * 32 + r, hc = I * 32 + c;
      tb[c][r] = (hr < n && hc < n) ? W[(long)hc * n + hr] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
      const int r = e & 31, c = e >> 5;
      const int gr = I * 32 + r, gc = J * 32 + c;
      if (gr < n && gc < n) W[(long)gc * n + gr] = (ta[c][r] + tb[r][c]) * 0.5;
      if (I != J) {
        const int hr = J * 32 + r, hc = I * 32 + c;
        if (hr < n && hc < n) W[(long)hc * n + hr] = (tb[c][r] + ta[r][c]) * 0.5;
      }
    }
  }
}

// X <- scale * (X + e I)  (psd_constraint.cc:19-22: the identity is scaled too)
__global__ void ShiftScaleKernel(int n, double* X, double e, double scale) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)n * n;
  if (idx >= total) return;
  double v = X[idx];
  if (idx % n == idx / n) v += e;
  X[idx] = v * scale;
}

// Y <- a*X + d*I   (element-wise, X may equal Y)
__global__ void ScaleAddDiagKernel(int n, const double* X, double a, double d, double* Y) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)n * n;
  if (idx >= total) return;
  double v = a * X[idx];
  if (idx % n == idx / n) v += d;
  Y[idx] = v;
}

// N <- V + U, D <- V - U
__global__ void SumDiffKernel(long total, const double* U, const double* V, double* N, double* D) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double u = U[i], v = V[i];
  N[i] = v + u;
  D[i] = v - u;
}

// XT = X^T (n x n), 32x32 tiles through shared memory.
__global__ void __launch_bounds__(256) TransposeKernel(int n, const double* __restrict__ X, double* XT) {
  __shared__ double t[32][33];
  const int I = blockIdx.x, J = blockIdx.y;
  for (int e = threadIdx.x; e < 1024; e += 256) {
    const int r = e & 31, c = e >> 5;
    const int gr = I * 32 + r, gc = J * 32 + c;
    t[c][r] = (gr < n && gc < n) ? X[(long)gc * n + gr] : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 1024; e += 256) {
    const int r = e & 31, c = e >> 5;
    const int gr = J * 32 + r, gc = I * 32 + c;  // XT[gr, gc] = X[gc, gr]
    if (gr < n && gc < n) XT[(long)gc * n + gr] = t[r][c];
  }
}

inline unsigned Blocks(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

int Transpose(cudaStream_t s, int n, const double* X, double* XT) {
  const int nt = (n + 31) / 32;
  CountLaunch(); TransposeKernel<<<dim3(nt, nt), 256, 0, s>>>(n, X, XT);
  return LaunchStatus();
}

int SetIdentity(cudaStream_t s, int n, double* W) {
  CountLaunch(); SetIdentityKernel<<<Blocks((long)n * n, 256), 256, 0, s>>>(n, W);
  return LaunchStatus();
}
int Symmetrize(cudaStream_t s, int n, double* W) {
  const int nt = (n + 31) / 32;
  const long pairs = (long)nt * (nt + 1) / 2;
  CountLaunch(); SymmetrizeKernel<<<(unsigned)min(pairs, (long)kNumSMs * 8), 256, 0, s>>>(n, W);
  return LaunchStatus();
}
int ShiftScale(cudaStream_t s, int n, double* X, double e, double scale) {
  CountLaunch(); ShiftScaleKernel<<<Blocks((long)n * n, 256), 256, 0, s>>>(n, X, e, scale);
  return LaunchStatus();
}
int ScaleAddDiag(cudaStream_t s, int n, const double* X, double a, double d, double* Y) {
  CountLaunch(); ScaleAddDiagKernel<<<Blocks((long)n * n, 256), 256, 0, s>>>(n, X, a, d, Y);
  return LaunchStatus();
}
int SumDiff(cudaStream_t s, long total, const double* U, const double* V, double* N, double* D) {
  CountLaunch(); SumDiffKernel<<<Blocks(total, 256), 256, 0, s>>>(total, U, V, N, D);
  return LaunchStatus();
}

}  // namespace cxb

using namespace cxb;

extern "C" {

int cxb_set_identity(void* stream, int n, double* dW) { return SetIdentity(AsStream(stream), n, dW); }

int cxb_dot(void* stream, long n, const double* dx, const double* dy, double* d_out) {
  const unsigned blocks = (unsigned)min((long)kMaxPartials, max(1L, (n + 1023) / 1024));
  CountLaunch(); DotKernel<<<min(blocks, (unsigned)(kNumSMs * 4)), 256, 0, AsStream(stream)>>>(n, dx, dy, d_out);
  return LaunchStatus();
}

int cxb_axpbypcz(void* stream, long n, double a, const double* dx, double b, double* dy, double c,
                 const double* dz) {
  if (n <= 0) return 0;
  CountLaunch(); AxpbypczKernel<<<Blocks(n, 256), 256, 0, AsStream(stream)>>>(n, a, dx, b, dy, c, dz);
  return LaunchStatus();
}

int cxb_copy_strided(void* stream, long n, const double* src, long incs, double* dst, long incd) {
  if (n <= 0) return 0;
  CountLaunch(); CopyStridedKernel<<<Blocks(n, 256), 256, 0, AsStream(stream)>>>(n, src, incs, dst, incd);
  return LaunchStatus();
}

int cxb_scatter_add_lower(void* stream, int mc, const double* dG, long ldg, const int* d_idx,
                          double* dH, long ldh) {
  if (mc <= 0) return 0;
  dim3 grid(Blocks(mc, 128), mc);
  CountLaunch(); ScatterAddLowerKernel<<<grid, 128, 0, AsStream(stream)>>>(mc, dG, ldg, d_idx, dH, ldh);
  return LaunchStatus();
}

int cxb_fill(void* stream, long n, double value, double* d_dst) {
  if (n <= 0) return 0;
  CountLaunch(); FillKernel<<<Blocks(n, 256), 256, 0, AsStream(stream)>>>(n, value, d_dst);
  return LaunchStatus();
}

int cxb_scatter_add_vec(void* stream, int n, const double* d_src, const int* d_idx, double* d_dst) {
  if (n <= 0) return 0;
  CountLaunch(); ScatterAddVecKernel<<<Blocks(n, 256), 256, 0, AsStream(stream)>>>(n, d_src, d_idx, d_dst);
  return LaunchStatus();
}

int cxb_gather_vec(void* stream, int n, const double* d_src, const int* d_idx, double* d_dst) {
  if (n <= 0) return 0;
  CountLaunch(); GatherVecKernel<<<Blocks(n, 256), 256, 0, AsStream(stream)>>>(n, d_src, d_idx, d_dst);
  return LaunchStatus();
}

int cxb_symv_lower(void* stream, int N, const double* dK, long ld, const double* dx, double* d_out) {
  if (N <= 0) return 0;
  CountLaunch(); SymvLowerKernel<<<Blocks(N, 128), 128, 0, AsStream(stream)>>>(N, dK, ld, dx, d_out);
  return LaunchStatus();
}

int cxb_affine_update(void* stream, int n, double* dW, const double* dWSW, double w_e) {
  const long total = (long)n * n;
  CountLaunch(); AffineUpdateKernel<<<Blocks(total, 256), 256, 0, AsStream(stream)>>>(total, dW, dWSW, w_e);
  return LaunchStatus();
}

int cxb_gemv_n(void* stream, long nn, int cols, const double* dAall, const double* d_coef,
               double* d_out) {
  if (nn <= 0) return 0;
  const bool vec2 = (nn % 2 == 0) && ((reinterpret_cast<uintptr_t>(dAall) & 15) == 0);
  if (vec2) {
    CountLaunch(); GemvNKernel<2><<<Blocks(nn / 2, 256), 256, 0, AsStream(stream)>>>(nn, cols, dAall, d_coef, d_out);
  } else {
    CountLaunch(); GemvNKernel<1><<<Blocks(nn, 256), 256, 0, AsStream(stream)>>>(nn, cols, dAall, d_coef, d_out);
  }
  return LaunchStatus();
}

int cxb_ws_reductions(void* stream, int n, const double* d_WS, double* d_out) {
  CountLaunch(); DiagStatsKernel<<<1, 1024, 0, AsStream(stream)>>>(n, d_WS, d_out);
  const int nt = (n + 31) / 32;
  const long pairs = (long)nt * (nt + 1) / 2;
  CountLaunch(); TraceSquareKernel<<<(unsigned)min(pairs, (long)kNumSMs * 4), 256, 0, AsStream(stream)>>>(n, d_WS,
                                                                                          d_out);
  return LaunchStatus();
}

}  // extern "C"
