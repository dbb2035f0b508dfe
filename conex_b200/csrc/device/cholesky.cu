// K3 / K5: blocked right-looking Cholesky of the dense Schur complement and the triangular solves.
//
// Replaces BlockCholeskyInPlace / ApplyBlockInverse(OfTranspose)InPlace
// (block_triangular_operations.cc:114-219) for one dense supernode. Structure per NB-wide panel:
//   1. PotrfDiagKernel   — one CTA factors the NB x NB diagonal block in shared memory
//                          (warp-shuffle free: column scale + rank-1 update per step);
//   2. TrsmPanelKernel   — CTAs of 32 rows solve X L11^T = A21 by true substitution in shared
//                          memory (backward stable; H can be very ill conditioned late in the IPM,
//                          so no explicit inverses are used);
//   3. DMMA SYRK update  — A22 -= L21 L21^T on the lower tiles only (gemm.cu, tensor cores).
// A non-positive pivot sets *info = 1 + column (Eigen::LLT::info() != Success in the reference).
#include "common.cuh"
#include "device_api.h"

namespace cxb {
namespace {

constexpr int kNB = 128;      // panel width
constexpr int kTrsmRows = 32;  // rows per TRSM CTA

// Factor the nb x nb block at H (ld) in place; lower triangle. One CTA, blockDim = 1024.
__global__ void __launch_bounds__(1024) PotrfDiagKernel(int nb, double* H, long ld, int col0,
                                                        int* info) {
  extern __shared__ double s[];  // nb x (nb+1), column-major with pitch nb+1
  const int P = nb + 1;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (*info != 0) return;  // an earlier panel already failed
  for (int e = tid; e < nb * nb; e += nt) {
    const int r = e % nb, c = e / nb;
    if (r >= c) s[c * P + r] = H[(long)c * ld + r];
  }
  __syncthreads();
  __shared__ int failed;
  if (tid == 0) failed = 0;
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    const double d = s[j * P + j];
    if (!(d > 0.0)) {
      if (tid == 0) {
        failed = 1;
        *info = col0 + j + 1;
      }
    }
    __syncthreads();
    if (failed) break;
    const double rd = sqrt(d);
    const double inv = 1.0 / rd;
    // scale column j
    for (int r = j + tid; r < nb; r += nt) {
      s[j * P + r] = (r == j) ? rd : s[j * P + r] * inv;
    }
    __syncthreads();
    // rank-1 update of the trailing lower triangle: s[r][c] -= l[r] * l[c], c > j, r >= c
    const int rem = nb - j - 1;
    const int total = rem * rem;
    for (int e = tid; e < total; e += nt) {
      const int r = j + 1 + e % rem, c = j + 1 + e / rem;
      if (r >= c) s[c * P + r] -= s[j * P + r] * s[j * P + c];
    }
    __syncthreads();
  }
  for (int e = tid; e < nb * nb; e += nt) {
    const int r = e % nb, c = e / nb;
    if (r >= c) H[(long)c * ld + r] = s[c * P + r];
  }
}

// Solves X * L11^T = A21 for a strip of kTrsmRows rows. L11: nb x nb lower (ld), A21 strip at
// A + row0 (ld), overwritten by X. blockDim = 256.
__global__ void __launch_bounds__(256) TrsmPanelKernel(int nb, const double* __restrict__ L11,
                                                       long ld, double* A21, int rows,
                                                       const int* info) {
  extern __shared__ double s[];
  if (*info != 0) return;
  const int PL = nb + 1;
  double* sl = s;                 // nb x nb lower, sl[c*PL + r] = L11[r][c]
  double* sx = s + nb * PL;       // strip: sx[c*(kTrsmRows+1) + r]
  constexpr int PX = kTrsmRows + 1;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * kTrsmRows;
  const int nrows = min(kTrsmRows, rows - row0);
  for (int e = tid; e < nb * nb; e += 256) {
    const int r = e % nb, c = e / nb;
    sl[c * PL + r] = (r >= c) ? L11[(long)c * ld + r] : 0.0;
  }
  for (int e = tid; e < nb * kTrsmRows; e += 256) {
    const int r = e % kTrsmRows, c = e / kTrsmRows;
    sx[c * PX + r] = (r < nrows) ? A21[(long)c * ld + row0 + r] : 0.0;
  }
  __syncthreads();
  const int r = tid % kTrsmRows;  // my row
  const int q = tid / kTrsmRows;  // column phase 0..7
  for (int j = 0; j < nb; j++) {
    // x[:, j] = a[:, j] / L[j][j]
    if (q == 0) sx[j * PX + r] /= sl[j * PL + j];
    __syncthreads();
    const double xj = sx[j * PX + r];
    // a[:, c] -= x[:, j] * L[c][j] for c > j; columns dealt round-robin over the 8 phases
    for (int c = j + 1 + q; c < nb; c += 8) sx[c * PX + r] -= xj * sl[j * PL + c];
    __syncthreads();
  }
  for (int e = tid; e < nb * kTrsmRows; e += 256) {
    const int rr = e % kTrsmRows, c = e / kTrsmRows;
    if (rr < nrows) A21[(long)c * ld + row0 + rr] = sx[c * PX + rr];
  }
}

// ---- triangular solves -----------------------------------------------------------------------
// Forward: solve L_kk x_k = b_k for one diagonal block and all right-hand sides. One CTA.
template <bool TRANS>
__global__ void __launch_bounds__(256) TrsvDiagKernel(int nb, const double* __restrict__ L, long ld,
                                                      double* X, long ldx, int nrhs) {
  extern __shared__ double s[];
  const int PL = nb + 1;
  double* sl = s;
  double* sx = s + nb * PL;  // nb x nrhs
  const int tid = threadIdx.x;
  for (int e = tid; e < nb * nb; e += 256) {
    const int r = e % nb, c = e / nb;
    if (r >= c) sl[c * PL + r] = L[(long)c * ld + r];
  }
  for (int e = tid; e < nb * nrhs; e += 256) sx[e] = X[(long)(e / nb) * ldx + e % nb];
  __syncthreads();
  if (!TRANS) {
    for (int j = 0; j < nb; j++) {
      if (tid < nrhs) sx[tid * nb + j] /= sl[j * PL + j];
      __syncthreads();
      for (int e = tid; e < (nb - j - 1) * nrhs; e += 256) {
        const int r = j + 1 + e % (nb - j - 1), k = e / (nb - j - 1);
        sx[k * nb + r] -= sl[j * PL + r] * sx[k * nb + j];
      }
      __syncthreads();
    }
  } else {
    for (int j = nb - 1; j >= 0; j--) {
      if (tid < nrhs) sx[tid * nb + j] /= sl[j * PL + j];
      __syncthreads();
      // x[r] -= L[j][r] * x[j] for r < j   (L^T[r][j] = L[j][r])
      for (int e = tid; e < j * nrhs; e += 256) {
        const int r = e % j, k = e / j;
        sx[k * nb + r] -= sl[r * PL + j] * sx[k * nb + j];
      }
      __syncthreads();
    }
  }
  for (int e = tid; e < nb * nrhs; e += 256) X[(long)(e / nb) * ldx + e % nb] = sx[e];
}

// Forward update: x[rows below] -= L[rows, kblock] * x_k. One thread per row; coalesced over rows.
__global__ void __launch_bounds__(256) TrsvUpdateFwdKernel(int rows, int nb,
                                                           const double* __restrict__ Lblk, long ld,
                                                           const double* __restrict__ xk,
                                                           double* xrest, long ldx, int nrhs) {
  __shared__ double sx[4 * kNB];
  for (int e = threadIdx.x; e < nb * nrhs; e += 256) sx[e] = xk[(long)(e / nb) * ldx + e % nb];
  __syncthreads();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  double acc[4] = {0, 0, 0, 0};
  for (int c = 0; c < nb; c++) {
    const double l = Lblk[(long)c * ld + r];
    for (int k = 0; k < nrhs; k++) acc[k] += l * sx[k * nb + c];
  }
  for (int k = 0; k < nrhs; k++) xrest[(long)k * ldx + r] -= acc[k];
}

// Backward update: x[0:cols] -= L[kblock rows, 0:cols]^T * x_k. One warp per column.
__global__ void __launch_bounds__(256) TrsvUpdateBwdKernel(int cols, int nb,
                                                           const double* __restrict__ Lrow, long ld,
                                                           const double* __restrict__ xk, double* x,
                                                           long ldx, int nrhs) {
  __shared__ double sx[4 * kNB];
  for (int e = threadIdx.x; e < nb * nrhs; e += 256) sx[e] = xk[(long)(e / nb) * ldx + e % nb];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= cols) return;
  double acc[4] = {0, 0, 0, 0};
  for (int r = lane; r < nb; r += 32) {
    const double l = Lrow[(long)c * ld + r];
    for (int k = 0; k < nrhs; k++) acc[k] += l * sx[k * nb + r];
  }
  for (int k = 0; k < nrhs; k++) {
    const double v = WarpSum(acc[k]);
    if (lane == 0) x[(long)k * ldx + c] -= v;
  }
}

__global__ void ResetInfoKernel(int* info) { *info = 0; }

}  // namespace
}  // namespace cxb

using namespace cxb;

extern "C" {

size_t cxb_potrf_worksize(int m) {
  (void)m;
  return 1;  // the current algorithm works fully in place
}

int cxb_potrf_lower(void* stream, int m, double* dH, long ldh, double* d_work, int* d_info) {
  (void)d_work;
  cudaStream_t s = AsStream(stream);
  if (m <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(PotrfDiagKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(double) * kNB * (kNB + 1)));
    cudaFuncSetAttribute(TrsmPanelKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(double) * (kNB * (kNB + 1) + kNB * (kTrsmRows + 1))));
    configured = true;
  }
  CountLaunch(); ResetInfoKernel<<<1, 1, 0, s>>>(d_info);
  for (int j0 = 0; j0 < m; j0 += kNB) {
    const int nb = min(kNB, m - j0);
    double* Hjj = dH + (long)j0 * ldh + j0;
    CountLaunch(); PotrfDiagKernel<<<1, 1024, sizeof(double) * nb * (nb + 1), s>>>(nb, Hjj, ldh, j0, d_info);
    const int rows = m - j0 - nb;
    if (rows > 0) {
      double* A21 = Hjj + nb;
      const size_t smem = sizeof(double) * ((size_t)nb * (nb + 1) + (size_t)nb * (kTrsmRows + 1));
      CountLaunch(); TrsmPanelKernel<<<(rows + kTrsmRows - 1) / kTrsmRows, 256, smem, s>>>(nb, Hjj, ldh, A21, rows,
                                                                            d_info);
      double* A22 = dH + (long)(j0 + nb) * ldh + (j0 + nb);
      // A22 -= L21 L21^T (lower tiles only)
      const int rc = Dgemm(s, false, true, rows, rows, nb, -1.0, A21, ldh, 0, A21, ldh, 0, 1.0, A22,
                           ldh, 0, 1, true);
      if (rc != 0) return rc;
    }
  }
  return LaunchStatus();
}

int cxb_potrs_lower(void* stream, int m, const double* dL, long ldl, double* dX, long ldx, int nrhs) {
  cudaStream_t s = AsStream(stream);
  if (m <= 0 || nrhs <= 0) return 0;
  if (nrhs > 4) return -1;
  static bool configured = false;
  if (!configured) {
    const int bytes = (int)(sizeof(double) * (kNB * (kNB + 1) + 4 * kNB));
    cudaFuncSetAttribute(TrsvDiagKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(TrsvDiagKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    configured = true;
  }
  const int nblk = (m + kNB - 1) / kNB;
  // forward: L z = b
  for (int k = 0; k < nblk; k++) {
    const int j0 = k * kNB, nb = min(kNB, m - j0);
    const double* Lkk = dL + (long)j0 * ldl + j0;
    const size_t smem = sizeof(double) * ((size_t)nb * (nb + 1) + (size_t)nb * nrhs);
    CountLaunch(); TrsvDiagKernel<false><<<1, 256, smem, s>>>(nb, Lkk, ldl, dX + j0, ldx, nrhs);
    const int rows = m - j0 - nb;
    if (rows > 0) {
      CountLaunch(); TrsvUpdateFwdKernel<<<(rows + 255) / 256, 256, 0, s>>>(rows, nb, Lkk + nb, ldl, dX + j0,
                                                             dX + j0 + nb, ldx, nrhs);
    }
  }
  // backward: L^T x = z
  for (int k = nblk - 1; k >= 0; k--) {
    const int j0 = k * kNB, nb = min(kNB, m - j0);
    const double* Lkk = dL + (long)j0 * ldl + j0;
    const size_t smem = sizeof(double) * ((size_t)nb * (nb + 1) + (size_t)nb * nrhs);
    CountLaunch(); TrsvDiagKernel<true><<<1, 256, smem, s>>>(nb, Lkk, ldl, dX + j0, ldx, nrhs);
    if (j0 > 0) {
      CountLaunch(); TrsvUpdateBwdKernel<<<(j0 + 7) / 8, 256, 0, s>>>(j0, nb, dL + j0, ldl, dX + j0, dX, ldx, nrhs);
    }
  }
  return LaunchStatus();
}

}  // extern "C"
