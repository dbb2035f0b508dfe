// Entry-sparse LMI operators (SURVEY.md §8d "structure caveat", §8f-1): when the constraint matrices
// A_i have only a few non-zero entries — MaxCut (A_i = -e_i e_i^T), Lovasz theta (A_e = E_ij + E_ji) —
// the Schur complement needs no matrix products at all:
//   H_ij  = tr(A_i W A_j W) = sum_{(p,q) in A_i} sum_{(r,s) in A_j} a_pq a_rs W[q,r] W[s,p]
//   AW_i  = tr(A_i W)       = sum_{(p,q) in A_i} a_pq W[q,p]
//   AQc_i = <A_i, W C W>    = sum_{(p,q) in A_i} a_pq (W C W)[q,p]
// i.e. O((sum_i nnz_i)^2) gathers from W instead of 4 m n^3 + m^2 n^2 flops, and the operator takes
// O(nnz) memory instead of m n^2 doubles. The reference has no such path (it stores dense matrices,
// hermitian_psd.cc:248-275); the arithmetic is the same sum in a different order.
//
// Storage (built by the host from the CONEX_UpdateLinearOperator stream):
//   offsets[m+1], rows[], cols[], vals[]   entries of A_i, both triangles listed explicitly
//   pos_ptr[npos+1], pos_index[npos], pos_var[], pos_val[]   the same entries grouped by matrix position
//   (column-major index c*n + r), for the deterministic slack accumulation.
#include "common.cuh"
#include "device_api.h"

namespace cxb {
namespace {

// One warp per pair (i >= j) of the augmented lower trapezoid: rows 0..m-1 = H, row m = AQc, row m+1 = AW.
__global__ void __launch_bounds__(256) SparseSchurKernel(int n, int m, const int* __restrict__ offsets,
                                                         const int* __restrict__ rows,
                                                         const int* __restrict__ cols,
                                                         const double* __restrict__ vals,
                                                         const double* __restrict__ W,
                                                         const double* __restrict__ WCW, double* Haug,
                                                         long ldh) {
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int j = blockIdx.y; j < m; j += gridDim.y) {
  const int i = j + (int)warp;  // rows j .. m+1 of column j
  if (i > m + 1) continue;
  const int jb = offsets[j], je = offsets[j + 1];
  double s = 0;
  if (i < m) {
    const int ib = offsets[i], ie = offsets[i + 1];
    const long ni = ie - ib, nj = je - jb;
    for (long t = lane; t < ni * nj; t += 32) {
      const int e = ib + (int)(t / nj), f = jb + (int)(t % nj);
      const int p = rows[e], q = cols[e], r = rows[f], sc = cols[f];
      s += vals[e] * vals[f] * W[(long)r * n + q] * W[(long)p * n + sc];
    }
  } else {
    const double* M = (i == m) ? WCW : W;  // AQc_j = <A_j, W C W>, AW_j = <A_j, W>
    for (int f = jb + lane; f < je; f += 32) s += vals[f] * M[(long)rows[f] * n + cols[f]];
  }
  s = WarpSum(s);
  if (lane == 0) Haug[(long)j * ldh + i] = s;
  }
}

// out = -k C, then out[pos] += sum over the entries at pos of y[var] * val (fixed order)
__global__ void SparseSlackInitKernel(long nn, double k, const double* __restrict__ C, double* out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nn) out[e] = -k * C[e];
}
__global__ void SparseSlackAddKernel(int npos, const int* __restrict__ pos_ptr,
                                     const long* __restrict__ pos_index, const int* __restrict__ pos_var,
                                     const double* __restrict__ pos_val, const double* __restrict__ y,
                                     double* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npos) return;
  double s = 0;
  for (int e = pos_ptr[t]; e < pos_ptr[t + 1]; e++) s += y[pos_var[e]] * pos_val[e];
  out[pos_index[t]] += s;
}

}  // namespace
}  // namespace cxb

using namespace cxb;

extern "C" {

int cxb_sparse_lmi_schur(void* stream, int n, int m, const int* d_offsets, const int* d_rows,
                         const int* d_cols, const double* d_vals, const double* dC, const double* dW,
                         double* d_work, double* dHaug, long ldh) {
  cudaStream_t s = AsStream(stream);
  if (n < 1 || m < 1 || ldh < m + 2) return -1;
  const long nn = (long)n * n;
  double* CW = d_work;         // C W
  double* WCW = d_work + nn;   // W C W
  int rc = Dgemm(s, false, false, n, n, n, 1.0, dC, n, 0, dW, n, 0, 0.0, CW, n, 0, 1, false);
  if (rc) return rc;
  if ((rc = Dgemm(s, false, false, n, n, n, 1.0, dW, n, 0, CW, n, 0, 0.0, WCW, n, 0, 1, false))) return rc;
  // one warp per entry of the lower trapezoid, column j = blockIdx.y
  const int warps_per_block = 8;
  dim3 grid((m + 2 + warps_per_block - 1) / warps_per_block, m < 65535 ? m : 65535);
  CountLaunch(); SparseSchurKernel<<<grid, 32 * warps_per_block, 0, s>>>(n, m, d_offsets, d_rows, d_cols, d_vals, dW,
                                                              WCW, dHaug, ldh);
  // column m: <c,Qc> = <C, W C W> (row m), <w,c> = <C, W> (row m + 1)
  if ((rc = cxb_dot(stream, nn, dC, WCW, dHaug + (long)m * ldh + m))) return rc;
  if ((rc = cxb_dot(stream, nn, dC, dW, dHaug + (long)m * ldh + m + 1))) return rc;
  return LaunchStatus();
}

int cxb_sparse_lmi_slack(void* stream, int n, int npos, const int* d_pos_ptr, const long* d_pos_index,
                         const int* d_pos_var, const double* d_pos_val, const double* dC, const double* dy,
                         double k, double* d_out) {
  cudaStream_t s = AsStream(stream);
  const long nn = (long)n * n;
  CountLaunch(); SparseSlackInitKernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(nn, k, dC, d_out);
  if (npos > 0) {
    CountLaunch(); SparseSlackAddKernel<<<(npos + 255) / 256, 256, 0, s>>>(npos, d_pos_ptr, d_pos_index, d_pos_var,
                                                                 d_pos_val, dy, d_out);
  }
  return LaunchStatus();
}

}  // extern "C"
