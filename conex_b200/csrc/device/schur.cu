// K1 + K2: Schur-complement assembly for one dense LMI block, H_ij = tr(A_i W A_j W), entirely
// on the FP64 tensor cores.
//
// Reference (dense_lmi_constraint.cc:62-103): per constraint two n x n x n GEMMs (A_i W, W A_i W)
// followed by a GEMV against the growing slab Avect[:, 0:i+1] — BLAS-2 and memory bound. Here:
//   K1  T_i = A_i W          one strided-batched DMMA GEMM per panel of constraints
//       B_i = W T_i          strided-batched, lower tiles only + mirrored store (B_i symmetric)
//   K2  Haug = Bmat^T Aall   one lower-trapezoid DMMA GEMM with K = n^2 (SYRK-style Gram)
// The affine term C rides along as matrix m of Aall and W as row m+1 of Bmat, so the same Gram
// launch also yields AQc_j = <W C W, A_j>, AW_j = <W, A_j>, <c,Qc> and <w,c> (rows m, m+1 of Haug).
#include "common.cuh"
#include "device_api.h"

extern "C" int cxb_schur_dense_lmi(void* stream, int n, int m, const double* dAall, const double* dW,
                                   double* dB, double* dT, int panel, double* dHaug, long ldh) {
  using namespace cxb;
  cudaStream_t s = AsStream(stream);
  if (n < 1 || m < 1 || panel < 1 || ldh < m + 2) return -1;
  const long nn = (long)n * n;
  const int total = m + 1;  // A_0..A_{m-1}, C
  for (int p0 = 0; p0 < total; p0 += panel) {
    const int pb = (panel < total - p0) ? panel : (total - p0);
    int rc = Dgemm(s, false, false, n, n, n, 1.0, dAall + (long)p0 * nn, n, nn, dW, n, 0, 0.0, dT, n,
                   nn, pb, false);
    if (rc) return rc;
    // B_i = W T_i is symmetric in exact arithmetic: compute the lower tiles only and mirror them
    // (3 n^3 instead of 4 n^3 flops per constraint, and B_i exactly symmetric).
    rc = DgemmEx(s, -1, 1, false, false, n, n, n, 1.0, dW, n, 0, dT, n, nn, 0.0, dB + (long)p0 * nn, n,
                 nn, pb, true, true);
    if (rc) return rc;
  }
  cudaMemcpyAsync(dB + (long)(m + 1) * nn, dW, sizeof(double) * nn, cudaMemcpyDeviceToDevice, s);
  return Dgemm(s, true, false, m + 2, m + 1, (int)nn, 1.0, dB, nn, 0, dAall, nn, 0, 0.0, dHaug, ldh, 0,
               1, true);
}

// Same result without ever materialising all of B: one row panel of scaled matrices at a time is
// formed and immediately contracted against the constraint matrices to its left
// (Haug[p0 + q, c] = <B_q, A_c>, c <= p0 + q). dBp: (panel + 1) * n * n doubles, dT: panel * n * n.
// Memory: A + two panels instead of 2 A — what lets config 5 (n = 1000, m = 20000, A = 160 GB) run on
// one 192 GB B200. Cost: A is re-read once per row panel ((m / panel) / 2 passes over A in total).
extern "C" int cxb_schur_dense_lmi_streamed(void* stream, int n, int m, const double* dAall,
                                            const double* dW, double* dBp, double* dT, int panel,
                                            double* dHaug, long ldh) {
  using namespace cxb;
  cudaStream_t s = AsStream(stream);
  if (n < 1 || m < 1 || panel < 1 || ldh < m + 2) return -1;
  const long nn = (long)n * n;
  const int total = m + 1;  // A_0..A_{m-1}, C
  for (int p0 = 0; p0 < total; p0 += panel) {
    const int pb = (panel < total - p0) ? panel : (total - p0);
    int rc = Dgemm(s, false, false, n, n, n, 1.0, dAall + (long)p0 * nn, n, nn, dW, n, 0, 0.0, dT, n,
                   nn, pb, false);
    if (rc) return rc;
    rc = DgemmEx(s, -1, 1, false, false, n, n, n, 1.0, dW, n, 0, dT, n, nn, 0.0, dBp, n, nn, pb, true,
                 true);
    if (rc) return rc;
    int rows = pb;
    if (p0 + pb == total) {  // last panel: W rides along as row m + 1 (AW_j = <W, A_j>, <w,c>)
      cudaMemcpyAsync(dBp + (long)pb * nn, dW, sizeof(double) * nn, cudaMemcpyDeviceToDevice, s);
      rows = pb + 1;
    }
    const int cols = (p0 + rows < total) ? (p0 + rows) : total;
    rc = DgemmEx(s, -1, 0, true, false, rows, cols, (int)nn, 1.0, dBp, nn, 0, dAall, nn, 0, 0.0,
                 dHaug + p0, ldh, 0, 1, true, false, p0);
    if (rc) return rc;
  }
  return LaunchStatus();
}
