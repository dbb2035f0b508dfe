/* conex-b200 device layer — the thin C ABI between the C++ host orchestration
 * (conex_b200/csrc/host) and the hand-written sm_100a CUDA kernels (conex_b200/csrc/device).
 *
 * Conventions
 *  - every pointer named d_* / marked "device" is a plain device pointer (cudaMalloc'ed by the
 *    caller, e.g. torch.Tensor.data_ptr()); no torch or CUDA types appear in the signatures;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *  - matrices are column-major FP64 with an explicit leading dimension;
 *  - every call returns 0 on success, a cudaError_t value (>0) if a launch failed, or -1 for bad
 *    arguments. Kernels are asynchronous on `stream`; results that the host needs are written to
 *    device memory and fetched by the caller.
 *
 * Each entry point names the reference code whose arithmetic it replaces.
 */
#ifndef CONEX_B200_DEVICE_H
#define CONEX_B200_DEVICE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- K1/K2/K3/K7/K8 workhorse: FP64 tensor-core (DMMA.8x8x4) GEMM --------------------------
 * C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],  b = 0..batch-1 (strided batches).
 * op(A) is M x K, op(B) is K x N. lower_only != 0 computes/stores only tiles and entries with
 * row >= col (SYRK-style Gram and Cholesky trailing updates).
 * Replaces Eigen GEMM at dense_lmi_constraint.cc:31-32,77-78, psd_constraint.cc:25,61,79,
 * exponential_map_pade.cc:13-17, block_triangular_operations.cc:205-216. */
int cxb_dgemm(void* stream, int transA, int transB, int M, int N, int K, double alpha,
              const double* dA, long lda, long strideA, const double* dB, long ldb, long strideB,
              double beta, double* dC, long ldc, long strideC, int batch, int lower_only);

/* Same with explicit control, used by the tuning harness (tools/gemm_tune.py) and the Schur assembly:
 * config = tile configuration (-1: chosen from the shape), splits = split-K factor (0: automatic,
 * deterministic fixed-order reduction), mirror != 0 (needs lower_only, M == N) also stores
 * C[col,row] so that the result is exactly symmetric; diag_off shifts the lower_only test to
 * row + diag_off >= col (a row panel starting diag_off rows below the top of a lower-triangular
 * result). */
int cxb_dgemm_ex(void* stream, int config, int splits, int transA, int transB, int M, int N, int K,
                 double alpha, const double* dA, long lda, long strideA, const double* dB, long ldb,
                 long strideB, double beta, double* dC, long ldc, long strideC, int batch,
                 int lower_only, int mirror, int diag_off);
/* Tile configuration used for large shapes when config < 0 (process-wide; tuning only). */
void cxb_set_default_gemm_config(int config);

/* ---- K1+K2: Schur complement of one dense LMI block -----------------------------------------
 * dAall: (m+1) contiguous column-major n x n matrices: A_0..A_{m-1} followed by C.
 * dW: n x n scaling point. dB: scratch of (m+2)*n*n doubles (receives W A_i W, W C W, W).
 * dT: scratch of panel*n*n doubles (panel >= 1 constraint matrices scaled per pass).
 * dHaug: (m+2) x (m+1) column-major, ld = ldh >= m+2. On return (lower trapezoid):
 *   Haug[i][j] = tr(W A_i W A_j) (i >= j, i,j < m)        -> H  (dense_lmi_constraint.cc:72-78)
 *   Haug[m][j] = <W C W, A_j> = AQc_j ; Haug[m][m] = <c,Qc>   (:80, :84-88)
 *   Haug[m+1][j] = <W, A_j> = AW_j  ; Haug[m+1][m] = <w,c>     (:79, :82)
 */
int cxb_schur_dense_lmi(void* stream, int n, int m, const double* dAall, const double* dW,
                        double* dB, double* dT, int panel, double* dHaug, long ldh);

/* Same result with bounded scratch: dBp holds (panel + 1) * n * n doubles (one row panel of scaled
 * matrices, contracted immediately), dT panel * n * n. Used when a second A-sized buffer does not fit. */
int cxb_schur_dense_lmi_streamed(void* stream, int n, int m, const double* dAall, const double* dW,
                                 double* dBp, double* dT, int panel, double* dHaug, long ldh);

/* ---- K3: blocked right-looking Cholesky, lower, in place (block_triangular_operations.cc:184-219,
 * Eigen::LLT). d_info (device int) is set to 0 on success or to (1 + index of the first
 * non-positive pivot). d_work: at least cxb_potrf_worksize(m) doubles. */
size_t cxb_potrf_worksize(int m);
int cxb_potrf_lower(void* stream, int m, double* dH, long ldh, double* d_work, int* d_info);

/* ---- K5: triangular solves with the Cholesky factor (block_triangular_operations.cc:114-182):
 * X <- L^{-T} L^{-1} X for nrhs right-hand sides (columns of dX, leading dimension ldx). */
int cxb_potrs_lower(void* stream, int m, const double* dL, long ldl, double* dX, long ldx, int nrhs);

/* ---- K6: negative slack  out = sum_j coef[j] * Aall[:, j]  (dense_lmi_constraint.cc:8-27);
 * Aall is nn x cols column-major (ld = nn), d_coef has `cols` entries (y followed by -k for C). */
int cxb_gemv_n(void* stream, long nn, int cols, const double* dAall, const double* d_coef,
               double* d_out);

/* ---- K7: two-sided Lanczos (approximate_eigenvalues.cc:178-239) ------------------------------
 * d_WS, d_W: n x n (ld n; W symmetric). Start vector: d_r (n entries) when d_col_index is NULL,
 * otherwise column (int)*d_col_index of the n x n matrix d_r (the index stays on the device, so no
 * host round trip is needed after the diag-argmax reduction). Runs at most num_iter steps with the reference's
 * breakdown test (beta^2 < 1e-6). Outputs: d_alpha[num_iter], d_beta[num_iter], d_count[0] = number
 * of valid beta entries (alpha has d_count+1 valid entries). d_work: cxb_lanczos_worksize(n)
 * doubles. */
size_t cxb_lanczos_worksize(int n);
int cxb_lanczos_two_sided(void* stream, int n, const double* d_WS, const double* d_W,
                          const double* d_r, const double* d_col_index, int num_iter,
                          double* d_alpha, double* d_beta, int* d_count, double* d_work);

/* ---- K7 reductions (psd_constraint.cc:63-80,107-127). d_out slots (all device doubles):
 *   out[0] = tr(WS), out[1] = sum_ij WS_ij WS_ji = tr(WS WS), out[2] = argmax_i WS_ii (as double),
 *   out[3] = max_i WS_ii. */
int cxb_ws_reductions(void* stream, int n, const double* d_WS, double* d_out);

/* ---- K8: geodesic update  W <- sym( pade33( scale * (WS + e_weight I) ) W )
 * (psd_constraint.cc:13-28, exponential_map_pade.cc:10-32). d_WS is destroyed.
 * d_work: cxb_geodesic_worksize(n) doubles; d_iwork: 2n ints (pivots + permutation). d_info: device int, 0 or
 * 1 + index of a zero pivot. */
size_t cxb_geodesic_worksize(int n);
int cxb_geodesic_update(void* stream, int n, double* d_W, double* d_WS, double e_weight,
                        double scale, double* d_work, int* d_iwork, int* d_info);

/* Padé map alone: d_out = pade33(d_X) (exponential_map_pade.cc:23-32). d_X is preserved. */
int cxb_pade_expm(void* stream, int n, const double* d_X, double* d_out, double* d_work,
                  int* d_iwork, int* d_info);

/* Partial-pivot LU solve A X = B (A n x n destroyed, B n x nrhs overwritten by X). */
int cxb_lu_solve(void* stream, int n, double* dA, long lda, int nrhs, double* dB, long ldb,
                 int* d_ipiv, int* d_info);

/* ---- small vector / matrix helpers used by the host loop -------------------------------------*/
/* W <- I (psd_constraint.cc:92-95) */
int cxb_set_identity(void* stream, int n, double* dW);
/* out[0] = x . y  (n entries) */
int cxb_dot(void* stream, long n, const double* dx, const double* dy, double* d_out);
/* y <- a*x + b*y + c*z (z may be NULL) */
int cxb_axpbypcz(void* stream, long n, double a, const double* dx, double b, double* dy, double c,
                 const double* dz);
/* strided copy: dst[i*incd] = src[i*incs] */
int cxb_copy_strided(void* stream, long n, const double* src, long incs, double* dst, long incd);
/* H[idx[a], idx[b]] (+)= G[a, b] for a >= b (lower) — supernodal_assembler.cc:103-165 in dense form */
int cxb_scatter_add_lower(void* stream, int mc, const double* dG, long ldg, const int* d_idx,
                          double* dH, long ldh);
/* dst[0..n) = value */
int cxb_fill(void* stream, long n, double value, double* d_dst);
/* dst[idx[i]] += src[i] / dst[i] = src[idx[i]]  (idx NULL = identity) — the clique gather/scatter of
 * constraint_manager.h:107-124 and cone_program.h:59-67 */
int cxb_scatter_add_vec(void* stream, int n, const double* d_src, const int* d_idx, double* d_dst);
int cxb_gather_vec(void* stream, int n, const double* d_src, const int* d_idx, double* d_dst);
/* W <- (1 + w_e) W + WSW  (psd_constraint.cc:33-43) */
int cxb_affine_update(void* stream, int n, double* dW, const double* dWSW, double w_e);

#ifdef __cplusplus
}
#endif
#endif
