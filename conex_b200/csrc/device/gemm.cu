// FP64 tensor-core GEMM for sm_100a: C = alpha * op(A) * op(B) + beta * C.
//
// The B200's FP64 tensor path is the legacy warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4);
// tcgen05 has no f64 kind, so there is no TMEM/UMMA involvement. The kernel is a classic
// multi-stage cp.async (LDGSTS) pipeline feeding register-tiled DMMA fragments:
//
//   * CTA tile BM x BN x BK, warp tile WM x WN -> (WM/8) x (WN/8) independent 8x8 accumulators per
//     warp, so every k-step issues (WM/8)*(WN/8) DMMAs for (WM/8)+(WN/8) 8-byte LDS per lane;
//   * both operand tiles are stored in shared memory with a pitch == 4 (mod 16) doubles, which
//     makes the 8x4 / 4x8 fragment loads bank-conflict free for either operand orientation
//     ("K-contiguous" = transposed A / plain B, "M/N-contiguous" = plain A / transposed B);
//   * ragged edges are zero-filled by cp.async's src-size operand, so no branches in the MMA loop;
//   * lower_only launches only the tiles on/below the diagonal (SYRK / Gram / Cholesky updates).
//
// Replaces the Eigen GEMM calls of the reference hot path (see include/conex_b200_device.h).
#include "common.cuh"
#include "device_api.h"

namespace cxb {
namespace {

struct GemmArgs {
  int M, N, K;
  double alpha, beta;
  const double* A;
  long lda, sA;
  const double* B;
  long ldb, sB;
  double* C;
  long ldc, sC;
  int lower;
  int tiles_m, tiles_n;
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKC, bool BKC, int VEC>
struct GemmCfg {
  static constexpr int kWarpsM = BM / WM;
  static constexpr int kWarpsN = BN / WN;
  static constexpr int kThreads = kWarpsM * kWarpsN * 32;
  static constexpr int kPitchA = AKC ? (BK + 4) : (BM + 4);
  static constexpr int kPitchB = BKC ? (BK + 4) : (BN + 4);
  static constexpr int kRowsA = AKC ? BM : BK;
  static constexpr int kRowsB = BKC ? BN : BK;
  static constexpr int kStageA = kRowsA * kPitchA;  // doubles
  static constexpr int kStageB = kRowsB * kPitchB;
  static constexpr size_t kSmemBytes = sizeof(double) * STAGES * (kStageA + kStageB);
};

// Copies one operand tile (ROWS x COLS logical, contiguous along COLS in global memory) into
// shared memory with row pitch PITCH. Element (r, c) lives at g[(r0 + r) * ld + c0 + c]; rows
// beyond `rmax` and columns beyond `cmax` are zero-filled.
template <int ROWS, int COLS, int PITCH, int VEC, int THREADS>
__device__ __forceinline__ void LoadTile(double* smem, const double* __restrict__ g, long ld,
                                         int r0, int c0, int rmax, int cmax, int tid) {
  constexpr int kChunks = COLS / VEC;
  constexpr int kTotal = ROWS * kChunks;
#pragma unroll
  for (int i = 0; i < (kTotal + THREADS - 1) / THREADS; i++) {
    const int c = tid + i * THREADS;
    if ((kTotal % THREADS != 0) && c >= kTotal) break;
    const int r = c / kChunks;
    const int cc = (c % kChunks) * VEC;
    const int gr = r0 + r, gc = c0 + cc;
    int valid = 0;
    if (gr < rmax) {
      valid = cmax - gc;
      valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
    }
    const double* src = valid ? (g + (long)gr * ld + gc) : g;
    double* dst = smem + r * PITCH + cc;
    if (VEC == 2) {
      CpAsync16(dst, src, valid * 8);
    } else {
      CpAsync8(dst, src, valid * 8);
    }
  }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKC, bool BKC, int VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
    DgemmKernel(const GemmArgs g) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, AKC, BKC, VEC>;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int NT = Cfg::kThreads;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::kStageA;

  // ---- tile coordinates ----
  int tm, tn;
  if (g.lower) {
    // linear index over tiles with tm >= tn (BM == BN)
    const int t = blockIdx.x;
    int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(r + 1) * (r + 2) / 2 <= t) r++;
    while ((long)r * (r + 1) / 2 > t) r--;
    tm = r;
    tn = t - r * (r + 1) / 2;
  } else {
    tm = blockIdx.x % g.tiles_m;
    tn = blockIdx.x / g.tiles_m;
  }
  const int m0 = tm * BM, n0 = tn * BN;
  if (m0 >= g.M || n0 >= g.N) return;
  const double* A = g.A + (long)blockIdx.z * g.sA;
  const double* B = g.B + (long)blockIdx.z * g.sB;
  double* C = g.C + (long)blockIdx.z * g.sC;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp % Cfg::kWarpsM) * WM;
  const int wn0 = (warp / Cfg::kWarpsM) * WN;
  const int gid = lane >> 2, tig = lane & 3;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (g.K + BK - 1) / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double* as = As + stage * Cfg::kStageA;
    double* bs = Bs + stage * Cfg::kStageB;
    if (AKC) {
      LoadTile<BM, BK, Cfg::kPitchA, VEC, NT>(as, A, g.lda, m0, k0, g.M, g.K, tid);
    } else {
      LoadTile<BK, BM, Cfg::kPitchA, VEC, NT>(as, A, g.lda, k0, m0, g.K, g.M, tid);
    }
    if (BKC) {
      LoadTile<BN, BK, Cfg::kPitchB, VEC, NT>(bs, B, g.ldb, n0, k0, g.N, g.K, tid);
    } else {
      LoadTile<BK, BN, Cfg::kPitchB, VEC, NT>(bs, B, g.ldb, k0, n0, g.K, g.N, tid);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) load_stage(s, s);
    CpAsyncCommit();
  }

  for (int kt = 0; kt < KT; kt++) {
    CpAsyncWait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk % STAGES, nk);
      CpAsyncCommit();
    }
    const double* as = As + (kt % STAGES) * Cfg::kStageA;
    const double* bs = Bs + (kt % STAGES) * Cfg::kStageB;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; i++) {
        a[i] = AKC ? as[(wm0 + i * 8 + gid) * Cfg::kPitchA + kk + tig]
                   : as[(kk + tig) * Cfg::kPitchA + wm0 + i * 8 + gid];
      }
#pragma unroll
      for (int j = 0; j < NI; j++) {
        b[j] = BKC ? bs[(wn0 + j * 8 + gid) * Cfg::kPitchB + kk + tig]
                   : bs[(kk + tig) * Cfg::kPitchB + wn0 + j * 8 + gid];
      }
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) Dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  CpAsyncWait<0>();

  // ---- epilogue ----
  const bool use_beta = g.beta != 0.0;
#pragma unroll
  for (int i = 0; i < MI; i++) {
    const int r = m0 + wm0 + i * 8 + gid;
    if (r >= g.M) continue;
#pragma unroll
    for (int j = 0; j < NI; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = n0 + wn0 + j * 8 + tig * 2 + e;
        if (c >= g.N) continue;
        if (g.lower && r < c) continue;
        double* p = C + (long)c * g.ldc + r;
        double v = g.alpha * acc[i][j][e];
        if (use_beta) v += g.beta * (*p);
        *p = v;
      }
    }
  }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool AKC, bool BKC, int VEC>
int LaunchCfg(cudaStream_t stream, GemmArgs g, int batch) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, AKC, BKC, VEC>;
  auto kernel = DgemmKernel<BM, BN, BK, WM, WN, STAGES, AKC, BKC, VEC>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    configured = true;
  }
  g.tiles_m = (g.M + BM - 1) / BM;
  g.tiles_n = (g.N + BN - 1) / BN;
  // lower_only: enumerate the full tile triangle over tiles_m block rows (tn <= tm); tiles whose
  // column block lies beyond N exit immediately in the kernel.
  const long tiles = g.lower ? (long)g.tiles_m * (g.tiles_m + 1) / 2 : (long)g.tiles_m * g.tiles_n;
  dim3 grid((unsigned)tiles, 1, (unsigned)batch);
  CountLaunch(); kernel<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(g);
  return LaunchStatus();
}

template <bool AKC, bool BKC, int VEC>
int LaunchLayout(cudaStream_t stream, const GemmArgs& g, int batch) {
  const bool small = (g.M <= 96 || g.N <= 96);
  if (small) {
    return LaunchCfg<64, 64, 16, 32, 32, 3, AKC, BKC, VEC>(stream, g, batch);
  }
  return LaunchCfg<128, 128, 16, 64, 32, 4, AKC, BKC, VEC>(stream, g, batch);
}

bool Aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int Dgemm(cudaStream_t stream, bool transA, bool transB, int M, int N, int K, double alpha,
          const double* A, long lda, long sA, const double* B, long ldb, long sB, double beta,
          double* C, long ldc, long sC, int batch, bool lower_only) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if (K < 0) return -1;
  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.alpha = alpha;
  g.beta = beta;
  g.A = A;
  g.lda = lda;
  g.sA = sA;
  g.B = B;
  g.ldb = ldb;
  g.sB = sB;
  g.C = C;
  g.ldc = ldc;
  g.sC = sC;
  g.lower = lower_only ? 1 : 0;
  g.tiles_m = g.tiles_n = 0;
  const bool vec2 = Aligned16(A) && Aligned16(B) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                    (sA % 2 == 0) && (sB % 2 == 0);
  const bool akc = transA, bkc = !transB;
#define CXB_DISPATCH(AK, BK_)                                        \
  if (akc == AK && bkc == BK_) {                                     \
    return vec2 ? LaunchLayout<AK, BK_, 2>(stream, g, batch)         \
                : LaunchLayout<AK, BK_, 1>(stream, g, batch);        \
  }
  CXB_DISPATCH(false, false)
  CXB_DISPATCH(false, true)
  CXB_DISPATCH(true, false)
  CXB_DISPATCH(true, true)
#undef CXB_DISPATCH
  return -1;
}

}  // namespace cxb

extern "C" int cxb_dgemm(void* stream, int transA, int transB, int M, int N, int K, double alpha,
                         const double* dA, long lda, long strideA, const double* dB, long ldb,
                         long strideB, double beta, double* dC, long ldc, long strideC, int batch,
                         int lower_only) {
  return cxb::Dgemm(cxb::AsStream(stream), transA != 0, transB != 0, M, N, K, alpha, dA, lda,
                    strideA, dB, ldb, strideB, beta, dC, ldc, strideC, batch, lower_only != 0);
}
