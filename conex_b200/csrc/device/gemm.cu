// FP64 tensor-core GEMM for sm_100a: C = alpha * op(A) * op(B) + beta * C.
//
// The B200's FP64 tensor path is the warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, one per
// 8 cycles per SM sub-partition = 37 TFLOP/s at 1965 MHz); tcgen05 has no f64 kind, so there is no
// TMEM/UMMA involvement. The kernel is a multi-stage cp.async (LDGSTS) pipeline feeding
// register-tiled DMMA fragments:
//
//   * CTA tile BM x BN x BK, warp tile WM x WN -> (WM/8) x (WN/8) independent 8x8 accumulators per
//     warp, so every k-step issues (WM/8)*(WN/8) DMMAs for (WM/8)+(WN/8) 8-byte LDS per lane;
//   * both operand tiles sit in shared memory with a pitch == 4 (mod 16) doubles, which makes the
//     8x4 / 4x8 fragment loads bank-conflict free for either operand orientation
//     ("K-contiguous" = transposed A / plain B, "M/N-contiguous" = plain A / transposed B);
//   * ragged edges are zero-filled by cp.async's src-size operand: no branches in the MMA loop;
//   * lower_only skips the tiles strictly above the diagonal (SYRK / Gram / Cholesky updates) and
//     `mirror` additionally stores the transposed entry, producing an exactly symmetric result
//     (used for W (A_i W), which is symmetric in exact arithmetic);
//   * split-K (deterministic): K is cut into `splits` fixed ranges, every range writes its partial
//     tile to a workspace and a second kernel sums the partials in a fixed order. Used by the Gram
//     contraction (K = n^2) when the lower tile triangle alone cannot fill 148 SMs.
//
// Several tile configurations are compiled; Dgemm() picks one from the shape (see PickConfig) and
// cxb_dgemm_ex lets the tuning harness (tools/gemm_tune.py) force one.
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
  int lower;    // only entries with row + diag_off >= col
  int diag_off; // row offset of this block relative to the diagonal (trapezoids / row panels)
  int mirror;   // also store C[col, row] (requires lower, M == N)
  int tiles_m, tiles_n;
  int tri;      // bit 0: op(B)[k, c] = 0 for k < c; bit 1: op(A)[r, k] = 0 for k < r: k-tiles that
                // lie entirely in the zero part are skipped
  int pack;     // store lower tiles (BM == BN) contiguously, off-diagonal ones scaled by sqrt(2)
  int splits;   // split-K factor (>1: partials go to `partials`, batch must be 1)
  int k_per_split;  // multiple of BK
  double* partials;  // splits x (M x N, ld = M)
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

// Streams one operand's tiles (ROWS x COLS logical, contiguous along COLS in global memory) into
// shared memory with row pitch PITCH. KC == true: rows index M (or N), columns index K;
// KC == false: rows index K, columns index M (or N). The per-thread column and the row of its
// first chunk never change, so the 64-bit address arithmetic is hoisted out of the k loop.
template <int ROWS, int COLS, int PITCH, int VEC, int THREADS, bool KC>
struct TileLoader {
  static constexpr int kChunksPerRow = COLS / VEC;
  static constexpr int kTotal = ROWS * kChunksPerRow;
  static constexpr int kIters = (kTotal + THREADS - 1) / THREADS;
  static_assert(THREADS % kChunksPerRow == 0 || kChunksPerRow % THREADS == 0,
                "a thread must keep one column across its chunks");
  static constexpr int kRowStep = (THREADS >= kChunksPerRow) ? THREADS / kChunksPerRow : 0;
  static_assert(kRowStep > 0, "tile row wider than the CTA");

  const double* base;   // &g[(mn0 + r_first) * ld + cc]  (KC) or &g[r_first * ld + mn0 + cc] (!KC)
  const double* safe;   // any valid address for fully masked chunks
  long row_stride;      // kRowStep * ld
  long ld;
  int r_first, cc;
  int mn_bytes;         // !KC: valid bytes of this thread's column chunk (fixed)
  unsigned row_mask;    // KC: bit i set when row r_first + i*kRowStep is inside the matrix
  int smem_off;         // r_first * PITCH + cc

  __device__ __forceinline__ void Init(const double* g, long ld_, int mn0, int mn_max, int tid) {
    ld = ld_;
    safe = g;
    r_first = tid / kChunksPerRow;
    cc = (tid % kChunksPerRow) * VEC;
    row_stride = (long)kRowStep * ld;
    smem_off = r_first * PITCH + cc;
    if (KC) {
      base = g + (long)(mn0 + r_first) * ld + cc;
      row_mask = 0;
#pragma unroll
      for (int i = 0; i < kIters; i++) {
        const int r = r_first + i * kRowStep;
        if (r < ROWS && mn0 + r < mn_max) row_mask |= (1u << i);
      }
      mn_bytes = 0;
    } else {
      base = g + (long)r_first * ld + mn0 + cc;
      int v = mn_max - (mn0 + cc);
      v = v < 0 ? 0 : (v > VEC ? VEC : v);
      mn_bytes = v * 8;
      row_mask = 0;
    }
  }

  // Issue the cp.asyncs of the k-tile starting at k0 (k < k_end are valid) into `smem`.
  __device__ __forceinline__ void Load(double* smem, int k0, int k_end) const {
    if (KC) {
      int v = k_end - (k0 + cc);
      v = v < 0 ? 0 : (v > VEC ? VEC : v);
      const int kbytes = v * 8;
      const double* p = base + k0;
#pragma unroll
      for (int i = 0; i < kIters; i++) {
        if ((kTotal % THREADS != 0) && (r_first + i * kRowStep >= ROWS)) break;
        const int bytes = ((row_mask >> i) & 1u) ? kbytes : 0;
        const double* src = bytes ? p + i * row_stride : safe;
        double* dst = smem + smem_off + i * kRowStep * PITCH;
        if (VEC == 2) CpAsync16(dst, src, bytes); else CpAsync8(dst, src, bytes);
      }
    } else {
      const double* p = base + (long)k0 * ld;
#pragma unroll
      for (int i = 0; i < kIters; i++) {
        const int r = r_first + i * kRowStep;
        if ((kTotal % THREADS != 0) && (r >= ROWS)) break;
        const int bytes = (k0 + r < k_end) ? mn_bytes : 0;
        const double* src = bytes ? p + i * row_stride : safe;
        double* dst = smem + smem_off + i * kRowStep * PITCH;
        if (VEC == 2) CpAsync16(dst, src, bytes); else CpAsync8(dst, src, bytes);
      }
    }
  }
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB, bool AKC, bool BKC, int VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
    DgemmKernel(const GemmArgs g) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, AKC, BKC, VEC>;
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int NT = Cfg::kThreads;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::kStageA;

  // ---- tile coordinates ----
  const int tm = blockIdx.x % g.tiles_m;
  const int tn = blockIdx.x / g.tiles_m;
  const int m0 = tm * BM, n0 = tn * BN;
  if (g.lower && m0 + BM - 1 + g.diag_off < n0) return;  // tile strictly above the diagonal
  const double* A = g.A + (long)blockIdx.z * g.sA;
  const double* B = g.B + (long)blockIdx.z * g.sB;
  double* C = g.C + (long)blockIdx.z * g.sC;
  int k_begin = blockIdx.y * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  if (g.tri & 1) k_begin = max(k_begin, (n0 / BK) * BK);
  if (g.tri & 2) k_begin = max(k_begin, (m0 / BK) * BK);

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

  using LoaderA = TileLoader<Cfg::kRowsA, AKC ? BK : BM, Cfg::kPitchA, VEC, NT, AKC>;
  using LoaderB = TileLoader<Cfg::kRowsB, BKC ? BK : BN, Cfg::kPitchB, VEC, NT, BKC>;
  LoaderA la;
  LoaderB lb;
  la.Init(A, g.lda, m0, g.M, tid);
  lb.Init(B, g.ldb, n0, g.N, tid);

  const int KT = (k_end - k_begin + BK - 1) / BK;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) {
      la.Load(As + s * Cfg::kStageA, k_begin + s * BK, k_end);
      lb.Load(Bs + s * Cfg::kStageB, k_begin + s * BK, k_end);
    }
    CpAsyncCommit();
  }

  // fragment offsets inside a stage (doubles)
  const int a_off = AKC ? (wm0 + gid) * Cfg::kPitchA + tig : tig * Cfg::kPitchA + wm0 + gid;
  const int b_off = BKC ? (wn0 + gid) * Cfg::kPitchB + tig : tig * Cfg::kPitchB + wn0 + gid;
  constexpr int kAStepI = AKC ? 8 * Cfg::kPitchA : 8;     // next 8 rows of the warp tile
  constexpr int kAStepK = AKC ? 1 : Cfg::kPitchA;          // next k
  constexpr int kBStepJ = BKC ? 8 * Cfg::kPitchB : 8;
  constexpr int kBStepK = BKC ? 1 : Cfg::kPitchB;

  int stage = 0;
  for (int kt = 0; kt < KT; kt++) {
    CpAsyncWait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      int ns = stage + STAGES - 1;
      if (ns >= STAGES) ns -= STAGES;
      if (nk < KT) {
        la.Load(As + ns * Cfg::kStageA, k_begin + nk * BK, k_end);
        lb.Load(Bs + ns * Cfg::kStageB, k_begin + nk * BK, k_end);
      }
      CpAsyncCommit();
    }
    const double* as = As + stage * Cfg::kStageA + a_off;
    const double* bs = Bs + stage * Cfg::kStageB + b_off;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = as[i * kAStepI + kk * kAStepK];
#pragma unroll
      for (int j = 0; j < NI; j++) b[j] = bs[j * kBStepJ + kk * kBStepK];
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) Dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (++stage == STAGES) stage = 0;
  }
  CpAsyncWait<0>();

  // ---- epilogue ----
  if (g.splits > 1) {
    double* P = g.partials + (long)blockIdx.y * g.M * g.N;
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
          if (g.lower && r + g.diag_off < c) continue;
          P[(long)c * g.M + r] = acc[i][j][e];
        }
      }
    }
    return;
  }
  if (g.pack) {
    // Tile (tm, tn), tm >= tn, goes to slot tn * T - tn (tn - 1) / 2 + (tm - tn) of BM * BN doubles
    // (column-major inside the tile). Entries beyond the matrix edge are exact zeros (zero-filled
    // operands) and are stored too, so the packed vector has a fixed length.
    const long slot = (long)tn * g.tiles_m - (long)tn * (tn - 1) / 2 + (tm - tn);
    double* P = C + slot * (BM * BN);
    const double sc = (tm == tn) ? g.alpha : g.alpha * 1.4142135623730951;
#pragma unroll
    for (int i = 0; i < MI; i++) {
      const int rl = wm0 + i * 8 + gid;
#pragma unroll
      for (int j = 0; j < NI; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int cl = wn0 + j * 8 + tig * 2 + e;
          P[cl * BM + rl] = sc * acc[i][j][e];
        }
      }
    }
    return;
  }
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
        if (g.lower && r + g.diag_off < c) continue;
        double* p = C + (long)c * g.ldc + r;
        double v = g.alpha * acc[i][j][e];
        if (use_beta) v += g.beta * (*p);
        *p = v;
        if (g.mirror && r != c) C[(long)r * g.ldc + c] = v;
      }
    }
  }
}

int g_max_ktiles_per_cta = 0;  // cxb_set_gemm_split_policy: 0 = split K only to fill the machine

// C = alpha * sum_s partials[s] + beta * C (fixed summation order s = 0, 1, ...).
__global__ void __launch_bounds__(256) SplitKReduceKernel(int M, int N, int splits,
                                                          const double* __restrict__ partials,
                                                          double alpha, double beta, double* C, long ldc,
                                                          int lower, int diag_off) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y;
  if (r >= M || (lower && r + diag_off < c)) return;
  const long mn = (long)M * N;
  const double* p = partials + (long)c * M + r;
  double s = 0;
  for (int k = 0; k < splits; k++) s += p[k * mn];
  double* out = C + (long)c * ldc + r;
  double v = alpha * s;
  if (beta != 0.0) v += beta * (*out);
  *out = v;
}

// Split-K partial sums live in stream-ordered memory (cudaMallocAsync / cudaFreeAsync on the launching
// stream): every program owns its stream (DeviceContext), programs may be driven from different host
// threads and devices, and a process-wide buffer would be shared by concurrent GEMMs. The default pool
// keeps the block cached between launches (release threshold set by DeviceContext), so this costs no
// driver call in steady state.
double* SplitKWorkspace(cudaStream_t stream, size_t doubles) {
  void* p = nullptr;
  if (cudaMallocAsync(&p, doubles * sizeof(double), stream) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return static_cast<double*>(p);
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB, bool AKC, bool BKC, int VEC>
int LaunchCfg(cudaStream_t stream, GemmArgs g, int batch, int splits) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, STAGES, AKC, BKC, VEC>;
  static_assert(MINB * (Cfg::kSmemBytes + 1024) <= 227 * 1024, "tile configuration exceeds shared memory");
  auto kernel = DgemmKernel<BM, BN, BK, WM, WN, STAGES, MINB, AKC, BKC, VEC>;
  constexpr int kSlots = kNumSMs * MINB;  // CTAs resident at once
  static std::atomic<unsigned long long> configured{0};
  if (FirstUseOnCurrentDevice(configured)) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
  }
  g.tiles_m = (g.M + BM - 1) / BM;
  g.tiles_n = (g.N + BN - 1) / BN;
  const long tiles = (long)g.tiles_m * g.tiles_n;
  long active = tiles;
  if (g.lower) {
    active = 0;
    for (int tn = 0; tn < g.tiles_n; tn++) {
      // smallest tm with tm*BM + BM - 1 + diag_off >= tn*BN
      const long num = (long)tn * BN - g.diag_off;
      const int f = num <= 0 ? 0 : (int)(num / BM);
      if (f < g.tiles_m) active += g.tiles_m - f;
    }
  }
  const int kt_total = (g.K + BK - 1) / BK;
  if (splits == 0) {
    // automatic: split only when the tiles cannot fill the machine once and K is deep
    splits = 1;
    const long waves1 = (active + kSlots - 1) / kSlots;
    if (batch == 1 && (double)active / (double)(waves1 * kSlots) < 0.9 && kt_total >= 64) {
      double best_eff = (double)active / (double)(waves1 * kSlots);
      for (int s = 2; s <= 128; s++) {
        if (kt_total / s < 32) break;
        const long ctas = active * s;
        const long waves = (ctas + kSlots - 1) / kSlots;
        const double eff = (double)ctas / (double)(waves * kSlots);
        if (eff > best_eff + 0.04) {
          best_eff = eff;
          splits = s;
        }
      }
    }
  }
  if (g_max_ktiles_per_cta > 0 && batch == 1 && kt_total > g_max_ktiles_per_cta) {
    // Deep contractions (the Gram over K = n^2 or the packed length): a CTA that walks 10^4 k-tiles drifts away from
    // the CTAs sharing its operand panels by more than L2 holds, and the panels are re-read from DRAM (K2 of C2:
    // 400 GB for 35 GB of operand, profiles/r02_e_c2_symmetric_assembly_ncu_full.txt). Shorter CTAs stay aligned.
    // Bounded by a partial-sum workspace of 768 MB.
    const long ws_cap = (768L << 20) / ((long)g.M * g.N * 8);
    int want = (kt_total + g_max_ktiles_per_cta - 1) / g_max_ktiles_per_cta;
    if (want > ws_cap) want = (int)ws_cap;
    if (want > splits) {
      int best_s = want;
      double best_eff = 0;
      for (int s = want; s < want + 8 && s <= ws_cap; s++) {
        const long ctas = active * s;
        const long waves = (ctas + kSlots - 1) / kSlots;
        const double eff = (double)ctas / (double)(waves * kSlots);
        if (eff > best_eff + 1e-9) {
          best_eff = eff;
          best_s = s;
        }
      }
      splits = best_s;
    }
  }
  if (batch != 1 || g.mirror || g.pack || g.tri) splits = 1;
  if (g.pack && BM != BN) return -1;
  if (splits > kt_total) splits = kt_total > 0 ? kt_total : 1;
  g.splits = splits;
  const int kt_per = (kt_total + splits - 1) / (splits > 0 ? splits : 1);
  g.k_per_split = (splits > 1) ? kt_per * BK : (g.K > 0 ? g.K : 1);
  if (splits > 1) {
    // drop empty trailing splits
    g.splits = splits = (kt_total + kt_per - 1) / kt_per;
    g.partials = SplitKWorkspace(stream, (size_t)splits * g.M * g.N);
    if (g.partials == nullptr) return (int)cudaErrorMemoryAllocation;
  }
  dim3 grid((unsigned)tiles, (unsigned)splits, (unsigned)batch);
  CountLaunch(); kernel<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(g);
  if (splits > 1) {
    dim3 rg((g.M + 255) / 256, g.N);
    CountLaunch(); SplitKReduceKernel<<<rg, 256, 0, stream>>>(g.M, g.N, splits, g.partials, g.alpha, g.beta,
                                                  g.C, g.ldc, g.lower, g.diag_off);
    cudaFreeAsync(g.partials, stream);
  }
  return LaunchStatus();
}

// Tile configurations (BM, BN, BK, WM, WN, STAGES, CTAs/SM):
//   0: 64x64, 4 warps of 32x32, 3 stages, >=3 CTAs/SM (small problems)
//   1: 128x128, 8 warps of 64x32, 4 stages, 1 CTA/SM (fewest L2 bytes per flop)
//   2: 64x64, 4 warps of 32x32, 2 stages, 4 CTAs/SM (<=128 registers)
//   3: 128x64, 8 warps of 32x32, 3 stages, 2 CTAs/SM
//   4: 128x128, BK = 32, 3 stages, 1 CTA/SM
//   5: 64x128, 8 warps of 32x32, 3 stages, 2 CTAs/SM
//   6: 64x128, 4 warps of 32x64, 3 stages, 2 CTAs/SM (the shape of cuBLAS's own sm_100 FP64 kernel: 32 DMMA per 12
//      fragment loads instead of 16 per 8)
//   7: 128x64, 4 warps of 64x32, 3 stages, 2 CTAs/SM
//   8: 64x128, 4 warps of 32x64, 2 stages, 3 CTAs/SM
//   9: 64x64, BK = 32, 4 warps of 32x32, 2 stages, 3 CTAs/SM (half the k-tile barriers of configuration 2)
// Several independent CTAs per SM keep the DMMA pipe busy across each other's barriers and
// fragment-load latencies (the profile of the one-CTA configurations shows `wait` and
// `short_scoreboard` stalls at every k-tile boundary).
template <bool AKC, bool BKC, int VEC>
int LaunchLayout(cudaStream_t stream, const GemmArgs& g, int batch, int config, int splits) {
  switch (config) {
    case 0: return LaunchCfg<64, 64, 16, 32, 32, 3, 3, AKC, BKC, VEC>(stream, g, batch, splits);
    case 1: return LaunchCfg<128, 128, 16, 64, 32, 4, 1, AKC, BKC, VEC>(stream, g, batch, splits);
    case 2: return LaunchCfg<64, 64, 16, 32, 32, 2, 4, AKC, BKC, VEC>(stream, g, batch, splits);
    case 3: return LaunchCfg<128, 64, 16, 32, 32, 3, 2, AKC, BKC, VEC>(stream, g, batch, splits);
    case 4: return LaunchCfg<128, 128, 32, 64, 32, 3, 1, AKC, BKC, VEC>(stream, g, batch, splits);
    case 5: return LaunchCfg<64, 128, 16, 32, 32, 3, 2, AKC, BKC, VEC>(stream, g, batch, splits);
    case 6: return LaunchCfg<64, 128, 16, 32, 64, 3, 2, AKC, BKC, VEC>(stream, g, batch, splits);
    case 7: return LaunchCfg<128, 64, 16, 64, 32, 3, 2, AKC, BKC, VEC>(stream, g, batch, splits);
    case 8: return LaunchCfg<64, 128, 16, 32, 64, 2, 3, AKC, BKC, VEC>(stream, g, batch, splits);
    case 9: return LaunchCfg<64, 64, 32, 32, 32, 2, 3, AKC, BKC, VEC>(stream, g, batch, splits);
    default: return -1;
  }
}

bool Aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int g_default_large_config = 2;

// Optional separate configuration for the deep A^T B contractions (the Gram matrices of the Schur assembly: K = n^2 or
// its packed length); -1 (default) = the same as the other large products. The 64 x 128 tile with 32 x 64 warp tiles (6)
// wins a stand-alone Gram of random data (28.4 vs 25.4-26.8 TFLOP/s of lower-triangle flops at m = 2000, K = 1e6,
// profiles/r02_o_gemm_wide_warp_tiles.jsonl) and LOSES inside the C2 step (assembly 986.7 vs 959.7 ms,
// profiles/r02_p_bench_c2_gram_config.txt), so it is not the default.
int g_gram_config = -1;

int PickConfig(bool transA, bool transB, int M, int N, int K) {
  if (M <= 96 || N <= 96) return 0;
  if (transA && !transB && K >= 16384 && g_gram_config >= 0) return g_gram_config;
  return g_default_large_config;
}

}  // namespace

int DgemmEx(cudaStream_t stream, int config, int splits, bool transA, bool transB, int M, int N, int K,
            double alpha, const double* A, long lda, long sA, const double* B, long ldb, long sB,
            double beta, double* C, long ldc, long sC, int batch, bool lower_only, bool mirror,
            int diag_off) {
  return DgemmStructured(stream, config, splits, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta,
                         C, ldc, sC, batch, lower_only, mirror, diag_off, 0, false);
}

long PackedSymmetricSize(int n) {
  const long t = (n + 63) / 64;
  return t * (t + 1) / 2 * 4096;
}

int DgemmStructured(cudaStream_t stream, int config, int splits, bool transA, bool transB, int M, int N,
                    int K, double alpha, const double* A, long lda, long sA, const double* B, long ldb,
                    long sB, double beta, double* C, long ldc, long sC, int batch, bool lower_only,
                    bool mirror, int diag_off, int tri, bool pack) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if (K < 0) return -1;
  if (mirror && (!lower_only || M != N || diag_off != 0)) return -1;
  if (pack && (!lower_only || M != N || diag_off != 0 || mirror || beta != 0.0)) return -1;
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
  g.diag_off = diag_off;
  g.mirror = mirror ? 1 : 0;
  g.tri = tri;
  g.pack = pack ? 1 : 0;
  g.tiles_m = g.tiles_n = 0;
  g.splits = 1;
  g.k_per_split = K;
  g.partials = nullptr;
  if (config < 0) config = PickConfig(transA, transB, M, N, K);
  if (pack && config != 0 && config != 2 && config != 9) config = 2;  // packing needs the 64 x 64 tile
  const bool vec2 = Aligned16(A) && Aligned16(B) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                    (sA % 2 == 0) && (sB % 2 == 0);
  const bool akc = transA, bkc = !transB;
#define CXB_DISPATCH(AK, BK_)                                                     \
  if (akc == AK && bkc == BK_) {                                                  \
    return vec2 ? LaunchLayout<AK, BK_, 2>(stream, g, batch, config, splits)      \
                : LaunchLayout<AK, BK_, 1>(stream, g, batch, config, splits);     \
  }
  CXB_DISPATCH(false, false)
  CXB_DISPATCH(false, true)
  CXB_DISPATCH(true, false)
  CXB_DISPATCH(true, true)
#undef CXB_DISPATCH
  return -1;
}

int Dgemm(cudaStream_t stream, bool transA, bool transB, int M, int N, int K, double alpha,
          const double* A, long lda, long sA, const double* B, long ldb, long sB, double beta,
          double* C, long ldc, long sC, int batch, bool lower_only) {
  return DgemmEx(stream, -1, 0, transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC,
                 batch, lower_only, false, 0);
}

void SetDefaultGemmConfig(int config) { g_default_large_config = config; }
void SetGramGemmConfig(int config) { g_gram_config = config; }

}  // namespace cxb

extern "C" int cxb_dgemm(void* stream, int transA, int transB, int M, int N, int K, double alpha,
                         const double* dA, long lda, long strideA, const double* dB, long ldb,
                         long strideB, double beta, double* dC, long ldc, long strideC, int batch,
                         int lower_only) {
  return cxb::Dgemm(cxb::AsStream(stream), transA != 0, transB != 0, M, N, K, alpha, dA, lda,
                    strideA, dB, ldb, strideB, beta, dC, ldc, strideC, batch, lower_only != 0);
}

extern "C" int cxb_dgemm_ex(void* stream, int config, int splits, int transA, int transB, int M, int N,
                            int K, double alpha, const double* dA, long lda, long strideA,
                            const double* dB, long ldb, long strideB, double beta, double* dC,
                            long ldc, long strideC, int batch, int lower_only, int mirror,
                            int diag_off) {
  return cxb::DgemmEx(cxb::AsStream(stream), config, splits, transA != 0, transB != 0, M, N, K, alpha,
                      dA, lda, strideA, dB, ldb, strideB, beta, dC, ldc, strideC, batch,
                      lower_only != 0, mirror != 0, diag_off);
}

extern "C" void cxb_set_default_gemm_config(int config) { cxb::SetDefaultGemmConfig(config); }
extern "C" void cxb_set_gram_gemm_config(int config) { cxb::SetGramGemmConfig(config); }
extern "C" void cxb_set_gemm_split_policy(int max_ktiles_per_cta) { cxb::g_max_ktiles_per_cta = max_ktiles_per_cta; }
