// Schur complement of one small dense LMI block on the FP64 tensor cores (DMMA m8n8k4), one CTA per program:
// the batched counterpart of device/schur.cu for blocks of order n <= 32 (BASELINE config 3: n = 20, m = 40).
//
//   phase 0  cp.async of the whole cone data [A_0 .. A_{m-1} C] (n^2 doubles each) into shared memory, in commit
//            groups of one matrix per warp; meanwhile warp 0 factors W = L L^T (left-looking, one lane per row);
//   phase A  every warp takes the matrices i = warp, warp + W, ...:  T = A_i L (k-steps above L's diagonal
//            skipped), written in place over A_i;  S = L^T T, lower tiles only (k-steps left of L^T's diagonal
//            skipped), packed into row i of X (lower triangle, off-diagonal entries scaled by sqrt 2, so that the
//            plain dot product of two rows is the trace inner product). Row m + 1 of X is the packed identity;
//   phase B  Gram  X X^T  (lower) in 16 x 16 blocks of four DMMA tiles, the packed length split across warps;
//            partial blocks are summed in a fixed order (deterministic) and written to G / AQc / AW / scal with the
//            conventions of small::PsdSchur (row m: <S_C, S_j>, row m + 1: <I, S_j>).
// Nothing but the cone data is read from HBM and nothing but the m (m + 1) / 2 + 2 m + 2 results is written: the
// scaled matrices never leave shared memory (the DFMA kernel this replaces round-tripped them through HBM and
// spent its time in ~80 barrier-separated phases).
// If W is not numerically positive definite the CTA falls back to small::PsdSchurClassic (the reference's formula).
#pragma once
#include "common.cuh"
#include "small_cone_math.cuh"
#include "team.cuh"

namespace cxb {
namespace psdmma {

// 16 warps per CTA (one CTA per SM: the scaled matrices of one program fill most of its shared memory): the phases are
// chains of shared-memory loads and DMMAs, so the warps of ONE program are all the latency hiding there is
// (8 warps: 32 us per program, profiles/r02_c_c3_launches_1024_programs.txt).
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;

// smallest p >= x with p == 4 (mod 16): row pitch (doubles) that makes the 8 x 4 / 4 x 8 DMMA fragment loads
// bank-conflict free in both orientations
__host__ __device__ constexpr int Pitch4Mod16(int x) {
  return ((x + 15) / 16 * 16 + 4) - 16 >= x ? ((x + 15) / 16 * 16 + 4) - 16 : (x + 15) / 16 * 16 + 4;
}

struct Layout {
  int n, m, kp, k4, pa, pl, px, nblk, ksplit;
  long off_l, off_x, off_a, off_p, total;  // doubles, after the 64-double reduction scratch
};

__host__ __device__ inline Layout MakeLayout(int n, int m) {
  Layout y;
  y.n = n;
  y.m = m;
  y.kp = n * (n + 1) / 2;
  y.k4 = (y.kp + 3) / 4 * 4;
  y.pa = Pitch4Mod16(n);
  y.pl = Pitch4Mod16(n);
  y.px = Pitch4Mod16(y.k4);
  const int mt = (m + 2 + 7) / 8;
  const int nb = (mt + 1) / 2;
  y.nblk = nb * (nb + 1) / 2;
  // split of the packed length: the (block, split) tasks should fill the warps evenly
  y.ksplit = 1;
  double best = 0;
  for (int sp = 1; sp <= 8; sp++) {
    const int tasks = y.nblk * sp;
    const int rounds = (tasks + kWarps - 1) / kWarps;
    const double eff = (double)tasks / (double)(rounds * kWarps);
    if (eff > best + 1e-9) {
      best = eff;
      y.ksplit = sp;
    }
  }
  const long a_size = (long)(m + 1) * n * y.pa + 16;
  const long p_size = (long)y.ksplit * y.nblk * 256;
  y.off_l = 64;
  y.off_x = y.off_l + (long)n * y.pl + 16;
  y.off_a = y.off_x + (long)(m + 2) * y.px;
  y.off_p = y.off_a;  // the partial Gram blocks reuse the operator's space
  y.total = y.off_a + (a_size > p_size ? a_size : p_size);
  return y;
}

__host__ inline bool Supported(int n, int m, size_t* smem_bytes) {
  if (n < 4 || n > 32 || (n % 4) != 0 || m < 1 || m + 2 > 1024) return false;
  const Layout y = MakeLayout(n, m);
  // the fallback (classic form, team code) runs in the same dynamic shared memory
  const long fallback = 64 + small::PsdSchurSmemDoubles(n, m, kThreads);
  const long total = y.total > fallback ? y.total : fallback;
  *smem_bytes = sizeof(double) * (size_t)total;
  return *smem_bytes <= 227 * 1024;
}

// The factor as the Schur kernel wants it in shared memory: L row-major with pitch pl, zeros above the diagonal, 16
// doubles of slack; one more double behind it is the "W is not positive definite" flag.
__host__ __device__ inline int LImageDoubles(int n) { return n * Pitch4Mod16(n) + 16; }

// One warp factors one W = L L^T (n <= 32; left-looking, lane = row, the image built in `sl`, then copied to `out`).
__device__ inline void PsdFactorWarp(int n, const double* __restrict__ W, double* sl, double* out) {
  const int lane = threadIdx.x & 31;
  const int pl = Pitch4Mod16(n);
  const int len = LImageDoubles(n);
  for (int e = lane; e < len; e += 32) {
    const int k = e / pl, c = e - k * pl;
    sl[e] = (k < n && c <= k && c < n) ? W[(long)c * n + k] : 0.0;
  }
  __syncwarp();
  const int r = lane;
  bool bad = false;
  for (int j = 0; j < n; j++) {
    double v0 = (r >= j && r < n) ? sl[r * pl + j] : 0.0, v1 = 0.0;
    if (r >= j && r < n) {
      int k = 0;
      for (; k + 1 < j; k += 2) {
        v0 -= sl[r * pl + k] * sl[j * pl + k];
        v1 -= sl[r * pl + k + 1] * sl[j * pl + k + 1];
      }
      if (k < j) v0 -= sl[r * pl + k] * sl[j * pl + k];
    }
    const double v = v0 + v1;
    const double d = __shfl_sync(0xffffffffu, v, j);
    if (!(d > 0.0)) {
      bad = true;
      break;
    }
    const double rd = sqrt(d);
    if (r >= j && r < n) sl[r * pl + j] = (r == j) ? rd : v / rd;
    __syncwarp();
  }
  __syncwarp();
  for (int e = lane; e < len; e += 32) out[e] = sl[e];
  if (lane == 0) out[len] = bad ? 1.0 : 0.0;
}

__device__ __forceinline__ void WaitPending(int pending) {
  // cp.async.wait_group takes an immediate: at most `pending` groups may still be in flight
  switch (pending < 0 ? 0 : (pending > 7 ? 7 : pending)) {
    case 0: CpAsyncWait<0>(); break;
    case 1: CpAsyncWait<1>(); break;
    case 2: CpAsyncWait<2>(); break;
    case 3: CpAsyncWait<3>(); break;
    case 4: CpAsyncWait<4>(); break;
    case 5: CpAsyncWait<5>(); break;
    case 6: CpAsyncWait<6>(); break;
    default: CpAsyncWait<7>(); break;
  }
}

// AC: (m + 1) column-major n x n matrices; W: n x n (global). sm: dynamic shared memory (MakeLayout(n, m).total
// doubles, at least the fallback's size). work: global scratch of the classic fallback.
template <int NT>  // NT = ceil(n / 8): 8-wide tiles per side
__device__ inline void PsdSchurMma(int n, int m, const double* __restrict__ AC, const double* __restrict__ W,
                                   const double* __restrict__ factor, double* work, double* sm, double* G, long ldg, double* AW, double* AQc,
                                   double* scal, bool acc) {
  const Layout y = MakeLayout(n, m);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int nn = n * n, pa = y.pa, pl = y.pl, px = y.px, kp = y.kp;
  double* sL = sm + y.off_l;   // L row-major: L(k, c) at sL[k * pl + c], zeros above the diagonal
  double* sX = sm + y.off_x;   // packed scaled matrices, row i at sX[i * px]
  double* sA = sm + y.off_a;   // matrix i, column c at sA[(i * n + c) * pa]
  double* sP = sm + y.off_p;

  // ---- phase 0: operator in flight, L = chol(W) ------------------------------------------------------------
  const int rounds = (m + 1 + kWarps - 1) / kWarps;
  // group 0: the Cholesky factor of W, computed by PsdFactorWarp in a launch of its own (one warp per program: inside
  // this CTA the 20 dependent columns of the factorisation would idle the other 15 warps and, at one CTA per SM, the SM)
  for (int q = tid; q < LImageDoubles(n) / 2; q += kThreads) CpAsync16(sL + 2 * q, factor + 2 * q, 16);
  CpAsyncCommit();
  {
    const int half = n / 2;  // 16-byte chunks per column
    for (int r = 0; r < rounds; r++) {
      const int i0 = r * kWarps;
      const int cnt = min(kWarps, m + 1 - i0);
      const int chunks = cnt * n * half;
      const double* src = AC + (long)i0 * nn;
      double* dst = sA + (long)i0 * n * pa;
      for (int q = tid; q < chunks; q += kThreads) {
        const int col = q / half, within = q - col * half;
        CpAsync16(dst + (long)col * pa + 2 * within, src + (long)col * n + 2 * within, 16);
      }
      CpAsyncCommit();
    }
  }
  // zero padding of the packed rows and the packed identity (row m + 1)
  for (int e = tid; e < (m + 2) * (px - kp); e += kThreads) {
    const int row = e / (px - kp), q = kp + e % (px - kp);
    sX[(long)row * px + q] = 0.0;
  }
  for (int e = tid; e < kp; e += kThreads) sX[(long)(m + 1) * px + e] = 0.0;
  __syncthreads();
  for (int c = tid; c < n; c += kThreads) sX[(long)(m + 1) * px + c * n - c * (c - 1) / 2] = 1.0;
  if (factor[LImageDoubles(n)] != 0.0) {  // W did not factor (uniform: one global word)
    CpAsyncWait<0>();
    __syncthreads();
    DeviceTeam t(sm);
    small::PsdSchurClassic(t, n, m, AC, W, work, sm + 64, G, ldg, AW, AQc, scal, acc);
    return;
  }

  // ---- phase A: S_i = L^T (A_i L), packed ------------------------------------------------------------------
  // the last of the NT 8-wide tiles per side may be ragged: its extra rows / columns re-read row / column n - 1
  // and are never stored
  const double kSqrt2 = 1.4142135623730951;
  for (int r = 0; r < rounds; r++) {
    WaitPending(rounds - 1 - r);  // group 0 (L) and the data groups 0..r have landed
    __syncthreads();
    const int i = r * kWarps + warp;
    if (i > m) continue;
    double* Ai = sA + (long)i * n * pa;
    double accT[NT][NT][2];
#pragma unroll
    for (int a = 0; a < NT; a++)
#pragma unroll
      for (int b = 0; b < NT; b++) accT[a][b][0] = accT[a][b][1] = 0.0;
    // T = A_i L: a = A(r0 + gid, k + tig) = Ai[(k + tig) * pa + r0 + gid] (A_i symmetric, column-major),
    //            b = L(k + tig, c0 + gid)
#pragma unroll
    for (int tc = 0; tc < NT; tc++) {
      for (int k = tc * 8; k < n; k += 4) {
        const double b = sL[(k + tig) * pl + min(tc * 8 + gid, n - 1)];
#pragma unroll
        for (int tr = 0; tr < NT; tr++) {
          const double a = Ai[(k + tig) * pa + min(tr * 8 + gid, n - 1)];
          Dmma884(accT[tr][tc][0], accT[tr][tc][1], a, b);
        }
      }
    }
    __syncwarp();
    // T(row, col) -> Ai[row * pa + col] (row = the k index of the next product)
#pragma unroll
    for (int tr = 0; tr < NT; tr++) {
#pragma unroll
      for (int tc = 0; tc < NT; tc++) {
        const int row = tr * 8 + gid, col = tc * 8 + tig * 2;
        if (row < n) {
          if (col < n) Ai[row * pa + col] = accT[tr][tc][0];
          if (col + 1 < n) Ai[row * pa + col + 1] = accT[tr][tc][1];
        }
      }
    }
    __syncwarp();
    // S = L^T T, tiles on or below the diagonal: a = L(k + tig, r0 + gid), b = T(k + tig, c0 + gid); L(k, r) = 0
    // for k < r: the k loop starts at the tile's first row
    double* Xi = sX + (long)i * px;
#pragma unroll
    for (int tr = 0; tr < NT; tr++) {
      double accS[NT][2];
#pragma unroll
      for (int b = 0; b < NT; b++) accS[b][0] = accS[b][1] = 0.0;
      for (int k = tr * 8; k < n; k += 4) {
        const double a = sL[(k + tig) * pl + min(tr * 8 + gid, n - 1)];
#pragma unroll
        for (int tc = 0; tc < NT; tc++) {
          if (tc > tr) break;
          const double b = Ai[(k + tig) * pa + min(tc * 8 + gid, n - 1)];
          Dmma884(accS[tc][0], accS[tc][1], a, b);
        }
      }
#pragma unroll
      for (int tc = 0; tc < NT; tc++) {
        if (tc > tr) break;
        const int row = tr * 8 + gid;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int col = tc * 8 + tig * 2 + e;
          if (row < n && col <= row) {
            Xi[col * n - col * (col - 1) / 2 + (row - col)] = (row == col) ? accS[tc][e] : kSqrt2 * accS[tc][e];
          }
        }
      }
    }
  }
  __syncthreads();  // X complete; the operator's space is free for the partial blocks

  // ---- phase B: Gram X X^T (lower), 16 x 16 blocks, packed length split across warps -----------------------
  const int MI = m + 2, MJ = m + 1;
  const int nb = ((MI + 7) / 8 + 1) / 2;
  const int ksteps = y.k4 / 4;
  const int per = (ksteps + y.ksplit - 1) / y.ksplit;
  for (int task = warp; task < y.nblk * y.ksplit; task += kWarps) {
    const int blk = task % y.nblk, ks = task / y.nblk;
    int bj = 0, rem = blk;
    while (rem >= nb - bj) {
      rem -= nb - bj;
      bj++;
    }
    const int bi = bj + rem;
    const double* xa0 = sX + (long)min(16 * bi + gid, MI - 1) * px + tig;
    const double* xa1 = sX + (long)min(16 * bi + 8 + gid, MI - 1) * px + tig;
    const double* xb0 = sX + (long)min(16 * bj + gid, MI - 1) * px + tig;
    const double* xb1 = sX + (long)min(16 * bj + 8 + gid, MI - 1) * px + tig;
    double c00[2] = {0, 0}, c01[2] = {0, 0}, c10[2] = {0, 0}, c11[2] = {0, 0};
    const int kbeg = ks * per, kend = min(ksteps, kbeg + per);
    for (int k = kbeg; k < kend; k++) {
      const double a0 = xa0[4 * k], a1 = xa1[4 * k], b0 = xb0[4 * k], b1 = xb1[4 * k];
      Dmma884(c00[0], c00[1], a0, b0);
      Dmma884(c01[0], c01[1], a0, b1);
      Dmma884(c10[0], c10[1], a1, b0);
      Dmma884(c11[0], c11[1], a1, b1);
    }
    // block stored column-major 16 x 16: entry (row, col) at [col * 16 + row]
    double* P = sP + ((long)ks * y.nblk + blk) * 256;
    const int col = tig * 2;
    P[(col + 0) * 16 + gid] = c00[0];
    P[(col + 1) * 16 + gid] = c00[1];
    P[(col + 8) * 16 + gid] = c01[0];
    P[(col + 9) * 16 + gid] = c01[1];
    P[(col + 0) * 16 + 8 + gid] = c10[0];
    P[(col + 1) * 16 + 8 + gid] = c10[1];
    P[(col + 8) * 16 + 8 + gid] = c11[0];
    P[(col + 9) * 16 + 8 + gid] = c11[1];
  }
  __syncthreads();
  for (int e = tid; e < y.nblk * 256; e += kThreads) {
    const int blk = e >> 8, in = e & 255;
    int bj = 0, rem = blk;
    while (rem >= nb - bj) {
      rem -= nb - bj;
      bj++;
    }
    const int bi = bj + rem;
    const int i = 16 * bi + (in & 15), j = 16 * bj + (in >> 4);
    if (i >= MI || j >= MJ || i < j) continue;
    double s = 0;
    for (int ks = 0; ks < y.ksplit; ks++) s += sP[((long)ks * y.nblk + blk) * 256 + in];
    if (i < m) {
      small::Accumulate(G + (long)j * ldg + i, s, acc);
    } else if (i == m) {
      small::Accumulate(j < m ? AQc + j : scal + 1, s, acc);
    } else {
      small::Accumulate(j < m ? AW + j : scal + 0, s, acc);
    }
  }
}

// ---- second layout: 8 warps per CTA, two CTAs per SM ---------------------------------------------------------------
// The kernel above keeps the whole operator of its program in shared memory (one CTA per SM), so its phases — wait
// for the data, scale, Gram, write — run strictly one after the other on that SM. Here a warp owns ONE slot of n x pa
// doubles: it streams its matrices i = warp, warp + 8, ... through it (cp.async by its own lanes, T = A_i L in
// place, S_i packed into X), so a CTA needs the packed X plus eight slots (100.7 KB at n = 20, m = 40) and two CTAs
// share an SM: while one waits for data or runs its Gram, the other scales. The Gram is not split along the packed
// length: a warp takes whole 16 x 16 blocks and writes its results straight from the accumulators.
constexpr int kWarps2 = 8;
constexpr int kThreads2 = kWarps2 * 32;

// (m + 1) mod 8 matrices are left for a last round in which only that many warps would work (m = 40: ONE warp, a
// sixth of the scaling phase with seven warps idle). When there are at most kCoopMax of them — and their extra slots
// still let two CTAs share an SM — they are loaded at the start and scaled by the whole CTA instead, one 8 x 8 tile of
// each product per warp.
constexpr int kCoopMax = 2;
struct Layout2 {
  int kp, k4, pa, pl, px, ncoop;
  long off_l, off_x, off_s, off_c, total;
};
__host__ __device__ inline Layout2 MakeLayout2(int n, int m) {
  Layout2 y;
  y.kp = n * (n + 1) / 2;
  y.k4 = (y.kp + 3) / 4 * 4;
  y.pa = Pitch4Mod16(n);
  y.pl = Pitch4Mod16(n);
  y.px = Pitch4Mod16(y.k4);
  y.off_l = 64;
  y.off_x = y.off_l + LImageDoubles(n);
  y.off_s = y.off_x + (long)(m + 2) * y.px;
  y.off_c = y.off_s + (long)kWarps2 * n * y.pa;
  y.ncoop = (m + 1) % kWarps2;
  if (y.ncoop > kCoopMax || m + 1 < kWarps2) y.ncoop = 0;
  if (sizeof(double) * (size_t)(y.off_c + (long)y.ncoop * n * y.pa + 16) > 113 * 1024) y.ncoop = 0;
  y.total = y.off_c + (long)y.ncoop * n * y.pa + 16;
  return y;
}
__host__ inline bool Supported2(int n, int m, size_t* smem_bytes) {
  if (n < 4 || n > 32 || (n % 4) != 0 || m < 1 || m + 2 > 1024) return false;
  const Layout2 y = MakeLayout2(n, m);
  const long fallback = 64 + small::PsdSchurSmemDoubles(n, m, kThreads2);
  const long total = y.total > fallback ? y.total : fallback;
  *smem_bytes = sizeof(double) * (size_t)total;
  return *smem_bytes <= 113 * 1024;  // two CTAs per SM
}

// N4 = n / 4: the order of the block is a compile-time constant, so the k loops of the two products unroll completely
// (their shared-memory loads are hoisted ahead of the DMMA chains instead of being issued and waited for k-step by
// k-step) and the index arithmetic of the loads, of the cp.async chunks and of the packing folds into immediates.
template <int N4>
__device__ inline void PsdSchurMma2(int m, const double* __restrict__ AC, const double* __restrict__ W,
                                    const double* __restrict__ factor, double* work, double* sm, double* G, long ldg,
                                    double* AW, double* AQc, double* scal, bool acc) {
  constexpr int n = 4 * N4, NT = (n + 7) / 8, nn = n * n, half = n / 2;
  constexpr int pa = Pitch4Mod16(n), pl = Pitch4Mod16(n);
  const Layout2 y = MakeLayout2(n, m);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int px = y.px, kp = y.kp;
  double* sL = sm + y.off_l;
  double* sX = sm + y.off_x;
  double* Ai = sm + y.off_s + (long)warp * n * pa;  // this warp's slot
  auto fetch = [&](int i) {  // matrix i -> the slot, by this warp's lanes
    const double* src = AC + (long)i * nn;
#pragma unroll
    for (int q0 = 0; q0 < n * half; q0 += 32) {
      const int q = q0 + lane;
      if (q < n * half) {
        const int col = q / half, within = q - col * half;
        CpAsync16(Ai + col * pa + 2 * within, src + col * n + 2 * within, 16);
      }
    }
    CpAsyncCommit();
  };
  for (int q = tid; q < LImageDoubles(n) / 2; q += kThreads2) CpAsync16(sL + 2 * q, factor + 2 * q, 16);
  const int mfull = m + 1 - y.ncoop;  // matrices 0 .. mfull-1 go through the warps' slots, the rest is shared work
  for (int q = tid; q < y.ncoop * n * half; q += kThreads2) {
    const int which = q / (n * half), e = q - which * n * half;
    const int col = e / half, within = e - col * half;
    CpAsync16(sm + y.off_c + ((long)which * n + col) * pa + 2 * within,
              AC + (long)(mfull + which) * nn + (long)col * n + 2 * within, 16);
  }
  CpAsyncCommit();
  if (warp < mfull) fetch(warp);
  for (int e = tid; e < (m + 2) * (px - kp); e += kThreads2) {
    const int row = e / (px - kp), q = kp + e % (px - kp);
    sX[(long)row * px + q] = 0.0;
  }
  for (int e = tid; e < kp; e += kThreads2) sX[(long)(m + 1) * px + e] = 0.0;
  __syncthreads();
  for (int c = tid; c < n; c += kThreads2) sX[(long)(m + 1) * px + c * n - c * (c - 1) / 2] = 1.0;
  CpAsyncWait<0>();
  __syncthreads();  // L, the first matrix of every warp, the shared matrices, the padding of X
  if (factor[LImageDoubles(n)] != 0.0) {
    DeviceTeam t(sm);
    small::PsdSchurClassic(t, n, m, AC, W, work, sm + 64, G, ldg, AW, AQc, scal, acc);
    return;
  }
  const double kSqrt2 = 1.4142135623730951;
  // fragment rows of the 8-row tiles, clamped into the block (rows >= n are computed and never stored)
  int frow[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) frow[t] = min(t * 8 + gid, n - 1);
  for (int i = warp; i < mfull; i += kWarps2) {
    // The slot is busy until S_i is done, so the next matrix cannot travel to shared memory yet: ask for it in L2 now,
    // and the cp.async issued after the second product pays an L2 round trip instead of a DRAM one.
    if (i + kWarps2 < mfull) {
      const double* next = AC + (long)(i + kWarps2) * nn;
      for (int off = lane * 16; off < nn; off += 512) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(next + off));
      }
    }
    double accT[NT][NT][2];
#pragma unroll
    for (int a = 0; a < NT; a++)
#pragma unroll
      for (int b = 0; b < NT; b++) accT[a][b][0] = accT[a][b][1] = 0.0;
#pragma unroll
    for (int tc = 0; tc < NT; tc++) {
#pragma unroll
      for (int k = tc * 8; k < n; k += 4) {
        const double b = sL[(k + tig) * pl + frow[tc]];
#pragma unroll
        for (int tr = 0; tr < NT; tr++) {
          const double a = Ai[(k + tig) * pa + frow[tr]];
          Dmma884(accT[tr][tc][0], accT[tr][tc][1], a, b);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int tr = 0; tr < NT; tr++) {
#pragma unroll
      for (int tc = 0; tc < NT; tc++) {
        const int row = tr * 8 + gid, col = tc * 8 + tig * 2;
        if (row < n) {
          if (col < n) Ai[row * pa + col] = accT[tr][tc][0];
          if (col + 1 < n) Ai[row * pa + col + 1] = accT[tr][tc][1];
        }
      }
    }
    __syncwarp();
    double accS[NT][NT][2];
#pragma unroll
    for (int tr = 0; tr < NT; tr++) {
#pragma unroll
      for (int b = 0; b < NT; b++) accS[tr][b][0] = accS[tr][b][1] = 0.0;
#pragma unroll
      for (int k = tr * 8; k < n; k += 4) {
        const double a = sL[(k + tig) * pl + frow[tr]];
#pragma unroll
        for (int tc = 0; tc <= tr; tc++) {
          const double b = Ai[(k + tig) * pa + frow[tc]];
          Dmma884(accS[tr][tc][0], accS[tr][tc][1], a, b);
        }
      }
    }
    __syncwarp();  // the slot is free: the next matrix travels while S_i is packed
    if (i + kWarps2 < mfull) fetch(i + kWarps2);
    double* Xi = sX + (long)i * px;
#pragma unroll
    for (int tr = 0; tr < NT; tr++) {
#pragma unroll
      for (int tc = 0; tc <= tr; tc++) {
        const int row = tr * 8 + gid;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int col = tc * 8 + tig * 2 + e;
          if (row < n && col <= row) {
            Xi[col * n - col * (col - 1) / 2 + (row - col)] =
                (row == col) ? accS[tr][tc][e] : kSqrt2 * accS[tr][tc][e];
          }
        }
      }
    }
    CpAsyncWait<0>();
    __syncwarp();
  }
  __syncthreads();  // rows 0 .. mfull-1 of X complete; every warp's slot is free
  for (int q = 0; q < y.ncoop; q++) {
    // the same two products, tile by tile: T = A L into the (free) slot of warp q, then S = L^T T into row mfull + q
    const double* Aq = sm + y.off_c + (long)q * n * pa;
    double* Tq = sm + y.off_s + (long)q * n * pa;
    for (int task = warp; task < NT * NT; task += kWarps2) {
      const int tr = task / NT, tc = task - tr * NT;
      const int ra = min(tr * 8 + gid, n - 1), rb = min(tc * 8 + gid, n - 1);
      double t0 = 0.0, t1 = 0.0;
      for (int k = tc * 8; k < n; k += 4) {
        const double b = sL[(k + tig) * pl + rb];
        const double a = Aq[(k + tig) * pa + ra];
        Dmma884(t0, t1, a, b);
      }
      const int row = tr * 8 + gid, col = tc * 8 + tig * 2;
      if (row < n) {
        if (col < n) Tq[row * pa + col] = t0;
        if (col + 1 < n) Tq[row * pa + col + 1] = t1;
      }
    }
    __syncthreads();
    double* Xi = sX + (long)(mfull + q) * px;
    for (int task = warp; task < NT * (NT + 1) / 2; task += kWarps2) {
      int tr = 0, tc = task;
      while (tc > tr) {
        tc -= tr + 1;
        tr++;
      }
      const int ra = min(tr * 8 + gid, n - 1), rb = min(tc * 8 + gid, n - 1);
      double s0 = 0.0, s1 = 0.0;
      for (int k = tr * 8; k < n; k += 4) {
        const double a = sL[(k + tig) * pl + ra];
        const double b = Tq[(k + tig) * pa + rb];
        Dmma884(s0, s1, a, b);
      }
      const int row = tr * 8 + gid;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int col = tc * 8 + tig * 2 + e;
        if (row < n && col <= row) {
          Xi[col * n - col * (col - 1) / 2 + (row - col)] = (row == col) ? (e ? s1 : s0) : kSqrt2 * (e ? s1 : s0);
        }
      }
    }
  }
  if (y.ncoop) __syncthreads();  // X complete

  // Gram X X^T (lower) in 16 x 16 blocks of four DMMA tiles, one block per warp and pass, results written straight
  // from the accumulators. (Dealing 8 x 8 tiles evenly to the warps was tried: the per-tile pointer selection cost more
  // issue slots than the better balance returned — 391 vs 353 us per launch, profiles/r02_l_*.)
  const int MI = m + 2, MJ = m + 1;
  const int nb = ((MI + 7) / 8 + 1) / 2;
  const int nblk = nb * (nb + 1) / 2;
  constexpr int ksteps = (n * (n + 1) / 2 + 3) / 4;
  for (int blk = warp; blk < nblk; blk += kWarps2) {
    int bj = 0, rem = blk;
    while (rem >= nb - bj) {
      rem -= nb - bj;
      bj++;
    }
    const int bi = bj + rem;
    const double* xa0 = sX + (long)min(16 * bi + gid, MI - 1) * px + tig;
    const double* xa1 = sX + (long)min(16 * bi + 8 + gid, MI - 1) * px + tig;
    const double* xb0 = sX + (long)min(16 * bj + gid, MI - 1) * px + tig;
    const double* xb1 = sX + (long)min(16 * bj + 8 + gid, MI - 1) * px + tig;
    double c00[2] = {0, 0}, c01[2] = {0, 0}, c10[2] = {0, 0}, c11[2] = {0, 0};
    const bool off_diagonal = bi != bj;
#pragma unroll 4
    for (int k = 0; k < ksteps; k++) {
      const double a0 = xa0[4 * k], a1 = xa1[4 * k], b0 = xb0[4 * k], b1 = xb1[4 * k];
      Dmma884(c00[0], c00[1], a0, b0);
      if (off_diagonal) Dmma884(c01[0], c01[1], a0, b1);  // above the diagonal inside a diagonal block
      Dmma884(c10[0], c10[1], a1, b0);
      Dmma884(c11[0], c11[1], a1, b1);
    }
    auto put = [&](int i, int j, double s) {
      if (i >= MI || j >= MJ || i < j) return;
      if (i < m) {
        small::Accumulate(G + (long)j * ldg + i, s, acc);
      } else if (i == m) {
        small::Accumulate(j < m ? AQc + j : scal + 1, s, acc);
      } else {
        small::Accumulate(j < m ? AW + j : scal + 0, s, acc);
      }
    };
    const int i0 = 16 * bi + gid, j0 = 16 * bj + tig * 2;
    put(i0, j0, c00[0]);
    put(i0, j0 + 1, c00[1]);
    put(i0, j0 + 8, c01[0]);
    put(i0, j0 + 9, c01[1]);
    put(i0 + 8, j0, c10[0]);
    put(i0 + 8, j0 + 1, c10[1]);
    put(i0 + 8, j0 + 8, c11[0]);
    put(i0 + 8, j0 + 9, c11[1]);
  }
}

}  // namespace psdmma
}  // namespace cxb
