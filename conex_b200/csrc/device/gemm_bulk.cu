// A/B arm of the DMMA GEMM with its operand tiles staged by the TMA engine's BULK copies
// (cp.async.bulk.shared.global + mbarrier complete_tx; SASS: UBLKCP / SYNCS) instead of per-thread cp.async (LDGSTS).
//
// BASELINE's north-star sketch names TMA for the A_i tiles. The tensor form (cp.async.bulk.tensor with a 2-D box)
// writes dense or XOR-swizzled rows; for 8-byte elements neither gives the pitch == 4 (mod 16) doubles that makes the
// 8 x 4 / 4 x 8 DMMA fragment loads conflict-free (dense: 4- and 8-way conflicts; SWIZZLE_128B: 2-way), so it would
// need a second shuffle pass through shared memory. The bulk form used here copies one contiguous tile row (512 B of
// A, 128 B of B) per instruction to an ARBITRARY 16-byte aligned shared address, i.e. straight into the padded
// layout of gemm.cu: same fragment loads, same arithmetic order (bit-identical results), 80 copy instructions per
// k-tile and CTA issued by one warp instead of 1024 LDGSTS spread over all threads, completion through an mbarrier.
// Plain NN case only (A column-major M x K, B column-major K x N, optional tri = 1: op(B) lower triangular) — the
// shape of K1's first GEMM, T_i = A_i L; everything else stays on gemm.cu. Measured result: DESIGN.md section 3.1.
#include "common.cuh"
#include "device_api.h"

namespace cxb {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, kStages = 2, kThreadsBulk = 128;
constexpr int kPitchA = BM + 4;  // As[k][m], 16 rows
constexpr int kPitchB = BK + 4;  // Bs[n][k], 64 rows
constexpr int kStageA = BK * kPitchA, kStageB = BN * kPitchB;
constexpr size_t kBulkSmem = sizeof(double) * kStages * (kStageA + kStageB) + 64;

struct BulkArgs {
  int M, N, K;
  const double* A;
  long lda, sA;
  const double* B;
  long ldb, sB;
  double* C;
  long ldc, sC;
  int tiles_m, tri;
};

__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void MbarInit(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(SmemAddr(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void BulkCopy(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   SmemAddr(smem)),
               "l"(gmem), "r"(bytes), "r"(SmemAddr(bar))
               : "memory");
}

__global__ void __launch_bounds__(kThreadsBulk, 4) DgemmBulkKernel(const BulkArgs g) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + kStages * kStageA;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * (kStageA + kStageB));
  const int tm = blockIdx.x % g.tiles_m, tn = blockIdx.x / g.tiles_m;
  const int m0 = tm * BM, n0 = tn * BN;
  const double* A = g.A + (long)blockIdx.z * g.sA;
  const double* B = g.B + (long)blockIdx.z * g.sB;
  double* C = g.C + (long)blockIdx.z * g.sC;
  int k_begin = 0;
  if (g.tri & 1) k_begin = (n0 / BK) * BK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp % 2) * 32, wn0 = (warp / 2) * 32;
  const int gid = lane >> 2, tig = lane & 3;
  const int rows_m = min(BM, g.M - m0), rows_n = min(BN, g.N - n0);
  // ragged edges: the bulk copies only ever write the valid part, the rest of both stages stays zero
  for (int e = tid; e < kStages * (kStageA + kStageB); e += kThreadsBulk) smem[e] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) MbarInit(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill before async-proxy writes
  const int KT = (g.K - k_begin + BK - 1) / BK;
  const uint32_t tx_bytes = (uint32_t)(BK * rows_m * 8 + rows_n * BK * 8);
  auto issue = [&](int kt, int stage) {  // warp 0: 16 column copies of A, 64 row copies of B
    const int k0 = k_begin + kt * BK;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of this stage
    if (lane == 0) MbarExpectTx(full + stage, tx_bytes);
    __syncwarp();
    if (lane < BK) {
      BulkCopy(As + stage * kStageA + lane * kPitchA, A + (long)(k0 + lane) * g.lda + m0, (uint32_t)(rows_m * 8),
               full + stage);
    }
    for (int r = lane; r < rows_n; r += 32) {
      BulkCopy(Bs + stage * kStageB + r * kPitchB, B + (long)(n0 + r) * g.ldb + k0, (uint32_t)(BK * 8), full + stage);
    }
  };
  if (warp == 0) {
    for (int s = 0; s < kStages - 1 && s < KT; s++) issue(s, s);
  }
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int a_off = tig * kPitchA + wm0 + gid;
  const int b_off = (wn0 + gid) * kPitchB + tig;
  for (int kt = 0; kt < KT; kt++) {
    const int stage = kt % kStages;
    MbarWait(full + stage, (uint32_t)((kt / kStages) & 1));
    __syncthreads();  // every warp is done with the stage that is refilled next
    if (warp == 0 && kt + kStages - 1 < KT) issue(kt + kStages - 1, (kt + kStages - 1) % kStages);
    const double* as = As + stage * kStageA + a_off;
    const double* bs = Bs + stage * kStageB + b_off;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = as[i * 8 + kk * kPitchA];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = bs[j * 8 * kPitchB + kk];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) Dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + wm0 + i * 8 + gid;
    if (r >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = n0 + wn0 + j * 8 + tig * 2 + e;
        if (c < g.N) C[(long)c * g.ldc + r] = acc[i][j][e];
      }
    }
  }
}

}  // namespace
}  // namespace cxb

// C = A B for a strided batch; tri = 1: B(k, c) = 0 for k < c (k-tiles above B's diagonal skipped). Requirements of the
// bulk copies: K % 16 == 0, M and the leading dimensions / batch strides even, 16-byte aligned bases.
extern "C" int cxb_dgemm_bulk(void* stream, int M, int N, int K, const double* dA, long lda, long strideA,
                              const double* dB, long ldb, long strideB, double* dC, long ldc, long strideC, int batch,
                              int tri) {
  using namespace cxb;
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if (K <= 0 || (K % BK) != 0 || (M & 1) || (lda & 1) || (ldb & 1) || (strideA & 1) || (strideB & 1) ||
      (reinterpret_cast<uintptr_t>(dA) & 15) || (reinterpret_cast<uintptr_t>(dB) & 15)) {
    return -1;
  }
  static std::atomic<unsigned long long> configured{0};
  if (FirstUseOnCurrentDevice(configured)) {
    cudaFuncSetAttribute(DgemmBulkKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBulkSmem);
  }
  BulkArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.A = dA;
  g.lda = lda;
  g.sA = strideA;
  g.B = dB;
  g.ldb = ldb;
  g.sB = strideB;
  g.C = dC;
  g.ldc = ldc;
  g.sC = strideC;
  g.tiles_m = (M + BM - 1) / BM;
  g.tri = tri;
  const int tiles_n = (N + BN - 1) / BN;
  dim3 grid((unsigned)(g.tiles_m * tiles_n), 1, (unsigned)batch);
  CountLaunch(); DgemmBulkKernel<<<grid, kThreadsBulk, kBulkSmem, AsStream(stream)>>>(g);
  return LaunchStatus();
}
