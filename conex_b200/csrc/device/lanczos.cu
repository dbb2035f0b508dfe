// K7: two-sided Lanczos on (WS, WS^T) in the W inner product — the step-size / mu eigen-bound
// estimate. Replaces AsymmetricLanczos (approximate_eigenvalues.cc:173-239) with the same
// recurrence, start vector, iteration count and breakdown test (beta^2 < 1e-6):
//
//   v1 = r; v0 = W r; V /= sqrt(v0.v1)
//   step j: u0 = WS v0; u1 = WS^T v1; alpha_j = v0.u1; U -= alpha_j V (+ beta_{j-1} Vprev)
//           beta_j^2 = u0.u1; stop if < 1e-6; Vprev = V; V = U / beta_j
//
// Bandwidth-bound: each step streams WS and WS^T once (both column-dot products, one warp per
// column, fully coalesced; for n = 2000 the two matrices stay resident in the 126 MB L2). The
// n-vector recurrence is executed by the last CTA to finish (ticket counter), so one launch per
// step suffices and the arithmetic order is fixed (deterministic).
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"
#include "device_api.h"

namespace cxb {
namespace {

struct LanczosState {
  int done;
  int count;
  unsigned int ticket;
  int pad;
};

// Layout of d_work (doubles): [WST n*n][v0][v1][u0][u1][p0][p1][state (4 doubles)]
struct LanczosBufs {
  double *WST, *v0, *v1, *u0, *u1, *p0, *p1;
  LanczosState* st;
};

__host__ __device__ inline LanczosBufs Carve(double* work, int n) {
  LanczosBufs b;
  const long nn = (long)n * n;
  const long np = (n + 3) & ~3L;
  b.WST = work;
  b.v0 = work + nn;
  b.v1 = b.v0 + np;
  b.u0 = b.v1 + np;
  b.u1 = b.u0 + np;
  b.p0 = b.u1 + np;
  b.p1 = b.p0 + np;
  b.st = reinterpret_cast<LanczosState*>(b.p1 + np);
  return b;
}

__device__ __forceinline__ double ColumnDot(const double* __restrict__ col,
                                            const double* __restrict__ x, int n, int lane) {
  double s = 0;
  for (int i = lane; i < n; i += 32) s += col[i] * x[i];
  return WarpSum(s);
}

// v0 = W^T r (= W r, W symmetric); the last CTA normalises: V /= sqrt(v0 . r).
__global__ void __launch_bounds__(256) LanczosInitKernel(int n, const double* __restrict__ W,
                                                         const double* __restrict__ rbase,
                                                         const double* __restrict__ col_index,
                                                         double* work) {
  __shared__ double scratch[33];
  __shared__ bool is_last;
  const LanczosBufs b = Carve(work, n);
  const double* __restrict__ r = col_index ? rbase + (long)(*col_index) * n : rbase;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c < n) {
    const double v = ColumnDot(W + (long)c * n, r, n, lane);
    if (lane == 0) b.v0[c] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&b.st->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(b.v0 + i) * r[i];
  s = BlockSum(s, scratch);
  const double scale = sqrt(s);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    b.v0[i] = __ldcg(b.v0 + i) / scale;
    b.v1[i] = r[i] / scale;
  }
  if (threadIdx.x == 0) {
    b.st->ticket = 0;
    b.st->done = 0;
    b.st->count = 0;
  }
}

__global__ void LanczosResetKernel(int n, double* work, int* count) {
  const LanczosBufs b = Carve(work, n);
  b.st->ticket = 0;
  b.st->done = 0;
  b.st->count = 0;
  count[0] = 0;
  count[1] = 0;  // set to 1 by the step that stops the recurrence
}

// One Lanczos step (index j).
// rel_tol == 0: the dense-LMI rule (stop when beta^2 < 1e-6, approximate_eigenvalues.cc:208);
// rel_tol > 0: the Hermitian rule (stop when beta^2 < rel_tol * <U, U> of the first, not yet
// orthogonalised, step — jordan_matrix_algebra.cc:421-433).
__global__ void __launch_bounds__(256) LanczosStepKernel(int n, int j, int num_iter,
                                                         const double* __restrict__ WS, double* work,
                                                         double* alpha, double* beta, int* count,
                                                         double rel_tol) {
  __shared__ double scratch[33];
  __shared__ bool is_last;
  const LanczosBufs b = Carve(work, n);
  if (b.st->done) return;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c < n) {
    const double a0 = ColumnDot(b.WST + (long)c * n, b.v0, n, lane);  // (WS v0)[c]
    const double a1 = ColumnDot(WS + (long)c * n, b.v1, n, lane);     // (WS^T v1)[c]
    if (lane == 0) {
      b.u0[c] = a0;
      b.u1[c] = a1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&b.st->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- n-vector recurrence, executed by one CTA in a fixed order ----
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += b.v0[i] * __ldcg(b.u1 + i);
  const double a = BlockSum(s, scratch);
  double* scaling = reinterpret_cast<double*>(b.st) + 2;
  if (rel_tol > 0 && j == 0) {
    s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(b.u0 + i) * __ldcg(b.u1 + i);
    const double raw = BlockSum(s, scratch);
    if (threadIdx.x == 0) *scaling = raw;
    __syncthreads();
  }
  const double threshold = (rel_tol > 0) ? rel_tol * (*scaling) : 1e-6;
  const double bprev = (j > 0) ? beta[j - 1] : 0.0;
  s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double x0 = __ldcg(b.u0 + i) - a * b.v0[i];
    double x1 = __ldcg(b.u1 + i) - a * b.v1[i];
    if (j > 0) {
      x0 -= bprev * b.p0[i];
      x1 -= bprev * b.p1[i];
    }
    b.u0[i] = x0;
    b.u1[i] = x1;
    s += x0 * x1;
  }
  const double b2 = BlockSum(s, scratch);
  bool stop = (j + 1 >= num_iter) || (b2 < threshold);
  if (!stop) {
    const double bj = sqrt(b2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double x0 = b.u0[i], x1 = b.u1[i];
      b.p0[i] = b.v0[i];
      b.p1[i] = b.v1[i];
      b.v0[i] = x0 / bj;
      b.v1[i] = x1 / bj;
    }
    if (threadIdx.x == 0) beta[j] = bj;
  }
  if (threadIdx.x == 0) {
    alpha[j] = a;
    b.st->ticket = 0;
    b.st->count = j;
    *count = j;
    if (stop) {
      b.st->done = 1;
      count[1] = 1;
    }
  }
}

}  // namespace

namespace {

// The chain of one transpose + one reset + one init + num_iter step launches.
int EnqueueLanczos(cudaStream_t s, int n, const double* d_WS, const double* d_W, const double* d_r,
                   const double* d_col_index, int num_iter, double* d_alpha, double* d_beta, int* d_count,
                   double* d_work, double rel_tol, int j_begin, int j_end) {
  const LanczosBufs b = Carve(d_work, n);
  const int grid = (n + 7) / 8;
  if (j_begin == 0) {
    int rc = Transpose(s, n, d_WS, b.WST);
    if (rc) return rc;
    CountLaunch(); LanczosResetKernel<<<1, 1, 0, s>>>(n, d_work, d_count);
    CountLaunch(); LanczosInitKernel<<<grid, 256, 0, s>>>(n, d_W, d_r, d_col_index, d_work);
  }
  for (int j = j_begin; j < j_end; j++) {
    CountLaunch(); LanczosStepKernel<<<grid, 256, 0, s>>>(n, j, num_iter, d_WS, d_work, d_alpha, d_beta, d_count, rel_tol);
  }
  return LaunchStatus();
}

// Long chains (n/2 dependent launches of ~10 us kernels) are launch-bound from the host: they are
// captured once per argument tuple into a CUDA graph and replayed with a single launch, so the
// device runs the whole recurrence back to back whatever the host is doing.
struct GraphKey {
  int n, num_iter, j_begin, j_end;
  const void *ws, *w, *r, *col, *alpha, *beta, *count, *work;
  double rel_tol;
  bool operator<(const GraphKey& o) const {
    return std::memcmp(this, &o, sizeof(GraphKey)) < 0;
  }
};
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  long launches = 0;
};
std::map<GraphKey, GraphEntry> g_graphs;
std::mutex g_graph_mutex;
constexpr int kGraphThreshold = 32;   // shorter chains are launched directly
constexpr size_t kMaxGraphs = 64;

}  // namespace
}  // namespace cxb

using namespace cxb;

extern "C" {

size_t cxb_lanczos_worksize(int n) { return (size_t)n * n + 6 * (size_t)((n + 3) & ~3) + 8; }

int cxb_lanczos_two_sided(void* stream, int n, const double* d_WS, const double* d_W,
                          const double* d_r, const double* d_col_index, int num_iter,
                          double* d_alpha, double* d_beta, int* d_count, double* d_work) {
  return cxb_lanczos_two_sided_ex(stream, n, d_WS, d_W, d_r, d_col_index, num_iter, d_alpha, d_beta, d_count,
                                  d_work, 0.0);
}

int cxb_lanczos_two_sided_ex(void* stream, int n, const double* d_WS, const double* d_W,
                             const double* d_r, const double* d_col_index, int num_iter,
                             double* d_alpha, double* d_beta, int* d_count, double* d_work,
                             double rel_tol) {
  return cxb_lanczos_two_sided_range(stream, n, d_WS, d_W, d_r, d_col_index, num_iter, 0, num_iter, d_alpha,
                                     d_beta, d_count, d_work, rel_tol);
}

int cxb_lanczos_two_sided_range(void* stream, int n, const double* d_WS, const double* d_W,
                                const double* d_r, const double* d_col_index, int num_iter, int j_begin,
                                int j_end, double* d_alpha, double* d_beta, int* d_count, double* d_work,
                                double rel_tol) {
  cudaStream_t s = AsStream(stream);
  if (n < 1 || num_iter < 1 || j_begin < 0 || j_end > num_iter || j_begin > j_end) return -1;
  if (j_end - j_begin < kGraphThreshold || s == nullptr) {
    return EnqueueLanczos(s, n, d_WS, d_W, d_r, d_col_index, num_iter, d_alpha, d_beta, d_count, d_work, rel_tol, j_begin, j_end);
  }
  GraphKey key;
  std::memset(&key, 0, sizeof(key));
  key.n = n;
  key.num_iter = num_iter;
  key.j_begin = j_begin;
  key.j_end = j_end;
  key.ws = d_WS;
  key.w = d_W;
  key.r = d_r;
  key.col = d_col_index;
  key.alpha = d_alpha;
  key.beta = d_beta;
  key.count = d_count;
  key.work = d_work;
  key.rel_tol = rel_tol;
  std::lock_guard<std::mutex> lock(g_graph_mutex);
  auto it = g_graphs.find(key);
  if (it == g_graphs.end()) {
    if (g_graphs.size() >= kMaxGraphs) {
      for (auto& e : g_graphs) cudaGraphExecDestroy(e.second.exec);
      g_graphs.clear();
    }
    const long before = g_launch_count.load();
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return EnqueueLanczos(s, n, d_WS, d_W, d_r, d_col_index, num_iter, d_alpha, d_beta, d_count, d_work, rel_tol, j_begin, j_end);
    }
    const int rc = EnqueueLanczos(s, n, d_WS, d_W, d_r, d_col_index, num_iter, d_alpha, d_beta, d_count, d_work, rel_tol, j_begin, j_end);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &graph);
    GraphEntry entry;
    entry.launches = g_launch_count.load() - before;
    g_launch_count.fetch_sub(entry.launches);
    if (rc != 0 || e != cudaSuccess || graph == nullptr ||
        cudaGraphInstantiate(&entry.exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return EnqueueLanczos(s, n, d_WS, d_W, d_r, d_col_index, num_iter, d_alpha, d_beta, d_count, d_work, rel_tol, j_begin, j_end);
    }
    cudaGraphDestroy(graph);
    it = g_graphs.emplace(key, entry).first;
  }
  g_launch_count.fetch_add(it->second.launches);
  const cudaError_t e = cudaGraphLaunch(it->second.exec, s);
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // extern "C"
