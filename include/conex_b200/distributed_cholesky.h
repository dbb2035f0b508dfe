// K3 across GPUs: right-looking Cholesky of the replicated Schur complement with block columns
// dealt 1-D block-cyclically to the ranks and every factored panel broadcast over NCCL
// (NVLink 5 / NVSwitch). The reference factors on one thread (block_triangular_operations.cc:184-219);
// SURVEY.md §8(e) items 4-5 ask for the multi-GPU form.
//
// Layout: H (N x N, lower triangle, leading dimension ld) is bit-identical on every rank when Factor()
// is entered (it comes out of one ncclAllReduce, or out of replicated deterministic kernels). Block
// column J = columns [J * block, (J + 1) * block) belongs to rank J % world. Only the owner updates
// and factors a block column; the factored panel is then broadcast into the same place of every
// rank's H, so on return every rank holds the complete factor L and the triangular solves (K5: HBM
// bound, m^2 * 8 B per sweep, no flops to share) run replicated without any further communication.
//
// Schedule (one list per rank, CholeskySchedule): panel J + 1 is updated, factored and put on the wire
// by its owner BEFORE the owner's other block columns receive panel J's update (look-ahead of one
// panel), and broadcasts run on a side stream, so the transfer and the panel's latency-bound
// kernels hide under the other ranks' DMMA trailing updates.
//
// Why columns 1-D and not the 2-D grid of the north-star sketch: a 2-D layout cuts the broadcast
// volume per rank from N^2/2 to N^2/sqrt(P) doubles, which matters on a network whose bisection is
// shared; through NVSwitch every rank receives at full link rate regardless of who sends, and the
// whole factor is 1.6 GB at N = 20000 (about 4 ms of link time against 80 ms of flops / P). The 1-D
// layout keeps every trailing update a single tall DMMA GEMM per owned block column.
#pragma once
#include <cuda_runtime_api.h>

#include <vector>

namespace conex {

inline int CholeskyOwner(int block_column, int world) { return block_column % world; }

struct CholeskyOp {
  enum Kind : int {
    kFactor = 0,     // factor panel `panel` in place (owner only)
    kBroadcast = 1,  // broadcast panel `panel` from rank `target` (every rank, side stream)
    kWait = 2,       // main stream waits until panel `panel` has arrived
    kUpdate = 3,     // block column `target` -= L[:, panel] L[target rows, panel]^T (owner of `target`)
  };
  int kind;
  int panel;
  int target;
};

// The ordered list of operations rank `rank` enqueues to factor an N x N matrix cut into block columns
// of `block` columns. Pure host logic (tests/test_sharding.py runs it under gloo).
std::vector<CholeskyOp> CholeskySchedule(int N, int block, int world, int rank);

// Process-wide policy: matrices of order < min_order are factored replicated (the panel chain is
// latency bound there and a broadcast per panel only adds to it).
struct DistributedCholeskyPolicy {
  int min_order = 4096;
  int block = 512;
};
DistributedCholeskyPolicy& DistributedCholeskyConfig();

class DistributedCholesky {
 public:
  DistributedCholesky() = default;
  ~DistributedCholesky();
  DistributedCholesky(const DistributedCholesky&) = delete;
  DistributedCholesky& operator=(const DistributedCholesky&) = delete;

  // Collective over Communicator::Get(). Enqueues everything on `stream` (and a private side
  // stream that `stream` has joined again on return). d_info: device int, 0 on success, non-zero on
  // every rank when any panel met a non-positive pivot (Eigen::LLT::info() != Success).
  void Factor(cudaStream_t stream, int N, double* dH, long ld, int* d_info, int block);

 private:
  void Prepare();
  cudaStream_t side_ = nullptr;
  cudaEvent_t factored_[2] = {nullptr, nullptr};
  cudaEvent_t arrived_[2] = {nullptr, nullptr};
};

}  // namespace conex
