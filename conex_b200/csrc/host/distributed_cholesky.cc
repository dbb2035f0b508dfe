#include "distributed_cholesky.h"

#include <algorithm>

#include "../../../include/conex_b200_device.h"
#include "communicator.h"
#include "device_runtime.h"

namespace conex {

std::vector<CholeskyOp> CholeskySchedule(int N, int block, int world, int rank) {
  std::vector<CholeskyOp> ops;
  if (N <= 0 || block <= 0 || world <= 0) return ops;
  const int nblk = (N + block - 1) / block;
  auto mine = [&](int J) { return CholeskyOwner(J, world) == rank; };
  if (mine(0)) ops.push_back({CholeskyOp::kFactor, 0, 0});
  ops.push_back({CholeskyOp::kBroadcast, 0, CholeskyOwner(0, world)});
  for (int J = 0; J < nblk; J++) {
    ops.push_back({CholeskyOp::kWait, J, 0});
    const int next = J + 1;
    if (next < nblk) {
      // look-ahead: the next panel leaves its owner before the owner's other columns are updated
      if (mine(next)) {
        ops.push_back({CholeskyOp::kUpdate, J, next});
        ops.push_back({CholeskyOp::kFactor, next, 0});
      }
      ops.push_back({CholeskyOp::kBroadcast, next, CholeskyOwner(next, world)});
    }
    for (int K = next + 1; K < nblk; K++) {
      if (mine(K)) ops.push_back({CholeskyOp::kUpdate, J, K});
    }
  }
  return ops;
}

DistributedCholeskyPolicy& DistributedCholeskyConfig() {
  static DistributedCholeskyPolicy policy;
  return policy;
}

DistributedCholesky::~DistributedCholesky() {
  for (auto& e : factored_) {
    if (e) cudaEventDestroy(e);
  }
  for (auto& e : arrived_) {
    if (e) cudaEventDestroy(e);
  }
  if (side_) cudaStreamDestroy(side_);
}

void DistributedCholesky::Prepare() {
  if (side_) return;
  CudaCheck(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking), "cudaStreamCreate");
  for (auto& e : factored_) CudaCheck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
  for (auto& e : arrived_) CudaCheck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
}

void DistributedCholesky::Factor(cudaStream_t stream, int N, double* dH, long ld, int* d_info, int block) {
  Communicator& comm = Communicator::Get();
  Prepare();
  block = std::max(1, std::min(block, cxb_potrf_max_panel()));
  void* s = reinterpret_cast<void*>(stream);
  DeviceCheck(cxb_potrf_begin(s, d_info), "cxb_potrf_begin");
  auto col0 = [&](int J) { return J * block; };
  auto width = [&](int J) { return std::min(block, N - J * block); };
  for (const CholeskyOp& op : CholeskySchedule(N, block, comm.world(), comm.rank())) {
    const int j0 = col0(op.panel), w = width(op.panel);
    switch (op.kind) {
      case CholeskyOp::kFactor:
        DeviceCheck(cxb_potrf_panel(s, N, j0, w, dH, ld, d_info), "cxb_potrf_panel");
        break;
      case CholeskyOp::kBroadcast: {
        // The side stream may touch the panel once everything enqueued so far on the main stream is
        // done: on the owner that is the panel's factorisation, on the others nothing that reads or
        // writes this block column (they do not own it).
        cudaEvent_t ready = factored_[op.panel % 2];
        CudaCheck(cudaEventRecord(ready, stream), "cudaEventRecord");
        CudaCheck(cudaStreamWaitEvent(side_, ready, 0), "cudaStreamWaitEvent");
        // Contiguous range from entry (j0, j0) to entry (N - 1, j0 + w - 1). The entries of the
        // following columns above the diagonal that it sweeps along are never read (lower storage),
        // and the two residual rows below row N - 1 hold the same values on every rank.
        const size_t count = static_cast<size_t>(w - 1) * static_cast<size_t>(ld) + static_cast<size_t>(N - j0);
        comm.Broadcast(dH + static_cast<long>(j0) * ld + j0, count, op.target, side_);
        CudaCheck(cudaEventRecord(arrived_[op.panel % 2], side_), "cudaEventRecord");
        break;
      }
      case CholeskyOp::kWait:
        CudaCheck(cudaStreamWaitEvent(stream, arrived_[op.panel % 2], 0), "cudaStreamWaitEvent");
        break;
      case CholeskyOp::kUpdate: {
        const int k0 = col0(op.target), wk = width(op.target);
        const double* Lk = dH + static_cast<long>(j0) * ld + k0;  // L[k0:N, j0:j0+w]
        DeviceCheck(cxb_dgemm(s, 0, 1, N - k0, wk, w, -1.0, Lk, ld, 0, Lk, ld, 0, 1.0,
                              dH + static_cast<long>(k0) * ld + k0, ld, 0, 1, 1),
                    "cxb_dgemm(trailing update)");
        break;
      }
    }
  }
  // A failed panel stops its owner's later panels (d_info != 0 there) while the other ranks factor
  // whatever arrives: only the flag matters then, and every rank must see it.
  comm.AllReduceMaxInt(d_info, 1, stream);
}

}  // namespace conex
