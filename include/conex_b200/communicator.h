// One-process-per-GPU communication for the sharded Newton step: a process-wide NCCL communicator
// over NVLink 5 / NVSwitch. The reference has no counterpart (single thread).
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that (a) libconex_b200.so loads and runs
// single-GPU programs on machines without NCCL and (b) inside a torch process the already loaded
// NCCL is reused instead of a second copy. The rendezvous is left to the launcher: rank 0 obtains
// a unique id (CONEXB200_CommGetUniqueId), distributes its 128 bytes by any means
// (torch.distributed, MPI, a file) and every rank calls CONEXB200_CommInitRank.
#pragma once
#include <cuda_runtime_api.h>

#include <cstddef>
#include <vector>

namespace conex {

class Communicator {
 public:
  static Communicator& Get();  // process-wide instance; world() == 1 until InitRank succeeds

  static constexpr int kUniqueIdBytes = 128;
  void GetUniqueId(char* out128);
  void InitRank(int world, int rank, const char* id128);
  void Destroy();

  int world() const { return world_; }
  int rank() const { return rank_; }
  bool distributed() const { return world_ > 1; }

  // All collectives are enqueued on `stream` and operate on FP64 device buffers.
  void AllReduceSum(double* buf, size_t count, cudaStream_t stream);
  void Broadcast(double* buf, size_t count, int root, cudaStream_t stream);
  // Element-wise maximum of device ints over the ranks (status flags that every rank must agree on).
  void AllReduceMaxInt(int* buf, size_t count, cudaStream_t stream);
  // One grouped exchange: send `send_count` doubles to `to` (skipped when send_count == 0) and
  // receive `recv_count` doubles from `from` (skipped when recv_count == 0).
  void SendRecv(const double* send, size_t send_count, int to, double* recv, size_t recv_count,
                int from, cudaStream_t stream);

 private:
  Communicator() = default;
  void Load();
  void* lib_ = nullptr;
  void* comm_ = nullptr;
  int world_ = 1;
  int rank_ = 0;
  struct Api;
  Api* api_ = nullptr;
};

// ---- the 1-D row partition of the Schur complement and the symmetric block-pair plan ----------
// Constraint indices [0, m) are cut into `world` contiguous balanced ranges; rank r owns
// [ShardBegin(m, world, r), ShardBegin(m, world, r + 1)).
inline int ShardBegin(int m, int world, int r) {
  return static_cast<int>((static_cast<long>(m) * r) / world);
}

// One off-diagonal task of rank `rank`: contract the local scaled matrices B_i, i in
// [row_begin, row_begin + row_count) (global indices inside the rank's own range) against the
// peer's constraint matrices A_j, j in [col_begin, col_begin + col_count) (global indices inside
// the peer's range). The result is the block H[i, j]; it lands in the lower triangle directly when
// peer < rank and transposed (as H[j, i]) when peer > rank.
struct PairTask {
  int peer;
  int row_begin, row_count;
  int col_begin, col_count;
};

// Every unordered pair of distinct ranks {p, q} is assigned exactly once: rank p takes the peers at
// cyclic distance 1 .. (world-1)/2 ahead of it; for even world the pairs at distance world/2 are
// split in halves (the lower rank takes the first half of its own rows against all of the peer,
// the higher rank takes all of its own rows against the second half of the lower rank's range), so
// that every rank contracts the same number of block entries. Together with each rank's own
// diagonal block (lower triangle) this covers the lower triangle of H exactly once.
std::vector<PairTask> ShardPlan(int m, int world, int rank);

// What rank `rank` must send at cyclic distance d (to rank (rank - d) mod world): the sub-range of
// its own constraint matrices the receiver's task needs. count == 0: nothing.
struct SendTask {
  int to;
  int begin, count;  // global constraint indices
};
SendTask ShardSend(int m, int world, int rank, int distance);

}  // namespace conex
