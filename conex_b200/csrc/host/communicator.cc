#include "communicator.h"

#include <dlfcn.h>

#include <cstring>
#include <stdexcept>
#include <string>

namespace conex {

namespace {
// Minimal mirror of the NCCL C API (nccl.h, stable since NCCL 2.7): only what is called here.
struct NcclUniqueId {
  char internal[128];
};
constexpr int kNcclFloat64 = 8;
constexpr int kNcclInt32 = 2;
constexpr int kNcclSum = 0;
constexpr int kNcclMax = 2;
}  // namespace

struct Communicator::Api {
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(void**, int, NcclUniqueId, int);
  int (*CommDestroy)(void*);
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
};

Communicator& Communicator::Get() {
  static Communicator instance;
  return instance;
}

void Communicator::Load() {
  if (api_) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    lib_ = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib_) break;
  }
  if (!lib_) throw std::runtime_error("conex-b200: libnccl.so.2 not found (needed for world > 1)");
  api_ = new Api;
  auto sym = [&](const char* name) {
    void* p = dlsym(lib_, name);
    if (!p) throw std::runtime_error(std::string("conex-b200: NCCL symbol missing: ") + name);
    return p;
  };
  api_->GetUniqueId = reinterpret_cast<decltype(api_->GetUniqueId)>(sym("ncclGetUniqueId"));
  api_->CommInitRank = reinterpret_cast<decltype(api_->CommInitRank)>(sym("ncclCommInitRank"));
  api_->CommDestroy = reinterpret_cast<decltype(api_->CommDestroy)>(sym("ncclCommDestroy"));
  api_->AllReduce = reinterpret_cast<decltype(api_->AllReduce)>(sym("ncclAllReduce"));
  api_->Broadcast = reinterpret_cast<decltype(api_->Broadcast)>(sym("ncclBroadcast"));
  api_->Send = reinterpret_cast<decltype(api_->Send)>(sym("ncclSend"));
  api_->Recv = reinterpret_cast<decltype(api_->Recv)>(sym("ncclRecv"));
  api_->GroupStart = reinterpret_cast<decltype(api_->GroupStart)>(sym("ncclGroupStart"));
  api_->GroupEnd = reinterpret_cast<decltype(api_->GroupEnd)>(sym("ncclGroupEnd"));
  api_->GetErrorString = reinterpret_cast<decltype(api_->GetErrorString)>(sym("ncclGetErrorString"));
}

#define NCCL_CHECK(call, what)                                                              \
  do {                                                                                      \
    const int rc_ = (call);                                                                 \
    if (rc_ != 0) {                                                                         \
      throw std::runtime_error(std::string("conex-b200: NCCL failure in ") + what + ": " +  \
                               api_->GetErrorString(rc_));                                  \
    }                                                                                       \
  } while (0)

void Communicator::GetUniqueId(char* out128) {
  Load();
  NcclUniqueId id;
  NCCL_CHECK(api_->GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, id.internal, kUniqueIdBytes);
}

void Communicator::InitRank(int world, int rank, const char* id128) {
  if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("conex-b200: bad world/rank");
  if (comm_) Destroy();
  if (world == 1) {
    world_ = 1;
    rank_ = 0;
    return;
  }
  Load();
  NcclUniqueId id;
  std::memcpy(id.internal, id128, kUniqueIdBytes);
  NCCL_CHECK(api_->CommInitRank(&comm_, world, id, rank), "ncclCommInitRank");
  world_ = world;
  rank_ = rank;
}

void Communicator::Destroy() {
  if (comm_ && api_) api_->CommDestroy(comm_);
  comm_ = nullptr;
  world_ = 1;
  rank_ = 0;
}

void Communicator::AllReduceSum(double* buf, size_t count, cudaStream_t stream) {
  if (world_ == 1 || count == 0) return;
  NCCL_CHECK(api_->AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, comm_, stream), "ncclAllReduce");
}

void Communicator::Broadcast(double* buf, size_t count, int root, cudaStream_t stream) {
  if (world_ == 1 || count == 0) return;
  NCCL_CHECK(api_->Broadcast(buf, buf, count, kNcclFloat64, root, comm_, stream), "ncclBroadcast");
}

void Communicator::AllReduceMaxInt(int* buf, size_t count, cudaStream_t stream) {
  if (world_ == 1 || count == 0) return;
  NCCL_CHECK(api_->AllReduce(buf, buf, count, kNcclInt32, kNcclMax, comm_, stream), "ncclAllReduce(max)");
}

void Communicator::SendRecv(const double* send, size_t send_count, int to, double* recv,
                            size_t recv_count, int from, cudaStream_t stream) {
  if (world_ == 1 || (send_count == 0 && recv_count == 0)) return;
  NCCL_CHECK(api_->GroupStart(), "ncclGroupStart");
  if (send_count) NCCL_CHECK(api_->Send(send, send_count, kNcclFloat64, to, comm_, stream), "ncclSend");
  if (recv_count) NCCL_CHECK(api_->Recv(recv, recv_count, kNcclFloat64, from, comm_, stream), "ncclRecv");
  NCCL_CHECK(api_->GroupEnd(), "ncclGroupEnd");
}

// ---- partition plan ---------------------------------------------------------------------------
namespace {
// The task of `rank` towards the peer at cyclic distance d ahead of it (1 <= d <= world/2).
bool TaskAtDistance(int m, int world, int rank, int d, PairTask* t) {
  const int peer = (rank + d) % world;
  const int rb = ShardBegin(m, world, rank), re = ShardBegin(m, world, rank + 1);
  const int cb = ShardBegin(m, world, peer), ce = ShardBegin(m, world, peer + 1);
  t->peer = peer;
  t->row_begin = rb;
  t->row_count = re - rb;
  t->col_begin = cb;
  t->col_count = ce - cb;
  if (world % 2 == 0 && d == world / 2) {
    if (rank < peer) {
      // lower rank: first half of its own rows against all of the peer
      t->row_count = (re - rb) / 2;
    } else {
      // higher rank: all of its own rows against the second half of the lower rank's range
      const int half = (ce - cb) / 2;
      t->col_begin = cb + half;
      t->col_count = (ce - cb) - half;
    }
  }
  return t->row_count > 0 && t->col_count > 0;
}
}  // namespace

std::vector<PairTask> ShardPlan(int m, int world, int rank) {
  std::vector<PairTask> plan;
  for (int d = 1; d <= world / 2; d++) {
    PairTask t;
    if (TaskAtDistance(m, world, rank, d, &t)) plan.push_back(t);
  }
  return plan;
}

SendTask ShardSend(int m, int world, int rank, int distance) {
  // The receiver is the rank whose peer at `distance` is us.
  SendTask s;
  s.to = ((rank - distance) % world + world) % world;
  s.begin = 0;
  s.count = 0;
  PairTask t;
  if (distance >= 1 && distance <= world / 2 && TaskAtDistance(m, world, s.to, distance, &t)) {
    s.begin = t.col_begin;
    s.count = t.col_count;
  }
  return s;
}

}  // namespace conex
