// Host-side plumbing for device-resident workspaces: RAII buffers, the per-program execution
// context (one CUDA stream + pinned staging + scalar slots) and error mapping. The product has no
// CPU fallback: every failure to reach the GPU throws, and the C ABI maps that to "not solved".
#pragma once
#include <cuda_runtime_api.h>

#include <chrono>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

namespace conex {

inline void CudaCheck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    throw std::runtime_error(std::string("conex-b200: CUDA failure in ") + what + ": " +
                             cudaGetErrorString(e));
  }
}
// Return codes of the cxb_* device layer (0 ok, >0 cudaError_t, <0 bad arguments).
inline void DeviceCheck(int rc, const char* what) {
  if (rc > 0) CudaCheck(static_cast<cudaError_t>(rc), what);
  if (rc < 0) throw std::runtime_error(std::string("conex-b200: invalid arguments to ") + what);
}

template <typename T>
class DeviceBuffer {
 public:
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t count) { Resize(count); }
  ~DeviceBuffer() { Release(); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : ptr_(o.ptr_), count_(o.count_) {
    o.ptr_ = nullptr;
    o.count_ = 0;
  }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    if (this != &o) {
      Release();
      ptr_ = o.ptr_;
      count_ = o.count_;
      o.ptr_ = nullptr;
      o.count_ = 0;
    }
    return *this;
  }
  // Contents are not preserved.
  void Resize(size_t count) {
    if (count == count_) return;
    Release();
    if (count > 0) {
      CudaCheck(cudaMalloc(reinterpret_cast<void**>(&ptr_), count * sizeof(T)), "cudaMalloc");
      count_ = count;
    }
  }
  void Reserve(size_t count) {
    if (count > count_) Resize(count);
  }
  void Release() {
    if (ptr_) cudaFree(ptr_);
    ptr_ = nullptr;
    count_ = 0;
  }
  T* get() const { return ptr_; }
  size_t size() const { return count_; }

 private:
  T* ptr_ = nullptr;
  size_t count_ = 0;
};

template <typename T>
class PinnedBuffer {
 public:
  PinnedBuffer() = default;
  explicit PinnedBuffer(size_t count) { Reserve(count); }
  ~PinnedBuffer() {
    if (ptr_) cudaFreeHost(ptr_);
  }
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  void Reserve(size_t count) {
    if (count <= count_) return;
    if (ptr_) cudaFreeHost(ptr_);
    ptr_ = nullptr;
    CudaCheck(cudaMallocHost(reinterpret_cast<void**>(&ptr_), count * sizeof(T)), "cudaMallocHost");
    count_ = count;
  }
  T* get() const { return ptr_; }
  T& operator[](size_t i) const { return ptr_[i]; }
  size_t size() const { return count_; }

 private:
  T* ptr_ = nullptr;
  size_t count_ = 0;
};

// Everything a cone plugin needs to launch work: the program's stream and small staging areas.
class DeviceContext {
 public:
  DeviceContext() {
    CudaCheck(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    // Stream-ordered scratch (cudaMallocAsync in the triangular solves) must stay cached in the pool:
    // with the default release threshold of 0 every stream synchronisation hands the memory back to
    // the driver, and re-acquiring it next to a 100 GB resident set was measured at 1.5 - 340 ms per
    // solve (CONEX_TRACE_EIGEN). Keep up to 1 GiB cached.
    int device = 0;
    cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long threshold = 1ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    pinned_.Reserve(4096);
    scalars_.Resize(64);
    flags_.Resize(16);
  }
  ~DeviceContext() {
    if (stream_) cudaStreamDestroy(stream_);
  }
  DeviceContext(const DeviceContext&) = delete;
  DeviceContext& operator=(const DeviceContext&) = delete;

  // How LMI blocks assemble their Schur complement: 0 = decide from free memory (symmetric form if it
  // fits), 1 = classic W A_i W keeping all scaled matrices, 2 = stream row panels (A + two panels of
  // scratch), 3 = symmetric form (packed L^T A_i L, fewest flops).
  int assembly_mode = 0;
  // Every rank of the process-wide Communicator holds this program with identical replicated state
  // and calls its solves in lock step (set by the sharded LMI constructors and by
  // CONEXB200_SetCollective): the factorisation may then be distributed (distributed_cholesky.h).
  bool collective = false;
  // Sharded LMI blocks exchange their scaled matrices through peer memory (CUDA IPC mappings pulled by copy engines
  // over NVLink) when every rank can map every other's; false: always ncclSend / ncclRecv (A/B, CONEXB200_SetPeerMemoryExchange).
  bool peer_memory_exchange = true;

  void* stream() const { return reinterpret_cast<void*>(stream_); }
  cudaStream_t cuda_stream() const { return stream_; }
  void Synchronize() const { CudaCheck(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"); }

  // 64 device doubles for reduction results, 16 device ints for info flags.
  double* scalars() const { return scalars_.get(); }
  int* flags() const { return flags_.get(); }
  PinnedBuffer<double>& pinned() { return pinned_; }

  void Upload(double* dst, const double* src, size_t count) const {
    CudaCheck(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, stream_),
              "H2D copy");
  }
  // Blocking download (synchronises the stream).
  void Download(double* dst, const double* src, size_t count) const {
    CudaCheck(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream_),
              "D2H copy");
    Synchronize();
  }
  void DownloadInts(int* dst, const int* src, size_t count) const {
    CudaCheck(cudaMemcpyAsync(dst, src, count * sizeof(int), cudaMemcpyDeviceToHost, stream_),
              "D2H copy");
    Synchronize();
  }
  void CopyOnDevice(double* dst, const double* src, size_t count) const {
    CudaCheck(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, stream_),
              "D2D copy");
  }
  void Zero(double* dst, size_t count) const {
    CudaCheck(cudaMemsetAsync(dst, 0, count * sizeof(double), stream_), "memset");
  }

 private:
  cudaStream_t stream_ = nullptr;
  PinnedBuffer<double> pinned_;
  DeviceBuffer<double> scalars_;
  DeviceBuffer<int> flags_;
};

// Wall-clock phase timer; the stream is synchronised on both sides so device work is included.
class PhaseTimer {
 public:
  PhaseTimer(const DeviceContext* ctx, double* accumulator, bool enabled)
      : ctx_(ctx), acc_(accumulator), enabled_(enabled) {
    if (enabled_) {
      ctx_->Synchronize();
      t0_ = std::chrono::high_resolution_clock::now();
    }
  }
  ~PhaseTimer() {
    if (enabled_) {
      try {
        ctx_->Synchronize();
      } catch (...) {
      }
      *acc_ += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0_).count();
    }
  }

 private:
  const DeviceContext* ctx_;
  double* acc_;
  bool enabled_;
  std::chrono::high_resolution_clock::time_point t0_;
};

}  // namespace conex
