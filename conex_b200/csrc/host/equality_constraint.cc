#include "equality_constraint.h"

namespace conex {

EqualityConstraints::EqualityConstraints(int rows, int nv, const double* A, const double* b)
    : rows_(rows), nv_(nv), dev_(std::make_shared<Device>()) {
  dev_->A.Resize(static_cast<size_t>(rows) * nv);
  dev_->b.Resize(rows);
  CudaCheck(cudaMemcpy(dev_->A.get(), A, sizeof(double) * rows * nv, cudaMemcpyHostToDevice), "upload of Aeq");
  CudaCheck(cudaMemcpy(dev_->b.get(), b, sizeof(double) * rows, cudaMemcpyHostToDevice), "upload of beq");
}

void ConstructSchurComplementSystem(EqualityConstraints* o, bool initialize, SchurComplementSystem* sys) {
  // equality_constraint.cc:13-28: G = [0 A^T; A 0] on (variables, multipliers) — only the lower
  // block is stored — and AQc = [0; b]. The assembler always asks for initialize == true.
  if (!initialize) {
    throw std::runtime_error("conex-b200: EqualityConstraints accumulates through the assembler");
  }
  cudaStream_t s = o->ctx_->cuda_stream();
  const int t = o->nv_ + o->rows_;
  CudaCheck(cudaMemsetAsync(sys->G.data, 0, sizeof(double) * sys->G.ld * t, s), "memset");
  CudaCheck(cudaMemcpy2DAsync(sys->G.data + o->nv_, sizeof(double) * sys->G.ld, o->dev_->A.get(),
                              sizeof(double) * o->rows_, sizeof(double) * o->rows_, o->nv_,
                              cudaMemcpyDeviceToDevice, s),
            "copy of Aeq");
  o->ctx_->Zero(sys->AW, t);
  o->ctx_->Zero(sys->AQc, t);
  o->ctx_->Zero(sys->scalars, 2);
  o->ctx_->CopyOnDevice(sys->AQc + o->nv_, o->dev_->b.get(), o->rows_);
}

}  // namespace conex
