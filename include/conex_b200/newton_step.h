// Plain data exchanged between the IPM driver and cone plugins — the device-resident counterpart
// of the reference's conex/newton_step.h:8-107. Names and meaning follow the reference; the only
// change is that `Ref` views DEVICE memory instead of being an Eigen::Map on the host.
#pragma once
#include <limits>

#include "device_runtime.h"

namespace conex {

// Column-major view of device memory (reference: using Ref = Eigen::Map<MatrixXd, Aligned>,
// newton_step.h:8-9). `ld` is the leading dimension.
struct Ref {
  double* data = nullptr;
  int rows = 0;
  int cols = 0;
  long ld = 0;
  Ref() {}
  Ref(double* d, int r, int c) : data(d), rows(r), cols(c), ld(r) {}
  Ref(double* d, int r, int c, long l) : data(d), rows(r), cols(c), ld(l) {}
  double* col(int j) const { return data + static_cast<long>(j) * ld; }
  size_t size() const { return static_cast<size_t>(rows) * cols; }
};

// reference newton_step.h:11-18
struct WeightedSlackEigenvalues {
  double limit = 0;
  double frobenius_norm_squared = 0;
  double trace = 0;
  double lambda_min = std::numeric_limits<double>::max();
  double lambda_max = -std::numeric_limits<double>::max();
  double rank = 0;
};

// reference newton_step.h:24-36
struct StepOptions {
  bool affine = true;
  double inv_sqrt_mu = 0;
  double c_weight = 0;  // step direction is e_weight * e + Q(w^{1/2})(A y - c_weight * c)
  double e_weight = 0;
  double step_size = 1;
};
struct StepInfo {
  double normsqrd = 0;
  double norminfd = 0;
};

// reference newton_step.h:51-107 (WorkspaceSchurComplement). All pointers are device pointers.
//
// Storage contract (differs from the reference, which stores a bare m x m matrix): G has leading
// dimension ldg = AugLd(m) >= m + 2 and at least m + 1 columns, i.e. two scratch rows below the m x m block.
// A plugin may use rows m and m+1 (the dense-LMI plugin lets its Gram launch drop <WCW, A_j> and
// <W, A_j> there). Only the lower triangle of the m x m block is meaningful
// (reference dense_lmi_constraint.cc:77, supernodal_assembler.cc:59-70).
struct WorkspaceSchurComplement {
  int m_ = 0;
  Ref G;                      // (m+2) x (m+1) storage, m x m lower triangle valid
  double* AW = nullptr;       // m
  double* AQc = nullptr;      // m
  double* scalars = nullptr;  // [0] = <w, c>, [1] = <c, Q(w) c>
  bool residual_only_ = false;

  static size_t size_of(int m, bool residual_only) {
    size_t s = 2 * Aligned(m) + 4;
    if (!residual_only) s += Aligned(static_cast<size_t>(AugLd(m)) * (m + 2));
    return s;
  }
  // Leading dimension of the augmented storage: m + 2 rounded up to even (16-byte aligned columns).
  static long AugLd(int m) { return (static_cast<long>(m) + 3) & ~1L; }
  static size_t Aligned(size_t n) { return (n + 3) & ~static_cast<size_t>(3); }  // memory_utils.h:4-12

  friend size_t SizeOf(const WorkspaceSchurComplement& o) { return size_of(o.m_, o.residual_only_); }
  friend void Initialize(WorkspaceSchurComplement* o, double* data) {
    o->AW = data;
    o->AQc = data + Aligned(o->m_);
    o->scalars = data + 2 * Aligned(o->m_);
    if (!o->residual_only_) {
      o->G = Ref(data + 2 * Aligned(o->m_) + 4, o->m_, o->m_, AugLd(o->m_));
    }
  }
};
using SchurComplementSystem = WorkspaceSchurComplement;

}  // namespace conex
