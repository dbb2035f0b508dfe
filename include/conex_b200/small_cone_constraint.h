// LP and second-order cone plugins with device-resident state — counterparts of the reference's
// LinearConstraint (conex/linear_constraint.{h,cc}, conex/linear_workspace.h) and SOCConstraint
// (conex/soc_constraint.{h,cc}, conex/workspace_soc.h):
//
//   LinearConstraint:  c - A y >= 0 elementwise,           A is n x m
//   SOCConstraint:     c - A y in the Lorentz cone of R^{n+1},  A is (n+1) x m
//
// Both run the batch-capable small-cone kernels (device/small_cones.cu) with a batch of one; the
// batched solver (batch_program.h) drives the same kernels for thousands of programs at a time.
// The operator can be given at construction (the reference's C++ constructors) or built entry by
// entry (CONEX_NewLinearInequality / CONEX_NewLorentzConeConstraint + CONEX_UpdateLinearOperator /
// CONEX_UpdateAffineTerm); the host copy is uploaded when it changed since the last use.
#pragma once
#include <memory>
#include <vector>

#include "../conex_b200_device.h"
#include "constraint.h"

namespace conex {

// The scaling point and temporaries of a small cone, carved from the program's device arena so
// that a warm start finds them (linear_workspace.h:10-41, workspace_soc.h:7-53).
struct WorkspaceSmallCone {
  WorkspaceSmallCone(int type, int n) : type_(type), n_(n) {}
  friend size_t SizeOf(const WorkspaceSmallCone& o) { return cxb_small_state_size(o.type_, o.n_); }
  friend void Initialize(WorkspaceSmallCone* o, double* data) {
    o->state = data;
    o->W = Ref(data, o->type_ == CXB_CONE_SOC ? o->n_ + 1 : o->n_, 1);
  }
  Ref W;                   // LP: w (n); SOC: (w0, w1) (n + 1)
  double* state = nullptr;
  int type_, n_;
};

class SmallConeConstraint {
 public:
  // data_rows = n (LP) or n + 1 (SOC); A: data_rows x m column-major or nullptr (zeros); c likewise.
  SmallConeConstraint(int type, int n, int m, const double* A, const double* c);

  WorkspaceSmallCone* workspace() { return &workspace_; }
  int number_of_variables() const { return m_; }
  void bind(DeviceContext* ctx) { ctx_ = ctx; }
  int rows() const { return rows_; }
  int type() const { return type_; }
  int order() const { return n_; }
  // Host copy of [A | c] (rows x (m + 1), column-major) — what the batched solver packs.
  const std::vector<double>& host_data() const { return *host_; }

  friend void SetIdentity(SmallConeConstraint* o);
  friend void ConstructSchurComplementSystem(SmallConeConstraint* o, bool initialize,
                                             SchurComplementSystem* sys);
  friend void PrepareStep(SmallConeConstraint* o, const StepOptions& opt, const Ref& y, StepInfo* info);
  friend bool TakeStep(SmallConeConstraint* o, const StepOptions& opt);
  friend void GetWeightedSlackEigenvalues(SmallConeConstraint* o, const Ref& y, double c_weight,
                                          WeightedSlackEigenvalues* p);
  // linear_constraint.cc:207-226, soc_constraint.cc:237-260
  friend bool UpdateLinearOperator(SmallConeConstraint* o, double val, int var, int r, int c, int dim);
  friend bool UpdateAffineTerm(SmallConeConstraint* o, double val, int r, int c, int dim);

 protected:
  cxb_small_cone Descriptor();
  struct Device;
  int type_, n_, m_, rows_;
  WorkspaceSmallCone workspace_;
  std::shared_ptr<std::vector<double>> host_;  // shared between copies of the plugin object
  std::shared_ptr<Device> dev_;
  DeviceContext* ctx_ = nullptr;
};

class LinearConstraint : public SmallConeConstraint {
 public:
  // reference linear_constraint.h:49-54: n rows, m variables, column-major A, affine term c
  LinearConstraint(int n, int m, const double* A, const double* c)
      : SmallConeConstraint(CXB_CONE_LP, n, m, A, c) {}
  friend int Rank(const LinearConstraint& o) { return o.n_; }
  // (declared here as well: the generic failing templates of constraint.h would otherwise win
  // overload resolution over the base-class friends)
  friend bool UpdateLinearOperator(LinearConstraint* o, double v, int var, int r, int c, int d) {
    return UpdateLinearOperator(static_cast<SmallConeConstraint*>(o), v, var, r, c, d);
  }
  friend bool UpdateAffineTerm(LinearConstraint* o, double v, int r, int c, int d) {
    return UpdateAffineTerm(static_cast<SmallConeConstraint*>(o), v, r, c, d);
  }
};

class SOCConstraint : public SmallConeConstraint {
 public:
  // Lorentz cone in R^{n+1} (soc_constraint.h:17-18); A is (n + 1) x m
  SOCConstraint(int n, int m, const double* A, const double* c)
      : SmallConeConstraint(CXB_CONE_SOC, n, m, A, c) {}
  friend int Rank(const SOCConstraint&) { return 2; }
  friend bool UpdateLinearOperator(SOCConstraint* o, double v, int var, int r, int c, int d) {
    return UpdateLinearOperator(static_cast<SmallConeConstraint*>(o), v, var, r, c, d);
  }
  friend bool UpdateAffineTerm(SOCConstraint* o, double v, int r, int c, int d) {
    return UpdateAffineTerm(static_cast<SmallConeConstraint*>(o), v, r, c, d);
  }
};

}  // namespace conex
