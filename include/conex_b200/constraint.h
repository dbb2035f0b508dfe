// The C++ cone plugin concept — device-resident counterpart of the reference's type-erased
// `Constraint` (conex/constraint.h:51-197). A cone type T plugs into Program::AddConstraint iff
// these are findable by argument-dependent lookup, exactly as in the reference:
//
//   void ConstructSchurComplementSystem(T*, bool initialize, SchurComplementSystem* sys);
//   void SetIdentity(T*);
//   void PrepareStep(T*, const StepOptions&, const Ref& y, StepInfo*);
//   bool TakeStep(T*, const StepOptions&);
//   void GetWeightedSlackEigenvalues(T*, const Ref& y, double c_weight, WeightedSlackEigenvalues*);
//   int  Rank(const T&);
//   members: WS* workspace();  int number_of_variables();
//   WS provides friends SizeOf(const WS&), Initialize(WS*, double* device_arena) and a member
//   `Ref W` (the scaling point, what CONEX_GetDualVariable copies out).
//
// Differences forced by the device boundary: `Ref` views device memory, `y` is a device vector
// (already gathered for the cone's clique), and T additionally provides
//   void bind(DeviceContext*)   — the stream/staging the cone must launch on.
// PrepareStep / GetWeightedSlackEigenvalues return host scalars, so they synchronise the stream.
#pragma once
#include <iostream>
#include <memory>

#include "newton_step.h"

namespace conex {

// reference conex/error_checking_macros.h:15-19
#define CONEX_DEMAND(x, msg)                                                   \
  if (!(x)) {                                                                  \
    std::cerr << __FILE__ << " line " << __LINE__ << ": " << msg << std::endl; \
    return 1;                                                                  \
  }

// Optional operations default to failure (reference constraint.h:13-28).
template <typename T>
bool UpdateLinearOperator(T*, double, int, int, int, int) {
  CONEX_DEMAND(false, "Constraint does not support updates of linear operator.");
}
template <typename T>
bool UpdateAffineTerm(T*, double, int, int, int) {
  CONEX_DEMAND(false, "Constraint does not support updates of affine term.");
}

// Type-erased workspace (reference conex/workspace.h:37-70).
class Workspace {
 public:
  template <typename T>
  explicit Workspace(T* t) : model_(std::make_unique<Model<T>>(t)) {}
  friend void Initialize(Workspace* o, double* device_data) { o->model_->do_initialize(device_data); }
  friend size_t SizeOf(const Workspace& o) { return o.model_->do_sizeof(); }

 private:
  struct Concept {
    virtual ~Concept() = default;
    virtual void do_initialize(double*) = 0;
    virtual size_t do_sizeof() const = 0;
  };
  template <typename T>
  struct Model final : Concept {
    explicit Model(T* t) : data(t) {}
    void do_initialize(double* p) override { Initialize(data, p); }
    size_t do_sizeof() const override { return SizeOf(*data); }
    T* data;
  };
  std::unique_ptr<Concept> model_;
};

class Constraint {
 public:
  template <typename Implementation>
  explicit Constraint(Implementation* t) : model_(std::make_unique<Model<Implementation>>(t)) {}

  friend void ConstructSchurComplementSystem(Constraint* o, bool initialize,
                                             SchurComplementSystem* sys) {
    o->model_->do_schur_complement(initialize, sys);
  }
  friend void SetIdentity(Constraint* o) { o->model_->do_set_identity(); }
  friend void PrepareStep(Constraint* o, const StepOptions& opt, const Ref& y, StepInfo* info) {
    o->model_->do_prepare_step(opt, y, info);
  }
  friend bool TakeStep(Constraint* o, const StepOptions& opt) { return o->model_->do_take_step(opt); }
  friend void GetWeightedSlackEigenvalues(Constraint* o, const Ref& y, double c_weight,
                                          WeightedSlackEigenvalues* p) {
    o->model_->do_weighted_slack_eigenvalues(y, c_weight, p);
  }
  friend int Rank(const Constraint& o) { return o.model_->do_rank(); }
  friend bool UpdateLinearOperator(Constraint* o, double val, int var, int row, int col, int dim) {
    return o->model_->do_update_linear_operator(val, var, row, col, dim);
  }
  friend bool UpdateAffineTerm(Constraint* o, double val, int row, int col, int dim) {
    return o->model_->do_update_affine_term(val, row, col, dim);
  }

  Workspace workspace() { return model_->do_get_workspace(); }
  int number_of_variables() { return model_->do_number_of_variables(); }
  void bind(DeviceContext* ctx) { model_->do_bind(ctx); }
  // Device view of the scaling point W (reference: get_dual_variable memcpy's workspace()->W).
  Ref dual_variable() { return model_->do_dual_variable(); }
  int dual_variable_size() {
    const Ref w = dual_variable();
    return w.rows * w.cols;
  }

 private:
  struct Concept {
    virtual ~Concept() = default;
    virtual void do_schur_complement(bool, SchurComplementSystem*) = 0;
    virtual void do_set_identity() = 0;
    virtual void do_prepare_step(const StepOptions&, const Ref&, StepInfo*) = 0;
    virtual bool do_take_step(const StepOptions&) = 0;
    virtual void do_weighted_slack_eigenvalues(const Ref&, double, WeightedSlackEigenvalues*) = 0;
    virtual int do_rank() = 0;
    virtual Workspace do_get_workspace() = 0;
    virtual int do_number_of_variables() = 0;
    virtual void do_bind(DeviceContext*) = 0;
    virtual Ref do_dual_variable() = 0;
    virtual bool do_update_linear_operator(double, int, int, int, int) = 0;
    virtual bool do_update_affine_term(double, int, int, int) = 0;
  };
  template <typename Implementation>
  struct Model final : Concept {
    explicit Model(Implementation* t) : data(t) {}
    void do_schur_complement(bool initialize, SchurComplementSystem* sys) override {
      ConstructSchurComplementSystem(data, initialize, sys);
    }
    void do_set_identity() override { SetIdentity(data); }
    void do_prepare_step(const StepOptions& opt, const Ref& y, StepInfo* info) override {
      PrepareStep(data, opt, y, info);
    }
    bool do_take_step(const StepOptions& opt) override { return TakeStep(data, opt); }
    void do_weighted_slack_eigenvalues(const Ref& y, double c_weight,
                                       WeightedSlackEigenvalues* p) override {
      GetWeightedSlackEigenvalues(data, y, c_weight, p);
    }
    int do_rank() override { return Rank(*data); }
    Workspace do_get_workspace() override { return Workspace(data->workspace()); }
    int do_number_of_variables() override { return data->number_of_variables(); }
    void do_bind(DeviceContext* ctx) override { data->bind(ctx); }
    Ref do_dual_variable() override { return data->workspace()->W; }
    bool do_update_linear_operator(double v, int var, int r, int c, int d) override {
      return UpdateLinearOperator(data, v, var, r, c, d);
    }
    bool do_update_affine_term(double v, int r, int c, int d) override {
      return UpdateAffineTerm(data, v, r, c, d);
    }
    Implementation* data;
  };
  std::unique_ptr<Concept> model_;
};

}  // namespace conex
