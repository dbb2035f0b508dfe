// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the cone plugins on conex's Newton-step hot path:
//   * the dense-LMI / PSD cone   (conex/psd_constraint.{h,cc}, conex/dense_lmi_constraint.{h,cc})
//   * the LP cone                (conex/linear_constraint.{h,cc}) — needed for the reference's
//     "Mixed" known-answer SDP (conex/test/test_sdp.cc:13-59)
// and of the dense math kernels below them (Padé map, two-sided Lanczos).
// Parity status: PINNED against the reference's own known-answer tests at their
// tolerances (see tests/test_oracle_golden.py and oracle/README.md); bit-level parity with
// Eigen is not pinned because the reference cannot be built here (Eigen 3.3.9 absent).
#pragma once
#include <vector>

#include "linalg.h"

namespace oracle {

// conex/newton_step.h:11-48 (plain data carried between driver and cones).
struct SlackEigenvalues {
  double frobenius_norm_squared = 0;
  double trace = 0;
  double lambda_min = 1.7976931348623157e308;   // newton_step.h:15-16: neutral for the min/max
  double lambda_max = -1.7976931348623157e308;  // aggregation (the equality cone leaves them)
  double rank = 0;
};
struct StepOptions {
  bool affine = true;
  double inv_sqrt_mu = 0;
  double c_weight = 0;
  double e_weight = 0;
  double step_size = 1;
};
struct StepInfo {
  double normsqrd = 0;
  double norminfd = 0;
};

// conex/newton_step.h:51-107. G is m x m (lower triangle meaningful), AW / AQc are m-vectors.
struct SchurSystem {
  int m = 0;
  bool residual_only = false;
  double* AW = nullptr;
  double* AQc = nullptr;
  View G;
  double inner_product_of_w_and_c = 0;
  double inner_product_of_c_and_Qc = 0;
  int SizeOf() const {
    int s = 2 * AlignedSize(m);
    if (!residual_only) s += AlignedSize(m * m);
    return s;
  }
  void Bind(double* data) {
    AW = data;
    AQc = data + AlignedSize(m);
    if (!residual_only) G = View(data + 2 * AlignedSize(m), m, m);
  }
  void SetZero();
};

// The role of conex::Constraint (conex/constraint.h:51-197): what the driver needs of a cone.
class Cone {
 public:
  virtual ~Cone() {}
  virtual int WorkspaceSize() const = 0;
  virtual void BindWorkspace(double* data) = 0;
  virtual void SetIdentity() = 0;
  virtual int Rank() const = 0;
  virtual int NumberOfVariables() const = 0;
  virtual void ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) = 0;
  virtual void PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) = 0;
  virtual bool TakeStep(const StepOptions& opt) = 0;
  virtual void GetWeightedSlackEigenvalues(const double* y, double c_weight,
                                           SlackEigenvalues* p) = 0;
  virtual const double* DualVariable() const = 0;
  virtual int DualVariableSize() const = 0;
  // The scaling point as one vector (LP: w; SOC: (w0, w1); PSD: vec(W)) — tests only.
  virtual int StateSize() const { return DualVariableSize(); }
  virtual void GetState(double* out) const {
    const double* w = DualVariable();
    for (int i = 0; i < DualVariableSize(); i++) out[i] = w[i];
  }
  virtual void SetState(const double* in) {
    double* w = const_cast<double*>(DualVariable());
    for (int i = 0; i < DualVariableSize(); i++) w[i] = in[i];
  }
};

// How the Gram rows of the Schur complement are formed.
enum class GramVariant {
  kAsWritten = 0,  // m GEMVs over growing slabs: dense_lmi_constraint.cc:75-78
  kBlas3 = 1,      // same numbers through one GEMM per 64-row panel (reordered summation)
  kSymmetric = 2,  // H_ij = <L^T A_i L, L^T A_j L> with W = L L^T: the same matrix through the Cholesky
                   // factor of W (the form the product's default assembly uses); NOT in the reference
};

// DenseLMIConstraint (conex/dense_lmi_constraint.h:24-41) on a PsdConstraint
// (conex/psd_constraint.h:39-65).
class DenseLmiCone : public Cone {
 public:
  // A: m contiguous column-major n x n matrices (interfaces/conex.cc:143-151); C: n x n.
  DenseLmiCone(int n, int m, const double* A, const double* C);
  int WorkspaceSize() const override { return 3 * AlignedSize(n_ * n_); }
  void BindWorkspace(double* data) override;
  void SetIdentity() override;
  int Rank() const override { return n_; }
  int NumberOfVariables() const override { return m_; }
  void ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) override;
  void PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) override;
  bool TakeStep(const StepOptions& opt) override;
  void GetWeightedSlackEigenvalues(const double* y, double c_weight,
                                   SlackEigenvalues* p) override;
  const double* DualVariable() const override { return W_.p; }
  int DualVariableSize() const override { return n_ * n_; }

  void ComputeNegativeSlack(double k, const double* y, View s) const;
  GramVariant gram_variant = GramVariant::kAsWritten;
  View W_, temp_1_, temp_2_;

 protected:
  void GeodesicUpdate(double scale, const StepOptions& opt, View WS);
  void AffineUpdate(double w_e, View WS);
  int n_, m_;
  std::vector<double> Avect_;  // n*n x m, column i = vec(A_i)
  std::vector<double> C_;      // n x n
};

// HermitianPsdConstraint<Real> (conex/hermitian_psd.{h,cc} on MatrixAlgebra<1>,
// conex/jordan_matrix_algebra.cc, conex/exponential_map.cc): the LMI built entry by entry through
// CONEX_NewLinearMatrixInequality / CONEX_UpdateLinearOperator. Same Schur complement as the dense
// LMI cone, but its own eigen-bound (random start vector from libc rand() like Eigen's Random(),
// n/2 + 1 Lanczos steps, relative breakdown test) and its own exponential (degree-2 Taylor with
// two squarings instead of the Padé map).
class HermitianLmiCone final : public DenseLmiCone {
 public:
  HermitianLmiCone(int n, int m);
  void PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) override;
  bool TakeStep(const StepOptions& opt) override;
  void GetWeightedSlackEigenvalues(const double* y, double c_weight, SlackEigenvalues* p) override;
  // hermitian_psd.cc:248-322 (real algebra: symmetric pair set)
  void SetOperatorEntry(int var, int r, int c, double v) {
    Avect_[(size_t)var * n_ * n_ + (size_t)c * n_ + r] = v;
    Avect_[(size_t)var * n_ * n_ + (size_t)r * n_ + c] = v;
  }
  void SetAffineEntry(int r, int c, double v) {
    C_[(size_t)c * n_ + r] = v;
    C_[(size_t)r * n_ + c] = v;
  }
  int order() const { return n_; }
};
// MatrixAlgebra<1>::ApproximateEigenvalues (jordan_matrix_algebra.cc:387-452): returns the Ritz
// values (ascending).
std::vector<double> HermitianLanczos(int n, const double* WS, const double* W, const double* r,
                                     int num_iter);
// DoExponentialMap<1> (exponential_map.cc:15-42): (I + X/4 + X^2/32)^4.
void ExponentialMapTaylor(int n, const double* X, double* result);

// LinearConstraint (conex/linear_constraint.h:14-89): c - A y >= 0 with A n x m.
class LinearCone final : public Cone {
 public:
  LinearCone(int n, int m, const double* A, const double* c);
  int WorkspaceSize() const override { return 3 * AlignedSize(n_) + AlignedSize(n_ * m_); }
  void BindWorkspace(double* data) override;
  void SetIdentity() override;
  int Rank() const override { return n_; }
  int NumberOfVariables() const override { return m_; }
  void ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) override;
  void PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) override;
  bool TakeStep(const StepOptions& opt) override;
  void GetWeightedSlackEigenvalues(const double* y, double c_weight,
                                   SlackEigenvalues* p) override;
  const double* DualVariable() const override { return W_; }
  int DualVariableSize() const override { return n_; }
  // incremental construction (linear_constraint.cc:207-226)
  int rows() const { return n_; }
  void SetOperatorEntry(int row, int var, double v) { A_[(size_t)var * n_ + row] = v; }
  void SetAffineEntry(int row, double v) { c_[row] = v; }

 private:
  void ComputeNegativeSlack(double k, const double* y, double* minus_s) const;
  int n_, m_;
  std::vector<double> A_, c_;
  double *W_ = nullptr, *temp_1_ = nullptr, *temp_2_ = nullptr, *WA_ = nullptr;
};

// SOCConstraint (conex/soc_constraint.{h,cc}, conex/workspace_soc.h): c - A y in the Lorentz cone
// of order n+1; A is (n+1) x m, scaling point w = (W0, W1).
class SocCone final : public Cone {
 public:
  SocCone(int n, int m, const double* A, const double* c);
  int WorkspaceSize() const override { return 5 * AlignedSize(n_); }  // workspace_soc.h:9-11
  void BindWorkspace(double* data) override;
  void SetIdentity() override;
  int Rank() const override { return 2; }
  int NumberOfVariables() const override { return m_; }
  void ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) override;
  void PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) override;
  bool TakeStep(const StepOptions& opt) override;
  void GetWeightedSlackEigenvalues(const double* y, double c_weight,
                                   SlackEigenvalues* p) override;
  const double* DualVariable() const override { return dual_.data(); }
  int DualVariableSize() const override { return n_ + 1; }
  void GetState(double* out) const override {
    out[0] = *W0_;
    for (int i = 0; i < n_; i++) out[1 + i] = W1_[i];
  }
  void SetState(const double* in) override {
    *W0_ = in[0];
    for (int i = 0; i < n_; i++) W1_[i] = in[1 + i];
  }
  // incremental construction (soc_constraint.cc:237-260)
  int order() const { return n_; }
  void SetOperatorEntry(int row, int var, double v) { A_[(size_t)var * (n_ + 1) + row] = v; }
  void SetAffineEntry(int row, double v) { c_[row] = v; }
  double* W0_ = nullptr;
  double* W1_ = nullptr;

 private:
  void ComputeNegativeSlack(double k, const double* y, double* minus_s) const;
  int n_, m_;
  std::vector<double> A_, c_;
  double* temp1_ = nullptr;
  double d0_ = 0;
  std::vector<double> dual_;  // the reference's dummy `W` member (workspace_soc.h:44)
};

// EqualityConstraints (conex/equality_constraint.{h,cc}): A x = b on `nv` variables with
// A.rows() multipliers appended to the clique (constraint_manager.h:71-86). Rank 0, no state.
class EqualityCone final : public Cone {
 public:
  EqualityCone(int rows, int nv, const double* A, const double* b)
      : rows_(rows), nv_(nv), A_(A, A + (size_t)rows * nv), b_(b, b + rows), lambda_(rows, 0.0) {}
  int WorkspaceSize() const override { return 0; }
  void BindWorkspace(double*) override {}
  void SetIdentity() override {}
  int Rank() const override { return 0; }
  int NumberOfVariables() const override { return nv_ + rows_; }  // size of its clique
  int NumberOfMultipliers() const { return rows_; }
  void ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) override;
  void PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) override;
  bool TakeStep(const StepOptions&) override { return true; }
  void GetWeightedSlackEigenvalues(const double*, double, SlackEigenvalues*) override {}
  const double* DualVariable() const override { return lambda_.data(); }
  int DualVariableSize() const override { return 0; }  // its workspace W is an empty map

 private:
  int rows_, nv_;
  std::vector<double> A_, b_, lambda_;
};

// --- dense math kernels -------------------------------------------------------------------
// [3/3] Padé approximation of exp(X), no scaling/squaring (conex/exponential_map_pade.cc:10-32).
void ExponentialMapPade(int n, const double* X, double* result);
// Two-sided Lanczos on (WS, WS^T) in the W inner product (conex/approximate_eigenvalues.cc:173-239).
// Returns the Ritz values (ascending).
std::vector<double> AsymmetricLanczos(int n, const double* WS, const double* W, const double* r,
                                      int num_iter);
// conex/approximate_eigenvalues.cc:241-256 with compressed == true.
std::vector<double> ApproximateEigenvalues(int n, const double* WS, const double* W,
                                           const double* r, int num_iter);
// Symmetric Lanczos, conex/approximate_eigenvalues.cc:147-171.
std::vector<double> SymmetricLanczos(int n, const double* A, const double* r0, int num_iter);

// conex/divergence.cc:96-111; returns -1 when neither branch yields a finite bound.
double DivergenceUpperBoundInverse(double divergence_upper_bound, const SlackEigenvalues& p);
// conex/divergence.cc:113-121.
double DivergenceUpperBound(double k, const SlackEigenvalues& p);

}  // namespace oracle
