// The dense-LMI / PSD cone plugin with device-resident state — counterpart of the reference's
// PsdConstraint (conex/psd_constraint.{h,cc}) and DenseLMIConstraint
// (conex/dense_lmi_constraint.{h,cc}).
//
//   c - sum_i y_i A_i  in  PSD(n),   A_i, c symmetric n x n.
//
// Device layout (HBM): `Aall` holds the m constraint matrices followed by the affine term C as
// m+1 contiguous column-major n x n blocks, i.e. the n^2 x (m+1) matrix [vec(A_0) .. vec(A_{m-1})
// vec(C)] (reference keeps three host copies: constraint_matrices_, constraint_matrices_vect_,
// constraint_affine_; dense_lmi_constraint.h:13-15). W, temp_1, temp_2 are carved from the
// program's device arena exactly like WorkspaceDensePSD (psd_constraint.h:9-37), so a warm start
// finds the iterate where the previous solve left it.
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "constraint.h"

namespace conex {

struct WorkspaceDensePSD {
  explicit WorkspaceDensePSD(int n) : n_(n) {}
  static size_t size_of(int n) {
    return 3 * WorkspaceSchurComplement::Aligned(static_cast<size_t>(n) * n);
  }
  friend size_t SizeOf(const WorkspaceDensePSD& o) { return size_of(o.n_); }
  friend void Initialize(WorkspaceDensePSD* o, double* data) {
    const size_t stride = WorkspaceSchurComplement::Aligned(static_cast<size_t>(o->n_) * o->n_);
    o->W = Ref(data, o->n_, o->n_);
    o->temp_1 = Ref(data + stride, o->n_, o->n_);
    o->temp_2 = Ref(data + 2 * stride, o->n_, o->n_);
  }
  Ref W, temp_1, temp_2;
  int n_;
};

class DenseLMIConstraint {
 public:
  // Host data: `A` = m contiguous column-major n x n matrices, `C` = n x n (copied to the device;
  // the caller may free them — reference interfaces/conex.cc:137-160).
  DenseLMIConstraint(int n, int m, const double* A, const double* C);
  // Device data (both pointers on the current device); copied device-to-device.
  struct DevicePointers {
    const double* A;
    const double* C;
  };
  DenseLMIConstraint(int n, int m, DevicePointers dev);
  // One rank's shard of a block whose m constraint matrices are partitioned over the ranks of the
  // process-wide Communicator (communicator.h: rank r owns [ShardBegin(m, world, r),
  // ShardBegin(m, world, r + 1))). `dev.A` holds only the local matrices, `dev.C` the full affine
  // term (replicated). W, y and the whole Newton system stay replicated and bit-identical on every
  // rank; only K1/K2 (assembly) and K6 (slack GEMV) are sharded — see DESIGN.md "Multi-GPU".
  struct Sharded {};
  DenseLMIConstraint(int n, int m_global, Sharded, DevicePointers dev);
  // Same, but the block only allocates its storage (this rank's shard of A followed by C) and the
  // caller fills it in place through mutable_device_matrices() before the first solve — no second
  // copy of a 160 GB operator ever exists.
  struct Uninitialized {};
  DenseLMIConstraint(int n, int m_global, Sharded, Uninitialized);
  double* mutable_device_matrices();  // local A_i blocks, then C
  int local_matrices() const { return m_local_; }

  // Operator given by the non-zero entries of its matrices (lower triangle, r >= c): the block then
  // runs the entry-sparse assembly and slack kernels (device/sparse_lmi.cu) and never allocates the
  // dense n^2 x m operator. C: dense n x n (host).
  struct Entry {
    int var, r, c;
    double val;
  };
  struct EntrySparse {};
  DenseLMIConstraint(int n, int m, EntrySparse);
  void LoadEntries(const std::vector<Entry>& lower_entries, const double* C);
  // Dense storage for a block created with the EntrySparse constructor (allocated on first use).
  void LoadDense(const std::vector<Entry>& lower_entries, const double* C);
  bool entry_sparse() const;
  // Sharded blocks: device milliseconds of the last assembly on this rank — {local K1 + diagonal block, stalls of
  // the compute stream waiting for a peer's chunk, off-diagonal contractions, all-reduce of H}. False if not sharded.
  bool shard_phase_milliseconds(double* out4) const;
  // How this block assembles its Schur complement: 0 undecided (before the first assembly), 1 classic
  // (all W A_i W kept), 2 row panels, 3 symmetric (packed L^T A_i L), 4 entry-sparse gathers.
  int assembly_form() const;
  // Switches the eigen-bound and the exponential to the rules of the reference's incremental LMI
  // (HermitianPsdConstraint<Real>, conex/hermitian_psd.cc): Lanczos started from an Eigen-style
  // Random(n, 1) vector drawn with libc rand(), n/2 + 1 steps, relative breakdown test
  // (jordan_matrix_algebra.cc:387-452); exp = (I + X/4 + X^2/32)^4 (exponential_map.cc:15-42).
  void set_hermitian_semantics(bool on) { hermitian_ = on; }

  WorkspaceDensePSD* workspace() { return &workspace_; }
  int number_of_variables() const { return m_; }
  int order() const { return n_; }
  void bind(DeviceContext* ctx) { ctx_ = ctx; }
  const double* device_matrices() const;  // Aall

  friend int Rank(const DenseLMIConstraint& o) { return o.n_; }
  friend void SetIdentity(DenseLMIConstraint* o);
  friend void ConstructSchurComplementSystem(DenseLMIConstraint* o, bool initialize,
                                             SchurComplementSystem* sys);
  friend void PrepareStep(DenseLMIConstraint* o, const StepOptions& opt, const Ref& y, StepInfo*);
  friend bool TakeStep(DenseLMIConstraint* o, const StepOptions& opt);
  friend void GetWeightedSlackEigenvalues(DenseLMIConstraint* o, const Ref& y, double c_weight,
                                          WeightedSlackEigenvalues* p);

 private:
  struct Storage;
  void EnsureScratch();
  void ExchangePeerHandles();  // sharded blocks: map the peers' scaled matrices (CUDA IPC)
  // minus_s = sum_i y_i A_i - k C  (reference ComputeNegativeSlack, dense_lmi_constraint.cc:22-27)
  void ComputeNegativeSlack(double k, const Ref& y, Ref* minus_s);
  // Two-sided Lanczos extreme Ritz values of WS w.r.t. W started from column `index` of R, plus
  // tr(WS) and tr(WS WS). Synchronises the stream.
  struct SpectrumEstimate {
    double ritz_min, ritz_max, trace, trace_of_square;
  };
  SpectrumEstimate EstimateSpectrum(const Ref& WS, const Ref& start_matrix);

  void AssembleSharded(SchurComplementSystem* sys);

  int n_;
  int m_;             // number of variables of the block (global)
  int m_local_;       // constraint matrices held by this rank (== m_ when not sharded)
  int row_begin_ = 0;  // global index of the first local matrix
  bool sharded_ = false;
  bool hermitian_ = false;
  WorkspaceDensePSD workspace_;
  std::shared_ptr<Storage> data_;
  DeviceContext* ctx_ = nullptr;
};

// The LMI built entry by entry through CONEX_NewLinearMatrixInequality / CONEX_UpdateLinearOperator /
// CONEX_UpdateAffineTerm — the reference's HermitianPsdConstraint<Real> (conex/hermitian_psd.{h,cc}).
// The host keeps the entries it is given (a map per variable, lower triangle) and the dense affine
// term; they are uploaded when they changed since the last use. Operators with few entries per matrix
// (MaxCut, Lovasz theta, ...) run the entry-sparse kernels — O((sum nnz)^2) gathers instead of
// 4 m n^3 + m^2 n^2 flops, O(nnz) memory instead of m n^2 doubles; dense ones the dense-LMI path. Both
// use the Hermitian eigen-bound and exponential rules.
class HermitianPsdConstraint : public DenseLMIConstraint {
 public:
  // order n, m variables (matrices of variables never updated stay zero)
  HermitianPsdConstraint(int n, int m);

  friend int Rank(const HermitianPsdConstraint& o) { return o.order(); }
  friend void SetIdentity(HermitianPsdConstraint* o) { SetIdentity(o->Synced()); }
  friend void ConstructSchurComplementSystem(HermitianPsdConstraint* o, bool initialize,
                                             SchurComplementSystem* sys) {
    ConstructSchurComplementSystem(o->Synced(), initialize, sys);
  }
  friend void PrepareStep(HermitianPsdConstraint* o, const StepOptions& opt, const Ref& y, StepInfo* info) {
    PrepareStep(o->Synced(), opt, y, info);
  }
  friend bool TakeStep(HermitianPsdConstraint* o, const StepOptions& opt) { return TakeStep(o->Synced(), opt); }
  friend void GetWeightedSlackEigenvalues(HermitianPsdConstraint* o, const Ref& y, double c_weight,
                                          WeightedSlackEigenvalues* p) {
    GetWeightedSlackEigenvalues(o->Synced(), y, c_weight, p);
  }
  // hermitian_psd.cc:248-322 for the real algebra
  friend bool UpdateLinearOperator(HermitianPsdConstraint* o, double val, int var, int r, int c, int dim);
  friend bool UpdateAffineTerm(HermitianPsdConstraint* o, double val, int r, int c, int dim);

 private:
  struct Host {
    std::vector<std::map<long, double>> entries;  // per variable: c * n + r (r >= c) -> value
    std::vector<double> C;
    bool dirty = true;
  };
  DenseLMIConstraint* Synced();
  std::shared_ptr<Host> host_;
};

}  // namespace conex
