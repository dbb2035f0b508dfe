// The IPM driver with device-resident workspaces — counterpart of the reference's
// conex/cone_program.{h,cc}, conex/constraint_manager.h, conex/workspace.h and (for one dense
// supernode) conex/kkt_solver.{h,cc} + conex/supernodal_assembler.{h,cc}.
//
// Host orchestration stays in C++ (loop control, mu rule, rescaling, termination, status); all
// matrices and vectors of the Newton step live in HBM and are touched only by the cxb_* kernels.
// Per iteration the host reads back O(10) scalars plus the Lanczos coefficients.
#pragma once
#include <any>
#include <list>
#include <memory>
#include <type_traits>
#include <vector>

#include "constraint.h"
#include "distributed_cholesky.h"
#include "equality_constraint.h"
#include "small_cone_constraint.h"

namespace conex {

enum : int {
  CONEX_INITIALIZATION_MODE_COLDSTART = 0,
  CONEX_INITIALIZATION_MODE_WARMSTART = 1,
};
enum : int { CONEX_LLT_FACTORIZATION = 0, CONEX_LDLT_FACTORIZATION = 1, CONEX_QR_FACTORIZATION = 2 };

// reference cone_program.h:17-38 — same fields, same defaults.
struct SolverConfiguration {
  int prepare_dual_variables = 0;
  int initialization_mode = 0;
  double inv_sqrt_mu_max = 1000;
  double minimum_mu = 1e-15;
  double maximum_mu = 1e4;
  double divergence_upper_bound = 1;
  int enable_line_search = 0;
  double dinf_upper_bound = 1;
  int final_centering_steps = 5;
  double final_centering_tolerance = .01;
  int initial_centering_steps_warmstart = 0;
  int initial_centering_steps_coldstart = 0;
  double warmstart_abort_threshold = 2;
  int max_iterations = 25;
  double infeasibility_threshold = 1e5;
  double kkt_error_tolerance = 1e10;
  int kkt_solver = 0;
  int enable_rescaling = 1;
  int iterative_refinement_iterations = 0;
};

// reference cone_program.h:40-45
struct ConexStatus {
  int solved = 0;
  int num_iterations = 0;
  int primal_infeasible = 0;
  int dual_infeasible = 0;
};

// Host-side iteration bookkeeping (reference workspace.h:73-117). The scalings persist across
// solves so that a warm start reuses them (cone_program.cc:343-357).
struct WorkspaceStats {
  std::vector<double> sqrt_inv_mu;
  double b_scaling = 1;
  double c_scaling = 1;
  int num_iter = 0;
  bool initialized = false;
};

// What REPORT() prints per iteration in the reference (cone_program.cc:456-468).
struct IterationRecord {
  double inv_sqrt_mu, mu, d_2, d_inf, by, cx, kkt_error, step_size;
  float milliseconds;  // device time of the whole Newton step (CUDA events on the program stream)
  float phase_ms[5];   // assemble, factor, mu, solve, update (same events)
  // the three terms cx is formed from, k cx b_s = 2 <c,w> + <AQc, y> - k c_s <c,Qc> (cone_program.cc:447-452): the last
  // two are ~k^2 times larger than their difference late in a solve, which is what bounds the accuracy of cx
  double cx_terms[3];
};
struct PhaseSeconds {
  double assemble = 0, factor = 0, solve = 0, update = 0, mu = 0;
};

// A cone, its clique and its slice of the assembly (reference cone_program.h:47-57 `Container`
// plus SupernodalAssembler, supernodal_assembler.h:56-131).
class Container {
 public:
  template <typename T>
  Container(const T& x, const std::vector<int>& vars)
      : obj(x), constraint(std::any_cast<T>(&obj)), variables(vars) {}
  std::any obj;
  Constraint constraint;
  std::vector<int> variables;
  SchurComplementSystem submatrix_data_;
  // The cone covers every variable in order: its G aliases the KKT matrix
  // (reference supernodal_assembler.cc:72-93 `direct_update`).
  bool direct_update = false;
  bool identity_clique = false;
  DeviceBuffer<int> d_variables;
  DeviceBuffer<double> y_clique;
};

// What the Newton driver needs from a KKT solver (reference kkt_solver.h:16-65).
class KKTSolver {
 public:
  virtual ~KKTSolver() = default;
  virtual void Bind(std::list<Container>* eqs) = 0;  // symbolic step
  virtual void Assemble() = 0;                       // reference kkt_solver.cc:164-170
  virtual bool Factor() = 0;                         // reference kkt_solver.cc:172-199
  virtual void SolveInPlace(Ref* b) const = 0;       // reference kkt_solver.cc:220-263
  // The assembled matrix as one dense N x N block (lower triangle valid) — for the export used by
  // the tests; the supernodal solver materialises it on demand.
  virtual Ref KKTMatrix() const = 0;
  virtual void SetSolverMode(int mode) = 0;
  virtual void SetIterativeRefinementIterations(int x) = 0;
  virtual void SetNumberOfMultipliers(int n) = 0;
  virtual int NumberOfSupernodes() const { return 1; }
};

// Dense stand-in for SupernodalKKTSolver: one supernode holding the whole Schur complement, factored
// by the blocked device Cholesky. Used whenever a cone couples all variables (every BASELINE
// configuration); programs whose cones act on small overlapping subsets get SupernodalKKTSolver
// (supernodal_kkt_solver.h).
class DenseKKTSolver : public KKTSolver {
 public:
  DenseKKTSolver(DeviceContext* ctx, int N);
  void Bind(std::list<Container>* eqs) override;  // symbolic step: who aliases H, who scatters
  void Assemble() override;
  bool Factor() override;                          // LLT mode; LDL^T with multipliers
  void SolveInPlace(Ref* b) const override;
  Ref KKTMatrix() const override { return Ref(H_.get(), N_, N_, ldh_); }
  void SetSolverMode(int mode) override { mode_ = mode; }
  void SetIterativeRefinementIterations(int x) override { iterative_refinement_iterations_ = x; }
  // Any multiplier block switches Factor()/SolveInPlace() to the regularised LDL^T
  // (reference kkt_solver.cc:180-193).
  void SetNumberOfMultipliers(int n) override { num_dual_ = n; }
  bool factorization_regularized() const { return factorization_regularized_; }

 private:
  void FactorLDLT();
  DistributedCholesky distributed_;  // used when the program is collective and N is large enough
  DeviceContext* ctx_;
  int N_;
  long ldh_;
  DeviceBuffer<double> H_;
  // LDL^T mode: P H P^T factored as L S L^T in Hp_, the pivot order, S, scratch
  int num_dual_ = 0;
  bool factorization_regularized_ = false;
  DeviceBuffer<double> Hp_, signs_, ldlt_work_, diag_;
  // iterative refinement (kkt_solver.cc:248-261): the assembled matrix and three N-vectors
  void SolveOnce(double* rhs) const;
  mutable DeviceBuffer<double> kkt_matrix_, refine_;
  DeviceBuffer<int> perm_;
  std::vector<int> host_perm_;
  std::list<Container>* eqs_ = nullptr;
  bool has_direct_ = false;
  int mode_ = CONEX_LLT_FACTORIZATION;
  int iterative_refinement_iterations_ = 0;
};

// What survives a solve and makes a warm start possible: the device arena (W and temporaries of every cone, the
// per-cone Schur systems, the residuals) and the iteration statistics, which the reference keeps INSIDE its arena
// (workspace.h:73-117) so that a second Program constructed on the same memory finds them
// (cone_program.h:106-109, test_warmstart.cc:47-79).
struct ProgramMemory {
  DeviceBuffer<double> data;
  WorkspaceStats stats;
};

class Program {
 public:
  explicit Program(int number_of_variables) : workspace_data_(&memory_), stats(memory_.stats) {
    SetNumberOfVariables(number_of_variables);
  }
  // Adopts the arena of another program (reference cone_program.h:106-109): after adding the same constraints in the
  // same order, a warm start continues from that program's iterate. `data` must outlive this object and must not be
  // used by two programs at the same time.
  Program(int number_of_variables, ProgramMemory* data) : workspace_data_(data), stats(data->stats) {
    SetNumberOfVariables(number_of_variables);
  }

  void SetNumberOfVariables(int m) {
    num_variables_ = m;
    linear_cost_.assign(m, 0.0);
  }
  int GetNumberOfVariables() const { return num_variables_; }
  // variables + multipliers of the equality constraints (reference constraint_manager.h:42-48)
  int SizeOfKKTSystem() const { return num_variables_ + num_dual_; }
  int NumberOfMultipliers() const { return num_dual_; }

  // reference cone_program.h:191-218 / constraint_manager.h:50-70.
  template <typename T>
  bool AddConstraint(T&& d) {
    std::vector<int> all(num_variables_);
    for (int i = 0; i < num_variables_; i++) all[i] = i;
    return AddConstraint(std::forward<T>(d), all);
  }
  template <typename T>
  bool AddConstraint(T&& d, const std::vector<int>& variables) {
    if (!VariablesAreUnique(variables)) return true;  // CONEX_FAILURE
    using Cone = std::decay_t<T>;
    std::vector<int> clique(variables);
    if constexpr (std::is_same<Cone, EqualityConstraints>::value) {
      // the multipliers become extra unknowns appended to the clique (constraint_manager.h:71-86)
      const int rows = d.SizeOfDualVariable();
      for (int i = 0; i < rows; i++) clique.push_back(num_variables_ + num_dual_ + i);
      num_dual_ += rows;
    }
    return AddToClique(static_cast<const Cone&>(d), clique);
  }
  template <typename Cone>
  bool AddToClique(const Cone& d, const std::vector<int>& variables) {
    // Only the LP cone (and equalities) implement PerformLineSearch in the reference
    // (linear_constraint.cc:82-104, equality_constraint.h:49-54); with any other cone the search
    // reports failure (constraint.h:24-28).
    if constexpr (!(std::is_same<Cone, LinearConstraint>::value || std::is_same<Cone, EqualityConstraints>::value)) {
      line_search_always_fails_ = true;
    }
    eqs.emplace_back(d, variables);
    eqs.back().constraint.bind(&ctx_);
    constraints_.push_back(&eqs.back().constraint);
    return false;  // CONEX_SUCCESS
  }

  int NumberOfConstraints() const { return static_cast<int>(eqs.size()); }
  bool LineSearchAlwaysFails() const { return line_search_always_fails_; }
  ConexStatus Status() const { return status_; }

  // Copies the (rescaled) scaling point of cone i to host memory (reference cone_program.h:120-134).
  void GetDualVariable(int i, double* host_out);
  int GetDualVariableSize(int i);

  bool AddLinearCost(const std::vector<double>& b);
  void ClearLinearCosts() { linear_cost_.assign(num_variables_, 0.0); }

  void InitializeWorkspace();  // reference cone_program.h:174-189 (device arena)

  DeviceContext ctx_;
  std::list<Container> eqs;  // std::list: addresses stay stable (constraint_manager.h:97-98)
  std::vector<Constraint*> constraints_;
  SchurComplementSystem sys;  // residual-only, program level (cone_program.cc:85-86)
  ProgramMemory memory_;            // this program's own arena (unused when another one was adopted)
  ProgramMemory* workspace_data_;   // the arena in use
  WorkspaceStats& stats;            // = workspace_data_->stats
  std::unique_ptr<KKTSolver> solver;
  // 0: choose from the clique structure (supernodal when it saves at least half of the dense
  // factorisation's flops), 1: one dense supernode, 2: supernodal (CONEXB200_SetKKTSolverKind)
  int kkt_solver_kind = 0;
  // what the current multifrontal solver was built from (Initialize keeps it across cold starts)
  std::vector<std::vector<int>> solver_cliques_;
  int solver_order_ = -1;
  int solver_kind_ = -1;
  bool solver_is_multifrontal_ = false;
  DeviceBuffer<double> vectors_;  // b, y, y2 (device copies of the host loop's m-vectors)
  bool is_initialized = false;
  ConexStatus status_;
  std::vector<double> linear_cost_;

  // diagnostics (not part of the reference ABI)
  std::vector<IterationRecord> log;
  PhaseSeconds seconds;
  bool timing_enabled = false;
  bool verbose = false;

 private:
  bool VariablesAreUnique(const std::vector<int>& x) const;
  int num_variables_ = 0;
  int num_dual_ = 0;
  bool line_search_always_fails_ = false;
};

// Pivot order of the regularised LDL^T from the diagonal alone (RLDLT.h:328-356); exposed for the
// host-logic tests.
std::vector<int> RldltPivotOrderForTest(const std::vector<double>& diag);

bool Initialize(Program& prog, const SolverConfiguration& config);
// Maximises -linear_cost (reference cone_program.cc:235-533). Writes m doubles to host memory.
bool Solve(Program& prog, const SolverConfiguration& config, double* primal_variable);
// reference cone_program.cc:547-552
bool Solve(const std::vector<double>& b, Program& prog, const SolverConfiguration& config,
           double* primal_variable);
// b = AW/2 at W = I: strictly feasible for both primal and dual (cone_program.cc:535-545).
std::vector<double> GetFeasibleObjective(Program* prog);
// Aggregates AW / AQc / {<w,c>, <c,Qc>} over all cones at the current iterate (after
// solver->Assemble()) and copies them to host memory. Used by CONEXB200_AssembleNewtonSystem.
void AssembleResidualsForExport(Program& prog, double* AW, double* AQc, double* scalars2);

}  // namespace conex
