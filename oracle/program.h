// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of conex's IPM driver: conex/cone_program.{h,cc}, conex/constraint_manager.h,
// conex/workspace.h, plus a dense stand-in for the supernodal KKT solver
// (conex/kkt_solver.cc, conex/block_triangular_operations.cc). The reference factors a
// chordal-sparse KKT matrix; this oracle assembles the same matrix densely (identical
// entries, identical solution up to rounding) because every BASELINE config is a single
// dense clique (SURVEY.md §8 a-K2'').
#pragma once
#include <memory>
#include <vector>

#include "cones.h"

namespace oracle {

// conex/cone_program.h:17-38 (same defaults).
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

// What REPORT() prints per iteration in the reference (conex/cone_program.cc:456-468),
// kept so tests can compare trajectories.
struct IterationRecord {
  double inv_sqrt_mu, mu, d_2, d_inf, by, cx, kkt_error, step_size;
};

struct PhaseSeconds {
  double assemble = 0, factor = 0, solve = 0, update = 0, mu = 0;
};

class Program {
 public:
  explicit Program(int m) : m_(m), linear_cost_(m, 0.0) {}
  void SetNumberOfVariables(int m) {
    m_ = m;
    linear_cost_.assign(m, 0.0);
  }
  int NumberOfVariables() const { return m_; }
  int NumberOfConstraints() const { return (int)cones_.size(); }
  // Returns false when `vars` has duplicates / out-of-range entries
  // (conex/constraint_manager.h:10-24,62-70).
  bool AddCone(std::unique_ptr<Cone> cone, const std::vector<int>& vars);
  bool AddCone(std::unique_ptr<Cone> cone);
  // conex/constraint_manager.h:71-94: the multipliers of A x = b become extra unknowns of the KKT
  // system, appended to the cone's clique. A is rows x vars.size() column-major.
  bool AddEquality(int rows, const double* A, const double* b, const std::vector<int>& vars);
  bool AddEquality(int rows, const double* A, const double* b);
  int SizeOfKKTSystem() const { return m_ + num_dual_; }

  // conex/cone_program.cc:235-533; b is maximised (Solve(b, prog, ...) :547-552).
  bool Maximize(const double* b, const SolverConfiguration& config, double* y);
  // conex/cone_program.cc:535-545: b = AW / 2 at W = I.
  std::vector<double> FeasibleObjective();
  // conex/cone_program.h:120-134
  void GetDualVariable(int i, double* x) const;
  int GetDualVariableSize(int i) const;

  struct Status {
    int solved = 0, num_iterations = 0, primal_infeasible = 0, dual_infeasible = 0;
  } status;
  std::vector<IterationRecord> log;
  std::vector<double> sqrt_inv_mu;  // stats->sqrt_inv_mu
  PhaseSeconds seconds;
  GramVariant gram_variant = GramVariant::kAsWritten;
  std::vector<double> H;  // dense KKT matrix, lower triangle, column-major m x m
  // residual-only program-level system (conex/cone_program.cc:85-86)
  SchurSystem sys;

  bool Initialize(const SolverConfiguration& config);
  void Assemble();
  Cone* cone(int i) { return cones_[i].get(); }

 private:
  bool Factor();
  void SolveInPlace(double* rhs) const;
  double ComputeMuFromDivergence(const SolverConfiguration& config, int rank,
                                 const std::vector<double>& b_scaled, double c_scaling,
                                 std::vector<double>* y);
  void GatherVars(int cone, const double* y, std::vector<double>* z) const;

  int m_;
  int num_dual_ = 0;
  std::vector<int> transpositions_;  // RLDLT pivots (LDLT mode: any equality present)
  // conex/kkt_solver.cc:174-178,229-261: with iterative refinement the assembled matrix is kept
  // and every solve is followed by `iterative_refinement_iterations_` correction solves.
  int iterative_refinement_iterations_ = 0;
  std::vector<double> kkt_matrix_;
  std::vector<std::unique_ptr<Cone>> cones_;
  std::vector<std::vector<int>> cliques_;
  std::vector<SchurSystem> cone_sys_;
  std::vector<double> arena_;  // conex/cone_program.h:174-189 — the warm-start state
  std::vector<double> linear_cost_;
  double* b_scaling_ = nullptr;
  double* c_scaling_ = nullptr;
  bool is_initialized_ = false;
  int stats_max_iter_ = 0;
};

}  // namespace oracle
