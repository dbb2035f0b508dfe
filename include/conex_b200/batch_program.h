// Many small, structurally identical cone programs advanced in lock step on one GPU — the
// "batched small-block SDP" regime (BASELINE config 3; SURVEY.md §8e: independent units, no
// collective). The reference solves such programs one after the other on the CPU, each through
// Solve() (conex/cone_program.cc:235-533); distinct Program handles are independent
// (interfaces/conex.cc has no shared state), which is what makes batching legal.
//
// Design: the cone data of all B programs is packed cone by cone into batched device arrays; every
// phase of the Newton step is ONE launch per cone with one CTA per program (device/small_cones.cu),
// the B small KKT systems are factored and solved by one launch each, and the per-program control
// flow of Solve() — centering schedule, mu rule, rescaling, step size, termination, status — runs on
// the host over arrays of B scalars fetched once per phase. Programs that have terminated are masked
// out of later launches, so every program follows exactly the iteration sequence it would follow
// alone.
#pragma once
#include <memory>
#include <vector>

#include "cone_program.h"

namespace conex {

struct BatchResult {
  int solved = 0;
  int num_iterations = 0;
  int primal_infeasible = 0;
  int dual_infeasible = 0;
  double by = 0, cx = 0, inv_sqrt_mu = 0;
  double b_scaling = 1, c_scaling = 1, d_inf = 0;
};

class BatchProgram {
 public:
  // All programs must have the same number of variables and the same list of cones (type, order),
  // every cone on all variables in order, no equality constraints. Throws otherwise.
  explicit BatchProgram(const std::vector<Program*>& programs);
  ~BatchProgram();

  int size() const { return batch_; }
  int number_of_variables() const { return m_; }
  // b, y: m x B column-major host arrays. Returns the number of solved programs.
  int Maximize(const double* b, const SolverConfiguration& config, double* y);
  const std::vector<BatchResult>& results() const { return results_; }
  // Device milliseconds of the last Maximize (CUDA events on the batch's stream) and of its
  // individual Newton steps.
  double milliseconds() const { return total_ms_; }
  const std::vector<float>& step_milliseconds() const { return step_ms_; }
  // Scaling point of cone `cone` of program `p`, rescaled like Program::GetDualVariable.
  int DualVariableSize(int cone) const;
  void GetDualVariable(int p, int cone, double* host_out);

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
  int batch_ = 0, m_ = 0;
  std::vector<BatchResult> results_;
  std::vector<float> step_ms_;
  double total_ms_ = 0;
};

}  // namespace conex
