// ORACLE — TEST INFRASTRUCTURE ONLY. See program.h.
#include "program.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>

namespace oracle {
namespace {

struct Timer {
  std::chrono::high_resolution_clock::time_point t0 = std::chrono::high_resolution_clock::now();
  double Seconds() const {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  }
};

double Norm2(const std::vector<double>& v) { return std::sqrt(Dot(v.size(), v.data(), v.data())); }
double Norm2(const double* v, int n) { return std::sqrt(Dot(n, v, v)); }

// conex/cone_program.cc:166-172
double MinimizeNormInf(const SlackEigenvalues& p) {
  return (p.lambda_min > 0) ? 2.0 / (p.lambda_min + p.lambda_max) : -1;
}

}  // namespace

bool Program::AddCone(std::unique_ptr<Cone> cone, const std::vector<int>& vars) {
  std::vector<int> seen(m_, 0);
  for (int v : vars) {
    if (v < 0 || v >= m_ || seen[v]++) return false;
  }
  cones_.push_back(std::move(cone));
  cliques_.push_back(vars);
  is_initialized_ = false;
  return true;
}

bool Program::AddCone(std::unique_ptr<Cone> cone) {
  std::vector<int> all(m_);
  for (int i = 0; i < m_; i++) all[i] = i;
  return AddCone(std::move(cone), all);
}

bool Program::AddEquality(int rows, const double* A, const double* b, const std::vector<int>& vars) {
  std::vector<int> seen(m_, 0);
  for (int v : vars) {
    if (v < 0 || v >= m_ || seen[v]++) return false;
  }
  std::vector<int> clique(vars);
  for (int i = 0; i < rows; i++) clique.push_back(m_ + num_dual_ + i);
  num_dual_ += rows;
  cones_.push_back(std::make_unique<EqualityCone>(rows, (int)vars.size(), A, b));
  cliques_.push_back(clique);
  is_initialized_ = false;
  return true;
}

bool Program::AddEquality(int rows, const double* A, const double* b) {
  std::vector<int> all(m_);
  for (int i = 0; i < m_; i++) all[i] = i;
  return AddEquality(rows, A, b, all);
}

void Program::GatherVars(int c, const double* y, std::vector<double>* z) const {
  // conex/cone_program.h:59-67
  const auto& cl = cliques_[c];
  z->resize(cl.size());
  for (size_t i = 0; i < cl.size(); i++) (*z)[i] = y[cl[i]];
}

bool Program::Initialize(const SolverConfiguration& config) {
  // conex/cone_program.cc:78-112 and conex/cone_program.h:174-189
  if (is_initialized_ && config.initialization_mode != 0) return true;
  cone_sys_.assign(cones_.size(), SchurSystem());
  size_t total = 0;
  for (size_t c = 0; c < cones_.size(); c++) {
    cone_sys_[c].m = cones_[c]->NumberOfVariables();
    cone_sys_[c].residual_only = false;
    total += cones_[c]->WorkspaceSize() + cone_sys_[c].SizeOf();
  }
  const size_t stats_offset = total;
  total += 2;  // c_scaling, b_scaling (conex/workspace.h:76-88; the per-iteration arrays live in
               // sqrt_inv_mu so that warm starts with a larger max_iterations stay in bounds)
  sys.m = SizeOfKKTSystem();
  sys.residual_only = true;
  const size_t sys_offset = total;
  total += sys.SizeOf();
  if (total > arena_.size()) arena_.resize(total);
  double* p = arena_.data();
  for (size_t c = 0; c < cones_.size(); c++) {
    cones_[c]->BindWorkspace(p);
    p += cones_[c]->WorkspaceSize();
    cone_sys_[c].Bind(p);
    p += cone_sys_[c].SizeOf();
  }
  c_scaling_ = arena_.data() + stats_offset;
  b_scaling_ = arena_.data() + stats_offset + 1;
  sys.Bind(arena_.data() + sys_offset);
  H.assign((size_t)SizeOfKKTSystem() * SizeOfKKTSystem(), 0.0);
  is_initialized_ = true;
  if (config.initialization_mode == 0) {
    *b_scaling_ = 1;
    *c_scaling_ = 1;
    for (auto& c : cones_) c->SetIdentity();
  }
  return true;
}

void Program::Assemble() {
  // conex/kkt_solver.cc:164-170 + conex/supernodal_assembler.cc:113-165 (dense equivalent),
  // then conex/constraint_manager.h:107-124.
  std::fill(H.begin(), H.end(), 0.0);
  sys.SetZero();
  for (size_t c = 0; c < cones_.size(); c++) {
    SchurSystem& s = cone_sys_[c];
    if (auto* lmi = dynamic_cast<DenseLmiCone*>(cones_[c].get())) lmi->gram_variant = gram_variant;
    cones_[c]->ConstructSchurComplementSystem(true, &s);
    const auto& cl = cliques_[c];
    for (size_t a = 0; a < cl.size(); a++) {
      for (size_t b = 0; b <= a; b++) {
        const int ga = std::max(cl[a], cl[b]), gb = std::min(cl[a], cl[b]);
        H[(size_t)gb * SizeOfKKTSystem() + ga] += s.G((int)a, (int)b);
      }
    }
    sys.inner_product_of_w_and_c += s.inner_product_of_w_and_c;
    sys.inner_product_of_c_and_Qc += s.inner_product_of_c_and_Qc;
    for (size_t a = 0; a < cl.size(); a++) {
      sys.AW[cl[a]] += s.AW[a];
      sys.AQc[cl[a]] += s.AQc[a];
    }
  }
}

bool Program::Factor() {
  // conex/kkt_solver.cc:172-199: Cholesky unless some cone carries multipliers, then the
  // regularised LDL^T, which never reports failure.
  const int N = SizeOfKKTSystem();
  if (iterative_refinement_iterations_ > 0) kkt_matrix_ = H;  // kkt_solver.cc:174-178 (lower triangle)
  if (num_dual_ == 0) return CholeskyLower(N, H.data(), N);
  LdltLower(N, H.data(), N, &transpositions_);
  return true;
}

void Program::SolveInPlace(double* rhs) const {
  // conex/kkt_solver.cc:220-263 for one dense supernode (identity permutation).
  const int N = SizeOfKKTSystem();
  auto solve_once = [&](double* v) {
    if (num_dual_ == 0) {
      SolveLower(N, H.data(), N, v, false);
      SolveLower(N, H.data(), N, v, true);
    } else {
      SolveLdlt(N, H.data(), N, transpositions_, v);
    }
  };
  if (iterative_refinement_iterations_ <= 0) {
    solve_once(rhs);
    return;
  }
  const std::vector<double> total_residual(rhs, rhs + N);
  solve_once(rhs);
  std::vector<double> r(N);
  for (int it = 0; it < iterative_refinement_iterations_; it++) {
    // residual = total_residual - K y with K given by its lower triangle (kkt_solver.cc:250-251)
    for (int i = 0; i < N; i++) {
      double s = 0;
      for (int j = 0; j < N; j++) {
        const double kij = (i >= j) ? kkt_matrix_[(size_t)j * N + i] : kkt_matrix_[(size_t)i * N + j];
        s += kij * rhs[j];
      }
      r[i] = total_residual[i] - s;
    }
    solve_once(r.data());
    for (int i = 0; i < N; i++) rhs[i] += r[i];
  }
}

double Program::ComputeMuFromDivergence(const SolverConfiguration& config, int rank,
                                        const std::vector<double>& b_scaled, double c_scaling,
                                        std::vector<double>* y) {
  // conex/cone_program.cc:173-214 with AQc := sys.AQc * c_scaling, b := b * b_scaling.
  for (int i = 0; i < SizeOfKKTSystem(); i++) (*y)[i] = sys.AQc[i] * c_scaling - b_scaled[i];
  SolveInPlace(y->data());
  // conex/cone_program.cc:31-57
  SlackEigenvalues p;
  p.frobenius_norm_squared = 0;
  p.trace = 0;
  p.lambda_max = -30000;
  p.lambda_min = 30000;
  std::vector<double> z;
  for (size_t c = 0; c < cones_.size(); c++) {
    GatherVars((int)c, y->data(), &z);
    SlackEigenvalues t;
    cones_[c]->GetWeightedSlackEigenvalues(z.data(), c_scaling, &t);
    p.lambda_max = std::max(p.lambda_max, t.lambda_max);
    p.lambda_min = std::min(p.lambda_min, t.lambda_min);
    p.frobenius_norm_squared += t.frobenius_norm_squared;
    p.trace += t.trace;
  }
  p.rank = rank;
  double k = DivergenceUpperBoundInverse(config.divergence_upper_bound * rank, p);
  if (k == -1) k = MinimizeNormInf(p);
  if (k < 0 && p.trace > 1e-12) {
    // conex/cone_program.cc:194-211: fall back to a norm bound.
    const double kstar = p.trace / p.frobenius_norm_squared;
    double norm_bound = 1.5 * (p.frobenius_norm_squared * kstar * kstar - 2 * p.trace * kstar + rank);
    norm_bound = std::min(norm_bound, rank * .7);
    const double a = p.frobenius_norm_squared, b = -2 * p.trace, c = rank - norm_bound;
    const double disc = b * b - 4 * a * c;
    k = (disc < 0) ? p.trace / p.frobenius_norm_squared : (-b + std::sqrt(disc)) / (2 * a);
  }
  return k;
}

std::vector<double> Program::FeasibleObjective() {
  Initialize(SolverConfiguration());
  Assemble();
  std::vector<double> b(m_);
  for (int i = 0; i < m_; i++) b[i] = .5 * sys.AW[i];
  return b;
}

bool Program::Maximize(const double* b_in, const SolverConfiguration& config, double* yout) {
  iterative_refinement_iterations_ = config.iterative_refinement_iterations;  // cone_program.cc:303-304
  const int mv = m_;                  // variables the caller sees
  const int m = SizeOfKKTSystem();    // + multipliers of the equality constraints
  status = Status();
  log.clear();
  seconds = PhaseSeconds();
  bool max_iters_reached = true;
  if (cones_.empty()) {
    // conex/cone_program.cc:266-271
    for (int i = 0; i < mv; i++) yout[i] = b_in[i] * std::numeric_limits<double>::infinity();
    return false;
  }
  Initialize(config);
  sqrt_inv_mu.assign(std::max(config.max_iterations, 1), 0.0);
  const bool warm = config.initialization_mode != 0;

  std::vector<double> b(m, 0.0), y(m, 0.0), b_scaled(m), z;  // cone_program.cc:294-296
  std::copy(b_in, b_in + mv, b.begin());
  double k = 0;  // inv_sqrt_mu
  double kmax = config.inv_sqrt_mu_max;
  double cx = 1, by = -1, kkt_error = 0;
  int rank = 0;
  for (auto& c : cones_) rank += c->Rank();
  int centering_steps = 0;
  bool warmstart_aborted = false;
  const int init_steps =
      warm ? config.initial_centering_steps_warmstart : config.initial_centering_steps_coldstart;
  double& b_scaling = *b_scaling_;
  double& c_scaling = *c_scaling_;
  int num_iter = 0;

  for (int i = 0; i < config.max_iterations; i++) {
    const bool initial_centering = i < init_steps;
    const bool final_centering = (k >= kmax) || (kkt_error > config.kkt_error_tolerance) ||
                                 i >= (config.max_iterations - config.final_centering_steps);
    const bool update_mu = (i == 0) || !(initial_centering || final_centering) || warmstart_aborted;
    warmstart_aborted = false;
    if (final_centering && centering_steps >= config.final_centering_steps) {
      max_iters_reached = (i >= config.max_iterations - 1);
      break;
    }
    {
      Timer t;
      Assemble();
      seconds.assemble += t.Seconds();
    }
    if (i < 1 && config.enable_rescaling) {
      if (!warm) {
        b_scaling = 1.0 / (1 + Norm2(b));
        c_scaling = 1.0 / (1 + Norm2(sys.AQc, m));
      }
      double mu_target = 1.0 / (kmax * kmax);
      mu_target *= (b_scaling * c_scaling);
      kmax = 1.0 / std::sqrt(mu_target);
    }
    {
      Timer t;
      const bool ok = Factor();
      seconds.factor += t.Seconds();
      if (!ok) {
        if (i == 0 && warm) {
          for (auto& c : cones_) c->SetIdentity();
          warmstart_aborted = true;
          continue;
        }
        status.solved = 0;
        return false;
      }
    }
    for (int j = 0; j < m; j++) b_scaled[j] = b[j] * b_scaling;
    if (update_mu) {
      Timer t;
      double temp = -1;
      if (config.enable_line_search) {
        // PSD cones do not implement PerformLineSearch (conex/constraint.h:24-28): the search
        // reports failure and the previous value is kept (conex/cone_program.cc:376-384).
        temp = k;
      }
      if (temp < 0) temp = ComputeMuFromDivergence(config, rank, b_scaled, c_scaling, &y);
      if (temp > 0) {
        k = temp;
      } else {
        k *= .5;
      }
      seconds.mu += t.Seconds();
    } else if (!initial_centering) {
      centering_steps++;
    }
    k = std::min(k, kmax);
    k = std::max(k, std::sqrt(1.0 / (1e-15 + config.maximum_mu)));

    for (int j = 0; j < m; j++) y[j] = k * (b_scaled[j] + sys.AQc[j] * c_scaling) - 2 * sys.AW[j];
    {
      Timer t;
      SolveInPlace(y.data());
      seconds.solve += t.Seconds();
    }
    StepOptions opt;
    opt.affine = false;
    opt.inv_sqrt_mu = k;
    opt.e_weight = 1;
    opt.c_weight = k * c_scaling;
    StepInfo info;
    info.normsqrd = 0;
    info.norminfd = -1;
    Timer tu;
    for (size_t c = 0; c < cones_.size(); c++) {
      // conex/cone_program.h:69-90
      GatherVars((int)c, y.data(), &z);
      StepInfo ci;
      cones_[c]->PrepareStep(opt, z.data(), &ci);
      info.norminfd = std::max(info.norminfd, ci.norminfd);
      info.normsqrd += ci.normsqrd;
    }
    opt.step_size = std::min(1.0, 2.0 / (info.norminfd * info.norminfd));
    if (i == 0 && warm && info.norminfd >= config.warmstart_abort_threshold) {
      for (auto& c : cones_) c->SetIdentity();
      warmstart_aborted = true;
    } else {
      for (auto& c : cones_) c->TakeStep(opt);
    }
    seconds.update += tu.Seconds();

    const double d_2 = std::sqrt(std::fabs(info.normsqrd));
    const double d_inf = std::fabs(info.norminfd);
    by = Dot(m, b.data(), y.data()) / (k * c_scaling);
    cx = 2 * sys.inner_product_of_w_and_c + Dot(m, sys.AQc, y.data()) -
         k * sys.inner_product_of_c_and_Qc * c_scaling;
    cx /= (k * b_scaling);
    double mu = 1.0 / k;
    mu *= mu;
    const double s_dot_x = mu * (rank - d_2 * d_2) / (b_scaling * c_scaling);
    mu = mu / (c_scaling * b_scaling);
    kkt_error = std::fabs(cx - by - s_dot_x) / s_dot_x;
    num_iter = i + 1;
    sqrt_inv_mu[i] = k;
    log.push_back({k, mu, d_2, d_inf, by, cx, kkt_error, opt.step_size});

    if ((final_centering || k >= kmax) && d_inf <= config.final_centering_tolerance) {
      max_iters_reached = false;
      break;
    }
  }

  status.num_iterations = num_iter;
  for (int j = 0; j < mv; j++) yout[j] = y[j];
  const double mu_final = (1.0 / k) * (1.0 / k);
  if (mu_final > config.infeasibility_threshold) {
    status.solved = 0;
    status.primal_infeasible = cx * k <= -.5;
    status.dual_infeasible = by * k >= .5;
  } else {
    status.solved = 1;
  }
  if (config.prepare_dual_variables) {
    // conex/cone_program.cc:500-516
    Assemble();
    Factor();
    std::vector<double> y2(m);
    for (int j = 0; j < m; j++) y2[j] = k * b[j] * b_scaling - sys.AW[j];
    SolveInPlace(y2.data());
    StepOptions opt;
    opt.affine = true;
    opt.inv_sqrt_mu = k;
    opt.e_weight = 0;
    opt.c_weight = 0;
    StepInfo info;
    for (size_t c = 0; c < cones_.size(); c++) {
      GatherVars((int)c, y2.data(), &z);
      cones_[c]->PrepareStep(opt, z.data(), &info);
    }
  }
  if (status.solved) {
    for (int j = 0; j < mv; j++) yout[j] = yout[j] / k / c_scaling;
    if (max_iters_reached) status.solved = 0;
  }
  return status.solved != 0;
}

int Program::GetDualVariableSize(int i) const {
  if (i < 0 || i >= (int)cones_.size()) return -1;
  return cones_[i]->DualVariableSize();
}

void Program::GetDualVariable(int i, double* x) const {
  if (i < 0 || i >= (int)cones_.size()) return;
  const int sz = cones_[i]->DualVariableSize();
  const double* w = cones_[i]->DualVariable();
  double scale = 1;
  if (!status.primal_infeasible && status.num_iterations > 0) {
    scale = 1.0 / (sqrt_inv_mu[status.num_iterations - 1] * (*b_scaling_));
  }
  for (int j = 0; j < sz; j++) x[j] = w[j] * scale;
}

}  // namespace oracle
