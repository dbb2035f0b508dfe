// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C ABI of the CPU oracle. Two groups of symbols:
//  (1) the subset of interfaces/conex.h (CONEX_*) that the hot path uses, with the reference's
//      argument meaning and return conventions (interfaces/conex.cc), so the same ctypes
//      harness can drive the oracle and the B200 library;
//  (2) ORACLE_* entry points exposing individual kernels (Schur assembly, Padé map, Lanczos,
//      Cholesky, mu rule, PSD step functions) for kernel-level parity tests.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "../include/conex.h"
#include "program.h"

using oracle::DenseLmiCone;
using oracle::LinearCone;
using oracle::SocCone;
using oracle::HermitianLmiCone;
using oracle::Program;

namespace {

oracle::SolverConfiguration Convert(const CONEX_SolverConfiguration* c) {
  // interfaces/conex.cc:65-90
  oracle::SolverConfiguration o;
  o.prepare_dual_variables = c->prepare_dual_variables;
  o.initialization_mode = c->initialization_mode;
  o.inv_sqrt_mu_max = c->inv_sqrt_mu_max;
  o.minimum_mu = c->minimum_mu;
  o.maximum_mu = c->maximum_mu;
  o.divergence_upper_bound = c->divergence_upper_bound;
  o.enable_line_search = c->enable_line_search;
  o.dinf_upper_bound = c->dinf_upper_bound;
  o.final_centering_steps = c->final_centering_steps;
  o.final_centering_tolerance = c->final_centering_tolerance;
  o.initial_centering_steps_warmstart = c->initial_centering_steps_warmstart;
  o.initial_centering_steps_coldstart = c->initial_centering_steps_coldstart;
  o.warmstart_abort_threshold = c->warmstart_abort_threshold;
  o.max_iterations = c->max_iterations;
  o.iterative_refinement_iterations = c->iterative_refinement_iterations;
  o.infeasibility_threshold = c->infeasibility_threshold;
  o.kkt_error_tolerance = c->kkt_error_tolerance;
  o.enable_rescaling = c->enable_rescaling;
  o.kkt_solver = c->kkt_solver;
  return o;
}

Program* Cast(void* p) { return static_cast<Program*>(p); }

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------- CONEX_* subset
void* CONEX_CreateConeProgram() { return new Program(0); }
void CONEX_DeleteConeProgram(void* p) { delete Cast(p); }

CONEX_STATUS CONEX_SetNumberOfVariables(void* p, int m) {
  // interfaces/conex.cc:399-407
  if (!p || m < 1) return CONEX_FAILURE;
  if (Cast(p)->NumberOfVariables() != 0) return CONEX_FAILURE;
  Cast(p)->SetNumberOfVariables(m);
  return CONEX_SUCCESS;
}

int CONEX_AddDenseLMIConstraint(void* p, const double* A, int Ar, int Ac, int m, const double* C,
                                int cr, int cc) {
  // interfaces/conex.cc:137-160
  (void)Ar;
  (void)Ac;
  (void)cr;
  Program* prog = Cast(p);
  const int id = prog->NumberOfConstraints();
  if (prog->NumberOfVariables() == 0) prog->SetNumberOfVariables(m);
  prog->AddCone(std::make_unique<DenseLmiCone>(cc, m, A, C));
  return id;
}

int CONEX_AddSparseLMIConstraint(void* p, const double* A, int Ar, int Ac, int num_vars,
                                 const double* C, int cr, int cc, const long* vars, int vars_rows) {
  // interfaces/conex.cc:162-188
  (void)Ar;
  (void)Ac;
  (void)cr;
  (void)vars_rows;
  Program* prog = Cast(p);
  const int id = prog->NumberOfConstraints();
  std::vector<int> v(num_vars);
  for (int i = 0; i < num_vars; i++) v[i] = (int)vars[i];
  prog->AddCone(std::make_unique<DenseLmiCone>(cc, num_vars, A, C), v);
  return id;
}

int CONEX_AddDenseLinearConstraint(void* p, const double* A, int Ar, int Ac, const double* c,
                                   int cr) {
  // interfaces/conex.cc:216-229
  (void)cr;
  Program* prog = Cast(p);
  const int id = prog->NumberOfConstraints();
  if (prog->NumberOfVariables() == 0) prog->SetNumberOfVariables(Ac);
  prog->AddCone(std::make_unique<LinearCone>(Ar, Ac, A, c));
  return id;
}

int CONEX_AddLinearInequalities(void* p, const double* A, int Ar, int Ac, const double* lb,
                                int num_lb, const double* ub, int num_ub) {
  // interfaces/conex.cc:190-215 + PreprocessLinearInequality (linear_constraint.cc:21-46):
  // rows with lb == ub become scaled equalities; finite bounds become scaled inequality rows.
  (void)num_lb;
  (void)num_ub;
  Program* prog = Cast(p);
  std::vector<std::vector<double>> ineq_rows, eq_rows;
  std::vector<double> ineq_rhs, eq_rhs;
  for (int i = 0; i < Ar; i++) {
    double rr = 0;
    for (int j = 0; j < Ac; j++) rr += A[(size_t)j * Ar + i] * A[(size_t)j * Ar + i];
    auto row = [&](double scale) {
      std::vector<double> r(Ac);
      for (int j = 0; j < Ac; j++) r[j] = scale * A[(size_t)j * Ar + i];
      return r;
    };
    if (lb[i] == ub[i]) {
      const double scale = 1.0 / std::sqrt(rr + ub[i] * ub[i]);
      eq_rows.push_back(row(scale));
      eq_rhs.push_back(scale * ub[i]);
    } else {
      if (ub[i] < 1e8) {
        const double scale = 1.0 / std::sqrt(rr + ub[i] * ub[i]);
        ineq_rows.push_back(row(scale));
        ineq_rhs.push_back(scale * ub[i]);
      }
      if (lb[i] > -1e8) {
        const double scale = 1.0 / std::sqrt(rr + lb[i] * lb[i]);
        ineq_rows.push_back(row(-scale));
        ineq_rhs.push_back(-scale * lb[i]);
      }
    }
  }
  auto pack = [&](const std::vector<std::vector<double>>& rows) {
    const int r = (int)rows.size();
    std::vector<double> M((size_t)r * Ac);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < Ac; j++) M[(size_t)j * r + i] = rows[i][j];
    return M;
  };
  if (!ineq_rows.empty()) {
    const auto M = pack(ineq_rows);
    prog->AddCone(std::make_unique<LinearCone>((int)ineq_rows.size(), Ac, M.data(), ineq_rhs.data()));
  }
  if (!eq_rows.empty()) {
    const auto M = pack(eq_rows);
    prog->AddEquality((int)eq_rows.size(), M.data(), eq_rhs.data());
  }
  return -1;
}

CONEX_STATUS CONEX_NewLorentzConeConstraint(void* p, int order, int* constraint_id) {
  // interfaces/conex.cc:384-397
  if (order < 1 || !constraint_id || !p) return CONEX_FAILURE;
  Program* prog = Cast(p);
  prog->AddCone(std::make_unique<SocCone>(order, prog->NumberOfVariables(), nullptr, nullptr));
  *constraint_id = prog->NumberOfConstraints() - 1;
  return CONEX_SUCCESS;
}

CONEX_STATUS CONEX_NewLinearMatrixInequality(void* p, int order, int hyper_complex_dim, int* constraint_id) {
  // interfaces/conex.cc:287-316; only the real algebra is restated
  if (order < 1 || !constraint_id || !p || hyper_complex_dim != 1) return CONEX_FAILURE;
  Program* prog = Cast(p);
  prog->AddCone(std::make_unique<HermitianLmiCone>(order, prog->NumberOfVariables()));
  *constraint_id = prog->NumberOfConstraints() - 1;
  return CONEX_SUCCESS;
}

CONEX_STATUS CONEX_NewLinearInequality(void* p, int num_rows, int* constraint_id) {
  // interfaces/conex.cc:318-329
  if (!constraint_id || !p) return CONEX_FAILURE;
  Program* prog = Cast(p);
  const int m = prog->NumberOfVariables();
  std::vector<double> A((size_t)num_rows * m, 0.0), c(num_rows, 0.0);
  prog->AddCone(std::make_unique<LinearCone>(num_rows, m, A.data(), c.data()));
  *constraint_id = prog->NumberOfConstraints() - 1;
  return CONEX_SUCCESS;
}

CONEX_STATUS CONEX_UpdateLinearOperator(void* p, int constraint, double value, int variable, int row,
                                        int col, int hyper_complex_dim) {
  // interfaces/conex.cc:365-373; linear_constraint.cc:207-216; soc_constraint.cc:237-247
  Program* prog = Cast(p);
  if (constraint < 0 || constraint >= prog->NumberOfConstraints()) return CONEX_FAILURE;
  if (variable < 0 || row < 0 || variable >= prog->NumberOfVariables()) return CONEX_FAILURE;
  if (auto* lmi = dynamic_cast<HermitianLmiCone*>(prog->cone(constraint))) {
    // hermitian_psd.cc:248-275
    if (hyper_complex_dim != 0 || col < 0 || row >= lmi->order() || col >= lmi->order()) return CONEX_FAILURE;
    lmi->SetOperatorEntry(variable, row, col, value);
    return CONEX_SUCCESS;
  }
  if (hyper_complex_dim != 0 || col != 0) return CONEX_FAILURE;
  if (auto* lp = dynamic_cast<LinearCone*>(prog->cone(constraint))) {
    if (row >= lp->rows()) return CONEX_FAILURE;
    lp->SetOperatorEntry(row, variable, value);
    return CONEX_SUCCESS;
  }
  if (auto* soc = dynamic_cast<SocCone*>(prog->cone(constraint))) {
    if (row > soc->order()) return CONEX_FAILURE;
    soc->SetOperatorEntry(row, variable, value);
    return CONEX_SUCCESS;
  }
  return CONEX_FAILURE;
}

CONEX_STATUS CONEX_UpdateAffineTerm(void* p, int constraint, double value, int row, int col,
                                    int hyper_complex_dim) {
  // interfaces/conex.cc:375-382; linear_constraint.cc:218-226; soc_constraint.cc:249-259
  Program* prog = Cast(p);
  if (constraint < 0 || constraint >= prog->NumberOfConstraints()) return CONEX_FAILURE;
  if (row < 0) return CONEX_FAILURE;
  if (auto* lmi = dynamic_cast<HermitianLmiCone*>(prog->cone(constraint))) {
    // hermitian_psd.cc:284-313
    if (hyper_complex_dim != 0 || col < 0 || row >= lmi->order() || col >= lmi->order()) return CONEX_FAILURE;
    lmi->SetAffineEntry(row, col, value);
    return CONEX_SUCCESS;
  }
  if (hyper_complex_dim != 0 || col != 0) return CONEX_FAILURE;
  if (auto* lp = dynamic_cast<LinearCone*>(prog->cone(constraint))) {
    if (row >= lp->rows()) return CONEX_FAILURE;
    lp->SetAffineEntry(row, value);
    return CONEX_SUCCESS;
  }
  if (auto* soc = dynamic_cast<SocCone*>(prog->cone(constraint))) {
    if (row > soc->order()) return CONEX_FAILURE;
    soc->SetAffineEntry(row, value);
    return CONEX_SUCCESS;
  }
  return CONEX_FAILURE;
}

int CONEX_Maximize(void* p, const double* b, int br, const CONEX_SolverConfiguration* config,
                   double* y, int yr) {
  // interfaces/conex.cc:93-105
  (void)br;
  (void)yr;
  return Cast(p)->Maximize(b, Convert(config), y) ? 1 : 0;
}

void CONEX_GetDualVariable(void* p, int i, double* x, int xr, int xc) {
  (void)xr;
  (void)xc;
  Cast(p)->GetDualVariable(i, x);
}

int CONEX_GetDualVariableSize(void* p, int i) { return Cast(p)->GetDualVariableSize(i); }

void CONEX_SetDefaultOptions(CONEX_SolverConfiguration* c) {
  // interfaces/conex.cc:231-257 (plus the two fields the reference leaves unset)
  if (!c) return;
  oracle::SolverConfiguration d;
  c->prepare_dual_variables = d.prepare_dual_variables;
  c->initialization_mode = d.initialization_mode;
  c->inv_sqrt_mu_max = d.inv_sqrt_mu_max;
  c->minimum_mu = d.minimum_mu;
  c->maximum_mu = d.maximum_mu;
  c->divergence_upper_bound = d.divergence_upper_bound;
  c->enable_line_search = d.enable_line_search;
  c->dinf_upper_bound = d.dinf_upper_bound;
  c->final_centering_steps = d.final_centering_steps;
  c->final_centering_tolerance = d.final_centering_tolerance;
  c->initial_centering_steps_warmstart = d.initial_centering_steps_warmstart;
  c->initial_centering_steps_coldstart = d.initial_centering_steps_coldstart;
  c->warmstart_abort_threshold = d.warmstart_abort_threshold;
  c->max_iterations = d.max_iterations;
  c->iterative_refinement_iterations = d.iterative_refinement_iterations;
  c->infeasibility_threshold = d.infeasibility_threshold;
  c->kkt_error_tolerance = d.kkt_error_tolerance;
  c->enable_rescaling = d.enable_rescaling;
  c->kkt_solver = d.kkt_solver;
}

void CONEX_GetIterationStats(void* p, CONEX_IterationStats* stats, int iter_circular) {
  // interfaces/conex.cc:259-285
  if (!p || !stats) return;
  Program* prog = Cast(p);
  int iter = iter_circular;
  if (iter < 0) iter = prog->status.num_iterations + iter;
  if (iter < 0 || iter >= prog->status.num_iterations) return;
  const double k = prog->sqrt_inv_mu[iter];
  stats->mu = 1.0 / (k * k);
  stats->iteration_number = iter;
}

// ---------------------------------------------------------------------------- ORACLE_* extras
// C++-only constructors of the reference, exposed for the tests:
// SOCConstraint(A, c) (soc_constraint.h:9-15): A is (n+1) x m column-major.
int ORACLE_AddSocConstraint(void* p, int n, int m, const double* A, const double* c) {
  Program* prog = Cast(p);
  const int id = prog->NumberOfConstraints();
  if (prog->NumberOfVariables() == 0) prog->SetNumberOfVariables(m);
  prog->AddCone(std::make_unique<SocCone>(n, m, A, c));
  return id;
}
// Program::AddConstraint(EqualityConstraints{A, b}[, vars]) (cone_program.h:193-217): A is
// rows x nvars column-major; vars == nullptr means all variables.
int ORACLE_AddEqualityConstraint(void* p, int rows, int nvars, const double* A, const double* b,
                                 const long* vars) {
  Program* prog = Cast(p);
  const int id = prog->NumberOfConstraints();
  bool ok;
  if (vars) {
    std::vector<int> v(nvars);
    for (int i = 0; i < nvars; i++) v[i] = (int)vars[i];
    ok = prog->AddEquality(rows, A, b, v);
  } else {
    ok = prog->AddEquality(rows, A, b);
  }
  return ok ? id : -1;
}
// Stand-alone cones for kernel-level parity tests. type: 0 LP (rows = n), 1 SOC (rows = n + 1),
// 2 dense LMI (rows = n * n). data = rows x (m + 1) column-major: the operator, then the affine term.
struct OracleCone {
  std::unique_ptr<oracle::Cone> cone;
  std::vector<double> workspace;
  int m;
};
void* ORACLE_ConeCreate(int type, int n, int m, const double* data) {
  auto* h = new OracleCone;
  h->m = m;
  if (type == 0) {
    h->cone = std::make_unique<LinearCone>(n, m, data, data + (size_t)n * m);
  } else if (type == 1) {
    h->cone = std::make_unique<SocCone>(n, m, data, data + (size_t)(n + 1) * m);
  } else {
    h->cone = std::make_unique<DenseLmiCone>(n, m, data, data + (size_t)n * n * m);
  }
  h->workspace.assign(h->cone->WorkspaceSize() + 8, 0.0);
  h->cone->BindWorkspace(h->workspace.data());
  h->cone->SetIdentity();
  return h;
}
void ORACLE_ConeDelete(void* h) { delete static_cast<OracleCone*>(h); }
int ORACLE_ConeStateSize(void* h) { return static_cast<OracleCone*>(h)->cone->StateSize(); }
void ORACLE_ConeGetState(void* h, double* w) { static_cast<OracleCone*>(h)->cone->GetState(w); }
void ORACLE_ConeSetState(void* h, const double* w) { static_cast<OracleCone*>(h)->cone->SetState(w); }
// G: m x m column-major (lower triangle written), AW, AQc: m, scal2 = {<w,c>, <c,Qc>}
void ORACLE_ConeSchur(void* h, double* G, double* AW, double* AQc, double* scal2) {
  auto* c = static_cast<OracleCone*>(h);
  const int m = c->m;
  oracle::SchurSystem sys;
  sys.m = m;
  std::vector<double> buf(sys.SizeOf(), 0.0);
  sys.Bind(buf.data());
  c->cone->ConstructSchurComplementSystem(true, &sys);
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) G[(size_t)j * m + i] = (i >= j) ? sys.G(i, j) : 0.0;
  std::memcpy(AW, sys.AW, sizeof(double) * m);
  std::memcpy(AQc, sys.AQc, sizeof(double) * m);
  scal2[0] = sys.inner_product_of_w_and_c;
  scal2[1] = sys.inner_product_of_c_and_Qc;
}
// out4 = {lambda_min, lambda_max, frobenius_norm_squared, trace}
void ORACLE_ConeEigen(void* h, const double* y, double c_weight, double* out4) {
  oracle::SlackEigenvalues p;
  static_cast<OracleCone*>(h)->cone->GetWeightedSlackEigenvalues(y, c_weight, &p);
  out4[0] = p.lambda_min;
  out4[1] = p.lambda_max;
  out4[2] = p.frobenius_norm_squared;
  out4[3] = p.trace;
}
void ORACLE_ConePrepare(void* h, const double* y, int affine, double c_weight, double e_weight,
                        double* out2) {
  oracle::StepOptions opt;
  opt.affine = affine != 0;
  opt.c_weight = c_weight;
  opt.e_weight = e_weight;
  oracle::StepInfo info;
  static_cast<OracleCone*>(h)->cone->PrepareStep(opt, y, &info);
  out2[0] = info.norminfd;
  out2[1] = info.normsqrd;
}
void ORACLE_ConeTakeStep(void* h, double step, double e_weight) {
  oracle::StepOptions opt;
  opt.affine = false;
  opt.e_weight = e_weight;
  opt.step_size = step;
  static_cast<OracleCone*>(h)->cone->TakeStep(opt);
}

void ORACLE_TaylorExpm(int n, const double* X, double* out) { oracle::ExponentialMapTaylor(n, X, out); }
int ORACLE_HermitianLanczos(int n, const double* WS, const double* W, const double* r, int num_iter,
                            double* out) {
  const auto e = oracle::HermitianLanczos(n, WS, W, r, num_iter);
  std::memcpy(out, e.data(), e.size() * sizeof(double));
  return (int)e.size();
}

int ORACLE_SizeOfKKTSystem(void* p) { return Cast(p)->SizeOfKKTSystem(); }
int ORACLE_NumberOfConstraints(void* p) { return Cast(p)->NumberOfConstraints(); }
// In-place RLDLT of the lower triangle of A (n x n); transpositions written as ints. Returns 1
// when no pivot was regularised.
int ORACLE_LdltLower(int n, double* A, int* transpositions) {
  std::vector<int> t;
  const bool ok = oracle::LdltLower(n, A, n, &t);
  for (int i = 0; i < n; i++) transpositions[i] = t[i];
  return ok ? 1 : 0;
}
void ORACLE_SolveLdlt(int n, const double* LD, const int* transpositions, double* x) {
  oracle::SolveLdlt(n, LD, n, std::vector<int>(transpositions, transpositions + n), x);
}

int ORACLE_BlasAvailable() { return oracle::BlasAvailable() ? 1 : 0; }
void ORACLE_ForcePlainLoops(int on) { oracle::ForcePlainLoops(on != 0); }
void ORACLE_SetBlasThreads(int n) { oracle::SetBlasThreads(n); }
int ORACLE_GetBlasThreads() { return oracle::GetBlasThreads(); }

void ORACLE_SetGramVariant(void* p, int v) {
  Cast(p)->gram_variant = (v == 2) ? oracle::GramVariant::kSymmetric
                                   : (v ? oracle::GramVariant::kBlas3 : oracle::GramVariant::kAsWritten);
}
void ORACLE_FeasibleObjective(void* p, double* b) {
  const auto v = Cast(p)->FeasibleObjective();
  std::memcpy(b, v.data(), v.size() * sizeof(double));
}
void ORACLE_GetStatus(void* p, int* out4) {
  const auto& s = Cast(p)->status;
  out4[0] = s.solved;
  out4[1] = s.num_iterations;
  out4[2] = s.primal_infeasible;
  out4[3] = s.dual_infeasible;
}
// out8 = {inv_sqrt_mu, mu, d_2, d_inf, by, cx, kkt_error, step_size}; returns 0 if out of range.
int ORACLE_GetIterationLog(void* p, int iter, double* out8) {
  const auto& log = Cast(p)->log;
  if (iter < 0 || iter >= (int)log.size()) return 0;
  const auto& r = log[iter];
  const double v[8] = {r.inv_sqrt_mu, r.mu, r.d_2, r.d_inf, r.by, r.cx, r.kkt_error, r.step_size};
  std::memcpy(out8, v, sizeof(v));
  return 1;
}
// out5 = seconds in {assemble, factor, solve, update, mu}
void ORACLE_GetPhaseSeconds(void* p, double* out5) {
  const auto& s = Cast(p)->seconds;
  out5[0] = s.assemble;
  out5[1] = s.factor;
  out5[2] = s.solve;
  out5[3] = s.update;
  out5[4] = s.mu;
}
// Assembles the Newton system at the program's current iterate (after Initialize / a solve) and
// copies it out: H (m x m, lower triangle valid), AW, AQc (m each), scalars {<w,c>, <c,Qc>}.
void ORACLE_AssembleNewtonSystem(void* p, int coldstart, double* H, double* AW, double* AQc,
                                 double* scalars2) {
  Program* prog = Cast(p);
  oracle::SolverConfiguration cfg;
  cfg.initialization_mode = coldstart ? 0 : 1;
  prog->Initialize(cfg);
  prog->Assemble();
  const int m = prog->SizeOfKKTSystem();
  std::memcpy(H, prog->H.data(), sizeof(double) * m * m);
  std::memcpy(AW, prog->sys.AW, sizeof(double) * m);
  std::memcpy(AQc, prog->sys.AQc, sizeof(double) * m);
  scalars2[0] = prog->sys.inner_product_of_w_and_c;
  scalars2[1] = prog->sys.inner_product_of_c_and_Qc;
}

// Schur complement of one dense LMI block at scaling point W (n x n). variant: 0 as-written,
// 1 BLAS-3. G is m x m column-major, lower triangle written.
void ORACLE_SchurDenseLMI(int n, int m, const double* A, const double* C, const double* W,
                          int variant, double* G, double* AW, double* AQc, double* scalars2) {
  DenseLmiCone cone(n, m, A, C);
  std::vector<double> ws(cone.WorkspaceSize());
  cone.BindWorkspace(ws.data());
  std::memcpy(cone.W_.p, W, sizeof(double) * n * n);
  cone.gram_variant = variant ? oracle::GramVariant::kBlas3 : oracle::GramVariant::kAsWritten;
  oracle::SchurSystem sys;
  sys.m = m;
  std::vector<double> buf(sys.SizeOf(), 0.0);
  sys.Bind(buf.data());
  cone.ConstructSchurComplementSystem(true, &sys);
  std::memset(G, 0, sizeof(double) * m * m);
  for (int j = 0; j < m; j++)
    for (int i = j; i < m; i++) G[(size_t)j * m + i] = sys.G(i, j);
  std::memcpy(AW, sys.AW, sizeof(double) * m);
  std::memcpy(AQc, sys.AQc, sizeof(double) * m);
  scalars2[0] = sys.inner_product_of_w_and_c;
  scalars2[1] = sys.inner_product_of_c_and_Qc;
}

// minus_s = sum_i y_i A_i - k C
void ORACLE_NegativeSlack(int n, int m, const double* A, const double* C, const double* y, double k,
                          double* minus_s) {
  DenseLmiCone cone(n, m, A, C);
  cone.ComputeNegativeSlack(k, y, oracle::View(minus_s, n, n));
}

// PrepareStep (+ optional TakeStep) of the PSD cone from a given W. W is updated in place when
// take_step != 0 or affine != 0. out4 = {norminfd, normsqrd, step_size, 0}. WS_out (n x n, may be
// null) receives temp_1 after PrepareStep (= W * minus_s).
void ORACLE_PsdStep(int n, int m, const double* A, const double* C, double* W, const double* y,
                    double c_weight, double e_weight, int affine, int take_step, double* out4,
                    double* WS_out) {
  DenseLmiCone cone(n, m, A, C);
  std::vector<double> ws(cone.WorkspaceSize());
  cone.BindWorkspace(ws.data());
  std::memcpy(cone.W_.p, W, sizeof(double) * n * n);
  oracle::StepOptions opt;
  opt.affine = affine != 0;
  opt.c_weight = c_weight;
  opt.e_weight = e_weight;
  oracle::StepInfo info;
  cone.PrepareStep(opt, y, &info);
  if (WS_out) std::memcpy(WS_out, cone.temp_1_.p, sizeof(double) * n * n);
  opt.step_size = 1;
  if (!affine) {
    opt.step_size = std::min(1.0, 2.0 / (info.norminfd * info.norminfd));
    if (take_step) cone.TakeStep(opt);
  }
  out4[0] = info.norminfd;
  out4[1] = info.normsqrd;
  out4[2] = opt.step_size;
  out4[3] = 0;
  std::memcpy(W, cone.W_.p, sizeof(double) * n * n);
}

// out4 = {lambda_min, lambda_max, frobenius_norm_squared, trace}
void ORACLE_PsdWeightedSlackEigenvalues(int n, int m, const double* A, const double* C,
                                        const double* W, const double* y, double c_weight,
                                        double* out4) {
  DenseLmiCone cone(n, m, A, C);
  std::vector<double> ws(cone.WorkspaceSize());
  cone.BindWorkspace(ws.data());
  std::memcpy(cone.W_.p, W, sizeof(double) * n * n);
  oracle::SlackEigenvalues p;
  cone.GetWeightedSlackEigenvalues(y, c_weight, &p);
  out4[0] = p.lambda_min;
  out4[1] = p.lambda_max;
  out4[2] = p.frobenius_norm_squared;
  out4[3] = p.trace;
}

void ORACLE_PadeExpm(int n, const double* X, double* out) { oracle::ExponentialMapPade(n, X, out); }

// Returns the number of Ritz values written to `out` (<= num_iter).
int ORACLE_ApproximateEigenvalues(int n, const double* WS, const double* W, const double* r,
                                  int num_iter, double* out) {
  const auto e = oracle::ApproximateEigenvalues(n, WS, W, r, num_iter);
  std::memcpy(out, e.data(), e.size() * sizeof(double));
  return (int)e.size();
}
int ORACLE_SymmetricLanczos(int n, const double* A, const double* r0, int num_iter, double* out) {
  const auto e = oracle::SymmetricLanczos(n, A, r0, num_iter);
  std::memcpy(out, e.data(), e.size() * sizeof(double));
  return (int)e.size();
}
void ORACLE_SymmetricEigenvalues(int n, const double* A, double* out) {
  const auto e = oracle::SymmetricEigenvalues(n, A, n);
  std::memcpy(out, e.data(), e.size() * sizeof(double));
}
int ORACLE_TridiagonalEigenvalues(int n, const double* alpha, const double* beta, double* out) {
  const auto e = oracle::TridiagonalEigenvalues(std::vector<double>(alpha, alpha + n),
                                                std::vector<double>(beta, beta + (n > 0 ? n - 1 : 0)));
  std::memcpy(out, e.data(), e.size() * sizeof(double));
  return (int)e.size();
}
int ORACLE_CholeskyLower(int n, double* A) { return oracle::CholeskyLower(n, A, n) ? 1 : 0; }
void ORACLE_SolveLower(int n, const double* L, double* x, int transpose) {
  oracle::SolveLower(n, L, n, x, transpose != 0);
}
// Solves H x = rhs from the lower triangle of SPD H (factor + two triangular solves).
int ORACLE_SolveSpd(int n, const double* H, double* x) {
  std::vector<double> L(H, H + (size_t)n * n);
  if (!oracle::CholeskyLower(n, L.data(), n)) return 0;
  oracle::SolveLower(n, L.data(), n, x, false);
  oracle::SolveLower(n, L.data(), n, x, true);
  return 1;
}
int ORACLE_LuSolve(int n, const double* A, int nrhs, double* B) {
  std::vector<double> a(A, A + (size_t)n * n);
  return oracle::LuSolve(n, a.data(), n, nrhs, B, n) ? 1 : 0;
}
void ORACLE_Gemm(int ta, int tb, int M, int N, int K, double alpha, const double* A, int lda,
                 const double* B, int ldb, double beta, double* C, int ldc) {
  oracle::Gemm(ta != 0, tb != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
double ORACLE_DivergenceUpperBoundInverse(double bound, double frob, double trace, double lmin,
                                          double lmax, double rank) {
  oracle::SlackEigenvalues p;
  p.frobenius_norm_squared = frob;
  p.trace = trace;
  p.lambda_min = lmin;
  p.lambda_max = lmax;
  p.rank = rank;
  return oracle::DivergenceUpperBoundInverse(bound, p);
}
double ORACLE_DivergenceUpperBound(double k, double frob, double trace, double lmin, double lmax,
                                   double rank) {
  oracle::SlackEigenvalues p;
  p.frobenius_norm_squared = frob;
  p.trace = trace;
  p.lambda_min = lmin;
  p.lambda_max = lmax;
  p.rank = rank;
  return oracle::DivergenceUpperBound(k, p);
}

}  // extern "C"
