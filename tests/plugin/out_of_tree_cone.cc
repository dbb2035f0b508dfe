// An OUT-OF-TREE cone type against the public C++ plugin concept (include/conex_b200/constraint.h — the
// device-resident counterpart of the reference's conex/constraint.h:51-197): nothing here is known to the
// library. Compiled by plain g++ against include/ only and linked with lib/libconex_b200_host.a, the way a C++
// user of the reference writes a cone against conex/*.h and adds it with Program::AddConstraint
// (conex/cone_program.h:191-218).
//
// The cone is the nonnegative orthant, c - A y >= 0, written independently of the library's LinearConstraint:
// its arithmetic runs on the HOST (state downloaded / uploaded through the DeviceContext it is bound to), which is
// all a test of the boundary needs. main() solves the same LP with this cone and with the library's own
// LinearConstraint and prints both solutions; tests/test_plugin.py compares them (and both with the oracle).
#include <cmath>
#include <cstdio>
#include <vector>

#include "conex_b200/cone_program.h"

namespace user {

using conex::Ref;

struct OrthantWorkspace {
  explicit OrthantWorkspace(int n) : n_(n) {}
  friend size_t SizeOf(const OrthantWorkspace& o) { return 2 * static_cast<size_t>((o.n_ + 3) & ~3); }
  friend void Initialize(OrthantWorkspace* o, double* device_arena) {
    o->W = Ref(device_arena, o->n_, 1);
    o->direction = device_arena + ((o->n_ + 3) & ~3);
  }
  Ref W;                        // the scaling point: what CONEX_GetDualVariable copies out
  double* direction = nullptr;  // d of the last PrepareStep, kept in the arena like the reference does
  int n_;
};

class OrthantCone {
 public:
  OrthantCone(int n, int m, const double* A, const double* c) : n_(n), m_(m), A_(A, A + n * m), c_(c, c + n), ws_(n) {}
  OrthantWorkspace* workspace() { return &ws_; }
  int number_of_variables() const { return m_; }
  void bind(conex::DeviceContext* ctx) { ctx_ = ctx; }

  std::vector<double> Get(const double* device, int count) const {
    std::vector<double> h(count);
    ctx_->Download(h.data(), device, count);
    return h;
  }
  void Put(double* device, const std::vector<double>& h) const {
    ctx_->Upload(device, h.data(), h.size());
    ctx_->Synchronize();  // h is a temporary
  }
  std::vector<double> NegativeSlack(const Ref& y, double k) const {
    const std::vector<double> yh = Get(y.data, m_);
    std::vector<double> s(n_);
    for (int r = 0; r < n_; r++) {
      double v = 0;
      for (int j = 0; j < m_; j++) v += A_[j * n_ + r] * yh[j];
      s[r] = v - k * c_[r];
    }
    return s;
  }

  int n_, m_;
  std::vector<double> A_, c_;  // A column-major n x m
  OrthantWorkspace ws_;
  conex::DeviceContext* ctx_ = nullptr;
};

int Rank(const OrthantCone& o) { return o.n_; }

void SetIdentity(OrthantCone* o) { o->Put(o->ws_.W.data, std::vector<double>(o->n_, 1.0)); }

void ConstructSchurComplementSystem(OrthantCone* o, bool initialize, conex::SchurComplementSystem* sys) {
  const int n = o->n_, m = o->m_;
  const std::vector<double> w = o->Get(o->ws_.W.data, n);
  const long ld = sys->G.ld;
  // G is read back only when accumulating (initialize == false)
  std::vector<double> G(static_cast<size_t>(ld) * m, 0.0), AW(m, 0.0), AQc(m, 0.0), sc(2, 0.0);
  if (!initialize) {
    G = o->Get(sys->G.data, static_cast<int>(ld * m));
    AW = o->Get(sys->AW, m);
    AQc = o->Get(sys->AQc, m);
    sc = o->Get(sys->scalars, 2);
  }
  for (int j = 0; j < m; j++) {
    for (int i = j; i < m; i++) {
      double s = 0;
      for (int r = 0; r < n; r++) s += (w[r] * o->A_[i * n + r]) * (w[r] * o->A_[j * n + r]);
      G[j * ld + i] += s;
    }
    for (int r = 0; r < n; r++) {
      AW[j] += o->A_[j * n + r] * w[r];
      AQc[j] += (w[r] * o->A_[j * n + r]) * (w[r] * o->c_[r]);
    }
  }
  for (int r = 0; r < n; r++) {
    sc[0] += w[r] * o->c_[r];
    sc[1] += (w[r] * o->c_[r]) * (w[r] * o->c_[r]);
  }
  o->Put(sys->G.data, G);
  o->Put(sys->AW, AW);
  o->Put(sys->AQc, AQc);
  o->Put(sys->scalars, sc);
}

void GetWeightedSlackEigenvalues(OrthantCone* o, const Ref& y, double c_weight, conex::WeightedSlackEigenvalues* p) {
  const std::vector<double> w = o->Get(o->ws_.W.data, o->n_);
  const std::vector<double> s = o->NegativeSlack(y, c_weight);
  double mn = 1e300, mx = -1e300, sq = 0, tr = 0;
  for (int r = 0; r < o->n_; r++) {
    const double v = w[r] * s[r];
    mn = std::fmin(mn, v);
    mx = std::fmax(mx, v);
    sq += v * v;
    tr += v;
  }
  p->lambda_min = -mx;
  p->lambda_max = -mn;
  p->frobenius_norm_squared = sq;
  p->trace = -tr;
}

void PrepareStep(OrthantCone* o, const conex::StepOptions& opt, const Ref& y, conex::StepInfo* info) {
  std::vector<double> w = o->Get(o->ws_.W.data, o->n_);
  if (opt.affine) {  // dual recovery: W <- W + W (W s)
    const std::vector<double> s = o->NegativeSlack(y, 0.0);
    for (int r = 0; r < o->n_; r++) w[r] += w[r] * (s[r] * w[r]);
    o->Put(o->ws_.W.data, w);
    info->norminfd = info->normsqrd = 0;
    return;
  }
  const std::vector<double> s = o->NegativeSlack(y, opt.c_weight);
  std::vector<double> d(o->n_);
  double ninf = 0, nsq = 0;
  for (int r = 0; r < o->n_; r++) {
    d[r] = s[r] * w[r] + opt.e_weight;
    ninf = std::fmax(ninf, std::fabs(d[r]));
    nsq += d[r] * d[r];
  }
  o->Put(o->ws_.direction, d);
  info->norminfd = ninf;
  info->normsqrd = nsq;
}

bool TakeStep(OrthantCone* o, const conex::StepOptions& opt) {
  std::vector<double> w = o->Get(o->ws_.W.data, o->n_);
  const std::vector<double> d = o->Get(o->ws_.direction, o->n_);
  for (int r = 0; r < o->n_; r++) w[r] *= std::exp(opt.step_size * d[r]);
  o->Put(o->ws_.W.data, w);
  return true;
}

}  // namespace user

int main() {
  // the LP of the reference's Python tests (interfaces/python/test/run_tests.py:48-60) plus two random rows
  const int n = 5, m = 2;
  const double A[n * m] = {1, 4, 1, 0.3, -0.7, /* column 1 */ 3, 1, 1, -0.2, 0.5};
  const double c[n] = {1, 1, 1, 2, 1.5};
  const std::vector<double> b = {6, 5};
  std::vector<std::vector<double>> ys;
  for (int variant = 0; variant < 2; variant++) {
    conex::Program prog(m);
    bool failed;
    if (variant == 0) {
      failed = prog.AddConstraint(user::OrthantCone(n, m, A, c));
    } else {
      failed = prog.AddConstraint(conex::LinearConstraint(n, m, A, c));
    }
    if (failed) return 2;
    conex::SolverConfiguration config;
    config.prepare_dual_variables = 1;
    std::vector<double> y(m, 0.0);
    if (!conex::Solve(b, prog, config, y.data())) return 3;
    std::vector<double> x(n);
    prog.GetDualVariable(0, x.data());
    std::printf("%s iterations %d y %.15e %.15e x", variant == 0 ? "out_of_tree" : "library", prog.Status().num_iterations,
                y[0], y[1]);
    for (double v : x) std::printf(" %.15e", v);
    std::printf("\n");
    ys.push_back(y);
  }
  return 0;
}
