// ORACLE — TEST INFRASTRUCTURE ONLY. See cones.h for scope and parity status.
#include "cones.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace oracle {

void SchurSystem::SetZero() {
  // conex/newton_step.h:83-91
  if (!residual_only) std::fill(G.p, G.p + G.size(), 0.0);
  std::fill(AW, AW + m, 0.0);
  std::fill(AQc, AQc + m, 0.0);
  inner_product_of_w_and_c = 0;
  inner_product_of_c_and_Qc = 0;
}

namespace {

double Trace(const View& X) {
  double t = 0;
  for (int i = 0; i < X.rows; i++) t += X(i, i);
  return t;
}

// conex/dense_lmi_constraint.cc:54-60 — column-by-column dot products.
double TraceInnerProduct(int n, const double* X, const double* Y) {
  double val = 0;
  for (int i = 0; i < n; i++) val += Dot(n, X + (size_t)i * n, Y + (size_t)i * n);
  return val;
}

void Store(bool initialize, double* dst, double v) {
  if (initialize) {
    *dst = v;
  } else {
    *dst += v;
  }
}

}  // namespace

// ============================================================================================
// Dense math kernels
// ============================================================================================

void ExponentialMapPade(int n, const double* X, double* result) {
  // conex/exponential_map_pade.cc:10-21: b = {120, 60, 12, 1}
  const size_t nn = (size_t)n * n;
  std::vector<double> even(nn), odd(nn), tmp(nn);
  Gemm(false, false, n, n, n, 1.0, X, n, X, n, 0.0, even.data(), n);  // X^2
  tmp = even;
  for (int i = 0; i < n; i++) tmp[(size_t)i * n + i] += 60.0;
  Gemm(false, false, n, n, n, 1.0, X, n, tmp.data(), n, 0.0, odd.data(), n);  // X (X^2 + 60 I)
  for (size_t k = 0; k < nn; k++) even[k] *= 12.0;
  for (int i = 0; i < n; i++) even[(size_t)i * n + i] += 120.0;
  // conex/exponential_map_pade.cc:23-32
  std::vector<double> denom(nn);
  for (size_t k = 0; k < nn; k++) {
    result[k] = odd[k] + even[k];  // numerator, overwritten by the solve
    denom[k] = -odd[k] + even[k];
  }
  if (!LuSolve(n, denom.data(), n, n, result, n)) {
    throw std::runtime_error("oracle: singular Pade denominator");
  }
}

std::vector<double> AsymmetricLanczos(int n, const double* WS, const double* W, const double* r,
                                      int num_iter) {
  // conex/approximate_eigenvalues.cc:178-239. State V = [v0 v1], U = [u0 u1];
  // <V, U> := V.col(0) . U.col(1)   (:173-176).
  std::vector<double> v0(n), v1(n), u0(n), u1(n), p0(n), p1(n);
  std::vector<double> alpha(std::max(num_iter, 1)), beta(std::max(num_iter - 1, 0));
  for (int i = 0; i < n; i++) v1[i] = r[i];
  Gemv(false, n, n, 1.0, W, n, r, 0.0, v0.data());
  const double scale = std::sqrt(Dot(n, v0.data(), v1.data()));
  for (int i = 0; i < n; i++) {
    v0[i] /= scale;
    v1[i] /= scale;
  }
  Gemv(false, n, n, 1.0, WS, n, v0.data(), 0.0, u0.data());
  Gemv(true, n, n, 1.0, WS, n, v1.data(), 0.0, u1.data());
  alpha[0] = Dot(n, v0.data(), u1.data());
  for (int i = 0; i < n; i++) {
    u0[i] -= alpha[0] * v0[i];
    u1[i] -= alpha[0] * v1[i];
  }
  int cnt = 0;
  for (int j = 1; j < num_iter; j++) {
    double b = Dot(n, u0.data(), u1.data());
    if (b < 1e-6) break;
    b = std::sqrt(b);
    beta[j - 1] = b;
    p0 = v0;
    p1 = v1;
    for (int i = 0; i < n; i++) {
      v0[i] = u0[i] / b;
      v1[i] = u1[i] / b;
    }
    Gemv(false, n, n, 1.0, WS, n, v0.data(), 0.0, u0.data());
    Gemv(true, n, n, 1.0, WS, n, v1.data(), 0.0, u1.data());
    alpha[j] = Dot(n, v0.data(), u1.data());
    for (int i = 0; i < n; i++) {
      u0[i] = u0[i] - alpha[j] * v0[i] - b * p0[i];
      u1[i] = u1[i] - alpha[j] * v1[i] - b * p1[i];
    }
    cnt++;
  }
  alpha.resize(cnt + 1);
  beta.resize(cnt);
  return TridiagonalEigenvalues(alpha, beta);
}

std::vector<double> ApproximateEigenvalues(int n, const double* WS, const double* W,
                                           const double* r, int num_iter) {
  // conex/approximate_eigenvalues.cc:241-256 (compressed branch)
  if (n == 1) return {WS[0]};
  return AsymmetricLanczos(n, WS, W, r, num_iter);
}

std::vector<double> SymmetricLanczos(int n, const double* A, const double* r0, int num_iter) {
  // conex/approximate_eigenvalues.cc:147-171
  std::vector<double> alpha(num_iter), beta(std::max(num_iter - 1, 0));
  std::vector<double> v(n), vprev(n), w(n), Av(n);
  const double nr = std::sqrt(Dot(n, r0, r0));
  for (int i = 0; i < n; i++) v[i] = r0[i] / nr;
  Gemv(false, n, n, 1.0, A, n, v.data(), 0.0, Av.data());
  alpha[0] = Dot(n, v.data(), Av.data());
  for (int i = 0; i < n; i++) w[i] = Av[i] - alpha[0] * v[i];
  for (int j = 1; j < num_iter; j++) {
    beta[j - 1] = std::sqrt(Dot(n, w.data(), w.data()));
    vprev = v;
    for (int i = 0; i < n; i++) v[i] = w[i] / beta[j - 1];
    Gemv(false, n, n, 1.0, A, n, v.data(), 0.0, Av.data());
    alpha[j] = Dot(n, v.data(), Av.data());
    for (int i = 0; i < n; i++) w[i] = Av[i] - alpha[j] * v[i] - beta[j - 1] * vprev[i];
  }
  return TridiagonalEigenvalues(alpha, beta);
}

// ============================================================================================
// mu selection: closed-form inverse of the divergence upper bound (conex/divergence.cc)
// ============================================================================================
namespace {

// Largest root x of (a x^2 + b x + c)/(2 - d x) = k   (conex/divergence.cc:11-23).
double RationalRoot(double a, double b, double c, double d, double k) {
  const double disc = b * b - 4 * a * c + 8 * a * k + 2 * b * d * k + std::pow(d * k, 2);
  return -(b + d * k - std::sqrt(disc)) / (2 * a);
}

// conex/divergence.cc:26-41
double LambdaMaxBranch(double bound, const SlackEigenvalues& p) {
  const double x = RationalRoot(p.frobenius_norm_squared, -2 * p.trace, p.rank, p.lambda_max, bound);
  const double lower = 2.0 / (p.lambda_max + p.lambda_min);
  return (x >= lower) ? x : -1;
}

// conex/divergence.cc:47-85: roots of  a k - b + n/k = c  kept if inside [0, 2/(lmax+lmin)].
double LambdaMinBranch(double bound, const SlackEigenvalues& p) {
  const double a = p.frobenius_norm_squared / p.lambda_min;
  const double b = 2 * p.trace / p.lambda_min;
  const double n = p.rank / p.lambda_min;
  const double c = bound;
  const double disc = b * b + 2 * b * c + c * c - 4 * a * n;
  const double r1 = (b + c + std::sqrt(disc)) / (2 * a);
  const double r2 = (b + c - std::sqrt(disc)) / (2 * a);
  const double upper = 2.0 / (p.lambda_max + p.lambda_min);
  double k = -1;
  if (!(disc < 0)) {
    if (r1 >= 0 && r1 <= upper) k = r1;
    if (r2 >= 0 && r2 <= upper && r2 > k) k = r2;
  }
  return k;
}

// conex/divergence.cc:85-94
bool BoundIsFinite(double k, const SlackEigenvalues& p) {
  double norm_inf = std::fabs(k * p.lambda_max - 1);
  norm_inf = std::max(norm_inf, std::fabs(k * p.lambda_min - 1));
  return norm_inf < 1;
}

}  // namespace

double DivergenceUpperBoundInverse(double bound, const SlackEigenvalues& p) {
  // conex/divergence.cc:96-111
  double k = -1;
  const double k1 = LambdaMinBranch(bound, p);
  const double k2 = LambdaMaxBranch(bound, p);
  if (BoundIsFinite(k1, p)) k = k1;
  if (k2 > k && BoundIsFinite(k2, p)) k = k2;
  return k;
}

double DivergenceUpperBound(double k, const SlackEigenvalues& p) {
  // conex/divergence.cc:113-121
  const double numerator = k * k * p.frobenius_norm_squared - 2 * k * p.trace + p.rank;
  double norm_inf = std::fabs(k * p.lambda_max - 1);
  norm_inf = std::max(norm_inf, std::fabs(k * p.lambda_min - 1));
  return numerator / (1 - norm_inf);
}

// ============================================================================================
// Dense LMI / PSD cone
// ============================================================================================

DenseLmiCone::DenseLmiCone(int n, int m, const double* A, const double* C)
    : n_(n), m_(m), Avect_(A, A + (size_t)n * n * m), C_(C, C + (size_t)n * n) {}

void DenseLmiCone::BindWorkspace(double* data) {
  // conex/psd_constraint.h:19-25
  const int stride = AlignedSize(n_ * n_);
  W_ = View(data, n_, n_);
  temp_1_ = View(data + stride, n_, n_);
  temp_2_ = View(data + 2 * stride, n_, n_);
}

void DenseLmiCone::SetIdentity() {
  // conex/psd_constraint.cc:92-95
  std::fill(W_.p, W_.p + W_.size(), 0.0);
  for (int i = 0; i < n_; i++) W_(i, i) = 1;
}

void DenseLmiCone::ComputeNegativeSlack(double k, const double* y, View s) const {
  // conex/dense_lmi_constraint.cc:8-27:  s = sum_i y_i A_i - k C   (column-by-column axpy order)
  const int nn = n_ * n_;
  Gemv(false, nn, m_, 1.0, Avect_.data(), nn, y, 0.0, s.p);
  for (int k2 = 0; k2 < nn; k2++) s.p[k2] -= k * C_[k2];
}

void DenseLmiCone::ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) {
  // conex/dense_lmi_constraint.cc:62-103
  const int n = n_, m = m_, nn = n * n;
  View AW = temp_1_, WAW = temp_2_;
  if (gram_variant == GramVariant::kAsWritten) {
    std::vector<double> row(m);
    for (int i = 0; i < m; i++) {
      const double* Ai = Avect_.data() + (size_t)i * nn;
      Gemm(false, false, n, n, n, 1.0, Ai, n, W_.p, n, 0.0, AW.p, n);     // :31
      Gemm(false, false, n, n, n, 1.0, W_.p, n, AW.p, n, 0.0, WAW.p, n);  // :32
      Gemv(true, nn, i + 1, 1.0, Avect_.data(), nn, WAW.p, 0.0, row.data());  // :77-78
      for (int j = 0; j <= i; j++) Store(initialize, &sys->G(i, j), row[j]);
      Store(initialize, &sys->AW[i], Trace(AW));                             // :79
      Store(initialize, &sys->AQc[i], TraceInnerProduct(n, C_.data(), WAW.p));  // :80
    }
  } else if (gram_variant == GramVariant::kSymmetric) {
    std::vector<double> L(W_.p, W_.p + nn), S((size_t)nn * (m + 1)), T(nn);
    CholeskyLower(n, L.data(), n);
    for (int j = 0; j < n; j++)
      for (int i = 0; i < j; i++) L[(size_t)j * n + i] = 0.0;
    for (int i = 0; i <= m; i++) {
      const double* Ai = (i < m) ? Avect_.data() + (size_t)i * nn : C_.data();
      Gemm(false, false, n, n, n, 1.0, Ai, n, L.data(), n, 0.0, T.data(), n);
      Gemm(true, false, n, n, n, 1.0, L.data(), n, T.data(), n, 0.0, S.data() + (size_t)i * nn, n);
    }
    std::vector<double> out((size_t)(m + 1) * (m + 1));
    Gemm(true, false, m + 1, m + 1, nn, 1.0, S.data(), nn, S.data(), nn, 0.0, out.data(), m + 1);
    for (int i = 0; i < m; i++) {
      for (int j = 0; j <= i; j++) Store(initialize, &sys->G(i, j), out[(size_t)j * (m + 1) + i]);
      double tr = 0;
      for (int r = 0; r < n; r++) tr += S[(size_t)i * nn + (size_t)r * n + r];
      Store(initialize, &sys->AW[i], tr);
      Store(initialize, &sys->AQc[i], out[(size_t)i * (m + 1) + m]);
    }
    double trc = 0;
    for (int r = 0; r < n; r++) trc += S[(size_t)m * nn + (size_t)r * n + r];
    Store(initialize, &sys->inner_product_of_w_and_c, trc);
    Store(initialize, &sys->inner_product_of_c_and_Qc, out[(size_t)m * (m + 1) + m]);
    return;
  } else {
    const int kPanel = 64;
    std::vector<double> B((size_t)nn * kPanel), out((size_t)kPanel * m);
    for (int i0 = 0; i0 < m; i0 += kPanel) {
      const int pb = std::min(kPanel, m - i0);
      for (int i = i0; i < i0 + pb; i++) {
        const double* Ai = Avect_.data() + (size_t)i * nn;
        double* Bi = B.data() + (size_t)(i - i0) * nn;
        Gemm(false, false, n, n, n, 1.0, Ai, n, W_.p, n, 0.0, AW.p, n);
        Gemm(false, false, n, n, n, 1.0, W_.p, n, AW.p, n, 0.0, Bi, n);
        Store(initialize, &sys->AW[i], Trace(AW));
        Store(initialize, &sys->AQc[i], TraceInnerProduct(n, C_.data(), Bi));
      }
      const int ncols = i0 + pb;
      Gemm(true, false, pb, ncols, nn, 1.0, B.data(), nn, Avect_.data(), nn, 0.0, out.data(), pb);
      for (int i = i0; i < i0 + pb; i++)
        for (int j = 0; j <= i; j++) Store(initialize, &sys->G(i, j), out[(size_t)j * pb + (i - i0)]);
    }
  }
  Store(initialize, &sys->inner_product_of_w_and_c, TraceInnerProduct(n, C_.data(), W_.p));  // :82
  View CW = AW, WCW = WAW;                                                                   // :84-86
  Gemm(false, false, n, n, n, 1.0, C_.data(), n, W_.p, n, 0.0, CW.p, n);
  Gemm(false, false, n, n, n, 1.0, W_.p, n, CW.p, n, 0.0, WCW.p, n);
  Store(initialize, &sys->inner_product_of_c_and_Qc, TraceInnerProduct(n, C_.data(), WCW.p));  // :87-88
}

void DenseLmiCone::GeodesicUpdate(double scale, const StepOptions& opt, View WS) {
  // conex/psd_constraint.cc:13-28
  const int n = n_;
  View expWS = temp_2_;
  for (int i = 0; i < n; i++) WS(i, i) += opt.e_weight;
  if (scale != 1.0) {
    for (size_t k = 0; k < WS.size(); k++) WS.p[k] *= scale;
  }
  ExponentialMapPade(n, WS.p, expWS.p);
  std::vector<double> prod(WS.size());
  Gemm(false, false, n, n, n, 1.0, expWS.p, n, W_.p, n, 0.0, prod.data(), n);  // W = expWS * W
  std::copy(prod.begin(), prod.end(), W_.p);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) WS(i, j) = W_(j, i);  // WS = W^T
  for (size_t k = 0; k < WS.size(); k++) W_.p[k] = (W_.p[k] + WS.p[k]) * 0.5;
}

void DenseLmiCone::AffineUpdate(double w_e, View WS) {
  // conex/psd_constraint.cc:33-43
  const int n = n_;
  View WSW = temp_2_;
  Gemm(false, false, n, n, n, 1.0, WS.p, n, W_.p, n, 0.0, WSW.p, n);
  if (w_e != 0) {
    for (size_t k = 0; k < W_.size(); k++) W_.p[k] *= (1 + w_e);
  }
  for (size_t k = 0; k < W_.size(); k++) W_.p[k] += WSW.p[k];
}

void DenseLmiCone::PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) {
  // conex/psd_constraint.cc:45-84. minus_s and WS alias temp_1 (:48-50).
  const int n = n_;
  View minus_s = temp_1_, WS = temp_1_, WSWS = temp_2_;
  ComputeNegativeSlack(opt.c_weight, y, minus_s);
  {
    // WS = W * minus_s with aliasing: Eigen evaluates into a temporary, then assigns.
    Gemm(false, false, n, n, n, 1.0, W_.p, n, minus_s.p, n, 0.0, temp_2_.p, n);
    std::copy(temp_2_.p, temp_2_.p + temp_2_.size(), WS.p);
  }
  if (opt.affine) {
    AffineUpdate(opt.e_weight, WS);
    return;
  }
  int index = 0;
  for (int i = 1; i < n; i++)
    if (WS(i, i) > WS(index, index)) index = i;
  // r = minus_s.col(index), which after the aliasing above is column `index` of WS (:66-69).
  std::vector<double> r(WS.col(index), WS.col(index) + n);
  const std::vector<double> eig = ApproximateEigenvalues(n, WS.p, W_.p, r.data(), n / 2);
  const double lambda_1 = std::fabs(opt.e_weight + eig.front());
  const double lambda_2 = std::fabs(opt.e_weight + eig.back());
  const double norminf = std::max(lambda_1, lambda_2);
  Gemm(false, false, n, n, n, 1.0, WS.p, n, WS.p, n, 0.0, WSWS.p, n);
  info->norminfd = norminf;
  info->normsqrd = Trace(WSWS) + 2 * Trace(WS) + n;
}

bool DenseLmiCone::TakeStep(const StepOptions& opt) {
  // conex/psd_constraint.cc:86-90
  GeodesicUpdate(opt.step_size, opt, temp_1_);
  return true;
}

void DenseLmiCone::GetWeightedSlackEigenvalues(const double* y, double c_weight,
                                               SlackEigenvalues* p) {
  // conex/psd_constraint.cc:97-128: minus_s = temp_1, WS = temp_2, WSWS = temp_1.
  const int n = n_;
  View minus_s = temp_1_, WSWS = temp_1_, WS = temp_2_;
  ComputeNegativeSlack(c_weight, y, minus_s);
  Gemm(false, false, n, n, n, 1.0, W_.p, n, minus_s.p, n, 0.0, WS.p, n);
  int index = 0;
  for (int i = 1; i < n; i++)
    if (WS(i, i) > WS(index, index)) index = i;
  std::vector<double> r(minus_s.col(index), minus_s.col(index) + n);
  const std::vector<double> eig = ApproximateEigenvalues(n, WS.p, W_.p, r.data(), n / 2);
  p->lambda_max = -eig.front();
  p->lambda_min = -eig.back();
  std::vector<double> prod(WS.size());
  Gemm(false, false, n, n, n, 1.0, WS.p, n, WS.p, n, 0.0, prod.data(), n);
  std::copy(prod.begin(), prod.end(), WSWS.p);
  p->frobenius_norm_squared = Trace(WSWS);
  p->trace = -Trace(WS);
}

// ============================================================================================
// LP cone
// ============================================================================================

LinearCone::LinearCone(int n, int m, const double* A, const double* c)
    : n_(n), m_(m), A_(A, A + (size_t)n * m), c_(c, c + n) {}

void LinearCone::BindWorkspace(double* data) {
  // conex/linear_workspace.h:20-28
  const int s = AlignedSize(n_);
  W_ = data;
  temp_1_ = data + s;
  temp_2_ = data + 2 * s;
  WA_ = data + 3 * s;
}

void LinearCone::SetIdentity() { std::fill(W_, W_ + n_, 1.0); }  // linear_constraint.cc:105

void LinearCone::ComputeNegativeSlack(double k, const double* y, double* minus_s) const {
  // conex/linear_constraint.cc:168-172
  Gemv(false, n_, m_, 1.0, A_.data(), n_, y, 0.0, minus_s);
  for (int i = 0; i < n_; i++) minus_s[i] -= c_[i] * k;
}

void LinearCone::ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) {
  // conex/linear_constraint.cc:177-205
  const int n = n_, m = m_;
  double* WC = temp_1_;
  for (int j = 0; j < m; j++)
    for (int i = 0; i < n; i++) WA_[(size_t)j * n + i] = W_[i] * A_[(size_t)j * n + i];
  double wc_sum = 0, wc_sq = 0;
  for (int i = 0; i < n; i++) {
    WC[i] = W_[i] * c_[i];
    wc_sum += WC[i];
    wc_sq += WC[i] * WC[i];
  }
  Store(initialize, &sys->inner_product_of_w_and_c, wc_sum);
  Store(initialize, &sys->inner_product_of_c_and_Qc, wc_sq);
  const double beta = initialize ? 0.0 : 1.0;
  Gemm(true, false, m, m, n, 1.0, WA_, n, WA_, n, beta, sys->G.p, sys->G.rows);
  Gemv(true, n, m, 1.0, A_.data(), n, W_, beta, sys->AW);
  Gemv(true, n, m, 1.0, WA_, n, WC, beta, sys->AQc);
}

void LinearCone::PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) {
  // conex/linear_constraint.cc:108-129
  if (!opt.affine) {
    double* d = temp_2_;
    ComputeNegativeSlack(opt.c_weight, y, d);
    double norminf = 0, normsq = 0;
    for (int i = 0; i < n_; i++) {
      d[i] = d[i] * W_[i] + opt.e_weight;
      norminf = std::max(norminf, std::fabs(d[i]));
      normsq += d[i] * d[i];
    }
    info->norminfd = norminf;
    info->normsqrd = normsq;
  } else {
    ComputeNegativeSlack(0, y, temp_1_);
    TakeStep(opt);
  }
}

bool LinearCone::TakeStep(const StepOptions& opt) {
  // conex/linear_constraint.cc:131-145 and :174-179 (affine)
  if (!opt.affine) {
    double* d = temp_2_;
    for (int i = 0; i < n_; i++) {
      double di = d[i];
      if (opt.step_size != 1) di *= opt.step_size;
      d[i] = std::exp(di);
      W_[i] *= d[i];
    }
  } else {
    double* minus_s = temp_1_;
    for (int i = 0; i < n_; i++) {
      const double sw = minus_s[i] * W_[i];
      minus_s[i] = sw;
      W_[i] += W_[i] * sw;
    }
  }
  return true;
}

void LinearCone::GetWeightedSlackEigenvalues(const double* y, double c_weight,
                                             SlackEigenvalues* p) {
  // conex/linear_constraint.cc:148-166
  double* minus_s = temp_1_;
  double* Ws = temp_2_;
  ComputeNegativeSlack(c_weight, y, minus_s);
  double mn = 0, mx = 0, sq = 0, sum = 0;
  for (int i = 0; i < n_; i++) {
    Ws[i] = W_[i] * minus_s[i];
    if (i == 0 || Ws[i] < mn) mn = Ws[i];
    if (i == 0 || Ws[i] > mx) mx = Ws[i];
    sq += Ws[i] * Ws[i];
    sum += Ws[i];
  }
  p->lambda_max = -mn;
  p->lambda_min = -mx;
  p->frobenius_norm_squared = sq;
  p->trace = -sum;
}

// ============================================================================================
// Second-order cone (spin factor algebra)
// ============================================================================================
namespace {

// SpectralDecompSpinFactor::Compute + Idempotents (soc_constraint.cc:22-63), applied to a scalar
// function f of the two eigenvalues: returns f(ev0) c0 + f(ev1) c1.
template <typename F>
void SpinFunction(int n, double x0, const double* x1, F f, double* out) {
  double nq = 0;
  for (int i = 0; i < n; i++) nq += x1[i] * x1[i];
  nq = std::sqrt(nq);
  const double f0 = f(x0 + nq), f1 = f(x0 - nq);
  out[0] = f0 * .5 + f1 * .5;
  for (int i = 0; i < n; i++) {
    const double q = (nq > 0) ? x1[i] / nq : 0.0;
    out[1 + i] = f0 * (.5 * q) + f1 * (-.5 * q);
  }
}

// Q(x) y = 2 <x,y> x - det(x) R y, R = diag(1,-1,...,-1) (soc_constraint.cc:115-128).
void QuadraticRepresentation(int order, const double* x, const double* y, double* out) {
  double det = x[0] * x[0], xy = 0;
  for (int i = 1; i < order; i++) det -= x[i] * x[i];
  for (int i = 0; i < order; i++) xy += x[i] * y[i];
  for (int i = 0; i < order; i++) {
    double z = det * y[i];
    if (i == 0) z *= -1;
    out[i] = (2 * xy) * x[i] + z;
  }
}

}  // namespace

SocCone::SocCone(int n, int m, const double* A, const double* c)
    : n_(n), m_(m), A_((size_t)(n + 1) * m, 0.0), c_(n + 1, 0.0), dual_(n + 1, 0.0) {
  if (A) A_.assign(A, A + (size_t)(n + 1) * m);
  if (c) c_.assign(c, c + n + 1);
}

void SocCone::BindWorkspace(double* data) {
  // conex/workspace_soc.h:13-21
  const int s = AlignedSize(n_);
  W1_ = data;
  temp1_ = data + s;
  W0_ = data + 4 * s;
}

void SocCone::SetIdentity() {
  *W0_ = 1;
  std::fill(W1_, W1_ + n_, 0.0);
}

void SocCone::ComputeNegativeSlack(double k, const double* y, double* minus_s) const {
  Gemv(false, n_ + 1, m_, 1.0, A_.data(), n_ + 1, y, 0.0, minus_s);
  for (int i = 0; i <= n_; i++) minus_s[i] -= c_[i] * k;
}

void SocCone::GetWeightedSlackEigenvalues(const double* y, double c_weight, SlackEigenvalues* p) {
  // conex/soc_constraint.cc:172-189
  const int n = n_;
  std::vector<double> minus_s(n + 1), wsqrt(n + 1), Ws(n + 1);
  ComputeNegativeSlack(c_weight, y, minus_s.data());
  SpinFunction(n, *W0_, W1_, [](double e) { return std::sqrt(e); }, wsqrt.data());
  QuadraticRepresentation(n + 1, wsqrt.data(), minus_s.data(), Ws.data());
  double nq = 0;
  for (int i = 1; i <= n; i++) nq += Ws[i] * Ws[i];
  nq = std::sqrt(nq);
  const double ev0 = Ws[0] + nq, ev1 = Ws[0] - nq;
  const double lmax = -std::min(ev0, ev1), lmin = -std::max(ev0, ev1);
  p->lambda_max = lmax;
  p->lambda_min = lmin;
  p->frobenius_norm_squared = lmax * lmax + lmin * lmin;
  p->trace = lmax + lmin;
}

void SocCone::PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) {
  // conex/soc_constraint.cc:213-230. NB: W is overwritten by its square root here and the
  // e_weight / affine options are not consulted.
  const int n = n_;
  std::vector<double> minus_s(n + 1), wsqrt(n + 1), d(n + 1);
  ComputeNegativeSlack(opt.c_weight, y, minus_s.data());
  SpinFunction(n, *W0_, W1_, [](double e) { return std::sqrt(e); }, wsqrt.data());
  *W0_ = wsqrt[0];
  std::copy(wsqrt.begin() + 1, wsqrt.end(), W1_);
  QuadraticRepresentation(n + 1, wsqrt.data(), minus_s.data(), d.data());
  d[0] += 1;
  std::copy(d.begin() + 1, d.end(), temp1_);
  d0_ = d[0];
  double nq = 0, sq = 0;
  for (int i = 1; i <= n; i++) nq += d[i] * d[i];
  for (int i = 0; i <= n; i++) sq += d[i] * d[i];
  nq = std::sqrt(nq);
  info->norminfd = std::max(std::fabs(d[0] + nq), std::fabs(d[0] - nq));
  info->normsqrd = 2 * sq;
}

bool SocCone::TakeStep(const StepOptions& opt) {
  // conex/soc_constraint.cc:191-211
  const int n = n_;
  std::vector<double> wsqrt(n + 1), d1(temp1_, temp1_ + n), expd(n + 1), wn(n + 1);
  double d0 = d0_;
  wsqrt[0] = *W0_;
  std::copy(W1_, W1_ + n, wsqrt.begin() + 1);
  if (opt.step_size != 1.0) {
    d0 *= opt.step_size;
    for (auto& v : d1) v *= opt.step_size;
  }
  SpinFunction(n, d0, d1.data(), [](double e) { return std::exp(e); }, expd.data());
  QuadraticRepresentation(n + 1, wsqrt.data(), expd.data(), wn.data());
  *W0_ = wn[0];
  std::copy(wn.begin() + 1, wn.end(), W1_);
  return true;
}

void SocCone::ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) {
  // conex/soc_constraint.cc:232-262
  const int n = n_, m = m_, o = n + 1;
  std::vector<double> wsqrt(o), W(o), WA((size_t)o * m), WC(o);
  SpinFunction(n, *W0_, W1_, [](double e) { return std::sqrt(e); }, wsqrt.data());
  W[0] = *W0_;
  std::copy(W1_, W1_ + n, W.begin() + 1);
  QuadraticRepresentation(o, wsqrt.data(), c_.data(), WC.data());
  for (int i = 0; i < m; i++)
    QuadraticRepresentation(o, wsqrt.data(), &A_[(size_t)i * o], &WA[(size_t)i * o]);
  const double beta = initialize ? 0.0 : 1.0;
  Gemm(true, false, m, m, o, 2.0, WA.data(), o, WA.data(), o, beta, sys->G.p, sys->G.rows);
  Gemv(true, o, m, 2.0, A_.data(), o, W.data(), beta, sys->AW);
  Gemv(true, o, m, 2.0, WA.data(), o, WC.data(), beta, sys->AQc);
  double sq = 0;
  for (int i = 0; i < o; i++) sq += WC[i] * WC[i];
  Store(initialize, &sys->inner_product_of_w_and_c, 2 * WC[0]);
  Store(initialize, &sys->inner_product_of_c_and_Qc, 2 * sq);
}

// ============================================================================================
// Equality constraints
// ============================================================================================

void EqualityCone::ConstructSchurComplementSystem(bool initialize, SchurSystem* sys) {
  // conex/equality_constraint.cc:13-28: G = [0 A^T; A 0] on (variables, multipliers), AQc = [0; b].
  const int t = nv_ + rows_;
  if (initialize) sys->SetZero();
  for (int j = 0; j < nv_; j++) {
    for (int i = 0; i < rows_; i++) {
      sys->G(nv_ + i, j) += A_[(size_t)j * rows_ + i];
      sys->G(j, nv_ + i) += A_[(size_t)j * rows_ + i];
    }
  }
  for (int i = 0; i < rows_; i++) sys->AQc[nv_ + i] += b_[i];
  (void)t;
}

void EqualityCone::PrepareStep(const StepOptions&, const double* y, StepInfo* info) {
  // conex/equality_constraint.cc:30-35
  for (int i = 0; i < rows_; i++) lambda_[i] = y[nv_ + i];
  info->normsqrd = 0;
  info->norminfd = 0;
}

// ============================================================================================
// Hermitian (real) LMI built incrementally
// ============================================================================================

std::vector<double> HermitianLanczos(int n, const double* WS, const double* W, const double* r,
                                     int num_iter) {
  // conex/jordan_matrix_algebra.cc:387-452 for d = 1
  std::vector<double> v0(n), v1(n), u0(n), u1(n), p0(n), p1(n);
  std::vector<double> alpha(std::max(num_iter, 1)), beta(std::max(num_iter - 1, 0));
  for (int i = 0; i < n; i++) v1[i] = r[i];
  Gemv(false, n, n, 1.0, W, n, r, 0.0, v0.data());
  const double nrm = std::sqrt(Dot(n, v0.data(), v1.data()));
  for (int i = 0; i < n; i++) {
    v0[i] *= 1.0 / nrm;
    v1[i] *= 1.0 / nrm;
  }
  Gemv(false, n, n, 1.0, WS, n, v0.data(), 0.0, u0.data());
  Gemv(true, n, n, 1.0, WS, n, v1.data(), 0.0, u1.data());
  const double scaling = Dot(n, u0.data(), u1.data());
  alpha[0] = Dot(n, v0.data(), u1.data());
  for (int i = 0; i < n; i++) {
    u0[i] -= alpha[0] * v0[i];
    u1[i] -= alpha[0] * v1[i];
  }
  int cnt = 0;
  for (int j = 1; j < num_iter; j++) {
    double b = Dot(n, u0.data(), u1.data());
    if (b < 1e-5 * scaling) break;
    b = std::sqrt(b);
    beta[j - 1] = b;
    p0 = v0;
    p1 = v1;
    for (int i = 0; i < n; i++) {
      v0[i] = u0[i] * (1.0 / b);
      v1[i] = u1[i] * (1.0 / b);
    }
    Gemv(false, n, n, 1.0, WS, n, v0.data(), 0.0, u0.data());
    Gemv(true, n, n, 1.0, WS, n, v1.data(), 0.0, u1.data());
    alpha[j] = Dot(n, v0.data(), u1.data());
    for (int i = 0; i < n; i++) {
      u0[i] = u0[i] - alpha[j] * v0[i] - b * p0[i];
      u1[i] = u1[i] - alpha[j] * v1[i] - b * p1[i];
    }
    cnt++;
  }
  alpha.resize(cnt + 1);
  beta.resize(cnt);
  return TridiagonalEigenvalues(alpha, beta);
}

void ExponentialMapTaylor(int n, const double* X, double* result) {
  // conex/exponential_map.cc:15-42 with squarings = 2, degree = 2
  const size_t nn = (size_t)n * n;
  std::vector<double> xpow(nn), y(nn), t(nn);
  for (size_t k = 0; k < nn; k++) xpow[k] = X[k] * 1.0 / 4.0;
  y = xpow;
  for (int i = 0; i < n; i++) y[(size_t)i * n + i] += 1;
  Gemm(false, false, n, n, n, 1.0, X, n, xpow.data(), n, 0.0, t.data(), n);
  for (size_t k = 0; k < nn; k++) {
    xpow[k] = t[k] * (1.0 / 2 * 1.0 / 4.0);
    y[k] += xpow[k];
  }
  Gemm(false, false, n, n, n, 1.0, y.data(), n, y.data(), n, 0.0, xpow.data(), n);
  Gemm(false, false, n, n, n, 1.0, xpow.data(), n, xpow.data(), n, 0.0, result, n);
}

namespace {
// Eigen::MatrixXd::Random(n, 1): n draws of -1 + 2 rand() / RAND_MAX (Eigen 3.3 random<double>()).
std::vector<double> EigenRandomVector(int n) {
  std::vector<double> r(n);
  for (int i = 0; i < n; i++) r[i] = -1.0 + 2.0 * double(std::rand()) / double(RAND_MAX);
  return r;
}
}  // namespace

HermitianLmiCone::HermitianLmiCone(int n, int m)
    : DenseLmiCone(n, m, std::vector<double>((size_t)n * n * m, 0.0).data(),
                   std::vector<double>((size_t)n * n, 0.0).data()) {}

void HermitianLmiCone::PrepareStep(const StepOptions& opt, const double* y, StepInfo* info) {
  // conex/hermitian_psd.cc:34-74: minus_s and WS are distinct matrices here
  const int n = n_;
  View minus_s = temp_1_, WS = temp_2_;
  ComputeNegativeSlack(opt.c_weight, y, minus_s);
  Gemm(false, false, n, n, n, 1.0, W_.p, n, minus_s.p, n, 0.0, WS.p, n);
  if (opt.affine) {
    std::vector<double> WSW(WS.size());
    Gemm(false, false, n, n, n, 1.0, WS.p, n, W_.p, n, 0.0, WSW.data(), n);
    if (opt.e_weight != 0) {
      for (size_t k = 0; k < W_.size(); k++) W_.p[k] *= (1 + opt.e_weight);
    }
    for (size_t k = 0; k < W_.size(); k++) W_.p[k] += WSW[k];
    return;
  }
  const auto r = EigenRandomVector(n);
  const auto eig = HermitianLanczos(n, WS.p, W_.p, r.data(), n / 2 + 1);
  const double lambda_1 = std::fabs(opt.e_weight + eig.front());
  const double lambda_2 = std::fabs(opt.e_weight + eig.back());
  std::vector<double> WSWS(WS.size());
  Gemm(false, false, n, n, n, 1.0, WS.p, n, WS.p, n, 0.0, WSWS.data(), n);
  double tr2 = 0, tr = 0;
  for (int i = 0; i < n; i++) {
    tr2 += WSWS[(size_t)i * n + i];
    tr += WS(i, i);
  }
  info->norminfd = std::max(lambda_1, lambda_2);
  info->normsqrd = tr2 + 2 * tr + n;
}

bool HermitianLmiCone::TakeStep(const StepOptions& opt) {
  // conex/hermitian_psd.cc:9-31
  const int n = n_;
  View WS = temp_2_;
  for (int i = 0; i < n; i++) WS(i, i) += opt.e_weight;
  if (opt.step_size != 1.0) {
    for (size_t k = 0; k < WS.size(); k++) WS.p[k] *= opt.step_size;
  }
  std::vector<double> expWS(WS.size()), prod(WS.size());
  ExponentialMapTaylor(n, WS.p, expWS.data());
  Gemm(false, false, n, n, n, 1.0, expWS.data(), n, W_.p, n, 0.0, prod.data(), n);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) W_(i, j) = (prod[(size_t)j * n + i] + prod[(size_t)i * n + j]) * .5;
  return true;
}

void HermitianLmiCone::GetWeightedSlackEigenvalues(const double* y, double c_weight, SlackEigenvalues* p) {
  // conex/hermitian_psd.cc:76-95
  const int n = n_;
  View minus_s = temp_1_, WS = temp_2_;
  ComputeNegativeSlack(c_weight, y, minus_s);
  Gemm(false, false, n, n, n, 1.0, W_.p, n, minus_s.p, n, 0.0, WS.p, n);
  const auto r = EigenRandomVector(n);
  const auto eig = HermitianLanczos(n, WS.p, W_.p, r.data(), n / 2 + 1);
  p->lambda_max = -eig.front();
  p->lambda_min = -eig.back();
  std::vector<double> WSWS(WS.size());
  Gemm(false, false, n, n, n, 1.0, WS.p, n, WS.p, n, 0.0, WSWS.data(), n);
  double tr2 = 0, tr = 0;
  for (int i = 0; i < n; i++) {
    tr2 += WSWS[(size_t)i * n + i];
    tr += WS(i, i);
  }
  p->frobenius_norm_squared = tr2;
  p->trace = -tr;
}

}  // namespace oracle
