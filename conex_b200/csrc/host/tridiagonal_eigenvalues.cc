#include "tridiagonal_eigenvalues.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace conex {
namespace {

// Number of eigenvalues strictly less than x (Sturm count with the usual pivot safeguard).
int CountBelow(const std::vector<double>& a, const std::vector<double>& b2, double x, double tiny) {
  int count = 0;
  double q = 1;
  for (size_t i = 0; i < a.size(); i++) {
    const double off = (i == 0) ? 0.0 : b2[i - 1] / q;
    q = a[i] - x - off;
    if (std::fabs(q) < tiny) q = -tiny;
    if (q < 0) count++;
  }
  return count;
}

// The k-th smallest eigenvalue (k = 0 .. n-1).
double KthEigenvalue(const std::vector<double>& a, const std::vector<double>& b2, int k, double lo,
                     double hi, double tiny) {
  for (int it = 0; it < 200; it++) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (CountBelow(a, b2, mid, tiny) > k) {
      hi = mid;
    } else {
      lo = mid;
    }
  }
  return 0.5 * (lo + hi);
}

struct Bounds {
  double lo, hi, tiny;
  std::vector<double> b2;
};

Bounds Gershgorin(const std::vector<double>& a, const std::vector<double>& b) {
  Bounds g;
  const size_t n = a.size();
  g.lo = std::numeric_limits<double>::max();
  g.hi = -std::numeric_limits<double>::max();
  g.b2.resize(b.size());
  double scale = 0;
  for (size_t i = 0; i < n; i++) {
    const double r = (i > 0 ? std::fabs(b[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(b[i]) : 0.0);
    g.lo = std::min(g.lo, a[i] - r);
    g.hi = std::max(g.hi, a[i] + r);
    scale = std::max(scale, std::fabs(a[i]) + r);
  }
  for (size_t i = 0; i < b.size(); i++) g.b2[i] = b[i] * b[i];
  const double pad = 4 * std::numeric_limits<double>::epsilon() * (scale + 1e-300) * (double)n;
  g.lo -= pad;
  g.hi += pad;
  g.tiny = std::numeric_limits<double>::min() / std::numeric_limits<double>::epsilon() +
           1e-30 * scale * scale;
  g.tiny = std::max(g.tiny, std::numeric_limits<double>::epsilon() * std::numeric_limits<double>::epsilon() * scale);
  return g;
}

}  // namespace

std::pair<double, double> ExtremeEigenvaluesOfTridiagonal(const std::vector<double>& alpha,
                                                          const std::vector<double>& beta) {
  const int n = static_cast<int>(alpha.size());
  if (n == 0) return {0.0, 0.0};
  if (n == 1) return {alpha[0], alpha[0]};
  const Bounds g = Gershgorin(alpha, beta);
  return {KthEigenvalue(alpha, g.b2, 0, g.lo, g.hi, g.tiny),
          KthEigenvalue(alpha, g.b2, n - 1, g.lo, g.hi, g.tiny)};
}

std::vector<double> EigenvaluesOfTridiagonal(const std::vector<double>& alpha,
                                             const std::vector<double>& beta) {
  const int n = static_cast<int>(alpha.size());
  std::vector<double> ev(n);
  if (n == 0) return ev;
  if (n == 1) {
    ev[0] = alpha[0];
    return ev;
  }
  const Bounds g = Gershgorin(alpha, beta);
  for (int k = 0; k < n; k++) ev[k] = KthEigenvalue(alpha, g.b2, k, g.lo, g.hi, g.tiny);
  return ev;
}

}  // namespace conex
