#include "divergence.h"

#include <algorithm>
#include <cmath>

namespace conex {
namespace {

// ||k*lambda - 1||_inf over the spectrum interval [lambda_min, lambda_max].
double NormInfOfScaledSlack(double k, const WeightedSlackEigenvalues& p) {
  return std::max(std::fabs(k * p.lambda_max - 1), std::fabs(k * p.lambda_min - 1));
}

// Branch where the bound's denominator is 2 - k*lambda_max: solve
// (F k^2 - 2 t k + r) / (2 - lambda_max k) = bound for its larger root (divergence.cc:11-41).
double BranchLambdaMax(double bound, const WeightedSlackEigenvalues& p) {
  const double F = p.frobenius_norm_squared, t = p.trace, r = p.rank, d = p.lambda_max;
  const double lin = -2 * t + d * bound;  // coefficient of k after clearing the denominator
  const double disc = 4 * t * t - 4 * F * r + 8 * F * bound - 4 * t * d * bound + d * d * bound * bound;
  const double root = (-lin + std::sqrt(disc)) / (2 * F);
  return (root >= 2.0 / (p.lambda_max + p.lambda_min)) ? root : -1;
}

// Branch where the denominator is k*lambda_min: solve (F k - 2 t + r / k) / lambda_min = bound,
// keep the largest root inside [0, 2/(lambda_max + lambda_min)] (divergence.cc:47-85).
double BranchLambdaMin(double bound, const WeightedSlackEigenvalues& p) {
  const double a = p.frobenius_norm_squared / p.lambda_min;
  const double s = 2 * p.trace / p.lambda_min + bound;  // (b + c) in the reference's notation
  const double disc = s * s - 4 * a * (p.rank / p.lambda_min);
  if (disc < 0) return -1;
  const double hi = 2.0 / (p.lambda_max + p.lambda_min);
  const double roots[2] = {(s + std::sqrt(disc)) / (2 * a), (s - std::sqrt(disc)) / (2 * a)};
  double k = -1;
  for (double x : roots) {
    if (x >= 0 && x <= hi && x > k) k = x;
  }
  return k;
}

}  // namespace

double DivergenceUpperBoundInverse(double bound, WeightedSlackEigenvalues& p) {
  const double k_min_branch = BranchLambdaMin(bound, p);
  const double k_max_branch = BranchLambdaMax(bound, p);
  double k = -1;
  if (NormInfOfScaledSlack(k_min_branch, p) < 1) k = k_min_branch;
  if (k_max_branch > k && NormInfOfScaledSlack(k_max_branch, p) < 1) k = k_max_branch;
  return k;
}

double DivergenceUpperBound(double k, WeightedSlackEigenvalues& p) {
  const double numerator = k * k * p.frobenius_norm_squared - 2 * k * p.trace + p.rank;
  return numerator / (1 - NormInfOfScaledSlack(k, p));
}

}  // namespace conex
