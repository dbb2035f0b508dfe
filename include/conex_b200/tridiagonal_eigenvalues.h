// Extreme eigenvalues of the Lanczos Jacobi matrix. The reference computes the whole spectrum with
// Eigen's SelfAdjointEigenSolver::computeFromTridiagonal and then consumes only min and max
// (approximate_eigenvalues.cc:235-237, psd_constraint.cc:72-77,115-116); here the two extremes are
// found directly by Sturm-sequence bisection to full double precision.
#pragma once
#include <utility>
#include <vector>

namespace conex {
// alpha: diagonal (k entries), beta: off-diagonal (k-1 entries). Returns {lambda_min, lambda_max}.
std::pair<double, double> ExtremeEigenvaluesOfTridiagonal(const std::vector<double>& alpha,
                                                          const std::vector<double>& beta);
// Full spectrum (ascending) by bisection; used by tests and small problems.
std::vector<double> EigenvaluesOfTridiagonal(const std::vector<double>& alpha,
                                             const std::vector<double>& beta);
}  // namespace conex
