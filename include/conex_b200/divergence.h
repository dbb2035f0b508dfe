// mu selection: closed-form inverse of the divergence upper bound (reference conex/divergence.cc).
#pragma once
#include "newton_step.h"

namespace conex {
// Largest k = 1/sqrt(mu) whose divergence bound stays below `divergence_upper_bound`; -1 if none
// (reference divergence.cc:96-111).
double DivergenceUpperBoundInverse(double divergence_upper_bound, WeightedSlackEigenvalues& p);
// reference divergence.cc:113-121
double DivergenceUpperBound(double k, WeightedSlackEigenvalues& p);
}  // namespace conex
