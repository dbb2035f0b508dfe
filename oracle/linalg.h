// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Dense column-major linear algebra used by the CPU restatement of conex's
// geodesic-IPM hot path. This file stands in for Eigen 3.3.9 (the reference's
// un-vendored numeric dependency, /root/reference/WORKSPACE:5-13). Every routine
// has a plain C++ loop implementation; when SciPy's bundled OpenBLAS can be
// dlopen'ed the big ones are routed to cblas/LAPACKE so the CPU baseline is timed
// with a tuned BLAS (the reference itself supports that: .bazelrc:17-34,
// interfaces/Makefile:7-12).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load anything built from oracle/. The product (conex_b200/) never does.
#pragma once
#include <cstddef>
#include <vector>

namespace oracle {

// y = number rounded up to a multiple of 4 (conex/memory_utils.h:4-12).
inline int AlignedSize(int n) { return (n % 4) ? n + 4 - (n % 4) : n; }

// Column-major matrix view on memory owned elsewhere (leading dimension == rows),
// the role Eigen::Map<MatrixXd> plays in the reference (conex/newton_step.h:8-9).
struct View {
  double* p = nullptr;
  int rows = 0;
  int cols = 0;
  View() {}
  View(double* data, int r, int c) : p(data), rows(r), cols(c) {}
  double& operator()(int i, int j) const { return p[(size_t)j * rows + i]; }
  double* col(int j) const { return p + (size_t)j * rows; }
  size_t size() const { return (size_t)rows * cols; }
};

// Owning column-major matrix.
struct Mat {
  int rows = 0;
  int cols = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)j * rows + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * rows + i]; }
  double* data() { return a.data(); }
  const double* data() const { return a.data(); }
  View view() { return View(a.data(), rows, cols); }
};

// --- BLAS backend control -------------------------------------------------------------
// Returns true when SciPy's OpenBLAS was found and is in use.
bool BlasAvailable();
// Force the plain-loop implementations (used by the tests to cross-check both backends).
void ForcePlainLoops(bool on);
void SetBlasThreads(int n);
int GetBlasThreads();

// --- BLAS-like kernels (column-major) -------------------------------------------------
// C = alpha * op(A) * op(B) + beta * C, op(X) = X or X^T. M,N,K are the op() shapes.
void Gemm(bool transA, bool transB, int M, int N, int K, double alpha,
          const double* A, int lda, const double* B, int ldb, double beta,
          double* C, int ldc);
// y = alpha * op(A) x + beta y, A is M x N (before op).
void Gemv(bool trans, int M, int N, double alpha, const double* A, int lda,
          const double* x, double beta, double* y);
double Dot(size_t n, const double* x, const double* y);

// In-place lower Cholesky A = L L^T (only the lower triangle is read or written).
// Returns false on a non-positive pivot — Eigen::LLT::info() != Success,
// conex/block_triangular_operations.cc:193-196.
bool CholeskyLower(int n, double* A, int lda);
// x <- L^{-1} x  or  x <- L^{-T} x  (triangularView<Lower>().solveInPlace,
// conex/block_triangular_operations.cc:125-128,169,181).
void SolveLower(int n, const double* L, int ldl, double* x, bool transpose);
// In-place diagonally pivoted, regularised LDL^T of a symmetric (possibly indefinite) matrix,
// restating Eigen::RLDLT::unblocked (conex/RLDLT.h:297-431): at step k the pivot is the largest
// |diagonal entry| of the *stored* trailing diagonal (left-looking algorithm: those entries are
// still the original ones), a symmetric swap brings it to position k, column k is formed from
// the k previous columns, and a pivot with |d| <= 1e-9 is replaced by +-1e-9. On return the
// strict lower triangle holds L, the diagonal holds D, transpositions[k] is the row swapped with
// k. Returns true when no pivot was regularised (RLDLT::regularization_used() == false).
bool LdltLower(int n, double* A, int lda, std::vector<int>* transpositions);
// x <- P^T L^{-T} D^{-1} L^{-1} P x (conex/block_triangular_operations.cc:222-312 for one block).
void SolveLdlt(int n, const double* LD, int lda, const std::vector<int>& transpositions, double* x);
// Solves A X = B with partial (row) pivoting; A and B are overwritten (B <- X).
// Stands in for Eigen partialPivLu().solve (conex/exponential_map_pade.cc:31).
// Returns false if a zero pivot is met.
bool LuSolve(int n, double* A, int lda, int nrhs, double* B, int ldb);
// Eigenvalues (ascending) of the symmetric tridiagonal matrix with diagonal `alpha`
// and off-diagonal `beta` (size alpha.size()-1). Stands in for
// SelfAdjointEigenSolver::computeFromTridiagonal(EigenvaluesOnly),
// conex/approximate_eigenvalues.cc:235-237.
std::vector<double> TridiagonalEigenvalues(std::vector<double> alpha,
                                           std::vector<double> beta);
// All eigenvalues (ascending) of a dense symmetric matrix (Jacobi rotations). Used by the
// tests only (exact spectra for the Lanczos known-answer tests).
std::vector<double> SymmetricEigenvalues(int n, const double* A, int lda);

}  // namespace oracle
