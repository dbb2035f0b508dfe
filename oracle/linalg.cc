// ORACLE — TEST INFRASTRUCTURE ONLY. See linalg.h.
#include "linalg.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#ifndef ORACLE_BLAS_PATH
#define ORACLE_BLAS_PATH ""
#endif

namespace oracle {
namespace {

// cblas / LAPACKE enums (values fixed by the CBLAS standard).
enum { kColMajor = 102, kNoTrans = 111, kTrans = 112, kLower = 122, kNonUnit = 131 };

struct Blas {
  void* handle = nullptr;
  void (*dgemm)(int, int, int, int, int, int, double, const double*, int,
                const double*, int, double, double*, int) = nullptr;
  void (*dgemv)(int, int, int, int, double, const double*, int, const double*, int,
                double, double*, int) = nullptr;
  void (*dtrsv)(int, int, int, int, int, const double*, int, double*, int) = nullptr;
  int (*dpotrf)(int, char, int, double*, int) = nullptr;
  int (*dgetrf)(int, int, int, double*, int, int*) = nullptr;
  int (*dgetrs)(int, char, int, int, const double*, int, const int*, double*,
                int) = nullptr;
  int (*dsterf)(int, double*, double*) = nullptr;
  void (*set_threads)(int) = nullptr;
  int (*get_threads)() = nullptr;
  bool ok = false;
};

Blas g_blas;
bool g_force_plain = false;
std::once_flag g_once;

template <typename F>
bool Load(void* h, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(h, name));
  return *out != nullptr;
}

void InitBlas() {
  const char* env = std::getenv("CONEX_ORACLE_BLAS");
  if (env && std::string(env) == "none") return;
  std::string path = env ? env : ORACLE_BLAS_PATH;
  if (path.empty()) return;
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) return;
  Blas b;
  b.handle = h;
  bool ok = Load(h, "scipy_cblas_dgemm", &b.dgemm) &&
            Load(h, "scipy_cblas_dgemv", &b.dgemv) &&
            Load(h, "scipy_cblas_dtrsv", &b.dtrsv) &&
            Load(h, "scipy_LAPACKE_dpotrf", &b.dpotrf) &&
            Load(h, "scipy_LAPACKE_dgetrf", &b.dgetrf) &&
            Load(h, "scipy_LAPACKE_dgetrs", &b.dgetrs) &&
            Load(h, "scipy_LAPACKE_dsterf", &b.dsterf) &&
            Load(h, "scipy_openblas_set_num_threads", &b.set_threads) &&
            Load(h, "scipy_openblas_get_num_threads", &b.get_threads);
  b.ok = ok;
  if (ok) g_blas = b;
}

bool UseBlas() {
  std::call_once(g_once, InitBlas);
  return g_blas.ok && !g_force_plain;
}

}  // namespace

bool BlasAvailable() {
  std::call_once(g_once, InitBlas);
  return g_blas.ok;
}
void ForcePlainLoops(bool on) { g_force_plain = on; }
void SetBlasThreads(int n) {
  if (BlasAvailable()) g_blas.set_threads(n);
}
int GetBlasThreads() { return BlasAvailable() ? g_blas.get_threads() : 1; }

double Dot(size_t n, const double* x, const double* y) {
  double s = 0;
  for (size_t i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}

void Gemm(bool ta, bool tb, int M, int N, int K, double alpha, const double* A,
          int lda, const double* B, int ldb, double beta, double* C, int ldc) {
  if (M == 0 || N == 0) return;
  if (UseBlas() && (size_t)M * N * K > 4096) {
    g_blas.dgemm(kColMajor, ta ? kTrans : kNoTrans, tb ? kTrans : kNoTrans, M, N, K,
                 alpha, A, lda, B, ldb, beta, C, ldc);
    return;
  }
  for (int j = 0; j < N; j++) {
    double* c = C + (size_t)j * ldc;
    if (beta == 0) {
      for (int i = 0; i < M; i++) c[i] = 0;
    } else if (beta != 1) {
      for (int i = 0; i < M; i++) c[i] *= beta;
    }
    for (int k = 0; k < K; k++) {
      const double b = alpha * (tb ? B[(size_t)k * ldb + j] : B[(size_t)j * ldb + k]);
      if (b == 0) continue;
      if (!ta) {
        const double* a = A + (size_t)k * lda;
        for (int i = 0; i < M; i++) c[i] += a[i] * b;
      } else {
        for (int i = 0; i < M; i++) c[i] += A[(size_t)i * lda + k] * b;
      }
    }
  }
}

void Gemv(bool trans, int M, int N, double alpha, const double* A, int lda,
          const double* x, double beta, double* y) {
  const int ylen = trans ? N : M;
  if (ylen == 0) return;
  if (UseBlas() && (size_t)M * N > 4096) {
    g_blas.dgemv(kColMajor, trans ? kTrans : kNoTrans, M, N, alpha, A, lda, x, 1, beta,
                 y, 1);
    return;
  }
  if (beta == 0) {
    for (int i = 0; i < ylen; i++) y[i] = 0;
  } else if (beta != 1) {
    for (int i = 0; i < ylen; i++) y[i] *= beta;
  }
  if (!trans) {
    for (int j = 0; j < N; j++) {
      const double xj = alpha * x[j];
      const double* a = A + (size_t)j * lda;
      for (int i = 0; i < M; i++) y[i] += a[i] * xj;
    }
  } else {
    for (int j = 0; j < N; j++) {
      y[j] += alpha * Dot(M, A + (size_t)j * lda, x);
    }
  }
}

bool CholeskyLower(int n, double* A, int lda) {
  if (UseBlas() && n > 16) {
    return g_blas.dpotrf(kColMajor, 'L', n, A, lda) == 0;
  }
  // Left-looking, column by column.
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * lda + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)k * lda + j] * A[(size_t)k * lda + j];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * lda + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)j * lda + i];
      for (int k = 0; k < j; k++) s -= A[(size_t)k * lda + i] * A[(size_t)k * lda + j];
      A[(size_t)j * lda + i] = s / d;
    }
  }
  return true;
}

bool LdltLower(int n, double* A, int lda, std::vector<int>* transpositions) {
  // conex/RLDLT.h:297-431 (rldlt_inplace<Lower>::unblocked), real scalars.
  auto a = [&](int i, int j) -> double& { return A[(size_t)j * lda + i]; };
  const double reg = 1e-9;
  transpositions->assign(n, 0);
  bool ok = true;
  if (n <= 1) {
    if (n == 1) {
      if (std::fabs(a(0, 0)) < reg) a(0, 0) = (a(0, 0) < 0) ? -reg : reg;  // :310-317 (returns true)
    }
    return true;
  }
  std::vector<double> temp(n);
  for (int k = 0; k < n; k++) {
    // largest |diagonal| of the trailing part, first occurrence (:330-334)
    int p = k;
    double best = std::fabs(a(k, k));
    for (int i = k + 1; i < n; i++) {
      if (std::fabs(a(i, i)) > best) {
        best = std::fabs(a(i, i));
        p = i;
      }
    }
    (*transpositions)[k] = p;
    if (p != k) {
      // symmetric swap restricted to the lower triangle (:336-353)
      for (int j = 0; j < k; j++) std::swap(a(k, j), a(p, j));
      for (int i = p + 1; i < n; i++) std::swap(a(i, k), a(i, p));
      std::swap(a(k, k), a(p, p));
      for (int i = k + 1; i < p; i++) std::swap(a(i, k), a(p, i));
    }
    const int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; j++) temp[j] = a(j, j) * a(k, j);  // D * A10^T (:365-366)
      double s = 0;
      for (int j = 0; j < k; j++) s += a(k, j) * temp[j];
      a(k, k) -= s;
      for (int i = 0; i < rs; i++) {
        double t = 0;
        for (int j = 0; j < k; j++) t += a(k + 1 + i, j) * temp[j];
        a(k + 1 + i, k) -= t;
      }
    }
    double akk = a(k, k);
    if (!(std::fabs(akk) > 1e-9)) {  // :378-389
      ok = false;
      a(k, k) = (akk < 0) ? -reg : reg;
      akk = a(k, k);
    }
    for (int i = 0; i < rs; i++) a(k + 1 + i, k) /= akk;
  }
  return ok;
}

void SolveLdlt(int n, const double* LD, int lda, const std::vector<int>& tr, double* x) {
  auto a = [&](int i, int j) { return LD[(size_t)j * lda + i]; };
  for (int k = 0; k < n; k++) std::swap(x[k], x[tr[k]]);          // P x
  for (int j = 0; j < n; j++)                                      // L^{-1}
    for (int i = j + 1; i < n; i++) x[i] -= a(i, j) * x[j];
  for (int j = 0; j < n; j++) x[j] /= a(j, j);                     // D^{-1}
  for (int j = n - 1; j >= 0; j--)                                 // L^{-T}
    for (int i = j + 1; i < n; i++) x[j] -= a(i, j) * x[i];
  for (int k = n - 1; k >= 0; k--) std::swap(x[k], x[tr[k]]);      // P^T
}

void SolveLower(int n, const double* L, int ldl, double* x, bool transpose) {
  if (UseBlas() && n > 16) {
    g_blas.dtrsv(kColMajor, kLower, transpose ? kTrans : kNoTrans, kNonUnit, n, L, ldl,
                 x, 1);
    return;
  }
  if (!transpose) {
    for (int j = 0; j < n; j++) {
      x[j] /= L[(size_t)j * ldl + j];
      const double xj = x[j];
      for (int i = j + 1; i < n; i++) x[i] -= L[(size_t)j * ldl + i] * xj;
    }
  } else {
    for (int j = n - 1; j >= 0; j--) {
      double s = x[j];
      for (int i = j + 1; i < n; i++) s -= L[(size_t)j * ldl + i] * x[i];
      x[j] = s / L[(size_t)j * ldl + j];
    }
  }
}

bool LuSolve(int n, double* A, int lda, int nrhs, double* B, int ldb) {
  if (UseBlas() && n > 16) {
    std::vector<int> piv(n);
    if (g_blas.dgetrf(kColMajor, n, n, A, lda, piv.data()) != 0) return false;
    return g_blas.dgetrs(kColMajor, 'N', n, nrhs, A, lda, piv.data(), B, ldb) == 0;
  }
  for (int k = 0; k < n; k++) {
    int p = k;
    double best = std::fabs(A[(size_t)k * lda + k]);
    for (int i = k + 1; i < n; i++) {
      const double v = std::fabs(A[(size_t)k * lda + i]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (best == 0) return false;
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(A[(size_t)j * lda + k], A[(size_t)j * lda + p]);
      for (int j = 0; j < nrhs; j++) std::swap(B[(size_t)j * ldb + k], B[(size_t)j * ldb + p]);
    }
    const double inv = 1.0 / A[(size_t)k * lda + k];
    for (int i = k + 1; i < n; i++) A[(size_t)k * lda + i] *= inv;
    for (int j = k + 1; j < n; j++) {
      const double akj = A[(size_t)j * lda + k];
      if (akj == 0) continue;
      for (int i = k + 1; i < n; i++) A[(size_t)j * lda + i] -= A[(size_t)k * lda + i] * akj;
    }
  }
  for (int r = 0; r < nrhs; r++) {
    double* b = B + (size_t)r * ldb;
    for (int j = 0; j < n; j++) {  // unit lower
      const double bj = b[j];
      for (int i = j + 1; i < n; i++) b[i] -= A[(size_t)j * lda + i] * bj;
    }
    for (int j = n - 1; j >= 0; j--) {  // upper
      b[j] /= A[(size_t)j * lda + j];
      const double bj = b[j];
      for (int i = 0; i < j; i++) b[i] -= A[(size_t)j * lda + i] * bj;
    }
  }
  return true;
}

std::vector<double> TridiagonalEigenvalues(std::vector<double> d, std::vector<double> e) {
  const int n = static_cast<int>(d.size());
  if (n == 0) return d;
  if (UseBlas() && n > 2) {
    e.resize(n, 0.0);
    g_blas.dsterf(n, d.data(), e.data());
    std::sort(d.begin(), d.end());
    return d;
  }
  // Implicit QL with Wilkinson shifts (eigenvalues only).
  e.resize(n, 0.0);
  for (int l = 0; l < n; l++) {
    int iter = 0;
    int m;
    do {
      for (m = l; m < n - 1; m++) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) break;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i];
          const double b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  std::sort(d.begin(), d.end());
  return d;
}

std::vector<double> SymmetricEigenvalues(int n, const double* Ain, int lda) {
  std::vector<double> a((size_t)n * n);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) a[(size_t)j * n + i] = Ain[(size_t)j * lda + i];
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < j; i++) off += a[(size_t)j * n + i] * a[(size_t)j * n + i];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) {
      for (int q = p + 1; q < n; q++) {
        const double apq = a[(size_t)q * n + p];
        if (apq == 0) continue;
        const double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
        const double theta = (aqq - app) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) {  // columns p,q
          const double akp = a[(size_t)p * n + k], akq = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * akp - s * akq;
          a[(size_t)q * n + k] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {  // rows p,q
          const double apk = a[(size_t)k * n + p], aqk = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * apk - s * aqk;
          a[(size_t)k * n + q] = s * apk + c * aqk;
        }
      }
    }
  }
  std::vector<double> ev(n);
  for (int i = 0; i < n; i++) ev[i] = a[(size_t)i * n + i];
  std::sort(ev.begin(), ev.end());
  return ev;
}

}  // namespace oracle
