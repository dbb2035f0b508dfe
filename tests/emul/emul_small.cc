// TEST INFRASTRUCTURE ONLY — never linked into libconex_b200.so.
//
// Executes the team-phase arithmetic of conex_b200/csrc/device/small_cone_math.cuh serially on the
// CPU ("HostTeam": every parallel loop runs in index order, every barrier is a no-op), behind the
// same argument lists as the cxb_small_* device entry points. The CPU test-suite uses it to check the
// index arithmetic and reference semantics of that header against the oracle where no GPU exists;
// the GPU tests then check the real kernels against the same oracle.
#include <cstddef>
#include <vector>

#include "../../conex_b200/csrc/device/small_cone_math.cuh"
#include "../../include/conex_b200_device.h"

namespace {

struct HostTeam {
  int size() const { return 128; }
  template <class F>
  void par(int n, F f) {
    for (int i = 0; i < n; i++) f(i);
  }
  template <class R, class F>
  void par2(int rows, int cols, R row_value, F f) {
    for (int i = 0; i < rows; i++) {
      const double rv = row_value(i);
      for (int c = 0; c < cols; c++) f(i, c, rv);
    }
  }
  template <class F>
  double sum(int n, F f) {
    double v = 0;
    for (int i = 0; i < n; i++) v += f(i);
    return v;
  }
  template <class F>
  double maxv(int n, F f) {
    double v = -1.7976931348623157e308;
    for (int i = 0; i < n; i++) v = f(i) > v ? f(i) : v;
    return v;
  }
  template <class F>
  double minv(int n, F f) {
    double v = 1.7976931348623157e308;
    for (int i = 0; i < n; i++) v = f(i) < v ? f(i) : v;
    return v;
  }
  template <class F>
  int argmax_first(int n, F f, double* vmax) {
    int best = 0;
    double bv = f(0);
    for (int i = 1; i < n; i++)
      if (f(i) > bv) {
        bv = f(i);
        best = i;
      }
    *vmax = bv;
    return best;
  }
  template <class F>
  int first_true(int n, F pred) {
    for (int i = 0; i < n; i++)
      if (pred(i)) return i;
    return n;
  }
  template <class F>
  double bcast(F f) {
    return f();
  }
  template <class F>
  void single(F f) {
    f();
  }
  template <class F>
  void warp0(F f) {
    f(*this);
  }
};

long Align4(long n) { return (n + 3) & ~3L; }
size_t PsdSmem(int n) { return 6 * (size_t)n * n + 8 * (size_t)n + 64 + 2 * cxb::small::kSections + 16; }

}  // namespace

using namespace cxb::small;

extern "C" {

// Number of (e, d) pairs on which the float-reciprocal Divider of the math header disagrees with integer division.
long emul_divider_mismatches(int max_divisor, int quotients) {
  long bad = 0;
  for (int d = 1; d <= max_divisor; d++) {
    const Divider v(d);
    const long limit = (long)d * quotients < 2147483000L ? (long)d * quotients : 2147483000L;
    const long step = d < 70 ? 1 : 7;
    for (long e = 0; e < limit; e += step) bad += v.quot((int)e) != e / d || v.rem((int)e) != e % d;
    for (long e = limit > 50000 ? limit - 50000 : 0; e < limit; e++) bad += v.quot((int)e) != e / d;
  }
  return bad;
}

int emul_small_set_identity(int batch, const cxb_small_cone* c) {
  HostTeam t;
  for (int p = 0; p < batch; p++) {
    double* st = c->state + p * c->state_stride;
    if (c->type == CXB_CONE_LP) {
      for (int i = 0; i < c->n; i++) st[i] = 1.0;
    } else if (c->type == CXB_CONE_SOC) {
      SocSetIdentity(t, c->n + 1, st);
    } else {
      PsdSetIdentity(t, c->n, st);
    }
  }
  return 0;
}

int emul_small_schur(int batch, const cxb_small_cone* c, double* G, long ldg, long gstride, double* AW,
                     double* AQc, long vstride, double* scal, long sstride, int accumulate) {
  HostTeam t;
  std::vector<double> sm(PsdSchurSmemDoubles(c->n, c->m, t.size()));
  for (int p = 0; p < batch; p++) {
    const double* data = c->data + p * c->data_stride;
    double* st = c->state + p * c->state_stride;
    double* work = c->work ? c->work + p * c->work_stride : nullptr;
    double *g = G + p * gstride, *aw = AW + p * vstride, *aq = AQc + p * vstride, *sc = scal + p * sstride;
    if (c->type == CXB_CONE_LP) {
      std::vector<double> staged(static_cast<size_t>(LpSchurSmemDoubles(c->n, c->m)));
      LpSchur(t, c->n, c->m, data, st, g, ldg, aw, aq, sc, accumulate != 0, staged.data());
    } else if (c->type == CXB_CONE_SOC) {
      SocSchur(t, c->n + 1, c->m, data, st, work, g, ldg, aw, aq, sc, accumulate != 0);
    } else {
      PsdSchur(t, c->n, c->m, data, st, work, sm.data(), g, ldg, aw, aq, sc, accumulate != 0);
    }
  }
  return 0;
}

int emul_small_eigen(int batch, const cxb_small_cone* c, const double* y, long ystride, double cw,
                     const double* cw_p, double* out4, long ostride) {
  HostTeam t;
  std::vector<double> sm(PsdSmem(c->n));
  for (int p = 0; p < batch; p++) {
    const double* data = c->data + p * c->data_stride;
    double* st = c->state + p * c->state_stride;
    double* work = c->work ? c->work + p * c->work_stride : nullptr;
    const double k = cw_p ? cw_p[p] : cw;
    if (c->type == CXB_CONE_LP) {
      const long np = Align4(c->n);
      LpEigen(t, c->n, c->m, data, y + p * ystride, k, st, st + np, st + 2 * np, out4 + p * ostride);
    } else if (c->type == CXB_CONE_SOC) {
      SocEigen(t, c->n + 1, c->m, data, y + p * ystride, k, st, work, out4 + p * ostride);
    } else {
      const long nnp = Align4((long)c->n * c->n);
      PsdEigen(t, c->n, c->m, data, y + p * ystride, k, st, st + nnp, st + 2 * nnp, sm.data(),
               out4 + p * ostride, c->packed ? c->packed + p * c->packed_stride : nullptr);
    }
  }
  return 0;
}

int emul_small_prepare(int batch, const cxb_small_cone* c, const double* y, long ystride, int affine,
                       double cw, const double* cw_p, double ew, double* out2, long ostride) {
  HostTeam t;
  std::vector<double> sm(PsdSmem(c->n));
  for (int p = 0; p < batch; p++) {
    const double* data = c->data + p * c->data_stride;
    double* st = c->state + p * c->state_stride;
    double* work = c->work ? c->work + p * c->work_stride : nullptr;
    const double k = cw_p ? cw_p[p] : cw;
    if (c->type == CXB_CONE_LP) {
      const long np = Align4(c->n);
      LpPrepare(t, c->n, c->m, data, y + p * ystride, affine != 0, k, ew, st, st + np, st + 2 * np,
                out2 + p * ostride);
    } else if (c->type == CXB_CONE_SOC) {
      const long op = Align4(c->n + 1);
      SocPrepare(t, c->n + 1, c->m, data, y + p * ystride, k, st, st + op, work, out2 + p * ostride);
    } else {
      const long nnp = Align4((long)c->n * c->n);
      PsdPrepare(t, c->n, c->m, data, y + p * ystride, affine != 0, k, ew, st, st + nnp, st + 2 * nnp,
                 sm.data(), out2 + p * ostride, c->packed ? c->packed + p * c->packed_stride : nullptr);
    }
  }
  return 0;
}

int emul_small_take_step(int batch, const cxb_small_cone* c, double step, const double* step_p, double ew,
                         int* info) {
  HostTeam t;
  std::vector<double> sm(PsdSmem(c->n));
  for (int p = 0; p < batch; p++) {
    double* st = c->state + p * c->state_stride;
    double* work = c->work ? c->work + p * c->work_stride : nullptr;
    const double s = step_p ? step_p[p] : step;
    if (c->type == CXB_CONE_LP) {
      const long np = Align4(c->n);
      LpTakeStep(t, c->n, s, st, st + 2 * np);
    } else if (c->type == CXB_CONE_SOC) {
      const long op = Align4(c->n + 1);
      SocTakeStep(t, c->n + 1, s, st, st + op, work);
    } else {
      const long nnp = Align4((long)c->n * c->n);
      PsdTakeStep(t, c->n, s, ew, st, st + 2 * nnp, sm.data(), info + p);
    }
  }
  return 0;
}

int emul_small_potrf(int batch, int N, double* H, long ldh, long hstride, int* info) {
  HostTeam t;
  std::vector<double> sm((size_t)N * N);
  for (int p = 0; p < batch; p++) SmallPotrf(t, N, H + p * hstride, ldh, sm.data(), info + p);
  return 0;
}

int emul_small_potrs(int batch, int N, const double* L, long ldl, long lstride, double* X, long xstride) {
  HostTeam t;
  std::vector<double> sm(N);
  for (int p = 0; p < batch; p++) SmallPotrs(t, N, L + p * lstride, ldl, X + p * xstride, sm.data());
  return 0;
}

}  // extern "C"
