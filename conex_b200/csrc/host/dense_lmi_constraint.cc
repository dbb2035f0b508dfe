#include "dense_lmi_constraint.h"

#include <algorithm>
#include <cmath>

#include "../../../include/conex_b200_device.h"
#include "tridiagonal_eigenvalues.h"

namespace conex {

// Device-side buffers owned by one LMI block. Shared between copies of the constraint object
// (the plugin concept requires copy-constructible types, reference cone_program.h:47-57).
struct DenseLMIConstraint::Storage {
  DeviceBuffer<double> Aall;     // n^2 x (m+1): constraint matrices then C
  DeviceBuffer<double> B;        // n^2 x (m+2): W A_i W, W C W, W        (K1 output / K2 operand)
  DeviceBuffer<double> T;        // panel scratch for A_i W
  DeviceBuffer<double> scratch;  // Lanczos / Padé / LU work
  DeviceBuffer<double> coef;     // [y; -k] for the slack GEMV
  DeviceBuffer<double> small;    // alpha | beta | reductions
  DeviceBuffer<int> iwork;       // LU pivots + permutation, Lanczos count, LU info
  int panel = 0;
};

namespace {
size_t Sq(int n) { return static_cast<size_t>(n) * n; }
}  // namespace

DenseLMIConstraint::DenseLMIConstraint(int n, int m, const double* A, const double* C)
    : n_(n), m_(m), workspace_(n), data_(std::make_shared<Storage>()) {
  data_->Aall.Resize(Sq(n) * (m + 1));
  CudaCheck(cudaMemcpy(data_->Aall.get(), A, sizeof(double) * Sq(n) * m, cudaMemcpyHostToDevice),
            "upload of LMI matrices");
  CudaCheck(cudaMemcpy(data_->Aall.get() + Sq(n) * m, C, sizeof(double) * Sq(n),
                       cudaMemcpyHostToDevice),
            "upload of LMI affine term");
}

DenseLMIConstraint::DenseLMIConstraint(int n, int m, DevicePointers dev)
    : n_(n), m_(m), workspace_(n), data_(std::make_shared<Storage>()) {
  data_->Aall.Resize(Sq(n) * (m + 1));
  CudaCheck(cudaMemcpy(data_->Aall.get(), dev.A, sizeof(double) * Sq(n) * m, cudaMemcpyDeviceToDevice),
            "copy of LMI matrices");
  CudaCheck(cudaMemcpy(data_->Aall.get() + Sq(n) * m, dev.C, sizeof(double) * Sq(n),
                       cudaMemcpyDeviceToDevice),
            "copy of LMI affine term");
}

const double* DenseLMIConstraint::device_matrices() const { return data_->Aall.get(); }

void DenseLMIConstraint::EnsureScratch() {
  Storage& d = *data_;
  if (d.panel != 0) return;
  const size_t nn = Sq(n_);
  // Panel of constraint matrices scaled per pass: as many as fit in ~1 GiB of scratch.
  const size_t budget = (size_t(1) << 27);  // doubles
  d.panel = static_cast<int>(std::max<size_t>(1, std::min<size_t>(m_ + 1, budget / nn)));
  d.T.Resize(nn * d.panel);
  d.B.Resize(nn * (m_ + 2));
  const size_t work = std::max(cxb_lanczos_worksize(n_), cxb_geodesic_worksize(n_));
  d.scratch.Resize(work);
  d.coef.Resize(m_ + 1);
  d.small.Resize(2 * (n_ / 2 + 2) + 8);
  d.iwork.Resize(2 * n_ + 8);
}

void SetIdentity(DenseLMIConstraint* o) {
  DeviceCheck(cxb_set_identity(o->ctx_->stream(), o->n_, o->workspace_.W.data), "cxb_set_identity");
}

void ConstructSchurComplementSystem(DenseLMIConstraint* o, bool initialize,
                                    SchurComplementSystem* sys) {
  // reference dense_lmi_constraint.cc:62-103; arithmetic described in device/schur.cu.
  o->EnsureScratch();
  auto& d = *o->data_;
  void* s = o->ctx_->stream();
  const int m = o->m_;
  if (!initialize) {
    throw std::runtime_error(
        "conex-b200: DenseLMIConstraint accumulates through the assembler (initialize == true)");
  }
  DeviceCheck(cxb_schur_dense_lmi(s, o->n_, m, d.Aall.get(), o->workspace_.W.data, d.B.get(),
                                  d.T.get(), d.panel, sys->G.data, sys->G.ld),
              "cxb_schur_dense_lmi");
  const long ld = sys->G.ld;
  // rows m and m+1 of the augmented Gram hold AQc, <c,Qc> and AW, <w,c>.
  DeviceCheck(cxb_copy_strided(s, m, sys->G.data + m, ld, sys->AQc, 1), "cxb_copy_strided");
  DeviceCheck(cxb_copy_strided(s, m, sys->G.data + m + 1, ld, sys->AW, 1), "cxb_copy_strided");
  DeviceCheck(cxb_copy_strided(s, 1, sys->G.data + (long)m * ld + m + 1, 1, sys->scalars, 1),
              "cxb_copy_strided");
  DeviceCheck(cxb_copy_strided(s, 1, sys->G.data + (long)m * ld + m, 1, sys->scalars + 1, 1),
              "cxb_copy_strided");
}

void DenseLMIConstraint::ComputeNegativeSlack(double k, const Ref& y, Ref* minus_s) {
  auto& d = *data_;
  ctx_->CopyOnDevice(d.coef.get(), y.data, m_);
  DeviceCheck(cxb_fill(ctx_->stream(), 1, -k, d.coef.get() + m_), "cxb_fill");
  DeviceCheck(cxb_gemv_n(ctx_->stream(), (long)Sq(n_), m_ + 1, d.Aall.get(), d.coef.get(),
                         minus_s->data),
              "cxb_gemv_n");
}

DenseLMIConstraint::SpectrumEstimate DenseLMIConstraint::EstimateSpectrum(const Ref& WS,
                                                                          const Ref& start_matrix) {
  // reference psd_constraint.cc:63-80 / :107-127 and approximate_eigenvalues.cc:241-256.
  auto& d = *data_;
  void* s = ctx_->stream();
  const int n = n_;
  const int num_iter = n / 2;
  const int cap = n / 2 + 2;
  double* alpha = d.small.get();
  double* beta = alpha + cap;
  double* red = beta + cap;  // 8 slots
  int* count = d.iwork.get() + 2 * n;
  DeviceCheck(cxb_ws_reductions(s, n, WS.data, red), "cxb_ws_reductions");
  if (n > 1 && num_iter >= 1) {
    DeviceCheck(cxb_lanczos_two_sided(s, n, WS.data, workspace_.W.data, start_matrix.data, red + 2,
                                      num_iter, alpha, beta, count, d.scratch.get()),
                "cxb_lanczos_two_sided");
  }
  std::vector<double> host(2 * cap + 8);
  int cnt = 0;
  CudaCheck(cudaMemcpyAsync(host.data(), d.small.get(), sizeof(double) * host.size(),
                            cudaMemcpyDeviceToHost, ctx_->cuda_stream()),
            "D2H Lanczos coefficients");
  if (n > 1) {
    ctx_->DownloadInts(&cnt, count, 1);
  } else {
    ctx_->Synchronize();
  }
  SpectrumEstimate e;
  e.trace = host[2 * cap + 0];
  e.trace_of_square = host[2 * cap + 1];
  if (n == 1) {
    e.ritz_min = e.ritz_max = e.trace;  // ApproximateEigenvalues returns WS itself for n == 1
  } else {
    std::vector<double> a(host.begin(), host.begin() + cnt + 1);
    std::vector<double> b(host.begin() + cap, host.begin() + cap + cnt);
    const auto mm = ExtremeEigenvaluesOfTridiagonal(a, b);
    e.ritz_min = mm.first;
    e.ritz_max = mm.second;
  }
  return e;
}

void PrepareStep(DenseLMIConstraint* o, const StepOptions& opt, const Ref& y, StepInfo* info) {
  // reference psd_constraint.cc:45-84. There minus_s and WS alias temp_1; here minus_s lives in
  // temp_1 and WS in temp_2 (WSWS is never formed: tr(WS WS) is an O(n^2) reduction).
  auto& w = o->workspace_;
  void* s = o->ctx_->stream();
  const int n = o->n_;
  Ref minus_s = w.temp_1, WS = w.temp_2;
  o->EnsureScratch();
  o->ComputeNegativeSlack(opt.c_weight, y, &minus_s);
  DeviceCheck(cxb_dgemm(s, 0, 0, n, n, n, 1.0, w.W.data, n, 0, minus_s.data, n, 0, 0.0, WS.data, n, 0,
                        1, 0),
              "cxb_dgemm(W*S)");
  if (opt.affine) {
    // AffineUpdate (psd_constraint.cc:33-43): W <- (1 + e) W + (WS) W
    DeviceCheck(cxb_dgemm(s, 0, 0, n, n, n, 1.0, WS.data, n, 0, w.W.data, n, 0, 0.0, minus_s.data, n,
                          0, 1, 0),
                "cxb_dgemm(WS*W)");
    DeviceCheck(cxb_affine_update(s, n, w.W.data, minus_s.data, opt.e_weight), "cxb_affine_update");
    return;
  }
  // The reference starts Lanczos from minus_s.col(argmax diag WS) *after* WS overwrote minus_s,
  // i.e. from a column of WS (psd_constraint.cc:48-50,66-69).
  const auto e = o->EstimateSpectrum(WS, WS);
  const double lambda_1 = std::fabs(opt.e_weight + e.ritz_min);
  const double lambda_2 = std::fabs(opt.e_weight + e.ritz_max);
  info->norminfd = std::max(lambda_1, lambda_2);
  info->normsqrd = e.trace_of_square + 2 * e.trace + n;
}

bool TakeStep(DenseLMIConstraint* o, const StepOptions& opt) {
  // reference psd_constraint.cc:86-90 -> GeodesicUpdate :13-28
  auto& d = *o->data_;
  auto& w = o->workspace_;
  DeviceCheck(cxb_geodesic_update(o->ctx_->stream(), o->n_, w.W.data, w.temp_2.data, opt.e_weight,
                                  opt.step_size, d.scratch.get(), d.iwork.get(),
                                  d.iwork.get() + 2 * o->n_ + 1),
              "cxb_geodesic_update");
  return true;
}

void GetWeightedSlackEigenvalues(DenseLMIConstraint* o, const Ref& y, double c_weight,
                                 WeightedSlackEigenvalues* p) {
  // reference psd_constraint.cc:97-128: here the start vector is a column of the true -S.
  auto& w = o->workspace_;
  void* s = o->ctx_->stream();
  const int n = o->n_;
  Ref minus_s = w.temp_1, WS = w.temp_2;
  o->EnsureScratch();
  o->ComputeNegativeSlack(c_weight, y, &minus_s);
  DeviceCheck(cxb_dgemm(s, 0, 0, n, n, n, 1.0, w.W.data, n, 0, minus_s.data, n, 0, 0.0, WS.data, n, 0,
                        1, 0),
              "cxb_dgemm(W*S)");
  const auto e = o->EstimateSpectrum(WS, minus_s);
  p->lambda_max = -e.ritz_min;
  p->lambda_min = -e.ritz_max;
  p->frobenius_norm_squared = e.trace_of_square;
  p->trace = -e.trace;
}

}  // namespace conex
