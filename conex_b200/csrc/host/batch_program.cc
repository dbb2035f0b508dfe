#include "batch_program.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../../include/conex_b200_device.h"
#include "dense_lmi_constraint.h"
#include "divergence.h"
#include "small_cone_constraint.h"

namespace conex {

namespace {

struct ConeSlot {
  int type = 0, n = 0, m = 0, rows = 0;
  size_t state_size = 0, work_size = 0;
  DeviceBuffer<double> data, state, work;
  DeviceBuffer<double> packed;  // LMI blocks: lower triangles of the matrices, for the slack passes
  cxb_small_cone desc;
  int rank = 0;
};

size_t Align4(size_t n) { return (n + 3) & ~static_cast<size_t>(3); }

}  // namespace

struct BatchProgram::Impl {
  DeviceContext ctx;
  std::vector<ConeSlot> cones;
  std::vector<cxb_small_cone> descs;  // the descriptors of `cones`, contiguous, for the fused launches
  int B = 0, m = 0;
  long ldh = 0;
  size_t vstride = 0;
  // per-program Newton system and vectors
  DeviceBuffer<double> H, AW, AQc, scal, b, y, y2;
  // per-program scalars exchanged with the host each phase
  DeviceBuffer<double> coef;   // 3 x B coefficient arrays | B c_weights | B step sizes
  DeviceBuffer<double> out;    // cone outputs: B x ncones x 4, then B x 4 dots/scalars
  DeviceBuffer<int> info;      // B Cholesky flags | B Padé flags
  DeviceBuffer<int> mask;      // B active | B mu-update
  PinnedBuffer<double> hcoef, hout;
  PinnedBuffer<int> hinfo, hmask;
  int rank = 0;

  void* s() const { return ctx.stream(); }
  void UploadCoef(size_t offset, size_t count) {
    CudaCheck(cudaMemcpyAsync(coef.get() + offset, hcoef.get() + offset, sizeof(double) * count,
                              cudaMemcpyHostToDevice, ctx.cuda_stream()),
              "H2D coefficients");
  }
  void UploadMask(size_t offset) {
    CudaCheck(cudaMemcpyAsync(mask.get() + offset, hmask.get() + offset, sizeof(int) * B,
                              cudaMemcpyHostToDevice, ctx.cuda_stream()),
              "H2D mask");
  }
  void DownloadOut(size_t count) {
    CudaCheck(cudaMemcpyAsync(hout.get(), out.get(), sizeof(double) * count, cudaMemcpyDeviceToHost,
                              ctx.cuda_stream()),
              "D2H outputs");
    ctx.Synchronize();
  }
  void Assemble(const int* active) {
    bool first = true;
    for (auto& c : cones) {
      DeviceCheck(cxb_small_schur(s(), B, &c.desc, H.get(), ldh, ldh * m, AW.get(), AQc.get(), vstride,
                                  scal.get(), 2, first ? 0 : 1, active),
                  "cxb_small_schur");
      first = false;
    }
  }
};

BatchProgram::BatchProgram(const std::vector<Program*>& programs) : impl_(std::make_unique<Impl>()) {
  Impl& d = *impl_;
  if (programs.empty()) throw std::runtime_error("conex-b200: empty batch");
  const int B = static_cast<int>(programs.size());
  batch_ = d.B = B;
  m_ = d.m = programs[0]->GetNumberOfVariables();
  const int m = m_;
  // ---- structure of program 0; every other program must match it ------------------------------
  struct Shape {
    int type, n;
  };
  auto shape_of = [&](Container& c, Shape* sh) {
    if (auto* lp = std::any_cast<LinearConstraint>(&c.obj)) {
      *sh = {CXB_CONE_LP, lp->order()};
    } else if (auto* soc = std::any_cast<SOCConstraint>(&c.obj)) {
      *sh = {CXB_CONE_SOC, soc->order()};
    } else if (auto* lmi = std::any_cast<DenseLMIConstraint>(&c.obj)) {
      *sh = {CXB_CONE_PSD, lmi->order()};
    } else {
      throw std::runtime_error("conex-b200: batched solves support LP, second-order and dense LMI cones");
    }
    if (static_cast<int>(c.variables.size()) != m) {
      throw std::runtime_error("conex-b200: batched solves need every cone on all variables");
    }
    for (int i = 0; i < m; i++) {
      if (c.variables[i] != i) throw std::runtime_error("conex-b200: batched solves need ordered cliques");
    }
  };
  std::vector<Shape> shapes;
  for (auto& c : programs[0]->eqs) {
    Shape sh;
    shape_of(c, &sh);
    shapes.push_back(sh);
  }
  if (shapes.empty()) throw std::runtime_error("conex-b200: batched programs have no constraints");
  const size_t nc = shapes.size();
  d.cones.resize(nc);
  d.rank = 0;
  for (size_t k = 0; k < nc; k++) {
    ConeSlot& c = d.cones[k];
    c.type = shapes[k].type;
    c.n = shapes[k].n;
    c.m = m;
    c.rows = c.type == CXB_CONE_LP ? c.n : (c.type == CXB_CONE_SOC ? c.n + 1 : c.n * c.n);
    c.rank = c.type == CXB_CONE_LP ? c.n : (c.type == CXB_CONE_SOC ? 2 : c.n);
    d.rank += c.rank;
    c.state_size = cxb_small_state_size(c.type, c.n);
    c.work_size = cxb_small_work_size(c.type, c.n, m);
    const size_t per = static_cast<size_t>(c.rows) * (m + 1);
    c.data.Resize(per * B);
    c.state.Resize(c.state_size * B);
    if (c.work_size) c.work.Resize(c.work_size * B);
    c.desc.type = c.type;
    c.desc.n = c.n;
    c.desc.m = m;
    c.desc.data = c.data.get();
    c.desc.data_stride = static_cast<long>(per);
    c.desc.state = c.state.get();
    c.desc.state_stride = static_cast<long>(c.state_size);
    c.desc.work = c.work_size ? c.work.get() : nullptr;
    c.desc.work_stride = static_cast<long>(c.work_size);
    c.desc.packed = nullptr;
    c.desc.packed_stride = 0;
  }
  // ---- pack the cone data -----------------------------------------------------------------------
  for (int p = 0; p < B; p++) {
    Program& prog = *programs[p];
    if (prog.GetNumberOfVariables() != m || prog.NumberOfMultipliers() != 0 || prog.eqs.size() != nc) {
      throw std::runtime_error("conex-b200: programs of a batch must have identical structure");
    }
    size_t k = 0;
    for (auto& c : prog.eqs) {
      Shape sh;
      shape_of(c, &sh);
      if (sh.type != shapes[k].type || sh.n != shapes[k].n) {
        throw std::runtime_error("conex-b200: programs of a batch must have identical structure");
      }
      ConeSlot& slot = d.cones[k];
      const size_t per = static_cast<size_t>(slot.rows) * (m + 1);
      double* dst = slot.data.get() + per * p;
      if (sh.type == CXB_CONE_PSD) {
        auto* lmi = std::any_cast<DenseLMIConstraint>(&c.obj);
        CudaCheck(cudaMemcpyAsync(dst, lmi->device_matrices(), sizeof(double) * per, cudaMemcpyDeviceToDevice,
                                  d.ctx.cuda_stream()),
                  "packing of an LMI block");
      } else {
        const auto& host = (sh.type == CXB_CONE_LP)
                               ? std::any_cast<LinearConstraint>(&c.obj)->host_data()
                               : std::any_cast<SOCConstraint>(&c.obj)->host_data();
        CudaCheck(cudaMemcpyAsync(dst, host.data(), sizeof(double) * per, cudaMemcpyHostToDevice, d.ctx.cuda_stream()),
                  "packing of a small cone");
      }
      k++;
    }
  }
  d.ctx.Synchronize();
  // ---- packed copies of the LMI operators (symmetric matrices: the slack passes read half the bytes) -------------
  {
    DeviceBuffer<int> flags;
    flags.Resize(nc);
    CudaCheck(cudaMemsetAsync(flags.get(), 0, sizeof(int) * nc, d.ctx.cuda_stream()), "memset");
    for (size_t k = 0; k < nc; k++) {
      ConeSlot& c = d.cones[k];
      if (c.type != CXB_CONE_PSD) continue;
      const size_t stride = Align4(static_cast<size_t>(m + 1) * (static_cast<size_t>(c.n) * (c.n + 1) / 2));
      c.packed.Resize(stride * B);
      DeviceCheck(cxb_small_pack_symmetric(d.s(), B, &c.desc, c.packed.get(), static_cast<long>(stride),
                                           flags.get() + k),
                  "cxb_small_pack_symmetric");
    }
    std::vector<int> asymmetric(nc, 0);
    d.ctx.DownloadInts(asymmetric.data(), flags.get(), nc);
    for (size_t k = 0; k < nc; k++) {
      ConeSlot& c = d.cones[k];
      if (c.type == CXB_CONE_PSD && !asymmetric[k]) {
        c.desc.packed = c.packed.get();
        c.desc.packed_stride =
            static_cast<long>(Align4(static_cast<size_t>(m + 1) * (static_cast<size_t>(c.n) * (c.n + 1) / 2)));
      } else {
        c.packed.Resize(0);
      }
    }
    for (auto& c : d.cones) d.descs.push_back(c.desc);
  }
  // ---- Newton-system storage ------------------------------------------------------------------------
  d.ldh = static_cast<long>(Align4(m));
  d.vstride = Align4(m);
  d.H.Resize(static_cast<size_t>(d.ldh) * m * B);
  d.AW.Resize(d.vstride * B);
  d.AQc.Resize(d.vstride * B);
  d.scal.Resize(2 * static_cast<size_t>(B));
  d.b.Resize(d.vstride * B);
  d.y.Resize(d.vstride * B);
  d.y2.Resize(d.vstride * B);
  d.coef.Resize(5 * static_cast<size_t>(B));
  d.out.Resize(static_cast<size_t>(B) * (4 * nc + 4));
  d.info.Resize(2 * static_cast<size_t>(B));
  d.mask.Resize(2 * static_cast<size_t>(B));
  d.hcoef.Reserve(std::max<size_t>(5 * static_cast<size_t>(B), d.vstride * B));
  d.hout.Reserve(std::max<size_t>(static_cast<size_t>(B) * (4 * nc + 4), d.vstride * B));
  d.hinfo.Reserve(2 * static_cast<size_t>(B));
  d.hmask.Reserve(2 * static_cast<size_t>(B));
  results_.assign(B, BatchResult());
}

BatchProgram::~BatchProgram() = default;

int BatchProgram::DualVariableSize(int cone) const {
  const ConeSlot& c = impl_->cones.at(cone);
  return c.type == CXB_CONE_PSD ? c.n * c.n : c.rows;
}

void BatchProgram::GetDualVariable(int p, int cone, double* host_out) {
  Impl& d = *impl_;
  const ConeSlot& c = d.cones.at(cone);
  const int sz = DualVariableSize(cone);
  d.ctx.Download(host_out, c.state.get() + c.state_size * p, sz);
  const BatchResult& r = results_.at(p);
  if (!r.primal_infeasible && r.num_iterations > 0) {
    // reference cone_program.h:120-134: W / (sqrt_inv_mu[last] * b_scaling)
    const double scale = r.inv_sqrt_mu * r.b_scaling;
    for (int j = 0; j < sz; j++) host_out[j] /= scale;
  }
}

namespace {

// Host state of one program of the batch: the locals of Solve() (reference cone_program.cc:276-309).
struct ProgramState {
  double k = 0, kmax = 0, cx = 1, by = -1, kkt_error = 0;
  double b_scaling = 1, c_scaling = 1, b_norm = 0, d_inf = 0;
  int centering_steps = 0, num_iter = 0;
  bool active = true, failed = false, max_iters_reached = true;
  bool initial_centering = false, final_centering = false, update_mu = false;
};

}  // namespace

int BatchProgram::Maximize(const double* b_in, const SolverConfiguration& cfg, double* y_out) {
  Impl& d = *impl_;
  const int B = batch_, m = m_;
  const size_t nc = d.cones.size();
  const size_t vs = d.vstride;
  if (cfg.initialization_mode != CONEX_INITIALIZATION_MODE_COLDSTART) {
    throw std::runtime_error("conex-b200: batched solves are cold starts");
  }
  if (cfg.kkt_solver == CONEX_QR_FACTORIZATION) {
    throw std::runtime_error("conex-b200: the QR KKT mode is not implemented on the device");
  }
  if (cfg.enable_line_search) {
    bool lp_only = true;
    for (const auto& c : d.cones) lp_only = lp_only && (c.type == CXB_CONE_LP);
    if (lp_only) throw std::runtime_error("conex-b200: enable_line_search on LP-only programs is not implemented");
  }
  void* s = d.s();
  cudaStream_t cs = d.ctx.cuda_stream();
  std::vector<ProgramState> st(B);
  results_.assign(B, BatchResult());
  step_ms_.clear();

  // ---- cost vectors, cold start -------------------------------------------------------------------
  for (int p = 0; p < B; p++) {
    double nrm = 0;
    for (int i = 0; i < m; i++) {
      const double v = b_in[static_cast<size_t>(p) * m + i];
      d.hcoef[p * vs + i] = v;
      nrm += v * v;
    }
    for (size_t i = m; i < vs; i++) d.hcoef[p * vs + i] = 0;
    st[p].b_norm = std::sqrt(nrm);
    st[p].kmax = cfg.inv_sqrt_mu_max;
  }
  CudaCheck(cudaMemcpyAsync(d.b.get(), d.hcoef.get(), sizeof(double) * vs * B, cudaMemcpyHostToDevice, cs),
            "H2D cost vectors");
  d.ctx.Synchronize();  // hcoef is reused below
  d.ctx.Zero(d.y.get(), vs * B);
  const int ncones = static_cast<int>(nc);
  DeviceCheck(cxb_small_set_identity_multi(s, B, ncones, d.descs.data(), nullptr), "cxb_small_set_identity_multi");
  cudaEvent_t ev_begin, ev_end, ev_a, ev_b;
  CudaCheck(cudaEventCreate(&ev_begin), "cudaEventCreate");
  CudaCheck(cudaEventCreate(&ev_end), "cudaEventCreate");
  CudaCheck(cudaEventCreate(&ev_a), "cudaEventCreate");
  CudaCheck(cudaEventCreate(&ev_b), "cudaEventCreate");
  CudaCheck(cudaEventRecord(ev_begin, cs), "cudaEventRecord");

  int* active = d.mask.get();
  int* mu_mask = d.mask.get() + B;
  double* coef_a = d.coef.get();
  double* coef_b = d.coef.get() + B;
  double* coef_c = d.coef.get() + 2 * static_cast<size_t>(B);
  double* coef_cw = d.coef.get() + 3 * static_cast<size_t>(B);
  double* coef_step = d.coef.get() + 4 * static_cast<size_t>(B);
  double* cone_out = d.out.get();                                   // B x nc x 4
  double* dots = d.out.get() + static_cast<size_t>(B) * 4 * nc;     // B x 4
  const long ostride = static_cast<long>(4 * nc);
  const int rank = d.rank;

  for (int i = 0; i < cfg.max_iterations; i++) {
    // ---- loop control of every program (reference cone_program.cc:312-336) -----------------------
    int nactive = 0;
    for (int p = 0; p < B; p++) {
      ProgramState& q = st[p];
      d.hmask[p] = 0;
      if (!q.active) continue;
      q.initial_centering = i < cfg.initial_centering_steps_coldstart;
      q.final_centering = (q.k >= q.kmax) || (q.kkt_error > cfg.kkt_error_tolerance) ||
                          (i >= cfg.max_iterations - cfg.final_centering_steps);
      q.update_mu = (i == 0) || !(q.initial_centering || q.final_centering);
      if (q.final_centering && q.centering_steps >= cfg.final_centering_steps) {
        q.max_iters_reached = (i >= cfg.max_iterations - 1);
        q.active = false;
        continue;
      }
      d.hmask[p] = 1;
      nactive++;
    }
    if (nactive == 0) break;
    d.UploadMask(0);
    CudaCheck(cudaEventRecord(ev_a, cs), "cudaEventRecord");

    // ---- assemble, (rescale,) factor ---------------------------------------------------------------
    d.Assemble(active);
    const bool rescale = (i == 0) && cfg.enable_rescaling;
    if (rescale) {
      DeviceCheck(cxb_batched_dot(s, B, m, d.AQc.get(), vs, d.AQc.get(), vs, dots, 4), "cxb_batched_dot");
    }
    DeviceCheck(cxb_small_potrf(s, B, m, d.H.get(), d.ldh, d.ldh * m, d.info.get(), active), "cxb_small_potrf");
    CudaCheck(cudaMemcpyAsync(d.hinfo.get(), d.info.get(), sizeof(int) * B, cudaMemcpyDeviceToHost, cs), "D2H info");
    if (rescale) {
      CudaCheck(cudaMemcpyAsync(d.hout.get(), dots, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, cs), "D2H norms");
    }
    d.ctx.Synchronize();
    bool mask_changed = false;
    bool any_mu = false;
    for (int p = 0; p < B; p++) {
      ProgramState& q = st[p];
      d.hmask[B + p] = 0;
      if (!q.active) continue;
      if (rescale) {
        // reference cone_program.cc:343-357 (cold start)
        q.b_scaling = 1.0 / (1 + q.b_norm);
        q.c_scaling = 1.0 / (1 + std::sqrt(d.hout[4 * p]));
        const double mu_target = (1.0 / (q.kmax * q.kmax)) * (q.b_scaling * q.c_scaling);
        q.kmax = 1.0 / std::sqrt(mu_target);
      }
      if (d.hinfo[p] != 0) {  // Factor() failed: the solve returns unsolved (cone_program.cc:360-371)
        q.active = false;
        q.failed = true;
        d.hmask[p] = 0;
        mask_changed = true;
        continue;
      }
      if (q.update_mu) {
        d.hmask[B + p] = 1;
        any_mu = true;
      }
    }
    if (mask_changed) d.UploadMask(0);

    // ---- mu from the divergence bound (reference cone_program.cc:173-214) --------------------------
    if (any_mu) {
      d.UploadMask(B);
      for (int p = 0; p < B; p++) {
        d.hcoef[p] = st[p].c_scaling;
        d.hcoef[B + p] = -st[p].b_scaling;
        d.hcoef[3 * static_cast<size_t>(B) + p] = st[p].c_scaling;
      }
      d.UploadCoef(0, 2 * static_cast<size_t>(B));
      d.UploadCoef(3 * static_cast<size_t>(B), B);
      DeviceCheck(cxb_batched_lincomb(s, B, m, coef_a, d.AQc.get(), vs, coef_b, d.b.get(), vs, nullptr, nullptr, 0,
                                      d.y.get(), vs, mu_mask),
                  "cxb_batched_lincomb");
      DeviceCheck(cxb_small_potrs(s, B, m, d.H.get(), d.ldh, d.ldh * m, d.y.get(), vs, mu_mask), "cxb_small_potrs");
      DeviceCheck(cxb_small_eigen_multi(s, B, ncones, d.descs.data(), d.y.get(), vs, 0.0, coef_cw, cone_out, ostride,
                                        mu_mask),
                  "cxb_small_eigen_multi");
      d.DownloadOut(static_cast<size_t>(B) * 4 * nc);
      for (int p = 0; p < B; p++) {
        ProgramState& q = st[p];
        if (!d.hmask[B + p]) continue;
        WeightedSlackEigenvalues w;
        w.frobenius_norm_squared = 0;
        w.trace = 0;
        w.lambda_max = -30000;
        w.lambda_min = 30000;
        for (size_t k = 0; k < nc; k++) {
          const double* o = d.hout.get() + (static_cast<size_t>(p) * nc + k) * 4;
          w.lambda_min = std::min(w.lambda_min, o[0]);
          w.lambda_max = std::max(w.lambda_max, o[1]);
          w.frobenius_norm_squared += o[2];
          w.trace += o[3];
        }
        w.rank = rank;
        double cand = -1;
        if (cfg.enable_line_search) cand = q.k;  // see NewtonDriver::Run
        if (cand < 0) {
          cand = DivergenceUpperBoundInverse(cfg.divergence_upper_bound * rank, w);
          if (cand == -1) cand = (w.lambda_min > 0) ? 2.0 / (w.lambda_min + w.lambda_max) : -1;
          if (cand < 0 && w.trace > 1e-12) {
            const double kstar = w.trace / w.frobenius_norm_squared;
            double norm_bound = 1.5 * (w.frobenius_norm_squared * kstar * kstar - 2 * w.trace * kstar + rank);
            norm_bound = std::min(norm_bound, rank * .7);
            const double qa = w.frobenius_norm_squared, qb = -2 * w.trace, qc = rank - norm_bound;
            const double disc = qb * qb - 4 * qa * qc;
            cand = (disc < 0) ? kstar : (-qb + std::sqrt(disc)) / (2 * qa);
          }
        }
        q.k = (cand > 0) ? cand : 0.5 * q.k;
      }
    }
    for (int p = 0; p < B; p++) {
      ProgramState& q = st[p];
      if (!q.active) continue;
      if (!q.update_mu && !q.initial_centering) q.centering_steps++;
      q.k = std::max(std::min(q.k, q.kmax), std::sqrt(1.0 / (1e-15 + cfg.maximum_mu)));
      d.hcoef[p] = q.k * q.b_scaling;
      d.hcoef[B + p] = q.k * q.c_scaling;
      d.hcoef[2 * static_cast<size_t>(B) + p] = -2.0;
      d.hcoef[3 * static_cast<size_t>(B) + p] = q.k * q.c_scaling;
    }

    // ---- Newton direction: y = H^{-1} (k (b b_s + AQc c_s) - 2 AW) (cone_program.cc:409-414) ------
    d.UploadCoef(0, 4 * static_cast<size_t>(B));
    DeviceCheck(cxb_batched_lincomb(s, B, m, coef_a, d.b.get(), vs, coef_b, d.AQc.get(), vs, coef_c, d.AW.get(), vs,
                                    d.y.get(), vs, active),
                "cxb_batched_lincomb");
    DeviceCheck(cxb_small_potrs(s, B, m, d.H.get(), d.ldh, d.ldh * m, d.y.get(), vs, active), "cxb_small_potrs");
    DeviceCheck(cxb_small_prepare_multi(s, B, ncones, d.descs.data(), d.y.get(), vs, 0, 0.0, coef_cw, 1.0, cone_out,
                                        ostride, active),
                "cxb_small_prepare_multi");
    d.DownloadOut(static_cast<size_t>(B) * 4 * nc);
    std::vector<double> normsq(B, 0.0);
    for (int p = 0; p < B; p++) {
      ProgramState& q = st[p];
      d.hcoef[4 * static_cast<size_t>(B) + p] = 1.0;
      if (!q.active) continue;
      double ninf = -1, nsq = 0;
      for (size_t k = 0; k < nc; k++) {
        const double* o = d.hout.get() + (static_cast<size_t>(p) * nc + k) * 4;
        ninf = std::max(ninf, o[0]);
        nsq += o[1];
      }
      q.d_inf = std::fabs(ninf);
      normsq[p] = nsq;
      d.hcoef[4 * static_cast<size_t>(B) + p] = std::min(1.0, 2.0 / (ninf * ninf));
    }
    d.UploadCoef(4 * static_cast<size_t>(B), B);
    DeviceCheck(cxb_small_take_step_multi(s, B, ncones, d.descs.data(), 1.0, coef_step, 1.0, d.info.get() + B, active),
                "cxb_small_take_step_multi");
    // ---- objectives (cone_program.cc:441-467) -------------------------------------------------------
    DeviceCheck(cxb_batched_dot(s, B, m, d.b.get(), vs, d.y.get(), vs, dots + 0, 4), "cxb_batched_dot");
    DeviceCheck(cxb_batched_dot(s, B, m, d.AQc.get(), vs, d.y.get(), vs, dots + 1, 4), "cxb_batched_dot");
    CudaCheck(cudaEventRecord(ev_b, cs), "cudaEventRecord");
    CudaCheck(cudaMemcpyAsync(d.hout.get(), dots, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, cs), "D2H dots");
    CudaCheck(cudaMemcpyAsync(d.hout.get() + 4 * static_cast<size_t>(B), d.scal.get(), sizeof(double) * 2 * B,
                              cudaMemcpyDeviceToHost, cs),
              "D2H scalars");
    d.ctx.Synchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_a, ev_b);
    step_ms_.push_back(ms);
    for (int p = 0; p < B; p++) {
      ProgramState& q = st[p];
      if (!q.active) continue;
      const double by_dot = d.hout[4 * p + 0], aqc_dot = d.hout[4 * p + 1];
      const double wc = d.hout[4 * static_cast<size_t>(B) + 2 * p], cqc = d.hout[4 * static_cast<size_t>(B) + 2 * p + 1];
      const double k = q.k;
      const double d_2 = std::sqrt(std::fabs(normsq[p]));
      q.by = by_dot / (k * q.c_scaling);
      q.cx = (2 * wc + aqc_dot - k * cqc * q.c_scaling) / (k * q.b_scaling);
      const double mu = (1.0 / k) * (1.0 / k);
      const double s_dot_x = mu * (rank - d_2 * d_2) / (q.b_scaling * q.c_scaling);
      q.kkt_error = std::fabs(q.cx - q.by - s_dot_x) / s_dot_x;
      q.num_iter = i + 1;
      if ((q.final_centering || k >= q.kmax) && q.d_inf <= cfg.final_centering_tolerance) {
        q.max_iters_reached = false;
        q.active = false;
      }
    }
  }

  // ---- results (cone_program.cc:486-530) -------------------------------------------------------------
  CudaCheck(cudaMemcpyAsync(d.hout.get(), d.y.get(), sizeof(double) * vs * B, cudaMemcpyDeviceToHost, cs), "D2H y");
  d.ctx.Synchronize();
  int solved_count = 0;
  bool any_recover = false;
  for (int p = 0; p < B; p++) {
    ProgramState& q = st[p];
    BatchResult& r = results_[p];
    r.num_iterations = q.num_iter;
    r.by = q.by;
    r.cx = q.cx;
    r.inv_sqrt_mu = q.k;
    r.b_scaling = q.b_scaling;
    r.c_scaling = q.c_scaling;
    r.d_inf = q.d_inf;
    double* yp = y_out + static_cast<size_t>(p) * m;
    for (int j = 0; j < m; j++) yp[j] = d.hout[p * vs + j];
    d.hmask[p] = 0;
    if (q.failed) {
      r.solved = 0;
      continue;
    }
    const double mu_final = (1.0 / q.k) * (1.0 / q.k);
    if (mu_final > cfg.infeasibility_threshold) {
      r.solved = 0;
      r.primal_infeasible = q.cx * q.k <= -.5;
      r.dual_infeasible = q.by * q.k >= .5;
    } else {
      r.solved = 1;
    }
    d.hmask[p] = 1;
    any_recover = true;
    if (r.solved) {
      for (int j = 0; j < m; j++) yp[j] = yp[j] / q.k / q.c_scaling;
      if (q.max_iters_reached) r.solved = 0;
    }
    solved_count += r.solved;
  }
  if (cfg.prepare_dual_variables && any_recover) {
    // reference cone_program.cc:500-516
    d.UploadMask(0);
    d.Assemble(active);
    DeviceCheck(cxb_small_potrf(s, B, m, d.H.get(), d.ldh, d.ldh * m, d.info.get(), active), "cxb_small_potrf");
    for (int p = 0; p < B; p++) {
      d.hcoef[p] = st[p].k * st[p].b_scaling;
      d.hcoef[B + p] = -1.0;
    }
    d.UploadCoef(0, 2 * static_cast<size_t>(B));
    DeviceCheck(cxb_batched_lincomb(s, B, m, coef_a, d.b.get(), vs, coef_b, d.AW.get(), vs, nullptr, nullptr, 0,
                                    d.y2.get(), vs, active),
                "cxb_batched_lincomb");
    DeviceCheck(cxb_small_potrs(s, B, m, d.H.get(), d.ldh, d.ldh * m, d.y2.get(), vs, active), "cxb_small_potrs");
    DeviceCheck(cxb_small_prepare_multi(s, B, ncones, d.descs.data(), d.y2.get(), vs, 1, 0.0, nullptr, 0.0, cone_out,
                                        ostride, active),
                "cxb_small_prepare_multi");
  }
  CudaCheck(cudaEventRecord(ev_end, cs), "cudaEventRecord");
  d.ctx.Synchronize();
  float total = 0;
  cudaEventElapsedTime(&total, ev_begin, ev_end);
  total_ms_ = total;
  cudaEventDestroy(ev_begin);
  cudaEventDestroy(ev_end);
  cudaEventDestroy(ev_a);
  cudaEventDestroy(ev_b);
  return solved_count;
}

}  // namespace conex
