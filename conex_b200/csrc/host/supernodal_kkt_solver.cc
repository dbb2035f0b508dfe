#include "supernodal_kkt_solver.h"

#include <algorithm>
#include <map>
#include <numeric>
#include <stdexcept>

#include "../../../include/conex_b200_device.h"

namespace conex {

// ================================================================================================
// Symbolic analysis (host)
// ================================================================================================

SupernodalStructure AnalyzeCliques(int N, const std::vector<std::vector<int>>& cliques_in) {
  SupernodalStructure out;
  out.N = N;
  out.dense_flops = static_cast<double>(N) * N * N / 3.0;
  out.position.assign(N, -1);
  out.node_of.assign(N, -1);

  // sorted, duplicate-free, non-empty cliques
  std::vector<std::vector<int>> cliques;
  for (const auto& c : cliques_in) {
    std::vector<int> s(c);
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    for (int v : s) {
      if (v < 0 || v >= N) throw std::runtime_error("conex-b200: clique variable out of range");
    }
    if (!s.empty()) cliques.push_back(std::move(s));
  }
  const int q = static_cast<int>(cliques.size());
  std::vector<std::vector<int>> cliques_of(N);
  for (int c = 0; c < q; c++) {
    for (int v : cliques[c]) cliques_of[v].push_back(c);
  }

  // ---- maximum-weight spanning forest of the clique graph (weight = |C_a n C_b|) -----------------
  // A variable shared by t cliques contributes to t (t - 1) / 2 pairs; beyond kAllPairs cliques only
  // the t - 1 pairs with the first of them are counted (a star: it still connects every clique that
  // holds the variable and keeps the tree shallow; any spanning forest yields a valid structure — the
  // fill step below does not rely on the weights), so that the analysis stays linear in the input when
  // thousands of cones share variables.
  constexpr size_t kAllPairs = 64;
  std::map<std::pair<int, int>, int> weight;
  for (int v = 0; v < N; v++) {
    const auto& cs = cliques_of[v];
    if (cs.size() <= kAllPairs) {
      for (size_t i = 0; i < cs.size(); i++) {
        for (size_t j = i + 1; j < cs.size(); j++) weight[{cs[i], cs[j]}]++;
      }
    } else {
      for (size_t i = 1; i < cs.size(); i++) weight[{cs[0], cs[i]}]++;
    }
  }
  struct Edge {
    int w, a, b;
  };
  std::vector<Edge> edges;
  edges.reserve(weight.size());
  for (const auto& kv : weight) edges.push_back({kv.second, kv.first.first, kv.first.second});
  std::sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) {
    if (x.w != y.w) return x.w > y.w;
    if (x.a != y.a) return x.a < y.a;
    return x.b < y.b;
  });
  std::vector<int> uf(q);
  std::iota(uf.begin(), uf.end(), 0);
  auto find = [&](int x) {
    while (uf[x] != x) {
      uf[x] = uf[uf[x]];
      x = uf[x];
    }
    return x;
  };
  std::vector<std::vector<int>> adj(q);
  for (const auto& e : edges) {
    const int ra = find(e.a), rb = find(e.b);
    if (ra == rb) continue;
    uf[ra] = rb;
    adj[e.a].push_back(e.b);
    adj[e.b].push_back(e.a);
  }
  for (auto& a : adj) std::sort(a.begin(), a.end());

  // ---- root every tree at its largest clique; parents, depths, post-order ------------------------
  std::vector<int> comp_root(q, -1);
  for (int c = 0; c < q; c++) {
    const int r = find(c);
    if (comp_root[r] < 0 || cliques[c].size() > cliques[comp_root[r]].size()) comp_root[r] = c;
  }
  std::vector<int> parent(q, -2), depth(q, 0), post;
  post.reserve(q);
  for (int c = 0; c < q; c++) {
    if (comp_root[find(c)] != c) continue;  // c is not the root of its tree
    // iterative DFS, children in ascending order
    std::vector<std::pair<int, size_t>> stack;
    parent[c] = -1;
    stack.push_back({c, 0});
    while (!stack.empty()) {
      auto& top = stack.back();
      const int x = top.first;
      if (top.second < adj[x].size()) {
        const int y = adj[x][top.second++];
        if (y == parent[x]) continue;
        parent[y] = x;
        depth[y] = depth[x] + 1;
        stack.push_back({y, 0});
      } else {
        post.push_back(x);
        stack.pop_back();
      }
    }
  }

  // ---- where every variable is eliminated, and which nodes carry it as a separator variable ------
  auto lca = [&](int a, int b) {
    while (a != b) {
      if (depth[a] < depth[b]) std::swap(a, b);
      a = parent[a];
    }
    return a;
  };
  std::vector<std::vector<int>> super(q), sep(q);
  std::vector<int> mark(q, -1);
  for (int v = 0; v < N; v++) {
    const auto& cs = cliques_of[v];
    if (cs.empty()) continue;
    int top = cs[0];
    for (size_t i = 1; i < cs.size(); i++) top = lca(top, cs[i]);
    super[top].push_back(v);
    mark[top] = v;
    for (int c : cs) {
      for (int x = c; mark[x] != v; x = parent[x]) {
        mark[x] = v;
        sep[x].push_back(v);
      }
    }
  }

  // ---- elimination order: post-order over the nodes, variables ascending inside a supernode ------
  int pos = 0;
  std::vector<int> node_index(q, -1);
  for (int c : post) {
    if (super[c].empty()) continue;  // pass-through node: nothing to eliminate here
    node_index[c] = static_cast<int>(out.supernodes.size());
    for (int v : super[c]) {
      out.position[v] = pos++;
      out.node_of[v] = node_index[c];
    }
    out.supernodes.push_back(super[c]);
    out.separators.push_back(sep[c]);
  }
  for (int v = 0; v < N; v++) {
    if (out.position[v] >= 0) continue;  // in no clique: a singleton supernode (zero pivot ahead)
    out.position[v] = pos++;
    out.node_of[v] = static_cast<int>(out.supernodes.size());
    out.supernodes.push_back({v});
    out.separators.push_back({});
  }
  const int nodes = static_cast<int>(out.supernodes.size());
  out.parent.assign(nodes, -1);
  for (int k = 0; k < nodes; k++) {
    auto& s = out.separators[k];
    std::sort(s.begin(), s.end(), [&](int a, int b) { return out.position[a] < out.position[b]; });
    if (!s.empty()) out.parent[k] = out.node_of[s.front()];
    const double sk = static_cast<double>(out.supernodes[k].size()), pk = static_cast<double>(s.size());
    out.factor_flops += sk * sk * sk / 3.0 + sk * sk * pk + sk * pk * pk;
  }
  return out;
}

// ================================================================================================
// Numeric part (device)
// ================================================================================================

SupernodalKKTSolver::SupernodalKKTSolver(DeviceContext* ctx, int N, SupernodalStructure structure)
    : ctx_(ctx), N_(N), st_(std::move(structure)) {
  const int nodes = static_cast<int>(st_.supernodes.size());
  fronts_meta_.resize(nodes);
  front_rows_.resize(nodes);
  long offset = 0, upd = 0, sepo = 0;
  size_t max_p = 0;
  std::vector<int> perm(N_), sep_pos;
  for (int k = 0; k < nodes; k++) {
    Front& f = fronts_meta_[k];
    f.s = static_cast<int>(st_.supernodes[k].size());
    f.p = static_cast<int>(st_.separators[k].size());
    f.offset = offset;
    f.first = st_.position[st_.supernodes[k].front()];
    f.update_offset = upd;
    f.sep_offset = sepo;
    offset += static_cast<long>(f.s + f.p) * f.s;
    upd += static_cast<long>(f.p) * (f.p + 1) / 2;
    sepo += f.p;
    max_p = std::max<size_t>(max_p, f.p);
    auto& rows = front_rows_[k];
    for (int v : st_.supernodes[k]) {
      rows.push_back(st_.position[v]);
      perm[st_.position[v]] = v;
    }
    for (int v : st_.separators[k]) {
      rows.push_back(st_.position[v]);
      sep_pos.push_back(st_.position[v]);
    }
  }
  total_ = offset;
  fronts_.Resize(static_cast<size_t>(std::max<long>(total_, 1)));
  schur_.Resize(std::max<size_t>(max_p * max_p, 1));
  x_.Resize(static_cast<size_t>(std::max(N_, 1)));
  perm_.Resize(static_cast<size_t>(std::max(N_, 1)));
  // Uploads are stream-ordered on the program's stream (pageable sources are staged before the call
  // returns): the kernels that read these lists run on that stream, which does not synchronise with
  // the legacy default stream.
  cudaStream_t stream = ctx_->cuda_stream();
  CudaCheck(cudaMemcpyAsync(perm_.get(), perm.data(), sizeof(int) * N_, cudaMemcpyHostToDevice, stream),
            "upload of the order");
  sep_pos_.Resize(std::max<size_t>(sep_pos.size(), 1));
  if (!sep_pos.empty()) {
    CudaCheck(cudaMemcpyAsync(sep_pos_.get(), sep_pos.data(), sizeof(int) * sep_pos.size(), cudaMemcpyHostToDevice,
                              stream),
              "upload of the separator positions");
  }
  // Schur-update destinations: the pair (a >= b) of front k's separator lives in the front that
  // eliminates the earlier of the two variables — an ancestor of k.
  std::vector<long> update(static_cast<size_t>(std::max<long>(upd, 1)), -1);
  for (int k = 0; k < nodes; k++) {
    const auto& s = st_.separators[k];
    long e = fronts_meta_[k].update_offset;
    for (size_t b = 0; b < s.size(); b++) {
      for (size_t a = b; a < s.size(); a++) {
        const long d = Locate(s[a], s[b]);
        if (d < 0) throw std::runtime_error("conex-b200: supernodal structure is not closed under elimination");
        update[e++] = d;
      }
    }
  }
  // leaves of the assembly tree and the lanes that factor them concurrently
  is_leaf_.assign(nodes, 1);
  for (int k = 0; k < nodes; k++) {
    if (st_.parent[k] >= 0) is_leaf_[st_.parent[k]] = 0;
  }
  for (int k = 0; k < nodes; k++) {
    if (is_leaf_[k] && fronts_meta_[k].p > 0) leaves_.push_back(k);  // isolated roots stay on the main stream
    else is_leaf_[k] = 0;
  }
  if (leaves_.size() >= 2) {
    lanes_.resize(std::min<size_t>(4, leaves_.size()));
    size_t leaf_p = 0;
    for (int k : leaves_) leaf_p = std::max<size_t>(leaf_p, fronts_meta_[k].p);
    for (auto& l : lanes_) {
      CudaCheck(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking), "cudaStreamCreate");
      CudaCheck(cudaEventCreateWithFlags(&l.factored, cudaEventDisableTiming), "cudaEventCreate");
      CudaCheck(cudaEventCreateWithFlags(&l.scattered, cudaEventDisableTiming), "cudaEventCreate");
      l.schur.Resize(std::max<size_t>(leaf_p * leaf_p, 1));
    }
    CudaCheck(cudaEventCreateWithFlags(&assembled_, cudaEventDisableTiming), "cudaEventCreate");
  } else {
    for (int k : leaves_) is_leaf_[k] = 0;
    leaves_.clear();
  }
  update_idx_.Resize(update.size());
  CudaCheck(cudaMemcpyAsync(update_idx_.get(), update.data(), sizeof(long) * update.size(), cudaMemcpyHostToDevice,
                            stream),
            "upload of the update lists");
  ctx_->Synchronize();
}

SupernodalKKTSolver::~SupernodalKKTSolver() {
  for (auto& l : lanes_) {
    if (l.factored) cudaEventDestroy(l.factored);
    if (l.scattered) cudaEventDestroy(l.scattered);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  if (assembled_) cudaEventDestroy(assembled_);
}

long SupernodalKKTSolver::Locate(int u, int v) const {
  int pu = st_.position[u], pv = st_.position[v];
  if (pu > pv) std::swap(pu, pv);  // pu is eliminated first: the entry sits in its column
  const int k = st_.node_of[pu == st_.position[u] ? u : v];
  const Front& f = fronts_meta_[k];
  const auto& rows = front_rows_[k];
  // supernode rows are consecutive positions, separator rows ascending: the whole list is sorted
  const auto it = std::lower_bound(rows.begin(), rows.end(), pv);
  if (it == rows.end() || *it != pv) return -1;
  const long r = it - rows.begin();
  const long c = pu - f.first;
  return f.offset + c * (f.s + f.p) + r;
}

void SupernodalKKTSolver::Bind(std::list<Container>* eqs) {
  eqs_ = eqs;
  cone_idx_.clear();
  for (auto& c : *eqs) {
    const int mc = static_cast<int>(c.variables.size());
    c.identity_clique = false;
    c.direct_update = false;
    c.submatrix_data_.m_ = mc;
    c.submatrix_data_.residual_only_ = false;  // every cone assembles into its own G
    c.d_variables.Resize(mc);
    CudaCheck(cudaMemcpyAsync(c.d_variables.get(), c.variables.data(), sizeof(int) * mc, cudaMemcpyHostToDevice,
                              ctx_->cuda_stream()),
              "upload of clique indices");
    c.y_clique.Resize(mc);
    std::vector<long> idx(static_cast<size_t>(mc) * (mc + 1) / 2);
    size_t e = 0;
    for (int b = 0; b < mc; b++) {
      for (int a = b; a < mc; a++) {
        const long d = Locate(c.variables[a], c.variables[b]);
        if (d < 0) throw std::runtime_error("conex-b200: cone couples variables outside the symbolic pattern");
        idx[e++] = d;
      }
    }
    cone_idx_.emplace_back(std::max<size_t>(idx.size(), 1));
    if (!idx.empty()) {
      CudaCheck(cudaMemcpyAsync(cone_idx_.back().get(), idx.data(), sizeof(long) * idx.size(), cudaMemcpyHostToDevice,
                                ctx_->cuda_stream()),
                "upload of the assembly list");
    }
  }
  ctx_->Synchronize();
}

void SupernodalKKTSolver::Assemble() {
  void* s = ctx_->stream();
  ctx_->Zero(fronts_.get(), static_cast<size_t>(total_));
  size_t i = 0;
  for (auto& c : *eqs_) {
    ConstructSchurComplementSystem(&c.constraint, true, &c.submatrix_data_);
    const auto& G = c.submatrix_data_.G;
    DeviceCheck(cxb_scatter_lower_indexed(s, c.submatrix_data_.m_, G.data, G.ld, cone_idx_[i].get(), 1.0, fronts_.get()),
                "cxb_scatter_lower_indexed(assembly)");
    i++;
  }
}

bool SupernodalKKTSolver::Factor() {
  if (mode_ == CONEX_QR_FACTORIZATION) {
    throw std::runtime_error("conex-b200: the QR KKT mode is not implemented on the device");
  }
  if (num_dual_ > 0) return FactorLDLT();  // reference kkt_solver.cc:180-193
  ldlt_factored_ = false;
  void* s = ctx_->stream();
  cudaStream_t main = ctx_->cuda_stream();
  int* info = ctx_->flags();
  DeviceCheck(cxb_potrf_begin(s, info), "cxb_potrf_begin");
  if (!lanes_.empty()) {
    CudaCheck(cudaEventRecord(assembled_, main), "cudaEventRecord");
    for (auto& l : lanes_) CudaCheck(cudaStreamWaitEvent(l.stream, assembled_, 0), "cudaStreamWaitEvent");
    for (size_t i = 0; i < lanes_.size(); i++) EnqueueLeaf(i, info);
  }
  size_t leaf_number = 0;
  for (size_t k = 0; k < fronts_meta_.size(); k++) {
    const Front& f = fronts_meta_[k];
    double* F = fronts_.get() + f.offset;
    const long ld = f.s + f.p;
    const double* schur = schur_.get();
    Lane* lane = nullptr;
    if (is_leaf_[k]) {
      // factored (or being factored) on its lane: only the update is applied here, in node order
      lane = &lanes_[leaf_number % lanes_.size()];
      CudaCheck(cudaStreamWaitEvent(main, lane->factored, 0), "cudaStreamWaitEvent");
      schur = lane->schur.get();
    } else {
      DeviceCheck(cxb_potrf_partial(s, f.s + f.p, f.s, F, ld, info), "cxb_potrf_partial");
      if (f.p == 0) continue;
      const double* L21 = F + f.s;
      DeviceCheck(cxb_dgemm(s, 0, 1, f.p, f.p, f.s, 1.0, L21, ld, 0, L21, ld, 0, 0.0, schur_.get(), f.p, 0, 1, 1),
                  "cxb_dgemm(Schur complement of a front)");
    }
    DeviceCheck(cxb_scatter_lower_indexed(s, f.p, schur, f.p, update_idx_.get() + f.update_offset, -1.0,
                                          fronts_.get()),
                "cxb_scatter_lower_indexed(update)");
    if (lane) {
      // the lane's scratch is free again: hand the lane its next leaf
      CudaCheck(cudaEventRecord(lane->scattered, main), "cudaEventRecord");
      if (leaf_number + lanes_.size() < leaves_.size()) EnqueueLeaf(leaf_number + lanes_.size(), info);
      leaf_number++;
    }
  }
  int host_info = 0;
  ctx_->DownloadInts(&host_info, info, 1);
  return host_info == 0;  // reference block_triangular_operations.cc:193-196
}

// Leaf number i (in node order) runs on lane i mod L: partial factorisation of its front and its
// Schur complement into the lane's scratch. Called only after the update of leaf i - L has been
// enqueued on the main stream (its `scattered` event orders the reuse of the scratch).
void SupernodalKKTSolver::EnqueueLeaf(size_t leaf_number, int* info) {
  Lane& lane = lanes_[leaf_number % lanes_.size()];
  const Front& f = fronts_meta_[leaves_[leaf_number]];
  double* F = fronts_.get() + f.offset;
  const long ld = f.s + f.p;
  void* ls = reinterpret_cast<void*>(lane.stream);
  if (leaf_number >= lanes_.size()) CudaCheck(cudaStreamWaitEvent(lane.stream, lane.scattered, 0), "cudaStreamWaitEvent");
  DeviceCheck(cxb_potrf_partial(ls, f.s + f.p, f.s, F, ld, info), "cxb_potrf_partial");
  // splits = 1: the split-K workspace of the GEMM layer is shared by all streams
  DeviceCheck(cxb_dgemm_ex(ls, -1, 1, 0, 1, f.p, f.p, f.s, 1.0, F + f.s, ld, 0, F + f.s, ld, 0, 0.0, lane.schur.get(),
                           f.p, 0, 1, 1, 0, 0),
              "cxb_dgemm(Schur complement of a leaf front)");
  CudaCheck(cudaEventRecord(lane.factored, lane.stream), "cudaEventRecord");
}

// Every front in node order on the main stream: pivot order of the supernode from the current diagonal of its block
// (one small D2H per front), the front permuted into fronts_p_, signed partial factorisation, Schur complement
// (G S) G^T scattered into the ancestors. The regularised factorisation never fails (kkt_solver.cc:187-193).
bool SupernodalKKTSolver::FactorLDLT() {
  void* s = ctx_->stream();
  if (fronts_p_.size() == 0) {
    size_t max_rows = 1, max_s = 1, max_ps = 1;
    for (const Front& f : fronts_meta_) {
      max_rows = std::max<size_t>(max_rows, f.s + f.p);
      max_s = std::max<size_t>(max_s, f.s);
      max_ps = std::max<size_t>(max_ps, static_cast<size_t>(f.p) * f.s);
    }
    fronts_p_.Resize(static_cast<size_t>(std::max<long>(total_, 1)));
    signs_.Resize(static_cast<size_t>(std::max(N_, 1)));
    pivots_.Resize(static_cast<size_t>(std::max(N_, 1)));
    ldlt_work_.Resize(max_rows * 128 + max_rows);
    gs_.Resize(max_ps);
    diag_.Resize(max_s);
    host_pivots_.assign(std::max(N_, 1), 0);
  }
  int* info = ctx_->flags() + 2;
  DeviceCheck(cxb_ldlt_begin(s, info), "cxb_ldlt_begin");
  std::vector<double> d;
  for (const Front& f : fronts_meta_) {
    double* F = fronts_.get() + f.offset;
    double* Fp = fronts_p_.get() + f.offset;
    const long ld = f.s + f.p;
    DeviceCheck(cxb_copy_strided(s, f.s, F, ld + 1, diag_.get(), 1), "cxb_copy_strided");
    d.resize(f.s);
    ctx_->Download(d.data(), diag_.get(), f.s);
    const std::vector<int> order = RldltPivotOrderForTest(d);
    std::copy(order.begin(), order.end(), host_pivots_.begin() + f.first);
    // (pageable source: staged before the call returns)
    CudaCheck(cudaMemcpyAsync(pivots_.get() + f.first, host_pivots_.data() + f.first, sizeof(int) * f.s,
                              cudaMemcpyHostToDevice, ctx_->cuda_stream()),
              "upload of a supernode's pivot order");
    DeviceCheck(cxb_front_permute(s, f.s + f.p, f.s, F, ld, pivots_.get() + f.first, Fp, ld), "cxb_front_permute");
    DeviceCheck(cxb_ldlt_partial(s, f.s + f.p, f.s, Fp, ld, signs_.get() + f.first, ldlt_work_.get(), info),
                "cxb_ldlt_partial");
    if (f.p == 0) continue;
    const double* G = Fp + f.s;
    DeviceCheck(cxb_scale_columns(s, f.p, f.s, G, ld, signs_.get() + f.first, gs_.get(), f.p), "cxb_scale_columns");
    DeviceCheck(cxb_dgemm(s, 0, 1, f.p, f.p, f.s, 1.0, gs_.get(), f.p, 0, G, ld, 0, 0.0, schur_.get(), f.p, 0, 1, 1),
                "cxb_dgemm(Schur complement of a front, LDL^T)");
    DeviceCheck(cxb_scatter_lower_indexed(s, f.p, schur_.get(), f.p, update_idx_.get() + f.update_offset, -1.0,
                                          fronts_.get()),
                "cxb_scatter_lower_indexed(update)");
  }
  int host_info[2] = {0, 0};
  ctx_->DownloadInts(host_info, info, 2);
  factorization_regularized_ = host_info[1] != 0;
  ldlt_factored_ = true;
  return true;
}

// x = Pi^T M^{-T} S M^{-1} Pi b with Pi = diag(P_k) and M the block lower-triangular factor [L_k; G_k].
void SupernodalKKTSolver::SolveLDLT(double* y) const {
  void* s = ctx_->stream();
  double* tmp = diag_.get();  // >= the largest supernode
  DeviceCheck(cxb_gather_vec(s, N_, y, perm_.get(), x_.get()), "cxb_gather_vec");
  for (const Front& f : fronts_meta_) {
    const double* Fp = fronts_p_.get() + f.offset;
    const long ld = f.s + f.p;
    double* xk = x_.get() + f.first;
    DeviceCheck(cxb_permute_vec(s, f.s, pivots_.get() + f.first, xk, tmp, 0), "cxb_permute_vec");
    ctx_->CopyOnDevice(xk, tmp, f.s);
    DeviceCheck(cxb_trsv_lower(s, f.s, Fp, ld, xk, 0), "cxb_trsv_lower");
    DeviceCheck(cxb_front_forward(s, f.p, f.s, Fp + f.s, ld, xk, sep_pos_.get() + f.sep_offset, x_.get()),
                "cxb_front_forward");
  }
  DeviceCheck(cxb_apply_signs(s, N_, signs_.get(), x_.get()), "cxb_apply_signs");
  for (auto it = fronts_meta_.rbegin(); it != fronts_meta_.rend(); ++it) {
    const Front& f = *it;
    const double* Fp = fronts_p_.get() + f.offset;
    const long ld = f.s + f.p;
    double* xk = x_.get() + f.first;
    DeviceCheck(cxb_front_backward(s, f.p, f.s, Fp + f.s, ld, xk, sep_pos_.get() + f.sep_offset, x_.get()),
                "cxb_front_backward");
    DeviceCheck(cxb_trsv_lower(s, f.s, Fp, ld, xk, 1), "cxb_trsv_lower");
    DeviceCheck(cxb_permute_vec(s, f.s, pivots_.get() + f.first, xk, tmp, 1), "cxb_permute_vec");
    ctx_->CopyOnDevice(xk, tmp, f.s);
  }
  ctx_->Zero(y, N_);
  DeviceCheck(cxb_scatter_add_vec(s, N_, x_.get(), perm_.get(), y), "cxb_scatter_add_vec");
}

void SupernodalKKTSolver::SolveInPlace(Ref* b) const {
  if (iterative_refinement_iterations_ > 0) {
    throw std::runtime_error("conex-b200: iterative refinement is implemented for the dense KKT solver only");
  }
  void* s = ctx_->stream();
  if (ldlt_factored_) {
    for (int col = 0; col < b->cols; col++) SolveLDLT(b->col(col));
    return;
  }
  for (int col = 0; col < b->cols; col++) {
    double* y = b->col(col);
    DeviceCheck(cxb_gather_vec(s, N_, y, perm_.get(), x_.get()), "cxb_gather_vec");
    for (const Front& f : fronts_meta_) {  // L z = P b
      const double* F = fronts_.get() + f.offset;
      const long ld = f.s + f.p;
      DeviceCheck(cxb_trsv_lower(s, f.s, F, ld, x_.get() + f.first, 0), "cxb_trsv_lower");
      DeviceCheck(cxb_front_forward(s, f.p, f.s, F + f.s, ld, x_.get() + f.first, sep_pos_.get() + f.sep_offset,
                                    x_.get()),
                  "cxb_front_forward");
    }
    for (auto it = fronts_meta_.rbegin(); it != fronts_meta_.rend(); ++it) {  // L^T x = z
      const Front& f = *it;
      const double* F = fronts_.get() + f.offset;
      const long ld = f.s + f.p;
      DeviceCheck(cxb_front_backward(s, f.p, f.s, F + f.s, ld, x_.get() + f.first, sep_pos_.get() + f.sep_offset,
                                     x_.get()),
                  "cxb_front_backward");
      DeviceCheck(cxb_trsv_lower(s, f.s, F, ld, x_.get() + f.first, 1), "cxb_trsv_lower");
    }
    ctx_->Zero(y, N_);
    DeviceCheck(cxb_scatter_add_vec(s, N_, x_.get(), perm_.get(), y), "cxb_scatter_add_vec");
  }
}

Ref SupernodalKKTSolver::KKTMatrix() const {
  // dense copy of the assembled fronts in the original variable order (export for the tests)
  std::vector<double> host(static_cast<size_t>(std::max<long>(total_, 1)));
  ctx_->Download(host.data(), fronts_.get(), static_cast<size_t>(total_));
  const long ld = WorkspaceSchurComplement::AugLd(N_);
  std::vector<double> dense(static_cast<size_t>(ld) * N_, 0.0);
  std::vector<int> var_at(N_);
  for (int v = 0; v < N_; v++) var_at[st_.position[v]] = v;
  for (size_t k = 0; k < fronts_meta_.size(); k++) {
    const Front& f = fronts_meta_[k];
    const auto& rows = front_rows_[k];
    for (int c = 0; c < f.s; c++) {
      for (int r = c; r < f.s + f.p; r++) {
        const int u = var_at[rows[r]], v = var_at[f.first + c];
        const double val = host[f.offset + static_cast<long>(c) * (f.s + f.p) + r];
        dense[static_cast<size_t>(std::min(u, v)) * ld + std::max(u, v)] = val;
      }
    }
  }
  dense_.Resize(dense.size());
  CudaCheck(cudaMemcpyAsync(dense_.get(), dense.data(), sizeof(double) * dense.size(), cudaMemcpyHostToDevice,
                            ctx_->cuda_stream()),
            "upload of the dense export");
  ctx_->Synchronize();
  return Ref(dense_.get(), N_, N_, ld);
}

}  // namespace conex
