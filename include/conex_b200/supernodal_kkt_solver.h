// Multi-supernode (chordal-sparse) KKT solver — SURVEY.md §8 row f4. Counterpart of the reference's
// SupernodalKKTSolver / clique_ordering / TriangularMatrixWorkspace / BlockCholeskyInPlace
// (kkt_solver.cc, clique_ordering.cc:111-343, triangular_matrix_workspace.cc:37-159,
// block_triangular_operations.cc:114-219), re-designed as a multifrontal method whose fronts are
// dense column blocks in HBM factored by the same DMMA kernels as the dense solver.
//
// Symbolic step (host, once per Initialize; AnalyzeCliques): the cones' variable sets ("cliques") are
// joined into a maximum-weight spanning forest (weight = size of the intersection, like the reference's
// PickCliqueOrder); every variable is eliminated at the lowest common ancestor of the cliques that
// contain it and is carried as a separator variable by every node on the paths up to there — the fill
// that makes the forest a tree decomposition whatever the cliques look like (reference FillIn). Node k
// then owns the dense front F_k = H[front_k, supernode_k], (s_k + p_k) x s_k column-major, rows =
// supernode variables followed by separator variables in elimination order.
//
// Numeric step (device): cones scatter their G into the fronts through precomputed destination lists;
// in post-order each front is partially factored (L11, L21 = A21 L11^{-T}), its Schur complement
// L21 L21^T is formed by one DMMA GEMM and subtracted from the ancestors' fronts through another
// destination list; the solves walk the same order with one substitution sweep and one L21 product per
// front. One stream, fixed order: deterministic. Elimination order differs from the reference's, which
// only reorders the summation (same factor up to rounding, same y).
#pragma once
#include <list>
#include <vector>

#include "cone_program.h"

namespace conex {

struct SupernodalStructure {
  int N = 0;
  // nodes in elimination (post-) order; empty supernodes are dropped
  std::vector<std::vector<int>> supernodes;   // original variable indices, eliminated in this order
  std::vector<std::vector<int>> separators;   // original variable indices, sorted by elimination position
  std::vector<int> parent;                    // node index of the node owning the first separator variable, -1: root
  std::vector<int> position;                  // position[v] = place of variable v in the elimination order
  std::vector<int> node_of;                   // node_of[v] = node that eliminates v
  double factor_flops = 0;                    // sum over fronts of the partial factorisation + Schur update
  double dense_flops = 0;                     // N^3 / 3
};

// Pure host logic (tests/test_supernodal.py drives it through CONEXB200_SupernodalAnalysis).
// Variables that appear in no clique become singleton supernodes at the end.
SupernodalStructure AnalyzeCliques(int N, const std::vector<std::vector<int>>& cliques);

class SupernodalKKTSolver : public KKTSolver {
 public:
  SupernodalKKTSolver(DeviceContext* ctx, int N, SupernodalStructure structure);
  ~SupernodalKKTSolver() override;
  void Bind(std::list<Container>* eqs) override;
  void Assemble() override;
  bool Factor() override;
  void SolveInPlace(Ref* b) const override;
  Ref KKTMatrix() const override;
  void SetSolverMode(int mode) override { mode_ = mode; }
  void SetIterativeRefinementIterations(int x) override { iterative_refinement_iterations_ = x; }
  void SetNumberOfMultipliers(int n) override { num_dual_ = n; }
  int NumberOfSupernodes() const override { return static_cast<int>(st_.supernodes.size()); }
  bool factorization_regularized() const { return factorization_regularized_; }

 private:
  struct Front {
    int s = 0, p = 0;        // supernode / separator sizes
    long offset = 0;         // of F_k in fronts_
    int first = 0;           // elimination position of its first variable
    long update_offset = 0;  // of its Schur-update destination list in update_idx_
    long sep_offset = 0;     // of its separator positions in sep_pos_
  };
  // Offset in fronts_ of entry (u, v) of the KKT matrix (original indices), -1 if it lies outside
  // the symbolic pattern (cannot happen for pairs inside one clique).
  long Locate(int u, int v) const;

  DeviceContext* ctx_;
  int N_;
  SupernodalStructure st_;
  std::vector<Front> fronts_meta_;
  std::vector<std::vector<int>> front_rows_;  // per node: elimination positions of its rows, ascending
  DeviceBuffer<double> fronts_;               // all F_k
  DeviceBuffer<long> update_idx_;             // per front: destinations of the lower triangle of L21 L21^T
  DeviceBuffer<int> sep_pos_;                 // per front: elimination positions of the separator rows
  DeviceBuffer<int> perm_;                    // perm_[position] = original variable
  mutable DeviceBuffer<double> schur_;        // largest p_k x p_k
  // Leaves of the assembly tree do not depend on each other: they are factored on side streams
  // ("lanes", round-robin) while the main stream applies their Schur updates in node order.
  struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t factored = nullptr, scattered = nullptr;
    DeviceBuffer<double> schur;
  };
  std::vector<Lane> lanes_;
  cudaEvent_t assembled_ = nullptr;
  std::vector<int> leaves_;                   // node indices without children, ascending
  std::vector<char> is_leaf_;
  void EnqueueLeaf(size_t leaf_number, int* info);
  // LDL^T over the fronts (programs with equality multipliers; reference BlockLDLTInPlace,
  // block_triangular_operations.cc:315-349): every supernode's diagonal block is factored as P^T L S L^T P with the
  // pivot order Eigen::RLDLT derives from the block's diagonal at that point, S = diag(+-1), regularised pivots as in
  // RLDLT.h:378-389. The permuted, factored fronts live in fronts_p_ (the Schur updates still go to fronts_).
  bool FactorLDLT();
  void SolveLDLT(double* y) const;
  DeviceBuffer<double> fronts_p_, signs_, ldlt_work_, gs_, diag_;
  DeviceBuffer<int> pivots_;                  // per supernode (at its first elimination position): its pivot order
  std::vector<int> host_pivots_;
  bool ldlt_factored_ = false;
  bool factorization_regularized_ = false;
  mutable DeviceBuffer<double> x_;            // right-hand side in elimination order
  mutable DeviceBuffer<double> dense_;        // KKTMatrix() export
  std::vector<DeviceBuffer<long>> cone_idx_;  // per cone: destinations of the lower triangle of its G
  std::list<Container>* eqs_ = nullptr;
  long total_ = 0;
  int mode_ = CONEX_LLT_FACTORIZATION;
  int iterative_refinement_iterations_ = 0;
  int num_dual_ = 0;
};

}  // namespace conex
