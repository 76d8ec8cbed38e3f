// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). Parity: unpinned by the reference.
//
// Restatement of the slice of g2o that NR-SLAM's drivers execute:
//   SparseOptimizer::initializeOptimization / optimize / update / push / pop
//       third_party/g2o/g2o/core/sparse_optimizer.cpp:171-285,392-470,548-627
//   OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale
//       third_party/g2o/g2o/core/optimization_algorithm_levenberg.cpp:57-174
//   BlockSolver::buildStructure / buildSystem / setLambda / restoreDiagonal / solve (non-Schur branch)
//       third_party/g2o/g2o/core/block_solver.hpp:108-159,329-341,495-603
//   BaseFixedSizedEdge::constructQuadraticForm      core/base_fixed_sized_edge.hpp:49-133
//   LinearSolverEigen (SimplicialLLT, exact sparse Cholesky)   solvers/eigen/linear_solver_eigen.h:92-188
//   LinearSolverDense (dense LDLT of the 6x6)                  solvers/dense/linear_solver_dense.h:56-104
// plus NR-SLAM's vertex/edge types (modules/optimization/*.cc, cited at each edge below).
//
// Eigen itself is not vendored in /root/reference (README.md:49-50 asks for ">= 3.1.0", unpinned). Its
// SimplicialLLT is an up-looking sparse Cholesky with an AMD ordering; any exact SPD solve agrees to
// rounding, so this file uses an up-looking LL^T (the algorithm published as CSparse cs_chol, Davis
// 2006, which SimplicialLLT follows) with a minimum-degree ordering on the block graph.
#pragma once
#include <array>
#include <vector>

#include "orc_math.h"

namespace orc {

enum VertexType { V_POSE = 0, V_POINT = 1 };

struct Vertex {
  int type = V_POINT;
  int dim = 3;
  bool fixed = false;
  SE3 pose{{0, 0, 0, 1}, {0, 0, 0}};
  double x[3] = {0, 0, 0};
  int hidx = -1;  // index among non-fixed active vertices, -1 otherwise
  int col = -1;   // first scalar column in the Hessian
  std::vector<std::array<double, 7>> stack;
  double A[36];   // diagonal Hessian block (dim x dim, row-major)
  double b[6];
};

enum EdgeType {
  E_REPROJ_ONLY_POSE = 0,   // optimization/reprojection_error_only_pose.cc:50-76
  E_REPROJ_DEFORM = 1,      // optimization/reprojection_error_with_deformation.cc:37-68
  E_REPROJ_BA = 2,          // optimization/reprojection_error.cc:32-64
  E_SPATIAL_DEFORM = 3,     // optimization/spatial_regularizer_with_deformation.cc:36-49
  E_POSITION_DEFORM = 4,    // optimization/position_regularizer_with_deformation.cc:31-57
  E_SPATIAL_FIXED = 5,      // optimization/spatial_regularizer_fixed.cc:32-43
  E_POSITION_BA = 6,        // optimization/position_regularizer.cc:32-61 (quirk E1)
  E_DAMPER_BA = 7,          // optimization/spatial_regularizer.cc:32-59
  E_REPROJ_ONLY_DEFORMATION = 8,  // optimization/reprojection_error_only_deformation.cc:33-39 — NO analytic Jacobian
                                  // (.h:40 commented out): g2o differentiates numerically, base_fixed_sized_edge.hpp:160-199
  E_SPATIAL_OBS = 9         // optimization/spatial_regularizer_with_observation.cc:33-51 (Jacobians +-w I, rotations ignored)
};

struct Edge {
  int type = 0;
  int nv = 1;
  int v[4] = {-1, -1, -1, -1};
  int dim = 2;
  int level = 0;
  double info = 1.0;    // every NR-SLAM edge uses a scalar multiple of the identity
  double delta = -1.0;  // Huber delta; <= 0 -> no robust kernel
  double meas[3] = {0, 0, 0};
  double Xw[3] = {0, 0, 0};            // landmark_world_ (reprojection edges with a fixed rest position)
  double rest1[3] = {0, 0, 0}, rest2[3] = {0, 0, 0};
  double weight = 1.0, k = 1.0;
  int ref_vertex = -1;                 // SpatialRegularizerFixed::flow_fixed (read live, no Jacobian)
  SE3 Ta{{0, 0, 0, 1}, {0, 0, 0}}, Tb{{0, 0, 0, 1}, {0, 0, 0}};  // E_SPATIAL_OBS: current_/next_world_transform_camera_
  double err[3] = {0, 0, 0};           // _error, updated only by compute_error (stale after a pop, like g2o)
  int hb[6] = {-1, -1, -1, -1, -1, -1};
  bool hbT[6] = {false, false, false, false, false, false};
};

struct OffBlock {
  int i, j;          // hessian indices, i < j
  int rows, cols;
  double m[36];
};

struct LMStats {
  int iterations = 0;        // LM iterations run (calls of solve())
  int trials = 0;            // damped solves
  int chol_fail = 0;
  double lambda = 0;
  double chi2_init = 0, chi2_final = 0;
  std::vector<double> chi2_trace;  // accepted chi2 after each LM iteration
  double t_order = 0, t_factor = 0, t_build = 0;  // seconds
};

class SparseChol;

class Optimizer {
 public:
  Optimizer(const Camera& cam, bool dense_solver) : cam_(cam), dense_(dense_solver) {}
  ~Optimizer();
  int add_vertex(const Vertex& v) { vertices.push_back(v); return (int)vertices.size() - 1; }
  int add_edge(const Edge& e) { edges.push_back(e); return (int)edges.size() - 1; }

  bool initialize_optimization(int level);
  int optimize(int iterations);

  void compute_error(Edge& e) const;
  double chi2(const Edge& e) const { double s = 0; for (int i = 0; i < e.dim; i++) s += e.err[i] * e.err[i]; return s * e.info; }
  // Jacobians: J[n] is dim x vdim(n) row-major
  void linearize(const Edge& e, double J[4][18]) const;

  std::vector<Vertex> vertices;
  std::vector<Edge> edges;
  LMStats stats;
  // experimentation knobs (never used by the parity drivers)
  bool use_pcg = false;
  double pcg_tol = 1e-10;
  int pcg_max_iter = 5000;
  long pcg_iters_total = 0;

 private:
  bool all_vertices_fixed(const Edge& e) const;
  void compute_active_errors();
  double active_robust_chi2() const;
  void build_structure();
  void build_system();
  bool solve_linear(double lambda);
  bool solve_pcg(double lambda);
  void update(const std::vector<double>& dx);
  void push();
  void pop();
  void discard_top();
  int lm_solve(int iteration);

  Camera cam_;
  bool dense_;
  std::vector<int> active_edges_, active_vertices_, index_mapping_;
  std::vector<OffBlock> blocks_;
  int n_scalar_ = 0;
  std::vector<double> b_, x_;
  double lambda_ = -1, ni_ = 2;
  SparseChol* chol_ = nullptr;
  bool structure_dirty_ = true;
};

// KAT hook: solve A x = b, A SPD given by scalar triplets of its upper triangle (block = block size used for
// the ordering). Returns 0 on success.
int sparse_solve_triplets(int n, int block, int nnz, const int* rows, const int* cols, const double* vals,
                          const double* b, double* x);

}  // namespace orc
