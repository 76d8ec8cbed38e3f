// nrs_engine.cuh — problem description handed to the persistent Levenberg–Marquardt kernel (nrs_engine.cu).
//
// One launch executes a whole driver "program" (reset / optimize(n) / re-level / final chi2) on the device:
// the g2o control flow of third_party/g2o/g2o/core/optimization_algorithm_levenberg.cpp:57-151 and the
// round / outlier logic of modules/optimization/g2o_optimization.cc run without a host round trip.
//
// Unknowns: F camera poses (6-dof, left-multiplicative exp update) and V point vertices (3-dof, additive).
// Every point vertex owns at most one reprojection edge (true for all three reference drivers: tracking has one
// deformation vertex per observation, BA one landmark vertex per (keyframe, landmark)), so the whole system is
// traversed "one row per point": the point's reprojection edge plus the regulariser edges incident to it.
#pragma once
#include <stdint.h>

#include "nrs_math.cuh"

namespace nrs {

enum Op : int {
  OP_RESET = 1,          // estimates <- seeds (g2o_optimization.cc:108-110,341-349)
  OP_CLEAR_LEVELS = 2,   // all edges to level 0
  OP_OPTIMIZE = 3,       // SparseOptimizer::optimize(arg) with the LM algorithm
  OP_RELEVEL_POSE = 4,   // g2o_optimization.cc:113-140 (stale-error semantics for inlier edges)
  OP_RELEVEL_DEFORM = 5, // g2o_optimization.cc:352-394
  OP_FINAL_CHI2 = 6      // fresh reprojection chi2 per point (g2o_optimization.cc:419-424)
};

enum SpringKind : int {
  SPRING_NONE = 0,
  SPRING_DEFORM = 1,  // PositionRegularizerWithDeformation: exact Jacobian, Huber
  SPRING_BA = 2       // PositionRegularizer: Jacobian quirk, no robust kernel
};

constexpr int kMaxOps = 16;
constexpr int kMaxWorld = 8;     // GPUs of one NVSwitch domain a sharded BA can span
constexpr int kTrace = 64;
constexpr int kSlotVals = 8;     // doubles per CTA per grid reduction
constexpr int kChunkVals = 28;   // doubles per chunk partial (21 H_pp + 6 b_p, padded)
#ifndef NRS_TPR
#define NRS_TPR 2
#endif
constexpr int kTPR = NRS_TPR;    // threads per point row inside the CG loop (a group of adjacent lanes, 2 or 4)
constexpr int kMaxRows = 128;    // point rows per chunk
constexpr int kMaxBlock = kTPR * kMaxRows;  // threads per CTA (544)
constexpr int kPrecBlock = 16;   // rows per dense preconditioner block (resident mode)
constexpr int kGatherVals = 12;  // doubles per CTA in the pushed reduction buffers of the cluster-native CG loop
constexpr int kCoarseN = 6 + 3 * 16;   // coarse unknowns: pose + one 3-dof aggregate per CTA of the cluster
constexpr int kCoarseS = 60;     // row stride of the coarse matrix (floats): float4 rows, conflict-free across lanes
constexpr int kRowBlk = 116;     // floats a CTA publishes for the coarse matrix: 16 x 6 (sym 3x3) + 18 (pose coupling) + 1 + pad

struct EngineStats {
  int lm_iterations, lm_trials, pcg_iterations, n_sweeps, n_chi2_passes, n_trace, pcg_fail, barriers;
  int xepochs, xfail;  // landmark-sharded runs: cross-device exchanges done by this launch / a peer never arrived
  double chi2_trace[kTrace];
  double lambda_final;
  long long prof[16];  // clock64 cycles per phase seen by CTA 0 / thread 0 (diagnostics)
};

struct Params {
  int F, V, P, D, n_chunks;
  int poses_fixed, points_fixed, spring_kind;
  Cam cam;
  double info_reproj, delta_reproj;    // delta <= 0: no robust kernel
  double info_spatial, delta_spatial;  // pair "spatial", damper and unary (fixed-reference) edges
  double info_spring, delta_spring, spring_k;
  float th2f, th3f;
  double lm_tau;
  int lm_max_trials;
  double pcg_tol;
  int pcg_max_iter;
  int n_ops;
  int op[kMaxOps];
  int op_arg[kMaxOps];

  // execution mode (chosen by the host from the problem size, nrs_api.cu: plan_launch)
  int cluster_mode;   // 1: the grid is ONE thread-block cluster -> hardware cluster barrier instead of the atomics one
  int resident;       // 1: one chunk per CTA; Jacobians, edge coefficients and the CG vectors of the chunk live in
                      //    shared memory for the duration of a solve, only the exchanged vector goes through L2
  int res_rows;       // shared-memory capacity: rows per chunk
  int res_inc;        // shared-memory capacity: pair incidences per chunk
  int block_prec;     // 1: dense kPrecBlock-row block-Jacobi preconditioner (resident mode only), else 3x3 blocks
  int wide;           // 1: launch the two-CTAs-per-SM variant (large windows: cooperative grid, nothing resident)
  int no_dsmem;       // 1: keep the general CG loop (exchange through L2) even where the cluster-native loop applies
  // halo exchange of the cluster-native CG loop: after every z = M^-1 r the owner of a row PUSHES it into the halo
  // buffer of every chunk whose regulariser edges read it (remote shared-memory stores, no remote loads)
  int coarse;               // 1: two-level preconditioner in the cluster-native loop (one 3-dof aggregate per chunk + the pose)
  int halo_rows;            // shared-memory capacity: halo rows per chunk
  const int* inc_halo;      // [2P] per incidence: halo slot of the neighbour in this chunk's buffer, -1 if in-chunk
  const int* push_ptr;      // [n_chunks + 1]
  const int* push_row;      // row (global index) to push
  const int* push_dst;      // target chunk * 65536 + slot
  const int* xinc_ptr;      // [n_chunks + 1] per chunk: its incidences whose neighbour lives in another chunk ...
  const int* xinc_idx;      // ... as indices into the incidence arrays, ascending (coarse-matrix assembly)

  // poses: 7 doubles each (q xyzw, t)
  double* pose;
  const double* pose_seed;
  int seed_via_f32;         // 1: the seed is the fp64 result of an earlier launch and passes through Sophus::SE3f (fp32)
                            //    like the reference's Frame does between the two tracking drivers
  // point vertices, 4-double stride
  double* x;
  const double* x_seed;
  double* x_bak;
  const double* rest;       // rest position added to the estimate (tracking) — zeros in BA
  const int* pt_kf;         // pose slot of the point's reprojection edge, -1 if it has none
  const double* uv;         // [2V] measured pixel
  unsigned char* rp_level;  // [V] reprojection edge level
  double* rp_chi2;          // [V] chi2 of the reprojection edge at its last evaluation
  // pair edges (two point vertices): spatial (weight w >= 0, <0: none) and/or spring (d0)
  const int* pair_i;
  const int* pair_j;
  const double* pair_w;
  const double* pair_d0;
  unsigned char* sp_level;  // [P] level of the spatial edge
  double* pc;               // [4P] linearised coefficients: s, u[3]   (H_ij = -(s I + u u^T))
  double* pcc;              // [P]  spring residual term c (gradient)
  const int* inc_ptr;       // [V+1] CSR of pair incidences per point
  const int* inc_other;     // neighbour vertex
  const int* inc_ent;       // pair id * 2 + (1 if this vertex is the pair's second endpoint)
  const int* inc_row;       // the point row an incidence belongs to
  // damper edges (four point vertices: i_k, j_k, i_k', j_k')
  const int* dmp_v;         // [4D]
  const double* dmp_w;      // [D]
  double* dc;               // [4D] s, g[3]
  const int* dinc_ptr;      // [V+1]
  const int* dinc_ent;      // damper id * 4 + role
  // unary edges to a fixed reference value (SpatialRegularizerFixed)
  const int* un_ptr;        // [V+1]
  const double* un_w;       // [U]
  const int* un_ref;        // [U] vertex whose estimate is the reference value
  int unary_on;             // unary edges take part (lost-point stage only)
  const unsigned char* pt_fixed;  // [V] per-vertex setFixed(true), nullptr: none
  // work partition: chunk = contiguous rows of one pose slot, at most blockDim.x rows
  const int* chunk_kf;
  const int* chunk_begin;
  const int* chunk_end;
  const int* kf_chunk_ptr;  // [F+1]
  // work arrays
  double* jac;         // [20V] A(2x6) B(2x3) omega
  double* dg;          // [8V]  sym 3x3 diagonal block (6) + unary s
  double* bvec;        // [4V]
  double* minv;        // [8V]
  double* xcg;         // [4V]
  double* rvec;        // [4V]
  double* pvec;        // [4V]
  double* qvec;        // [4V]
  double* zvec;        // [4V] preconditioned residual — the one vector neighbours read during the CG loop
  double* chunk_part;  // [2][n_chunks * kChunkVals]
  double* slots;       // [2][G * kSlotVals]
  // wide CG loop (Engine<true>::pcg_wide): every CTA owns ONE contiguous row range of equal length, cut into segments
  // at the pose-slot boundaries; the pose partials of the matvec are reduced per segment, not per 128-row chunk
  int wide_prefetch;       // 1: L1 prefetch of the damper records in the matvec pass
  int n_wseg;
  const int* wseg_ptr;     // [G + 1] segments of CTA b
  const int* wseg_begin;   // [n_wseg] rows of the segment
  const int* wseg_end;
  const int* wseg_kf;      // [n_wseg] pose slot of the segment
  const int* kf_wseg_ptr;  // [F + 1] the segments of a pose slot are contiguous (rows are slot-major)
  double* wseg_part;       // [2][n_wseg * 8] pose partials of the CG matvec per segment
  double* wvec;            // [4V] row-local part of the next matvec ((lambda + s) z + B^T t), written where z is
  double* wrec;            // [2P * 4] s, u[3] of the pair edge in incidence order (expanded once per solve)
  double* wdrec;           // [4D * 4] per damper incidence, one 32-byte record: the three other vertices ordered
                           //          (+, -, -) as int32 and the coefficient
  unsigned long long* bar;
  EngineStats* stats;

  // ---- landmark-sharded BA over several GPUs (world > 1). Every rank owns the rows of its landmarks
  // and keeps read-only halo copies of the rows its regulariser edges reach on other ranks. The exchange buffers of
  // all ranks are peer-mapped (CUDA IPC): owners PUSH the halo values and their partial sums over NVLink, then signal.
  int world, rank;
  int xfused;                      // 1: grid reduction and exchange in one synchronisation (xreduce_impl)
  int xstride;                     // doubles per (parity, source rank) record of the reduction buffer: 8 + 27 F
  unsigned long long xepoch0;      // exchanges completed by earlier launches (the flags count up monotonically)
  unsigned long long xtimeout_ns;  // a peer that does not arrive within this time aborts the exchange (no hang)
  unsigned long long* xflag[kMaxWorld];  // rank r's flags [world], indexed by the signalling rank
  double* xred[kMaxWorld];         // rank r's reduction records [2][world][xstride]
  double* xz[kMaxWorld];           // rank r's z vector (its rows, then its halo rows) = that rank's zvec
  double* xx[kMaxWorld];           // rank r's estimates, same indexing                 = that rank's x
  const int* xp_ptr;               // [V+1] per owned row: the halo copies to refresh ...
  const int* xp_dst;               // ... as rank << 26 | row index on that rank
  const unsigned char* pair_cnt;   // [P] this rank counts the pair edge's chi2 (nullptr: every edge)
  const unsigned char* dmp_cnt;    // [D]
  int* xabort;
};

// Host-side launch. Returns cudaError_t as int. grid/block/smem chosen by the caller (plan_launch).
int launch_engine(const Params& p, int grid, int block, size_t smem, cudaStream_t stream);
// Shared memory needed for F poses (+ the resident chunk state when res_rows > 0).
size_t engine_smem_bytes(int F, int res_rows, int res_inc, int block_prec);
// Shared memory of the wide variant (no per-row CG records, H_pp aliases the linearisation's row records).
size_t engine_smem_bytes_wide(int F);
// Extra shared memory of the cluster-native loop: pushed halo rows and the coarse (aggregate) level.
size_t engine_smem_extra(int res_inc, int halo_rows, int coarse);
// Largest co-resident grid for a cooperative launch / largest cluster that can be scheduled (0 if none).
int engine_max_grid(int block, size_t smem, int wide = 0);
int engine_max_cluster(int block, size_t smem);

}  // namespace nrs
