// nrs_direct.cu — persistent Levenberg–Marquardt kernel of the per-frame pose + deformation tracking problem with an
// EXACT sparse LL^T solve of every damped system (sm_100a, cooperative grid of 2^depth CTAs).
//
// What it replaces (reference paths relative to /root/reference):
//   CameraPoseAndDeformationOptimization, main rounds          modules/optimization/g2o_optimization.cc:338-395
//   OptimizationAlgorithmLevenberg::solve                      third_party/g2o/g2o/core/optimization_algorithm_levenberg.cpp:57-174
//   BlockSolver::buildSystem / setLambda / solve               third_party/g2o/g2o/core/block_solver.hpp:329-341,495-603
//   LinearSolverEigen::solve (Eigen::SimplicialLLT)            third_party/g2o/g2o/solvers/eigen/linear_solver_eigen.h:92-136
//   the three edge types of the tracking graph (cited at each formula)
//
// Why a second engine: a tracking frame is ~35 damped solves of a 6006-unknown SPD system. The
// preconditioned-CG engine (nrs_engine.cu) needs ~90 dependent iterations per solve on ONE 16-CTA cluster; here the
// system is factorised exactly, like the reference does, by a multifrontal LL^T over a nested-dissection tree whose
// independent subtrees run on up to 128 SMs (nrs_direct_plan.h, nrs_direct_core.cuh). One launch still runs the
// whole driver program (reset / optimize(n) / re-level / final chi2) without a host round trip.
//
// Work distribution outside the solve: point row i is linearised by CTA (i mod G), a group of LPR adjacent lanes
// splitting the row's regulariser incidences; every pair edge is evaluated by both endpoints (no barrier between an
// edge pass and a row pass), its chi2 counted and its Hessian block published by the first endpoint only. Grid sums
// go through per-CTA slots added in CTA order, so every CTA derives bit-identical scalars and takes the same branches.
#include <cooperative_groups.h>
#include <float.h>
#include <stdio.h>

#include <cuda/ptx>

#include "nrs_direct.cuh"

namespace nrs {

namespace {

__device__ __forceinline__ unsigned long long d_ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void d_red_release_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 dld3(const double* base, int i) {  // written by other CTAs during the launch
  const double2 a = __ldcg(reinterpret_cast<const double2*>(base + 4 * (size_t)i));
  const double b = __ldcg(base + 4 * (size_t)i + 2);
  return D3{a.x, a.y, b};
}
__device__ __forceinline__ D3 dld3c(const double* base, int i) {  // read-only input
  const double2 a = __ldg(reinterpret_cast<const double2*>(base + 4 * (size_t)i));
  const double b = __ldg(base + 4 * (size_t)i + 2);
  return D3{a.x, a.y, b};
}
__device__ __forceinline__ void dst3(double* base, int i, const D3& v) {
  *reinterpret_cast<double2*>(base + 4 * (size_t)i) = make_double2(v.x, v.y);
  base[4 * (size_t)i + 2] = v.z;
}

constexpr int kDBlock = 256;
// shared-memory doubles (even) of the pivot inverses + look-ahead staging / of the unscaled panel rows
__host__ __device__ inline size_t direct_sw_doubles(int max_nv) {
  return 6 * (size_t)(((max_nv + 1) & ~1) + 1) + 20 * direct::kPanel + 2;
}
__host__ __device__ inline size_t direct_sv_doubles(int max_rows) {
  return ((size_t)max_rows * 3 * (3 * direct::kPanel + 1) + 2) & ~(size_t)1;
}
constexpr int kRec = 16;  // doubles per pose-block row record: A (12), omega, -, we0, we1

struct DirectEngine {
  const DirectParams& Q;
  const Params& P;
  double *s_pose, *s_pose_bak, *s_scal, *s_w, *s_v, *s_hpp, *s_red, *s_rec, *s_path, *s_z, *sp;
  uint64_t* s_mbar;    // mbarrier of the bulk copies (TMA) of the backward substitution
  unsigned mphase;
  int tid, G, cta;
  int rpc, lpr, slot, lane;  // rows per CTA, lanes per row, this thread's row slot / lane inside the row group
  unsigned long long gen, nsolve;
  double lambda, ni;
  int lm_iters, lm_trials, n_sweeps, n_chi2, n_trace, n_fail, fail_seen;
  long long prof[16];
#ifdef NRS_DIRECT_PLEV
  long long plev[32];  // per tree level: cycles in stage AB [d], stage C [8 + d], backward [16 + d]
#endif

  __device__ DirectEngine(const DirectParams& q, double* sm) : Q(q), P(q.P) {
    tid = threadIdx.x;
    G = gridDim.x;
    cta = blockIdx.x;
    double* p = sm;  // every piece has an even number of doubles: sp stays 16-byte aligned for the bulk copies
    s_pose = p; p += 8;
    s_pose_bak = p; p += 8;
    s_scal = p; p += 8;
    s_mbar = reinterpret_cast<uint64_t*>(p); p += 2;
    s_w = p; p += direct_sw_doubles(Q.max_nv);
    s_v = p; p += direct_sv_doubles(Q.max_rows);
    s_hpp = p; p += 28;
    s_red = p; p += 27 * 8;
    s_path = p; p += (Q.pl.max_path + 1) & ~1;
    s_z = p; p += (Q.scratch_z + 1) & ~1;
    sp = p;
    s_rec = sp;  // the pose-block row records of the linearisation live in the (then idle) panel buffer
    mphase = 0;
    if (tid == 0) {
      cuda::ptx::mbarrier_init(s_mbar, 1);
      cuda::ptx::fence_proxy_async();
    }
    __syncthreads();
    rpc = (Q.pl.V + G - 1) / G;
    lpr = 16;
    while (lpr > 1 && lpr * rpc > kDBlock) lpr >>= 1;
    slot = tid / lpr;
    lane = tid - slot * lpr;
    gen = 0;
    nsolve = 0;
    lambda = -1;
    ni = 2;
    lm_iters = lm_trials = n_sweeps = n_chi2 = n_trace = n_fail = 0;
    fail_seen = 0;
    for (int i = 0; i < 16; i++) prof[i] = 0;
#ifdef NRS_DIRECT_PLEV
    for (int i = 0; i < 32; i++) plev[i] = 0;
#endif
  }

  __device__ __forceinline__ void barrier() {
    gen++;
    const long long t0 = clock64();
    __syncthreads();
    if (tid == 0) {
      d_red_release_add_u64(P.bar, 1ULL);
      const unsigned long long target = gen * (unsigned long long)G;
      while (d_ld_acquire_u64(P.bar) < target) {
      }
    }
    __syncthreads();
    prof[12] += clock64() - t0;
  }

  // Barrier over the `count` CTAs that work on one tree node: arrivals are counted on the node's own counter, so
  // independent subtrees never wait for each other (only the LM-level reductions are grid-wide).
  __device__ __forceinline__ void team_barrier(unsigned long long* ctr, unsigned long long count) {
    const long long t0 = clock64();
    __syncthreads();
    if (tid == 0) {
      d_red_release_add_u64(ctr, 1ULL);
      const unsigned long long target = nsolve * count;
      while (d_ld_acquire_u64(ctr) < target) {
      }
    }
    __syncthreads();
    prof[12] += clock64() - t0;
  }

  // the row this thread group works on (-1: none)
  __device__ __forceinline__ int my_row() const {
    if (slot >= rpc) return -1;
    const int i = slot * G + cta;
    return i < Q.pl.V ? i : -1;  // unknown rows only (fixed rows sit behind them)
  }

  // sum over the lanes of a row group (fixed tree)
  __device__ __forceinline__ double group_sum(double v) const {
    for (int off = lpr >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }

  // Grid reduction of two per-thread values (sum, and sum or max): slots summed in CTA order. One grid barrier.
  __device__ __forceinline__ void grid_reduce2(double a, double b, bool bmax) {
    const int par = gen & 1;
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      const double o = __shfl_xor_sync(0xffffffffu, b, off);
      b = bmax ? fmax(b, o) : b + o;
    }
    const int warp = tid >> 5, ln = tid & 31;
    if (ln == 0) {
      s_red[2 * warp] = a;
      s_red[2 * warp + 1] = b;
    }
    __syncthreads();
    if (tid == 0) {
      double sa = 0, sb = bmax ? -DBL_MAX : 0.0;
      for (int w = 0; w < kDBlock / 32; w++) {
        sa += s_red[2 * w];
        sb = bmax ? fmax(sb, s_red[2 * w + 1]) : sb + s_red[2 * w + 1];
      }
      double2* o = reinterpret_cast<double2*>(Q.dslots + ((size_t)par * G + cta) * 2);
      *o = make_double2(sa, sb);
    }
    barrier();
    if (tid < 32) {
      double sa = 0, sb = bmax ? -DBL_MAX : 0.0;
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = min(ln + 32 * u, G - 1);
        v[u] = __ldcg(reinterpret_cast<const double2*>(Q.dslots + ((size_t)par * G + c) * 2));
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (ln + 32 * u < G) {
          sa += v[u].x;
          sb = bmax ? fmax(sb, v[u].y) : sb + v[u].y;
        }
      for (int off = 16; off > 0; off >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, off);
        const double o = __shfl_xor_sync(0xffffffffu, sb, off);
        sb = bmax ? fmax(sb, o) : sb + o;
      }
      if (ln == 0) {
        s_scal[0] = sa;
        s_scal[1] = sb;
      }
    }
    __syncthreads();
  }

  //   ReprojectionErrorWithDeformation::computeError  optimization/reprojection_error_with_deformation.cc:37-50
  __device__ __forceinline__ double reproj_error(int i, const D3& xi, double pc[3], double err[2]) {
    const D3 r = dld3c(P.rest, i);
    const double Xw[3] = {xi.x + r.x, xi.y + r.y, xi.z + r.z};
    pose_map(s_pose, Xw, pc);
    float u, v;
    project_f(P.cam, (float)pc[0], (float)pc[1], (float)pc[2], u, v);
    const double2 z = __ldg(reinterpret_cast<const double2*>(P.uv) + i);
    err[0] = z.x - (double)u;
    err[1] = z.y - (double)v;
    return (err[0] * err[0] + err[1] * err[1]) * P.info_reproj;
  }

  // One pair edge in its canonical orientation (i = first endpoint). Returns chi2 contribution (robustified).
  //   spatial : SpatialRegularizerWithDeformation   optimization/spatial_regularizer_with_deformation.cc:36-49
  //   spring  : PositionRegularizerWithDeformation  optimization/position_regularizer_with_deformation.cc:31-57
  template <bool LIN>
  __device__ __forceinline__ double pair_edge(int e, int i, int j, const D3& xi, const D3& xj, double& s, double u[3],
                                              double& c) {
    double chi = 0;
    s = 0;
    u[0] = u[1] = u[2] = 0;
    c = 0;
    const double w = __ldg(P.pair_w + e);
    if (w >= 0 && P.sp_level[e] == 0) {
      const double e0 = w * (xi.x - xj.x), e1 = w * (xi.y - xj.y), e2 = w * (xi.z - xj.z);
      const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
      double rho, drho;
      huber(c2, P.delta_spatial, rho, drho);
      chi += rho;
      s = drho * P.info_spatial * w * w;
    }
    {
      const D3 ri = dld3c(P.rest, i), rj = dld3c(P.rest, j);
      const double c1x = ri.x + xi.x, c1y = ri.y + xi.y, c1z = ri.z + xi.z;
      const double c2x = rj.x + xj.x, c2y = rj.y + xj.y, c2z = rj.z + xj.z;
      const double dx = c1x - c2x, dy = c1y - c2y, dz = c1z - c2z;
      const double dist = sqrt(dx * dx + dy * dy + dz * dz);
      const double d0 = __ldg(P.pair_d0 + e);
      const double err = P.spring_k * (dist - d0) / d0;
      const double ch = err * err * P.info_spring;
      double rho, drho;
      huber(ch, P.delta_spring, rho, drho);
      chi += rho;
      if (LIN) {
        const double aa = P.spring_k / (2 * d0 * dist);
        const double j0 = aa * (2 * c1x - 2 * c2x), j1 = aa * (2 * c1y - 2 * c2y), j2 = aa * (2 * c1z - 2 * c2z);
        const double sw = sqrt(drho * P.info_spring);
        u[0] = sw * j0;
        u[1] = sw * j1;
        u[2] = sw * j2;
        c = sw * err;
      }
    }
    return chi;
  }

  // ================================================================================================
  // Linearisation (LIN) or chi2 evaluation of the whole graph at the current estimate. Returns this thread's chi2
  // partial; LIN also returns the largest diagonal entry seen and publishes this CTA's pose-block partial.
  // ================================================================================================
  template <bool LIN>
  __device__ void graph_pass(double& chi, double& maxd) {
    const int i = my_row();
    double D[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    bool rec_written = false;
    if (i >= 0) {
      const D3 xi = dld3(P.x, i);
      if (lane == 0 && __ldg(P.pt_kf + i) >= 0 && P.rp_level[i] == 0) {
        double pc[3], err[2];
        const double c2 = reproj_error(i, xi, pc, err);
        double rho, drho;
        huber(c2, P.delta_reproj, rho, drho);
        chi += rho;
        P.rp_chi2[i] = c2;
        if (LIN) {
          // linearizeOplus: J_pose = -J_pi [ -[p]x | I ], J_point = -J_pi R
          //   optimization/reprojection_error_with_deformation.cc:52-68
          float Jf[6];
          projection_jacobian_f(P.cam, (float)pc[0], (float)pc[1], (float)pc[2], Jf);
          double Jp[6], A[12], B[6];
#pragma unroll
          for (int k = 0; k < 6; k++) Jp[k] = -(double)Jf[k];
          const double x = pc[0], y = pc[1], z = pc[2];
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const double a = Jp[r * 3], bb = Jp[r * 3 + 1], cc = Jp[r * 3 + 2];
            A[r * 6 + 0] = bb * (-z) + cc * y;
            A[r * 6 + 1] = a * z + cc * (-x);
            A[r * 6 + 2] = a * (-y) + bb * x;
            A[r * 6 + 3] = a;
            A[r * 6 + 4] = bb;
            A[r * 6 + 5] = cc;
          }
          const double omega = drho * P.info_reproj;
          const double we0 = -omega * err[0], we1 = -omega * err[1];
          double R[9];
          quat_to_R(s_pose, R);
#pragma unroll
          for (int r = 0; r < 2; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++)
              B[r * 3 + cc] = Jp[r * 3] * R[cc] + Jp[r * 3 + 1] * R[3 + cc] + Jp[r * 3 + 2] * R[6 + cc];
          D[0] = omega * (B[0] * B[0] + B[3] * B[3]);
          D[1] = omega * (B[0] * B[1] + B[3] * B[4]);
          D[2] = omega * (B[0] * B[2] + B[3] * B[5]);
          D[3] = omega * (B[1] * B[1] + B[4] * B[4]);
          D[4] = omega * (B[1] * B[2] + B[4] * B[5]);
          D[5] = omega * (B[2] * B[2] + B[5] * B[5]);
          b[0] = B[0] * we0 + B[3] * we1;
          b[1] = B[1] * we0 + B[4] * we1;
          b[2] = B[2] * we0 + B[5] * we1;
          if (Q.pl.np) {
            double* cp = Q.cpl + 18 * (size_t)i;
#pragma unroll
            for (int p = 0; p < 6; p++)
#pragma unroll
              for (int cc = 0; cc < 3; cc++) cp[3 * p + cc] = omega * (A[p] * B[cc] + A[6 + p] * B[3 + cc]);
            double* rec = s_rec + kRec * slot;
#pragma unroll
            for (int k = 0; k < 12; k++) rec[k] = A[k];
            rec[12] = omega;
            rec[14] = we0;
            rec[15] = we1;
            rec_written = true;
          }
        }
      }
      if (LIN && lane == 0 && !rec_written && Q.pl.np) {
        double* cp = Q.cpl + 18 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 18; k++) cp[k] = 0.0;
      }
      if (P.unary_on) {
        // SpatialRegularizerFixed  optimization/spatial_regularizer_fixed.cc:32-43 — the reference value is read
        // live from another (fixed) vertex and carries no Jacobian; the lanes of the row group split the edges
        for (int a = __ldg(P.un_ptr + i) + lane; a < __ldg(P.un_ptr + i + 1); a += lpr) {
          const double w = __ldg(P.un_w + a);
          const D3 rf = dld3(P.x, __ldg(P.un_ref + a));
          const double d0 = xi.x - rf.x, d1 = xi.y - rf.y, d2 = xi.z - rf.z;
          const double c2 = w * w * (d0 * d0 + d1 * d1 + d2 * d2) * P.info_spatial;
          double rho, drho;
          huber(c2, P.delta_spatial, rho, drho);
          chi += rho;
          if (LIN) {
            const double s = drho * P.info_spatial * w * w;
            D[0] += s;
            D[3] += s;
            D[5] += s;
            b[0] -= s * d0;
            b[1] -= s * d1;
            b[2] -= s * d2;
          }
        }
      }
      // regulariser incidences of the row, split over the lanes of the group
      const int a1 = __ldg(P.inc_ptr + i + 1);
      for (int a = __ldg(P.inc_ptr + i) + lane; a < a1; a += lpr) {
        const int o = __ldg(P.inc_other + a), ent = __ldg(P.inc_ent + a);
        const int e = ent >> 1;
        const bool second = ent & 1;
        // the first endpoint counts the edge's chi2 — unless it is a fixed row (then this row is the only unknown)
        const bool counts = !second || o >= Q.pl.V;
        if (!LIN && !counts) continue;
        const D3 xo = dld3(P.x, o);
        double s, u[3], c;
        const double ch = second ? pair_edge<LIN>(e, o, i, xo, xi, s, u, c) : pair_edge<LIN>(e, i, o, xi, xo, s, u, c);
        if (counts) chi += ch;
        if (LIN) {
          D[0] += s + u[0] * u[0];
          D[1] += u[0] * u[1];
          D[2] += u[0] * u[2];
          D[3] += s + u[1] * u[1];
          D[4] += u[1] * u[2];
          D[5] += s + u[2] * u[2];
          const double sg = second ? -c : c;
          b[0] -= s * (xi.x - xo.x) + sg * u[0];
          b[1] -= s * (xi.y - xo.y) + sg * u[1];
          b[2] -= s * (xi.z - xo.z) + sg * u[2];
          if (!second) {
            double2* pcw = reinterpret_cast<double2*>(P.pc + 4 * (size_t)e);
            pcw[0] = make_double2(s, u[0]);
            pcw[1] = make_double2(u[1], u[2]);
          }
        }
      }
    }
    if (LIN) {
      // the lanes of a group live in one warp: reduce D and b over them
#pragma unroll
      for (int k = 0; k < 6; k++) D[k] = group_sum(D[k]);
#pragma unroll
      for (int k = 0; k < 3; k++) b[k] = group_sum(b[k]);
      if (i >= 0 && lane == 0) {
        double2* d = reinterpret_cast<double2*>(P.dg + 8 * (size_t)i);
        d[0] = make_double2(D[0], D[1]);
        d[1] = make_double2(D[2], D[3]);
        d[2] = make_double2(D[4], D[5]);
        dst3(P.bvec, i, D3{b[0], b[1], b[2]});
        maxd = fmax(maxd, fmax(fabs(D[0]), fmax(fabs(D[3]), fabs(D[5]))));
      }
      if (lane == 0 && slot < rpc && !rec_written) {
        double* rec = s_rec + kRec * slot;
#pragma unroll
        for (int k = 0; k < kRec; k++) rec[k] = 0.0;
      }
      __syncthreads();
      // pose-block partial of this CTA: 21 entries of H_pp (upper) + 6 of b_p, rows added in slot order
      if (tid < 27 && Q.pl.np) {
        const int v = tid;
        int a = 0, c = 0;
        if (v < 21) {
          int t = v;
          while (t >= 6 - a) {
            t -= 6 - a;
            a++;
          }
          c = a + t;
        } else {
          a = v - 21;
        }
        double s = 0;
        for (int r = 0; r < rpc; r++) {
          const double* rec = s_rec + kRec * r;
          if (v < 21)
            s += rec[12] * (rec[a] * rec[c] + rec[6 + a] * rec[6 + c]);
          else
            s += rec[a] * rec[14] + rec[6 + a] * rec[15];
        }
        Q.hpp_part[(size_t)cta * 28 + v] = s;
      }
    }
  }

  // Sum of the CTAs' pose-block partials in CTA order -> s_hpp (every CTA); CTA 0 publishes it for the root front.
  __device__ void reduce_hpp() {
    for (int tt = tid; tt < 27 * 8; tt += kDBlock) {
      const int v = tt % 27, g = tt / 27;
      double s = 0;
      for (int c0 = g; c0 < G; c0 += 64) {
        double o[8];
#pragma unroll
        for (int u = 0; u < 8; u++) o[u] = __ldcg(Q.hpp_part + (size_t)min(c0 + 8 * u, G - 1) * 28 + v);
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (c0 + 8 * u < G) s += o[u];
      }
      s_red[tt] = s;
    }
    __syncthreads();
    if (tid < 27) {
      double s = 0;
#pragma unroll
      for (int g = 0; g < 8; g++) s += s_red[27 * g + tid];
      s_hpp[tid] = s;
      if (cta == 0) Q.hpp[tid] = s;
    }
    __syncthreads();
  }

  // ================================================================================================
  // Exact solve of (H + lambda I) delta = b. Returns false when a pivot was not positive (uniform over the grid).
  // ================================================================================================
  __device__ bool solve() {
    direct::Sys sys;
    sys.dg = P.dg;
    sys.cpl = Q.cpl;
    sys.bvec = P.bvec;
    sys.pc = P.pc;
    sys.hpp = Q.hpp;
    sys.inc_ptr = P.inc_ptr;
    sys.inc_ent = P.inc_ent;
    sys.inc_pos = Q.inc_pos;
    sys.inc_row = P.inc_row;
    sys.lambda = lambda;
    const direct::Thr th{tid, kDBlock};
    const int depth = Q.pl.depth;
    nsolve++;
    for (int d = depth; d >= 0; d--) {
      const int t = (1 << d) + (cta >> (depth - d));
      const unsigned long long R = (unsigned long long)(G >> d);
      const long long t0 = clock64();
      direct::stage_ab(Q.pl, sys, cta, d, sp, s_w, s_v, th, prof);
      const long long t1 = clock64();
      prof[0] += t1 - t0;
#ifdef NRS_DIRECT_PLEV
      plev[d] += t1 - t0;
#endif
      team_barrier(Q.tbar + 2 * t, R);  // every member's L21 rows (and the leader's L11) are in global memory
      if (d > 0) {
        const long long t2 = clock64();
        direct::stage_c(Q.pl, cta, d, sp, th);
        const long long t2b = clock64();
        prof[1] += t2b - t2;
#ifdef NRS_DIRECT_PLEV
        plev[8 + d] += t2b - t2;
#endif
        team_barrier(Q.tbar + 2 * (t >> 1) + 1, 2 * R);  // both children's update matrices are complete
      }
    }
    const int f = __ldcg(Q.pl.fail);
    const bool failed = f != fail_seen;
    fail_seen = f;
    if (failed) return false;
    const long long t3 = clock64();
    for (int d = 0; d <= depth; d++) {
      const long long t4 = clock64();
      direct::backward_front(Q.pl, cta, d, s_path, sp, s_z, P.xcg, Q.dpose, th, prof, s_mbar, &mphase);
#ifdef NRS_DIRECT_PLEV
      plev[16 + d] += clock64() - t4;
#else
      (void)t4;
#endif
    }
    prof[2] += clock64() - t3;
    return true;
  }

  // ================================================================================================
  // One LM iteration (optimization_algorithm_levenberg.cpp:57-151). Returns true on g2o's "Terminate".
  // ================================================================================================
  __device__ bool lm_iteration(int iteration) {
    const long long tl0 = clock64();
    barrier();  // estimates / levels written by other CTAs are visible
    double chi = 0, maxd = 0;
    graph_pass<true>(chi, maxd);
    grid_reduce2(chi, maxd, true);
    double currentChi = s_scal[0];
    double maxDiag = s_scal[1];
    if (Q.pl.np) reduce_hpp();
    n_sweeps++;
    if (iteration == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:153-165
      if (Q.pl.np)
        for (int a = 0; a < 6; a++) maxDiag = fmax(maxDiag, fabs(s_hpp[direct::sym6i(a, a)]));
      lambda = P.lm_tau * maxDiag;
      ni = 2;
    }
    prof[5] += clock64() - tl0;
    double rho = 0;
    int qmax = 0;
    const int nv_root = __ldg(Q.pl.nv + 1);
    do {
      const long long ts0 = clock64();
      const bool solved = solve();
      const long long ts1 = clock64();
      prof[6] += ts1 - ts0;
      lm_trials++;
      if (!solved) n_fail++;
      double tchi = 0, scale = 0;
      if (solved) {
        barrier();  // delta of every front is in global memory
        const int i = my_row();
        if (i >= 0 && lane == 0) {  // push + update (sparse_optimizer.cpp:457-470)
          const D3 d = dld3(P.xcg, i), bb = dld3(P.bvec, i);
          D3 x = dld3(P.x, i);
          dst3(P.x_bak, i, x);
          x.x += d.x; x.y += d.y; x.z += d.z;
          dst3(P.x, i, x);
          scale += d.x * (lambda * d.x + bb.x) + d.y * (lambda * d.y + bb.y) + d.z * (lambda * d.z + bb.z);
        }
        const double* dp = s_path + 3 * (nv_root - 2);  // the root owns the pose: last 6 entries of its solution
        if (tid == 0 && Q.pl.np) {
          for (int t = 0; t < 7; t++) s_pose_bak[t] = s_pose[t];
          double dl[6];
          for (int t = 0; t < 6; t++) dl[t] = dp[t];
          pose_oplus(s_pose, dl);
          if (cta == 0)
            for (int t = 0; t < 6; t++) scale += dl[t] * (lambda * dl[t] + s_hpp[21 + t]);
        }
        barrier();  // updated estimates are read across CTAs by the regulariser edges
        double dummy = 0;
        graph_pass<false>(tchi, dummy);
        n_chi2++;
      }
      grid_reduce2(tchi, scale, false);
      const double tempChi = solved ? s_scal[0] : DBL_MAX;
      scale = s_scal[1] + 1e-3;
      rho = (currentChi - tempChi) / scale;
      if (rho > 0 && isfinite(tempChi)) {
        const double t3 = 2 * rho - 1;
        double alpha = 1. - t3 * t3 * t3;
        alpha = fmin(alpha, 2. / 3.);
        const double scaleFactor = fmax(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        if (solved) {  // pop
          const int i = my_row();
          if (i >= 0 && lane == 0) dst3(P.x, i, dld3(P.x_bak, i));
          __syncthreads();
          if (tid == 0 && Q.pl.np)
            for (int t = 0; t < 7; t++) s_pose[t] = s_pose_bak[t];
          __syncthreads();
        }
        if (!isfinite(lambda)) break;
      }
      prof[7] += clock64() - ts1;
      qmax++;
    } while (rho < 0 && qmax < P.lm_max_trials);
    lm_iters++;
    if (cta == 0 && tid == 0 && n_trace < kTrace) P.stats->chi2_trace[n_trace] = currentChi;
    n_trace++;
    return (qmax == P.lm_max_trials || rho == 0 || !isfinite(lambda));
  }

  __device__ void run() {
    const long long trun0 = clock64();
    if (tid < 7) s_pose[tid] = P.pose[tid];
    __syncthreads();
    const int gsz = G * kDBlock, gt = cta * kDBlock + tid;
    for (int o = 0; o < P.n_ops; o++) {
      const int op = P.op[o], arg = P.op_arg[o];
      switch (op) {
        case OP_RESET: {
          barrier();
          for (int i = gt; i < P.V; i += gsz) dst3(P.x, i, dld3c(P.x_seed, i));
          __syncthreads();
          if (P.seed_via_f32) {
            if (tid == 0) {  // g2o_optimization.cc:144-145 then :69-71 — the pose crosses the Frame as Sophus::SE3f
              double sd[7];
              float sf[7];
              for (int t = 0; t < 7; t++) sd[t] = __ldcg(P.pose_seed + t);
              pose_to_f7(sd, sf);
              pose_from_f7(sf, s_pose);
            }
          } else if (tid < 7) {
            s_pose[tid] = P.pose_seed[tid];
          }
          __syncthreads();
        } break;
        case OP_CLEAR_LEVELS: {
          barrier();
          for (int i = gt; i < P.V; i += gsz) P.rp_level[i] = 0;
          for (int e = gt; e < P.P; e += gsz) P.sp_level[e] = 0;
        } break;
        case OP_OPTIMIZE: {
          for (int it = 0; it < arg; it++)
            if (lm_iteration(it)) break;
        } break;
        case OP_RELEVEL_DEFORM: {
          // g2o_optimization.cc:352-394 — reprojection edges by chi2 > 5.99; every spatial edge ends on its own
          // chi2 > 0.584 test.
          barrier();
          for (int i = gt; i < P.V; i += gsz) {
            double pc[3], err[2];
            const double c2 = reproj_error(i, dld3(P.x, i), pc, err);
            P.rp_chi2[i] = c2;
            P.rp_level[i] = ((float)c2 > P.th2f) ? 1 : 0;
          }
          for (int e = gt; e < P.P; e += gsz) {
            const double w = P.pair_w[e];
            if (w < 0) continue;
            const D3 xi = dld3(P.x, P.pair_i[e]), xj = dld3(P.x, P.pair_j[e]);
            const double e0 = w * (xi.x - xj.x), e1 = w * (xi.y - xj.y), e2 = w * (xi.z - xj.z);
            const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
            P.sp_level[e] = (c2 > (double)P.th3f) ? 1 : 0;
          }
        } break;
        case OP_FINAL_CHI2: {
          barrier();
          for (int i = gt; i < P.V; i += gsz) {
            double pc[3], err[2];
            P.rp_chi2[i] = reproj_error(i, dld3(P.x, i), pc, err);
          }
        } break;
        default: break;
      }
    }
    __syncthreads();
#ifdef NRS_DIRECT_PLEV
    if (Q.plev && tid == 0)
      for (int i = 0; i < 32; i++) Q.plev[(size_t)cta * 32 + i] = plev[i];
#endif
    if (cta == 0) {
      if (tid < 7) P.pose[tid] = s_pose[tid];
      if (tid == 0) {
        EngineStats* st = P.stats;
        st->lm_iterations = lm_iters;
        st->lm_trials = lm_trials;
        st->pcg_iterations = 0;
        st->n_sweeps = n_sweeps;
        st->n_chi2_passes = n_chi2;
        st->n_trace = n_trace < kTrace ? n_trace : kTrace;
        st->pcg_fail = n_fail;
        st->barriers = (int)gen;
        st->xepochs = 0;
        st->xfail = 0;
        st->lambda_final = lambda;
        prof[15] = clock64() - trun0;
        for (int i = 0; i < 16; i++) st->prof[i] = prof[i];
      }
    }
  }
};

__global__ void __launch_bounds__(kDBlock, 1) nrs_track_direct_kernel(const __grid_constant__ DirectParams q) {
  extern __shared__ __align__(16) double nrs_dsmem[];
  DirectEngine eng(q, nrs_dsmem);
  eng.run();
}

}  // namespace

size_t direct_smem_bytes(int max_path, int scratch_z, int max_nv, int max_rows, size_t panel_doubles) {
  size_t d = 8 + 8 + 8 + 2 + direct_sw_doubles(max_nv) + direct_sv_doubles(max_rows) + 28 + 27 * 8 +
             ((max_path + 1) & ~1) + ((scratch_z + 1) & ~1) +
             (panel_doubles > (size_t)kRec * kDBlock ? panel_doubles : (size_t)kRec * kDBlock) + 2;
  return d * sizeof(double);
}

int direct_block_threads() { return kDBlock; }

int direct_max_grid(size_t smem) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(nrs_track_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrs_track_direct_kernel, kDBlock, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return sms * per_sm;
}

int launch_direct(const DirectParams& q, int grid, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024) {
    // per device: the attribute belongs to the function in the CURRENT context
    const cudaError_t e =
        cudaFuncSetAttribute(nrs_track_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaError_t e = cudaMemsetAsync(q.P.bar, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(q.pl.fail, 0, sizeof(int), stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(q.tbar, 0, sizeof(unsigned long long) * 2 * (size_t)((2 << q.pl.depth) + 1), stream);
  if (e != cudaSuccess) return (int)e;
  void* args[] = {const_cast<DirectParams*>(&q)};
  return (int)cudaLaunchCooperativeKernel((const void*)nrs_track_direct_kernel, dim3(grid), dim3(kDBlock), args, smem,
                                          stream);
}

}  // namespace nrs
