// nrs_engine.cu — persistent, cooperative Levenberg–Marquardt kernel for sm_100a.
//
// What it replaces (reference paths relative to /root/reference):
//   SparseOptimizer::optimize / computeActiveErrors / update / push / pop   third_party/g2o/g2o/core/sparse_optimizer.cpp:62-114,392-470
//   OptimizationAlgorithmLevenberg::solve (lambda control, gain ratio)      third_party/g2o/g2o/core/optimization_algorithm_levenberg.cpp:57-174
//   BlockSolver::buildSystem / setLambda / solve                           third_party/g2o/g2o/core/block_solver.hpp:329-341,495-603
//   BaseFixedSizedEdge::constructQuadraticForm                             third_party/g2o/g2o/core/base_fixed_sized_edge.hpp:49-133
//   the ten edge types of modules/optimization/*.cc (cited at each formula)
//   the round / re-levelling logic of modules/optimization/g2o_optimization.cc:100-140,338-395
//
// Design: one launch runs a whole driver program. Each CTA owns "chunks" of point rows; a row is a
// point vertex with its reprojection edge and the regulariser edges incident to it, so the normal equations are
// applied matrix-free, row by row, without atomics and in a fixed summation order. The reference factorises
// H + lambda*I exactly (sparse LL^T); here the damped system is solved by block-Jacobi preconditioned CG whose
// vectors stay in HBM/L2 and whose 6-dof pose blocks are replicated in every CTA's shared memory. CTAs meet at a
// global-memory barrier (release/acquire on one counter); reductions go through per-CTA slots summed in a fixed
// order, so every CTA derives bit-identical scalars and takes the same branches.
#include <cooperative_groups.h>
#include <float.h>
#include <stdio.h>

#include "nrs_engine.cuh"

namespace nrs {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 ld3(const double* base, int i) {  // L2-coherent (data written by other CTAs)
  const double2 a = __ldcg(reinterpret_cast<const double2*>(base + 4 * (size_t)i));
  const double b = __ldcg(base + 4 * (size_t)i + 2);
  return V3{a.x, a.y, b};
}
__device__ __forceinline__ void st3(double* base, int i, const V3& v) {
  *reinterpret_cast<double2*>(base + 4 * (size_t)i) = make_double2(v.x, v.y);
  base[4 * (size_t)i + 2] = v.z;
}
__device__ __forceinline__ V3 ld3c(const double* base, int i) {  // read-only input
  const double2 a = __ldg(reinterpret_cast<const double2*>(base + 4 * (size_t)i));
  const double b = __ldg(base + 4 * (size_t)i + 2);
  return V3{a.x, a.y, b};
}

__device__ __forceinline__ int sym6(int a, int c) { return a * 6 - (a * (a - 1)) / 2 + (c - a); }  // a <= c

// Inverse of the SPD 6x6 (upper-packed H + lambda I) by Cholesky; out = full 36. Returns false if not SPD.
__device__ bool invert6(const double* Hu, double lambda, double* out) {
  double L[36];
  for (int i = 0; i < 36; i++) L[i] = 0;
  for (int j = 0; j < 6; j++) {
    double s = Hu[sym6(j, j)] + lambda;
    for (int k = 0; k < j; k++) s -= L[j * 6 + k] * L[j * 6 + k];
    if (!(s > 0)) return false;
    const double d = sqrt(s);
    L[j * 6 + j] = d;
    for (int i = j + 1; i < 6; i++) {
      double t = Hu[sym6(j, i)];
      for (int k = 0; k < j; k++) t -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = t / d;
    }
  }
  for (int c = 0; c < 6; c++) {
    double y[6];
    for (int i = 0; i < 6; i++) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
      y[i] = s / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * y[k];
      y[i] = s / L[i * 6 + i];
    }
    for (int i = 0; i < 6; i++) out[i * 6 + c] = y[i];
  }
  return true;
}

struct Engine {
  const Params& P;
  // shared memory
  double *s_pose, *s_pose_bak, *s_H, *s_M, *s_bp, *s_xp, *s_rp, *s_zp, *s_pp, *s_qp, *s_red, *s_scal;
  int* s_flag;
  unsigned gen;
  int tid, nthr;
  // LM state (uniform)
  double lambda, ni;
  int lm_iters, lm_trials, pcg_iters, n_sweeps, n_chi2, n_trace, pcg_fail;
  long long prof[16];

  __device__ Engine(const Params& p, double* sm) : P(p) {
    tid = threadIdx.x;
    nthr = blockDim.x;
    const int F = p.F;
    s_pose = sm;            sm += 7 * F;
    s_pose_bak = sm;        sm += 7 * F;
    s_H = sm;               sm += 21 * F;
    s_M = sm;               sm += 36 * F;
    s_bp = sm;              sm += 6 * F;
    s_xp = sm;              sm += 6 * F;
    s_rp = sm;              sm += 6 * F;
    s_zp = sm;              sm += 6 * F;
    s_pp = sm;              sm += 6 * F;
    s_qp = sm;              sm += 6 * F;
    s_red = sm;             sm += 32 * kChunkVals;
    s_scal = sm;            sm += 32;
    s_flag = reinterpret_cast<int*>(sm);
    gen = 0;
    lambda = -1;
    ni = 2;
    lm_iters = lm_trials = pcg_iters = n_sweeps = n_chi2 = n_trace = pcg_fail = 0;
    for (int i = 0; i < 16; i++) prof[i] = 0;
  }

  // ---- grid-wide barrier: release-add on one counter, acquire-poll until every CTA of this generation arrived
  __device__ __forceinline__ void barrier() {
    __syncthreads();
    gen++;
    if (tid == 0) {
      const long long t0 = clock64();
      __threadfence();
      red_release_add_u64(P.bar, 1ULL);
      const unsigned long long target = (unsigned long long)gen * gridDim.x;
      while (ld_acquire_u64(P.bar) < target) {
      }
      __threadfence();
      prof[0] += clock64() - t0;
    }
    __syncthreads();
  }

  // ---- block reduction of NV per-thread values (fixed order); result in dst[0..NV)
  template <int NV>
  __device__ __forceinline__ void block_reduce(double (&v)[NV], double* dst) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    }
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) s_red[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (tid < NV) {
      double s = 0;
      for (int w = 0; w < nw; w++) s += s_red[w * NV + tid];
      dst[tid] = s;
    }
    __syncthreads();
  }

  // ---- grid reduction of n (<= 4) values: v[k] are per-thread partials. maxmask bit k: max instead of sum.
  // Result (identical in every CTA) lands in s_scal[0..n). Includes one grid barrier.
  template <int N>
  __device__ __forceinline__ void grid_reduce(double (&v)[N], unsigned maxmask) {
    const int par = gen & 1;
    // block level
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v[k], off);
        v[k] = ((maxmask >> k) & 1) ? fmax(v[k], o) : v[k] + o;
      }
    }
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < N; k++) s_red[warp * N + k] = v[k];
    }
    __syncthreads();
    if (tid < N) {
      double s = s_red[tid];
      for (int w = 1; w < nw; w++) s = ((maxmask >> tid) & 1) ? fmax(s, s_red[w * N + tid]) : s + s_red[w * N + tid];
      P.slots[((size_t)par * gridDim.x + blockIdx.x) * kSlotVals + tid] = s;
    }
    barrier();
    if (tid < 32 * N) {
      const int k = tid >> 5;
      const bool mx = (maxmask >> k) & 1;
      double s = mx ? -DBL_MAX : 0.0;
      for (int c = lane; c < (int)gridDim.x; c += 32) {
        const double o = __ldcg(P.slots + ((size_t)par * gridDim.x + c) * kSlotVals + k);
        s = mx ? fmax(s, o) : s + o;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, s, off);
        s = mx ? fmax(s, o) : s + o;
      }
      if (lane == 0) s_scal[k] = s;
    }
    __syncthreads();
  }

  // ================================================================================================
  // Edge-parallel pass over pair and damper edges: chi2 (and, when LIN, the linearised coefficients).
  //   spatial  : SpatialRegularizerWithDeformation   optimization/spatial_regularizer_with_deformation.cc:36-49
  //   spring   : PositionRegularizerWithDeformation  optimization/position_regularizer_with_deformation.cc:31-57
  //              PositionRegularizer (quirk E1)      optimization/position_regularizer.cc:32-61
  //   damper   : SpatialRegularizer                  optimization/spatial_regularizer.cc:32-59
  // ================================================================================================
  template <bool LIN>
  __device__ void edges_pass(double& chi) {
    const int gsz = gridDim.x * nthr;
    for (int e = blockIdx.x * nthr + tid; e < P.P; e += gsz) {
      const int i = P.pair_i[e], j = P.pair_j[e];
      const V3 xi = ld3(P.x, i), xj = ld3(P.x, j);
      double s = 0, u0 = 0, u1 = 0, u2 = 0, c = 0;
      const double w = P.pair_w[e];
      // an edge whose vertices are all fixed is not part of the active set (sparse_optimizer.cpp:232-246)
      const bool live = !(P.pt_fixed && P.pt_fixed[i] && P.pt_fixed[j]);
      if (live && w >= 0 && P.sp_level[e] == 0) {
        const double e0 = w * (xi.x - xj.x), e1 = w * (xi.y - xj.y), e2 = w * (xi.z - xj.z);
        const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
        double rho, drho;
        huber(c2, P.delta_spatial, rho, drho);
        chi += rho;
        s = drho * P.info_spatial * w * w;
      }
      if (live && P.spring_kind != SPRING_NONE) {
        const V3 ri = ld3c(P.rest, i), rj = ld3c(P.rest, j);
        const double c1x = ri.x + xi.x, c1y = ri.y + xi.y, c1z = ri.z + xi.z;
        const double c2x = rj.x + xj.x, c2y = rj.y + xj.y, c2z = rj.z + xj.z;
        const double dx = c1x - c2x, dy = c1y - c2y, dz = c1z - c2z;
        const double dist = sqrt(dx * dx + dy * dy + dz * dz);
        const double d0 = P.pair_d0[e];
        const double err = P.spring_k * (dist - d0) / d0;
        const double ch = err * err * P.info_spring;
        double rho, drho;
        huber(ch, P.delta_spring, rho, drho);
        chi += rho;
        if (LIN) {
          double j0, j1, j2;
          if (P.spring_kind == SPRING_DEFORM) {
            const double aa = P.spring_k / (2 * d0 * dist);
            j0 = aa * (2 * c1x - 2 * c2x);
            j1 = aa * (2 * c1y - 2 * c2y);
            j2 = aa * (2 * c1z - 2 * c2z);
          } else {
            const double kd = P.spring_k / d0, dcs = 1.0 / sqrt(dist);
            j0 = kd * dcs * (2.0 * dx);
            j1 = kd * dcs * (2.0 * dy);
            j2 = kd * dcs * (2.0 * dz);
          }
          const double sw = sqrt(drho * P.info_spring);
          u0 = sw * j0;
          u1 = sw * j1;
          u2 = sw * j2;
          c = sw * err;
        }
      }
      if (LIN) {
        double2* o = reinterpret_cast<double2*>(P.pc + 8 * (size_t)e);
        o[0] = make_double2(s, u0);
        o[1] = make_double2(u1, u2);
        o[2] = make_double2(c, 0.0);
      }
    }
    for (int e = blockIdx.x * nthr + tid; e < P.D; e += gsz) {
      const int4 v = *reinterpret_cast<const int4*>(P.dmp_v + 4 * (size_t)e);
      const V3 a = ld3(P.x, v.x), b = ld3(P.x, v.y), an = ld3(P.x, v.z), bn = ld3(P.x, v.w);
      const double w = P.dmp_w[e];
      const double e0 = w * ((an.x - a.x) - (bn.x - b.x));
      const double e1 = w * ((an.y - a.y) - (bn.y - b.y));
      const double e2 = w * ((an.z - a.z) - (bn.z - b.z));
      const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
      double rho, drho;
      huber(c2, P.delta_spatial, rho, drho);
      chi += rho;
      if (LIN) {
        const double g = drho * P.info_spatial * w;
        double2* o = reinterpret_cast<double2*>(P.dc + 4 * (size_t)e);
        o[0] = make_double2(g * w, g * e0);
        o[1] = make_double2(g * e1, g * e2);
      }
    }
  }

  // Reprojection error of point row i at the current estimate; returns chi2 (info * |e|^2).
  //   ReprojectionErrorWithDeformation::computeError  optimization/reprojection_error_with_deformation.cc:37-50
  //   ReprojectionError::computeError                 optimization/reprojection_error.cc:32-44
  //   ReprojectionErrorOnlyPose::computeError         optimization/reprojection_error_only_pose.cc:50-58
  __device__ __forceinline__ double reproj_error(int i, int kf, const V3& xi, double pc[3], double err[2]) {
    const V3 r = ld3c(P.rest, i);
    const double Xw[3] = {xi.x + r.x, xi.y + r.y, xi.z + r.z};
    pose_map(s_pose + 7 * kf, Xw, pc);
    float u, v;
    project_f(P.cam, (float)pc[0], (float)pc[1], (float)pc[2], u, v);
    const double2 z = __ldg(reinterpret_cast<const double2*>(P.uv) + i);
    err[0] = z.x - (double)u;
    err[1] = z.y - (double)v;
    return (err[0] * err[0] + err[1] * err[1]) * P.info_reproj;
  }

  // ================================================================================================
  // Row pass: reprojection edge of each point + gather of the incident regulariser coefficients.
  // LIN: stores Jacobians, diagonal blocks, gradient; reduces the pose blocks per chunk. Always: chi2.
  // ================================================================================================
  template <bool LIN>
  __device__ void rows_pass(double& chi, double& maxd, int par) {
    for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
      const int i = P.chunk_begin[c] + tid;
      const bool valid = i < P.chunk_end[c];
      double red[27];
      if (LIN) {
#pragma unroll
        for (int k = 0; k < 27; k++) red[k] = 0;
      }
      if (valid) {
        const V3 xi = ld3(P.x, i);
        const int kf = P.pt_kf[i];
        double D[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        double A[12], B[6], omega = 0;
        const bool var = !P.points_fixed && !(P.pt_fixed && P.pt_fixed[i]);  // this point is an unknown
        if (kf >= 0 && P.rp_level[i] == 0 && (var || !P.poses_fixed)) {
          double pc[3], err[2];
          const double c2 = reproj_error(i, kf, xi, pc, err);
          double rho, drho;
          huber(c2, P.delta_reproj, rho, drho);
          chi += rho;
          P.rp_chi2[i] = c2;
          if (LIN) {
            // linearizeOplus: J_pose = -J_pi * [ -[p]x | I ],  J_point = -J_pi * R
            //   optimization/reprojection_error_with_deformation.cc:52-68, reprojection_error.cc:46-64,
            //   reprojection_error_only_pose.cc:60-76
            float Jf[6];
            projection_jacobian_f(P.cam, (float)pc[0], (float)pc[1], (float)pc[2], Jf);
            double Jp[6];
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
            omega = drho * P.info_reproj;
            const double we0 = -omega * err[0], we1 = -omega * err[1];
            if (!P.poses_fixed) {
              int t = 0;
#pragma unroll
              for (int a = 0; a < 6; a++)
#pragma unroll
                for (int cc = a; cc < 6; cc++) red[t++] = omega * (A[a] * A[cc] + A[6 + a] * A[6 + cc]);
#pragma unroll
              for (int a = 0; a < 6; a++) red[21 + a] = A[a] * we0 + A[6 + a] * we1;
            }
            if (var) {
              double R[9];
              quat_to_R(s_pose + 7 * kf, R);
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
            } else {
#pragma unroll
              for (int k = 0; k < 6; k++) B[k] = 0;
            }
          }
        }
        if (LIN) {
          double* jo = P.jac + 20 * (size_t)i;
          if (omega == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) A[k] = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) B[k] = 0;
          }
#pragma unroll
          for (int k = 0; k < 6; k++) reinterpret_cast<double2*>(jo)[k] = make_double2(A[2 * k], A[2 * k + 1]);
#pragma unroll
          for (int k = 0; k < 3; k++) reinterpret_cast<double2*>(jo)[6 + k] = make_double2(B[2 * k], B[2 * k + 1]);
          reinterpret_cast<double2*>(jo)[9] = make_double2(omega, 0.0);
        }
        if (!var && !P.points_fixed && LIN) {
          double2* d = reinterpret_cast<double2*>(P.dg + 8 * (size_t)i);
          d[0] = d[1] = d[2] = d[3] = make_double2(0.0, 0.0);
          st3(P.bvec, i, V3{0, 0, 0});
        }
        if (var) {
          double su = 0;  // unary diagonal weight
          if (P.unary_on) {
            // SpatialRegularizerFixed  optimization/spatial_regularizer_fixed.cc:32-43 — the reference value is
            // read live from another vertex and carries no Jacobian
            for (int a = P.un_ptr[i]; a < P.un_ptr[i + 1]; a++) {
              const double w = P.un_w[a];
              const V3 rf = ld3(P.x, P.un_ref[a]);
              const double d0 = xi.x - rf.x, d1 = xi.y - rf.y, d2 = xi.z - rf.z;
              const double c2 = w * w * (d0 * d0 + d1 * d1 + d2 * d2) * P.info_spatial;
              double rho, drho;
              huber(c2, P.delta_spatial, rho, drho);
              chi += rho;
              if (LIN) {
                const double s = drho * P.info_spatial * w * w;
                su += s;
                b[0] -= s * d0;
                b[1] -= s * d1;
                b[2] -= s * d2;
              }
            }
          }
          if (LIN) {
            for (int a = P.inc_ptr[i]; a < P.inc_ptr[i + 1]; a++) {
              const int other = P.inc_other[a], ent = P.inc_ent[a];
              const double2* cf = reinterpret_cast<const double2*>(P.pc + 8 * (size_t)(ent >> 1));
              const double2 c0 = __ldcg(cf), c1 = __ldcg(cf + 1);
              const double cc = __ldcg(P.pc + 8 * (size_t)(ent >> 1) + 4);
              const double s = c0.x, u0 = c0.y, u1 = c1.x, u2 = c1.y;
              const V3 xo = ld3(P.x, other);
              D[0] += s + u0 * u0;
              D[1] += u0 * u1;
              D[2] += u0 * u2;
              D[3] += s + u1 * u1;
              D[4] += u1 * u2;
              D[5] += s + u2 * u2;
              const double sg = (ent & 1) ? -cc : cc;
              b[0] -= s * (xi.x - xo.x) + sg * u0;
              b[1] -= s * (xi.y - xo.y) + sg * u1;
              b[2] -= s * (xi.z - xo.z) + sg * u2;
            }
            if (P.D > 0) {
              for (int a = P.dinc_ptr[i]; a < P.dinc_ptr[i + 1]; a++) {
                const int ent = P.dinc_ent[a];
                const double2* cf = reinterpret_cast<const double2*>(P.dc + 4 * (size_t)(ent >> 2));
                const double2 c0 = __ldcg(cf), c1 = __ldcg(cf + 1);
                const int role = ent & 3;
                const double sg = (role == 1 || role == 2) ? 1.0 : -1.0;  // J = (-w, +w, +w, -w) I
                D[0] += c0.x;
                D[3] += c0.x;
                D[5] += c0.x;
                b[0] -= sg * c0.y;
                b[1] -= sg * c1.x;
                b[2] -= sg * c1.y;
              }
            }
            D[0] += su;
            D[3] += su;
            D[5] += su;
            double2* d = reinterpret_cast<double2*>(P.dg + 8 * (size_t)i);
            d[0] = make_double2(D[0], D[1]);
            d[1] = make_double2(D[2], D[3]);
            d[2] = make_double2(D[4], D[5]);
            d[3] = make_double2(su, 0.0);
            st3(P.bvec, i, V3{b[0], b[1], b[2]});
            maxd = fmax(maxd, fmax(fabs(D[0]), fmax(fabs(D[3]), fabs(D[5]))));
          }
        }
      }
      if (LIN && !P.poses_fixed) {
        block_reduce<27>(red, P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals);
      }
    }
  }

  // own-row iteration helper
  template <typename Fn>
  __device__ __forceinline__ void for_rows(Fn fn) {
    for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
      const int i = P.chunk_begin[c] + tid;
      if (i < P.chunk_end[c]) fn(i);
    }
  }

  // ================================================================================================
  // Block-Jacobi preconditioned CG on (H + lambda I) delta = b.  Result: xcg rows, s_xp poses.
  // Returns false on breakdown (treated like g2o's failed linear solve,
  // optimization_algorithm_levenberg.cpp:102-121).
  // ================================================================================================
  __device__ bool pcg() {
    const int F = P.F;
    const bool pts = !P.points_fixed, pos = !P.poses_fixed;
    // ---- preconditioner
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    if (pos) {
      for (int k = tid; k < F; k += nthr)
        if (!invert6(s_H + 21 * k, lambda, s_M + 36 * k)) *s_flag = 1;
    }
    double rz_part[1] = {0};
    if (pts) {
      for_rows([&](int i) {
        if (P.pt_fixed && P.pt_fixed[i]) {  // no unknowns: keep p = z = 0 for this row
          double2* mo = reinterpret_cast<double2*>(P.minv + 8 * (size_t)i);
          mo[0] = mo[1] = mo[2] = make_double2(0.0, 0.0);
          st3(P.rvec, i, V3{0, 0, 0});
          st3(P.xcg, i, V3{0, 0, 0});
          st3(P.rec + 8 * (size_t)P.V, i * 2, V3{0, 0, 0});
          return;
        }
        const double2* d = reinterpret_cast<const double2*>(P.dg + 8 * (size_t)i);
        const double2 d0 = d[0], d1 = d[1], d2 = d[2];
        const double a = d0.x + lambda, b = d0.y, c = d1.x, e = d1.y + lambda, f = d2.x, g = d2.y + lambda;
        // symmetric 3x3 inverse by cofactors
        const double C00 = e * g - f * f, C01 = c * f - b * g, C02 = b * f - c * e;
        const double det = a * C00 + b * C01 + c * C02;
        const double id = 1.0 / det;
        const double m00 = C00 * id, m01 = C01 * id, m02 = C02 * id;
        const double m11 = (a * g - c * c) * id, m12 = (b * c - a * f) * id, m22 = (a * e - b * b) * id;
        double2* mo = reinterpret_cast<double2*>(P.minv + 8 * (size_t)i);
        mo[0] = make_double2(m00, m01);
        mo[1] = make_double2(m02, m11);
        mo[2] = make_double2(m12, m22);
        const V3 r = ld3(P.bvec, i);
        const V3 z{m00 * r.x + m01 * r.y + m02 * r.z, m01 * r.x + m11 * r.y + m12 * r.z,
                   m02 * r.x + m12 * r.y + m22 * r.z};
        st3(P.rvec, i, r);
        st3(P.xcg, i, V3{0, 0, 0});
        st3(P.rec + 8 * (size_t)P.V, i * 2, z);  // rec[1][i].z  (record stride 8 = 2 x 4)
        rz_part[0] += r.x * z.x + r.y * z.y + r.z * z.z;
      });
    }
    __syncthreads();
    if (*s_flag) return false;  // uniform: every CTA inverts the same blocks
    double rz_pose = 0;
    if (pos) {
      for (int t = tid; t < 6 * F; t += nthr) {
        s_rp[t] = s_bp[t];
        s_xp[t] = 0;
        s_pp[t] = 0;
      }
      __syncthreads();
      for (int t = tid; t < 6 * F; t += nthr) {
        const int k = t / 6, a = t % 6;
        double s = 0;
        for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_rp[6 * k + c];
        s_zp[t] = s;
      }
      __syncthreads();
      for (int t = 0; t < 6 * F; t++) rz_pose += s_rp[t] * s_zp[t];
    }
    grid_reduce<1>(rz_part, 0);
    double rz = s_scal[0] + rz_pose;
    const double rz0 = rz;
    if (!(rz0 > 0)) return isfinite(rz0);  // b == 0: delta = 0
    const double stop = P.pcg_tol * P.pcg_tol * rz0;
    double beta = 0;
    bool ok = true;
    int it = 0;
    for (; it < P.pcg_max_iter; it++) {
      const int bufR = (it & 1) ^ 1, bufW = it & 1;
      const double* recR = P.rec + (size_t)bufR * 8 * P.V;
      double* recW = P.rec + (size_t)bufW * 8 * P.V;
      const bool first = (it == 0);
      const int par = gen & 1;
      if (pos) {
        for (int t = tid; t < 6 * F; t += nthr) s_pp[t] = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
        __syncthreads();
      }
      // ---- q = (H + lambda I) p, row by row
      const long long tm0 = clock64();
      double pq_part[1] = {0};
      for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
        const int i = P.chunk_begin[c] + tid;
        const bool valid = i < P.chunk_end[c];
        double red[6] = {0, 0, 0, 0, 0, 0};
        if (valid) {
          V3 pi{0, 0, 0};
          if (pts) {
            const V3 zi = ld3(recR, 2 * i);
            if (first) {
              pi = zi;
            } else {
              const V3 po = ld3(recR, 2 * i + 1);
              pi = V3{zi.x + beta * po.x, zi.y + beta * po.y, zi.z + beta * po.z};
            }
            st3(recW, 2 * i + 1, pi);
          }
          double q0 = lambda * pi.x, q1 = lambda * pi.y, q2 = lambda * pi.z;
          const int kf = P.pt_kf[i];
          if (kf >= 0) {
            const double2* jo = reinterpret_cast<const double2*>(P.jac + 20 * (size_t)i);
            const double omega = jo[9].x;
            if (omega != 0) {
              double A[12], B[6];
#pragma unroll
              for (int k = 0; k < 6; k++) {
                const double2 t = jo[k];
                A[2 * k] = t.x;
                A[2 * k + 1] = t.y;
              }
#pragma unroll
              for (int k = 0; k < 3; k++) {
                const double2 t = jo[6 + k];
                B[2 * k] = t.x;
                B[2 * k + 1] = t.y;
              }
              double jp0 = 0, jp1 = 0;
              if (pos) {
                const double* pk = s_pp + 6 * kf;
#pragma unroll
                for (int a = 0; a < 6; a++) {
                  jp0 += A[a] * pk[a];
                  jp1 += A[6 + a] * pk[a];
                }
              }
              const double t0 = omega * (jp0 + B[0] * pi.x + B[1] * pi.y + B[2] * pi.z);
              const double t1 = omega * (jp1 + B[3] * pi.x + B[4] * pi.y + B[5] * pi.z);
              q0 += B[0] * t0 + B[3] * t1;
              q1 += B[1] * t0 + B[4] * t1;
              q2 += B[2] * t0 + B[5] * t1;
              if (pos) {
#pragma unroll
                for (int a = 0; a < 6; a++) red[a] = A[a] * t0 + A[6 + a] * t1;
                pq_part[0] += jp0 * t0 + jp1 * t1;
              }
            }
          }
          if (pts) {
            for (int a = P.inc_ptr[i]; a < P.inc_ptr[i + 1]; a++) {
              const int other = P.inc_other[a], ent = P.inc_ent[a];
              const double2* cf = reinterpret_cast<const double2*>(P.pc + 8 * (size_t)(ent >> 1));
              const double2 c0 = cf[0], c1 = cf[1];
              const V3 zo = ld3(recR, 2 * other);
              V3 po = zo;
              if (!first) {
                const V3 pp = ld3(recR, 2 * other + 1);
                po = V3{zo.x + beta * pp.x, zo.y + beta * pp.y, zo.z + beta * pp.z};
              }
              const double dx = pi.x - po.x, dy = pi.y - po.y, dz = pi.z - po.z;
              const double ud = c0.y * dx + c1.x * dy + c1.y * dz;
              q0 += c0.x * dx + c0.y * ud;
              q1 += c0.x * dy + c1.x * ud;
              q2 += c0.x * dz + c1.y * ud;
            }
            if (P.D > 0) {
              for (int a = P.dinc_ptr[i]; a < P.dinc_ptr[i + 1]; a++) {
                const int ent = P.dinc_ent[a];
                const int4 v = *reinterpret_cast<const int4*>(P.dmp_v + 4 * (size_t)(ent >> 2));
                const double s = P.dc[4 * (size_t)(ent >> 2)];
                const int vv[4] = {v.x, v.y, v.z, v.w};
                double r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
                for (int m = 0; m < 4; m++) {
                  const V3 zo = ld3(recR, 2 * vv[m]);
                  V3 po = zo;
                  if (!first) {
                    const V3 pp = ld3(recR, 2 * vv[m] + 1);
                    po = V3{zo.x + beta * pp.x, zo.y + beta * pp.y, zo.z + beta * pp.z};
                  }
                  const double sg = (m == 1 || m == 2) ? 1.0 : -1.0;
                  r0 += sg * po.x;
                  r1 += sg * po.y;
                  r2 += sg * po.z;
                }
                const int role = ent & 3;
                const double sg = ((role == 1 || role == 2) ? 1.0 : -1.0) * s;
                q0 += sg * r0;
                q1 += sg * r1;
                q2 += sg * r2;
              }
            }
            const double su = P.dg[8 * (size_t)i + 6];
            q0 += su * pi.x;
            q1 += su * pi.y;
            q2 += su * pi.z;
            st3(P.qvec, i, V3{q0, q1, q2});
            pq_part[0] += pi.x * q0 + pi.y * q1 + pi.z * q2;
          }
        }
        if (pos) block_reduce<6>(red, P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals);
      }
      const long long tm1 = clock64();
      grid_reduce<1>(pq_part, 0);
      const long long tm2 = clock64();
      double pq = s_scal[0];
      if (pos) {
        for (int t = tid; t < 6 * F; t += nthr) {
          const int k = t / 6, a = t % 6;
          double s = lambda * s_pp[t];
          for (int c = P.kf_chunk_ptr[k]; c < P.kf_chunk_ptr[k + 1]; c++)
            s += __ldcg(P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals + a);
          s_qp[t] = s;
        }
        __syncthreads();
        for (int t = 0; t < 6 * F; t++) pq += lambda * s_pp[t] * s_pp[t];
      }
      if (!(pq > 0) || !isfinite(pq)) {
        ok = false;
        break;
      }
      const double alpha = rz / pq;
      double rzn_part[1] = {0};
      if (pts) {
        for_rows([&](int i) {
          const V3 pi = ld3(recW, 2 * i + 1);
          const V3 qi = ld3(P.qvec, i);
          V3 xi = ld3(P.xcg, i), ri = ld3(P.rvec, i);
          xi.x += alpha * pi.x; xi.y += alpha * pi.y; xi.z += alpha * pi.z;
          ri.x -= alpha * qi.x; ri.y -= alpha * qi.y; ri.z -= alpha * qi.z;
          const double2* mo = reinterpret_cast<const double2*>(P.minv + 8 * (size_t)i);
          const double2 m0 = mo[0], m1 = mo[1], m2 = mo[2];
          const V3 z{m0.x * ri.x + m0.y * ri.y + m1.x * ri.z, m0.y * ri.x + m1.y * ri.y + m2.x * ri.z,
                     m1.x * ri.x + m2.x * ri.y + m2.y * ri.z};
          st3(P.xcg, i, xi);
          st3(P.rvec, i, ri);
          st3(recW, 2 * i, z);
          rzn_part[0] += ri.x * z.x + ri.y * z.y + ri.z * z.z;
        });
      }
      double rzn_pose = 0;
      if (pos) {
        for (int t = tid; t < 6 * F; t += nthr) {
          s_xp[t] += alpha * s_pp[t];
          s_rp[t] -= alpha * s_qp[t];
        }
        __syncthreads();
        for (int t = tid; t < 6 * F; t += nthr) {
          const int k = t / 6, a = t % 6;
          double s = 0;
          for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_rp[6 * k + c];
          s_zp[t] = s;
        }
        __syncthreads();
        for (int t = 0; t < 6 * F; t++) rzn_pose += s_rp[t] * s_zp[t];
      }
      const long long tm3 = clock64();
      grid_reduce<1>(rzn_part, 0);
      const long long tm4 = clock64();
      prof[1] += tm1 - tm0;  // matvec pass
      prof[2] += tm2 - tm1;  // pq reduce (incl. barrier)
      prof[3] += tm3 - tm2;  // update pass
      prof[4] += tm4 - tm3;  // rz reduce (incl. barrier)
      const double rzn = s_scal[0] + rzn_pose;
      if (!isfinite(rzn)) {
        ok = false;
        it++;
        break;
      }
      beta = rzn / rz;
      rz = rzn;
      if (rz <= stop) {
        it++;
        break;
      }
    }
    pcg_iters += it;
    return ok;
  }

  // ================================================================================================
  // One LM iteration. Returns true when the optimisation must terminate (g2o "Terminate").
  // ================================================================================================
  __device__ bool lm_iteration(int iteration) {
    const int F = P.F;
    const bool pts = !P.points_fixed, pos = !P.poses_fixed;
    barrier();  // estimates written by other CTAs (restore / reset) are visible
    double acc[2] = {0, 0};  // chi2, max diagonal
    edges_pass<true>(acc[0]);
    barrier();
    const int par = gen & 1;
    rows_pass<true>(acc[0], acc[1], par);
    grid_reduce<2>(acc, 2u);
    double currentChi = s_scal[0];
    double maxDiag = s_scal[1];
    n_sweeps++;
    if (pos) {
      for (int t = tid; t < 27 * F; t += nthr) {
        const int k = t / 27, v = t % 27;
        double s = 0;
        for (int c = P.kf_chunk_ptr[k]; c < P.kf_chunk_ptr[k + 1]; c++)
          s += __ldcg(P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals + v);
        if (v < 21)
          s_H[21 * k + v] = s;
        else
          s_bp[6 * k + v - 21] = s;
      }
      __syncthreads();
      if (iteration == 0)
        for (int k = 0; k < F; k++)
          for (int a = 0; a < 6; a++) maxDiag = fmax(maxDiag, fabs(s_H[21 * k + sym6(a, a)]));
    }
    if (iteration == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:153-165
      lambda = P.lm_tau * maxDiag;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      const bool solved = pcg();
      lm_trials++;
      if (!solved) pcg_fail++;
      // push + update (sparse_optimizer.cpp:457-470), scale = delta^T (lambda delta + b)
      double acc2[2] = {0, 0};
      if (solved) {
        if (pts) {
          for_rows([&](int i) {
            const V3 d = ld3(P.xcg, i), bb = ld3(P.bvec, i);
            V3 x = ld3(P.x, i);
            st3(P.x_bak, i, x);
            x.x += d.x; x.y += d.y; x.z += d.z;
            st3(P.x, i, x);
            acc2[1] += d.x * (lambda * d.x + bb.x) + d.y * (lambda * d.y + bb.y) + d.z * (lambda * d.z + bb.z);
          });
        }
        if (pos) {
          for (int t = tid; t < 7 * F; t += nthr) s_pose_bak[t] = s_pose[t];
          __syncthreads();
          for (int k = tid; k < F; k += nthr) pose_oplus(s_pose + 7 * k, s_xp + 6 * k);
          __syncthreads();
        }
        barrier();
        edges_pass<false>(acc2[0]);
        double dummy = 0;
        rows_pass<false>(acc2[0], dummy, 0);
        n_chi2++;
      }
      grid_reduce<2>(acc2, 0);
      double tempChi = solved ? s_scal[0] : DBL_MAX;
      double scale = s_scal[1];
      if (pos && solved)
        for (int t = 0; t < 6 * F; t++) scale += s_xp[t] * (lambda * s_xp[t] + s_bp[t]);
      scale += 1e-3;
      rho = (currentChi - tempChi) / scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, 2. / 3.);
        const double scaleFactor = fmax(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        if (solved) {  // pop
          if (pts) for_rows([&](int i) { st3(P.x, i, ld3(P.x_bak, i)); });
          if (pos) {
            __syncthreads();
            for (int t = tid; t < 7 * F; t += nthr) s_pose[t] = s_pose_bak[t];
            __syncthreads();
          }
        }
        if (!isfinite(lambda)) break;
      }
      qmax++;
    } while (rho < 0 && qmax < P.lm_max_trials);
    lm_iters++;
    if (blockIdx.x == 0 && tid == 0 && n_trace < kTrace) P.stats->chi2_trace[n_trace] = currentChi;
    n_trace++;
    return (qmax == P.lm_max_trials || rho == 0 || !isfinite(lambda));
  }

  __device__ void run() {
    const int F = P.F;
    const long long trun0 = clock64();
    for (int t = tid; t < 7 * F; t += nthr) s_pose[t] = P.pose[t];
    __syncthreads();
    for (int o = 0; o < P.n_ops; o++) {
      const int op = P.op[o], arg = P.op_arg[o];
      switch (op) {
        case OP_RESET: {
          barrier();
          for_rows([&](int i) { st3(P.x, i, ld3c(P.x_seed, i)); });
          __syncthreads();
          for (int t = tid; t < 7 * F; t += nthr) s_pose[t] = P.pose_seed[t];
          __syncthreads();
        } break;
        case OP_CLEAR_LEVELS: {
          barrier();
          const int gsz = gridDim.x * nthr;
          for (int i = blockIdx.x * nthr + tid; i < P.V; i += gsz) P.rp_level[i] = 0;
          for (int e = blockIdx.x * nthr + tid; e < P.P; e += gsz) P.sp_level[e] = 0;
        } break;
        case OP_OPTIMIZE: {
          for (int it = 0; it < arg; it++)
            if (lm_iteration(it)) break;
        } break;
        case OP_RELEVEL_POSE: {
          // g2o_optimization.cc:113-134 — inlier edges keep the error of the last evaluated trial (stale after a
          // rejected step), outlier edges are re-evaluated at the final pose.
          barrier();
          for_rows([&](int i) {
            const int kf = P.pt_kf[i];
            if (kf < 0) return;
            double c2;
            if (P.rp_level[i] == 0) {
              c2 = P.rp_chi2[i];
            } else {
              double pc[3], err[2];
              c2 = reproj_error(i, kf, ld3(P.x, i), pc, err);
              P.rp_chi2[i] = c2;
            }
            P.rp_level[i] = ((float)c2 > P.th2f) ? 1 : 0;
          });
        } break;
        case OP_RELEVEL_DEFORM: {
          // g2o_optimization.cc:352-394 — reprojection edges by chi2 > 5.99; every spatial edge ends on its own
          // chi2 > 0.584 test.
          barrier();
          for_rows([&](int i) {
            const int kf = P.pt_kf[i];
            if (kf < 0) return;
            double pc[3], err[2];
            const double c2 = reproj_error(i, kf, ld3(P.x, i), pc, err);
            P.rp_chi2[i] = c2;
            P.rp_level[i] = ((float)c2 > P.th2f) ? 1 : 0;
          });
          const int gsz = gridDim.x * nthr;
          for (int e = blockIdx.x * nthr + tid; e < P.P; e += gsz) {
            const double w = P.pair_w[e];
            if (w < 0) continue;
            const V3 xi = ld3(P.x, P.pair_i[e]), xj = ld3(P.x, P.pair_j[e]);
            const double e0 = w * (xi.x - xj.x), e1 = w * (xi.y - xj.y), e2 = w * (xi.z - xj.z);
            const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
            P.sp_level[e] = (c2 > (double)P.th3f) ? 1 : 0;
          }
        } break;
        case OP_FINAL_CHI2: {
          barrier();
          for_rows([&](int i) {
            const int kf = P.pt_kf[i];
            if (kf < 0) return;
            double pc[3], err[2];
            P.rp_chi2[i] = reproj_error(i, kf, ld3(P.x, i), pc, err);
          });
        } break;
        default: break;
      }
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      for (int t = tid; t < 7 * F; t += nthr) P.pose[t] = s_pose[t];
      if (tid == 0) {
        EngineStats* st = P.stats;
        st->lm_iterations = lm_iters;
        st->lm_trials = lm_trials;
        st->pcg_iterations = pcg_iters;
        st->n_sweeps = n_sweeps;
        st->n_chi2_passes = n_chi2;
        st->n_trace = n_trace < kTrace ? n_trace : kTrace;
        st->pcg_fail = pcg_fail;
        st->barriers = (int)gen;
        st->lambda_final = lambda;
        prof[15] = clock64() - trun0;
        for (int i = 0; i < 16; i++) st->prof[i] = prof[i];
      }
    }
  }
};

__global__ void __launch_bounds__(256, 1) nrs_lm_kernel(const __grid_constant__ Params p) {
  extern __shared__ double nrs_smem[];
  Engine eng(p, nrs_smem);
  eng.run();
}

}  // namespace

size_t engine_smem_bytes(int F, int block) {
  (void)block;
  return sizeof(double) * ((size_t)F * (7 + 7 + 21 + 36 + 6 * 6) + 32 * kChunkVals + 32) + 16;
}

int engine_max_grid(int block, size_t smem) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(nrs_lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrs_lm_kernel, block, smem);
  return sms * per_sm;
}

int launch_engine(const Params& p, int grid, int block, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nrs_lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaError_t e = cudaMemsetAsync(p.bar, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return (int)e;
  void* args[] = {const_cast<Params*>(&p)};
  e = cudaLaunchCooperativeKernel((const void*)nrs_lm_kernel, dim3(grid), dim3(block), args, smem, stream);
  return (int)e;
}

}  // namespace nrs
