// nrs_direct_core.cuh — numeric half of the exact sparse LL^T used by the tracking solve (multifrontal, dense fronts).
//
// What it replaces: Eigen::SimplicialLLT::factorize / solve behind g2o's LinearSolverEigen
//   third_party/g2o/g2o/solvers/eigen/linear_solver_eigen.h:92-136, called once per Levenberg trial from
//   third_party/g2o/g2o/core/block_solver.hpp:329-341 via optimization_algorithm_levenberg.cpp:97-116.
//
// The elimination tree comes from nrs_direct_plan.h (geometric nested dissection, heap-numbered complete binary tree,
// all index lists in 3x3-block units). 2^depth CTAs walk it leaf-to-root in lock-step; a node of level d is worked on
// by a TEAM of 2^(depth-d) CTAs:
//   stage AB  every team member assembles the node's pivot block F11 (original entries + the children's update
//             matrices, pulled through inverse index maps — no atomics, fixed order) and ITS share of the boundary
//             rows F21 (row k belongs to member k mod R), factorises the tall panel [F11 ; F21_mine] right-looking in
//             shared memory (F11 redundantly: no intra-team sync), stores L11 (leader) / its L21 rows to global;
//   -- grid barrier --
//   stage C   member r computes the rows k = r (mod R) of the update matrix U = sum_children - L21 L21^T.
//   -- grid barrier --
// The right-hand side is the LAST boundary row of every front (augmented matrix), so the forward substitution is part
// of the factorisation: the rhs row of a node's panel ends up holding y_own^T and the rhs row of U the reduced rhs.
// The backward substitution needs no synchronisation at all: every CTA walks ITS root-to-leaf path and recomputes the
// (bit-identical) solution of every front on it into a shared-memory path vector.
// The factorisation is a block L D L^T with 3x3 pivot blocks (no square roots: one reciprocal per pivot block): the
// stored factor is the unit lower block-triangular L', the diagonal block positions hold the pivot blocks D_k.
//
// The file compiles for the device and — with NRS_DIRECT_HOST_EMULATION — for the host with one emulated thread per
// CTA (the code between two syncs is race free), which is how the CPU test suite checks the arithmetic without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef NRS_DIRECT_HOST_EMULATION
#include <cuda/ptx>
#endif

#ifdef NRS_DIRECT_HOST_EMULATION
#define NRS_DD inline
#define NRS_DSYNC() ((void)0)
#define NRS_DSYNCWARP() ((void)0)
#define NRS_DLDCG(p) (*(p))
#define NRS_DLDG(p) (*(p))
#define NRS_DFAIL(p) (++*(p))
#define NRS_DCLOCK() 0LL
#define NRS_DRSQRT(x) (1.0 / sqrt(x))
#define NRS_DRCP(x) (1.0 / (x))
#else
#define NRS_DD __device__ __forceinline__
#define NRS_DSYNC() __syncthreads()
#define NRS_DSYNCWARP() __syncwarp()
#define NRS_DLDCG(p) __ldcg(p)
#define NRS_DLDG(p) __ldg(p)
#define NRS_DFAIL(p) atomicAdd((p), 1)
#define NRS_DCLOCK() clock64()
#define NRS_DRSQRT(x) rsqrt(x)
#define NRS_DRCP(x) nrs::direct::fast_rcp(x)
#endif

namespace nrs {
namespace direct {

// Device view of the plan (nrs_direct_plan.h) plus the factor storage.
struct Plan {
  int V, depth, G, max_path;
  int np;  // pose pseudo-vertices owned by the root (2, or 0 when the pose is fixed)
  const int *vb, *nv, *nbv, *bnd_ptr, *bnd, *bpath, *path_off, *inv_ptr, *inv;
  const long long *p_off, *u_off;
  double* panel;  // per node: (3 nv + 3 nbv) x 3 nv, row-major
  double* upd;    // per node: 3 nbv x 3 nbv, row-major, lower block triangle valid
  int* fail;      // bumped when a pivot is not positive (the reference's "Cholesky failure": solve() returns false)
};

// The linearised system H delta = b in global memory (written by the linearisation pass of nrs_direct.cu).
struct Sys {
  const double* dg;     // [8V]  symmetric 3x3 diagonal block: xx xy xz yy yz zz
  const double* cpl;    // [18V] pose coupling H_{pose, v}: 6 x 3 row-major
  const double* bvec;   // [4V]
  const double* pc;     // [4P]  pair block: H_ij = -(s I + u u^T)
  const double* hpp;    // [27]  H_pp upper packed (21) + b_p (6)
  const int *inc_ptr, *inc_ent, *inc_pos, *inc_row;
  double lambda;
};

struct Thr {
  int tid, nthr;
};

#ifndef NRS_DIRECT_HOST_EMULATION
// 1 / d to full double precision for normal positive d: hardware approximation + two Newton steps (the IEEE division
// of CUDA costs several times more and sits on the pivot chain of the factorisation).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}
#endif

constexpr int kPanel = 6;  // vertex columns per panel of the blocked factorisation

NRS_DD int sym6i(int a, int c) { return a * 6 - (a * (a - 1)) / 2 + (c - a); }  // a <= c

// Cholesky of a symmetric 3x3 (lower part read) and the inverse of its factor: W = L^-1 (lower).
NRS_DD bool chol3_inv(const double a00, const double a10, const double a11, const double a20, const double a21,
                      const double a22, double W[6]) {
  bool ok = a00 > 0;
  const double r0 = ok ? NRS_DRSQRT(a00) : 1.0;
  const double l10 = a10 * r0, l20 = a20 * r0;
  const double d1 = a11 - l10 * l10;
  ok = ok && d1 > 0;
  const double r1 = d1 > 0 ? NRS_DRSQRT(d1) : 1.0;
  const double l21 = (a21 - l20 * l10) * r1;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  ok = ok && d2 > 0;
  const double r2 = d2 > 0 ? NRS_DRSQRT(d2) : 1.0;
  W[0] = r0;                            // w00
  W[1] = -l10 * r0 * r1;                // w10
  W[2] = r1;                            // w11
  W[4] = -l21 * r1 * r2;                // w21
  W[3] = -(l20 * r0 + l21 * W[1]) * r2; // w20
  W[5] = r2;                            // w22
  return ok;
}

// Inverse of a symmetric positive definite 3x3 (lower part read) by cofactors: M = A^-1, symmetric, packed like the
// input (00 10 11 20 21 22). false when a leading principal minor is not positive.
NRS_DD bool inv3_spd(const double a00, const double a10, const double a11, const double a20, const double a21,
                     const double a22, double M[6]) {
  const double c00 = a11 * a22 - a21 * a21, c10 = a21 * a20 - a10 * a22, c20 = a10 * a21 - a11 * a20;
  const double c22 = a00 * a11 - a10 * a10;
  const double det = a00 * c00 + a10 * c10 + a20 * c20;
  const bool ok = a00 > 0 && c22 > 0 && det > 0;
  const double r = ok ? NRS_DRCP(det) : 1.0;
  M[0] = c00 * r;
  M[1] = c10 * r;
  M[2] = (a00 * a22 - a20 * a20) * r;
  M[3] = c20 * r;
  M[4] = (a10 * a20 - a00 * a21) * r;
  M[5] = c22 * r;
  return ok;
}

struct Front {
  int t, d, r, R;
  int nv, nbv, ns, ld;
  int nmy;         // boundary rows of this team member
  int rows_local;  // nv + nmy
};

NRS_DD Front front_of(const Plan& pl, int g, int d) {
  Front f;
  f.d = d;
  f.t = (1 << d) + (g >> (pl.depth - d));
  f.R = pl.G >> d;
  f.r = g & (f.R - 1);
  f.nv = NRS_DLDG(pl.nv + f.t);
  f.nbv = NRS_DLDG(pl.nbv + f.t);
  f.ns = 3 * f.nv;
  f.ld = f.ns | 1;
  f.nmy = (f.nbv > f.r) ? (f.nbv - f.r + f.R - 1) / f.R : 0;
  f.rows_local = f.nv + f.nmy;
  return f;
}

// local panel row of front position fp for this team member, -1 if the row belongs to another member
NRS_DD int local_row(const Front& f, int fp) {
  if (fp < f.nv) return fp;
  const int k = fp - f.nv;
  return ((k & (f.R - 1)) == f.r) ? f.nv + (k - f.r) / f.R : -1;
}
NRS_DD int front_pos(const Front& f, int li) { return li < f.nv ? li : f.nv + f.r + (li - f.nv) * f.R; }

// Factorisation of the diagonal region [ka, kb) x [ka, kb) (at most kPanel vertex blocks a side) of the panel in sp by
// ONE warp (nl lanes, warp-level syncs only): block L D L^T sweep with 3x3 pivots — per column M_k = D_k^-1 (cofactors,
// one reciprocal; kept in s_m), then block(li, j) -= B'(li, k) M_k B'(j, k)^T on the unscaled blocks B' — and at the
// end every block below the diagonal becomes L' = B' M_k. The diagonal blocks keep the pivots D_k.
NRS_DD void diag_region(const Plan& pl, double* sp, int ld, int ka, int kb, double* s_m, int lane, int nl) {
  const int nr = kb - ka;
  const int nblk = nr * (nr + 1) / 2;  // lower block triangle, row-major: (0,0) (1,0) (1,1) (2,0) ...
  for (int k = ka; k < kb; k++) {
    const double* P = sp + (size_t)(3 * k) * ld + 3 * k;
    double M[6];
    if (!inv3_spd(P[0], P[ld], P[ld + 1], P[2 * ld], P[2 * ld + 1], P[2 * ld + 2], M) && lane == 0) NRS_DFAIL(pl.fail);
    if (lane == 0)
      for (int i = 0; i < 6; i++) s_m[6 * k + i] = M[i];
    for (int q = lane; q < nblk; q += nl) {
      int ii = 0;
      while ((ii + 1) * (ii + 2) / 2 <= q) ii++;
      const int jj = q - ii * (ii + 1) / 2;
      const int li = ka + ii, j = ka + jj;
      if (j <= k) continue;
      const double* A = sp + (size_t)(3 * li) * ld + 3 * k;
      const double* B = sp + (size_t)(3 * j) * ld + 3 * k;
      double* o = sp + (size_t)(3 * li) * ld + 3 * j;
      double b[9];
      for (int t = 0; t < 3; t++) {
        b[3 * t] = B[t * ld];
        b[3 * t + 1] = B[t * ld + 1];
        b[3 * t + 2] = B[t * ld + 2];
      }
      for (int rr = 0; rr < 3; rr++) {
        const double a0 = A[rr * ld], a1 = A[rr * ld + 1], a2 = A[rr * ld + 2];
        const double t0 = a0 * M[0] + a1 * M[1] + a2 * M[3];
        const double t1 = a0 * M[1] + a1 * M[2] + a2 * M[4];
        const double t2 = a0 * M[3] + a1 * M[4] + a2 * M[5];
        for (int cc = 0; cc < 3; cc++) o[rr * ld + cc] -= t0 * b[3 * cc] + t1 * b[3 * cc + 1] + t2 * b[3 * cc + 2];
      }
    }
    NRS_DSYNCWARP();
  }
  for (int q = lane; q < nblk; q += nl) {
    int ii = 0;
    while ((ii + 1) * (ii + 2) / 2 <= q) ii++;
    const int kk = q - ii * (ii + 1) / 2;
    if (kk >= ii) continue;
    const double* M = s_m + 6 * (ka + kk);
    double* o = sp + (size_t)(3 * (ka + ii)) * ld + 3 * (ka + kk);
    for (int rr = 0; rr < 3; rr++) {
      const double b0 = o[rr * ld], b1 = o[rr * ld + 1], b2 = o[rr * ld + 2];
      o[rr * ld] = b0 * M[0] + b1 * M[1] + b2 * M[3];
      o[rr * ld + 1] = b0 * M[1] + b1 * M[2] + b2 * M[4];
      o[rr * ld + 2] = b0 * M[3] + b1 * M[4] + b2 * M[5];
    }
  }
  NRS_DSYNCWARP();
}

#ifndef NRS_DIRECT_HOST_EMULATION
// Device version of [special update +] diag_region for the look-ahead warp: the critical latency chain of the whole
// factorisation, so it is written for the fewest dependent instructions. Lane l < nblk owns ONE block (I, J) of the
// region's lower block triangle in registers from the rank update through the pivot sweep to the final scaling;
// per column the owners of column k publish their blocks to a small staging area (double-buffered by column parity),
// everybody reads pivot / A / B from there. kPanel <= 7 (28 blocks <= 32 lanes).
//   update: nc > 0 -> blk -= V(ka + I, 0:nc) L'(ka + J, k0cols)^T first (the rank-3 kPanel update of the previous panel)
__device__ __forceinline__ void diag_region_dev(const Plan& pl, double* sp, int ld, int ka, int kb, double* s_m,
                                                double* s_x, int lane, const double* s_v, int kvs, int k0, int nc) {
  const int nr = kb - ka;
  const int nblk = nr * (nr + 1) / 2;
  const bool act = lane < nblk;
  const int I = act ? (lane >= 1) + (lane >= 3) + (lane >= 6) + (lane >= 10) + (lane >= 15) + (lane >= 21) : 0;
  const int J = act ? lane - I * (I + 1) / 2 : 0;
  double* o = sp + (size_t)(3 * (ka + I)) * ld + 3 * (ka + J);
  double b[9];
#pragma unroll
  for (int rr = 0; rr < 3; rr++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) b[3 * rr + cc] = act ? o[rr * ld + cc] : 0.0;
  if (nc > 0 && act) {
    const double* A = s_v + (size_t)(ka + I) * (3 * kvs);
    const double* B = sp + (size_t)(3 * (ka + J)) * ld + 3 * k0;
#pragma unroll 4
    for (int c = 0; c < nc; c++) {
      const double a0 = A[c], a1 = A[kvs + c], a2 = A[2 * kvs + c];
      const double b0 = B[c], b1 = B[ld + c], b2 = B[2 * ld + c];
      b[0] -= a0 * b0; b[1] -= a0 * b1; b[2] -= a0 * b2;
      b[3] -= a1 * b0; b[4] -= a1 * b1; b[5] -= a1 * b2;
      b[6] -= a2 * b0; b[7] -= a2 * b1; b[8] -= a2 * b2;
    }
  }
  for (int k = 0; k < nr; k++) {
    double* xs = s_x + 10 * kPanel * (k & 1);  // [kPanel][10]
    if (act && J == k) {
      double* x = xs + 10 * I;
#pragma unroll
      for (int i = 0; i < 9; i++) x[i] = b[i];
    }
    __syncwarp();
    const double* P = xs + 10 * k;
    double M[6];
    const bool ok = inv3_spd(P[0], P[3], P[4], P[6], P[7], P[8], M);
    if (lane == 0) {
      if (!ok) NRS_DFAIL(pl.fail);
#pragma unroll
      for (int i = 0; i < 6; i++) s_m[6 * (ka + k) + i] = M[i];
    }
    if (act && J == k) {  // this block is final: pivot as is, below the diagonal L' = B' M
      if (I > k) {
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
          const double b0 = b[3 * rr], b1 = b[3 * rr + 1], b2 = b[3 * rr + 2];
          b[3 * rr] = b0 * M[0] + b1 * M[1] + b2 * M[3];
          b[3 * rr + 1] = b0 * M[1] + b1 * M[2] + b2 * M[4];
          b[3 * rr + 2] = b0 * M[3] + b1 * M[4] + b2 * M[5];
        }
      }
#pragma unroll
      for (int rr = 0; rr < 3; rr++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) o[rr * ld + cc] = b[3 * rr + cc];
    } else if (act && J > k) {
      const double* A = xs + 10 * I;
      const double* B = xs + 10 * J;
#pragma unroll
      for (int rr = 0; rr < 3; rr++) {
        const double a0 = A[3 * rr], a1 = A[3 * rr + 1], a2 = A[3 * rr + 2];
        const double t0 = a0 * M[0] + a1 * M[1] + a2 * M[3];
        const double t1 = a0 * M[1] + a1 * M[2] + a2 * M[4];
        const double t2 = a0 * M[3] + a1 * M[4] + a2 * M[5];
#pragma unroll
        for (int cc = 0; cc < 3; cc++) b[3 * rr + cc] -= t0 * B[3 * cc] + t1 * B[3 * cc + 1] + t2 * B[3 * cc + 2];
      }
    }
  }
  __syncwarp();
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// Stage AB: assemble + factorise + store. sp: rows_local*3 x ld doubles; s_w: 6 nv doubles (M_k); s_v: rows_local x
// (3 kPanel + 1) doubles per scalar row (the unscaled panel rows the trailing update multiplies with).
// A member without boundary rows that is not the leader has nothing to do.
// ---------------------------------------------------------------------------------------------------------------
NRS_DD void stage_ab(const Plan& pl, const Sys& sys, int g, int d, double* sp, double* s_w, double* s_v, Thr th,
                     long long* pf = nullptr) {
  const long long c0 = NRS_DCLOCK();
  const Front f = front_of(pl, g, d);
  if (f.nmy == 0 && f.r != 0) return;
  const int V = pl.V, ld = f.ld, nv = f.nv, rows = f.rows_local;
  const int vb = NRS_DLDG(pl.vb + f.t);
  const int T = (2 << pl.depth) - 1;
  // (1) zero
  for (int q = th.tid; q < 3 * rows * ld; q += th.nthr) sp[q] = 0.0;
  NRS_DSYNC();
  // (2) original entries of the pivot columns: one thread per own vertex
  const int fp_rhs = nv + f.nbv - 1;
  const int fp_pose = (f.t == 1) ? nv - 2 : nv + f.nbv - 3;
  const int li_rhs = local_row(f, fp_rhs);
  for (int j = th.tid; j < nv; j += th.nthr) {
    const int v = vb + j;
    double* col = sp + 3 * j;
    if (v < V) {
      const double* D = sys.dg + 8 * (size_t)v;
      const double d0 = NRS_DLDCG(D), d1 = NRS_DLDCG(D + 1), d2 = NRS_DLDCG(D + 2), d3 = NRS_DLDCG(D + 3),
                   d4 = NRS_DLDCG(D + 4), d5 = NRS_DLDCG(D + 5);
      double* b = col + (size_t)(3 * j) * ld;
      b[0] = d0 + sys.lambda; b[1] = d1; b[2] = d2;
      b[ld] = d1; b[ld + 1] = d3 + sys.lambda; b[ld + 2] = d4;
      b[2 * ld] = d2; b[2 * ld + 1] = d4; b[2 * ld + 2] = d5 + sys.lambda;
      for (int a = 0; a < pl.np; a++) {
        const int li = local_row(f, fp_pose + a);
        if (li < 0) continue;
        double* o = col + (size_t)(3 * li) * ld;
        const double* c = sys.cpl + 18 * (size_t)v + 9 * a;
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) o[rr * ld + cc] = NRS_DLDCG(c + 3 * rr + cc);
      }
      if (li_rhs >= 0) {
        double* o = col + (size_t)(3 * li_rhs) * ld;
        const double* bb = sys.bvec + 4 * (size_t)v;
        o[0] = NRS_DLDCG(bb);
        o[1] = NRS_DLDCG(bb + 1);
        o[2] = NRS_DLDCG(bb + 2);
      }
    } else {  // pose pseudo-vertex (root only): a = 0 rotation, 1 translation
      const int a = v - V;
      double* b = col + (size_t)(3 * j) * ld;
      for (int rr = 0; rr < 3; rr++)
        for (int cc = 0; cc < 3; cc++) {
          const int p = 3 * a + rr, q = 3 * a + cc;
          b[rr * ld + cc] = NRS_DLDCG(sys.hpp + sym6i(p < q ? p : q, p < q ? q : p)) + (rr == cc ? sys.lambda : 0.0);
        }
      if (a == 0) {  // block (translation, rotation)
        double* o = col + (size_t)(3 * (j + 1)) * ld;
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) o[rr * ld + cc] = NRS_DLDCG(sys.hpp + sym6i(cc, 3 + rr));
      }
      if (li_rhs >= 0) {
        double* o = col + (size_t)(3 * li_rhs) * ld;
        for (int cc = 0; cc < 3; cc++) o[cc] = NRS_DLDCG(sys.hpp + 21 + 3 * a + cc);
      }
    }
  }
  {  // pair blocks of the pivot columns: one thread per incidence of an own vertex (they are contiguous)
    const int npts = (vb + nv <= V) ? nv : V - vb;
    const int a0 = npts > 0 ? NRS_DLDG(sys.inc_ptr + vb) : 0, a1 = npts > 0 ? NRS_DLDG(sys.inc_ptr + vb + npts) : 0;
    for (int a = a0 + th.tid; a < a1; a += th.nthr) {
      const int fp = NRS_DLDG(sys.inc_pos + a);
      if (fp < 0) continue;
      const int li = local_row(f, fp);
      if (li < 0) continue;
      const int j = NRS_DLDG(sys.inc_row + a) - vb;
      const double* c = sys.pc + 4 * (size_t)(NRS_DLDG(sys.inc_ent + a) >> 1);
      const double s = NRS_DLDCG(c), u0 = NRS_DLDCG(c + 1), u1 = NRS_DLDCG(c + 2), u2 = NRS_DLDCG(c + 3);
      double* o = sp + 3 * j + (size_t)(3 * li) * ld;
      o[0] = -(s + u0 * u0); o[1] = -(u0 * u1); o[2] = -(u0 * u2);
      o[ld] = -(u1 * u0); o[ld + 1] = -(s + u1 * u1); o[ld + 2] = -(u1 * u2);
      o[2 * ld] = -(u2 * u0); o[2 * ld + 1] = -(u2 * u1); o[2 * ld + 2] = -(s + u2 * u2);
    }
  }
  NRS_DSYNC();
  const long long c1 = NRS_DCLOCK();
  // (3) children's update matrices (pull: every panel block gathers from both children)
  if (2 * f.t <= T) {
    for (int c = 2 * f.t; c <= 2 * f.t + 1; c++) {
      const int* iv = pl.inv + NRS_DLDG(pl.inv_ptr + c);
      const int ldc = 3 * NRS_DLDG(pl.nbv + c);
      const double* U = pl.upd + NRS_DLDG(pl.u_off + c);
      for (int q = th.tid; q < rows * nv; q += th.nthr) {
        const int li = q / nv, j = q - li * nv;
        if (li < j) continue;
        const int cj = NRS_DLDG(iv + j);
        if (cj < 0) continue;
        const int ci = NRS_DLDG(iv + front_pos(f, li));
        if (ci < 0) continue;
        const double* u = U + (size_t)(3 * ci) * ldc + 3 * cj;
        double* o = sp + (size_t)(3 * li) * ld + 3 * j;
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) o[rr * ld + cc] += NRS_DLDCG(u + (size_t)rr * ldc + cc);
      }
      NRS_DSYNC();
    }
  }
  const long long c2 = NRS_DCLOCK();
  // (4) blocked right-looking block L D L^T of the tall panel, kPanel vertex columns (3 kPanel scalars) per step and
  // TWO block syncs per step. With [k0, k1) the current panel and [k1, k2) the next one:
  //   R  every row block below the panel's diagonal region solves against it (thread per row, registers):
  //      B'_k = A_k - sum_{k' < k} B'_k' L'(k, k')^T, L'_k = B'_k M_k; the unscaled B' goes to the scratch s_v;
  //   T  rank-(3 kPanel) update of everything right of the panel, block(li, j) -= B'(li, panel) L'(j, panel)^T, in 3x6
  //      register tiles (18 accumulators, 9 shared-memory loads per 18 FMAs) — while warp 0 LOOKS AHEAD: it updates
  //      the next panel's diagonal region first and factorises it (diag_region), so the pivot chain of panel p+1
  //      overlaps the bulk update of panel p.
  constexpr int kVS = 3 * kPanel + 1;                              // row stride of s_v (odd: conflict-free)
  const int nl = th.nthr < 32 ? th.nthr : 32;                      // lanes of the look-ahead warp
  const bool la_warp = th.tid < nl;                                // this thread belongs to it
  const int gt = (th.nthr >= 64) ? th.tid - 32 : th.tid;           // bulk threads: everybody else (all, if one warp)
  const int gn = (th.nthr >= 64) ? th.nthr - 32 : th.nthr;
#ifndef NRS_DIRECT_HOST_EMULATION
  static_assert(kPanel <= 7, "diag_region_dev keeps one block of the region per lane (28 blocks <= 32 lanes)");
  double* s_x = s_w + 6 * (((nv + 1) & ~1) + 1);  // staging area of the look-ahead warp: 2 x [kPanel][10] doubles
  if (la_warp && nv > 0)
    diag_region_dev(pl, sp, ld, 0, nv < kPanel ? nv : kPanel, s_w, s_x, th.tid, s_v, kVS, 0, 0);
#else
  if (la_warp && nv > 0) diag_region(pl, sp, ld, 0, nv < kPanel ? nv : kPanel, s_w, th.tid, nl);
#endif
  NRS_DSYNC();
  for (int k0 = 0; k0 < nv; k0 += kPanel) {
    const int k1 = (k0 + kPanel < nv) ? k0 + kPanel : nv;
    const int k2 = (k1 + kPanel < nv) ? k1 + kPanel : nv;
    const int nk = k1 - k0;
    // ---- R
    const long long r0 = NRS_DCLOCK();
    for (int q = 3 * k1 + th.tid; q < 3 * rows; q += th.nthr) {  // one thread per scalar row of the panel
      const int li = q / 3;
      double* o = sp + (size_t)q * ld + 3 * k0;
      double* vo = s_v + (size_t)q * kVS;
      double X[kPanel][3];
#pragma unroll
      for (int kk = 0; kk < kPanel; kk++)
        if (kk < nk) {
          X[kk][0] = o[3 * kk];
          X[kk][1] = o[3 * kk + 1];
          X[kk][2] = o[3 * kk + 2];
        }
#pragma unroll
      for (int kk = 0; kk < kPanel; kk++) {
        if (kk < nk) {
          const int k = k0 + kk;
#pragma unroll
          for (int kp = 0; kp < kPanel; kp++) {
            if (kp < kk) {  // B'_kk -= B'_kp L'(k, k0 + kp)^T
              const double* Lk = sp + (size_t)(3 * k) * ld + 3 * (k0 + kp);
#pragma unroll
              for (int cc = 0; cc < 3; cc++)
                X[kk][cc] -= X[kp][0] * Lk[cc * ld] + X[kp][1] * Lk[cc * ld + 1] + X[kp][2] * Lk[cc * ld + 2];
            }
          }
        }
      }
#pragma unroll
      for (int kk = 0; kk < kPanel; kk++)
        if (kk < nk) {
          const double* M = s_w + 6 * (k0 + kk);
          const double b0 = X[kk][0], b1 = X[kk][1], b2 = X[kk][2];
          vo[3 * kk] = b0;
          vo[3 * kk + 1] = b1;
          vo[3 * kk + 2] = b2;
          o[3 * kk] = b0 * M[0] + b1 * M[1] + b2 * M[3];
          o[3 * kk + 1] = b0 * M[1] + b1 * M[2] + b2 * M[4];
          o[3 * kk + 2] = b0 * M[3] + b1 * M[4] + b2 * M[5];
        }
      (void)li;
    }
    NRS_DSYNC();
    const long long r1 = NRS_DCLOCK();
    if (pf) pf[13] += r1 - r0;
    if (k1 >= nv) break;
    const int nc = 3 * nk;
    // ---- T, look-ahead part: blocks (li, j) of the next diagonal region, k1 <= j <= li < k2, then its factorisation
#ifndef NRS_DIRECT_HOST_EMULATION
    if (la_warp) {
      diag_region_dev(pl, sp, ld, k1, k2, s_w, s_x, th.tid, s_v, kVS, k0, nc);
      if (pf) pf[14] += NRS_DCLOCK() - r1;
    }
#else
    if (la_warp) {
      const int nr = k2 - k1;
      const int nblk = nr * (nr + 1) / 2;
      for (int q = th.tid; q < nblk; q += nl) {
        int ii = 0;
        while ((ii + 1) * (ii + 2) / 2 <= q) ii++;
        const int jj = q - ii * (ii + 1) / 2;
        const int li = k1 + ii, j = k1 + jj;
        const double* A = s_v + (size_t)li * (3 * kVS);
        const double* B = sp + (size_t)(3 * j) * ld + 3 * k0;
        double* o = sp + (size_t)(3 * li) * ld + 3 * j;
        double acc[9];
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) acc[3 * rr + cc] = o[rr * ld + cc];
        for (int c = 0; c < nc; c++) {
          const double a0 = A[c], a1 = A[kVS + c], a2 = A[2 * kVS + c];
          const double b0 = B[c], b1 = B[ld + c], b2 = B[2 * ld + c];
          acc[0] -= a0 * b0; acc[1] -= a0 * b1; acc[2] -= a0 * b2;
          acc[3] -= a1 * b0; acc[4] -= a1 * b1; acc[5] -= a1 * b2;
          acc[6] -= a2 * b0; acc[7] -= a2 * b1; acc[8] -= a2 * b2;
        }
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) o[rr * ld + cc] = acc[3 * rr + cc];
      }
      NRS_DSYNCWARP();
      diag_region(pl, sp, ld, k1, k2, s_w, th.tid, nl);
    }
#endif
    // ---- T, bulk: rows li >= k2, column pairs from k1
    if (!la_warp || th.nthr < 64) {
      const int ntj = (nv - k1 + 1) >> 1, w = rows - k2;
      for (int q = gt; q < ntj * w; q += gn) {
        const int jp = q / w, ii = q - jp * w;
        const int j = k1 + 2 * jp, li = k2 + ii;
        if (li < j) continue;
        const bool two = (j + 1 < nv) && (li >= j + 1);
        double* o = sp + (size_t)(3 * li) * ld + 3 * j;
        double acc[18];
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 6; cc++) acc[6 * rr + cc] = (cc < 3 || two) ? o[rr * ld + cc] : 0.0;
        const double* A = s_v + (size_t)li * (3 * kVS);
        const double* B = sp + (size_t)(3 * j) * ld + 3 * k0;
        for (int c = 0; c < nc; c++) {
          const double a0 = A[c], a1 = A[kVS + c], a2 = A[2 * kVS + c];
          const double b0 = B[c], b1 = B[ld + c], b2 = B[2 * ld + c];
          acc[0] -= a0 * b0; acc[1] -= a0 * b1; acc[2] -= a0 * b2;
          acc[6] -= a1 * b0; acc[7] -= a1 * b1; acc[8] -= a1 * b2;
          acc[12] -= a2 * b0; acc[13] -= a2 * b1; acc[14] -= a2 * b2;
          if (two) {
            const double b3 = B[3 * ld + c], b4 = B[4 * ld + c], b5 = B[5 * ld + c];
            acc[3] -= a0 * b3; acc[4] -= a0 * b4; acc[5] -= a0 * b5;
            acc[9] -= a1 * b3; acc[10] -= a1 * b4; acc[11] -= a1 * b5;
            acc[15] -= a2 * b3; acc[16] -= a2 * b4; acc[17] -= a2 * b5;
          }
        }
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < (two ? 6 : 3); cc++) o[rr * ld + cc] = acc[6 * rr + cc];
      }
    }
    NRS_DSYNC();
  }
  const long long c3 = NRS_DCLOCK();
  // (5) store: the leader writes L11, every member its boundary rows
  double* Pg = pl.panel + NRS_DLDG(pl.p_off + f.t);
  const int ns = f.ns;
  const int first = (f.r == 0) ? 0 : 3 * nv;
  for (int q = first * ns + th.tid; q < 3 * rows * ns; q += th.nthr) {
    const int row = q / ns, c = q - row * ns;
    const int li = row / 3, a = row - 3 * li;
    Pg[(size_t)(3 * front_pos(f, li) + a) * ns + c] = sp[(size_t)row * ld + c];
  }
  if (pf) {
    const long long c4 = NRS_DCLOCK();
    pf[3] += c1 - c0;
    pf[4] += c2 - c1;
    pf[8] += c3 - c2;
    pf[9] += c4 - c3;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Stage C: rows k = r (mod R) of U = sum_children U_c - (L21' D) L21'^T. sp: 3 (kmax + 1) x ld doubles (L21'), then
// 3 nmy x ld doubles (this member's rows times D).
// ---------------------------------------------------------------------------------------------------------------
NRS_DD void stage_c(const Plan& pl, int g, int d, double* sp, Thr th) {
  const Front f = front_of(pl, g, d);
  if (f.nmy == 0) return;
  const int ld = f.ld, nv = f.nv, ns = f.ns, nbv = f.nbv;
  const int T = (2 << pl.depth) - 1;
  const int kmax = f.r + (f.nmy - 1) * f.R;
  const double* Pg = pl.panel + NRS_DLDG(pl.p_off + f.t) + (size_t)(3 * nv) * ns;
  for (int q = th.tid; q < 3 * (kmax + 1) * ns; q += th.nthr) {
    const int row = q / ns, c = q - row * ns;
    sp[(size_t)row * ld + c] = NRS_DLDCG(Pg + q);
  }
  NRS_DSYNC();
  // V = L21'(my rows) D: one thread per (my row, pivot block); the pivots sit on the diagonal of L11
  double* sv = sp + (size_t)(3 * (kmax + 1)) * ld;
  {
    const double* Pd = pl.panel + NRS_DLDG(pl.p_off + f.t);
    for (int q = th.tid; q < f.nmy * nv; q += th.nthr) {
      const int m = q / nv, kk = q - m * nv;
      const int k = f.r + m * f.R;
      const double* Dk = Pd + (size_t)(3 * kk) * ns + 3 * kk;
      const double d00 = NRS_DLDCG(Dk), d10 = NRS_DLDCG(Dk + ns), d11 = NRS_DLDCG(Dk + ns + 1),
                   d20 = NRS_DLDCG(Dk + 2 * (size_t)ns), d21 = NRS_DLDCG(Dk + 2 * (size_t)ns + 1),
                   d22 = NRS_DLDCG(Dk + 2 * (size_t)ns + 2);
      const double* Lr = sp + (size_t)(3 * k) * ld + 3 * kk;
      double* o = sv + (size_t)(3 * m) * ld + 3 * kk;
      for (int rr = 0; rr < 3; rr++) {
        const double b0 = Lr[rr * ld], b1 = Lr[rr * ld + 1], b2 = Lr[rr * ld + 2];
        o[rr * ld] = b0 * d00 + b1 * d10 + b2 * d20;
        o[rr * ld + 1] = b0 * d10 + b1 * d11 + b2 * d21;
        o[rr * ld + 2] = b0 * d20 + b1 * d21 + b2 * d22;
      }
    }
  }
  NRS_DSYNC();
  const bool kids = 2 * f.t <= T;
  const int* iv0 = nullptr;
  const int* iv1 = nullptr;
  const double *U0 = nullptr, *U1 = nullptr;
  int ld0 = 0, ld1 = 0;
  if (kids) {
    iv0 = pl.inv + NRS_DLDG(pl.inv_ptr + 2 * f.t) + nv;
    iv1 = pl.inv + NRS_DLDG(pl.inv_ptr + 2 * f.t + 1) + nv;
    U0 = pl.upd + NRS_DLDG(pl.u_off + 2 * f.t);
    U1 = pl.upd + NRS_DLDG(pl.u_off + 2 * f.t + 1);
    ld0 = 3 * NRS_DLDG(pl.nbv + 2 * f.t);
    ld1 = 3 * NRS_DLDG(pl.nbv + 2 * f.t + 1);
  }
  double* Ug = pl.upd + NRS_DLDG(pl.u_off + f.t);
  const int ldu = 3 * nbv;
  for (int q = th.tid; q < f.nmy * nbv; q += th.nthr) {
    const int m = q / nbv, j = q - m * nbv;
    const int k = f.r + m * f.R;
    if (j > k) continue;
    double acc[9];
    for (int i = 0; i < 9; i++) acc[i] = 0.0;
    if (kids) {
      const int c0k = NRS_DLDG(iv0 + k), c0j = NRS_DLDG(iv0 + j);
      if (c0k >= 0 && c0j >= 0) {
        const double* u = U0 + (size_t)(3 * c0k) * ld0 + 3 * c0j;
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) acc[3 * rr + cc] += NRS_DLDCG(u + (size_t)rr * ld0 + cc);
      }
      const int c1k = NRS_DLDG(iv1 + k), c1j = NRS_DLDG(iv1 + j);
      if (c1k >= 0 && c1j >= 0) {
        const double* u = U1 + (size_t)(3 * c1k) * ld1 + 3 * c1j;
        for (int rr = 0; rr < 3; rr++)
          for (int cc = 0; cc < 3; cc++) acc[3 * rr + cc] += NRS_DLDCG(u + (size_t)rr * ld1 + cc);
      }
    }
    const double* A = sv + (size_t)(3 * m) * ld;
    const double* B = sp + (size_t)(3 * j) * ld;
    double s[9];
    for (int i = 0; i < 9; i++) s[i] = 0.0;
    for (int c = 0; c < ns; c++) {
      const double a0 = A[c], a1 = A[ld + c], a2 = A[2 * ld + c];
      const double b0 = B[c], b1 = B[ld + c], b2 = B[2 * ld + c];
      s[0] += a0 * b0; s[1] += a0 * b1; s[2] += a0 * b2;
      s[3] += a1 * b0; s[4] += a1 * b1; s[5] += a1 * b2;
      s[6] += a2 * b0; s[7] += a2 * b1; s[8] += a2 * b2;
    }
    double* o = Ug + (size_t)(3 * k) * ldu + 3 * j;
    for (int rr = 0; rr < 3; rr++)
      for (int cc = 0; cc < 3; cc++) o[(size_t)rr * ldu + cc] = acc[3 * rr + cc] - s[3 * rr + cc];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward substitution of the front of level d on CTA g's path. s_path: the solution of the ancestors' own
// vertices (and, on return, of this front's). sp: 3 nv x ld (L11); s_z: ns + nparts * ns doubles.
// The leader also writes the point rows of delta (4-double stride) and the pose delta.
// ---------------------------------------------------------------------------------------------------------------
NRS_DD void backward_front(const Plan& pl, int g, int d, double* s_path, double* sp, double* s_z, double* delta,
                           double* dpose, Thr th, long long* pf = nullptr, uint64_t* mbar = nullptr,
                           unsigned* mphase = nullptr) {
  const long long c0 = NRS_DCLOCK();
  const Front f = front_of(pl, g, d);
  const int nv = f.nv, ns = f.ns, nbv = f.nbv;
  const int ld = ns;  // L11 is copied as it lies in global memory (rows are read along columns: no bank conflicts)
  if (nv == 0) return;
  const double* Pg = pl.panel + NRS_DLDG(pl.p_off + f.t);
  // L11 -> shared memory. Device: ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) issued by thread 0; it
  // overlaps the L21^T x product below, which streams L21 from L2.
#ifndef NRS_DIRECT_HOST_EMULATION
  const unsigned bytes = ((unsigned)(ns * ns) * 8u + 15u) & ~15u;
  if (th.tid == 0) {
    cuda::ptx::fence_proxy_async();  // earlier generic-proxy accesses of sp / the panel are ordered before the copy
    cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, mbar,
                                         bytes);
    cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, sp, Pg, bytes, mbar);
  }
#else
  for (int q = th.tid; q < ns * ns; q += th.nthr) sp[q] = Pg[q];
  (void)mbar;
  (void)mphase;
#endif
  // z = y - L21^T x_boundary: groups of threads split the boundary rows, lanes the columns
  const int nrows = 3 * (nbv - 1);
  const int lanes = th.nthr >= 32 ? 32 : th.nthr;
  const int ngrp = th.nthr / lanes, grp = th.tid / lanes, lane = th.tid - grp * lanes;
  const double* L21 = Pg + (size_t)ns * ns;
  const int* bp = pl.bpath + NRS_DLDG(pl.bnd_ptr + f.t);
  double* part = s_z + ns;
  for (int c = lane; c < ns; c += lanes) {
    double s = 0;
    for (int row = grp; row < nrows; row += 4 * ngrp) {  // four independent loads in flight per thread
      double l[4], x[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int r = row + u * ngrp;
        const int rc = r < nrows ? r : nrows - 1;
        const int k = rc / 3;
        l[u] = NRS_DLDCG(L21 + (size_t)rc * ns + c);
        x[u] = (r < nrows) ? s_path[NRS_DLDG(bp + k) + (rc - 3 * k)] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) s += l[u] * x[u];
    }
    part[grp * ns + c] = s;
  }
  NRS_DSYNC();
  const long long c1 = NRS_DCLOCK();
  const double* y = L21 + (size_t)nrows * ns;  // first scalar row of the rhs block row
  for (int c = th.tid; c < ns; c += th.nthr) {
    double s = NRS_DLDCG(y + c);
    for (int w = 0; w < ngrp; w++) s -= part[w * ns + c];
    s_z[c] = s;
  }
#ifndef NRS_DIRECT_HOST_EMULATION
  {  // the bulk copy has landed (bounded spin: a lost copy must not hang the grid)
    const unsigned ph = *mphase;
    for (int spin = 0; spin < (1 << 22); spin++)
      if (cuda::ptx::mbarrier_try_wait_parity(mbar, ph)) break;
    *mphase = ph ^ 1u;
  }
#endif
  NRS_DSYNC();
  // L11'^T x = z with the unit lower block-triangular L11', block row by block row from the bottom (the rhs row of the
  // panel already holds D^-1 L'^-1 b, see the header). Rows whose update fits one warp's lanes are done by warp 0
  // alone with warp-level syncs; the others wait at one block sync.
  double* xo = s_path + NRS_DLDG(pl.path_off + f.t);
  const int wl = th.nthr >= 32 ? 32 : th.nthr;
  int i = nv - 1;
  for (; i >= 0 && 3 * i > 4 * wl; i--) {
    const double x0 = s_z[3 * i], x1 = s_z[3 * i + 1], x2 = s_z[3 * i + 2];
    const double* Li = sp + (size_t)(3 * i) * ld;
    for (int c = th.tid; c < 3 * i; c += th.nthr) s_z[c] -= Li[c] * x0 + Li[ld + c] * x1 + Li[2 * ld + c] * x2;
    if (th.tid == 0) {
      xo[3 * i] = x0;
      xo[3 * i + 1] = x1;
      xo[3 * i + 2] = x2;
    }
    NRS_DSYNC();
  }
  if (th.tid < wl) {
    for (; i >= 0; i--) {
      const double x0 = s_z[3 * i], x1 = s_z[3 * i + 1], x2 = s_z[3 * i + 2];
      const double* Li = sp + (size_t)(3 * i) * ld;
      for (int c = th.tid; c < 3 * i; c += wl) s_z[c] -= Li[c] * x0 + Li[ld + c] * x1 + Li[2 * ld + c] * x2;
      if (th.tid == 0) {
        xo[3 * i] = x0;
        xo[3 * i + 1] = x1;
        xo[3 * i + 2] = x2;
      }
      NRS_DSYNCWARP();
    }
  }
  NRS_DSYNC();
  if (pf) {
    pf[10] += c1 - c0;
    pf[11] += NRS_DCLOCK() - c1;
  }
  if (f.r == 0) {
    const int vb = NRS_DLDG(pl.vb + f.t);
    for (int q = th.tid; q < ns; q += th.nthr) {
      const int j = q / 3, a = q - 3 * j, v = vb + j;
      if (v < pl.V)
        delta[4 * (size_t)v + a] = xo[q];
      else
        dpose[3 * (v - pl.V) + a] = xo[q];
    }
  }
}

// Shared-memory doubles backward_front needs besides sp and s_path.
NRS_DD int backward_scratch(int max_ns, int nthr) { return max_ns * (1 + (nthr >= 32 ? nthr / 32 : 1)); }

}  // namespace direct
}  // namespace nrs
