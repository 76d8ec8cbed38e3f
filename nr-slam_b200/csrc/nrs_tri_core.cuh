// nrs_tri_core.cuh — one DeformableTriangulation problem solved by one CTA.
//
// Reference: modules/optimization/g2o_optimization.cc:559-814 (driver), optimization/
// reprojection_error_only_deformation.cc:33-39 (numerically differentiated by g2o, core/base_fixed_sized_edge.hpp:
// 160-199), optimization/spatial_regularizer_with_observation.cc:33-51, utilities/geometry_toolbox.cc:31-78,
// third_party/Sophus/sophus/{so3,se3}.hpp (fp32 group operations), third_party/g2o/g2o/core/
// optimization_algorithm_levenberg.cpp:57-165 (LM control flow).
//
// Structure of the problem (what makes one CTA per candidate the right granularity): T <= 48 point vertices (the
// candidate's position in every frame of its track, camera coordinates), T reprojection edges (unary) and one
// SpatialRegularizerWithObservation edge per (frame pair a < b, neighbour valid in a, b and the first frame). Every
// spatial edge has Jacobians +-I, so the Hessian is  omega (L (x) I3) + blockdiag(4 Jr^T Jr) + lambda I  with L the
// integer Laplacian of the pair counts c_ab: it is NEVER assembled edge by edge — the counts are computed once, the
// dense (3T)^2 lower triangle is generated in shared memory per trial and factorised there (right-looking LL^T,
// fp64). Residuals are evaluated edge by edge (they involve the real rotations), spread over the CTA.
//
// The code between two TRI_SYNC()s is data-race free, so the same source compiled for the host with one "thread"
// (tests/emul/tri_emul.cc, TEST INFRASTRUCTURE) executes the identical arithmetic sequentially: that is how the
// algorithm is checked against the oracle on machines without a GPU. The product only ever runs the CUDA build.
#pragma once
#include "nrs_math.cuh"

#if defined(__CUDA_ARCH__)
#define TRI_TID ((int)threadIdx.x)
#define TRI_NT ((int)blockDim.x)
#define TRI_SYNC() __syncthreads()
#define TRI_DEV __device__ __forceinline__
#elif defined(NRS_TRI_HOST_EMULATION)
#define TRI_TID 0
#define TRI_NT 1
#define TRI_SYNC() ((void)0)
#define TRI_DEV inline
#else
#define TRI_TID 0
#define TRI_NT 1
#define TRI_SYNC() ((void)0)
#define TRI_DEV __device__ __forceinline__
#endif

// fp64 product / sum with one rounding each (no FMA contraction) where two code paths must produce identical bits
#if defined(__CUDA_ARCH__)
#define TRI_DMUL(a, b) __dmul_rn((a), (b))
#define TRI_DADD(a, b) __dadd_rn((a), (b))
#define TRI_FFS(m) __ffs((int)(m))
#else
#define TRI_FFS(m) __builtin_ffs((int)(m))
#define TRI_DMUL(a, b) ((a) * (b))
#define TRI_DADD(a, b) ((a) + (b))
#endif

namespace nrs {
namespace tri {

constexpr int kMaxTrack = 48;  // NRSLAM_B200_TRI_MAX_TRACK
constexpr int kNB = 12;        // NRSLAM_B200_TRI_MAX_NB
constexpr int kThreads = 128;

enum {
  ST_OK = 0, ST_TOO_CLOSE = 1, ST_HIGH_REPROJ_FIRST = 2, ST_HIGH_REPROJ_SECOND = 3, ST_LOW_PARALLAX = 4,
  ST_NO_NEIGHBOURS = 5, ST_NEGATIVE_DEPTH = 6, ST_BAD_NEIGHBOURS = 7, ST_HIGH_ERROR = 8, ST_NAN = 9,
  ST_SHORT_TRACK = 10, ST_NOT_RIGID = 11, ST_RIGID_PARALLAX = 12
};

// The rigid branch of Mapping::LandmarkTriangulation (mapping/mapping.cc:115-185) rides on the two-view quantities the
// deformable pre-checks compute anyway. enabled == 0: DeformableTriangulation alone (nrslam_b200_tri_run).
struct RigidArgs {
  int enabled;
  int min_track;         // TrackLenght(candidate) >= 5 gate of the deformable branch (mapping.cc:94)
  float rad_per_pixel;   // Mapping::Options::rad_per_pixel
  int rigid_ok;          // TemporalBuffer::CheckRigidity(first, last, 0.004) of this candidate's track (mapping.cc:123)
  float* out;            // [3] rigid position
  int* status;           // ST_OK / ST_TOO_CLOSE / ST_NOT_RIGID / ST_RIGID_PARALLAX
};

// ---- fp32 Sophus / Eigen restatements (no FMA contraction: NRS_F* are single-rounding intrinsics on the device) ----
struct SE3f {
  float q[4];  // x y z w
  float t[3];
};
// so3.hpp:388-397
TRI_DEV void rot_f(const float* q, const float* p, float* o) {
  float u0 = NRS_FS(NRS_FM(q[1], p[2]), NRS_FM(q[2], p[1]));
  float u1 = NRS_FS(NRS_FM(q[2], p[0]), NRS_FM(q[0], p[2]));
  float u2 = NRS_FS(NRS_FM(q[0], p[1]), NRS_FM(q[1], p[0]));
  u0 = NRS_FA(u0, u0);
  u1 = NRS_FA(u1, u1);
  u2 = NRS_FA(u2, u2);
  const float c0 = NRS_FS(NRS_FM(q[1], u2), NRS_FM(q[2], u1));
  const float c1 = NRS_FS(NRS_FM(q[2], u0), NRS_FM(q[0], u2));
  const float c2 = NRS_FS(NRS_FM(q[0], u1), NRS_FM(q[1], u0));
  o[0] = NRS_FA(NRS_FA(p[0], NRS_FM(q[3], u0)), c0);
  o[1] = NRS_FA(NRS_FA(p[1], NRS_FM(q[3], u1)), c1);
  o[2] = NRS_FA(NRS_FA(p[2], NRS_FM(q[3], u2)), c2);
}
TRI_DEV void normalize_q(float* q) {  // so3.hpp:318-325
  const float len = sqrtf(NRS_FA(NRS_FA(NRS_FA(NRS_FM(q[0], q[0]), NRS_FM(q[1], q[1])), NRS_FM(q[2], q[2])), NRS_FM(q[3], q[3])));
  for (int i = 0; i < 4; i++) q[i] = NRS_FD(q[i], len);
}
TRI_DEV SE3f inverse_f(const SE3f& T) {  // se3.hpp:222-225
  SE3f r;
  r.q[0] = -T.q[0];
  r.q[1] = -T.q[1];
  r.q[2] = -T.q[2];
  r.q[3] = T.q[3];
  normalize_q(r.q);
  const float mt[3] = {NRS_FM(T.t[0], -1.f), NRS_FM(T.t[1], -1.f), NRS_FM(T.t[2], -1.f)};
  rot_f(r.q, mt, r.t);
  return r;
}
TRI_DEV SE3f mul_f(const SE3f& a, const SE3f& b) {  // se3.hpp:302-306, so3.hpp:346-369
  SE3f r;
  const float *A = a.q, *B = b.q;
  r.q[3] = NRS_FS(NRS_FS(NRS_FS(NRS_FM(A[3], B[3]), NRS_FM(A[0], B[0])), NRS_FM(A[1], B[1])), NRS_FM(A[2], B[2]));
  r.q[0] = NRS_FS(NRS_FA(NRS_FA(NRS_FM(A[3], B[0]), NRS_FM(A[0], B[3])), NRS_FM(A[1], B[2])), NRS_FM(A[2], B[1]));
  r.q[1] = NRS_FS(NRS_FA(NRS_FA(NRS_FM(A[3], B[1]), NRS_FM(A[1], B[3])), NRS_FM(A[2], B[0])), NRS_FM(A[0], B[2]));
  r.q[2] = NRS_FS(NRS_FA(NRS_FA(NRS_FM(A[3], B[2]), NRS_FM(A[2], B[3])), NRS_FM(A[0], B[1])), NRS_FM(A[1], B[0]));
  normalize_q(r.q);
  float rt[3];
  rot_f(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = NRS_FA(a.t[i], rt[i]);
  return r;
}
TRI_DEV void map_f(const SE3f& T, const float* p, float* o) {
  rot_f(T.q, p, o);
  for (int i = 0; i < 3; i++) o[i] = NRS_FA(o[i], T.t[i]);
}
TRI_DEV void quat_to_R_f(const float* q, float* R) {
  const float tx = NRS_FM(2.f, q[0]), ty = NRS_FM(2.f, q[1]), tz = NRS_FM(2.f, q[2]);
  const float twx = NRS_FM(tx, q[3]), twy = NRS_FM(ty, q[3]), twz = NRS_FM(tz, q[3]);
  const float txx = NRS_FM(tx, q[0]), txy = NRS_FM(ty, q[0]), txz = NRS_FM(tz, q[0]);
  const float tyy = NRS_FM(ty, q[1]), tyz = NRS_FM(tz, q[1]), tzz = NRS_FM(tz, q[2]);
  R[0] = NRS_FS(1.f, NRS_FA(tyy, tzz)); R[1] = NRS_FS(txy, twz);              R[2] = NRS_FA(txz, twy);
  R[3] = NRS_FA(txy, twz);              R[4] = NRS_FS(1.f, NRS_FA(txx, tzz)); R[5] = NRS_FS(tyz, twx);
  R[6] = NRS_FS(txz, twy);              R[7] = NRS_FA(tyz, twx);              R[8] = NRS_FS(1.f, NRS_FA(txx, tyy));
}
TRI_DEV float sqn_f(const float* v) { return NRS_FA(NRS_FA(NRS_FM(v[0], v[0]), NRS_FM(v[1], v[1])), NRS_FM(v[2], v[2])); }
TRI_DEV float norm_f(const float* v) { return sqrtf(sqn_f(v)); }
TRI_DEV void normalized_f(const float* v, float* o) {
  const float n2 = sqn_f(v);
  if (n2 > 0.f) {
    const float n = sqrtf(n2);
    for (int i = 0; i < 3; i++) o[i] = NRS_FD(v[i], n);
  } else {
    for (int i = 0; i < 3; i++) o[i] = v[i];
  }
}
TRI_DEV void cross_f(const float* a, const float* b, float* o) {
  o[0] = NRS_FS(NRS_FM(a[1], b[2]), NRS_FM(a[2], b[1]));
  o[1] = NRS_FS(NRS_FM(a[2], b[0]), NRS_FM(a[0], b[2]));
  o[2] = NRS_FS(NRS_FM(a[0], b[1]), NRS_FM(a[1], b[0]));
}
// calibration/pin_hole.cc:33-38 ; calibration/kannala_brandt_8.cc:52-85
TRI_DEV void unproject_f(const Cam& c, float u, float v, float* ray) {
  if (c.model == 0) {
    ray[0] = NRS_FD(NRS_FS(u, c.p[2]), c.p[0]);
    ray[1] = NRS_FD(NRS_FS(v, c.p[3]), c.p[1]);
    ray[2] = 1.f;
    return;
  }
  const float pwx = NRS_FD(NRS_FS(u, c.p[2]), c.p[0]), pwy = NRS_FD(NRS_FS(v, c.p[3]), c.p[1]);
  const float theta_d = sqrtf(NRS_FA(NRS_FM(pwx, pwx), NRS_FM(pwy, pwy)));
  float th = 0.f;
  if (theta_d > 1e-8) {
    float theta = theta_d;
    for (int j = 0; j < 10; j++) {
      const float t2 = NRS_FM(theta, theta), t4 = NRS_FM(t2, t2), t6 = NRS_FM(t4, t2), t8 = NRS_FM(t4, t4);
      const float k0 = NRS_FM(c.p[4], t2), k1 = NRS_FM(c.p[5], t4), k2 = NRS_FM(c.p[6], t6), k3 = NRS_FM(c.p[7], t8);
      const float num = NRS_FS(NRS_FM(theta, NRS_FA(NRS_FA(NRS_FA(NRS_FA(1.f, k0), k1), k2), k3)), theta_d);
      const float den = NRS_FA(NRS_FA(NRS_FA(NRS_FA(1.f, NRS_FM(3.f, k0)), NRS_FM(5.f, k1)), NRS_FM(7.f, k2)), NRS_FM(9.f, k3));
      const float fix = NRS_FD(num, den);
      theta = NRS_FS(theta, fix);
      if (fabsf(fix) < 1e-6f) break;
    }
    th = theta;
  }
  ray[0] = NRS_FD(NRS_FM(sinf(th), pwx), theta_d);
  ray[1] = NRS_FD(NRS_FM(sinf(th), pwy), theta_d);
  ray[2] = cosf(th);
}
// geometry_toolbox.cc:46-78
TRI_DEV void triangulate_mid_point(const float* ray_1, const float* ray_2, const SE3f& cam1, const SE3f& cam2, float* X) {
  float f0h[3], f1h[3];
  normalized_f(ray_1, f0h);
  normalized_f(ray_2, f1h);
  const SE3f T10 = mul_f(cam2, inverse_f(cam1));
  float R[9], Rf0[3];
  quat_to_R_f(T10.q, R);
  for (int r = 0; r < 3; r++)
    Rf0[r] = NRS_FA(NRS_FA(NRS_FM(R[r * 3], f0h[0]), NRS_FM(R[r * 3 + 1], f0h[1])), NRS_FM(R[r * 3 + 2], f0h[2]));
  float p[3], q[3], rr[3];
  cross_f(Rf0, f1h, p);
  cross_f(Rf0, T10.t, q);
  cross_f(f1h, T10.t, rr);
  const float pn = norm_f(p), qn = norm_f(q), rn = norm_f(rr);
  const float s1 = NRS_FD(qn, NRS_FA(qn, rn)), s2 = NRS_FD(rn, pn);
  float x1[3];
  for (int i = 0; i < 3; i++) x1[i] = NRS_FM(s1, NRS_FA(T10.t[i], NRS_FM(s2, NRS_FA(Rf0[i], f1h[i]))));
  map_f(inverse_f(cam2), x1, X);
}
TRI_DEV float sq_reproj(const float* a, float bu, float bv) {
  const float ex = NRS_FS(a[0], bu), ey = NRS_FS(a[1], bv);
  return NRS_FA(NRS_FM(ex, ex), NRS_FM(ey, ey));
}
TRI_DEV SE3f load_pose(const float* p) {
  SE3f T;
  for (int i = 0; i < 4; i++) T.q[i] = p[i];
  for (int i = 0; i < 3; i++) T.t[i] = p[4 + i];
  return T;
}

// ---- shared-memory work space of one candidate -------------------------------------------------------------------
struct Work {
  double* W;      // packed lower triangle of H + lambda I, n (n + 1) / 2
  double* x;      // 3T estimates (camera coordinates of each frame)
  double* xbak;   // 3T
  double* dx;     // 3T last solution (stale after a failed factorisation, like g2o's _x)
  double* b;      // 3T
  double* r;      // 3T right-hand side scratch of the triangular solves
  double* ld;     // 3T diagonal of L (kept apart so a column of the factorisation needs two barriers, not three)
  double* wpos;   // 3T world positions T_wc x
  double* err_r;  // 2T reprojection errors
  double* Jr;     // 6T numeric reprojection Jacobians (2x3 row-major)
  double* Hr;     // 6T 4 Jr^T Jr (xx xy xz yy yz zz)
  double* Twc;    // 7T world_T_camera as g2o::SE3Quat (q xyzw, t)
  double* part;   // 5 * kThreads partial sums
  double* scal;   // 8 scalars: 0 lambda, 1 ni, 2 currentChi, 3 tempChi, 4 rho
  float* nbp;     // 3 kNB T neighbour world positions
  int* ictl;      // 8 control words: 0 status, 1 chol ok, 2 accept, 3 continue, 4 terminate, 5 qmax
  int* deg;       // T
  int* sfail;     // T seed failure codes
  int* perm;      // 3T: coupled solves — permuted position -> dof (decoupled vertices component-major first, coupled last)
  unsigned char* cnt;  // T*T pair counts
  unsigned char* nbv;  // kNB T validity
  unsigned short* vmask;  // T: bit j = neighbour j valid in this frame AND in the first one (edge test = one AND)
};

NRS_HD size_t work_bytes(int T) {
  const size_t n = 3 * (size_t)T;
  size_t d = n * (n + 1) / 2 + 7 * n + 2 * T + 6 * T + 6 * T + 7 * T + 5 * kThreads + 8;
  size_t bytes = d * sizeof(double);
  bytes += (size_t)3 * kNB * T * sizeof(float);
  bytes += (size_t)(8 + 5 * T) * sizeof(int);
  bytes += (size_t)T * T + (size_t)kNB * T;
  bytes = (bytes + 1) & ~(size_t)1;
  bytes += 2 * (size_t)T;
  return (bytes + 15) & ~(size_t)15;
}
TRI_DEV Work carve(void* smem, int T) {
  const size_t n = 3 * (size_t)T;
  Work w;
  double* d = reinterpret_cast<double*>(smem);
  w.W = d; d += n * (n + 1) / 2;
  w.x = d; d += n;
  w.xbak = d; d += n;
  w.dx = d; d += n;
  w.b = d; d += n;
  w.r = d; d += n;
  w.ld = d; d += n;
  w.wpos = d; d += n;
  w.err_r = d; d += 2 * T;
  w.Jr = d; d += 6 * T;
  w.Hr = d; d += 6 * T;
  w.Twc = d; d += 7 * T;
  w.part = d; d += 5 * kThreads;
  w.scal = d; d += 8;
  float* f = reinterpret_cast<float*>(d);
  w.nbp = f; f += 3 * kNB * T;
  int* ip = reinterpret_cast<int*>(f);
  w.ictl = ip; ip += 8;
  w.deg = ip; ip += T;
  w.sfail = ip; ip += T;
  w.perm = ip; ip += 3 * T;
  unsigned char* c = reinterpret_cast<unsigned char*>(ip);
  w.cnt = c; c += (size_t)T * T;
  w.nbv = c; c += (size_t)kNB * T;
  c = reinterpret_cast<unsigned char*>((reinterpret_cast<size_t>(c) + 1) & ~(size_t)1);
  w.vmask = reinterpret_cast<unsigned short*>(c);
  return w;
}

TRI_DEV double& Wel(const Work& w, int i, int j) { return w.W[(size_t)i * (i + 1) / 2 + j]; }  // i >= j

// Residuals at the current estimate. Fills err_r and wpos, leaves in part[] per-thread partial sums; returns (to every
// thread) chi2 = sum 4 |e_r|^2 + sum omega |e_s|^2. with_b: also accumulates the spatial part of b (b_a -= omega e for
// the edge's first vertex, b_b += omega e for its second: Jacobians +I / -I). n_bad (optional, thread 0's view valid
// for all): number of spatial edges with chi2 > th_bad.
TRI_DEV double evaluate(const Cam& cam, const Work& w, int T, int n_nb, const float* uv, double omega, bool with_b,
                        double th_bad, int* n_bad_out) {
  const int tid = TRI_TID, nt = TRI_NT;
  for (int k = tid; k < T; k += nt) {
    const double* xk = w.x + 3 * k;
    float pu, pv;
    project_f(cam, (float)xk[0], (float)xk[1], (float)xk[2], pu, pv);
    w.err_r[2 * k] = (double)uv[2 * k] - (double)pu;
    w.err_r[2 * k + 1] = (double)uv[2 * k + 1] - (double)pv;
    pose_map(w.Twc + 7 * k, xk, w.wpos + 3 * k);
  }
  TRI_SYNC();
  // spatial edges: vertex v, slice s of the other vertices
  int S = nt / T;
  if (S < 1) S = 1;
  for (int slot = tid; slot < T * S; slot += nt) {
    const int v = slot / S, s = slot - v * S;
    double bx = 0, by = 0, bz = 0, chi = 0;
    int nbad = 0;
    for (int u = s; u < T; u += S) {
      if (u == v || w.cnt[v * T + u] == 0) continue;
      const int a = u < v ? u : v, bb = u < v ? v : u;  // edge (a, bb), a < bb
      const double* wa = w.wpos + 3 * a;
      const double* wb = w.wpos + 3 * bb;
      const double D0 = wb[0] - wa[0], D1 = wb[1] - wa[1], D2 = wb[2] - wa[2];
      for (unsigned m = (unsigned)w.vmask[a] & (unsigned)w.vmask[bb]; m; m &= m - 1) {
        const int j = TRI_FFS(m) - 1;  // ascending neighbour index, like the byte tests it replaces
        const float* pa = w.nbp + (a * kNB + j) * 3;
        const float* pb = w.nbp + (bb * kNB + j) * 3;
        const double e0 = (double)NRS_FS(pb[0], pa[0]) - D0;
        const double e1 = (double)NRS_FS(pb[1], pa[1]) - D1;
        const double e2 = (double)NRS_FS(pb[2], pa[2]) - D2;
        if (v == a) {  // first vertex of the edge: owns its chi2
          const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * omega;
          chi += c2;
          if (c2 > th_bad) nbad++;
          bx -= omega * e0;
          by -= omega * e1;
          bz -= omega * e2;
        } else {
          bx += omega * e0;
          by += omega * e1;
          bz += omega * e2;
        }
      }
    }
    double* p = w.part + 5 * (size_t)slot;
    p[0] = bx;
    p[1] = by;
    p[2] = bz;
    p[3] = chi;
    p[4] = (double)nbad;
  }
  TRI_SYNC();
  if (with_b) {
    for (int v = tid; v < T; v += nt) {
      double bx = 0, by = 0, bz = 0;
      for (int s = 0; s < S && v * S + s < (T * S); s++) {
        const double* p = w.part + 5 * (size_t)(v * S + s);
        bx += p[0];
        by += p[1];
        bz += p[2];
      }
      w.b[3 * v] = bx;
      w.b[3 * v + 1] = by;
      w.b[3 * v + 2] = bz;
    }
  }
  // every thread sums the same values in the same order -> identical chi2 everywhere, no broadcast needed
  double chi = 0;
  for (int k = 0; k < T; k++) chi += (w.err_r[2 * k] * w.err_r[2 * k] + w.err_r[2 * k + 1] * w.err_r[2 * k + 1]) * 4.0;
  int nbad = 0;
  const int slots = T * S;
  for (int s = 0; s < slots; s++) {
    chi += w.part[5 * (size_t)s + 3];
    nbad += (int)w.part[5 * (size_t)s + 4];
  }
  if (n_bad_out) *n_bad_out = nbad;
  TRI_SYNC();
  return chi;
}

// In-place LL^T of the packed lower triangle; false when a pivot is <= 0 (Eigen::SimplicialLLT's failure test).
TRI_DEV bool cholesky(const Work& w, int n) {
  const int tid = TRI_TID, nt = TRI_NT;
  const int KW = nt >= 16 ? 16 : nt, RW = nt / KW;
  const int tk = tid % KW, tr = tid / KW;
  for (int j = 0; j < n; j++) {
    const double d = Wel(w, j, j);  // final since the barrier that closed the previous trailing update; never overwritten
    if (d <= 0) return false;  // uniform: every thread reads the same value
    const double l = sqrt(d);
    for (int i = j + 1 + tid; i < n; i += nt) Wel(w, i, j) /= l;
    if (tid == 0) w.ld[j] = l;
    TRI_SYNC();
    if (tr < RW)
      for (int i = j + 1 + tr; i < n; i += RW) {
        const double lij = Wel(w, i, j);
        if (lij == 0.0) continue;
        for (int k = j + 1 + tk; k <= i; k += KW) Wel(w, i, k) -= lij * Wel(w, k, j);
      }
    TRI_SYNC();
  }
  return true;
}

// dx = P^T (L L^T)^-1 P b for the factor of the PERMUTED matrix (perm[position] = dof). Forward substitution reads b-updates in r and writes y into dx; backward substitution updates dx
// in place and writes the solution into r (a value is never overwritten in the step that reads it: one barrier per
// column); the solution is copied to dx at the end.
TRI_DEV void solve(const Work& w, int n, const int* perm) {
  const int tid = TRI_TID, nt = TRI_NT;
  for (int i = tid; i < n; i += nt) w.r[i] = w.b[perm[i]];
  TRI_SYNC();
  for (int j = 0; j < n; j++) {
    const double yj = w.r[j] / w.ld[j];
    for (int i = j + 1 + tid; i < n; i += nt) w.r[i] -= Wel(w, i, j) * yj;
    if (tid == 0) w.dx[j] = yj;
    TRI_SYNC();
  }
  for (int j = n - 1; j >= 0; j--) {
    const double xj = w.dx[j] / w.ld[j];
    for (int i = tid; i < j; i += nt) w.dx[i] -= Wel(w, j, i) * xj;
    if (tid == 0) w.r[j] = xj;
    TRI_SYNC();
  }
  for (int i = tid; i < n; i += nt) w.dx[perm[i]] = w.r[i];
  TRI_SYNC();
}

// The same for the decoupled system (every numeric reprojection Jacobian is zero, the usual case): H + lambda I =
// (omega L + lambda I) (x) I3, so ONE T x T factor serves the x, y and z right-hand sides. The interleaved 3T x 3T
// factorisation would compute exactly these numbers (its extra terms are products with exact zeros), with three times
// the dependent steps.
TRI_DEV void solve3(const Work& w, int T) {
  const int tid = TRI_TID, nt = TRI_NT;
  const int n = 3 * T;
  for (int i = tid; i < n; i += nt) w.r[i] = w.b[i];
  TRI_SYNC();
  for (int j = 0; j < T; j++) {
    const double l = w.ld[j];
    const double y0 = w.r[3 * j] / l, y1 = w.r[3 * j + 1] / l, y2 = w.r[3 * j + 2] / l;
    for (int q = 3 * (j + 1) + tid; q < n; q += nt) {
      const int i = q / 3, c = q - 3 * i;
      w.r[q] -= Wel(w, i, j) * (c == 0 ? y0 : (c == 1 ? y1 : y2));
    }
    if (tid == 0) {
      w.dx[3 * j] = y0;
      w.dx[3 * j + 1] = y1;
      w.dx[3 * j + 2] = y2;
    }
    TRI_SYNC();
  }
  for (int j = T - 1; j >= 0; j--) {
    const double l = w.ld[j];
    const double x0 = w.dx[3 * j] / l, x1 = w.dx[3 * j + 1] / l, x2 = w.dx[3 * j + 2] / l;
    for (int q = tid; q < 3 * j; q += nt) {
      const int i = q / 3, c = q - 3 * i;
      w.dx[q] -= Wel(w, j, i) * (c == 0 ? x0 : (c == 1 ? x1 : x2));
    }
    if (tid == 0) {
      w.r[3 * j] = x0;
      w.r[3 * j + 1] = x1;
      w.r[3 * j + 2] = x2;
    }
    TRI_SYNC();
  }
  for (int i = tid; i < n; i += nt) w.dx[i] = w.r[i];
  TRI_SYNC();
}

// One candidate. uv [2T], pose [7T], nb_pos [3 kNB T], nb_valid [kNB T] are the candidate's slices (global memory).
TRI_DEV void solve_candidate(const Cam& cam, int T, const float* uv, const float* pose, int n_nb, const float* nb_pos,
                             const unsigned char* nb_valid, void* smem, float* out, int* status_out, int* iters_out,
                             const RigidArgs rg) {
  const int tid = TRI_TID, nt = TRI_NT;
  const int n = 3 * T;
  const Work w = carve(smem, T);
  const double omega = (double)NRS_FD(1.0f, NRS_FM(0.1f, 0.1f));  // info_spatial, float arithmetic (:696-697)

  // ---- stage the neighbour data, two-view pre-checks (:569-635) -----------------------------------------------
  for (int i = tid; i < 3 * kNB * T; i += nt) w.nbp[i] = nb_pos[i];
  for (int i = tid; i < kNB * T; i += nt) w.nbv[i] = nb_valid[i];
  if (tid == 0) {
    int st = ST_OK, rst = ST_OK;
    float rX[3] = {0.f, 0.f, 0.f};
    if (n_nb <= 0) {
      st = ST_TOO_CLOSE;   // g2o_optimization.cc:569-571 ; "Close features" for both branches, mapping.cc:90-94
      rst = ST_TOO_CLOSE;
    } else {
      const float* cur_uv = uv;
      const float* prev_uv = uv + 2 * (T - 1);
      float cu[3], pu[3], cur_ray[3], prev_ray[3];
      unproject_f(cam, cur_uv[0], cur_uv[1], cu);
      unproject_f(cam, prev_uv[0], prev_uv[1], pu);
      normalized_f(cu, cur_ray);
      normalized_f(pu, prev_ray);
      const SE3f cur_T = load_pose(pose), prev_T = load_pose(pose + 7 * (T - 1));
      float X[3], pc_cur[3], pc_prev[3], cu_u, cu_v, pr_u, pr_v;
      triangulate_mid_point(prev_ray, cur_ray, prev_T, cur_T, X);
      map_f(cur_T, X, pc_cur);
      project_f(cam, pc_cur[0], pc_cur[1], pc_cur[2], cu_u, cu_v);
      map_f(prev_T, X, pc_prev);
      project_f(cam, pc_prev[0], pc_prev[1], pc_prev[2], pr_u, pr_v);
      const double e_cur = (double)sq_reproj(cur_uv, cu_u, cu_v), e_prev = (double)sq_reproj(prev_uv, pr_u, pr_v);
      const SE3f ci = inverse_f(cur_T), pi = inverse_f(prev_T);
      const float n1[3] = {NRS_FS(X[0], ci.t[0]), NRS_FS(X[1], ci.t[1]), NRS_FS(X[2], ci.t[2])};
      const float n2[3] = {NRS_FS(X[0], pi.t[0]), NRS_FS(X[1], pi.t[1]), NRS_FS(X[2], pi.t[2])};
      const float dot = NRS_FA(NRS_FA(NRS_FM(n1[0], n2[0]), NRS_FM(n1[1], n2[1])), NRS_FM(n1[2], n2[2]));
      const float c = NRS_FD(dot, NRS_FM(norm_f(n1), norm_f(n2)));
      const float parallax = acosf(fminf(c, 1.f));
      // DeformableTriangulation's order (:618-635): first camera, second camera, parallax
      if (e_cur > 5.991) st = ST_HIGH_REPROJ_FIRST;
      else if (e_prev > 5.991) st = ST_HIGH_REPROJ_SECOND;
      else if ((double)parallax < 0.0025 * 5.0) st = ST_LOW_PARALLAX;
      if (rg.enabled) {
        if (T < rg.min_track) st = ST_SHORT_TRACK;   // mapping.cc:94,111-113: DeformableTriangulation is not called
        // rigid branch (mapping.cc:122-183): rigidity, parallax window, depth and reprojection in BOTH views;
        // NaNs fall through every comparison exactly like in the reference
        const float lo = NRS_FM(rg.rad_per_pixel, 10.f), hi = NRS_FM(rg.rad_per_pixel, 20.f);
        if (!rg.rigid_ok) rst = ST_NOT_RIGID;
        else if (parallax < lo || parallax > hi) rst = ST_RIGID_PARALLAX;
        else if (pc_prev[2] < 0) rst = ST_RIGID_PARALLAX;
        else if (e_prev > 5.991) rst = ST_RIGID_PARALLAX;
        else if (pc_cur[2] < 0) rst = ST_RIGID_PARALLAX;
        else if (e_cur > 5.991) rst = ST_RIGID_PARALLAX;
        if (rst == ST_OK) {
          rX[0] = X[0];
          rX[1] = X[1];
          rX[2] = X[2];
        }
      }
    }
    if (rg.enabled) {
      rg.out[0] = rX[0];
      rg.out[1] = rX[1];
      rg.out[2] = rX[2];
      *rg.status = rst;
    }
    w.ictl[0] = st;
  }
  TRI_SYNC();
  if (w.ictl[0] != ST_OK) {
    if (tid == 0) {
      *status_out = w.ictl[0];
      *iters_out = 0;
      out[0] = out[1] = out[2] = 0.f;
    }
    return;
  }

  // ---- seeds, world_T_camera, pair counts (:637-690, :696-748) -------------------------------------------------
  for (int k = tid; k < T; k += nt) {
    const SE3f cT = load_pose(pose + 7 * k);
    float depth = 0.f;
    int m = 0;
    for (int j = 0; j < n_nb; j++) {
      if (!w.nbv[k * kNB + j]) continue;
      float pc[3];
      map_f(cT, w.nbp + (k * kNB + j) * 3, pc);
      depth = NRS_FA(depth, pc[2]);
      m++;
    }
    int f = 0;
    if (m == 0) {
      f = ST_NO_NEIGHBOURS;
    } else {
      depth = NRS_FD(depth, (float)m);
      if (depth < 0) f = ST_NEGATIVE_DEPTH;
    }
    w.sfail[k] = f;
    float ray[3];
    unproject_f(cam, uv[2 * k], uv[2 * k + 1], ray);
    for (int i = 0; i < 3; i++) {
      w.x[3 * k + i] = (double)NRS_FM(ray[i], depth);
      w.dx[3 * k + i] = 0.0;
    }
    const SE3f inv = inverse_f(cT);
    double* Tw = w.Twc + 7 * k;
    for (int i = 0; i < 4; i++) Tw[i] = inv.q[i];
    for (int i = 0; i < 3; i++) Tw[4 + i] = inv.t[i];
    pose_normalize(Tw);  // g2o::SE3Quat ctor
  }
  for (int k = tid; k < T; k += nt) {
    unsigned m = 0;
    for (int j = 0; j < n_nb; j++) m |= (w.nbv[k * kNB + j] && w.nbv[j]) ? (1u << j) : 0u;
    w.vmask[k] = (unsigned short)m;
  }
  for (int pr = tid; pr < T * T; pr += nt) {
    const int a = pr / T, bq = pr - a * T;
    int c = 0;
    if (a != bq)
      for (int j = 0; j < n_nb; j++) c += (w.nbv[a * kNB + j] && w.nbv[bq * kNB + j] && w.nbv[j]) ? 1 : 0;
    w.cnt[pr] = (unsigned char)c;
  }
  TRI_SYNC();
  if (tid == 0) {
    int st = ST_OK;
    for (int k = 0; k < T && st == ST_OK; k++) st = w.sfail[k];
    w.ictl[0] = st;
  }
  for (int a = tid; a < T; a += nt) {
    int d = 0;
    for (int u = 0; u < T; u++) d += w.cnt[a * T + u];
    w.deg[a] = d;
  }
  TRI_SYNC();
  if (w.ictl[0] != ST_OK) {
    if (tid == 0) {
      *status_out = w.ictl[0];
      *iters_out = 0;
      out[0] = out[1] = out[2] = 0.f;
    }
    return;
  }
  int n_reg = 0;
  for (int a = 0; a < T; a++) n_reg += w.deg[a];
  n_reg /= 2;

  // ---- optimizer.optimize(10): g2o Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:57-165) -------------
  int iters = 0;
  double lambda = 0, ni = 2;
  for (int it = 0; it < 10; it++) {
    double currentChi = evaluate(cam, w, T, n_nb, uv, omega, true, 1e300, nullptr);
    // numeric Jacobian of the reprojection edge (base_fixed_sized_edge.hpp:160-199), one thread per (vertex, column)
    for (int kd = tid; kd < n; kd += nt) {
      const int k = kd / 3, d = kd - 3 * k;
      const double delta = 1e-9, scalar = 1 / (2 * delta);
      double xp[3] = {w.x[3 * k], w.x[3 * k + 1], w.x[3 * k + 2]};
      double xm[3] = {xp[0], xp[1], xp[2]};
      xp[d] = xp[d] + delta;
      xm[d] = xm[d] + (-delta);
      float upu, upv, umu, umv;
      project_f(cam, (float)xp[0], (float)xp[1], (float)xp[2], upu, upv);
      project_f(cam, (float)xm[0], (float)xm[1], (float)xm[2], umu, umv);
      const double mu = (double)uv[2 * k], mv = (double)uv[2 * k + 1];
      w.Jr[6 * k + d] = scalar * ((mu - (double)upu) - (mu - (double)umu));
      w.Jr[6 * k + 3 + d] = scalar * ((mv - (double)upv) - (mv - (double)umv));
    }
    if (tid == 0) w.ictl[1] = 0;
    TRI_SYNC();
    for (int k = tid; k < T; k += nt) {
      const double* J = w.Jr + 6 * k;
      double* H = w.Hr + 6 * k;
      int q = 0;
      bool nz = false;
      for (int a = 0; a < 3; a++)
        for (int c = a; c < 3; c++) {
          const double h = (J[a] * 4.0) * J[c] + (J[3 + a] * 4.0) * J[3 + c];
          H[q++] = h;
          nz = nz || (h != 0.0);
        }
      if (nz) w.ictl[1] = 1;  // some reprojection block couples x, y, z: full 3T x 3T system this iteration
      const double we0 = -4.0 * w.err_r[2 * k], we1 = -4.0 * w.err_r[2 * k + 1];
      for (int c = 0; c < 3; c++) w.b[3 * k + c] += J[c] * we0 + J[3 + c] * we1;
    }
    TRI_SYNC();
    // Coupled linearisation: order the decoupled vertices first, component-major (three T' x T' diagonal blocks with
    // nothing between them), and the vertices whose reprojection block is non-zero last. The dense factorisation of
    // the permuted matrix then skips the structurally zero rows (L(i,j) == 0 test): its work is three T'^3 / 6
    // blocks plus a thin dense border instead of (3T)^3 / 6.
    if (w.ictl[1] != 0 && tid == 0) {
      int Tp = 0;
      for (int k = 0; k < T; k++) {
        const double* H = w.Hr + 6 * k;
        const bool nzk = H[0] != 0.0 || H[1] != 0.0 || H[2] != 0.0 || H[3] != 0.0 || H[4] != 0.0 || H[5] != 0.0;
        Tp += nzk ? 0 : 1;
      }
      int r = 0, sdx = 0;
      for (int k = 0; k < T; k++) {
        const double* H = w.Hr + 6 * k;
        const bool nzk = H[0] != 0.0 || H[1] != 0.0 || H[2] != 0.0 || H[3] != 0.0 || H[4] != 0.0 || H[5] != 0.0;
        if (!nzk) {
          for (int c = 0; c < 3; c++) w.perm[c * Tp + r] = 3 * k + c;
          r++;
        } else {
          for (int c = 0; c < 3; c++) w.perm[3 * Tp + 3 * sdx + c] = 3 * k + c;
          sdx++;
        }
      }
    }
    TRI_SYNC();
    if (it == 0) {  // computeLambdaInit :153-165 — every thread computes the same value
      double md = 0;
      for (int k = 0; k < T; k++) {
        const double dg = omega * (double)w.deg[k];
        md = fmax(md, fabs(w.Hr[6 * k] + dg));
        md = fmax(md, fabs(w.Hr[6 * k + 3] + dg));
        md = fmax(md, fabs(w.Hr[6 * k + 5] + dg));
      }
      lambda = 1e-5 * md;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    bool again;
    do {
      for (int i = tid; i < n; i += nt) w.xbak[i] = w.x[i];
      const bool coupled = w.ictl[1] != 0;
      // H + lambda I, packed lower
      if (!coupled) {
        for (int i = tid; i < T; i += nt) {
          double* row = w.W + (size_t)i * (i + 1) / 2;
          for (int j = 0; j < i; j++) row[j] = -omega * (double)w.cnt[i * T + j];
          row[i] = TRI_DADD(TRI_DMUL(omega, (double)w.deg[i]), lambda);
        }
      } else
      for (int pi = tid; pi < n; pi += nt) {
        const int i = w.perm[pi];
        const int a = i / 3, ci = i - 3 * a;
        double* row = w.W + (size_t)pi * (pi + 1) / 2;
        for (int pj = 0; pj <= pi; pj++) {
          const int j = w.perm[pj];
          const int bq = j / 3, cj = j - 3 * bq;
          double v = 0;
          if (a == bq) {
            const int lo = cj < ci ? cj : ci, hi = cj < ci ? ci : cj;
            const int q = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
            v = w.Hr[6 * a + q];
            if (ci == cj) v += TRI_DADD(TRI_DMUL(omega, (double)w.deg[a]), lambda);  // same roundings as the decoupled path
          } else if (ci == cj) {
            v = -omega * (double)w.cnt[a * T + bq];
          }
          row[pj] = v;
        }
      }
      TRI_SYNC();
      const bool ok2 = cholesky(w, coupled ? n : T);
      TRI_SYNC();
      if (ok2) {
        if (coupled) solve(w, n, w.perm); else solve3(w, T);
      }
      for (int i = tid; i < n; i += nt) w.x[i] += w.dx[i];
      TRI_SYNC();
      double tempChi = evaluate(cam, w, T, n_nb, uv, omega, false, 1e300, nullptr);
      if (!ok2) tempChi = 1.7976931348623157e308;
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < n; j++) scale += w.dx[j] * (lambda * w.dx[j] + w.b[j]);
      scale += 1e-3;
      rho /= scale;
      bool brk = false;
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
        TRI_SYNC();
        for (int i = tid; i < n; i += nt) w.x[i] = w.xbak[i];
        if (!isfinite(lambda)) brk = true;
      }
      TRI_SYNC();
      if (!brk) qmax++;
      again = !brk && rho < 0 && qmax < 10;
    } while (again);
    iters++;
    if (qmax == 10 || rho == 0 || !isfinite(lambda)) break;
  }

  // ---- acceptance tests and the result (:765-813) --------------------------------------------------------------
  int bad_edges = 0;
  evaluate(cam, w, T, n_nb, uv, omega, false, (double)7.815f, &bad_edges);
  if (tid == 0) {
    int st = ST_OK;
    if ((double)((float)bad_edges / (float)n_reg) > 0.5) st = ST_BAD_NEIGHBOURS;  // 0/0 = NaN passes, like the reference
    if (st == ST_OK) {
      int n_bad = 0;
      for (int k = 0; k < T; k++) {
        const double c2 = (w.err_r[2 * k] * w.err_r[2 * k] + w.err_r[2 * k + 1] * w.err_r[2 * k + 1]) * 4.0;
        if (c2 > 5.99 * 10) n_bad++;
      }
      if ((double)((float)n_bad / (float)T) > 0.5) st = ST_HIGH_ERROR;
    }
    float o[3] = {0.f, 0.f, 0.f};
    if (st == ST_OK) {
      const float depth = (float)w.x[3 * (T - 1) + 2];
      float ray[3];
      unproject_f(cam, uv[2 * (T - 1)], uv[2 * (T - 1) + 1], ray);
      const float z = ray[2];
      for (int i = 0; i < 3; i++) ray[i] = NRS_FD(ray[i], z);
      const float pl[3] = {NRS_FM(ray[0], depth), NRS_FM(ray[1], depth), NRS_FM(ray[2], depth)};
      map_f(inverse_f(load_pose(pose + 7 * (T - 1))), pl, o);
      if (isnan(o[0]) || isnan(o[1]) || isnan(o[2])) st = ST_NAN;
    }
    out[0] = o[0];
    out[1] = o[1];
    out[2] = o[2];
    *status_out = st;
    *iters_out = iters;
  }
}

}  // namespace tri
}  // namespace nrs
