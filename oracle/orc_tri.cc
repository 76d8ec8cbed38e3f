// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). Parity: unpinned by the reference (NR-SLAM ships no
// tests or golden outputs for this function).
//
// CPU restatement of DeformableTriangulation (modules/optimization/g2o_optimization.cc:559-814) over a flattened
// TemporalBuffer view, one candidate at a time exactly like Mapping::LandmarkTriangulation calls it
// (modules/mapping/mapping.cc:88-113), plus RegularizationGraph::UpdateVertex over a list of vertices
// (modules/map/regularization_graph.cc:89-146 driven by g2o_optimization.cc:458-474).
// fp32 geometry follows the vendored Sophus (third_party/Sophus/sophus/so3.hpp:318-325,346-397, se3.hpp:222-225,
// 302-306) and modules/utilities/geometry_toolbox.cc:31-78; the LM runs through orc::Optimizer (orc_lm.cc) with the
// exact sparse Cholesky, edge at a time, in the reference's insertion order.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/nrslam_b200.h"
#include "orc_lm.h"

namespace {
using namespace orc;

struct SE3f {
  float q[4];  // x y z w
  float t[3];
};

// so3.hpp:388-397  uv = q.vec x p ; uv += uv ; p + w uv + q.vec x uv
void rot_f(const float q[4], const float p[3], float o[3]) {
  float uv[3] = {q[1] * p[2] - q[2] * p[1], q[2] * p[0] - q[0] * p[2], q[0] * p[1] - q[1] * p[0]};
  uv[0] += uv[0];
  uv[1] += uv[1];
  uv[2] += uv[2];
  const float c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; i++) o[i] = (p[i] + q[3] * uv[i]) + c[i];
}
// so3.hpp:318-325
void normalize_q(float q[4]) {
  const float len = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= len;
}
// se3.hpp:222-225: invR = SO3(conj) (normalised by the ctor, so3.hpp:528-534), t' = invR * (t * -1)
SE3f inverse_f(const SE3f& T) {
  SE3f r;
  r.q[0] = -T.q[0];
  r.q[1] = -T.q[1];
  r.q[2] = -T.q[2];
  r.q[3] = T.q[3];
  normalize_q(r.q);
  const float mt[3] = {T.t[0] * -1.f, T.t[1] * -1.f, T.t[2] * -1.f};
  rot_f(r.q, mt, r.t);
  return r;
}
// se3.hpp:302-306 with so3.hpp:346-369 (plain quaternion product, normalised by the SO3 ctor)
SE3f mul_f(const SE3f& a, const SE3f& b) {
  SE3f r;
  const float *A = a.q, *B = b.q;
  r.q[3] = A[3] * B[3] - A[0] * B[0] - A[1] * B[1] - A[2] * B[2];
  r.q[0] = A[3] * B[0] + A[0] * B[3] + A[1] * B[2] - A[2] * B[1];
  r.q[1] = A[3] * B[1] + A[1] * B[3] + A[2] * B[0] - A[0] * B[2];
  r.q[2] = A[3] * B[2] + A[2] * B[3] + A[0] * B[1] - A[1] * B[0];
  normalize_q(r.q);
  float rt[3];
  rot_f(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  return r;
}
void map_f(const SE3f& T, const float p[3], float o[3]) {  // se3.hpp: so3() * p + translation()
  rot_f(T.q, p, o);
  for (int i = 0; i < 3; i++) o[i] += T.t[i];
}
void quat_to_R_f(const float q[4], float R[9]) {  // Eigen::Quaternion::toRotationMatrix
  const float tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
float norm_f(const float v[3]) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
void normalized_f(const float v[3], float o[3]) {  // Eigen normalized(): v / sqrt(squaredNorm) when > 0
  const float n2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
  if (n2 > 0.f) {
    const float n = std::sqrt(n2);
    for (int i = 0; i < 3; i++) o[i] = v[i] / n;
  } else {
    for (int i = 0; i < 3; i++) o[i] = v[i];
  }
}
void cross_f(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// calibration/pin_hole.cc:33-38 ; calibration/kannala_brandt_8.cc:52-85
void unproject_f(const Camera& c, float u, float v, float ray[3]) {
  if (c.model == 0) {
    ray[0] = (u - c.p[2]) / c.p[0];
    ray[1] = (v - c.p[3]) / c.p[1];
    ray[2] = 1.f;
    return;
  }
  const float pwx = (u - c.p[2]) / c.p[0], pwy = (v - c.p[3]) / c.p[1];
  const float theta_d = sqrtf(pwx * pwx + pwy * pwy);
  float th = 0.f;  // the reference leaves it uninitialised when theta_d <= 1e-8 (UB); 0 here
  if (theta_d > 1e-8) {
    float theta = theta_d;
    for (int j = 0; j < 10; j++) {
      const float t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
      const float k0t2 = c.p[4] * t2, k1t4 = c.p[5] * t4, k2t6 = c.p[6] * t6, k3t8 = c.p[7] * t8;
      const float fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) /
                        (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8);
      theta = theta - fix;
      if (fabsf(fix) < 1e-6f) break;  // precision_ = 1e-6 (kannala_brandt_8.h)
    }
    th = theta;
  }
  ray[0] = sinf(th) * pwx / theta_d;
  ray[1] = sinf(th) * pwy / theta_d;
  ray[2] = cosf(th);
}

// geometry_toolbox.cc:46-78 (the adequacy test :66-71 computes values it never uses)
void triangulate_mid_point(const float ray_1[3], const float ray_2[3], const SE3f& cam1, const SE3f& cam2, float X[3]) {
  float f0h[3], f1h[3];
  normalized_f(ray_1, f0h);
  normalized_f(ray_2, f1h);
  const SE3f T10 = mul_f(cam2, inverse_f(cam1));
  float R[9], Rf0[3];
  quat_to_R_f(T10.q, R);
  for (int r = 0; r < 3; r++) Rf0[r] = (R[r * 3] * f0h[0] + R[r * 3 + 1] * f0h[1]) + R[r * 3 + 2] * f0h[2];
  float p[3], q[3], rr[3];
  cross_f(Rf0, f1h, p);
  cross_f(Rf0, T10.t, q);
  cross_f(f1h, T10.t, rr);
  const float pn = norm_f(p), qn = norm_f(q), rn = norm_f(rr);
  const float s1 = qn / (qn + rn), s2 = rn / pn;
  float x1[3];
  for (int i = 0; i < 3; i++) x1[i] = s1 * (T10.t[i] + s2 * (Rf0[i] + f1h[i]));
  map_f(inverse_f(cam2), x1, X);
}

float sq_reproj(const float a[2], const float b[2]) {  // geometry_toolbox.cc:31-36
  const float ex = a[0] - b[0], ey = a[1] - b[1];
  return ex * ex + ey * ey;
}
float rays_parallax(const float a[3], const float b[3]) {  // geometry_toolbox.cc:38-44
  const float c = ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) / (norm_f(a) * norm_f(b));
  return acosf(std::fmin(c, 1.f));
}

SE3f load_pose(const float* p) {
  SE3f T;
  for (int i = 0; i < 4; i++) T.q[i] = p[i];
  for (int i = 0; i < 3; i++) T.t[i] = p[4 + i];
  return T;
}

// g2o::SE3Quat(inverse().unit_quaternion().cast<double>(), inverse().translation().cast<double>())
// (g2o_optimization.cc:696-698; the SE3Quat ctor normalises, se3quat.h:63-69)
SE3 world_T_cam_g2o(const SE3f& cam_T_world) {
  const SE3f inv = inverse_f(cam_T_world);
  SE3 T;
  for (int i = 0; i < 4; i++) T.q[i] = inv.q[i];
  for (int i = 0; i < 3; i++) T.t[i] = inv.t[i];
  normalize_rotation(T);
  return T;
}

// One candidate. order != 0 visits the neighbours in reverse when the spatial edges are created (a different but
// equally valid summation order; used by the tests to measure how reproducible the function is under re-association).
int triangulate_one(const Camera& cam, int T, const float* uv, const float* pose, int n_nb, const float* nb_pos,
                    const uint8_t* nb_valid, int order, float out[3], int* lm_iters) {
  const int NB = NRSLAM_B200_TRI_MAX_NB;
  if (lm_iters) *lm_iters = 0;
  if (n_nb <= 0) return NRSLAM_B200_TRI_TOO_CLOSE;  // :569-571
  // "current" = track.front() (oldest), "previous" = track.back() (latest)  — :591-592
  const float* cur_uv = uv;
  const float* prev_uv = uv + 2 * (T - 1);
  float cur_ray_u[3], prev_ray_u[3], cur_ray[3], prev_ray[3];
  unproject_f(cam, cur_uv[0], cur_uv[1], cur_ray_u);
  unproject_f(cam, prev_uv[0], prev_uv[1], prev_ray_u);
  normalized_f(cur_ray_u, cur_ray);
  normalized_f(prev_ray_u, prev_ray);
  const SE3f cur_T = load_pose(pose), prev_T = load_pose(pose + 7 * (T - 1));
  float X[3];
  triangulate_mid_point(prev_ray, cur_ray, prev_T, cur_T, X);  // :604-606
  float pc[3], proj[2];
  map_f(cur_T, X, pc);
  project_f(cam, pc, proj);
  if (sq_reproj(cur_uv, proj) > 5.991) return NRSLAM_B200_TRI_HIGH_REPROJ_FIRST;
  map_f(prev_T, X, pc);
  project_f(cam, pc, proj);
  if (sq_reproj(prev_uv, proj) > 5.991) return NRSLAM_B200_TRI_HIGH_REPROJ_SECOND;
  {
    const SE3f ci = inverse_f(cur_T), pi = inverse_f(prev_T);
    const float n1[3] = {X[0] - ci.t[0], X[1] - ci.t[1], X[2] - ci.t[2]};
    const float n2[3] = {X[0] - pi.t[0], X[1] - pi.t[1], X[2] - pi.t[2]};
    const float parallax = rays_parallax(n1, n2);
    if (parallax < 0.0025 * 5.f) return NRSLAM_B200_TRI_LOW_PARALLAX;
  }

  Optimizer opt(cam, /*dense_solver=*/false);  // BlockSolverX + LinearSolverEigen (:578-583)
  const double info_reproj = 1.0f / (0.5f * 0.5f);      // :587-588
  std::vector<SE3f> cam_T(T);
  for (int k = 0; k < T; k++) {  // :637-690
    cam_T[k] = load_pose(pose + 7 * k);
    float depth_seed = 0.f;
    int n = 0;
    for (int j = 0; j < n_nb; j++) {
      if (!nb_valid[(size_t)k * NB + j]) continue;
      float pcn[3];
      map_f(cam_T[k], nb_pos + ((size_t)k * NB + j) * 3, pcn);
      depth_seed += pcn[2];
      n++;
    }
    if (n == 0) return NRSLAM_B200_TRI_NO_NEIGHBOURS;
    depth_seed /= (float)n;
    if (depth_seed < 0) return NRSLAM_B200_TRI_NEGATIVE_DEPTH;
    float ray[3];
    unproject_f(cam, uv[2 * k], uv[2 * k + 1], ray);
    Vertex v;
    v.type = V_POINT;
    v.dim = 3;
    for (int i = 0; i < 3; i++) v.x[i] = (double)(ray[i] * depth_seed);
    opt.add_vertex(v);
    Edge e;
    e.type = E_REPROJ_ONLY_DEFORMATION;
    e.nv = 1;
    e.v[0] = k;
    e.dim = 2;
    e.info = info_reproj;
    e.meas[0] = uv[2 * k];
    e.meas[1] = uv[2 * k + 1];
    opt.add_edge(e);
  }
  const double info_spatial = 1.0f / (0.1f * 0.1f);  // :696-697 (float arithmetic, then widened)
  std::vector<int> reg_edges;
  std::vector<SE3> world_T_cam(T);
  for (int k = 0; k < T; k++) world_T_cam[k] = world_T_cam_g2o(cam_T[k]);
  for (int a = 0; a < T; a++)
    for (int b = a + 1; b < T; b++)
      for (int jj = 0; jj < n_nb; jj++) {
        const int j = order ? n_nb - 1 - jj : jj;
        if (!nb_valid[(size_t)a * NB + j] || !nb_valid[(size_t)b * NB + j] || !nb_valid[j]) continue;  // :726-729
        const float* pa = nb_pos + ((size_t)a * NB + j) * 3;
        const float* pb = nb_pos + ((size_t)b * NB + j) * 3;
        Edge e;
        e.type = E_SPATIAL_OBS;
        e.nv = 2;
        e.v[0] = a;
        e.v[1] = b;
        e.dim = 3;
        e.info = info_spatial;
        e.weight = 1.0f;
        for (int i = 0; i < 3; i++) e.meas[i] = (double)(pb[i] - pa[i]);  // flow, fp32 (:731)
        e.Ta = world_T_cam[a];
        e.Tb = world_T_cam[b];
        reg_edges.push_back(opt.add_edge(e));
      }
  // optimizer.edges().size() == 0 cannot happen (T >= 1 reprojection edges)
  opt.initialize_optimization(0);
  opt.optimize(10);  // :763
  if (lm_iters) *lm_iters = opt.stats.iterations;

  int bad_edges = 0;
  for (int id : reg_edges) {
    Edge& e = opt.edges[id];
    opt.compute_error(e);
    if (opt.chi2(e) > 7.815f) bad_edges++;  // th_huber_3dof_squared is a float (:692)
  }
  if ((float)bad_edges / (float)reg_edges.size() > 0.5) return NRSLAM_B200_TRI_BAD_NEIGHBOURS;  // 0/0 = NaN passes
  int n_bad = 0;
  for (int k = 0; k < T; k++) {
    Edge& e = opt.edges[k];
    opt.compute_error(e);
    if (opt.chi2(e) > 5.99 * 10) n_bad++;
  }
  if ((float)n_bad / (float)T > 0.5) return NRSLAM_B200_TRI_HIGH_ERROR;

  const float current_depth = (float)opt.vertices[T - 1].x[2];  // :798-799
  float ray[3];
  unproject_f(cam, prev_uv[0], prev_uv[1], ray);
  const float z = ray[2];
  for (int i = 0; i < 3; i++) ray[i] /= z;
  const float pl[3] = {ray[0] * current_depth, ray[1] * current_depth, ray[2] * current_depth};
  map_f(inverse_f(prev_T), pl, out);
  if (std::isnan(out[0]) || std::isnan(out[1]) || std::isnan(out[2])) return NRSLAM_B200_TRI_NAN;  // mapping.cc:98
  return NRSLAM_B200_TRI_OK;
}
// Rigid branch of Mapping::LandmarkTriangulation for one candidate (modules/mapping/mapping.cc:115-185).
int rigid_one(const Camera& cam, int T, const float* uv, const float* pose, int n_nb, int rigid_ok, float rad_per_pixel,
              float out[3]) {
  out[0] = out[1] = out[2] = 0.f;
  if (n_nb <= 0) return NRSLAM_B200_TRI_TOO_CLOSE;                      // :90-94 "Close features"
  if (!rigid_ok) return NRSLAM_B200_TRI_NOT_RIGID;                       // :122-125
  const float* cur_uv = uv;                                              // track.front()  (:118-119)
  const float* prev_uv = uv + 2 * (T - 1);                               // track.back()
  float cu[3], pu[3], cur_ray[3], prev_ray[3];
  unproject_f(cam, cur_uv[0], cur_uv[1], cu);
  unproject_f(cam, prev_uv[0], prev_uv[1], pu);
  normalized_f(cu, cur_ray);
  normalized_f(pu, prev_ray);
  const SE3f cur_T = load_pose(pose), prev_T = load_pose(pose + 7 * (T - 1));
  float X[3];
  triangulate_mid_point(prev_ray, cur_ray, prev_T, cur_T, X);            // :137-139
  const SE3f ci = inverse_f(cur_T), pi = inverse_f(prev_T);
  const float n1[3] = {X[0] - ci.t[0], X[1] - ci.t[1], X[2] - ci.t[2]};
  const float n2[3] = {X[0] - pi.t[0], X[1] - pi.t[1], X[2] - pi.t[2]};
  const float parallax = rays_parallax(n1, n2);
  if (parallax < rad_per_pixel * 10.f || parallax > rad_per_pixel * 20.f) return NRSLAM_B200_TRI_RIGID_PARALLAX;
  float pc[3], proj[2];
  map_f(prev_T, X, pc);                                                  // :158-168
  if (pc[2] < 0) return NRSLAM_B200_TRI_RIGID_PARALLAX;
  project_f(cam, pc, proj);
  if (sq_reproj(prev_uv, proj) > 5.991) return NRSLAM_B200_TRI_RIGID_PARALLAX;
  map_f(cur_T, X, pc);                                                   // :170-181
  if (pc[2] < 0) return NRSLAM_B200_TRI_RIGID_PARALLAX;
  project_f(cam, pc, proj);
  if (sq_reproj(cur_uv, proj) > 5.991) return NRSLAM_B200_TRI_RIGID_PARALLAX;
  out[0] = X[0];
  out[1] = X[1];
  out[2] = X[2];
  return NRSLAM_B200_TRI_OK;
}
}  // namespace

extern "C" {

// Mapping::LandmarkTriangulation's per-candidate work and its vote (mapping/mapping.cc:65-205), sequential.
int orc_landmark_triangulation_frame(const nrslam_b200_camera* cam_, int32_t n_cand, const int32_t* track_ptr,
                                     const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                                     const float* nb_pos, const uint8_t* nb_valid, const uint8_t* rigid_ok,
                                     float rad_per_pixel, int32_t min_track, float* deform_pos, int32_t* deform_status,
                                     float* rigid_pos, int32_t* rigid_status, float* selected_pos, uint8_t* selected) {
  Camera cam;
  cam.model = cam_->model;
  for (int i = 0; i < 8; i++) cam.p[i] = cam_->params[i];
  const int NB = NRSLAM_B200_TRI_MAX_NB;
  int n_rigid = 0, n_def = 0;
  for (int c = 0; c < n_cand; c++) {
    const int e0 = track_ptr[c], T = track_ptr[c + 1] - e0;
    const float* uv = track_uv + 2 * (size_t)e0;
    const float* pose = track_pose + 7 * (size_t)e0;
    float d[3] = {0, 0, 0}, r[3] = {0, 0, 0};
    int ds;
    if (n_neighbours[c] <= 0) {
      ds = NRSLAM_B200_TRI_TOO_CLOSE;
    } else if (T >= min_track) {                                         // :94
      ds = triangulate_one(cam, T, uv, pose, n_neighbours[c], nb_pos + (size_t)e0 * NB * 3, nb_valid + (size_t)e0 * NB,
                           0, d, nullptr);
      if (ds != NRSLAM_B200_TRI_OK) d[0] = d[1] = d[2] = 0.f;
    } else {
      ds = NRSLAM_B200_TRI_SHORT_TRACK;                                  // :111-113
    }
    if (ds == NRSLAM_B200_TRI_OK) n_def++;                               // :100-103 (NaN results were mapped to an error)
    const int rs = rigid_one(cam, T, uv, pose, n_neighbours[c], rigid_ok[c], rad_per_pixel, r);
    if (rs == NRSLAM_B200_TRI_OK) n_rigid++;                             // :184-186
    for (int i = 0; i < 3; i++) {
      deform_pos[3 * (size_t)c + i] = d[i];
      rigid_pos[3 * (size_t)c + i] = r[i];
    }
    deform_status[c] = ds;
    rigid_status[c] = rs;
  }
  for (int c = 0; c < n_cand; c++) {                                     // :188-212
    const float* pick = nullptr;
    if (n_rigid > 1.5 * n_def) {
      if (rigid_status[c] == NRSLAM_B200_TRI_OK) pick = rigid_pos + 3 * (size_t)c;
    } else if (n_def >= 1.5 * n_rigid) {
      if (deform_status[c] == NRSLAM_B200_TRI_OK) pick = deform_pos + 3 * (size_t)c;
    }
    if (pick && (std::isnan(pick[0]) || std::isnan(pick[1]) || std::isnan(pick[2]))) pick = nullptr;  // :210-213
    selected[c] = pick ? 1 : 0;
    for (int i = 0; i < 3; i++) selected_pos[3 * (size_t)c + i] = pick ? pick[i] : 0.f;
  }
  return 0;
}

int orc_deformable_triangulation(const nrslam_b200_camera* cam_, int32_t n_cand, const int32_t* track_ptr,
                                 const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                                 const float* nb_pos, const uint8_t* nb_valid, int32_t order, float* position_out,
                                 int32_t* status_out, int32_t* lm_iterations_out) {
  Camera cam;
  cam.model = cam_->model;
  for (int i = 0; i < 8; i++) cam.p[i] = cam_->params[i];
  const int NB = NRSLAM_B200_TRI_MAX_NB;
  for (int c = 0; c < n_cand; c++) {
    const int e0 = track_ptr[c], T = track_ptr[c + 1] - e0;
    float out[3] = {0, 0, 0};
    int it = 0;
    status_out[c] = triangulate_one(cam, T, track_uv + 2 * (size_t)e0, track_pose + 7 * (size_t)e0, n_neighbours[c],
                                    nb_pos + (size_t)e0 * NB * 3, nb_valid + (size_t)e0 * NB, order, out, &it);
    for (int i = 0; i < 3; i++) position_out[3 * (size_t)c + i] = out[i];
    if (lm_iterations_out) lm_iterations_out[c] = it;
  }
  return 0;
}

// The loop of g2o_optimization.cc:458-474 over RegularizationGraph::UpdateVertex (regularization_graph.cc:130-146),
// sequential, in the caller's order.
int orc_graph_update_vertices(nrslam_b200_graph* g, int32_t n, const int32_t* vertices, const float* pos,
                              int32_t* good_out) {
  for (int i = 0; i < n; i++) {
    const int v = vertices[i];
    int n_good = 0;
    const float* p1 = pos + 3 * (size_t)v;
    for (int p = g->rowptr[v]; p < g->rowptr[v + 1]; p++) {
      const float* p2 = pos + 3 * (size_t)g->col[p];
      const int e = g->eid[p];
      const float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
      const float distance = std::sqrt((dx * dx + dy * dy) + dz * dz);  // (p1 - p2).norm()
      if (distance > g->max_distance[e]) g->max_distance[e] = distance;
      if (distance < g->min_distance[e]) g->min_distance[e] = distance;
      const float dm = g->max_distance[e], s = g->weight_sigma;
      g->weight[e] = std::exp(-(dm * dm) / (2 * s * s));  // geometry_toolbox.cc:26-28
      if (std::fabs((g->max_distance[e] - g->min_distance[e]) / g->min_distance[e]) > g->stretching_th)
        g->status[e] = NRSLAM_EDGE_BAD;
      else
        n_good++;
    }
    good_out[i] = n_good;
  }
  return 0;
}

}  // extern "C"
