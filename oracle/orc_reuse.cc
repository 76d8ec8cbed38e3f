// ORACLE — test infrastructure only (see orc_math.h): CPU restatement of Tracking::PointReuse,
// modules/tracking/tracking.cc:394-506, on top of the LucasKanadeTracker restatement (orc_klt.cc). PARITY UNPINNED:
// the reference ships no test or stored output for this function.
//
// Follows the reference statement by statement:
//   :397-414  every map point that has no 3-D observation in the frame (or that the optimiser reported lost) is
//             projected through the frame pose (Sophus::SE3f * Vector3f in fp32 = Eigen::Quaternion::_transformVector:
//             v + w t + q x t with t = 2 q x v; Eigen is un-vendored, its published formula is restated), kept when
//             depth >= 0 and the projection lies inside the image;
//   :421-458  a fresh tracker with maxLevel 1 gets the stored PhotometricInformation of each candidate
//             (InsertPhotometricInformation) and tracks with the projection as initial flow, minSSIM 0.75;
//   :460-479  survivors (status TRACKED_WITH_3D) are accepted unless SquaredReprojectionError(projection, keypoint) > 5.99
//             (utilities/geometry_toolbox.cc:30-35).
// The reference walks an absl::flat_hash_set (unspecified order); candidates are taken in ascending index, like the
// product (documented deviation, results are order independent: every candidate is tracked on its own).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/nrslam_b200.h"
#include "orc_math.h"

extern "C" {
void* orc_klt_create(int win, int max_level, int max_iters, float eps, float min_eig);
void orc_klt_destroy(void* p);
int orc_klt_insert_patch(void* p, float x, float y, const int16_t* gray, const int16_t* grad, const float* mean,
                         const float* mean2, const uint8_t* valid);
int orc_klt_track(void* p, const uint8_t* img, int w, int h, int pitch, int n, float* pts_io, uint8_t* status_io,
                  int use_initial_flow, float min_ssim, const uint8_t* mask, int mask_pitch, int* n_tracked);

int orc_point_reuse(const nrslam_b200_camera* cam, const float* pose, const uint8_t* image, int32_t width,
                    int32_t height, int32_t pitch, const uint8_t* mask, int32_t mask_pitch, int32_t n,
                    const float* X_world, const uint8_t* in_frame, const uint8_t* forced, int32_t max_iters,
                    float epsilon, float min_eig_threshold, const int16_t* gray, const int16_t* grad,
                    const float* mean, const float* mean2, const uint8_t* valid, int32_t* cand_out, float* seed_out,
                    float* uv_out, uint8_t* status_out, uint8_t* accepted_out, int32_t* n_cand_out,
                    int32_t* n_reused_out) {
  orc::Camera c;
  c.model = cam->model;
  for (int i = 0; i < 8; i++) c.p[i] = cam->params[i];
  const float q[3] = {pose[0], pose[1], pose[2]}, w = pose[3];
  const int win = 21;
  const size_t A = (size_t)win * win;
  void* klt = orc_klt_create(win, 1, max_iters, epsilon, min_eig_threshold);  // :424-426
  std::vector<int> cand;
  std::vector<float> seeds;
  for (int i = 0; i < n; i++) {
    if (in_frame[i] && !(forced && forced[i])) continue;  // :398 / the ids handed in by the caller
    const float* v = X_world + 3 * (size_t)i;
    float t[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    for (int a = 0; a < 3; a++) t[a] += t[a];
    const float cr[3] = {q[1] * t[2] - q[2] * t[1], q[2] * t[0] - q[0] * t[2], q[0] * t[1] - q[1] * t[0]};
    float pc[3];
    for (int a = 0; a < 3; a++) pc[a] = ((v[a] + w * t[a]) + cr[a]) + pose[4 + a];
    if (pc[2] < 0) continue;  // :403-405
    float uv[2];
    orc::project_f(c, pc, uv);
    if (!(uv[0] >= 0 && uv[0] < (float)width && uv[1] >= 0 && uv[1] < (float)height)) continue;  // :409-412
    cand.push_back(i);
    seeds.push_back(uv[0]);
    seeds.push_back(uv[1]);
    orc_klt_insert_patch(klt, uv[0], uv[1], gray + (size_t)i * 2 * A, grad + (size_t)i * 4 * A, mean + 2 * (size_t)i,
                         mean2 + 2 * (size_t)i, valid + 2 * (size_t)i);  // :446-449
  }
  const int m = (int)cand.size();
  *n_cand_out = m;
  *n_reused_out = 0;
  if (m == 0) {  // :452-454
    orc_klt_destroy(klt);
    return 0;
  }
  std::vector<float> pts(seeds);
  std::vector<uint8_t> st(m, NRSLAM_TRACKED_WITH_3D);
  int n_tracked = 0;
  orc_klt_track(klt, image, width, height, pitch, m, pts.data(), st.data(), 1, 0.75f, mask, mask_pitch, &n_tracked);
  orc_klt_destroy(klt);
  for (int j = 0; j < m; j++) {
    cand_out[j] = cand[j];
    if (seed_out) { seed_out[2 * j] = seeds[2 * j]; seed_out[2 * j + 1] = seeds[2 * j + 1]; }
    uv_out[2 * j] = pts[2 * j];
    uv_out[2 * j + 1] = pts[2 * j + 1];
    if (status_out) status_out[j] = st[j];
    const float errx = seeds[2 * j] - pts[2 * j], erry = seeds[2 * j + 1] - pts[2 * j + 1];
    accepted_out[j] = (st[j] == NRSLAM_TRACKED_WITH_3D && !(errx * errx + erry * erry > 5.99f)) ? 1 : 0;  // :474-479
    *n_reused_out += accepted_out[j];
  }
  return 0;
}
}
