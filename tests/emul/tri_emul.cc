// TEST INFRASTRUCTURE — never part of libnrslam_b200.so and never reachable from the product path.
// Compiles the per-candidate DeformableTriangulation routine of nr-slam_b200/csrc/nrs_tri_core.cuh for the host with
// ONE emulated thread (TRI_NT = 1, barriers are no-ops): the code between two barriers is data-race free, so this
// executes the same arithmetic the CTA does, sequentially. tests/test_tri_emulation.py compares it with the oracle so
// that the algorithm (not the parallel schedule) is verified on machines without a GPU.
#define NRS_TRI_HOST_EMULATION 1
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/nrslam_b200.h"
#include "../../nr-slam_b200/csrc/nrs_tri_core.cuh"

// rigid_ok == nullptr: DeformableTriangulation only (nrslam_b200_tri_run); otherwise the frame mode of
// nrslam_b200_tri_run_frame without the host-side vote.
extern "C" int tri_emul_run_frame(const nrslam_b200_camera* cam_, int32_t n_cand, const int32_t* track_ptr,
                                  const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                                  const float* nb_pos, const uint8_t* nb_valid, const uint8_t* rigid_ok,
                                  float rad_per_pixel, int32_t min_track, float* position_out, int32_t* status_out,
                                  int32_t* lm_iterations_out, float* rigid_out, int32_t* rigid_status) {
  nrs::Cam cam;
  cam.model = cam_->model;
  for (int i = 0; i < 8; i++) cam.p[i] = cam_->params[i];
  std::vector<double> smem;
  for (int c = 0; c < n_cand; c++) {
    const int e0 = track_ptr[c], T = track_ptr[c + 1] - e0;
    if (T < 1 || T > nrs::tri::kMaxTrack) return -3;
    smem.assign(nrs::tri::work_bytes(T) / 8 + 2, 0.0);
    int st = -1, it = 0;
    nrs::tri::RigidArgs rg;
    rg.enabled = rigid_ok != nullptr;
    rg.min_track = min_track;
    rg.rad_per_pixel = rad_per_pixel;
    rg.rigid_ok = rg.enabled ? rigid_ok[c] : 0;
    rg.out = rg.enabled ? rigid_out + 3 * (size_t)c : nullptr;
    rg.status = rg.enabled ? rigid_status + c : nullptr;
    nrs::tri::solve_candidate(cam, T, track_uv + 2 * (size_t)e0, track_pose + 7 * (size_t)e0, n_neighbours[c],
                              nb_pos + (size_t)e0 * nrs::tri::kNB * 3, nb_valid + (size_t)e0 * nrs::tri::kNB,
                              smem.data(), position_out + 3 * (size_t)c, &st, &it, rg);
    status_out[c] = st;
    if (lm_iterations_out) lm_iterations_out[c] = it;
  }
  return 0;
}

extern "C" int tri_emul_run(const nrslam_b200_camera* cam_, int32_t n_cand, const int32_t* track_ptr,
                            const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                            const float* nb_pos, const uint8_t* nb_valid, float* position_out, int32_t* status_out,
                            int32_t* lm_iterations_out) {
  return tri_emul_run_frame(cam_, n_cand, track_ptr, track_uv, track_pose, n_neighbours, nb_pos, nb_valid, nullptr, 0.f,
                            1, position_out, status_out, lm_iterations_out, nullptr, nullptr);
}
