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

extern "C" int tri_emul_run(const nrslam_b200_camera* cam_, int32_t n_cand, const int32_t* track_ptr,
                            const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                            const float* nb_pos, const uint8_t* nb_valid, float* position_out, int32_t* status_out,
                            int32_t* lm_iterations_out) {
  nrs::Cam cam;
  cam.model = cam_->model;
  for (int i = 0; i < 8; i++) cam.p[i] = cam_->params[i];
  std::vector<double> smem;
  for (int c = 0; c < n_cand; c++) {
    const int e0 = track_ptr[c], T = track_ptr[c + 1] - e0;
    if (T < 1 || T > nrs::tri::kMaxTrack) return -3;
    smem.assign(nrs::tri::work_bytes(T) / 8 + 2, 0.0);
    int st = -1, it = 0;
    nrs::tri::solve_candidate(cam, T, track_uv + 2 * (size_t)e0, track_pose + 7 * (size_t)e0, n_neighbours[c],
                              nb_pos + (size_t)e0 * nrs::tri::kNB * 3, nb_valid + (size_t)e0 * nrs::tri::kNB,
                              smem.data(), position_out + 3 * (size_t)c, &st, &it);
    status_out[c] = st;
    if (lm_iterations_out) lm_iterations_out[c] = it;
  }
  return 0;
}
