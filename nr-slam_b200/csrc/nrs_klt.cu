// nrs_klt.cu — LucasKanadeTracker entry points (modules/matching/lucas_kanade_tracker.h:55-92).
// Placeholder until the per-patch kernel lands: every entry point reports "not implemented" (-6) loudly;
// nothing here computes on the CPU.
#include "nrs_host.h"

#define NRSLAM_B200_ERR_UNIMPLEMENTED (-6)

struct nrslam_b200_klt {
  nrslam_b200_ctx* ctx;
};

extern "C" {
int nrslam_b200_klt_create(nrslam_b200_ctx* ctx, int32_t, int32_t, int32_t, float, float, nrslam_b200_klt** out) {
  if (out) *out = nullptr;
  if (ctx) ctx->err = "KLT kernels not implemented yet";
  return NRSLAM_B200_ERR_UNIMPLEMENTED;
}
void nrslam_b200_klt_destroy(nrslam_b200_klt* klt) { delete klt; }
int nrslam_b200_klt_set_reference(nrslam_b200_klt*, const uint8_t*, int32_t, int32_t, int32_t, int32_t, const float*,
                                  const uint8_t*, int32_t) {
  return NRSLAM_B200_ERR_UNIMPLEMENTED;
}
int nrslam_b200_klt_track(nrslam_b200_klt*, const uint8_t*, int32_t, int32_t, int32_t, int32_t, float*, uint8_t*,
                          int32_t, float, const uint8_t*, int32_t, int32_t*) {
  return NRSLAM_B200_ERR_UNIMPLEMENTED;
}
int nrslam_b200_klt_get_patch(nrslam_b200_klt*, int32_t, int16_t*, int16_t*, float*, float*, uint8_t*) {
  return NRSLAM_B200_ERR_UNIMPLEMENTED;
}
int nrslam_b200_klt_insert_patch(nrslam_b200_klt*, float, float, const int16_t*, const int16_t*, const float*,
                                 const float*, const uint8_t*) {
  return NRSLAM_B200_ERR_UNIMPLEMENTED;
}
int nrslam_b200_klt_clear(nrslam_b200_klt*) { return NRSLAM_B200_ERR_UNIMPLEMENTED; }
int32_t nrslam_b200_klt_num_points(const nrslam_b200_klt*) { return 0; }
int nrslam_b200_klt_retrack(nrslam_b200_klt*, float*) { return NRSLAM_B200_ERR_UNIMPLEMENTED; }
}
