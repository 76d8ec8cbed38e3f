// TEST INFRASTRUCTURE: a recording stand-in for libnrslam_b200.so so that the shim can be driven on a machine without
// a GPU. Every entry point validates the buffers the shim hands over (CSR invariants, ordering contracts of
// include/nrslam_b200.h) and returns a deterministic, recognisable result the driver then looks for in the
// Frame / Map / KeyFrame objects. Nothing here is product code.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "nrslam_b200.h"

struct nrslam_b200_ctx { int dummy; };
struct nrslam_b200_tri { int dummy; };
struct nrslam_b200_klt { int win, levels, n; std::vector<float> pts; };

static int g_errors = 0;
#define EXPECT(c) do { if (!(c)) { fprintf(stderr, "mock_abi: expectation failed: %s (line %d)\n", #c, __LINE__); g_errors++; } } while (0)

static void check_graph(const nrslam_b200_graph* g) {
  EXPECT(g && g->n_vertices >= 0 && g->rowptr && g->rowptr[0] == 0);
  std::vector<int> seen(g->n_edges, 0);
  for (int v = 0; v < g->n_vertices; v++) {
    EXPECT(g->rowptr[v + 1] >= g->rowptr[v]);
    for (int a = g->rowptr[v]; a < g->rowptr[v + 1]; a++) {
      EXPECT(g->col[a] >= 0 && g->col[a] < g->n_vertices && g->col[a] != v);
      if (a > g->rowptr[v]) EXPECT(g->col[a] > g->col[a - 1]);  // neighbours ascending
      EXPECT(g->eid[a] >= 0 && g->eid[a] < g->n_edges);
      seen[g->eid[a]]++;
      bool back = false;  // the same record from the other endpoint
      for (int b = g->rowptr[g->col[a]]; b < g->rowptr[g->col[a] + 1]; b++)
        if (g->col[b] == v && g->eid[b] == g->eid[a]) back = true;
      EXPECT(back);
    }
  }
  for (int e = 0; e < g->n_edges; e++) EXPECT(seen[e] == 2);
  EXPECT(g->rowptr[g->n_vertices] == 2 * g->n_edges);
  EXPECT(g->weight_sigma > 0 && std::fabs(g->stretching_th - 1.1f) < 1e-6f);
}

extern "C" {
int mock_errors(void) { return g_errors; }
int nrslam_b200_abi_version(void) { return NRSLAM_B200_ABI_VERSION; }
int nrslam_b200_create(const nrslam_b200_options*, nrslam_b200_ctx** out) { *out = new nrslam_b200_ctx{0}; return 0; }
void nrslam_b200_destroy(nrslam_b200_ctx* c) { delete c; }

int nrslam_b200_pose_only(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                          const float* X, float* pose_io, uint8_t*, nrslam_b200_stats*) {
  EXPECT(ctx && cam && cam->model == 0 && cam->params[0] == 500.f && n > 0 && uv && X);
  const float nq = pose_io[0] * pose_io[0] + pose_io[1] * pose_io[1] + pose_io[2] * pose_io[2] + pose_io[3] * pose_io[3];
  EXPECT(std::fabs(nq - 1.f) < 1e-5f);
  pose_io[4] += 0.25f;  // recognisable result
  return 0;
}

int nrslam_b200_pose_deform(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                            const float* X_rest, const int32_t* point_vertex, const int8_t* vfs, nrslam_b200_graph* g,
                            float scale, float* pose_io, float* last, float*, float* X_out, float*, uint8_t* status_out,
                            float* median, int32_t* lost_out, int32_t* n_lost, nrslam_b200_stats*) {
  EXPECT(ctx && cam && n > 0 && uv && X_rest && point_vertex && vfs && last && scale > 0);
  check_graph(g);
  for (int i = 0; i < n; i++) {
    EXPECT(point_vertex[i] >= 0 && point_vertex[i] < g->n_vertices);
    EXPECT(vfs[point_vertex[i]] == NRSLAM_TRACKED_WITH_3D);
    for (int k = 0; k < 3; k++) X_out[3 * i + k] = X_rest[3 * i + k] + 0.5f;
    status_out[i] = (i % 3 == 2) ? NRSLAM_TRACKED : NRSLAM_TRACKED_WITH_3D;
    last[3 * point_vertex[i]] += 1.0f;  // SetLastWorldPosition for this vertex
  }
  int nl = 0;
  for (int v = 0; v < g->n_vertices; v++)
    if (vfs[v] == NRSLAM_TRACKED) {  // the frame's TRACKED points come back as "lost"
      lost_out[nl++] = v;
      last[3 * v + 1] += 2.0f;
    }
  *n_lost = nl;
  for (int e = 0; e < g->n_edges; e++) {  // what UpdateVertex would touch
    g->weight[e] *= 0.5f;
    g->max_distance[e] += 1.0f;
    if (e == 0) g->status[e] = NRSLAM_EDGE_BAD;
  }
  pose_io[5] += 0.75f;
  *median = 0.125f;
  return 0;
}

int nrslam_b200_local_ba(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n_kf, float* kf_pose_io,
                         int32_t n_obs, const int32_t* obs_kf, const int32_t* obs_vertex, const float* uv, float* X_io,
                         const nrslam_b200_graph* g, float scale, int32_t iterations, nrslam_b200_stats*) {
  EXPECT(ctx && cam && n_kf >= 3 && n_kf <= 5 && n_obs > 0 && uv && scale > 0 && iterations <= 0);
  check_graph(g);
  for (int o = 0; o < n_obs; o++) {
    EXPECT(obs_kf[o] >= 0 && obs_kf[o] < n_kf);
    if (o) EXPECT(obs_kf[o] >= obs_kf[o - 1]);  // grouped by keyframe, oldest first
    EXPECT(obs_vertex[o] >= 0 && obs_vertex[o] < g->n_vertices);
    X_io[3 * o + 2] += 0.01f * (obs_kf[o] + 1);
  }
  for (int k = 0; k < n_kf; k++) {
    EXPECT(std::fabs(kf_pose_io[7 * k + 4] - (float)k) < 1e-6f);  // the driver encodes the age in tx: oldest = 0
    kf_pose_io[7 * k + 6] += 1.0f;
  }
  return 0;
}

int nrslam_b200_tri_create(nrslam_b200_ctx*, nrslam_b200_tri** out) { *out = new nrslam_b200_tri{0}; return 0; }
void nrslam_b200_tri_destroy(nrslam_b200_tri* t) { delete t; }
int nrslam_b200_tri_run(nrslam_b200_tri* tri, const nrslam_b200_camera* cam, int32_t n_cand, const int32_t* track_ptr,
                        const float* track_uv, const float* track_pose, const int32_t* n_nb, const float* nb_pos,
                        const uint8_t* nb_valid, float, float* pos, int32_t* status, int32_t*) {
  EXPECT(tri && cam && n_cand == 1 && track_ptr[0] == 0 && track_uv && track_pose && nb_pos && nb_valid);
  const int T = track_ptr[1];
  EXPECT(T == 6 && n_nb[0] == 2);
  int valid = 0;
  for (int f = 0; f < T; f++)
    for (int k = 0; k < NRSLAM_B200_TRI_MAX_NB; k++) valid += nb_valid[f * NRSLAM_B200_TRI_MAX_NB + k];
  EXPECT(valid == 2 * T - 1);  // the driver leaves one (frame, neighbour) position out
  if (track_uv[0] < 0) { status[0] = NRSLAM_B200_TRI_LOW_PARALLAX; return 0; }
  pos[0] = 1; pos[1] = 2; pos[2] = 3;
  status[0] = NRSLAM_B200_TRI_OK;
  return 0;
}

int nrslam_b200_klt_create(nrslam_b200_ctx* ctx, int32_t win, int32_t max_level, int32_t, float, float,
                           nrslam_b200_klt** out) {
  EXPECT(ctx && win == 21);
  *out = new nrslam_b200_klt{win, max_level + 1, 0, {}};
  return 0;
}
void nrslam_b200_klt_destroy(nrslam_b200_klt* k) { delete k; }
int nrslam_b200_klt_set_reference(nrslam_b200_klt* k, const uint8_t* image, int32_t w, int32_t h, int32_t pitch,
                                  int32_t n, const float* xy, const uint8_t*, int32_t) {
  EXPECT(k && image && w == 64 && h == 48 && pitch >= w && xy);
  k->n = n;
  k->pts.assign(xy, xy + 2 * n);
  return 0;
}
int nrslam_b200_klt_track(nrslam_b200_klt* k, const uint8_t* image, int32_t, int32_t, int32_t, int32_t n, float* pts,
                          uint8_t* status, int32_t, float min_ssim, const uint8_t*, int32_t, int32_t* n_tracked) {
  EXPECT(k && image && n == k->n && min_ssim > 0);
  int t = 0;
  for (int i = 0; i < n; i++) {
    pts[2 * i] = k->pts[2 * i] + 1.5f;
    pts[2 * i + 1] = k->pts[2 * i + 1] - 0.5f;
    if (i == 1) status[i] = NRSLAM_BAD; else t++;
  }
  *n_tracked = t;
  return 0;
}
int nrslam_b200_klt_get_patch(nrslam_b200_klt* k, int32_t idx, int16_t* gray, int16_t* grad, float* mean, float* mean2,
                              uint8_t* valid) {
  EXPECT(k && idx >= 0 && idx < k->n);
  const int px = k->win * k->win;
  for (int l = 0; l < k->levels; l++) {
    valid[l] = l < k->levels - 1;
    mean[l] = 10.f * l + idx;
    mean2[l] = 100.f * l;
    for (int p = 0; p < px; p++) gray[l * px + p] = (int16_t)(l * 1000 + p);
    for (int p = 0; p < 2 * px; p++) grad[l * 2 * px + p] = (int16_t)(-p);
  }
  return 0;
}
int nrslam_b200_klt_insert_patch(nrslam_b200_klt* k, float x, float y, const int16_t* gray, const int16_t* grad,
                                 const float* mean, const float*, const uint8_t* valid) {
  const int px = k->win * k->win;
  EXPECT(valid[0] == 1 && valid[k->levels - 1] == 0 && gray[px + 5] == 1005 && grad[7] == -7 && mean[1] == 10.f);
  k->pts.push_back(x);
  k->pts.push_back(y);
  k->n++;
  return 0;
}
int nrslam_b200_klt_clear(nrslam_b200_klt* k) { k->n = 0; k->pts.clear(); return 0; }
}
