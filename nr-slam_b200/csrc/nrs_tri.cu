// nrs_tri.cu — batched DeformableTriangulation (one CTA per candidate) and the batched RegularizationGraph update.
//
// Reference: modules/optimization/g2o_optimization.cc:559-814 called per candidate from
// modules/mapping/mapping.cc:88-113; modules/map/regularization_graph.cc:89-146 called per accepted point from
// g2o_optimization.cc:458-474. The per-candidate routine lives in nrs_tri_core.cuh.
#include <algorithm>
#include <cmath>
#include <string>

#include "nrs_host.h"
#include "nrs_tri_core.cuh"

namespace {
using namespace nrs;

__global__ void __launch_bounds__(tri::kThreads)
nrs_tri_kernel(Cam cam, int n_cand, const int* __restrict__ track_ptr, const float* __restrict__ track_uv,
               const float* __restrict__ track_pose, const int* __restrict__ n_neighbours,
               const float* __restrict__ nb_pos, const unsigned char* __restrict__ nb_valid,
               const int* __restrict__ order, float* __restrict__ position_out, int* __restrict__ status_out,
               int* __restrict__ iters_out, const unsigned char* __restrict__ rigid_ok, float rad_per_pixel,
               int min_track, float* __restrict__ rigid_out, int* __restrict__ rigid_status) {
  extern __shared__ __align__(16) unsigned char tri_smem[];
  // longest tracks first (order[] sorts the candidates by descending track length) so the tail of the grid is cheap
  const int c = order[blockIdx.x];
  const int e0 = track_ptr[c], T = track_ptr[c + 1] - e0;
  tri::RigidArgs rg;
  rg.enabled = rigid_ok != nullptr;
  rg.min_track = min_track;
  rg.rad_per_pixel = rad_per_pixel;
  rg.rigid_ok = rg.enabled ? rigid_ok[c] : 0;
  rg.out = rg.enabled ? rigid_out + 3 * (size_t)c : nullptr;
  rg.status = rg.enabled ? rigid_status + c : nullptr;
  tri::solve_candidate(cam, T, track_uv + 2 * (size_t)e0, track_pose + 7 * (size_t)e0, n_neighbours[c],
                       nb_pos + (size_t)e0 * tri::kNB * 3, nb_valid + (size_t)e0 * tri::kNB, tri_smem,
                       position_out + 3 * (size_t)c, status_out + c, iters_out + c, rg);
}

// One thread per CSR entry of an updated vertex (regularization_graph.cc:107-128 UpdateConnection). Reads the edge
// records of the PREVIOUS state (in_*), writes the new state (out_*, pre-initialised with a copy of in_*): the
// positions are fixed during the loop of g2o_optimization.cc:458-474, so an edge between two updated vertices gets
// the same values from both sides and the sequential loop's result does not depend on its order. Only the smaller
// updated endpoint stores.
__global__ void nrs_graph_update_kernel(int n_entries, const int* __restrict__ ent_vertex_slot,
                                        const int* __restrict__ ent_csr, const int* __restrict__ vertices,
                                        const int* __restrict__ col, const int* __restrict__ eid,
                                        const unsigned char* __restrict__ is_updated, const float* __restrict__ pos,
                                        const float* __restrict__ in_min, const float* __restrict__ in_max,
                                        float sigma, float stretching_th, float* __restrict__ out_w,
                                        float* __restrict__ out_min, float* __restrict__ out_max,
                                        unsigned char* __restrict__ out_status, int* __restrict__ good) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_entries) return;
  const int slot = ent_vertex_slot[t], p = ent_csr[t];
  const int v = vertices[slot], u = col[p], e = eid[p];
  const float dx = NRS_FS(pos[3 * v], pos[3 * u]), dy = NRS_FS(pos[3 * v + 1], pos[3 * u + 1]),
              dz = NRS_FS(pos[3 * v + 2], pos[3 * u + 2]);
  const float distance = sqrtf(NRS_FA(NRS_FA(NRS_FM(dx, dx), NRS_FM(dy, dy)), NRS_FM(dz, dz)));
  float mx = in_max[e], mn = in_min[e];
  if (distance > mx) mx = distance;
  if (distance < mn) mn = distance;
  const bool bad = fabsf(NRS_FD(NRS_FS(mx, mn), mn)) > stretching_th;
  if (!bad) atomicAdd(good + slot, 1);
  if (!is_updated[u] || v < u) {
    // InterpolationWeight (geometry_toolbox.cc:26-28): expf of a float argument. fp64 exp rounded to fp32 is the
    // correctly rounded expf (up to double-rounding cases of probability ~2^-29); glibc's expf is faithfully rounded
    // and differs from it by 1 ulp for 0.07 % of the arguments (measured), which is the stated bar for this attribute.
    const float arg = NRS_FD(-NRS_FM(mx, mx), NRS_FM(NRS_FM(2.f, sigma), sigma));
    out_w[e] = (float)exp((double)arg);
    out_min[e] = mn;
    out_max[e] = mx;
    if (bad) out_status[e] = NRSLAM_EDGE_BAD;
  }
}

// GetEdges as a segmented top-k (regularization_graph.cc:61-87). One CTA per listed vertex: every entry of the row gets
// the number of entries that sort before it (keys are unique: the neighbour is part of the key); ranks below top_k are
// written, and the list is cut at the smallest rank whose weight is < min_weight.
__device__ __forceinline__ unsigned long long edge_key(unsigned status, float weight, int col) {
  const unsigned wb = 0xFFFFFFFFu - __float_as_uint(weight);  // non-negative floats order like their bit patterns
  return ((unsigned long long)(status & 3u) << 62) | ((unsigned long long)wb << 30) | (unsigned long long)(col & 0x3FFFFFFF);
}
__global__ void nrs_graph_keys_kernel(int nnz, const int* __restrict__ col, const int* __restrict__ eid,
                                      const float* __restrict__ weight, const unsigned char* __restrict__ status,
                                      unsigned long long* __restrict__ keys) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) keys[p] = edge_key(status[eid[p]], weight[eid[p]], col[p]);
}
__global__ void __launch_bounds__(128)
nrs_graph_topk_kernel(const int* __restrict__ vertices, const int* __restrict__ rowptr, const int* __restrict__ eid,
                      const float* __restrict__ weight, const unsigned long long* __restrict__ keys, float min_w,
                      int top_k, int* __restrict__ out_entries, int* __restrict__ out_count) {
  __shared__ int s_cut;
  __shared__ unsigned long long s_keys[1024];
  const int v = vertices[blockIdx.x];
  const int p0 = rowptr[v], d = rowptr[v + 1] - p0;
  if (threadIdx.x == 0) s_cut = d;
  for (int k = threadIdx.x; k < top_k; k += blockDim.x) out_entries[(size_t)blockIdx.x * top_k + k] = -1;
  const bool staged = d <= 1024;
  if (staged)
    for (int k = threadIdx.x; k < d; k += blockDim.x) s_keys[k] = keys[p0 + k];
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const unsigned long long mine = staged ? s_keys[k] : keys[p0 + k];
    int rank = 0;
    if (staged)
      for (int q = 0; q < d; q++) rank += (s_keys[q] < mine) ? 1 : 0;
    else
      for (int q = 0; q < d; q++) rank += (keys[p0 + q] < mine) ? 1 : 0;
    if (rank < top_k) out_entries[(size_t)blockIdx.x * top_k + rank] = p0 + k;
    if (weight[eid[p0 + k]] < min_w) atomicMin(&s_cut, rank);
  }
  __syncthreads();
  if (threadIdx.x == 0) out_count[blockIdx.x] = s_cut;
  // entries at or behind the cut are not part of the list
  for (int k = s_cut + threadIdx.x; k < top_k; k += blockDim.x) out_entries[(size_t)blockIdx.x * top_k + k] = -1;
}

int tfail(nrslam_b200_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
#define TRI_CUDA(ctx, call)                                                                            \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return tfail(ctx, NRSLAM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
  } while (0)
}  // namespace

struct nrslam_b200_tri {
  nrslam_b200_ctx* ctx = nullptr;
  nrs::Arena in, out;
  // staged batch
  int n_cand = 0, t_max = 0;
  size_t smem = 0;
  nrs::Cam cam;
  size_t o_ptr = 0, o_uv = 0, o_pose = 0, o_nnb = 0, o_pos = 0, o_val = 0, o_order = 0;
  size_t o_out = 0, o_status = 0, o_iters = 0;
  // frame mode (tri_run_frame): rigid branch inputs / outputs
  bool frame = false;
  size_t o_rok = 0, o_rout = 0, o_rstatus = 0;
  float rad_per_pixel = 0.f;
  int min_track = 1;
  float last_ms = 0.f;
  bool staged = false;
};

namespace {
int tri_launch(nrslam_b200_tri* t) {
  nrslam_b200_ctx* ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  TRI_CUDA(ctx, cudaFuncSetAttribute(nrs_tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem));
  TRI_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  nrs_tri_kernel<<<t->n_cand, tri::kThreads, t->smem, st>>>(
      t->cam, t->n_cand, t->in.d<int>(t->o_ptr), t->in.d<float>(t->o_uv), t->in.d<float>(t->o_pose),
      t->in.d<int>(t->o_nnb), t->in.d<float>(t->o_pos), t->in.d<unsigned char>(t->o_val), t->in.d<int>(t->o_order),
      t->out.d<float>(t->o_out), t->out.d<int>(t->o_status), t->out.d<int>(t->o_iters),
      t->frame ? t->in.d<unsigned char>(t->o_rok) : nullptr, t->rad_per_pixel, t->min_track,
      t->frame ? t->out.d<float>(t->o_rout) : nullptr, t->frame ? t->out.d<int>(t->o_rstatus) : nullptr);
  TRI_CUDA(ctx, cudaGetLastError());
  TRI_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
  return 0;
}
}  // namespace

extern "C" {

int nrslam_b200_tri_create(nrslam_b200_ctx* ctx, nrslam_b200_tri** out) {
  if (out) *out = nullptr;
  if (!ctx || !out) return NRSLAM_B200_ERR_ARG;
  nrslam_b200_tri* t = new nrslam_b200_tri();
  t->ctx = ctx;
  *out = t;
  return 0;
}

void nrslam_b200_tri_destroy(nrslam_b200_tri* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->stream);
  delete t;
}

// Stage a batch, run the kernel, bring the results into the pinned output arena. rigid_ok == nullptr: deformable only.
static int tri_stage_and_run(nrslam_b200_tri* t, const nrslam_b200_camera* cam, int32_t n_cand, const int32_t* track_ptr,
                             const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                             const float* nb_pos, const uint8_t* nb_valid, const uint8_t* rigid_ok,
                             float rad_per_pixel, int32_t min_track) {
  if (!t) return NRSLAM_B200_ERR_ARG;
  nrslam_b200_ctx* ctx = t->ctx;
  if (!cam || n_cand <= 0 || !track_ptr)
    return tfail(ctx, NRSLAM_B200_ERR_ARG, "tri_run: bad argument");
  if (!track_uv || !track_pose || !n_neighbours || !nb_pos || !nb_valid)
    return tfail(ctx, NRSLAM_B200_ERR_ARG, "tri_run: bad argument");
  int t_max = 0;
  for (int c = 0; c < n_cand; c++) {
    const int T = track_ptr[c + 1] - track_ptr[c];
    if (T < 1 || T > NRSLAM_B200_TRI_MAX_TRACK || n_neighbours[c] < 0 || n_neighbours[c] > NRSLAM_B200_TRI_MAX_NB)
      return tfail(ctx, NRSLAM_B200_ERR_ARG, "tri_run: track length or neighbour count out of range");
    t_max = std::max(t_max, T);
  }
  if (track_ptr[0] != 0) return tfail(ctx, NRSLAM_B200_ERR_ARG, "tri_run: track_ptr[0] must be 0");
  const size_t E = (size_t)track_ptr[n_cand];
  TRI_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;

  const size_t need_in = (n_cand + 1 + 2 * (size_t)n_cand) * 4 + E * (2 + 7 + 3 * tri::kNB) * 4 + E * tri::kNB + n_cand + 4096;
  if (!t->in.reserve(need_in, true)) return tfail(ctx, NRSLAM_B200_ERR_ALLOC, "tri_run: input arena allocation failed");
  t->o_ptr = t->in.take<int>(n_cand + 1);
  t->o_nnb = t->in.take<int>(n_cand);
  t->o_order = t->in.take<int>(n_cand);
  t->o_uv = t->in.take<float>(2 * E);
  t->o_pose = t->in.take<float>(7 * E);
  t->o_pos = t->in.take<float>(3 * tri::kNB * E);
  t->o_val = t->in.take<unsigned char>(tri::kNB * E);
  memcpy(t->in.h<int>(t->o_ptr), track_ptr, (n_cand + 1) * sizeof(int));
  memcpy(t->in.h<int>(t->o_nnb), n_neighbours, n_cand * sizeof(int));
  memcpy(t->in.h<float>(t->o_uv), track_uv, 2 * E * sizeof(float));
  memcpy(t->in.h<float>(t->o_pose), track_pose, 7 * E * sizeof(float));
  memcpy(t->in.h<float>(t->o_pos), nb_pos, 3 * tri::kNB * E * sizeof(float));
  memcpy(t->in.h<unsigned char>(t->o_val), nb_valid, tri::kNB * E);
  t->frame = rigid_ok != nullptr;
  t->rad_per_pixel = rad_per_pixel;
  t->min_track = min_track;
  if (t->frame) {
    t->o_rok = t->in.take<unsigned char>(n_cand);
    memcpy(t->in.h<unsigned char>(t->o_rok), rigid_ok, n_cand);
  }
  int* order = t->in.h<int>(t->o_order);
  for (int c = 0; c < n_cand; c++) order[c] = c;
  std::stable_sort(order, order + n_cand, [&](int a, int b) {
    return track_ptr[a + 1] - track_ptr[a] > track_ptr[b + 1] - track_ptr[b];
  });
  const size_t need_out = (size_t)n_cand * 9 * 4 + 8192;
  if (!t->out.reserve(need_out, true)) return tfail(ctx, NRSLAM_B200_ERR_ALLOC, "tri_run: output arena allocation failed");
  t->o_out = t->out.take<float>(3 * (size_t)n_cand);
  t->o_status = t->out.take<int>(n_cand);
  t->o_iters = t->out.take<int>(n_cand);
  if (t->frame) {
    t->o_rout = t->out.take<float>(3 * (size_t)n_cand);
    t->o_rstatus = t->out.take<int>(n_cand);
  }
  t->n_cand = n_cand;
  t->t_max = t_max;
  t->smem = tri::work_bytes(t_max);
  t->cam.model = cam->model;
  for (int i = 0; i < 8; i++) t->cam.p[i] = cam->params[i];
  t->staged = false;

  TRI_CUDA(ctx, cudaMemcpyAsync(t->in.dev(), t->in.host(), t->in.used(), cudaMemcpyHostToDevice, st));
  const int rc = tri_launch(t);
  if (rc) return rc;
  TRI_CUDA(ctx, cudaMemcpyAsync(t->out.host(), t->out.dev(), t->out.used(), cudaMemcpyDeviceToHost, st));
  TRI_CUDA(ctx, cudaStreamSynchronize(st));
  TRI_CUDA(ctx, cudaEventElapsedTime(&t->last_ms, ctx->ev0, ctx->ev1));
  t->staged = true;
  return 0;
}

int nrslam_b200_tri_run(nrslam_b200_tri* t, const nrslam_b200_camera* cam, int32_t n_cand, const int32_t* track_ptr,
                        const float* track_uv, const float* track_pose, const int32_t* n_neighbours,
                        const float* nb_pos, const uint8_t* nb_valid, float scale, float* position_out,
                        int32_t* status_out, int32_t* lm_iterations_out) {
  (void)scale;  // unused by the reference body as well (g2o_optimization.cc:559-814 never reads it)
  if (!t) return NRSLAM_B200_ERR_ARG;
  if (n_cand < 0 || !position_out || !status_out) return tfail(t->ctx, NRSLAM_B200_ERR_ARG, "tri_run: bad argument");
  if (n_cand == 0) return 0;
  const int rc = tri_stage_and_run(t, cam, n_cand, track_ptr, track_uv, track_pose, n_neighbours, nb_pos, nb_valid,
                                   nullptr, 0.f, 1);
  if (rc) return rc;
  memcpy(position_out, t->out.h<float>(t->o_out), 3 * (size_t)n_cand * sizeof(float));
  memcpy(status_out, t->out.h<int>(t->o_status), n_cand * sizeof(int));
  if (lm_iterations_out) memcpy(lm_iterations_out, t->out.h<int>(t->o_iters), n_cand * sizeof(int));
  return 0;
}

int nrslam_b200_tri_run_frame(nrslam_b200_tri* t, const nrslam_b200_camera* cam, int32_t n_cand,
                              const int32_t* track_ptr, const float* track_uv, const float* track_pose,
                              const int32_t* n_neighbours, const float* nb_pos, const uint8_t* nb_valid,
                              const uint8_t* rigid_ok, float rad_per_pixel, int32_t min_track, float scale,
                              float* deform_pos_out, int32_t* deform_status_out, float* rigid_pos_out,
                              int32_t* rigid_status_out, float* selected_pos_out, uint8_t* selected_out) {
  (void)scale;
  if (!t) return NRSLAM_B200_ERR_ARG;
  if (n_cand < 0 || !rigid_ok || !deform_pos_out || !deform_status_out || !rigid_pos_out || !rigid_status_out ||
      !selected_pos_out || !selected_out)
    return tfail(t->ctx, NRSLAM_B200_ERR_ARG, "tri_run_frame: bad argument");
  if (n_cand == 0) return 0;
  const int rc = tri_stage_and_run(t, cam, n_cand, track_ptr, track_uv, track_pose, n_neighbours, nb_pos, nb_valid,
                                   rigid_ok, rad_per_pixel, min_track);
  if (rc) return rc;
  const float* dp = t->out.h<float>(t->o_out);
  const int* ds = t->out.h<int>(t->o_status);
  const float* rp = t->out.h<float>(t->o_rout);
  const int* rs = t->out.h<int>(t->o_rstatus);
  memcpy(deform_pos_out, dp, 3 * (size_t)n_cand * sizeof(float));
  memcpy(deform_status_out, ds, n_cand * sizeof(int));
  memcpy(rigid_pos_out, rp, 3 * (size_t)n_cand * sizeof(float));
  memcpy(rigid_status_out, rs, n_cand * sizeof(int));
  // the vote of mapping.cc:188-212 (host: two counts and a per-candidate pick)
  int n_rigid = 0, n_def = 0;
  for (int c = 0; c < n_cand; c++) {
    n_rigid += rs[c] == NRSLAM_B200_TRI_OK;
    n_def += ds[c] == NRSLAM_B200_TRI_OK;
  }
  for (int c = 0; c < n_cand; c++) {
    const float* pick = nullptr;
    if (n_rigid > 1.5 * n_def) {
      if (rs[c] == NRSLAM_B200_TRI_OK) pick = rp + 3 * (size_t)c;
    } else if (n_def >= 1.5 * n_rigid) {
      if (ds[c] == NRSLAM_B200_TRI_OK) pick = dp + 3 * (size_t)c;
    }
    if (pick && (std::isnan(pick[0]) || std::isnan(pick[1]) || std::isnan(pick[2]))) pick = nullptr;
    selected_out[c] = pick ? 1 : 0;
    for (int i = 0; i < 3; i++) selected_pos_out[3 * (size_t)c + i] = pick ? pick[i] : 0.f;
  }
  return 0;
}

float nrslam_b200_tri_last_ms(const nrslam_b200_tri* t) { return t ? t->last_ms : 0.f; }

int nrslam_b200_tri_rerun(nrslam_b200_tri* t, float* gpu_ms_out) {
  if (!t || !t->staged) return tfail(t ? t->ctx : nullptr, NRSLAM_B200_ERR_ARG, "tri_rerun: nothing staged");
  nrslam_b200_ctx* ctx = t->ctx;
  TRI_CUDA(ctx, cudaSetDevice(ctx->device));
  const int rc = tri_launch(t);
  if (rc) return rc;
  TRI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  TRI_CUDA(ctx, cudaEventElapsedTime(&t->last_ms, ctx->ev0, ctx->ev1));
  if (gpu_ms_out) *gpu_ms_out = t->last_ms;
  return 0;
}

int nrslam_b200_graph_update_vertices(nrslam_b200_ctx* ctx, nrslam_b200_graph* g, int32_t n, const int32_t* vertices,
                                      const float* positions, int32_t* good_out) {
  if (!ctx || !g || n < 0 || (n > 0 && (!vertices || !good_out)) || !positions)
    return tfail(ctx, NRSLAM_B200_ERR_ARG, "graph_update_vertices: bad argument");
  if (n == 0) return 0;
  const int V = g->n_vertices, E = g->n_edges;
  std::vector<unsigned char> upd(V, 0);
  size_t n_ent = 0;
  for (int i = 0; i < n; i++) {
    const int v = vertices[i];
    if (v < 0 || v >= V || upd[v]) return tfail(ctx, NRSLAM_B200_ERR_ARG, "graph_update_vertices: bad or repeated vertex");
    upd[v] = 1;
    n_ent += (size_t)(g->rowptr[v + 1] - g->rowptr[v]);
  }
  TRI_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  nrs::Arena& in = ctx->graph_in;
  nrs::Arena& out = ctx->graph_out;
  const size_t nnz = (size_t)g->rowptr[V];
  const size_t need_in = (2 * n_ent + n + 2 * nnz + 3 * (size_t)V + 2 * (size_t)E) * 4 + V + 8192;
  const size_t need_out = ((size_t)3 * E + n) * 4 + E + 8192;
  if (!in.reserve(need_in, true) || !out.reserve(need_out, true))
    return tfail(ctx, NRSLAM_B200_ERR_ALLOC, "graph_update_vertices: allocation failed");
  const size_t o_slot = in.take<int>(n_ent), o_csr = in.take<int>(n_ent), o_vert = in.take<int>(n),
               o_col = in.take<int>(nnz), o_eid = in.take<int>(nnz), o_pos = in.take<float>(3 * (size_t)V),
               o_min = in.take<float>(E), o_max = in.take<float>(E), o_upd = in.take<unsigned char>(V);
  const size_t q_w = out.take<float>(E), q_min = out.take<float>(E), q_max = out.take<float>(E),
               q_good = out.take<int>(n), q_status = out.take<unsigned char>(E);
  {
    int* slot = in.h<int>(o_slot);
    int* csr = in.h<int>(o_csr);
    size_t k = 0;
    for (int i = 0; i < n; i++)
      for (int p = g->rowptr[vertices[i]]; p < g->rowptr[vertices[i] + 1]; p++) {
        slot[k] = i;
        csr[k++] = p;
      }
  }
  memcpy(in.h<int>(o_vert), vertices, n * sizeof(int));
  memcpy(in.h<int>(o_col), g->col, nnz * sizeof(int));
  memcpy(in.h<int>(o_eid), g->eid, nnz * sizeof(int));
  memcpy(in.h<float>(o_pos), positions, 3 * (size_t)V * sizeof(float));
  memcpy(in.h<float>(o_min), g->min_distance, E * sizeof(float));
  memcpy(in.h<float>(o_max), g->max_distance, E * sizeof(float));
  memcpy(in.h<unsigned char>(o_upd), upd.data(), V);
  memcpy(out.h<float>(q_w), g->weight, E * sizeof(float));
  memcpy(out.h<float>(q_min), g->min_distance, E * sizeof(float));
  memcpy(out.h<float>(q_max), g->max_distance, E * sizeof(float));
  memset(out.h<int>(q_good), 0, n * sizeof(int));
  memcpy(out.h<unsigned char>(q_status), g->status, E);
  TRI_CUDA(ctx, cudaMemcpyAsync(in.dev(), in.host(), in.used(), cudaMemcpyHostToDevice, st));
  TRI_CUDA(ctx, cudaMemcpyAsync(out.dev(), out.host(), out.used(), cudaMemcpyHostToDevice, st));
  if (n_ent > 0) {
    const int blk = 256, grd = (int)((n_ent + blk - 1) / blk);
    nrs_graph_update_kernel<<<grd, blk, 0, st>>>((int)n_ent, in.d<int>(o_slot), in.d<int>(o_csr), in.d<int>(o_vert),
                                                 in.d<int>(o_col), in.d<int>(o_eid), in.d<unsigned char>(o_upd),
                                                 in.d<float>(o_pos), in.d<float>(o_min), in.d<float>(o_max),
                                                 g->weight_sigma, g->stretching_th, out.d<float>(q_w),
                                                 out.d<float>(q_min), out.d<float>(q_max),
                                                 out.d<unsigned char>(q_status), out.d<int>(q_good));
    TRI_CUDA(ctx, cudaGetLastError());
  }
  TRI_CUDA(ctx, cudaMemcpyAsync(out.host(), out.dev(), out.used(), cudaMemcpyDeviceToHost, st));
  TRI_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(g->weight, out.h<float>(q_w), E * sizeof(float));
  memcpy(g->min_distance, out.h<float>(q_min), E * sizeof(float));
  memcpy(g->max_distance, out.h<float>(q_max), E * sizeof(float));
  memcpy(g->status, out.h<unsigned char>(q_status), E);
  memcpy(good_out, out.h<int>(q_good), n * sizeof(int));
  return 0;
}

int nrslam_b200_graph_get_edges_batch(nrslam_b200_ctx* ctx, const nrslam_b200_graph* g, int32_t n,
                                      const int32_t* vertices, int32_t top_k, int32_t* out_entries,
                                      int32_t* out_count) {
  if (!ctx || !g || n < 0 || top_k < 1 || (n > 0 && (!vertices || !out_entries || !out_count)))
    return tfail(ctx, NRSLAM_B200_ERR_ARG, "graph_get_edges_batch: bad argument");
  if (n == 0) return 0;
  const int V = g->n_vertices, E = g->n_edges;
  for (int i = 0; i < n; i++)
    if (vertices[i] < 0 || vertices[i] >= V) return tfail(ctx, NRSLAM_B200_ERR_ARG, "graph_get_edges_batch: bad vertex");
  TRI_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  nrs::Arena& in = ctx->graph_in;
  nrs::Arena& out = ctx->graph_out;
  const size_t nnz = (size_t)g->rowptr[V];
  const size_t need_in = ((size_t)n + V + 1 + 2 * nnz + E) * 4 + E + 8192;
  const size_t need_out = ((size_t)n * top_k + n) * 4 + nnz * 8 + 8192;
  if (!in.reserve(need_in, true) || !out.reserve(need_out, true))
    return tfail(ctx, NRSLAM_B200_ERR_ALLOC, "graph_get_edges_batch: allocation failed");
  const size_t o_vert = in.take<int>(n), o_row = in.take<int>(V + 1), o_col = in.take<int>(nnz), o_eid = in.take<int>(nnz),
               o_w = in.take<float>(E), o_st = in.take<unsigned char>(E);
  const size_t q_ent = out.take<int>((size_t)n * top_k), q_cnt = out.take<int>(n);
  const size_t q_ret = out.used();  // only the part above travels back
  const size_t q_keys = out.take<unsigned long long>(nnz);
  memcpy(in.h<int>(o_vert), vertices, n * sizeof(int));
  memcpy(in.h<int>(o_row), g->rowptr, (V + 1) * sizeof(int));
  memcpy(in.h<int>(o_col), g->col, nnz * sizeof(int));
  memcpy(in.h<int>(o_eid), g->eid, nnz * sizeof(int));
  memcpy(in.h<float>(o_w), g->weight, E * sizeof(float));
  memcpy(in.h<unsigned char>(o_st), g->status, E);
  TRI_CUDA(ctx, cudaMemcpyAsync(in.dev(), in.host(), in.used(), cudaMemcpyHostToDevice, st));
  // min_weight_ = InterpolationWeight(1.5 sigma, sigma) (regularization_graph.cc:28-31), evaluated on the host like the
  // constructor does
  const float s = g->weight_sigma, dm = (float)(s * 1.5);
  const float min_w = std::exp(-(dm * dm) / (2 * s * s));
  if (nnz > 0)
    nrs_graph_keys_kernel<<<(int)((nnz + 255) / 256), 256, 0, st>>>((int)nnz, in.d<int>(o_col), in.d<int>(o_eid),
                                                                   in.d<float>(o_w), in.d<unsigned char>(o_st),
                                                                   out.d<unsigned long long>(q_keys));
  nrs_graph_topk_kernel<<<n, 128, 0, st>>>(in.d<int>(o_vert), in.d<int>(o_row), in.d<int>(o_eid), in.d<float>(o_w),
                                           out.d<unsigned long long>(q_keys), min_w, top_k, out.d<int>(q_ent),
                                           out.d<int>(q_cnt));
  TRI_CUDA(ctx, cudaGetLastError());
  TRI_CUDA(ctx, cudaMemcpyAsync(out.host(), out.dev(), q_ret, cudaMemcpyDeviceToHost, st));
  TRI_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(out_entries, out.h<int>(q_ent), (size_t)n * top_k * sizeof(int));
  memcpy(out_count, out.h<int>(q_cnt), n * sizeof(int));
  return 0;
}

}  // extern "C"
