// ORACLE — TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library, and only as the checker / reported baseline.
// Parity status: UNPINNED by the reference (no NR-SLAM tests exist); see orc_math.h.
//
// CPU restatement of the three optimisation drivers of modules/optimization/g2o_optimization.cc and of
// RegularizationGraph::GetEdges / UpdateVertex (modules/map/regularization_graph.cc:71-146).
// The entry points mirror include/nrslam_b200.h argument for argument (minus the ctx) so the parity tests
// feed identical buffers to both sides.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/nrslam_b200.h"
#include "orc_lm.h"

using namespace orc;

static double wall_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// utilities/geometry_toolbox.cc:26-28 (float arithmetic, std::exp(float))
static float interpolation_weight(float distance, float sigma) {
  return std::exp(-(distance * distance) / (2 * sigma * sigma));
}

static float min_weight(const nrslam_b200_graph* g) {
  // regularization_graph.cc:28-31: InterpolationWeight(weight_sigma * 1.5 [double -> float], weight_sigma)
  return interpolation_weight((float)(g->weight_sigma * 1.5), g->weight_sigma);
}

// regularization_graph.cc:61-87. Returns CSR entry indices. Tie-break on the neighbour index (E13).
static std::vector<int> get_edges(const nrslam_b200_graph* g, int vertex) {
  std::vector<int> ent;
  for (int p = g->rowptr[vertex]; p < g->rowptr[vertex + 1]; p++) ent.push_back(p);
  std::sort(ent.begin(), ent.end(), [&](int a, int b) {
    const int ea = g->eid[a], eb = g->eid[b];
    if (g->status[ea] != g->status[eb]) return g->status[ea] < g->status[eb];
    if (g->weight[ea] != g->weight[eb]) return g->weight[ea] > g->weight[eb];
    return g->col[a] < g->col[b];
  });
  const float mw = min_weight(g);
  std::vector<int> good;
  for (int p : ent) {
    if (g->weight[g->eid[p]] < mw) break;
    good.push_back(p);
  }
  return good;
}

// regularization_graph.cc:89-146
static int update_vertex(nrslam_b200_graph* g, int vertex, const float* pos) {
  int n_good = 0;
  const float* p1 = pos + 3 * vertex;
  for (int p = g->rowptr[vertex]; p < g->rowptr[vertex + 1]; p++) {
    const float* p2 = pos + 3 * g->col[p];
    const int e = g->eid[p];
    const float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
    const float distance = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (distance > g->max_distance[e]) g->max_distance[e] = distance;
    if (distance < g->min_distance[e]) g->min_distance[e] = distance;
    g->weight[e] = interpolation_weight(g->max_distance[e], g->weight_sigma);
    if (std::fabs((g->max_distance[e] - g->min_distance[e]) / g->min_distance[e]) > g->stretching_th) {
      g->status[e] = NRSLAM_EDGE_BAD;
    } else {
      n_good++;
    }
  }
  return n_good;
}

static void fill_stats(nrslam_b200_stats* st, const LMStats& s, double t0) {
  if (!st) return;
  st->lm_iterations += s.iterations;
  st->lm_trials += s.trials;
  st->lambda_final = s.lambda;
  for (double c : s.chi2_trace)
    if (st->n_trace < NRSLAM_B200_TRACE) st->chi2_trace[st->n_trace++] = c;
  st->host_ms = (float)((wall_s() - t0) * 1e3);
}

static Camera to_cam(const nrslam_b200_camera* c) {
  Camera cam;
  cam.model = c->model;
  for (int i = 0; i < 8; i++) cam.p[i] = c->params[i];
  return cam;
}

extern "C" {

void orc_default_options(nrslam_b200_options* o) {
  o->th_huber_2dof_sq = 5.99f;
  o->th_huber_3dof_sq = 0.584f;
  o->sigma_reprojection = 0.5f;
  o->sigma_position = 0.1f;
  o->sigma_spatial_factor = 0.1f;
  o->spring_k = 1.1f;
  o->regularizers_per_point = 10;
  o->pose_only_iterations[0] = o->pose_only_iterations[1] = o->pose_only_iterations[2] = 10;
  o->pose_deform_iterations[0] = o->pose_deform_iterations[1] = 10;
  o->lost_iterations = 10;
  o->ba_iterations = 5;
  o->lm_max_trials = 10;
  o->lm_tau = 1e-5;
  o->pcg_rel_tol = 1e-8;
  o->pcg_max_iterations = 2000;
  o->device = 0;
  o->grid_ctas = 0;
}

int32_t orc_graph_get_edges(const nrslam_b200_graph* g, int32_t vertex, int32_t* out, int32_t cap) {
  std::vector<int> e = get_edges(g, vertex);
  int n = std::min<int>((int)e.size(), cap);
  for (int i = 0; i < n; i++) out[i] = e[i];
  return (int32_t)e.size();
}

int32_t orc_graph_update_vertex(nrslam_b200_graph* g, int32_t vertex, const float* positions) {
  return update_vertex(g, vertex, positions);
}

// ---------------------------------------------------------------------------------------------
// g2o_optimization.cc:50-146
// ---------------------------------------------------------------------------------------------
int orc_pose_only(const nrslam_b200_options* opt, const nrslam_b200_camera* cam_, int32_t n, const float* uv,
                  const float* X, float* pose_io, uint8_t* inlier_out, nrslam_b200_stats* stats) {
  double t0 = wall_s();
  if (stats) *stats = nrslam_b200_stats{};
  Camera cam = to_cam(cam_);
  Optimizer optz(cam, /*dense*/ true);
  const float th2 = opt->th_huber_2dof_sq;
  const float th_huber_2dof = std::sqrt(th2);  // :64 (float sqrt)
  Vertex pv;
  pv.type = V_POSE;
  pv.dim = 6;
  const SE3 seed = se3_from_f7(pose_io);
  pv.pose = seed;
  optz.add_vertex(pv);
  for (int i = 0; i < n; i++) {
    Edge e;
    e.type = E_REPROJ_ONLY_POSE;
    e.nv = 1;
    e.v[0] = 0;
    e.dim = 2;
    e.info = 1.0;  // :90 Identity
    e.delta = th_huber_2dof;
    e.meas[0] = uv[2 * i];
    e.meas[1] = uv[2 * i + 1];
    for (int k = 0; k < 3; k++) e.Xw[k] = X[3 * i + k];
    optz.add_edge(e);
  }
  std::vector<char> inliers(n, 1);
  for (int it = 0; it < 3; it++) {
    optz.vertices[0].pose = seed;  // :108-110
    optz.initialize_optimization(0);
    optz.optimize(opt->pose_only_iterations[it]);
    for (int i = 0; i < n; i++) {
      Edge& e = optz.edges[i];
      if (!inliers[i]) optz.compute_error(e);  // :120-122 (inlier edges keep the error of the last trial)
      const float chi_squared = (float)optz.chi2(e);
      if (chi_squared > th2) {
        inliers[i] = 0;
        e.level = 1;
      } else {
        inliers[i] = 1;
        e.level = 0;
      }
      // :137-139 setRobustKernel(0) when it == 2: after the last optimize => no effect (E17)
    }
  }
  se3_to_f7(optz.vertices[0].pose, pose_io);
  if (inlier_out)
    for (int i = 0; i < n; i++) inlier_out[i] = inliers[i];
  if (stats) {
    stats->n_reproj_edges = n;
    stats->n_poses = 1;
  }
  fill_stats(stats, optz.stats, t0);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// g2o_optimization.cc:148-557
// debug_pcg: experiment knob (0 = exact Cholesky = the oracle; 1 = block-Jacobi PCG at opt->pcg_rel_tol)
// ---------------------------------------------------------------------------------------------
int orc_pose_deform_ex(const nrslam_b200_options* opt, const nrslam_b200_camera* cam_, int32_t n,
                       const float* uv, const float* X_rest, const int32_t* point_vertex,
                       const int8_t* vfs, nrslam_b200_graph* g, float scale, float* pose_io,
                       float* last_pos, float* deformation_out, float* X_out, float* chi2_out,
                       uint8_t* status_out, float* median_out, int32_t* lost_out, int32_t* n_lost_out,
                       nrslam_b200_stats* stats, int debug_pcg, double* timing_out) {
  double t0 = wall_s();
  if (stats) *stats = nrslam_b200_stats{};
  if (n_lost_out) *n_lost_out = 0;
  Camera cam = to_cam(cam_);
  Optimizer optz(cam, /*dense*/ false);
  optz.use_pcg = debug_pcg != 0;
  optz.pcg_tol = opt->pcg_rel_tol;
  optz.pcg_max_iter = opt->pcg_max_iterations;

  const int M = g->n_vertices;
  Vertex pv;
  pv.type = V_POSE;
  pv.dim = 6;
  const SE3 seed = se3_from_f7(pose_io);
  pv.pose = seed;
  optz.add_vertex(pv);
  std::vector<int> opt_index(M, -1);  // mappoint_id_to_index (:177,191)
  for (int i = 0; i < n; i++) {
    Vertex v;
    optz.add_vertex(v);  // id i+1, origin
    opt_index[point_vertex[i]] = i;
  }

  const int regularizers_per_point = opt->regularizers_per_point;
  const float th2 = opt->th_huber_2dof_sq;
  const float th_huber_2dof = std::sqrt(th2);
  const float th3 = opt->th_huber_3dof_sq;
  const float th_huber_3dof = std::sqrt(th3);
  const float sigma_reprojection = opt->sigma_reprojection;
  const float info_reprojection = 1.0f / (sigma_reprojection * sigma_reprojection);
  const float sigma_position = opt->sigma_position;
  const float info_position = 1.0f / (sigma_position * sigma_position);
  const float sigma_spatial = (float)((double)opt->sigma_spatial_factor * scale);  // :209 `0.1 * scale`
  const float info_spatial = 1.0f / (sigma_spatial * sigma_spatial);

  std::vector<std::unordered_map<int, int>> spatial(n);  // idx -> (idx_other -> edge index)
  std::vector<int> reproj(n);
  std::set<int> lost_ordered;
  double t_build0 = wall_s();
  for (int idx = 0; idx < n; idx++) {
    Edge e;
    e.type = E_REPROJ_DEFORM;
    e.nv = 2;
    e.v[0] = 0;
    e.v[1] = idx + 1;
    e.dim = 2;
    e.info = info_reprojection;
    e.delta = th_huber_2dof;
    e.meas[0] = uv[2 * idx];
    e.meas[1] = uv[2 * idx + 1];
    for (int k = 0; k < 3; k++) e.Xw[k] = X_rest[3 * idx + k];
    reproj[idx] = optz.add_edge(e);

    const int vtx = point_vertex[idx];
    std::vector<int> redges = get_edges(g, vtx);
    int n_regularizers = 0;
    for (int p : redges) {
      const int other = g->col[p], ge = g->eid[p];
      if (n_regularizers > regularizers_per_point || g->status[ge] == NRSLAM_EDGE_BAD) break;  // :258-261
      if (vfs[other] < 0 || vfs[other] != NRSLAM_TRACKED_WITH_3D) {                             // :264-273
        if (vfs[other] >= 0 && vfs[other] != NRSLAM_JUST_TRIANGULATED) lost_ordered.insert(other);
        continue;
      }
      const int idx_other = opt_index[other];
      if (spatial[idx].count(idx_other)) continue;  // :277-279
      Edge s;
      s.type = E_SPATIAL_DEFORM;
      s.nv = 2;
      s.v[0] = idx + 1;
      s.v[1] = idx_other + 1;
      s.dim = 3;
      s.info = info_spatial;
      s.delta = th_huber_3dof;
      s.weight = g->weight[ge];
      int si = optz.add_edge(s);
      spatial[idx][idx_other] = si;
      spatial[idx_other][idx] = si;
      n_regularizers++;
      Edge q;
      q.type = E_POSITION_DEFORM;
      q.nv = 2;
      q.v[0] = idx + 1;
      q.v[1] = idx_other + 1;
      q.dim = 1;
      q.info = info_position;
      q.delta = th_huber_3dof;
      q.meas[0] = g->first_distance[ge];
      for (int k = 0; k < 3; k++) {
        q.rest1[k] = X_rest[3 * idx + k];
        q.rest2[k] = X_rest[3 * idx_other + k];
      }
      q.k = 1.1f;  // :328  (float literal stored in a double)
      if (opt->spring_k != 1.1f) q.k = opt->spring_k;
      optz.add_edge(q);
    }
  }
  double t_build = wall_s() - t_build0;

  std::vector<char> inliers(n, 1);
  for (int it = 0; it < 2; it++) {
    optz.vertices[0].pose = seed;
    for (int i = 0; i < n; i++) optz.vertices[i + 1].x[0] = optz.vertices[i + 1].x[1] = optz.vertices[i + 1].x[2] = 0;
    optz.initialize_optimization(0);
    optz.optimize(opt->pose_deform_iterations[it]);
    for (int idx = 0; idx < n; idx++) {
      Edge& re = optz.edges[reproj[idx]];
      optz.compute_error(re);
      const float chi_squared = (float)optz.chi2(re);
      const int lvl = chi_squared > th2 ? 1 : 0;
      inliers[idx] = lvl == 0;
      re.level = lvl;
      for (auto& kv : spatial[idx]) optz.edges[kv.second].level = lvl;
      for (auto& kv : spatial[idx]) {  // :385-393 — own chi2 decides (E7)
        Edge& se = optz.edges[kv.second];
        optz.compute_error(se);
        se.level = (optz.chi2(se) > th3) ? 1 : 0;
      }
    }
  }

  se3_to_f7(optz.vertices[0].pose, pose_io);  // :398-399

  std::vector<float> mags(n);
  std::vector<std::array<float, 3>> def(n);
  for (int idx = 0; idx < n; idx++) {
    for (int k = 0; k < 3; k++) def[idx][k] = (float)optz.vertices[idx + 1].x[k];
    mags[idx] = std::sqrt(def[idx][0] * def[idx][0] + def[idx][1] * def[idx][1] + def[idx][2] * def[idx][2]);
    if (deformation_out)
      for (int k = 0; k < 3; k++) deformation_out[3 * idx + k] = def[idx][k];
  }
  std::vector<float> sorted = mags;
  std::sort(sorted.begin(), sorted.end());
  const float q1 = sorted[(int)(sorted.size() * 0.25f)];
  const float q3 = sorted[(int)(sorted.size() * 0.75f)];
  const float iqr = q3 - q1;
  const float th_ = 1.5f * iqr;

  for (int idx = 0; idx < n; idx++) {
    Edge& re = optz.edges[reproj[idx]];
    optz.compute_error(re);
    const float chi_squared = (float)optz.chi2(re);
    if (chi2_out) chi2_out[idx] = chi_squared;
    uint8_t status = NRSLAM_TRACKED_WITH_3D;
    if (chi_squared > th2) {
      inliers[idx] = 0;
      status = NRSLAM_TRACKED;
    }
    if (X_out)
      for (int k = 0; k < 3; k++) X_out[3 * idx + k] = X_rest[3 * idx + k];
    if (mags[idx] >= q3 + th_) {
      status = NRSLAM_TRACKED;
      if (status_out) status_out[idx] = status;
      continue;
    }
    optz.vertices[idx + 1].fixed = true;  // :439
    for (int k = 0; k < 3; k++) {
      const float cur = def[idx][k] + X_rest[3 * idx + k];
      if (X_out) X_out[3 * idx + k] = cur;
      last_pos[3 * point_vertex[idx] + k] = cur;  // :446
    }
    if (status_out) status_out[idx] = status;
  }
  {
    std::vector<float> m2 = mags;
    const int median_idx = (int)m2.size() / 2;
    std::nth_element(m2.begin(), m2.begin() + median_idx, m2.end());
    if (median_out) *median_out = m2[median_idx];
  }
  // :458-474
  for (int idx = 0; idx < n; idx++) {
    if (!inliers[idx]) continue;
    int good = update_vertex(g, point_vertex[idx], last_pos);
    if (good < regularizers_per_point * 0.5) {
      if (status_out) status_out[idx] = NRSLAM_BAD;
    }
  }
  if (stats) {
    stats->n_reproj_edges = n;
    stats->n_pair_edges = ((int)optz.edges.size() - n) / 2;
    stats->n_points = n;
    stats->n_poses = 1;
    stats->stage_ms = (float)(t_build * 1e3);
  }
  if (timing_out) {
    timing_out[0] = t_build;
    timing_out[1] = optz.stats.t_order;
    timing_out[2] = optz.stats.t_factor;
    timing_out[3] = optz.stats.t_build;
    timing_out[4] = (double)optz.pcg_iters_total;
  }
  if (lost_ordered.empty()) {
    fill_stats(stats, optz.stats, t0);
    return 0;
  }
  // :480-555
  std::vector<std::pair<int, int>> lost_vertex_idx;  // (graph vertex, optimizer vertex)
  for (int lost : lost_ordered) {
    Vertex v;
    int vid = optz.add_vertex(v);
    lost_vertex_idx.push_back({lost, vid});
    std::vector<int> redges = get_edges(g, lost);
    int n_regularizers = 0;
    for (int p : redges) {
      if (n_regularizers > 10) break;
      const int other = g->col[p];
      if (opt_index[other] < 0) continue;
      Edge s;
      s.type = E_SPATIAL_FIXED;
      s.nv = 1;
      s.v[0] = vid;
      s.dim = 3;
      s.info = info_spatial;
      s.delta = th_huber_3dof;
      s.weight = g->weight[g->eid[p]];
      s.ref_vertex = opt_index[other] + 1;
      optz.add_edge(s);
      n_regularizers++;
    }
  }
  optz.vertices[0].fixed = true;
  if (optz.initialize_optimization(0)) optz.optimize(opt->lost_iterations);
  int nl = 0;
  for (auto& lv : lost_vertex_idx) {
    for (int k = 0; k < 3; k++) last_pos[3 * lv.first + k] = (float)optz.vertices[lv.second].x[k] + last_pos[3 * lv.first + k];
    if (lost_out) lost_out[nl] = lv.first;
    nl++;
  }
  if (n_lost_out) *n_lost_out = nl;
  if (stats) stats->n_fixed_edges = (int)optz.edges.size() - n - 2 * stats->n_pair_edges;
  fill_stats(stats, optz.stats, t0);
  if (timing_out) {
    timing_out[1] = optz.stats.t_order;
    timing_out[2] = optz.stats.t_factor;
    timing_out[3] = optz.stats.t_build;
    timing_out[4] = (double)optz.pcg_iters_total;
  }
  return 0;
}

int orc_pose_deform(const nrslam_b200_options* opt, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                    const float* X_rest, const int32_t* point_vertex, const int8_t* vfs, nrslam_b200_graph* g,
                    float scale, float* pose_io, float* last_pos, float* deformation_out, float* X_out,
                    float* chi2_out, uint8_t* status_out, float* median_out, int32_t* lost_out,
                    int32_t* n_lost_out, nrslam_b200_stats* stats) {
  return orc_pose_deform_ex(opt, cam, n, uv, X_rest, point_vertex, vfs, g, scale, pose_io, last_pos,
                            deformation_out, X_out, chi2_out, status_out, median_out, lost_out, n_lost_out, stats,
                            0, nullptr);
}

// ---------------------------------------------------------------------------------------------
// g2o_optimization.cc:880-1161
// ---------------------------------------------------------------------------------------------
int orc_local_ba_ex(const nrslam_b200_options* opt, const nrslam_b200_camera* cam_, int32_t F, float* kf_pose_io,
                    int32_t O, const int32_t* obs_kf, const int32_t* obs_vertex, const float* uv, float* X_io,
                    const nrslam_b200_graph* g, float scale, int32_t iterations, nrslam_b200_stats* stats,
                    int debug_pcg, double* timing_out) {
  double t0 = wall_s();
  if (stats) *stats = nrslam_b200_stats{};
  if (F < 3) return NRSLAM_B200_NUM_TOO_FEW;  // :922-924
  if (iterations <= 0) iterations = opt->ba_iterations;
  Camera cam = to_cam(cam_);
  Optimizer optz(cam, false);
  optz.use_pcg = debug_pcg != 0;
  optz.pcg_tol = opt->pcg_rel_tol;
  optz.pcg_max_iter = opt->pcg_max_iterations;
  const int M = g->n_vertices;
  // pose vertices: ids = KF ids (ascending with age order), all below the point ids (:911,928)
  for (int k = 0; k < F; k++) {
    Vertex pv;
    pv.type = V_POSE;
    pv.dim = 6;
    pv.pose = se3_from_f7(kf_pose_io + 7 * k);
    optz.add_vertex(pv);
  }
  // inserted_landmarks[kf][mappoint] (:927-952)
  std::vector<int> inserted((size_t)F * M, -1);
  std::vector<int> kf_begin(F + 1, 0);
  for (int o = 0; o < O; o++) kf_begin[obs_kf[o] + 1]++;
  for (int k = 0; k < F; k++) kf_begin[k + 1] += kf_begin[k];
  for (int o = 0; o < O; o++) {
    Vertex v;
    for (int k = 0; k < 3; k++) v.x[k] = X_io[3 * o + k];
    int vid = optz.add_vertex(v);
    inserted[(size_t)obs_kf[o] * M + obs_vertex[o]] = vid;
  }
  const int regularizers_per_point = opt->regularizers_per_point;
  const float th2 = opt->th_huber_2dof_sq;
  const float th_huber_2dof = std::sqrt(th2);
  const float th3 = opt->th_huber_3dof_sq;
  const float th_huber_3dof = std::sqrt(th3);
  const float sigma_reprojection = opt->sigma_reprojection;
  const float info_reprojection = 1.0f / (sigma_reprojection * sigma_reprojection);
  const float sigma_position = opt->sigma_position;
  const float info_position = 1.0f / (sigma_position * sigma_position);
  const float sigma_spatial = (float)((double)opt->sigma_spatial_factor * scale);
  const float info_spatial = 1.0f / (sigma_spatial * sigma_spatial);

  std::vector<std::vector<int>> edge_cache(M);
  std::vector<char> cached(M, 0);
  std::unordered_set<uint64_t> spring_edges, dumper_edges;
  int n_spring = 0, n_damper = 0;
  double tb0 = wall_s();
  for (int k = 0; k < F; k++) {
    const int kn = (k + 1 < F) ? k + 1 : -1;  // next = next newer KF (:986-992)
    for (int o = kf_begin[k]; o < kf_begin[k + 1]; o++) {
      const int mp = obs_vertex[o];
      const int lidx = inserted[(size_t)k * M + mp];
      Edge e;
      e.type = E_REPROJ_BA;
      e.nv = 2;
      e.v[0] = k;
      e.v[1] = lidx;
      e.dim = 2;
      e.info = info_reprojection;
      e.delta = th_huber_2dof;
      e.meas[0] = uv[2 * o];
      e.meas[1] = uv[2 * o + 1];
      optz.add_edge(e);
      if (!cached[mp]) {
        edge_cache[mp] = get_edges(g, mp);
        cached[mp] = 1;
      }
      const std::vector<int>& redges = edge_cache[mp];
      int n_regularizers = 0;
      for (int p : redges) {
        const int other = g->col[p], ge = g->eid[p];
        if (n_regularizers > regularizers_per_point || g->status[ge] == NRSLAM_EDGE_BAD) break;
        const int oidx = inserted[(size_t)k * M + other];
        if (oidx < 0) continue;
        const uint64_t a = std::min(mp, other), b = std::max(mp, other);
        const uint64_t key = (a * (uint64_t)M + b) * (uint64_t)F + k;
        if (spring_edges.count(key)) {
          n_regularizers++;
          continue;
        }
        spring_edges.insert(key);
        Edge q;
        q.type = E_POSITION_BA;
        q.nv = 2;
        q.v[0] = lidx;
        q.v[1] = oidx;
        q.dim = 1;
        q.info = info_position;
        q.delta = -1;  // no robust kernel (:1057-1071)
        q.meas[0] = g->first_distance[ge];
        q.k = 1.1f;
        if (opt->spring_k != 1.1f) q.k = opt->spring_k;
        optz.add_edge(q);
        n_spring++;
        n_regularizers++;
      }
      if (kn >= 0) {
        const int nlidx = inserted[(size_t)kn * M + mp];
        if (nlidx < 0) continue;
        int n_reg2 = 0;
        for (int p : redges) {
          const int other = g->col[p], ge = g->eid[p];
          if (n_reg2 > regularizers_per_point || g->status[ge] == NRSLAM_EDGE_BAD) break;
          const int oidx = inserted[(size_t)k * M + other];
          const int noidx = inserted[(size_t)kn * M + other];
          if (oidx < 0 || noidx < 0) continue;
          const uint64_t a = std::min(mp, other), b = std::max(mp, other);
          const uint64_t key = (a * (uint64_t)M + b) * (uint64_t)F + k;  // (pair, k, k+1): k identifies it
          if (dumper_edges.count(key)) {
            n_reg2++;
            continue;
          }
          dumper_edges.insert(key);
          Edge s;
          s.type = E_DAMPER_BA;
          s.nv = 4;
          s.v[0] = lidx;
          s.v[1] = oidx;
          s.v[2] = nlidx;
          s.v[3] = noidx;
          s.dim = 3;
          s.info = info_spatial;
          s.delta = th_huber_3dof;
          s.weight = g->weight[ge];
          optz.add_edge(s);
          n_damper++;
          n_reg2++;
        }
      }
    }
  }
  double t_build = wall_s() - tb0;
  optz.initialize_optimization(0);
  optz.optimize(iterations);
  for (int k = 0; k < F; k++) se3_to_f7(optz.vertices[k].pose, kf_pose_io + 7 * k);
  for (int o = 0; o < O; o++)
    for (int k = 0; k < 3; k++) X_io[3 * o + k] = (float)optz.vertices[F + o].x[k];
  if (stats) {
    stats->n_reproj_edges = O;
    stats->n_spring_edges = n_spring;
    stats->n_damper_edges = n_damper;
    stats->n_points = O;
    stats->n_poses = F;
    stats->stage_ms = (float)(t_build * 1e3);
  }
  if (timing_out) {
    timing_out[0] = t_build;
    timing_out[1] = optz.stats.t_order;
    timing_out[2] = optz.stats.t_factor;
    timing_out[3] = optz.stats.t_build;
    timing_out[4] = (double)optz.pcg_iters_total;
  }
  fill_stats(stats, optz.stats, t0);
  return 0;
}

int orc_local_ba(const nrslam_b200_options* opt, const nrslam_b200_camera* cam, int32_t F, float* kf_pose_io,
                 int32_t O, const int32_t* obs_kf, const int32_t* obs_vertex, const float* uv, float* X_io,
                 const nrslam_b200_graph* g, float scale, int32_t iterations, nrslam_b200_stats* stats) {
  return orc_local_ba_ex(opt, cam, F, kf_pose_io, O, obs_kf, obs_vertex, uv, X_io, g, scale, iterations, stats, 0,
                         nullptr);
}

// ---------------------------------------------------------------------------------------------
// Small probes used by the oracle self-tests (tests/test_oracle_*.py)
// ---------------------------------------------------------------------------------------------
void orc_project(const nrslam_b200_camera* cam, const double* X, double* uv, double* J6) {
  Camera c = to_cam(cam);
  project_d(c, X, uv);
  if (J6) projection_jacobian_d(c, X, J6);
}
void orc_huber(double e, double delta, double* rho3) { huber(e, delta, rho3); }
void orc_se3_exp_mul(const double* u6, const double* q_t7, double* out7) {
  SE3 T;
  for (int i = 0; i < 4; i++) T.q[i] = q_t7[i];
  for (int i = 0; i < 3; i++) T.t[i] = q_t7[4 + i];
  SE3 r = se3_mul(se3_exp(u6), T);
  for (int i = 0; i < 4; i++) out7[i] = r.q[i];
  for (int i = 0; i < 3; i++) out7[4 + i] = r.t[i];
}

// Evaluate one edge: type, pose (q,t), up to 4 point estimates, parameters -> error, Jacobians.
// vals: [meas(3), Xw(3), rest1(3), rest2(3), weight, k]
void orc_edge_eval(const nrslam_b200_camera* cam, int type, const double* pose7, const double* pts12,
                   const double* vals, double* err3, double* J /*4 x 18*/) {
  Optimizer optz(to_cam(cam), false);
  Vertex pv;
  pv.type = V_POSE;
  pv.dim = 6;
  for (int i = 0; i < 4; i++) pv.pose.q[i] = pose7[i];
  for (int i = 0; i < 3; i++) pv.pose.t[i] = pose7[4 + i];
  optz.add_vertex(pv);
  for (int p = 0; p < 4; p++) {
    Vertex v;
    for (int k = 0; k < 3; k++) v.x[k] = pts12[3 * p + k];
    optz.add_vertex(v);
  }
  Edge e;
  e.type = type;
  for (int k = 0; k < 3; k++) {
    e.meas[k] = vals[k];
    e.Xw[k] = vals[3 + k];
    e.rest1[k] = vals[6 + k];
    e.rest2[k] = vals[9 + k];
  }
  e.weight = vals[12];
  e.k = vals[13];
  switch (type) {
    case E_REPROJ_ONLY_POSE: e.nv = 1; e.v[0] = 0; e.dim = 2; break;
    case E_REPROJ_DEFORM:
    case E_REPROJ_BA: e.nv = 2; e.v[0] = 0; e.v[1] = 1; e.dim = 2; break;
    case E_SPATIAL_DEFORM: e.nv = 2; e.v[0] = 1; e.v[1] = 2; e.dim = 3; break;
    case E_POSITION_DEFORM:
    case E_POSITION_BA: e.nv = 2; e.v[0] = 1; e.v[1] = 2; e.dim = 1; break;
    case E_SPATIAL_FIXED: e.nv = 1; e.v[0] = 1; e.ref_vertex = 2; e.dim = 3; break;
    case E_DAMPER_BA: e.nv = 4; e.v[0] = 1; e.v[1] = 2; e.v[2] = 3; e.v[3] = 4; e.dim = 3; break;
  }
  optz.compute_error(e);
  for (int k = 0; k < 3; k++) err3[k] = e.err[k];
  double Jt[4][18] = {{0}};
  optz.linearize(e, Jt);
  for (int a = 0; a < 4; a++)
    for (int k = 0; k < 18; k++) J[a * 18 + k] = Jt[a][k];
}

// Solve a block-sparse SPD system given as scalar triplets of the upper triangle — used to pin the
// sparse Cholesky against g2o's known-answer test (unit_test/solver/linear_solver_test.cpp:73-87).
int orc_sparse_solve(int n, int block, int nnz, const int* rows, const int* cols, const double* vals,
                     const double* b, double* x) {
  return orc::sparse_solve_triplets(n, block, nnz, rows, cols, vals, b, x);
}

}  // extern "C"
