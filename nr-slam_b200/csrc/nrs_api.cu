// nrs_api.cu — C ABI of libnrslam_b200 (include/nrslam_b200.h): context management and the host side of the
// three optimisation drivers. The host does what the reference does on the host as bookkeeping — neighbour
// selection from the regularisation graph, de-duplication, status / IQR gating, graph refresh — in the
// reference's order so that index bookkeeping is bit-exact; all arithmetic of the optimisation itself runs in
// the persistent sm_100a kernel of nrs_engine.cu. There is no CPU fallback.
//
// Reference (paths relative to /root/reference):
//   modules/optimization/g2o_optimization.cc:50-146    CameraPoseOptimization
//   modules/optimization/g2o_optimization.cc:148-557   CameraPoseAndDeformationOptimization
//   modules/optimization/g2o_optimization.cc:880-1161  LocalDeformableBundleAdjustment
//   modules/map/regularization_graph.cc:28-31,61-146   min weight, GetEdges, UpdateConnection, UpdateVertex
//   modules/utilities/geometry_toolbox.cc:26-28        InterpolationWeight
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <malloc.h>
#include <cstring>
#include <cstdio>
#include <numeric>
#include <set>
#include <thread>
#include <unordered_map>
#include <vector>

#include "nrs_host.h"

using namespace nrs;

namespace {

double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Host-side phase timer (NRSLAM_B200_HOSTPROF=1): where the end-to-end time of a call goes besides the kernel.
struct HostProf {
  bool on = getenv("NRSLAM_B200_HOSTPROF") != nullptr;
  double t = wall_ms();
  std::string line;
  void mark(const char* what) {
    if (!on) return;
    const double n = wall_ms();
    char buf[64];
    snprintf(buf, sizeof(buf), " %s %.3f", what, n - t);
    line += buf;
    t = n;
  }
  void print(const char* who) {
    if (on) fprintf(stderr, "[nrs host] %s:%s\n", who, line.c_str());
  }
};

// Host threads for the keyframe-parallel staging of large BA windows: min(items, cores, 16), capped by
// NRSLAM_B200_HOST_THREADS (one process per GPU shares the host: set it to cores / ranks there). 1 for small problems.
int host_threads(int items, int problem_size) {
  if (problem_size < 4096) return 1;
  static const int cap = [] {
    const char* e = getenv("NRSLAM_B200_HOST_THREADS");
    const int v = e ? atoi(e) : 16;
    return v < 1 ? 1 : v;
  }();
  const int hw = (int)std::thread::hardware_concurrency();
  return std::max(1, std::min(std::min(items, hw > 0 ? hw : 1), cap));
}

// fn(begin, end) over [0, N) on n_threads short-lived threads (BA-sized loops: a spawn costs ~30 us).
template <typename Fn>
void par_ranges(int n_threads, size_t N, Fn fn) {
  if (n_threads <= 1 || N < 8192) {
    fn((size_t)0, N);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back([=] { fn(N * t / n_threads, N * (t + 1) / n_threads); });
  fn((size_t)0, N / n_threads);
  for (auto& x : th) x.join();
}

// Host threads for the per-frame staging loops of the tracking path (a few thousand independent items of ~0.3 us):
// at most 4, none below 512 items; NRSLAM_B200_HOST_THREADS caps it like the BA staging.
int frame_threads(int items) {
  if (items < 512) return 1;
  static const int cap = [] {
    const char* e = getenv("NRSLAM_B200_HOST_THREADS");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  const int hw = (int)std::thread::hardware_concurrency();
  return std::max(1, std::min(cap, hw > 1 ? hw - 1 : 1));
}

// fn(t, n_threads) on n_threads threads (the caller is thread 0); returns when all are done.
template <typename Fn>
void run_threads(int n_threads, Fn fn) {
  if (n_threads <= 1) {
    fn(0, 1);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(n_threads - 1);
  for (int t = 1; t < n_threads; t++) pool.emplace_back(fn, t, n_threads);
  fn(0, n_threads);
  for (auto& th : pool) th.join();
}

int fail(nrslam_b200_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define NRS_CUDA(ctx, call)                                                                          \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(ctx, NRSLAM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// RegularizationGraph helpers
// ---------------------------------------------------------------------------------------------------
inline float interpolation_weight(float distance, float sigma) {  // geometry_toolbox.cc:26-28
  return std::exp(-(distance * distance) / (2 * sigma * sigma));
}
inline float graph_min_weight(const nrslam_b200_graph* g) {  // regularization_graph.cc:28-31
  return interpolation_weight((float)(g->weight_sigma * 1.5), g->weight_sigma);
}

// GetEdges (regularization_graph.cc:61-87): CSR entries of `vertex` ordered by (status asc, weight desc,
// neighbour asc — the documented tie-break for the reference's unstable std::sort), cut at the first
// weight < min_weight. Appends to `out`, returns the count.
int graph_sorted_entries(const nrslam_b200_graph* g, int vertex, float min_w, std::vector<int>& out) {
  const size_t base = out.size();
  for (int p = g->rowptr[vertex]; p < g->rowptr[vertex + 1]; p++) out.push_back(p);
  std::sort(out.begin() + base, out.end(), [g](int a, int b) {
    const int ea = g->eid[a], eb = g->eid[b];
    if (g->status[ea] != g->status[eb]) return g->status[ea] < g->status[eb];
    if (g->weight[ea] != g->weight[eb]) return g->weight[ea] > g->weight[eb];
    return g->col[a] < g->col[b];
  });
  size_t keep = base;
  while (keep < out.size() && !(g->weight[g->eid[out[keep]]] < min_w)) keep++;
  out.resize(keep);
  return (int)(keep - base);
}

int graph_update_vertex(nrslam_b200_graph* g, int vertex, const float* pos) {  // regularization_graph.cc:89-146
  int n_good = 0;
  const float* p1 = pos + 3 * (size_t)vertex;
  for (int p = g->rowptr[vertex]; p < g->rowptr[vertex + 1]; p++) {
    const float* p2 = pos + 3 * (size_t)g->col[p];
    const int e = g->eid[p];
    const float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
    const float distance = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (distance > g->max_distance[e]) g->max_distance[e] = distance;
    if (distance < g->min_distance[e]) g->min_distance[e] = distance;
    g->weight[e] = interpolation_weight(g->max_distance[e], g->weight_sigma);
    if (std::fabs((g->max_distance[e] - g->min_distance[e]) / g->min_distance[e]) > g->stretching_th)
      g->status[e] = NRSLAM_EDGE_BAD;
    else
      n_good++;
  }
  return n_good;
}

// ---------------------------------------------------------------------------------------------------
// Host description of one engine problem, then staging into HBM.
// ---------------------------------------------------------------------------------------------------
struct HostProblem {
  int F = 0, V = 0;
  bool poses_fixed = false, points_fixed = false;
  int spring_kind = SPRING_NONE;
  Cam cam;
  double info_reproj = 1, delta_reproj = -1, info_spatial = 1, delta_spatial = -1, info_spring = 1, delta_spring = -1,
         spring_k = 1.1f;
  float th2f = 5.99f, th3f = 0.584f;
  std::vector<double> pose_seed;        // 7F
  std::vector<double> x_seed, rest;     // 4V
  std::vector<double> uv;               // 2V
  std::vector<int> pt_kf;               // V
  std::vector<int> pair_i, pair_j;
  std::vector<double> pair_w, pair_d0;
  std::vector<int> dmp_v;               // 4D
  std::vector<double> dmp_w;
  std::vector<int> un_ptr;              // V+1 or empty
  std::vector<double> un_w;             // U
  std::vector<int> un_ref;              // U
  bool want_fixed_flags = false;        // reserve a per-vertex fixed-flag array (filled later)
  std::vector<unsigned char> fixed0;    // per-vertex setFixed(true) known at staging time (lost-point stage)
  std::vector<unsigned char> rp_level0, sp_level0;  // edge levels carried over from an earlier program (else 0)
  bool unary_on = false;                // the unary (fixed-reference) edges take part
  bool plain_jacobi = false;            // 3x3 block-Jacobi instead of the dense 16-row blocks (tiny, anchored problems)
  int n_sort = 0;                       // leading rows of every pose-slot group that may be re-ordered spatially
  int n_stage1 = 0;                     // rows of the first launch (tracking: without the lost-point rows); 0 = all
  std::vector<int> kf_begin;            // F+1: rows of pose slot k are [kf_begin[k], kf_begin[k+1]) (rows w/o pose: slot 0)
  int n_halo = 0;                       // landmark-sharded BA: trailing rows owned by other ranks (read-only copies)
  bool sharded = false;
  bool direct = false;                  // tracking: exact multifrontal L D L^T engine (nrs_direct.cu)
  const std::vector<int32_t>* plan_key = nullptr;  // identity of the rows (map points in frame order): enables plan re-use
  const double* pose_seed_dev = nullptr;  // seed poses already on the device (result of an earlier launch on the ctx stream)
  int n_unknown = 0;                    // direct engine: rows [n_unknown, V) are fixed vertices (0: every row is an unknown)
  std::vector<int> ops, op_args;
};

// Copies into the pinned input arena can be deferred and run by the staging threads in one go (stage_problem): the
// arena fill of a tracking frame is ~1 MB in two dozen pieces, 0.3 ms on one thread.
struct CopyJob {
  void* dst;
  const void* src;
  size_t bytes;
};
thread_local std::vector<CopyJob>* t_defer = nullptr;

template <typename T>
size_t put(Arena& a, const std::vector<T>& v, size_t min_elems = 1) {
  const size_t n = std::max(v.size(), min_elems);
  const size_t off = a.take<T>(n);
  if (!v.empty()) {
    const size_t bytes = v.size() * sizeof(T);
    if (t_defer && bytes >= 4096)
      t_defer->push_back(CopyJob{a.h<T>(off), v.data(), bytes});  // the source outlives flush_copies()
    else
      memcpy(a.h<T>(off), v.data(), bytes);
  }
  return off;
}

// Runs the deferred copies on the pool in 32 KB pieces (dynamic distribution).
void flush_copies(HostPool& pool, std::vector<CopyJob>& jobs, int spawn_threads = 1) {
  t_defer = nullptr;
  if (jobs.empty()) return;
  if (pool.threads() == 1 && spawn_threads > 1) {  // BA window: no persistent pool, megabytes to copy
    size_t total = 0;
    for (const CopyJob& j : jobs) total += j.bytes;
    std::vector<size_t> acc(jobs.size() + 1, 0);
    for (size_t k = 0; k < jobs.size(); k++) acc[k + 1] = acc[k] + jobs[k].bytes;
    par_ranges(spawn_threads, total, [&](size_t b, size_t e) {
      for (size_t k = 0; k < jobs.size(); k++) {
        const size_t lo = std::max(b, acc[k]), hi = std::min(e, acc[k + 1]);
        if (lo < hi)
          memcpy(static_cast<char*>(jobs[k].dst) + (lo - acc[k]), static_cast<const char*>(jobs[k].src) + (lo - acc[k]), hi - lo);
      }
    });
    jobs.clear();
    return;
  }
  constexpr size_t kPiece = 32 * 1024;
  std::vector<CopyJob> pieces;
  for (const CopyJob& j : jobs)
    for (size_t o = 0; o < j.bytes; o += kPiece)
      pieces.push_back(CopyJob{static_cast<char*>(j.dst) + o, static_cast<const char*>(j.src) + o, std::min(kPiece, j.bytes - o)});
  std::atomic<size_t> next{0};
  pool.run([&](int, int) {
    for (size_t k = next.fetch_add(1, std::memory_order_relaxed); k < pieces.size();
         k = next.fetch_add(1, std::memory_order_relaxed))
      memcpy(pieces[k].dst, pieces[k].src, pieces[k].bytes);
  });
  jobs.clear();
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// 30-bit Morton code of a point inside the bounding box [lo, lo + ext]
inline uint32_t morton3(const double* p, const double* lo, const double* inv_ext) {
  uint32_t code = 0;
  uint32_t q[3];
  for (int a = 0; a < 3; a++) {
    double t = (p[a] - lo[a]) * inv_ext[a];
    t = std::min(std::max(t, 0.0), 1.0);
    q[a] = (uint32_t)(t * 1023.0);
  }
  for (int b = 9; b >= 0; b--)
    for (int a = 0; a < 3; a++) code = (code << 1) | ((q[a] >> b) & 1u);
  return code;
}

// Re-order the first hp.n_sort rows of every pose-slot group along a Morton curve of their positions and rewrite
// every row index of the problem. row_of[old] = new.
void sort_rows(HostProblem& hp, std::vector<int>& row_of, const std::vector<int>* forced_old_of_new = nullptr) {
  const int V = hp.V;
  row_of.resize(V);
  std::iota(row_of.begin(), row_of.end(), 0);
  if (!forced_old_of_new && (hp.n_sort <= 1 || hp.points_fixed)) return;
  std::vector<int> old_of_new(V);
  std::iota(old_of_new.begin(), old_of_new.end(), 0);
  bool rows_done = false;  // the per-keyframe sort below permutes the row arrays itself
  if (forced_old_of_new) {
    // elimination order of the exact solve (nrs_direct_plan.h); it may cover a prefix (the unknown rows) only
    std::copy(forced_old_of_new->begin(), forced_old_of_new->end(), old_of_new.begin());
  } else {
  // keyframes are sorted independently: host threads take them round robin on large windows (same result)
  const int n_threads = host_threads(hp.F, V);
  auto work = [&](int tix) {
  std::vector<std::pair<uint32_t, int>> keyed;
  std::vector<double> tmp;
  std::vector<int> tmpi;
  for (int k = tix; k < hp.F; k += n_threads) {
    const int b = hp.kf_begin[k], e = std::min(hp.kf_begin[k + 1], hp.n_sort);
    if (e - b < 2) continue;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    auto pos = [&](int i, int a) { return hp.rest[4 * (size_t)i + a] + hp.x_seed[4 * (size_t)i + a]; };
    for (int i = b; i < e; i++)
      for (int a = 0; a < 3; a++) {
        lo[a] = std::min(lo[a], pos(i, a));
        hi[a] = std::max(hi[a], pos(i, a));
      }
    // one common scale for the three axes keeps the curve isotropic
    double ext = 0;
    for (int a = 0; a < 3; a++) ext = std::max(ext, hi[a] - lo[a]);
    const double inv = ext > 0 ? 1.0 / ext : 0.0;
    const double inv_ext[3] = {inv, inv, inv};
    keyed.clear();
    for (int i = b; i < e; i++) {
      const double p[3] = {pos(i, 0), pos(i, 1), pos(i, 2)};
      keyed.emplace_back(morton3(p, lo, inv_ext), i);
    }
    std::sort(keyed.begin(), keyed.end());
    for (int t = 0; t < e - b; t++) old_of_new[b + t] = keyed[t].second;
    // rows only move inside their keyframe: permute them here, through a temporary that stays in cache (the
    // whole-array gathers cost three extra passes over 16 MB and six thread spawns on a 200k-row window)
    const int nk = e - b;
    tmp.resize(4 * (size_t)nk);
    auto gather = [&](std::vector<double>& v, int stride) {
      for (int t = 0; t < nk; t++)
        for (int a = 0; a < stride; a++) tmp[(size_t)stride * t + a] = v[(size_t)stride * keyed[t].second + a];
      std::copy(tmp.begin(), tmp.begin() + (size_t)stride * nk, v.begin() + (size_t)stride * b);
    };
    gather(hp.x_seed, 4);
    gather(hp.rest, 4);
    gather(hp.uv, 2);
    tmpi.resize(nk);
    for (int t = 0; t < nk; t++) tmpi[t] = hp.pt_kf[keyed[t].second];
    std::copy(tmpi.begin(), tmpi.end(), hp.pt_kf.begin() + b);
    for (int t = 0; t < nk; t++) row_of[keyed[t].second] = b + t;
  }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  rows_done = true;
  }
  const int pt = host_threads(64, V);  // the gathers below are independent per row / per edge
  if (!rows_done) {
  par_ranges(pt, (size_t)V, [&](size_t b, size_t e) {
    for (size_t nw = b; nw < e; nw++) row_of[old_of_new[nw]] = (int)nw;
  });
  auto permute = [&](std::vector<double>& v, int stride) {
    std::vector<double> o(v.size());
    par_ranges(pt, (size_t)V, [&](size_t b, size_t e) {
      for (size_t nw = b; nw < e; nw++)
        for (int a = 0; a < stride; a++) o[(size_t)stride * nw + a] = v[(size_t)stride * old_of_new[nw] + a];
    });
    v.swap(o);
  };
  permute(hp.x_seed, 4);
  permute(hp.rest, 4);
  permute(hp.uv, 2);
  {
    std::vector<int> o(V);
    for (int nw = 0; nw < V; nw++) o[nw] = hp.pt_kf[old_of_new[nw]];
    hp.pt_kf.swap(o);
  }
  }
  par_ranges(pt, hp.pair_i.size(), [&](size_t b, size_t e) {
    for (size_t t = b; t < e; t++) {
      hp.pair_i[t] = row_of[hp.pair_i[t]];
      hp.pair_j[t] = row_of[hp.pair_j[t]];
    }
  });
  par_ranges(pt, hp.dmp_v.size(), [&](size_t b, size_t e) {
    for (size_t t = b; t < e; t++) hp.dmp_v[t] = row_of[hp.dmp_v[t]];
  });
  if (forced_old_of_new) {  // per-row extras of the lost-point problem
    auto permute_u8 = [&](std::vector<unsigned char>& v) {
      if (v.empty()) return;
      std::vector<unsigned char> o(V);
      for (int nw = 0; nw < V; nw++) o[nw] = v[old_of_new[nw]];
      v.swap(o);
    };
    permute_u8(hp.fixed0);
    permute_u8(hp.rp_level0);
    if (!hp.un_ptr.empty()) {
      std::vector<int> np(V + 1, 0), nr;
      std::vector<double> nwt;
      nr.reserve(hp.un_ref.size());
      nwt.reserve(hp.un_w.size());
      for (int nw = 0; nw < V; nw++) {
        const int o = old_of_new[nw];
        for (int a = hp.un_ptr[o]; a < hp.un_ptr[o + 1]; a++) {
          nr.push_back(row_of[hp.un_ref[a]]);
          nwt.push_back(hp.un_w[a]);
        }
        np[nw + 1] = (int)nr.size();
      }
      hp.un_ptr.swap(np);
      hp.un_ref.swap(nr);
      hp.un_w.swap(nwt);
    }
  }
}

// Launch plan of one engine launch over the rows [0, kf_begin[F]).
struct Plan {
  int V = 0, n_chunks = 0, grid = 0, block = 0, cluster_mode = 0, resident = 0, res_rows = 0, res_inc = 0,
      block_prec = 0, wide = 0;
  size_t smem = 0;
  std::vector<int> chunk_kf, chunk_begin, chunk_end, kf_chunk_ptr;
  const int *d_chunk_kf = nullptr, *d_chunk_begin = nullptr, *d_chunk_end = nullptr, *d_kf_chunk_ptr = nullptr;
  // wide CG loop: one contiguous row range per CTA, cut into segments at the pose-slot boundaries
  std::vector<int> wseg_ptr, wseg_begin, wseg_end, wseg_kf, kf_wseg_ptr;
  const int *d_wseg_ptr = nullptr, *d_wseg_begin = nullptr, *d_wseg_end = nullptr, *d_wseg_kf = nullptr,
            *d_kf_wseg_ptr = nullptr;
  // halo push lists of the cluster-native CG loop (built when the plan qualifies for it)
  int halo_rows = 0, coarse = 0;
  std::vector<int> inc_halo, push_ptr, push_row, push_dst, xinc_ptr, xinc_idx;
  const int *d_inc_halo = nullptr, *d_push_ptr = nullptr, *d_push_row = nullptr, *d_push_dst = nullptr;
  const int *d_xinc_ptr = nullptr, *d_xinc_idx = nullptr;
};

// For every chunk: the out-of-chunk rows its incidences read (its halo) and, for every owner chunk, which rows to push
// where. Deterministic (sorted) so that runs are reproducible.
void build_halo(Plan& pl, const std::vector<int>& inc_ptr, const std::vector<int>& inc_other) {
  const int nc = pl.n_chunks;
  pl.inc_halo.assign(inc_other.size(), -1);
  std::vector<int> chunk_of(pl.V, 0);
  for (int c = 0; c < nc; c++)
    for (int i = pl.chunk_begin[c]; i < pl.chunk_end[c]; i++) chunk_of[i] = c;
  std::vector<std::vector<std::pair<int, int>>> pushes(nc);  // owner -> (row, target * 65536 + slot)
  pl.halo_rows = 0;
  for (int c = 0; c < nc; c++) {
    std::vector<int> rows;
    for (int a = inc_ptr[pl.chunk_begin[c]]; a < inc_ptr[pl.chunk_end[c]]; a++) {
      const int o = inc_other[a];
      if (o < pl.chunk_begin[c] || o >= pl.chunk_end[c]) rows.push_back(o);
    }
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    pl.halo_rows = std::max(pl.halo_rows, (int)rows.size());
    for (int a = inc_ptr[pl.chunk_begin[c]]; a < inc_ptr[pl.chunk_end[c]]; a++) {
      const int o = inc_other[a];
      if (o < pl.chunk_begin[c] || o >= pl.chunk_end[c])
        pl.inc_halo[a] = (int)(std::lower_bound(rows.begin(), rows.end(), o) - rows.begin());
    }
    for (size_t k = 0; k < rows.size(); k++) pushes[chunk_of[rows[k]]].emplace_back(rows[k], c * 65536 + (int)k);
  }
  pl.xinc_ptr.assign(nc + 1, 0);
  pl.xinc_idx.clear();
  for (int c = 0; c < nc; c++) {
    for (int a = inc_ptr[pl.chunk_begin[c]]; a < inc_ptr[pl.chunk_end[c]]; a++)
      if (pl.inc_halo[a] >= 0) pl.xinc_idx.push_back(a);
    pl.xinc_ptr[c + 1] = (int)pl.xinc_idx.size();
  }
  pl.push_ptr.assign(nc + 1, 0);
  for (int c = 0; c < nc; c++) {
    std::sort(pushes[c].begin(), pushes[c].end());
    pl.push_ptr[c + 1] = pl.push_ptr[c] + (int)pushes[c].size();
    for (auto& pr : pushes[c]) {
      pl.push_row.push_back(pr.first);
      pl.push_dst.push_back(pr.second);
    }
  }
}

// The path is bound by synchronisation latency: prefer ONE thread-block cluster (hardware barrier) whenever the
// rows fit a cluster's threads; otherwise a cooperative grid with the atomics barrier.
int make_plan(nrslam_b200_ctx* ctx, const HostProblem& hp, const std::vector<int>& kf_begin,
              const std::vector<int>& inc_ptr, Plan& pl, const std::vector<int>* dinc_ptr = nullptr) {
  const int F = hp.F;
  pl.V = kf_begin[F];
  if (ctx->max_cluster < 0) ctx->max_cluster = engine_max_cluster(kMaxBlock, 200 * 1024);
  int ncl = env_int("NRSLAM_B200_CLUSTER", ctx->max_cluster);
  ncl = std::min(ncl, ctx->max_cluster);
  auto count_chunks = [&](int rows) {
    int n = 0;
    for (int k = 0; k < F; k++) n += (kf_begin[k + 1] - kf_begin[k] + rows - 1) / rows;
    return n;
  };
  int chunk_rows = kMaxRows, cluster_mode = 0;
  if (ncl >= 2 && !hp.sharded) {  // a sharded rank synchronises with its peers through the cooperative-grid path
    for (int rows = 32; rows <= kMaxRows; rows += 8) {
      if (count_chunks(rows) <= ncl) {
        chunk_rows = rows;
        cluster_mode = 1;
        break;
      }
    }
  }
  // experiment (off): pick the chunk size that minimises passes x chunk rows of the busiest CTA. Measured on
  // configs[2]: no gain — a chunk pass is a fixed chain of dependent gathers whatever its row count, so time follows
  // the number of passes, and smaller chunks add pose-partial records to sum (96-row chunks 67.0 vs 66.1 Mcycles)
  if (!cluster_mode && !hp.points_fixed && env_int("NRSLAM_B200_BALANCE", 0)) {
    const size_t smem0 = engine_smem_bytes(F, 0, 0, 0);
    const int g1 = engine_max_grid(kMaxBlock, smem0, 0), g2 = engine_max_grid(kMaxBlock, smem0, 1);
    if (count_chunks(kMaxRows) > std::max(std::max(g1, g2), 1)) {
      const bool want_wide = env_int("NRSLAM_B200_WIDE", 1) != 0;
      long best = -1;
      for (int rows = env_int("NRSLAM_B200_BALANCE_MIN", 88); rows <= kMaxRows; rows += 8) {
        const int nc = count_chunks(rows);
        const int blk = std::min(kMaxBlock, (kTPR * rows + 31) / 32 * 32);
        const int G = std::max(1, std::max(engine_max_grid(blk, smem0, 0), want_wide ? engine_max_grid(blk, smem0, 1) : 0));
        const long passes = (nc + G - 1) / G;
        const long cost = passes * (rows + 6);  // + a per-pass overhead in row units
        if (best < 0 || cost < best) {
          best = cost;
          chunk_rows = rows;
        }
      }
    }
  }
  // kTPR threads per row inside the CG loop; pose-only problems (no CG) run one thread per row
  int block = hp.points_fixed ? std::max((chunk_rows + 31) / 32 * 32, 128) : kTPR * chunk_rows;
  block = std::min(kMaxBlock, (block + 31) / 32 * 32);
  // chunks: contiguous rows of one pose slot, at most chunk_rows rows each
  pl.kf_chunk_ptr.assign(F + 1, 0);
  for (int k = 0; k < F; k++) {
    pl.kf_chunk_ptr[k] = (int)pl.chunk_kf.size();
    for (int b = kf_begin[k]; b < kf_begin[k + 1]; b += chunk_rows) {
      pl.chunk_kf.push_back(k);
      pl.chunk_begin.push_back(b);
      pl.chunk_end.push_back(std::min(b + chunk_rows, kf_begin[k + 1]));
    }
  }
  pl.kf_chunk_ptr[F] = (int)pl.chunk_kf.size();
  const int n_chunks = (int)pl.chunk_kf.size();
  // resident mode: one chunk per CTA and the chunk state fits shared memory
  int res_rows = 0, res_inc = 0;
  for (int c = 0; c < n_chunks; c++) {
    res_rows = std::max(res_rows, pl.chunk_end[c] - pl.chunk_begin[c]);
    res_inc = std::max(res_inc, inc_ptr[pl.chunk_end[c]] - inc_ptr[pl.chunk_begin[c]]);
  }
  res_rows = (res_rows + kPrecBlock - 1) / kPrecBlock * kPrecBlock;
  res_inc = std::max(res_inc, 1);
  int resident = 0, block_prec = 0, grid = 0, wide = 0;
  size_t smem = engine_smem_bytes(F, 0, 0, 0);
  if (smem > 200 * 1024) return fail(ctx, NRSLAM_B200_ERR_ARG, "too many poses for the shared-memory pose blocks");
  const size_t kSmemBudget = 220 * 1024;
  if (!hp.points_fixed && !hp.sharded && env_int("NRSLAM_B200_RESIDENT", 1)) {
    const size_t s1 = engine_smem_bytes(F, res_rows, res_inc, 1), s0 = engine_smem_bytes(F, res_rows, res_inc, 0);
    const bool want_prec = env_int("NRSLAM_B200_BLOCKPREC", 1) != 0 && !hp.plain_jacobi;
    const size_t s = (want_prec && s1 <= kSmemBudget) ? s1 : s0;
    if (s <= kSmemBudget) {
      const int mg = cluster_mode ? n_chunks : engine_max_grid(block, s);
      if (n_chunks <= mg) {
        resident = 1;
        block_prec = (s == s1 && want_prec) ? 1 : 0;
        smem = s;
        grid = n_chunks;
      }
    }
  }
  if (!resident) {
    if (cluster_mode) {
      grid = n_chunks;
    } else {
      int max_grid = engine_max_grid(block, smem);
      if (max_grid <= 0) return fail(ctx, NRSLAM_B200_ERR_CUDA, "kernel cannot be made resident");
      // more chunks than one CTA per SM can take in one pass: the two-CTAs-per-SM variant hides the gather latency
      if (!hp.points_fixed && n_chunks > max_grid && env_int("NRSLAM_B200_WIDE", 1)) {
        const size_t smem_w = engine_smem_bytes_wide(F);
        const int mg2 = engine_max_grid(block, smem_w, 1);
        if (mg2 > max_grid) {
          wide = 1;
          max_grid = mg2;
          smem = smem_w;
        }
      }
      grid = std::max(1, std::min(n_chunks, max_grid));
      if (ctx->opt.grid_ctas > 0) grid = std::min(grid, ctx->opt.grid_ctas);
      const int g = env_int("NRSLAM_B200_GRID", 0);
      if (g > 0) grid = std::min(max_grid, g);
    }
  }
  if (cluster_mode) {
    // the cluster must be schedulable with this shared-memory size; otherwise fall back to the cooperative grid
    const int mc = engine_max_cluster(block, smem);
    if (mc < grid) {
      cluster_mode = 0;
      const int max_grid = engine_max_grid(block, smem);
      if (max_grid < grid) return fail(ctx, NRSLAM_B200_ERR_CUDA, "kernel cannot be made resident");
    }
  }
  pl.n_chunks = n_chunks; pl.grid = grid; pl.block = block; pl.cluster_mode = cluster_mode; pl.resident = resident;
  pl.res_rows = res_rows; pl.res_inc = res_inc; pl.block_prec = block_prec; pl.smem = smem; pl.wide = wide;
  if (wide && !cluster_mode) {
    // CTA b owns rows [V b / G, V (b + 1) / G); a segment = the part of that range inside one pose slot. Rows are
    // slot-major, so the segments of a slot are contiguous in this numbering.
    // Ranges are balanced by gather work (1 per row + 1 per pair incidence + 2 per damper incidence), not by rows.
    const int Vr = pl.V;
    std::vector<long long> wsum(Vr + 1, 0);
    for (int i = 0; i < Vr; i++)
      wsum[i + 1] = wsum[i] + 2 + (inc_ptr[i + 1] - inc_ptr[i]) + 2LL * (dinc_ptr ? (*dinc_ptr)[i + 1] - (*dinc_ptr)[i] : 0);
    std::vector<int> cut(grid + 1, 0);
    for (int b = 1; b < grid; b++)
      cut[b] = (int)(std::lower_bound(wsum.begin(), wsum.end(), wsum[Vr] * b / grid) - wsum.begin());
    cut[grid] = Vr;
    for (int b = 1; b <= grid; b++) cut[b] = std::max(cut[b], cut[b - 1]);
    pl.wseg_ptr.assign(grid + 1, 0);
    pl.kf_wseg_ptr.assign(F + 1, 0);
    int k = 0;
    for (int b = 0; b < grid; b++) {
      int r = cut[b];
      const int r1 = cut[b + 1];
      while (r < r1) {
        while (k < F && kf_begin[k + 1] <= r) k++;
        const int e = std::min(r1, kf_begin[std::min(k, F - 1) + 1]);
        pl.wseg_begin.push_back(r);
        pl.wseg_end.push_back(e);
        pl.wseg_kf.push_back(std::min(k, F - 1));
        pl.kf_wseg_ptr[std::min(k, F - 1) + 1]++;
        r = e;
      }
      pl.wseg_ptr[b + 1] = (int)pl.wseg_begin.size();
    }
    for (int q = 0; q < F; q++) pl.kf_wseg_ptr[q + 1] += pl.kf_wseg_ptr[q];
  }
  return 0;
}

void apply_plan(Params& p, const Plan& pl) {
  p.V = pl.V; p.n_chunks = pl.n_chunks;
  p.cluster_mode = pl.cluster_mode; p.resident = pl.resident; p.res_rows = pl.res_rows; p.res_inc = pl.res_inc;
  p.block_prec = pl.block_prec; p.wide = pl.wide;
  p.chunk_kf = pl.d_chunk_kf; p.chunk_begin = pl.d_chunk_begin; p.chunk_end = pl.d_chunk_end;
  p.kf_chunk_ptr = pl.d_kf_chunk_ptr;
  p.n_wseg = (int)pl.wseg_begin.size();
  p.wide_prefetch = env_int("NRSLAM_B200_WIDE_PREFETCH", 0);  // measured: configs[3] -1.8 %, configs[2] +2 %
  p.wseg_ptr = pl.d_wseg_ptr; p.wseg_begin = pl.d_wseg_begin; p.wseg_end = pl.d_wseg_end; p.wseg_kf = pl.d_wseg_kf;
  p.kf_wseg_ptr = pl.d_kf_wseg_ptr;
  p.coarse = pl.coarse;
  p.halo_rows = pl.halo_rows; p.inc_halo = pl.d_inc_halo; p.push_ptr = pl.d_push_ptr; p.push_row = pl.d_push_row;
  p.push_dst = pl.d_push_dst;
  p.xinc_ptr = pl.d_xinc_ptr; p.xinc_idx = pl.d_xinc_idx;
}

int stage_problem(nrslam_b200_ctx* ctx, Staged& st, HostProblem& hp) {
  const int F = hp.F, V = hp.V, P = (int)hp.pair_i.size(), D = (int)hp.dmp_w.size();
  const int U = (int)hp.un_w.size();
  st.valid = false;
  st.use_direct = false;
  HostProf hprof;
  // ---- exact-solve engine: symbolic analysis first, its elimination order becomes the row order
  DirectPlanHost dplan_local;
  const bool cacheable = hp.plan_key != nullptr && env_int("NRSLAM_B200_PLAN_CACHE", 1) != 0;
  DirectPlanHost& dplan = cacheable ? ctx->plan_cache.plan : dplan_local;
  st.plan_reused = false;
  const int Vu = (hp.n_unknown > 0 && hp.n_unknown < V) ? hp.n_unknown : V;  // rows [Vu, V) are fixed vertices
  bool direct = hp.direct && F == 1 && D == 0 && !hp.points_fixed && !hp.sharded && hp.n_stage1 == 0 && Vu >= 1 &&
                (hp.poses_fixed || U == 0);
  if (direct && !hp.fixed0.empty())
    for (int r = 0; r < V && direct; r++) direct = (hp.fixed0[r] != 0) == (r >= Vu);
  if (direct && hp.fixed0.empty() && Vu < V) direct = false;
  size_t dsmem = 0;
  int dscratch = 0, dmaxnv = 0;
  if (direct) {
    const int depth = std::min(direct_depth(Vu, ctx->sm_count, hp.poses_fixed ? env_int("NRSLAM_B200_LOST_LEAF", 8) : 8),
                               env_int("NRSLAM_B200_DIRECT_DEPTH", 7));
    bool reuse = false;
    if (cacheable && Vu == V) {
      PlanCache& pc = ctx->plan_cache;
      if (pc.valid && pc.plan.V == V && pc.depth == depth && pc.np == (hp.poses_fixed ? 0 : 2) && pc.key == *hp.plan_key) {
        if (pc.pair_i == hp.pair_i && pc.pair_j == hp.pair_j) {
          reuse = true;  // the very same structure
        } else {         // every pair inside the adjacency the plan was built from?
          reuse = true;
          for (int e = 0; e < P && reuse; e++) {
            const uint64_t a = (uint64_t)std::min(hp.pair_i[e], hp.pair_j[e]), b = (uint64_t)std::max(hp.pair_i[e], hp.pair_j[e]);
            reuse = std::binary_search(pc.pairs.begin(), pc.pairs.end(), (a << 32) | b);
          }
        }
      }
      if (reuse) {
        pc.reuses++;
      } else {
        build_direct_plan(V, hp.uv.data(), hp.pair_i, hp.pair_j, depth, dplan, !hp.poses_fixed);
        pc.valid = true;
        pc.depth = depth;
        pc.np = hp.poses_fixed ? 0 : 2;
        pc.key = *hp.plan_key;
        pc.pair_i = hp.pair_i;
        pc.pair_j = hp.pair_j;
        pc.pairs.resize(P);
        for (int e = 0; e < P; e++) {
          const uint64_t a = (uint64_t)std::min(hp.pair_i[e], hp.pair_j[e]), b = (uint64_t)std::max(hp.pair_i[e], hp.pair_j[e]);
          pc.pairs[e] = (a << 32) | b;
        }
        std::sort(pc.pairs.begin(), pc.pairs.end());
        pc.builds++;
      }
      st.plan_reused = reuse;
    } else if (Vu == V) {
      build_direct_plan(V, hp.uv.data(), hp.pair_i, hp.pair_j, depth, dplan, !hp.poses_fixed);
    } else {  // only pairs between two unknowns couple unknowns
      std::vector<int> pi, pj;
      for (int e = 0; e < P; e++)
        if (hp.pair_i[e] < Vu && hp.pair_j[e] < Vu) {
          pi.push_back(hp.pair_i[e]);
          pj.push_back(hp.pair_j[e]);
        }
      build_direct_plan(Vu, hp.uv.data(), pi, pj, depth, dplan, !hp.poses_fixed);
    }
    int max_ns = 0;
    for (int t = 1; t <= dplan.n_nodes; t++) max_ns = std::max(max_ns, 3 * dplan.nv[t]);
    dscratch = max_ns * (1 + direct_block_threads() / 32);
    dmaxnv = max_ns / 3;
    dsmem = direct_smem_bytes(dplan.max_path, dscratch, dmaxnv, dplan.max_rows, dplan.smem_doubles);
    // the busiest team member's panel must fit one SM, the whole grid must be co-resident, one row group per thread
    if (dsmem > 226 * 1024 || direct_max_grid(dsmem) < dplan.G ||
        (Vu + dplan.G - 1) / dplan.G > direct_block_threads())
      direct = false;
    if (getenv("NRSLAM_B200_DIRECT_DEBUG")) {
      fprintf(stderr, "[nrs direct] V %d depth %d G %d fits %d: smem %zu B (panel %zu doubles, scratch %d, max_nv %d, "
              "max_rows %d, max_path %d) separators:", Vu, depth, dplan.G, direct ? 1 : 0, dsmem, (size_t)dplan.smem_doubles,
              dscratch, dmaxnv, dplan.max_rows, dplan.max_path);
      for (int t = 1; t <= dplan.n_nodes && t < 16; t++) fprintf(stderr, " %d", dplan.nv[t]);
      fprintf(stderr, " | boundaries:");
      for (int t = 1; t <= dplan.n_nodes && t < 16; t++) fprintf(stderr, " %d", dplan.nbv[t]);
      fprintf(stderr, "\n");
    }
  }
  hprof.mark("direct_plan");
  sort_rows(hp, st.row_of, direct ? &dplan.old_of_new : nullptr);
  hprof.mark("sort_rows");

  std::vector<int> dinc_ptr(V + 1, 0), dinc_ent(4 * (size_t)D);
  auto damper_csr = [&] {
    for (size_t t = 0; t < 4 * (size_t)D; t++) dinc_ptr[hp.dmp_v[t] + 1]++;
    for (int i = 0; i < V; i++) dinc_ptr[i + 1] += dinc_ptr[i];
    std::vector<int> w(dinc_ptr.begin(), dinc_ptr.end() - 1);
    for (size_t t = 0; t < 4 * (size_t)D; t++) dinc_ent[w[hp.dmp_v[t]]++] = (int)t;  // id*4 + role
  };
  std::thread damper_thread;  // the damper lists do not depend on the pair lists: second host thread on BA windows
  if (D >= 4096 && host_threads(2, V) > 1) damper_thread = std::thread(damper_csr);
  // incidence lists
  std::vector<int> inc_ptr(V + 1, 0), inc_other(2 * (size_t)P), inc_ent(2 * (size_t)P), inc_row(2 * (size_t)P);
  for (int e = 0; e < P; e++) {
    inc_ptr[hp.pair_i[e] + 1]++;
    inc_ptr[hp.pair_j[e] + 1]++;
  }
  for (int i = 0; i < V; i++) inc_ptr[i + 1] += inc_ptr[i];
  {
    std::vector<int> w(inc_ptr.begin(), inc_ptr.end() - 1);
    for (int e = 0; e < P; e++) {
      int a = w[hp.pair_i[e]]++;
      inc_other[a] = hp.pair_j[e];
      inc_ent[a] = 2 * e;
      inc_row[a] = hp.pair_i[e];
      a = w[hp.pair_j[e]]++;
      inc_other[a] = hp.pair_i[e];
      inc_ent[a] = 2 * e + 1;
      inc_row[a] = hp.pair_j[e];
    }
  }
  if (damper_thread.joinable()) damper_thread.join(); else damper_csr();

  hprof.mark("incidences");
  // ---- launch plans. Tracking stages its lost-point rows behind the optimised ones: the main rounds run on the
  // first hp.n_stage1 rows only (plan A), the lost-point stage on all rows (plan B).
  const int V1 = hp.sharded ? V - hp.n_halo : (hp.n_stage1 > 0 && hp.n_stage1 < V && F == 1) ? hp.n_stage1 : V;
  Plan planA, planB;
  {
    std::vector<int> kfb(hp.kf_begin);
    kfb[F] = V1;
    const int rc = make_plan(ctx, hp, kfb, inc_ptr, planA, &dinc_ptr);
    if (rc) return rc;
    if (V1 < V && !hp.sharded) {
      const int rc2 = make_plan(ctx, hp, hp.kf_begin, inc_ptr, planB, &dinc_ptr);
      if (rc2) return rc2;
    }
  }
  hprof.mark("make_plan");
  const int n_chunks = planA.n_chunks;
  const int max_chunks = std::max(planA.n_chunks, planB.n_chunks), max_grid_used = std::max(planA.grid, planB.grid);
  for (Plan* pl : {&planA, &planB}) {
    // the cluster-native CG loop (nrs_engine.cu: pcg_cluster) exchanges z through pushed halos
    if (pl->n_chunks == 0 || !pl->cluster_mode || !pl->resident || F != 1 || D != 0 || hp.points_fixed) continue;
    if (!env_int("NRSLAM_B200_HALO_PUSH", 1) || pl->n_chunks >= 65536) continue;
    if (direct && pl == &planA) continue;  // the exact-solve engine runs this plan: no CG loop, no halos
    build_halo(*pl, inc_ptr, inc_other);
    // the coarse level needs the dense block preconditioner's layout and at most 16 aggregates
    int coarse = env_int("NRSLAM_B200_COARSE", 1) && pl->block_prec && pl->n_chunks <= 16 && pl->halo_rows < 255 * 256;
    size_t extra = engine_smem_extra(pl->res_inc, pl->halo_rows, coarse);
    if (coarse && pl->smem + extra > 226 * 1024) {
      coarse = 0;
      extra = engine_smem_extra(pl->res_inc, pl->halo_rows, 0);
    }
    pl->coarse = coarse;
    if (pl->halo_rows >= 65536 || pl->smem + extra > 226 * 1024) {
      pl->coarse = 0;
      pl->halo_rows = 0;  // does not fit: the general CG loop (exchange through L2) runs instead
      pl->inc_halo.clear();
      pl->push_ptr.clear();
      pl->push_row.clear();
      pl->push_dst.clear();
      pl->xinc_ptr.clear();
      pl->xinc_idx.clear();
      continue;
    }
    pl->smem += extra;
  }

  hprof.mark("build_halo");
  // ---- input arena
  size_t need = 0;
  auto sz = [&](size_t bytes) { need += ((bytes + 255) & ~size_t(255)) + 256; };
  sz(7 * F * 8); sz(4 * (size_t)V * 8); sz(4 * (size_t)V * 8); sz(2 * (size_t)V * 8); sz((size_t)V * 4);
  sz((size_t)P * 4); sz((size_t)P * 4); sz((size_t)P * 8); sz((size_t)P * 8);
  sz(((size_t)V + 1) * 4); sz(2 * (size_t)P * 4); sz(2 * (size_t)P * 4); sz(2 * (size_t)P * 4);
  sz(4 * (size_t)D * 4); sz((size_t)D * 8); sz(((size_t)V + 1) * 4); sz(4 * (size_t)D * 4);
  sz(((size_t)V + 1) * 4); sz((size_t)U * 8); sz((size_t)U * 4); sz((size_t)V);
  sz((size_t)V); sz((size_t)V); sz((size_t)P);
  sz((size_t)max_chunks * 4 * 3 * 2); sz(((size_t)F + 1) * 4 * 2);
  for (Plan* pl : {&planA, &planB}) {
    sz(pl->wseg_ptr.size() * 4); sz(pl->wseg_begin.size() * 4); sz(pl->wseg_end.size() * 4); sz(pl->wseg_kf.size() * 4);
    sz(pl->kf_wseg_ptr.size() * 4);
    sz(pl->inc_halo.size() * 4); sz(pl->push_ptr.size() * 4); sz(pl->push_row.size() * 4); sz(pl->push_dst.size() * 4); sz(pl->xinc_ptr.size() * 4); sz(pl->xinc_idx.size() * 4);
  }
  std::vector<int> inc_pos;
  if (direct) {
    direct_inc_pos(dplan, inc_ptr, inc_other, inc_pos, &ctx->pool);
    const size_t T1 = (size_t)dplan.n_nodes + 2;
    sz(T1 * 4); sz(T1 * 4); sz(T1 * 4); sz(T1 * 4); sz(T1 * 4); sz(T1 * 4); sz(T1 * 8); sz(T1 * 8);
    sz(dplan.bnd.size() * 4); sz(dplan.bpath.size() * 4); sz(dplan.inv.size() * 4); sz(inc_pos.size() * 4);
  }
  need += 8192;
  if (!st.in.reserve(need, true)) return fail(ctx, NRSLAM_B200_ERR_ALLOC, "input arena allocation failed");
  std::vector<CopyJob> copy_jobs;
  struct DeferGuard {  // every return path leaves put() in its immediate mode
    ~DeferGuard() { t_defer = nullptr; }
  } defer_guard;
  if (ctx->pool.threads() > 1 || V >= 4096) t_defer = &copy_jobs;
  Arena& in = st.in;
  Params& p = st.p;
  memset(&p, 0, sizeof(p));
  p.F = F; p.V = V; p.P = P; p.D = D; p.n_chunks = n_chunks;  // V / chunks: overwritten by apply_plan below
  p.poses_fixed = hp.poses_fixed; p.points_fixed = hp.points_fixed; p.spring_kind = hp.spring_kind;
  p.cam = hp.cam;
  p.info_reproj = hp.info_reproj; p.delta_reproj = hp.delta_reproj;
  p.info_spatial = hp.info_spatial; p.delta_spatial = hp.delta_spatial;
  p.info_spring = hp.info_spring; p.delta_spring = hp.delta_spring; p.spring_k = hp.spring_k;
  p.th2f = hp.th2f; p.th3f = hp.th3f;
  p.lm_tau = ctx->opt.lm_tau; p.lm_max_trials = ctx->opt.lm_max_trials;
  p.pcg_tol = ctx->opt.pcg_rel_tol; p.pcg_max_iter = ctx->opt.pcg_max_iterations;
  p.no_dsmem = env_int("NRSLAM_B200_NO_DSMEM", 0);
  p.n_ops = (int)hp.ops.size();
  if (p.n_ops > kMaxOps) return fail(ctx, NRSLAM_B200_ERR_ARG, "program too long");
  for (int i = 0; i < p.n_ops; i++) {
    p.op[i] = hp.ops[i];
    p.op_arg[i] = hp.op_args[i];
  }
  p.pose_seed = in.d<double>(put(in, hp.pose_seed));
  if (hp.pose_seed_dev) {
    p.pose_seed = hp.pose_seed_dev;
    p.seed_via_f32 = 1;
  }
  p.x_seed = in.d<double>(put(in, hp.x_seed));
  p.rest = in.d<double>(put(in, hp.rest));
  p.uv = in.d<double>(put(in, hp.uv));
  p.pt_kf = in.d<int>(put(in, hp.pt_kf));
  p.pair_i = in.d<int>(put(in, hp.pair_i));
  p.pair_j = in.d<int>(put(in, hp.pair_j));
  p.pair_w = in.d<double>(put(in, hp.pair_w));
  p.pair_d0 = in.d<double>(put(in, hp.pair_d0));
  p.inc_ptr = in.d<int>(put(in, inc_ptr));
  p.inc_other = in.d<int>(put(in, inc_other));
  p.inc_ent = in.d<int>(put(in, inc_ent));
  p.inc_row = in.d<int>(put(in, inc_row));
  p.dmp_v = in.d<int>(put(in, hp.dmp_v, 4));
  p.dmp_w = in.d<double>(put(in, hp.dmp_w));
  p.dinc_ptr = in.d<int>(put(in, dinc_ptr));
  p.dinc_ent = in.d<int>(put(in, dinc_ent));
  if (!hp.un_ptr.empty()) {
    p.un_ptr = in.d<int>(put(in, hp.un_ptr));
    p.un_w = in.d<double>(put(in, hp.un_w));
    p.un_ref = in.d<int>(put(in, hp.un_ref));
  }
  if (hp.want_fixed_flags) {
    st.o_fixed = in.take<unsigned char>(V);
    memset(in.h<unsigned char>(st.o_fixed), 0, V);
  }
  if (!hp.fixed0.empty()) p.pt_fixed = in.d<unsigned char>(put(in, hp.fixed0));
  p.unary_on = hp.unary_on ? 1 : 0;
  const unsigned char *d_rp0 = nullptr, *d_sp0 = nullptr;
  if (!hp.rp_level0.empty()) d_rp0 = in.d<unsigned char>(put(in, hp.rp_level0));
  if (!hp.sp_level0.empty()) d_sp0 = in.d<unsigned char>(put(in, hp.sp_level0));
  for (Plan* pl : {&planA, &planB}) {
    if (pl->n_chunks == 0) continue;
    pl->d_chunk_kf = in.d<int>(put(in, pl->chunk_kf));
    pl->d_chunk_begin = in.d<int>(put(in, pl->chunk_begin));
    pl->d_chunk_end = in.d<int>(put(in, pl->chunk_end));
    pl->d_kf_chunk_ptr = in.d<int>(put(in, pl->kf_chunk_ptr));
    if (!pl->wseg_ptr.empty()) {
      pl->d_wseg_ptr = in.d<int>(put(in, pl->wseg_ptr));
      pl->d_wseg_begin = in.d<int>(put(in, pl->wseg_begin));
      pl->d_wseg_end = in.d<int>(put(in, pl->wseg_end));
      pl->d_wseg_kf = in.d<int>(put(in, pl->wseg_kf));
      pl->d_kf_wseg_ptr = in.d<int>(put(in, pl->kf_wseg_ptr));
    }
    if (!pl->push_ptr.empty()) {
      pl->d_inc_halo = in.d<int>(put(in, pl->inc_halo));
      pl->d_push_ptr = in.d<int>(put(in, pl->push_ptr));
      pl->d_push_row = in.d<int>(put(in, pl->push_row));
      pl->d_push_dst = in.d<int>(put(in, pl->push_dst));
      pl->d_xinc_ptr = in.d<int>(put(in, pl->xinc_ptr));
      pl->d_xinc_idx = in.d<int>(put(in, pl->xinc_idx));
    }
  }
  if (direct) {
    direct::Plan& dp = st.dq.pl;
    dp.V = Vu; dp.np = dplan.np; dp.depth = dplan.depth; dp.G = dplan.G; dp.max_path = dplan.max_path;
    dp.vb = in.d<int>(put(in, dplan.vb));
    dp.nv = in.d<int>(put(in, dplan.nv));
    dp.nbv = in.d<int>(put(in, dplan.nbv));
    dp.bnd_ptr = in.d<int>(put(in, dplan.bnd_ptr));
    dp.bnd = in.d<int>(put(in, dplan.bnd));
    dp.bpath = in.d<int>(put(in, dplan.bpath));
    dp.path_off = in.d<int>(put(in, dplan.path_off));
    dp.inv_ptr = in.d<int>(put(in, dplan.inv_ptr));
    dp.inv = in.d<int>(put(in, dplan.inv));
    dp.p_off = in.d<long long>(put(in, dplan.p_off));
    dp.u_off = in.d<long long>(put(in, dplan.u_off));
    st.dq.inc_pos = in.d<int>(put(in, inc_pos));
  }
  flush_copies(ctx->pool, copy_jobs, host_threads(64, V));
  apply_plan(p, planA);
  st.h2d_bytes = in.used();
  st.block = planA.block;
  st.smem = planA.smem;
  st.grid = planA.grid;
  st.has_plan2 = planB.n_chunks > 0;
  if (st.has_plan2) {
    st.block2 = planB.block;
    st.smem2 = planB.smem;
    st.grid2 = planB.grid;
  }

  hprof.mark("fill_in_arena");
  // ---- results + work arrays
  Arena& out = st.out;
  size_t oneed = 7 * F * 8 + 4 * (size_t)V * 8 + (size_t)V * 8 + V + P + sizeof(EngineStats) + 8 * 256 + 4096;
  if (!out.reserve(oneed, true)) return fail(ctx, NRSLAM_B200_ERR_ALLOC, "output arena allocation failed");
  st.o_pose = out.take<double>(7 * F);
  st.o_x = out.take<double>(4 * (size_t)V);
  st.o_chi2 = out.take<double>(V);
  st.o_rp_level = out.take<unsigned char>(V);
  st.o_sp_level = out.take<unsigned char>(std::max(P, 1));
  st.o_stats = out.take<EngineStats>(1);
  st.d2h_bytes = out.used();
  p.pose = out.d<double>(st.o_pose);
  p.x = out.d<double>(st.o_x);
  p.rp_chi2 = out.d<double>(st.o_chi2);
  p.rp_level = out.d<unsigned char>(st.o_rp_level);
  p.sp_level = out.d<unsigned char>(st.o_sp_level);
  p.stats = out.d<EngineStats>(st.o_stats);

  Arena& wk = st.work;
  size_t wneed = (size_t)V * 8 * (4 + 20 + 8 + 4 + 8 + 4 * 5) + (size_t)P * 40 + (size_t)D * 32 +
                 2 * (size_t)max_chunks * kChunkVals * 8 + 2 * (size_t)max_grid_used * kSlotVals * 8 + 32 * 256 + 4096;
  const size_t max_wseg = std::max(planA.wseg_begin.size(), planB.wseg_begin.size());
  if (max_wseg > 0) wneed += 2 * max_wseg * 8 * 8 + 2 * (size_t)P * 32 + 4 * (size_t)D * 32 + 4 * (size_t)V * 8 + 8 * 256;
  if (direct)
    wneed += ((size_t)dplan.p_total + (size_t)dplan.u_total + 18 * (size_t)V + (size_t)dplan.G * (28 + 4 + 32) + 2 * (size_t)dplan.n_nodes + 72) * 8 +
             16 * 256;
  if (!wk.reserve(wneed, false)) return fail(ctx, NRSLAM_B200_ERR_ALLOC, "work arena allocation failed");
  p.x_bak = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.jac = wk.d<double>(wk.take<double>(20 * (size_t)V));
  p.dg = wk.d<double>(wk.take<double>(8 * (size_t)V));
  p.bvec = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.minv = wk.d<double>(wk.take<double>(8 * (size_t)V));
  p.xcg = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.rvec = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.pvec = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.qvec = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.zvec = wk.d<double>(wk.take<double>(4 * (size_t)V));
  p.pc = wk.d<double>(wk.take<double>(4 * (size_t)std::max(P, 1)));
  p.pcc = wk.d<double>(wk.take<double>((size_t)std::max(P, 1)));
  p.dc = wk.d<double>(wk.take<double>(4 * (size_t)std::max(D, 1)));
  p.chunk_part = wk.d<double>(wk.take<double>(2 * (size_t)max_chunks * kChunkVals));
  p.slots = wk.d<double>(wk.take<double>(2 * (size_t)max_grid_used * kSlotVals));
  if (max_wseg > 0) {
    p.wseg_part = wk.d<double>(wk.take<double>(2 * max_wseg * 8));
    p.wvec = wk.d<double>(wk.take<double>(4 * (size_t)V));
    p.wrec = wk.d<double>(wk.take<double>(8 * (size_t)std::max(P, 1)));
    p.wdrec = wk.d<double>(wk.take<double>(16 * (size_t)std::max(D, 1)));
  }
  p.bar = ctx->bar;
  if (direct) {
    DirectParams& q = st.dq;
    q.pl.panel = wk.d<double>(wk.take<double>((size_t)dplan.p_total + 1));
    q.pl.upd = wk.d<double>(wk.take<double>((size_t)dplan.u_total + 1));
    q.pl.fail = wk.d<int>(wk.take<int>(4));
    q.cpl = wk.d<double>(wk.take<double>(18 * (size_t)V));
    q.hpp_part = wk.d<double>(wk.take<double>(28 * (size_t)dplan.G));
    q.hpp = wk.d<double>(wk.take<double>(28));
    q.dslots = wk.d<double>(wk.take<double>(4 * (size_t)dplan.G));
    q.dpose = wk.d<double>(wk.take<double>(8));
    q.scratch_z = dscratch;
    q.max_nv = dmaxnv;
    q.max_rows = dplan.max_rows;
    q.tbar = wk.d<unsigned long long>(wk.take<unsigned long long>(2 * (size_t)dplan.n_nodes + 8));
    q.plev = wk.d<long long>(wk.take<long long>(32 * (size_t)dplan.G));
    q.P = p;
    st.use_direct = true;
    st.dgrid = dplan.G;
    st.dfactor_doubles = dplan.p_total;
    st.dupdate_doubles = dplan.u_total;
    st.dsmem = dsmem;
  }
  if (st.has_plan2) {
    st.p2 = p;
    apply_plan(st.p2, planB);
  }

  NRS_CUDA(ctx, cudaMemcpyAsync(in.dev(), in.host(), in.used(), cudaMemcpyHostToDevice, ctx->stream));
  // dg[.][6] (unary diagonal weight) is read by the matvec even when no unary edges exist
  NRS_CUDA(ctx, cudaMemsetAsync(wk.dev(), 0, wk.used(), ctx->stream));
  NRS_CUDA(ctx, cudaMemsetAsync(out.dev(), 0, out.used(), ctx->stream));
  if (d_rp0) NRS_CUDA(ctx, cudaMemcpyAsync(p.rp_level, d_rp0, hp.rp_level0.size(), cudaMemcpyDeviceToDevice, ctx->stream));
  if (d_sp0) NRS_CUDA(ctx, cudaMemcpyAsync(p.sp_level, d_sp0, hp.sp_level0.size(), cudaMemcpyDeviceToDevice, ctx->stream));
  hprof.mark("enqueue_copies");
  hprof.print("stage_problem");
  st.valid = true;
  return 0;
}

// Launch the staged program WITHOUT waiting: the kernel and the copy of its (small) statistics are queued on the ctx
// stream between the second pair of events; finish_async() collects them after a later synchronisation.
int launch_async(nrslam_b200_ctx* ctx, Staged& st) {
  if (!st.valid) return fail(ctx, NRSLAM_B200_ERR_ARG, "no staged problem");
  NRS_CUDA(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
  const int rc = st.use_direct ? launch_direct(st.dq, st.dgrid, st.dsmem, ctx->stream)
                               : launch_engine(st.p, st.grid, st.block, st.smem, ctx->stream);
  if (rc != 0)
    return fail(ctx, NRSLAM_B200_ERR_CUDA, std::string("engine launch: ") + cudaGetErrorString((cudaError_t)rc));
  NRS_CUDA(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
  NRS_CUDA(ctx, cudaMemcpyAsync(st.out.host(), st.out.dev(), st.out.used(), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

void finish_async(nrslam_b200_ctx* ctx, Staged& st, nrslam_b200_stats* stats) {
  if (!stats) return;
  const EngineStats* es = st.out.h<EngineStats>(st.o_stats);
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
  stats->gpu_ms += ms;
  stats->lm_iterations += es->lm_iterations;
  stats->lm_trials += es->lm_trials;
  stats->pcg_iterations += es->pcg_iterations;
  stats->n_sweeps += es->n_sweeps;
  stats->n_chi2_passes += es->n_chi2_passes;
  stats->kernel_launches += 1;
  stats->grid_ctas = st.use_direct ? st.dgrid : st.grid;
  stats->block_threads = st.use_direct ? direct_block_threads() : st.block;
  stats->d2h_bytes += (int64_t)st.out.used();
  for (int i = 0; i < es->n_trace && stats->n_trace < NRSLAM_B200_TRACE; i++)
    stats->chi2_trace[stats->n_trace++] = es->chi2_trace[i];
  stats->lambda_final = es->lambda_final;
}

// run_staged in two halves: queue the launch and the copy back (between the events ev0 / ev1), then wait and collect.
// The caller may do host work in between.
int run_staged_begin(nrslam_b200_ctx* ctx, Staged& st, bool copy_back = true, const Params* override_params = nullptr,
                     bool plan2 = false) {
  if (!st.valid) return fail(ctx, NRSLAM_B200_ERR_ARG, "no staged problem");
  NRS_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  const bool direct = st.use_direct && !override_params && !plan2;
  const int rc = direct ? launch_direct(st.dq, st.dgrid, st.dsmem, ctx->stream)
                        : launch_engine(override_params ? *override_params : st.p, plan2 ? st.grid2 : st.grid,
                                        plan2 ? st.block2 : st.block, plan2 ? st.smem2 : st.smem, ctx->stream);
  if (rc != 0)
    return fail(ctx, NRSLAM_B200_ERR_CUDA, std::string("engine launch: ") + cudaGetErrorString((cudaError_t)rc));
  NRS_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  if (copy_back)
    NRS_CUDA(ctx, cudaMemcpyAsync(st.out.host(), st.out.dev(), st.out.used(), cudaMemcpyDeviceToHost, ctx->stream));
  else
    NRS_CUDA(ctx, cudaMemcpyAsync(st.out.h<char>(st.o_stats), st.out.d<char>(st.o_stats), sizeof(EngineStats),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

int run_staged_end(nrslam_b200_ctx* ctx, Staged& st, nrslam_b200_stats* stats, bool copy_back = true,
                   const Params* override_params = nullptr, bool plan2 = false) {
  const bool direct = st.use_direct && !override_params && !plan2;
  NRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) {
    const EngineStats* es = st.out.h<EngineStats>(st.o_stats);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    stats->gpu_ms += ms;
    stats->lm_iterations += es->lm_iterations;
    stats->lm_trials += es->lm_trials;
    stats->pcg_iterations += es->pcg_iterations;
    stats->n_sweeps += es->n_sweeps;
    stats->n_chi2_passes += es->n_chi2_passes;
    stats->kernel_launches += 1;
    if (!plan2) {
      stats->grid_ctas = direct ? st.dgrid : st.grid;
      stats->block_threads = direct ? direct_block_threads() : st.block;
    }
    stats->d2h_bytes += copy_back ? (int64_t)st.out.used() : (int64_t)sizeof(EngineStats);
    stats->solve_failures += es->pcg_fail;
    if (direct) {
      stats->direct_solves += es->lm_trials;
      stats->factor_doubles = st.dfactor_doubles;
      stats->update_doubles = st.dupdate_doubles;
      stats->plan_reused += st.plan_reused ? 1 : 0;
    }
    for (int i = 0; i < es->n_trace && stats->n_trace < NRSLAM_B200_TRACE; i++)
      stats->chi2_trace[stats->n_trace++] = es->chi2_trace[i];
    stats->lambda_final = es->lambda_final;
    if (getenv("NRSLAM_B200_PROF")) {
      fprintf(stderr, "[nrs prof] grid %d block %d barriers %d cycles:", direct ? st.dgrid : st.grid,
              direct ? direct_block_threads() : st.block, es->barriers);
      for (int i = 0; i < 16; i++) fprintf(stderr, " %lld", es->prof[i]);
      fprintf(stderr, "\n");
      if (direct && st.dq.plev && getenv("NRSLAM_B200_PLEV")) {
        std::vector<long long> lv(32 * (size_t)st.dgrid);
        cudaMemcpy(lv.data(), st.dq.plev, lv.size() * 8, cudaMemcpyDeviceToHost);
        const char* nm[3] = {"stageAB", "stageC", "backward"};
        for (int ph = 0; ph < 3; ph++)
          for (int d = 0; d <= st.dq.pl.depth; d++) {
            long long mx = 0, sum = 0;
            int arg = 0;
            for (int g = 0; g < st.dgrid; g++) {
              const long long v = lv[32 * (size_t)g + 8 * ph + d];
              sum += v;
              if (v > mx) { mx = v; arg = g; }
            }
            fprintf(stderr, "[nrs plev] %s level %d: max %lld (cta %d) mean %lld cta0 %lld\n", nm[ph], d, mx, arg,
                    sum / st.dgrid, lv[8 * ph + d]);
          }
      }
    }
  }
  return 0;
}

// Launch the staged program, bring the results back, fill stats. Blocks until done.
int run_staged(nrslam_b200_ctx* ctx, Staged& st, nrslam_b200_stats* stats, bool copy_back = true,
               const Params* override_params = nullptr, bool plan2 = false) {
  const int rc = run_staged_begin(ctx, st, copy_back, override_params, plan2);
  if (rc) return rc;
  return run_staged_end(ctx, st, stats, copy_back, override_params, plan2);
}

Cam to_cam(const nrslam_b200_camera* c) {
  Cam cam;
  cam.model = c->model;
  for (int i = 0; i < 8; i++) cam.p[i] = c->params[i];
  return cam;
}

}  // namespace

extern "C" {

void nrslam_b200_default_options(nrslam_b200_options* o) {
  if (!o) return;
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
  if (const char* e = getenv("NRSLAM_B200_PCG_TOL")) o->pcg_rel_tol = atof(e);  // experiments only
  o->pcg_max_iterations = 2000;
  const char* lr = getenv("LOCAL_RANK");
  o->device = lr ? atoi(lr) : 0;
  o->grid_ctas = 0;
}

int nrslam_b200_abi_version(void) { return NRSLAM_B200_ABI_VERSION; }

int nrslam_b200_create(const nrslam_b200_options* opt, nrslam_b200_ctx** out) {
  if (!out) return NRSLAM_B200_ERR_ARG;
  *out = nullptr;
  nrslam_b200_options o;
  if (opt)
    o = *opt;
  else
    nrslam_b200_default_options(&o);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return NRSLAM_B200_ERR_NO_DEVICE;
  if (o.device < 0 || o.device >= n_dev) return NRSLAM_B200_ERR_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, o.device) != cudaSuccess) return NRSLAM_B200_ERR_CUDA;
  if (prop.major != 10) return NRSLAM_B200_ERR_NO_DEVICE;  // sm_100a cubin only
  if (cudaSetDevice(o.device) != cudaSuccess) return NRSLAM_B200_ERR_CUDA;
  // The staging of a BA window allocates ~60 MB of host vectors per call. Above glibc's mmap threshold every one of
  // them is a fresh mapping whose pages fault in on first touch and are unmapped again on free: 10 of the 19 ms of
  // host staging on configs[3]. Serving them from the heap and not trimming it keeps the pages across calls.
  // (Process-wide allocator knobs: NRSLAM_B200_MALLOC_TUNE=0 leaves them alone.)
  if (env_int("NRSLAM_B200_MALLOC_TUNE", 1)) {
    mallopt(M_MMAP_THRESHOLD, 512 << 20);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
  }
  nrslam_b200_ctx* ctx = new nrslam_b200_ctx();
  ctx->opt = o;
  ctx->device = o.device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->ev2) != cudaSuccess || cudaEventCreate(&ctx->ev3) != cudaSuccess ||
      cudaMalloc(&ctx->bar, 256) != cudaSuccess) {
    delete ctx;
    return NRSLAM_B200_ERR_CUDA;
  }
  *out = ctx;
  return 0;
}

void nrslam_b200_destroy(nrslam_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& s : ctx->staged) {
    s.in.release();
    s.work.release();
    s.out.release();
  }
  for (int r = 0; r < nrs::kMaxWorld; r++)
    if (ctx->shard.peer[r] && r != ctx->shard.rank) cudaIpcCloseMemHandle(ctx->shard.peer[r]);
  if (ctx->shard.local) cudaFree(ctx->shard.local);
  for (auto& s : ctx->staged) s.xin.release();
  if (ctx->bar) cudaFree(ctx->bar);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->ev3) cudaEventDestroy(ctx->ev3);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* nrslam_b200_last_error(const nrslam_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nrslam_b200_device_info(const nrslam_b200_ctx* ctx, int32_t* device, int32_t* sm_count) {
  if (!ctx) return NRSLAM_B200_ERR_ARG;
  if (device) *device = ctx->device;
  if (sm_count) *sm_count = ctx->sm_count;
  return 0;
}

int32_t nrslam_b200_graph_get_edges(const nrslam_b200_graph* g, int32_t vertex, int32_t* out_entries,
                                    int32_t capacity) {
  if (!g || vertex < 0 || vertex >= g->n_vertices) return -1;
  std::vector<int> ent;
  const int n = graph_sorted_entries(g, vertex, graph_min_weight(g), ent);
  for (int i = 0; i < n && i < capacity; i++) out_entries[i] = ent[i];
  return n;
}

// ---- RegularizationGraph::AddEdge / SetSigma as an owning store (regularization_graph.cc:33-55) --------------------
}  // extern "C"
struct nrslam_b200_graph_store {
  float sigma = 0, stretching = 0;
  int n_vertices = 0;
  std::vector<int32_t> ev1, ev2;  // undirected edge -> endpoints (v1 < v2)
  std::vector<float> weight, first_distance, min_distance, max_distance;
  std::vector<uint8_t> status;
  std::unordered_map<uint64_t, int32_t> index;  // (v1 << 32 | v2) -> edge id
  bool dirty = true;
  std::vector<int32_t> rowptr, col, eid;
};
extern "C" {
int nrslam_b200_graph_store_create(float weight_sigma, float stretching_th, nrslam_b200_graph_store** out) {
  if (!out || !(weight_sigma > 0)) return NRSLAM_B200_ERR_ARG;
  *out = new nrslam_b200_graph_store();
  (*out)->sigma = weight_sigma;
  (*out)->stretching = stretching_th;
  return 0;
}

void nrslam_b200_graph_store_destroy(nrslam_b200_graph_store* store) { delete store; }

int nrslam_b200_graph_store_set_sigma(nrslam_b200_graph_store* store, float weight_sigma) {
  if (!store || !(weight_sigma > 0)) return NRSLAM_B200_ERR_ARG;
  store->sigma = weight_sigma;  // min_weight follows from sigma (graph_min_weight)
  return 0;
}

int nrslam_b200_graph_store_add_edges(nrslam_b200_graph_store* st, int32_t n, const int32_t* v1, const int32_t* v2,
                                      const float* rel) {
  if (!st || n < 0 || (n > 0 && (!v1 || !v2 || !rel))) return NRSLAM_B200_ERR_ARG;
  for (int k = 0; k < n; k++)
    if (v1[k] < 0 || v2[k] < 0 || v1[k] == v2[k]) return NRSLAM_B200_ERR_ARG;  // nothing is modified on an error
  for (int k = 0; k < n; k++) {
    const int32_t a = std::min(v1[k], v2[k]), b = std::max(v1[k], v2[k]);
    const float* r = rel + 3 * (size_t)k;
    const float distance = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);  // Eigen::Vector3f::norm()
    const uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    auto it = st->index.find(key);
    int32_t e;
    if (it == st->index.end()) {
      e = (int32_t)st->ev1.size();
      st->index.emplace(key, e);
      st->ev1.push_back(a);
      st->ev2.push_back(b);
      st->weight.push_back(0);
      st->first_distance.push_back(0);
      st->min_distance.push_back(0);
      st->max_distance.push_back(0);
      st->status.push_back(0);
      st->dirty = true;
      st->n_vertices = std::max(st->n_vertices, b + 1);
    } else {
      e = it->second;  // graph_[a][b] = edge: the pair's record is replaced
    }
    st->weight[e] = interpolation_weight(distance, st->sigma);
    st->first_distance[e] = st->min_distance[e] = st->max_distance[e] = distance;
    st->status[e] = NRSLAM_EDGE_NEUTRAL;
  }
  return 0;
}

int nrslam_b200_graph_store_view(nrslam_b200_graph_store* st, nrslam_b200_graph* out) {
  if (!st || !out) return NRSLAM_B200_ERR_ARG;
  if (st->dirty) {
    const int M = st->n_vertices, E = (int)st->ev1.size();
    st->rowptr.assign(M + 1, 0);
    for (int e = 0; e < E; e++) {
      st->rowptr[st->ev1[e] + 1]++;
      st->rowptr[st->ev2[e] + 1]++;
    }
    for (int v = 0; v < M; v++) st->rowptr[v + 1] += st->rowptr[v];
    std::vector<std::pair<int32_t, int32_t>> ent(2 * (size_t)E);  // (neighbour, edge) per CSR slot
    {
      std::vector<int32_t> w(st->rowptr.begin(), st->rowptr.end() - 1);
      for (int e = 0; e < E; e++) {
        ent[w[st->ev1[e]]++] = {st->ev2[e], e};
        ent[w[st->ev2[e]]++] = {st->ev1[e], e};
      }
    }
    st->col.resize(2 * (size_t)E);
    st->eid.resize(2 * (size_t)E);
    for (int v = 0; v < M; v++) {  // row entries ascending by neighbour (the btree_map order)
      std::sort(ent.begin() + st->rowptr[v], ent.begin() + st->rowptr[v + 1]);
      for (int p = st->rowptr[v]; p < st->rowptr[v + 1]; p++) {
        st->col[p] = ent[p].first;
        st->eid[p] = ent[p].second;
      }
    }
    st->dirty = false;
  }
  out->n_vertices = st->n_vertices;
  out->n_edges = (int32_t)st->ev1.size();
  out->rowptr = st->rowptr.data();
  out->col = st->col.data();
  out->eid = st->eid.data();
  out->weight = st->weight.data();
  out->first_distance = st->first_distance.data();
  out->min_distance = st->min_distance.data();
  out->max_distance = st->max_distance.data();
  out->status = st->status.data();
  out->weight_sigma = st->sigma;
  out->stretching_th = st->stretching;
  return 0;
}

int32_t nrslam_b200_graph_update_vertex(nrslam_b200_graph* g, int32_t vertex, const float* positions) {
  if (!g || vertex < 0 || vertex >= g->n_vertices || !positions) return -1;
  return graph_update_vertex(g, vertex, positions);
}

// =====================================================================================================
// CameraPoseOptimization — g2o_optimization.cc:50-146
// =====================================================================================================
}  // extern "C"

namespace {
int pose_deform_impl(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                     const float* X_rest, const int32_t* point_vertex, const int8_t* vfs, nrslam_b200_graph* g,
                     float scale, float* pose_io, float* last_pos, float* deformation_out, float* X_out,
                     float* chi2_out, uint8_t* status_out, float* median_deformation_out, int32_t* lost_vertex_out,
                     int32_t* n_lost_out, nrslam_b200_stats* stats, const double* pose_seed_dev);

// async: stage + launch only (results are collected by the caller after a later synchronisation of the ctx stream)
int pose_only_impl(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv, const float* X,
                   float* pose_io, uint8_t* inlier_out, nrslam_b200_stats* stats, bool async) {
  if (!ctx || !cam || !uv || !X || !pose_io || n < 0) return fail(ctx, NRSLAM_B200_ERR_ARG, "pose_only: bad argument");
  const double t0 = wall_ms();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n == 0) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "pose_only: no observations");
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  const nrslam_b200_options& opt = ctx->opt;
  HostProblem hp;
  hp.F = 1;
  hp.V = n;
  hp.points_fixed = true;
  hp.cam = to_cam(cam);
  hp.info_reproj = 1.0;                                    // :90 Matrix2d::Identity()
  hp.delta_reproj = (double)std::sqrt(opt.th_huber_2dof_sq);  // :64 float sqrt
  hp.th2f = opt.th_huber_2dof_sq;
  hp.pose_seed.resize(7);
  pose_from_f7(pose_io, hp.pose_seed.data());
  hp.x_seed.assign(4 * (size_t)n, 0.0);
  hp.rest.resize(4 * (size_t)n);
  hp.uv.resize(2 * (size_t)n);
  hp.pt_kf.assign(n, 0);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) hp.rest[4 * (size_t)i + k] = X[3 * (size_t)i + k];
    hp.rest[4 * (size_t)i + 3] = 0;
    hp.uv[2 * (size_t)i] = uv[2 * (size_t)i];
    hp.uv[2 * (size_t)i + 1] = uv[2 * (size_t)i + 1];
  }
  hp.kf_begin = {0, n};
  hp.ops.push_back(OP_CLEAR_LEVELS);
  hp.op_args.push_back(0);
  for (int it = 0; it < 3; it++) {
    hp.ops.push_back(OP_RESET);
    hp.op_args.push_back(0);
    hp.ops.push_back(OP_OPTIMIZE);
    hp.op_args.push_back(opt.pose_only_iterations[it]);
    hp.ops.push_back(OP_RELEVEL_POSE);
    hp.op_args.push_back(0);
  }
  Staged& st = ctx->staged[0];
  int rc = stage_problem(ctx, st, hp);
  if (rc) return rc;
  const double t1 = wall_ms();
  if (stats) stats->h2d_bytes += (int64_t)st.h2d_bytes;
  if (async) {
    if (stats) {
      stats->n_reproj_edges = n;
      stats->n_poses = 1;
      stats->stage_ms = (float)(t1 - t0);
    }
    return launch_async(ctx, st);
  }
  rc = run_staged(ctx, st, stats);
  if (rc) return rc;
  const double* pose = st.out.h<double>(st.o_pose);
  for (int i = 0; i < 7; i++)
    if (!std::isfinite(pose[i])) return fail(ctx, NRSLAM_B200_NUM_NONFINITE, "pose_only: non-finite pose");
  pose_to_f7(pose, pose_io);
  if (inlier_out) {
    const unsigned char* lvl = st.out.h<unsigned char>(st.o_rp_level);
    for (int i = 0; i < n; i++) inlier_out[i] = lvl[i] == 0;
  }
  if (stats) {
    stats->n_reproj_edges = n;
    stats->n_poses = 1;
    stats->stage_ms = (float)(t1 - t0);
    stats->host_ms = (float)(wall_ms() - t0);
  }
  return 0;
}
}  // namespace

extern "C" {
int nrslam_b200_pose_only(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                          const float* X, float* pose_io, uint8_t* inlier_out, nrslam_b200_stats* stats) {
  return pose_only_impl(ctx, cam, n, uv, X, pose_io, inlier_out, stats, false);
}

int nrslam_b200_pose_deform(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                            const float* X_rest, const int32_t* point_vertex,
                            const int8_t* vfs, nrslam_b200_graph* g, float scale, float* pose_io,
                            float* last_pos, float* deformation_out, float* X_out, float* chi2_out,
                            uint8_t* status_out, float* median_deformation_out, int32_t* lost_vertex_out,
                            int32_t* n_lost_out, nrslam_b200_stats* stats) {
  return pose_deform_impl(ctx, cam, n, uv, X_rest, point_vertex, vfs, g, scale, pose_io, last_pos, deformation_out,
                          X_out, chi2_out, status_out, median_deformation_out, lost_vertex_out, n_lost_out, stats,
                          nullptr);
}

// Tracking::TrackCameraAndDeformation after the data association (tracking.cc:291-330): CameraPoseOptimization, then
// CameraPoseAndDeformationOptimization seeded with its result, on the same TRACKED_WITH_3D points. Same results as the
// two calls in sequence (bit for bit: the pose+deformation kernel reads the seed pose straight from the pose-only
// kernel's output in HBM), but the host staging of the second problem overlaps the first kernel.
int nrslam_b200_track_pose_and_deform(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                                      const float* X_rest, const int32_t* point_vertex, const int8_t* vfs,
                                      nrslam_b200_graph* g, float scale, float* pose_io, float* last_pos,
                                      float* deformation_out, float* X_out, float* chi2_out, uint8_t* status_out,
                                      float* median_deformation_out, int32_t* lost_vertex_out, int32_t* n_lost_out,
                                      float* pose_only_out, uint8_t* pose_only_inlier_out,
                                      nrslam_b200_stats* stats_pose_only, nrslam_b200_stats* stats) {
  if (!ctx || !pose_io) return fail(ctx, NRSLAM_B200_ERR_ARG, "track_pose_and_deform: bad argument");
  float seed[7];
  memcpy(seed, pose_io, sizeof(seed));
  if (n > 0 && frame_threads(n) > 1) {  // the staging threads wake up while the pose-only problem is staged
    ctx->pool.start(frame_threads(n) - 1);
    ctx->pool.begin();
  }
  int rc = pose_only_impl(ctx, cam, n, uv, X_rest, seed, nullptr, stats_pose_only, true);
  if (rc) {
    ctx->pool.end();
    return rc;
  }
  nrs::Staged& st0 = ctx->staged[0];
  rc = pose_deform_impl(ctx, cam, n, uv, X_rest, point_vertex, vfs, g, scale, pose_io, last_pos, deformation_out, X_out,
                        chi2_out, status_out, median_deformation_out, lost_vertex_out, n_lost_out, stats,
                        st0.out.d<double>(st0.o_pose));
  // the stream has been synchronised by the second stage: the pose-only results are on the host
  finish_async(ctx, st0, stats_pose_only);
  const double* pose0 = st0.out.h<double>(st0.o_pose);
  if (pose_only_out) pose_to_f7(pose0, pose_only_out);
  if (pose_only_inlier_out) {
    const unsigned char* lvl = st0.out.h<unsigned char>(st0.o_rp_level);
    for (int i = 0; i < n; i++) pose_only_inlier_out[i] = lvl[i] == 0;
  }
  if (rc) return rc;
  for (int i = 0; i < 7; i++)
    if (!std::isfinite(pose0[i])) return fail(ctx, NRSLAM_B200_NUM_NONFINITE, "pose_only: non-finite pose");
  return 0;
}
}  // extern "C"

namespace {
// =====================================================================================================
// CameraPoseAndDeformationOptimization — g2o_optimization.cc:148-557
// =====================================================================================================
int pose_deform_impl(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t n, const float* uv,
                     const float* X_rest, const int32_t* point_vertex, const int8_t* vfs, nrslam_b200_graph* g,
                     float scale, float* pose_io, float* last_pos, float* deformation_out, float* X_out,
                     float* chi2_out, uint8_t* status_out, float* median_deformation_out, int32_t* lost_vertex_out,
                     int32_t* n_lost_out, nrslam_b200_stats* stats, const double* pose_seed_dev) {
  if (!ctx || !cam || !uv || !X_rest || !point_vertex || !vfs || !g || !pose_io || !last_pos || n < 0)
    return fail(ctx, NRSLAM_B200_ERR_ARG, "pose_deform: bad argument");
  const double t0 = wall_ms();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_lost_out) *n_lost_out = 0;
  if (n == 0) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "pose_deform: no observations");
  if (frame_threads(n) > 1) {  // wake the staging threads now: they spin until the regulariser selection is done
    ctx->pool.start(frame_threads(n) - 1);
    ctx->pool.begin();
  }
  struct SpinWindow {  // whatever path leaves this function, the workers go back to sleep
    HostPool& p;
    ~SpinWindow() { p.end(); }
  } spin_window{ctx->pool};
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  const nrslam_b200_options& opt = ctx->opt;
  const int M = g->n_vertices;
  for (int i = 0; i < n; i++)
    if (point_vertex[i] < 0 || point_vertex[i] >= M) return fail(ctx, NRSLAM_B200_ERR_ARG, "pose_deform: bad vertex");
  HostProf hprof;

  const int regularizers_per_point = opt.regularizers_per_point;
  const float th2 = opt.th_huber_2dof_sq, th3 = opt.th_huber_3dof_sq;
  const float info_reprojection = 1.0f / (opt.sigma_reprojection * opt.sigma_reprojection);  // :203-204
  const float info_position = 1.0f / (opt.sigma_position * opt.sigma_position);              // :206-207
  const float sigma_spatial = (float)((double)opt.sigma_spatial_factor * scale);             // :209
  const float info_spatial = 1.0f / (sigma_spatial * sigma_spatial);
  const float min_w = graph_min_weight(g);

  HostProblem hp;
  hp.F = 1;
  hp.V = n;
  hp.cam = to_cam(cam);
  hp.spring_kind = SPRING_DEFORM;
  hp.info_reproj = info_reprojection;
  hp.delta_reproj = (double)std::sqrt(th2);
  hp.info_spatial = info_spatial;
  hp.delta_spatial = (double)std::sqrt(th3);
  hp.info_spring = info_position;
  hp.delta_spring = (double)std::sqrt(th3);
  hp.spring_k = (double)opt.spring_k;  // :328 float literal 1.1f widened
  hp.th2f = th2;
  hp.th3f = th3;
  hp.pose_seed.resize(7);
  pose_from_f7(pose_io, hp.pose_seed.data());
  hp.pose_seed_dev = pose_seed_dev;
  hp.x_seed.assign(4 * (size_t)n, 0.0);  // deformation vertices start at the origin (:188,345-348)
  hp.rest.resize(4 * (size_t)n);
  hp.uv.resize(2 * (size_t)n);
  hp.pt_kf.assign(n, 0);
  std::vector<int> opt_index(M, -1);  // mappoint_id_to_index (:177,191)
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) hp.rest[4 * (size_t)i + k] = X_rest[3 * (size_t)i + k];
    hp.rest[4 * (size_t)i + 3] = 0;
    hp.uv[2 * (size_t)i] = uv[2 * (size_t)i];
    hp.uv[2 * (size_t)i + 1] = uv[2 * (size_t)i + 1];
    opt_index[point_vertex[i]] = i;
  }
  hp.kf_begin = {0, n};

  hprof.mark("fill_rows");
  // ---- regulariser selection (:251-336). A pair is owned by the first endpoint that reaches it.
  std::vector<unsigned char> pair_stamp(g->n_edges, 0);  // spatial_connections_ids de-duplication, keyed by graph edge
  std::vector<unsigned char> lost_flag(M, 0);            // absl::btree_set<ID> (:222): flags now, ascending ids below
  std::vector<int> ent;
  // GetEdges of every point (sort of its connections, regularization_graph.cc:71-87) is independent of the other
  // points: host threads fill the sorted lists, the order-dependent selection below walks them sequentially
  std::vector<int> ent_ptr(n + 1, 0), ent_cnt(n, 0);
  for (int idx = 0; idx < n; idx++)
    ent_ptr[idx + 1] = ent_ptr[idx] + (g->rowptr[point_vertex[idx] + 1] - g->rowptr[point_vertex[idx]]);
  std::vector<int> ent_flat(ent_ptr[n]);
  ctx->pool.run([&](int t, int nt) {
    std::vector<int> loc;
    const int b = (int)((long long)n * t / nt), e = (int)((long long)n * (t + 1) / nt);
    for (int idx = b; idx < e; idx++) {
      loc.clear();
      ent_cnt[idx] = graph_sorted_entries(g, point_vertex[idx], min_w, loc);
      std::copy(loc.begin(), loc.end(), ent_flat.begin() + ent_ptr[idx]);
    }
  });
  {
    const int32_t *gcol = g->col, *geid = g->eid;
    const uint8_t* gstatus = g->status;
    const float *gweight = g->weight, *gfirst = g->first_distance;
    const size_t cap = (size_t)std::min<long long>(ent_ptr[n], (long long)n * (regularizers_per_point + 1));
    hp.pair_i.reserve(cap);
    hp.pair_j.reserve(cap);
    hp.pair_w.reserve(cap);
    hp.pair_d0.reserve(cap);
    for (int idx = 0; idx < n; idx++) {
      int n_regularizers = 0;
      const int* el = ent_flat.data() + ent_ptr[idx];
      for (int q = 0, nq = ent_cnt[idx]; q < nq; q++) {
        const int pe = el[q];
        const int other = gcol[pe], ge = geid[pe];
        if (n_regularizers > regularizers_per_point || gstatus[ge] == NRSLAM_EDGE_BAD) break;  // :258-261
        const int vs = vfs[other];
        if (vs != NRSLAM_TRACKED_WITH_3D) {  // :264-273 (negative: not in the frame)
          if (vs >= 0 && vs != NRSLAM_JUST_TRIANGULATED) lost_flag[other] = 1;
          continue;
        }
        const int idx_other = opt_index[other];
        if (idx_other < 0) continue;   // inconsistent input: TRACKED_WITH_3D in the frame but not optimised
        if (pair_stamp[ge]) continue;  // :277-279
        pair_stamp[ge] = 1;
        hp.pair_i.push_back(idx);
        hp.pair_j.push_back(idx_other);
        hp.pair_w.push_back((double)gweight[ge]);
        hp.pair_d0.push_back((double)gfirst[ge]);
        n_regularizers++;
      }
    }
  }
  // ---- lost neighbours (:476-537): one extra vertex each, unary SpatialRegularizerFixed edges to at most 11
  // optimised neighbours; they are optimised by a second, compact problem after the graph refresh below (the
  // graph is only modified by UpdateVertex, whose weights / statuses the reference does see when it queries
  // GetEdges for the lost points — so the unary edge list is built after the refresh).
  hprof.mark("select_regularisers");
  std::vector<int> lost_list;
  for (int v = 0; v < M; v++)
    if (lost_flag[v]) lost_list.push_back(v);
  const int n_lost = (int)lost_list.size();
  hp.n_sort = n;
  hp.direct = env_int("NRSLAM_B200_DIRECT", 1) != 0;
  const std::vector<int32_t> plan_key(point_vertex, point_vertex + n);
  hp.plan_key = &plan_key;
  hp.ops = {OP_CLEAR_LEVELS, OP_RESET, OP_OPTIMIZE, OP_RELEVEL_DEFORM, OP_RESET, OP_OPTIMIZE, OP_RELEVEL_DEFORM,
            OP_FINAL_CHI2};
  hp.op_args = {0, 0, opt.pose_deform_iterations[0], 0, 0, opt.pose_deform_iterations[1], 0, 0};

  Staged& st = ctx->staged[1];
  int rc = stage_problem(ctx, st, hp);
  if (rc) return rc;
  hprof.mark("stage");
  const double t1 = wall_ms();
  if (stats) stats->h2d_bytes += (int64_t)st.h2d_bytes;
  ctx->pool.end();  // the staging threads sleep while the main rounds run on the device ...
  rc = run_staged(ctx, st, stats);
  if (rc) return rc;
  if (frame_threads(n) > 1) ctx->pool.begin();  // ... and spin again for the graph refresh and the lost-point staging
  hprof.mark("run1");

  const double* pose = st.out.h<double>(st.o_pose);
  for (int i = 0; i < 7; i++)
    if (!std::isfinite(pose[i])) return fail(ctx, NRSLAM_B200_NUM_NONFINITE, "pose_deform: non-finite pose");
  pose_to_f7(pose, pose_io);  // :398-399
  const double* xd = st.out.h<double>(st.o_x);
  const double* chi2d = st.out.h<double>(st.o_chi2);
  const std::vector<int>& row_of = st.row_of;  // caller row -> engine row

  // ---- deformation magnitudes, IQR gate (:401-455)
  std::vector<float> mags(n), def(3 * (size_t)n);
  for (int idx = 0; idx < n; idx++) {
    float* d = &def[3 * (size_t)idx];
    for (int k = 0; k < 3; k++) d[k] = (float)xd[4 * (size_t)row_of[idx] + k];
    mags[idx] = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (deformation_out)
      for (int k = 0; k < 3; k++) deformation_out[3 * (size_t)idx + k] = d[k];
  }
  // order statistics of the sorted magnitudes (:411-416 sorts a copy; the two elements are the same by selection)
  std::vector<float> sorted(mags);
  const int iq1 = (int)(sorted.size() * 0.25f), iq3 = (int)(sorted.size() * 0.75f);
  std::nth_element(sorted.begin(), sorted.begin() + iq3, sorted.end());
  const float q3 = sorted[iq3];
  std::nth_element(sorted.begin(), sorted.begin() + iq1, sorted.begin() + iq3 + 1);
  const float q1 = sorted[iq1];
  const float th_ = 1.5f * (q3 - q1);
  std::vector<char> inliers(n, 1);
  std::vector<uint8_t> status(n, NRSLAM_TRACKED_WITH_3D);
  std::vector<unsigned char> fixed(n, 0);
  for (int idx = 0; idx < n; idx++) {
    const float chi_squared = (float)chi2d[row_of[idx]];
    if (chi2_out) chi2_out[idx] = chi_squared;
    if (chi_squared > th2) {
      inliers[idx] = 0;
      status[idx] = NRSLAM_TRACKED;
    }
    if (X_out)
      for (int k = 0; k < 3; k++) X_out[3 * (size_t)idx + k] = X_rest[3 * (size_t)idx + k];
    if (mags[idx] >= q3 + th_) {
      status[idx] = NRSLAM_TRACKED;
      continue;
    }
    fixed[row_of[idx]] = 1;  // :439 setFixed(true)
    for (int k = 0; k < 3; k++) {
      const float cur = def[3 * (size_t)idx + k] + X_rest[3 * (size_t)idx + k];
      if (X_out) X_out[3 * (size_t)idx + k] = cur;
      last_pos[3 * (size_t)point_vertex[idx] + k] = cur;  // :446
    }
  }
  if (median_deformation_out) {
    std::vector<float> m2(mags);
    const int median_idx = (int)m2.size() / 2;
    std::nth_element(m2.begin(), m2.begin() + median_idx, m2.end());
    *median_deformation_out = m2[median_idx];
  }
  hprof.mark("gate");
  // ---- regularisation-graph refresh (:458-474)
  // Positions are fixed during the loop and UpdateConnection is idempotent (max / min / weight / "set BAD" from the
  // two endpoint positions only, regularization_graph.cc:89-128), so the loop is order independent: host threads take
  // ranges of points. An edge with two accepted endpoints is written twice with identical values — through relaxed
  // atomics-free stores of the same bits, which is benign on every platform this library targets.
  // The lost-point stage below reads the refreshed weights / statuses of the edges at the LOST vertices only, so the
  // points adjacent to a lost vertex are refreshed first and the rest of the loop runs while the lost-point kernel
  // is on the device (same result: the loop is order independent).
  std::vector<char> prio(n, n_lost > 0 ? 0 : 1);
  for (int v = 0; v < n_lost; v++)
    for (int pe = g->rowptr[lost_list[v]]; pe < g->rowptr[lost_list[v] + 1]; pe++)
      if (opt_index[g->col[pe]] >= 0) prio[opt_index[g->col[pe]]] = 1;
  auto refresh = [&](char which) {
    ctx->pool.run([&](int t, int nt) {
      const int b = (int)((long long)n * t / nt), e = (int)((long long)n * (t + 1) / nt);
      for (int idx = b; idx < e; idx++) {
        if (!inliers[idx] || prio[idx] != which) continue;
        const int good = graph_update_vertex(g, point_vertex[idx], last_pos);
        if (good < regularizers_per_point * 0.5) status[idx] = NRSLAM_BAD;
      }
    });
  };
  refresh(1);
  bool rest_done = n_lost == 0;
  hprof.mark("update_vertex");
  if (stats) {
    stats->n_reproj_edges = n;
    stats->n_pair_edges = (int)hp.pair_i.size();
    stats->n_points = n;
    stats->n_poses = 1;
    stats->stage_ms = (float)(t1 - t0);
  }

  // ---- lost-point stage (:476-555). The reference re-runs the optimiser on the same graph with the pose and
  // the accepted deformation vertices fixed: the gated-out deformation vertices stay free and are re-optimised
  // alongside the lost points (their results are discarded), sharing one lambda / gain-ratio sequence.
  if (n_lost > 0) {
    // Staged as its own COMPACT problem: only vertices that are still free (gated-out deformation vertices, lost
    // points) and the fixed vertices their edges read are rows; an edge between two fixed vertices is not part of
    // g2o's active set (sparse_optimizer.cpp:232-246), so dropping it changes nothing. ~10x fewer rows than the frame.
    const unsigned char* rp_lvl = st.out.h<unsigned char>(st.o_rp_level);
    const unsigned char* sp_lvl = st.out.h<unsigned char>(st.o_sp_level);
    const float min_w2 = graph_min_weight(g);
    // unary edges of the lost points (:476-537), references by engine row of the main problem
    std::vector<int> un_ptr_l(n_lost + 1, 0), un_ref_row;
    std::vector<double> un_w_l;
    for (int v = 0; v < n_lost; v++) {
      ent.clear();
      graph_sorted_entries(g, lost_list[v], min_w2, ent);
      int n_regularizers = 0;
      for (int pe : ent) {
        if (n_regularizers > 10) break;  // :497 literal
        const int other = g->col[pe];
        if (opt_index[other] < 0) continue;  // :502-504
        un_w_l.push_back((double)g->weight[g->eid[pe]]);
        un_ref_row.push_back(row_of[opt_index[other]]);
        n_regularizers++;
      }
      un_ptr_l[v + 1] = (int)un_w_l.size();
    }
    const int n_un = (int)un_w_l.size();
    if (n_un > 0) {
      // compact row numbering: free deformation rows, lost rows, then the fixed rows they touch
      std::vector<int> crow(n, -1), rows;  // main engine row -> compact row; compact -> main engine row
      for (int r = 0; r < n; r++)
        if (!fixed[r]) {
          crow[r] = (int)rows.size();
          rows.push_back(r);
        }
      const int n_free = (int)rows.size();
      const int lost0 = n_free;  // compact rows [lost0, lost0 + n_lost) are the lost points
      auto touch = [&](int r) {
        if (crow[r] < 0) {
          crow[r] = n_lost + (int)rows.size();
          rows.push_back(r);
        }
      };
      const int Pm = (int)hp.pair_i.size();
      std::vector<int> keep;
      for (int e = 0; e < Pm; e++) {
        const int i = hp.pair_i[e], j = hp.pair_j[e];
        if (fixed[i] && fixed[j]) continue;
        touch(i);
        touch(j);
        keep.push_back(e);
      }
      for (int r : un_ref_row) touch(r);
      // crow of the non-free rows was assigned with the lost block already skipped
      const int Vc = n_lost + (int)rows.size();
      auto cidx = [&](int r) { return crow[r]; };
      HostProblem h2;
      h2.F = 1;
      h2.V = Vc;
      h2.cam = hp.cam;
      h2.poses_fixed = true;
      h2.unary_on = true;
      h2.plain_jacobi = env_int("NRSLAM_B200_LOST_JACOBI", 1) != 0;
      h2.spring_kind = hp.spring_kind;
      h2.info_reproj = hp.info_reproj; h2.delta_reproj = hp.delta_reproj;
      h2.info_spatial = hp.info_spatial; h2.delta_spatial = hp.delta_spatial;
      h2.info_spring = hp.info_spring; h2.delta_spring = hp.delta_spring; h2.spring_k = hp.spring_k;
      h2.th2f = hp.th2f; h2.th3f = hp.th3f;
      h2.pose_seed.assign(pose, pose + 7);
      h2.x_seed.assign(4 * (size_t)Vc, 0.0);
      h2.rest.assign(4 * (size_t)Vc, 0.0);
      h2.uv.assign(2 * (size_t)Vc, 0.0);
      h2.pt_kf.assign(Vc, -1);
      h2.fixed0.assign(Vc, 0);
      h2.rp_level0.assign(Vc, 0);
      for (size_t t = 0; t < rows.size(); t++) {
        const int r = rows[t], c = cidx(r);
        for (int k = 0; k < 4; k++) {
          h2.x_seed[4 * (size_t)c + k] = xd[4 * (size_t)r + k];
          h2.rest[4 * (size_t)c + k] = hp.rest[4 * (size_t)r + k];
        }
        h2.uv[2 * (size_t)c] = hp.uv[2 * (size_t)r];
        h2.uv[2 * (size_t)c + 1] = hp.uv[2 * (size_t)r + 1];
        h2.pt_kf[c] = 0;
        h2.fixed0[c] = fixed[r];
        h2.rp_level0[c] = rp_lvl[r];
      }
      for (int e : keep) {
        h2.pair_i.push_back(cidx(hp.pair_i[e]));
        h2.pair_j.push_back(cidx(hp.pair_j[e]));
        h2.pair_w.push_back(hp.pair_w[e]);
        h2.pair_d0.push_back(hp.pair_d0[e]);
        h2.sp_level0.push_back(sp_lvl[e]);
      }
      h2.un_ptr.assign(Vc + 1, 0);
      for (int c = 0; c < Vc; c++) {
        const int v = c - lost0;
        h2.un_ptr[c + 1] = h2.un_ptr[c] + ((v >= 0 && v < n_lost) ? un_ptr_l[v + 1] - un_ptr_l[v] : 0);
      }
      h2.un_w = un_w_l;
      h2.un_ref.resize(n_un);
      for (int t = 0; t < n_un; t++) h2.un_ref[t] = cidx(un_ref_row[t]);
      h2.kf_begin = {0, Vc};
      h2.n_sort = 0;
      // exact-solve engine: unknowns are the free deformation rows and the lost rows (a prefix of the rows); a lost
      // point has no pixel of its own, so the dissection places it at its first reference vertex
      h2.direct = env_int("NRSLAM_B200_DIRECT", 1) != 0 && env_int("NRSLAM_B200_DIRECT_LOST", 1) != 0;
      h2.n_unknown = n_free + n_lost;
      for (int v = 0; v < n_lost; v++)
        if (un_ptr_l[v + 1] > un_ptr_l[v]) {
          const int ref = h2.un_ref[un_ptr_l[v]];
          h2.uv[2 * (size_t)(lost0 + v)] = h2.uv[2 * (size_t)ref];
          h2.uv[2 * (size_t)(lost0 + v) + 1] = h2.uv[2 * (size_t)ref + 1];
        }
      h2.ops = {OP_RESET, OP_OPTIMIZE};
      h2.op_args = {0, opt.lost_iterations};
      Staged& st2 = ctx->staged[3];
      rc = stage_problem(ctx, st2, h2);
      if (rc) return rc;
      if (stats) stats->h2d_bytes += (int64_t)st2.h2d_bytes;
      hprof.mark("lost_setup");
      rc = run_staged_begin(ctx, st2);
      if (rc) return rc;
      refresh(0);  // the rest of the graph refresh, behind the lost-point kernel
      rest_done = true;
      rc = run_staged_end(ctx, st2, stats);
      if (rc) return rc;
      hprof.mark("run2");
      const double* xl = st2.out.h<double>(st2.o_x);
      for (int v = 0; v < n_lost; v++)
        for (int k = 0; k < 3; k++)
          last_pos[3 * (size_t)lost_list[v] + k] =
              (float)xl[4 * (size_t)st2.row_of[lost0 + v] + k] + last_pos[3 * (size_t)lost_list[v] + k];  // :544-552
    }
    for (int v = 0; v < n_lost; v++)
      if (lost_vertex_out) lost_vertex_out[v] = lost_list[v];
    if (n_lost_out) *n_lost_out = n_lost;
    if (stats) stats->n_fixed_edges = n_un;
  }
  if (!rest_done) refresh(0);
  if (status_out) memcpy(status_out, status.data(), n);
  hprof.print("pose_deform");
  if (stats) stats->host_ms = (float)(wall_ms() - t0);
  return 0;
}
}  // namespace

// =====================================================================================================
// LocalDeformableBundleAdjustment — g2o_optimization.cc:880-1161
// =====================================================================================================

namespace {
// Graph construction of the BA window (vertices, reprojection edges, springs, dampers) in the reference's order.
int build_ba_problem(nrslam_b200_ctx* ctx, const nrslam_b200_options& opt, const nrslam_b200_camera* cam, int32_t F,
                     const float* kf_pose_io, int32_t O, const int32_t* obs_kf, const int32_t* obs_vertex,
                     const float* uv, const float* X_io, const nrslam_b200_graph* g, float scale, int32_t iterations,
                     HostProblem& hp) {
  if (iterations <= 0) iterations = opt.ba_iterations;
  const int M = g->n_vertices;
  const int regularizers_per_point = opt.regularizers_per_point;
  const float th2 = opt.th_huber_2dof_sq, th3 = opt.th_huber_3dof_sq;
  const float info_reprojection = 1.0f / (opt.sigma_reprojection * opt.sigma_reprojection);
  const float info_position = 1.0f / (opt.sigma_position * opt.sigma_position);
  const float sigma_spatial = (float)((double)opt.sigma_spatial_factor * scale);
  const float info_spatial = 1.0f / (sigma_spatial * sigma_spatial);
  const float min_w = graph_min_weight(g);

  hp.F = F;
  hp.V = O;
  hp.cam = to_cam(cam);
  hp.spring_kind = SPRING_BA;
  hp.info_reproj = info_reprojection;
  hp.delta_reproj = (double)std::sqrt(th2);
  hp.info_spatial = info_spatial;
  hp.delta_spatial = (double)std::sqrt(th3);
  hp.info_spring = info_position;
  hp.delta_spring = -1;  // :1057-1071 no robust kernel on the springs
  hp.spring_k = (double)opt.spring_k;
  hp.th2f = th2;
  hp.th3f = th3;
  hp.pose_seed.resize(7 * (size_t)F);
  for (int k = 0; k < F; k++) pose_from_f7(kf_pose_io + 7 * k, hp.pose_seed.data() + 7 * k);
  hp.x_seed.resize(4 * (size_t)O);
  hp.rest.assign(4 * (size_t)O, 0.0);
  hp.uv.resize(2 * (size_t)O);
  hp.pt_kf.resize(O);
  hp.kf_begin.assign(F + 1, 0);
  {
    std::atomic<int> bad{0};
    par_ranges(host_threads(64, O), (size_t)O, [&](size_t ob, size_t oe) {  // rows are independent: threads take ranges
      for (size_t o = ob; o < oe; o++) {
        if (obs_kf[o] < 0 || obs_kf[o] >= F || (o > 0 && obs_kf[o] < obs_kf[o - 1]) || obs_vertex[o] < 0 ||
            obs_vertex[o] >= M) {
          bad.store(1, std::memory_order_relaxed);
          return;
        }
        for (int k = 0; k < 3; k++) hp.x_seed[4 * o + k] = X_io[3 * o + k];
        hp.x_seed[4 * o + 3] = 0;
        hp.uv[2 * o] = uv[2 * o];
        hp.uv[2 * o + 1] = uv[2 * o + 1];
        hp.pt_kf[o] = obs_kf[o];
      }
    });
    if (bad.load()) return fail(ctx, NRSLAM_B200_ERR_ARG, "local_ba: observations must be grouped by keyframe slot");
    for (int o = 0; o < O; o++) hp.kf_begin[obs_kf[o] + 1]++;
  }
  for (int k = 0; k < F; k++) hp.kf_begin[k + 1] += hp.kf_begin[k];
  HostProf hprof;
  hprof.mark("rows");

  // ---- springs inside a keyframe, dampers to the next newer keyframe (:982-1136)
  // sorted neighbour lists are cached per map point; (pair, keyframe) de-duplication is keyed by graph edge
  // The sorted list of a map point is stored as (neighbour, edge) pairs, already cut at the first weight < min_weight
  // (GetEdges) and at the first BAD edge (the loops' break; BAD sorts last), so the scans below read sequential memory.
  // Entries of map point mp live at [rowptr[mp], rowptr[mp] + nb_cnt[mp]) of nb_pair (capacity = its degree), so the
  // observed map points are sorted independently of each other: host threads take ranges of them.
  std::vector<int> nb_cnt(M, 0);
  std::vector<int> nb_pair(2 * (size_t)g->rowptr[M]);  // 2 ints per entry: neighbour vertex, undirected edge id
  const int* nb_ptr = g->rowptr;
  {
    std::vector<char> seen(M, 0);
    for (int o = 0; o < O; o++) seen[obs_vertex[o]] = 1;
    par_ranges(host_threads(64, O), (size_t)M, [&](size_t mb, size_t me) {
      std::vector<std::pair<unsigned long long, int>> keyed;
      for (size_t mq = mb; mq < me; mq++) {
        const int mp = (int)mq;
        if (!seen[mp]) continue;
        // GetEdges order (regularization_graph.cc:61-87): status asc, weight desc, neighbour asc — one 64-bit key
        // (weights are non-negative floats: their bit patterns order like their values)
        keyed.clear();
        for (int p = g->rowptr[mp]; p < g->rowptr[mp + 1]; p++) {
          const int ge = g->eid[p];
          unsigned wb;
          const float wgt = g->weight[ge];
          memcpy(&wb, &wgt, 4);
          const unsigned long long key = ((unsigned long long)(g->status[ge] & 3u) << 62) |
                                         ((unsigned long long)(0xFFFFFFFFu - wb) << 30) |
                                         (unsigned long long)(g->col[p] & 0x3FFFFFFF);
          keyed.emplace_back(key, p);
        }
        std::sort(keyed.begin(), keyed.end());
        int* dst = nb_pair.data() + 2 * (size_t)g->rowptr[mp];
        int cnt = 0;
        for (const auto& kp : keyed) {
          const int ge = g->eid[kp.second];
          if (g->weight[ge] < min_w || g->status[ge] == NRSLAM_EDGE_BAD) break;  // GetEdges cut ; :1035-1037 break
          dst[2 * cnt] = g->col[kp.second];
          dst[2 * cnt + 1] = ge;
          cnt++;
        }
        nb_cnt[mp] = cnt;
      }
    });
  }
  hprof.mark("nb_lists");
  // Keyframes are independent here (a spring joins two points of ONE keyframe, a damper reads keyframes k and k + 1,
  // the reference's de-duplication maps are per keyframe): host threads take keyframes round robin, every keyframe
  // fills its own edge lists in the reference's order, and the lists are concatenated in keyframe order afterwards, so
  // the result is identical to the sequential loop whatever the thread count.
  struct KfEdges {
    std::vector<int> pair_i, pair_j, dmp_v;
    std::vector<double> pair_d0, dmp_w;
  };
  std::vector<KfEdges> per_kf(F);
  const int n_threads = host_threads(F, O);
  auto work = [&](int tix) {
    std::vector<int> cur(M, -1), nxt(M, -1);  // inserted_landmarks[k][mappoint] -> row
    std::vector<int> spring_stamp(g->n_edges, -1), damper_stamp(g->n_edges, -1);
    KfEdges out;  // per-thread scratch that keeps its capacity: the keyframe's lists are stored at their exact size
                  // (worst-case reserves per keyframe were 8x oversized and paid for in page faults)
    for (int k = tix; k < F; k += n_threads) {
      const bool has_next = k + 1 < F;
      out.pair_i.clear(); out.pair_j.clear(); out.pair_d0.clear(); out.dmp_v.clear(); out.dmp_w.clear();
      for (int o = hp.kf_begin[k]; o < hp.kf_begin[k + 1]; o++) cur[obs_vertex[o]] = o;
      if (has_next)
        for (int o = hp.kf_begin[k + 1]; o < hp.kf_begin[k + 2]; o++) nxt[obs_vertex[o]] = o;
      for (int o = hp.kf_begin[k]; o < hp.kf_begin[k + 1]; o++) {
        const int mp = obs_vertex[o];
        const int* ents = nb_pair.data() + 2 * (size_t)nb_ptr[mp];
        const int ne = nb_cnt[mp];
        int n_regularizers = 0;
        for (int t = 0; t < ne; t++) {
          const int other = ents[2 * t], ge = ents[2 * t + 1];
          if (n_regularizers > regularizers_per_point) break;  // :1035-1037 (BAD edges are already cut off)
          const int oidx = cur[other];
          if (oidx < 0) continue;
          if (spring_stamp[ge] == k) {  // :1049-1052 already inserted from the other endpoint
            n_regularizers++;
            continue;
          }
          spring_stamp[ge] = k;
          out.pair_i.push_back(o);
          out.pair_j.push_back(oidx);
          out.pair_d0.push_back((double)g->first_distance[ge]);
          n_regularizers++;
        }
        if (has_next) {
          const int nlidx = nxt[mp];
          if (nlidx < 0) continue;
          int n_reg2 = 0;
          for (int t = 0; t < ne; t++) {
            const int other = ents[2 * t], ge = ents[2 * t + 1];
            if (n_reg2 > regularizers_per_point) break;
            const int oidx = cur[other], noidx = nxt[other];
            if (oidx < 0 || noidx < 0) continue;
            if (damper_stamp[ge] == k) {
              n_reg2++;
              continue;
            }
            damper_stamp[ge] = k;
            out.dmp_v.push_back(o);
            out.dmp_v.push_back(oidx);
            out.dmp_v.push_back(nlidx);
            out.dmp_v.push_back(noidx);
            out.dmp_w.push_back((double)g->weight[ge]);
            n_reg2++;
          }
        }
      }
      for (int o = hp.kf_begin[k]; o < hp.kf_begin[k + 1]; o++) cur[obs_vertex[o]] = -1;
      if (has_next)
        for (int o = hp.kf_begin[k + 1]; o < hp.kf_begin[k + 2]; o++) nxt[obs_vertex[o]] = -1;
      per_kf[k] = out;
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  hprof.mark("kf_edges");
  // concatenation in keyframe order: offsets by prefix sums, the copies on the same threads
  std::vector<size_t> poff(F + 1, 0), doff(F + 1, 0);
  for (int k = 0; k < F; k++) {
    poff[k + 1] = poff[k] + per_kf[k].pair_i.size();
    doff[k + 1] = doff[k] + per_kf[k].dmp_w.size();
  }
  const size_t np = poff[F], nd = doff[F];
  hp.pair_i.resize(np); hp.pair_j.resize(np); hp.pair_d0.resize(np);
  hp.dmp_v.resize(4 * nd); hp.dmp_w.resize(nd);
  auto concat = [&](int tix) {
    for (int k = tix; k < F; k += n_threads) {
      const KfEdges& e = per_kf[k];
      std::copy(e.pair_i.begin(), e.pair_i.end(), hp.pair_i.begin() + poff[k]);
      std::copy(e.pair_j.begin(), e.pair_j.end(), hp.pair_j.begin() + poff[k]);
      std::copy(e.pair_d0.begin(), e.pair_d0.end(), hp.pair_d0.begin() + poff[k]);
      std::copy(e.dmp_v.begin(), e.dmp_v.end(), hp.dmp_v.begin() + 4 * doff[k]);
      std::copy(e.dmp_w.begin(), e.dmp_w.end(), hp.dmp_w.begin() + doff[k]);
    }
  };
  if (n_threads == 1) {
    concat(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(concat, t);
    concat(0);
    for (auto& th : pool) th.join();
  }
  hp.pair_w.assign(np, -1.0);
  hprof.mark("concat");
  hprof.print("build_ba_problem");
  hp.ops = {OP_CLEAR_LEVELS, OP_RESET, OP_OPTIMIZE};
  hp.op_args = {0, 0, iterations};
  hp.n_sort = O;

  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Landmark sharding of a BA window over `world` ranks. Every rank runs this on the FULL problem and
// derives the same partition, so no metadata is exchanged at run time:
//   - landmarks are ordered along a Morton curve of their first observed position and cut into `world` contiguous
//     ranges of equal observation count; a rank owns every per-keyframe copy of its landmarks;
//   - a rank's problem = its rows (keyframe-major, Morton order inside a keyframe) + read-only halo copies of the
//     rows of other ranks that its springs / dampers touch; an edge shared by two ranks is linearised by both and
//     its chi2 is counted by the owner of its first endpoint;
//   - push lists tell the owner of a row which halo copies to refresh (rank << 26 | row index on that rank).
// ---------------------------------------------------------------------------------------------------
struct ShardPlan {
  std::vector<int> row_of;           // caller observation -> row of the globally sorted problem
  std::vector<int> owner;            // [O] owner rank of a sorted row
  std::vector<int> loc_of;           // [O] index of a sorted row among its owner's rows
  std::vector<int> n_own;            // [world]
  std::vector<std::vector<int>> halo;  // [world] sorted rows (global sorted index) held as halo copies
  std::vector<int> xp_ptr, xp_dst;   // push lists of `rank`
  std::vector<unsigned char> pair_cnt, dmp_cnt;
};

int shard_problem(HostProblem& G, const int32_t* obs_vertex, int n_vertices, int rank, int world, HostProblem& L,
                  ShardPlan& sp) {
  const int O = G.V, F = G.F;
  G.n_sort = O;
  sort_rows(G, sp.row_of);
  std::vector<int> lm_of_row(O);
  for (int o = 0; o < O; o++) lm_of_row[sp.row_of[o]] = obs_vertex[o];
  // ---- landmark order: Morton code of the first (oldest keyframe) observed position, ties by id
  std::vector<int> first_row(n_vertices, -1), n_obs(n_vertices, 0);
  for (int r = 0; r < O; r++) {
    const int m = lm_of_row[r];
    if (first_row[m] < 0) first_row[m] = r;
    n_obs[m]++;
  }
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int m = 0; m < n_vertices; m++) {
    if (first_row[m] < 0) continue;
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], G.x_seed[4 * (size_t)first_row[m] + a]);
      hi[a] = std::max(hi[a], G.x_seed[4 * (size_t)first_row[m] + a]);
    }
  }
  double ext = 0;
  for (int a = 0; a < 3; a++) ext = std::max(ext, hi[a] - lo[a]);
  const double inv = ext > 0 ? 1.0 / ext : 0.0;
  const double inv_ext[3] = {inv, inv, inv};
  std::vector<std::pair<uint32_t, int>> keyed;
  for (int m = 0; m < n_vertices; m++)
    if (first_row[m] >= 0) keyed.emplace_back(morton3(&G.x_seed[4 * (size_t)first_row[m]], lo, inv_ext), m);
  std::sort(keyed.begin(), keyed.end());
  std::vector<int> lm_owner(n_vertices, -1);
  {
    long long seen = 0;
    for (auto& km : keyed) {
      // the landmark goes to the rank whose range contains the midpoint of its observations
      const long long mid = 2 * seen + n_obs[km.second];
      int r = (int)((mid * world) / (2 * (long long)O));
      lm_owner[km.second] = std::min(std::max(r, 0), world - 1);
      seen += n_obs[km.second];
    }
  }
  sp.owner.resize(O);
  sp.loc_of.resize(O);
  sp.n_own.assign(world, 0);
  for (int r = 0; r < O; r++) {
    sp.owner[r] = lm_owner[lm_of_row[r]];
    sp.loc_of[r] = sp.n_own[sp.owner[r]]++;
  }
  // ---- halo sets of every rank
  sp.halo.assign(world, {});
  const int P = (int)G.pair_i.size(), D = (int)G.dmp_w.size();
  for (int e = 0; e < P; e++) {
    const int i = G.pair_i[e], j = G.pair_j[e];
    if (sp.owner[i] != sp.owner[j]) {
      sp.halo[sp.owner[i]].push_back(j);
      sp.halo[sp.owner[j]].push_back(i);
    }
  }
  for (int e = 0; e < D; e++) {
    const int* v = &G.dmp_v[4 * (size_t)e];
    const int oa = sp.owner[v[0]], ob = sp.owner[v[1]];
    if (oa != ob) {
      sp.halo[oa].push_back(v[1]);
      sp.halo[oa].push_back(v[3]);
      sp.halo[ob].push_back(v[0]);
      sp.halo[ob].push_back(v[2]);
    }
  }
  for (auto& h : sp.halo) {
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
  }
  // ---- this rank's problem
  const int n_own = sp.n_own[rank], n_halo = (int)sp.halo[rank].size();
  auto local_of = [&](int row) {
    if (sp.owner[row] == rank) return sp.loc_of[row];
    const auto& h = sp.halo[rank];
    return n_own + (int)(std::lower_bound(h.begin(), h.end(), row) - h.begin());
  };
  L = HostProblem();
  L.F = F;
  L.V = n_own + n_halo;
  L.n_halo = n_halo;
  L.sharded = true;
  L.poses_fixed = G.poses_fixed;
  L.points_fixed = G.points_fixed;
  L.spring_kind = G.spring_kind;
  L.cam = G.cam;
  L.info_reproj = G.info_reproj; L.delta_reproj = G.delta_reproj;
  L.info_spatial = G.info_spatial; L.delta_spatial = G.delta_spatial;
  L.info_spring = G.info_spring; L.delta_spring = G.delta_spring; L.spring_k = G.spring_k;
  L.th2f = G.th2f; L.th3f = G.th3f;
  L.pose_seed = G.pose_seed;
  L.ops = G.ops;
  L.op_args = G.op_args;
  L.n_sort = 0;
  L.x_seed.assign(4 * (size_t)L.V, 0.0);
  L.rest.assign(4 * (size_t)L.V, 0.0);
  L.uv.assign(2 * (size_t)L.V, 0.0);
  L.pt_kf.assign(L.V, -1);
  L.kf_begin.assign(F + 1, 0);
  for (int r = 0; r < O; r++) {
    if (sp.owner[r] != rank) continue;
    const int li = sp.loc_of[r];
    for (int a = 0; a < 4; a++) L.x_seed[4 * (size_t)li + a] = G.x_seed[4 * (size_t)r + a];
    L.uv[2 * (size_t)li] = G.uv[2 * (size_t)r];
    L.uv[2 * (size_t)li + 1] = G.uv[2 * (size_t)r + 1];
    L.pt_kf[li] = G.pt_kf[r];
    L.kf_begin[G.pt_kf[r] + 1]++;
  }
  for (int k = 0; k < F; k++) L.kf_begin[k + 1] += L.kf_begin[k];
  for (int h = 0; h < n_halo; h++)
    for (int a = 0; a < 4; a++) L.x_seed[4 * (size_t)(n_own + h) + a] = G.x_seed[4 * (size_t)sp.halo[rank][h] + a];
  for (int e = 0; e < P; e++) {
    const int i = G.pair_i[e], j = G.pair_j[e];
    if (sp.owner[i] != rank && sp.owner[j] != rank) continue;
    L.pair_i.push_back(local_of(i));
    L.pair_j.push_back(local_of(j));
    L.pair_w.push_back(G.pair_w[e]);
    L.pair_d0.push_back(G.pair_d0[e]);
    sp.pair_cnt.push_back(sp.owner[i] == rank ? 1 : 0);
  }
  for (int e = 0; e < D; e++) {
    const int* v = &G.dmp_v[4 * (size_t)e];
    if (sp.owner[v[0]] != rank && sp.owner[v[1]] != rank) continue;
    for (int a = 0; a < 4; a++) L.dmp_v.push_back(local_of(v[a]));
    L.dmp_w.push_back(G.dmp_w[e]);
    sp.dmp_cnt.push_back(sp.owner[v[0]] == rank ? 1 : 0);
  }
  // ---- push lists: for every rank r that holds one of my rows as a halo copy
  std::vector<std::vector<int>> dst(n_own);
  for (int r = 0; r < world; r++) {
    if (r == rank) continue;
    const auto& h = sp.halo[r];
    for (int s = 0; s < (int)h.size(); s++)
      if (sp.owner[h[s]] == rank) dst[sp.loc_of[h[s]]].push_back((r << 26) | (sp.n_own[r] + s));
  }
  sp.xp_ptr.assign(L.V + 1, 0);
  sp.xp_dst.clear();
  for (int i = 0; i < n_own; i++) {
    for (int d : dst[i]) sp.xp_dst.push_back(d);
    sp.xp_ptr[i + 1] = (int)sp.xp_dst.size();
  }
  for (int i = n_own; i < L.V; i++) sp.xp_ptr[i + 1] = sp.xp_ptr[n_own];
  return 0;
}
}  // namespace

extern "C" {

int nrslam_b200_shard_partition(const nrslam_b200_options* opt_in, int32_t world, int32_t F, const float* kf_pose,
                                int32_t O, const int32_t* obs_kf, const int32_t* obs_vertex, const float* uv,
                                const float* X, const nrslam_b200_graph* g, float scale, int32_t* owner_out,
                                int32_t* n_own_out, int32_t* n_halo_out, int32_t* n_push_out,
                                int32_t* n_edges_out) {
  if (!kf_pose || !obs_kf || !obs_vertex || !uv || !X || !g || !owner_out || world < 1 || world > nrs::kMaxWorld ||
      F < 1 || O < 1)
    return NRSLAM_B200_ERR_ARG;
  nrslam_b200_options opt;
  if (opt_in) opt = *opt_in; else nrslam_b200_default_options(&opt);
  nrslam_b200_camera cam;
  memset(&cam, 0, sizeof(cam));
  for (int r = 0; r < world; r++) {
    HostProblem G, L;
    const int rc = build_ba_problem(nullptr, opt, &cam, F, kf_pose, O, obs_kf, obs_vertex, uv, X, g, scale, 0, G);
    if (rc) return rc;
    ShardPlan sp;
    shard_problem(G, obs_vertex, g->n_vertices, r, world, L, sp);
    if (r == 0)
      for (int o = 0; o < O; o++) owner_out[o] = sp.owner[sp.row_of[o]];
    if (n_own_out) n_own_out[r] = sp.n_own[r];
    if (n_halo_out) n_halo_out[r] = L.n_halo;
    if (n_push_out) n_push_out[r] = (int)sp.xp_dst.size();
    if (n_edges_out) {
      n_edges_out[3 * r] = (int)L.pair_i.size();
      n_edges_out[3 * r + 1] = (int)L.dmp_w.size();
      int cnt = 0;
      for (unsigned char c : sp.pair_cnt) cnt += c;
      for (unsigned char c : sp.dmp_cnt) cnt += c;
      n_edges_out[3 * r + 2] = cnt;  // edges whose chi2 this rank counts
    }
  }
  return 0;
}

int nrslam_b200_shard_init(nrslam_b200_ctx* ctx, int32_t rank, int32_t world, int32_t max_rows, int32_t max_poses,
                           unsigned char* handle_out) {
  if (!ctx || !handle_out || world < 2 || world > nrs::kMaxWorld || rank < 0 || rank >= world || max_rows < 1 ||
      max_poses < 1)
    return fail(ctx, NRSLAM_B200_ERR_ARG, "shard_init: bad argument");
  Shard& sh = ctx->shard;
  if (sh.local) return fail(ctx, NRSLAM_B200_ERR_ARG, "shard_init: already initialised");
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  sh.rank = rank;
  sh.world = world;
  sh.max_rows = max_rows;
  sh.max_poses = max_poses;
  sh.xstride = 8 + 27 * max_poses;
  sh.off_abort = 8 * nrs::kMaxWorld;
  sh.off_red = 256;
  sh.off_z = sh.off_red + sizeof(double) * 2 * (size_t)world * sh.xstride;
  sh.off_z = (sh.off_z + 255) & ~size_t(255);
  sh.off_x = sh.off_z + sizeof(double) * 4 * (size_t)max_rows;
  sh.bytes = sh.off_x + sizeof(double) * 4 * (size_t)max_rows;
  NRS_CUDA(ctx, cudaMalloc(&sh.local, sh.bytes));
  NRS_CUDA(ctx, cudaMemset(sh.local, 0, sh.bytes));
  NRS_CUDA(ctx, cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  NRS_CUDA(ctx, cudaIpcGetMemHandle(&h, sh.local));
  static_assert(sizeof(h) == NRSLAM_B200_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

int nrslam_b200_shard_attach(nrslam_b200_ctx* ctx, const unsigned char* handles) {
  if (!ctx || !handles || !ctx->shard.local) return fail(ctx, NRSLAM_B200_ERR_ARG, "shard_attach: bad argument");
  Shard& sh = ctx->shard;
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < sh.world; r++) {
    if (r == sh.rank) {
      sh.peer[r] = sh.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
    NRS_CUDA(ctx, cudaIpcOpenMemHandle(&sh.peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  sh.attached = true;
  sh.broken = false;
  sh.epoch = 0;
  return 0;
}

int nrslam_b200_local_ba_sharded(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t F, float* kf_pose_io,
                                 int32_t O, const int32_t* obs_kf, const int32_t* obs_vertex, const float* uv,
                                 float* X_io, const nrslam_b200_graph* g, float scale, int32_t iterations,
                                 int32_t* owner_out, nrslam_b200_stats* stats) {
  if (!ctx || !cam || !kf_pose_io || !obs_kf || !obs_vertex || !uv || !X_io || !g || O < 0)
    return fail(ctx, NRSLAM_B200_ERR_ARG, "local_ba_sharded: bad argument");
  Shard& sh = ctx->shard;
  if (!sh.attached) return fail(ctx, NRSLAM_B200_ERR_ARG, "local_ba_sharded: call shard_init / shard_attach first");
  if (sh.broken) return fail(ctx, NRSLAM_B200_ERR_CUDA, "local_ba_sharded: an earlier exchange timed out; re-create the context");
  const double t0 = wall_ms();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (F < 3) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "local_ba: fewer than 3 keyframes");
  if (O == 0) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "local_ba: no observations");
  if (F > sh.max_poses) return fail(ctx, NRSLAM_B200_ERR_ARG, "local_ba_sharded: more keyframes than shard_init reserved");
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  HostProblem G, hp;
  {
    const int brc = build_ba_problem(ctx, ctx->opt, cam, F, kf_pose_io, O, obs_kf, obs_vertex, uv, X_io, g, scale,
                                     iterations, G);
    if (brc) return brc;
  }
  ShardPlan sp;
  shard_problem(G, obs_vertex, g->n_vertices, sh.rank, sh.world, hp, sp);
  const int n_own = sp.n_own[sh.rank];
  // The call is a collective: every rank derives the same partition, so the size checks run over ALL ranks and every
  // rank leaves with the same code before anybody launches (a lone early exit would leave the peers spinning in the
  // exchange until its timeout and mark their shard state broken).
  for (int r = 0; r < sh.world; r++) {
    if (sp.n_own[r] + (int)sp.halo[r].size() > sh.max_rows)
      return fail(ctx, NRSLAM_B200_ERR_ARG, "local_ba_sharded: more rows than shard_init reserved (on some rank)");
    if (sp.n_own[r] == 0) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "local_ba_sharded: a rank owns no observation");
  }
  Staged& st = ctx->staged[2];
  int rc = stage_problem(ctx, st, hp);
  if (rc) return rc;
  // ---- this rank's exchange state
  Arena& xin = st.xin;
  const size_t xneed = (sp.xp_ptr.size() + sp.xp_dst.size() + 2) * 4 + sp.pair_cnt.size() + sp.dmp_cnt.size() + 8 * 256;
  if (!xin.reserve(xneed, true)) return fail(ctx, NRSLAM_B200_ERR_ALLOC, "exchange arena allocation failed");
  Params& p = st.p;
  p.xp_ptr = xin.d<int>(put(xin, sp.xp_ptr));
  p.xp_dst = xin.d<int>(put(xin, sp.xp_dst));
  p.pair_cnt = xin.d<unsigned char>(put(xin, sp.pair_cnt));
  p.dmp_cnt = xin.d<unsigned char>(put(xin, sp.dmp_cnt));
  p.world = sh.world;
  p.xfused = env_int("NRSLAM_B200_XFUSED", 1);
  p.rank = sh.rank;
  p.xstride = sh.xstride;
  p.xepoch0 = sh.epoch;
  p.xtimeout_ns = (unsigned long long)env_int("NRSLAM_B200_XTIMEOUT_MS", 20000) * 1000000ULL;
  for (int r = 0; r < sh.world; r++) {
    char* base = static_cast<char*>(sh.peer[r]);
    p.xflag[r] = reinterpret_cast<unsigned long long*>(base);
    p.xred[r] = reinterpret_cast<double*>(base + sh.off_red);
    p.xz[r] = reinterpret_cast<double*>(base + sh.off_z);
    p.xx[r] = reinterpret_cast<double*>(base + sh.off_x);
  }
  char* mine = static_cast<char*>(sh.local);
  p.xabort = reinterpret_cast<int*>(mine + sh.off_abort);
  p.zvec = p.xz[sh.rank];
  p.x = p.xx[sh.rank];
  NRS_CUDA(ctx, cudaMemcpyAsync(xin.dev(), xin.host(), xin.used(), cudaMemcpyHostToDevice, ctx->stream));
  NRS_CUDA(ctx, cudaMemsetAsync(p.xabort, 0, sizeof(int), ctx->stream));
  // halo estimates start at their seeds (the owners push every later change); the kernel's RESET seeds the own rows
  if (hp.n_halo > 0)
    NRS_CUDA(ctx, cudaMemcpyAsync(p.x + 4 * (size_t)n_own, p.x_seed + 4 * (size_t)n_own,
                                  sizeof(double) * 4 * (size_t)hp.n_halo, cudaMemcpyDeviceToDevice, ctx->stream));
  const double t1 = wall_ms();
  if (stats) stats->h2d_bytes += (int64_t)st.h2d_bytes + (int64_t)xin.used();
  rc = run_staged(ctx, st, stats);
  st.valid = false;  // the staged program is bound to this call's exchange epoch: no resolve()
  const EngineStats* es = st.out.h<EngineStats>(st.o_stats);
  if (rc) {
    sh.broken = true;
    return rc;
  }
  sh.epoch += (unsigned long long)es->xepochs;
  if (es->xfail) {
    sh.broken = true;
    return fail(ctx, NRSLAM_B200_ERR_CUDA, "local_ba_sharded: a peer rank did not arrive at an exchange (timeout)");
  }
  std::vector<double> xd(4 * (size_t)n_own);
  NRS_CUDA(ctx, cudaMemcpy(xd.data(), p.x, sizeof(double) * 4 * (size_t)n_own, cudaMemcpyDeviceToHost));
  const double* pose = st.out.h<double>(st.o_pose);
  for (int i = 0; i < 7 * F; i++)
    if (!std::isfinite(pose[i])) return fail(ctx, NRSLAM_B200_NUM_NONFINITE, "local_ba: non-finite pose");
  for (int k = 0; k < F; k++) pose_to_f7(pose + 7 * k, kf_pose_io + 7 * k);
  for (int o = 0; o < O; o++) {
    const int row = sp.row_of[o];
    if (owner_out) owner_out[o] = sp.owner[row];
    if (sp.owner[row] != sh.rank) continue;  // rows of other ranks stay untouched: the caller gathers them
    for (int k = 0; k < 3; k++) X_io[3 * (size_t)o + k] = (float)xd[4 * (size_t)sp.loc_of[row] + k];
  }
  if (stats) {
    stats->n_reproj_edges = n_own;
    stats->n_spring_edges = (int)hp.pair_i.size();
    stats->n_damper_edges = (int)hp.dmp_w.size();
    stats->n_points = n_own;
    stats->n_poses = F;
    stats->d2h_bytes += (int64_t)(sizeof(double) * 4 * (size_t)n_own);
    stats->stage_ms = (float)(t1 - t0);
    stats->host_ms = (float)(wall_ms() - t0);
  }
  return 0;
}

int nrslam_b200_local_ba(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, int32_t F, float* kf_pose_io,
                         int32_t O, const int32_t* obs_kf, const int32_t* obs_vertex, const float* uv,
                         float* X_io, const nrslam_b200_graph* g, float scale, int32_t iterations,
                         nrslam_b200_stats* stats) {
  if (!ctx || !cam || !kf_pose_io || !obs_kf || !obs_vertex || !uv || !X_io || !g || O < 0)
    return fail(ctx, NRSLAM_B200_ERR_ARG, "local_ba: bad argument");
  const double t0 = wall_ms();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (F < 3) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "local_ba: fewer than 3 keyframes");  // :922-924
  if (O == 0) return fail(ctx, NRSLAM_B200_NUM_TOO_FEW, "local_ba: no observations");
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  HostProf hprof;
  HostProblem hp;
  {
    const int brc = build_ba_problem(ctx, ctx->opt, cam, F, kf_pose_io, O, obs_kf, obs_vertex, uv, X_io, g, scale,
                                     iterations, hp);
    if (brc) return brc;
  }
  hprof.mark("build_graph");
  Staged& st = ctx->staged[2];
  int rc = stage_problem(ctx, st, hp);
  if (rc) return rc;
  hprof.mark("stage");
  const double t1 = wall_ms();
  if (stats) stats->h2d_bytes += (int64_t)st.h2d_bytes;
  rc = run_staged(ctx, st, stats);
  if (rc) return rc;
  hprof.mark("run");
  hprof.print("local_ba");
  const double* pose = st.out.h<double>(st.o_pose);
  const double* xd = st.out.h<double>(st.o_x);
  for (int i = 0; i < 7 * F; i++)
    if (!std::isfinite(pose[i])) return fail(ctx, NRSLAM_B200_NUM_NONFINITE, "local_ba: non-finite pose");
  for (int k = 0; k < F; k++) pose_to_f7(pose + 7 * k, kf_pose_io + 7 * k);  // :1146-1151
  par_ranges(host_threads(64, O), (size_t)O, [&](size_t ob, size_t oe) {
    for (size_t o = ob; o < oe; o++)
      for (int k = 0; k < 3; k++) X_io[3 * o + k] = (float)xd[4 * (size_t)st.row_of[o] + k];  // :1153-1160
  });
  if (stats) {
    stats->n_reproj_edges = O;
    stats->n_spring_edges = (int)hp.pair_i.size();
    stats->n_damper_edges = (int)hp.dmp_w.size();
    stats->n_points = O;
    stats->n_poses = F;
    stats->stage_ms = (float)(t1 - t0);
    stats->host_ms = (float)(wall_ms() - t0);
  }
  return 0;
}

int nrslam_b200_resolve(nrslam_b200_ctx* ctx, int32_t which, nrslam_b200_stats* stats) {
  if (!ctx || which < 0 || which > 3) return fail(ctx, NRSLAM_B200_ERR_ARG, "resolve: bad argument");
  NRS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (stats) memset(stats, 0, sizeof(*stats));
  const double t0 = wall_ms();
  const int rc = run_staged(ctx, ctx->staged[which], stats, /*copy_back=*/false);
  if (stats) stats->host_ms = (float)(wall_ms() - t0);
  return rc;
}

}  // extern "C"
