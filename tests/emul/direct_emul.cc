// TEST INFRASTRUCTURE — never part of libnrslam_b200.so and never reachable from the product path.
// Compiles the multifrontal LL^T of nr-slam_b200/csrc/nrs_direct_core.cuh for the host with ONE emulated thread per
// CTA and runs the CTAs of a level one after the other where the kernel has a grid barrier. Together with the symbolic
// analysis (nrs_direct_plan.h, plain C++) this executes the same arithmetic as the kernel, so the CPU suite can check
// plan + numeric code against a dense solve without a GPU (tests/test_direct_emulation.py).
#define NRS_DIRECT_HOST_EMULATION 1
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../nr-slam_b200/csrc/nrs_direct_core.cuh"
#include "../../nr-slam_b200/csrc/nrs_direct_plan.h"

// Inputs in CALLER row order; the function permutes them like the product's staging does.
//   dg [V][6], cpl [V][18], b [V][3], pairs (pair_i, pair_j, pc [P][4]), hpp [27], lambda.
// Outputs: delta [V][3] (caller order), dpose [6], stats[0..7] = depth, G, p_total, u_total, smem_doubles, max_path,
// max front scalars, fail.
extern "C" int direct_emul_solve(int32_t V, const double* uv, int32_t P, const int32_t* pair_i, const int32_t* pair_j,
                                 const double* dg, const double* cpl, const double* b, const double* pc,
                                 const double* hpp, double lambda, int32_t depth, double* delta, double* dpose,
                                 int64_t* stats) {
  using namespace nrs;
  std::vector<int> pi(pair_i, pair_i + P), pj(pair_j, pair_j + P);
  DirectPlanHost pl;
  if (depth < 0) depth = direct_depth(V, 128);
  build_direct_plan(V, uv, pi, pj, depth, pl);
  std::vector<int> new_of_old(V);
  for (int n = 0; n < V; n++) new_of_old[pl.old_of_new[n]] = n;
  for (auto& v : pi) v = new_of_old[v];
  for (auto& v : pj) v = new_of_old[v];
  // incidence lists exactly like stage_problem (nrs_api.cu)
  std::vector<int> inc_ptr(V + 1, 0), inc_other(2 * (size_t)P), inc_ent(2 * (size_t)P), inc_row(2 * (size_t)P);
  for (int e = 0; e < P; e++) {
    inc_ptr[pi[e] + 1]++;
    inc_ptr[pj[e] + 1]++;
  }
  for (int i = 0; i < V; i++) inc_ptr[i + 1] += inc_ptr[i];
  {
    std::vector<int> w(inc_ptr.begin(), inc_ptr.end() - 1);
    for (int e = 0; e < P; e++) {
      int a = w[pi[e]]++;
      inc_other[a] = pj[e];
      inc_ent[a] = 2 * e;
      inc_row[a] = pi[e];
      a = w[pj[e]]++;
      inc_other[a] = pi[e];
      inc_ent[a] = 2 * e + 1;
      inc_row[a] = pj[e];
    }
  }
  std::vector<int> inc_pos;
  direct_inc_pos(pl, inc_ptr, inc_other, inc_pos);
  std::vector<double> dg8(8 * (size_t)V, 0.0), cplp(18 * (size_t)V), b4(4 * (size_t)V, 0.0);
  for (int n = 0; n < V; n++) {
    const int o = pl.old_of_new[n];
    for (int k = 0; k < 6; k++) dg8[8 * (size_t)n + k] = dg[6 * (size_t)o + k];
    for (int k = 0; k < 18; k++) cplp[18 * (size_t)n + k] = cpl[18 * (size_t)o + k];
    for (int k = 0; k < 3; k++) b4[4 * (size_t)n + k] = b[3 * (size_t)o + k];
  }
  std::vector<double> panel((size_t)pl.p_total + 1, 0.0), upd((size_t)pl.u_total + 1, 0.0);
  int fail = 0;
  direct::Plan dp;
  dp.np = pl.np; dp.V = V; dp.depth = pl.depth; dp.G = pl.G; dp.max_path = pl.max_path;
  dp.vb = pl.vb.data(); dp.nv = pl.nv.data(); dp.nbv = pl.nbv.data(); dp.bnd_ptr = pl.bnd_ptr.data();
  dp.bnd = pl.bnd.data(); dp.bpath = pl.bpath.data(); dp.path_off = pl.path_off.data();
  dp.inv_ptr = pl.inv_ptr.data(); dp.inv = pl.inv.data(); dp.p_off = pl.p_off.data(); dp.u_off = pl.u_off.data();
  dp.panel = panel.data(); dp.upd = upd.data(); dp.fail = &fail;
  direct::Sys sys;
  sys.dg = dg8.data(); sys.cpl = cplp.data(); sys.bvec = b4.data(); sys.pc = pc; sys.hpp = hpp;
  sys.inc_ptr = inc_ptr.data(); sys.inc_ent = inc_ent.data(); sys.inc_pos = inc_pos.data(); sys.inc_row = inc_row.data(); sys.lambda = lambda;
  int max_ns = 0;
  for (int t = 1; t <= pl.n_nodes; t++) max_ns = std::max(max_ns, 3 * pl.nv[t]);
  std::vector<double> sp(pl.smem_doubles + 16), sw(2 * (size_t)max_ns + 16), sv((size_t)pl.max_rows * 3 * (3 * direct::kPanel + 1) + 8), spath(pl.max_path + 8), sz(2 * (size_t)max_ns + 8);
  const direct::Thr th{0, 1};
  for (int d = pl.depth; d >= 0; d--) {
    for (int g = 0; g < pl.G; g++) direct::stage_ab(dp, sys, g, d, sp.data(), sw.data(), sv.data(), th);
    if (d > 0)
      for (int g = 0; g < pl.G; g++) direct::stage_c(dp, g, d, sp.data(), th);
  }
  std::vector<double> d4(4 * (size_t)V, 0.0);
  double dps[6] = {0, 0, 0, 0, 0, 0};
  std::vector<double> ref4;
  for (int g = 0; g < pl.G; g++) {
    std::vector<double> dg4(4 * (size_t)V, 0.0);
    for (int d = 0; d <= pl.depth; d++)
      direct::backward_front(dp, g, d, spath.data(), sp.data(), sz.data(), d4.data(), dps, th);
  }
  for (int n = 0; n < V; n++)
    for (int k = 0; k < 3; k++) delta[3 * (size_t)pl.old_of_new[n] + k] = d4[4 * (size_t)n + k];
  for (int k = 0; k < 6; k++) dpose[k] = dps[k];
  if (stats) {
    int maxf = 0;
    for (int t = 1; t <= pl.n_nodes; t++) maxf = std::max(maxf, 3 * (pl.nv[t] + pl.nbv[t]));
    stats[0] = pl.depth; stats[1] = pl.G; stats[2] = pl.p_total; stats[3] = pl.u_total;
    stats[4] = (int64_t)pl.smem_doubles; stats[5] = pl.max_path; stats[6] = maxf; stats[7] = fail;
  }
  return fail ? 1 : 0;
}

// Symbolic analysis only: stats[0..7] = depth, G, p_total, u_total, smem_doubles, max_path, root separator vertices
// (pose pseudo-vertices included), largest separator below the root.
extern "C" int direct_emul_plan(int32_t V, const double* uv, int32_t P, const int32_t* pair_i, const int32_t* pair_j,
                                int32_t depth, int64_t* stats) {
  using namespace nrs;
  std::vector<int> pi(pair_i, pair_i + P), pj(pair_j, pair_j + P);
  DirectPlanHost pl;
  if (depth < 0) depth = direct_depth(V, 128);
  build_direct_plan(V, uv, pi, pj, depth, pl);
  int below = 0;
  for (int t = 2; t <= pl.n_nodes; t++) below = std::max(below, pl.nv[t]);
  stats[0] = pl.depth; stats[1] = pl.G; stats[2] = pl.p_total; stats[3] = pl.u_total;
  stats[4] = (int64_t)pl.smem_doubles; stats[5] = pl.max_path; stats[6] = pl.nv[1]; stats[7] = below;
  return 0;
}
