// nrs_direct_plan.h — host-side symbolic analysis of the exact sparse factorisation used by the tracking solve
// (nrs_direct.cu). Pure C++ (no CUDA): also compiled into the CPU test harness (tests/emul/direct_emul.cc).
//
// What it replaces: the symbolic half of Eigen::SimplicialLLT + block-AMD ordering behind g2o's LinearSolverEigen
//   third_party/g2o/g2o/solvers/eigen/linear_solver_eigen.h:92-188 (analyzePattern / computeSymbols),
//   third_party/g2o/g2o/core/block_solver.hpp:329-341 (solve of the un-marginalised system).
// The reference orders with approximate minimum degree and factorises column by column on one core. Here the
// deformation vertices of a frame live on a 2-D sheet seen by the camera, so a GEOMETRIC nested dissection of their
// pixel coordinates gives a balanced elimination tree whose independent subtrees are handed to different SMs:
//
//   * the vertex set is bisected recursively (median cut along the wider pixel axis) `depth` times; the vertices of one
//     side that touch the other side are covered greedily (max cut-degree first) and become the SEPARATOR of that tree
//     node; leaves keep what is left. Tree nodes are heap-numbered (root 1, children 2t / 2t+1), every leaf sits at
//     depth `depth`, so level d has 2^d nodes and 2^depth CTAs can walk the tree in lock-step (a team of
//     2^(depth-d) CTAs per node of level d);
//   * rows are renumbered in post-order (left subtree, right subtree, own vertices), which IS the elimination order;
//     the 6 pose unknowns are two 3-dof pseudo-vertices V, V+1 owned by the root and eliminated last, and the
//     right-hand side rides along as pseudo-vertex V+2 that is never eliminated (augmented-matrix forward solve);
//   * a node's FRONT is [its own vertices ; its boundary], the boundary being every later-eliminated vertex an own
//     vertex or a descendant is coupled to (always including pose and rhs). Fronts are dense (relaxed supernodes).
//
// All index lists are in vertex (3x3 block) units.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace nrs {

struct DirectPlanHost {
  int V = 0;        // point rows (deformation vertices)
  int depth = 0;    // leaves at this depth
  int n_nodes = 0;  // 2^(depth+1) - 1, nodes are 1..n_nodes
  int G = 1;        // CTAs = 2^depth
  std::vector<int> old_of_new;  // [V] elimination order: new row -> caller row
  // per node t (index 0 unused)
  std::vector<int> vb, nv, nbv;     // first own vertex, own count (root: + 2 pose), boundary count (incl. pose, rhs)
  std::vector<int> bnd_ptr, bnd;    // boundary vertex ids, ascending (V, V+1 = pose; V+2 = rhs)
  std::vector<int> bpath;           // parallel to bnd: index (scalars) into the root-to-node path vector, -1 for rhs
  std::vector<int> path_off;        // per node: scalars owned by its strict ancestors (root first)
  std::vector<int> inv_ptr, inv;    // per node t >= 2: [nv[p] + nbv[p]] parent front position -> own boundary position / -1
  std::vector<long long> p_off, u_off;  // offsets (doubles) of the node's panel / update matrix
  long long p_total = 0, u_total = 0;
  int max_path = 0;           // scalars of the longest root-to-leaf path (own vertices)
  size_t smem_doubles = 0;    // shared-memory doubles the numeric kernel needs for this plan
  std::vector<int> owner;     // [V] tree node owning a row
};

namespace direct_detail {

struct Builder {
  const double* uv;
  const std::vector<std::vector<int>>* adj;  // by caller row
  int depth;
  std::vector<std::vector<int>> own;  // per node: caller rows
  std::vector<int> side;              // scratch: 0 none, 1 left, 2 right (by caller row)

  void split(int t, int d, std::vector<int>& ids) {
    if (d == depth) {
      own[t] = ids;
      return;
    }
    std::vector<int> left, right;
    if (ids.size() >= 2) {
      double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
      for (int i : ids)
        for (int a = 0; a < 2; a++) {
          lo[a] = std::min(lo[a], uv[2 * (size_t)i + a]);
          hi[a] = std::max(hi[a], uv[2 * (size_t)i + a]);
        }
      const int ax = (hi[0] - lo[0] >= hi[1] - lo[1]) ? 0 : 1;
      // median cut, ties by caller row: deterministic
      std::vector<int> order(ids);
      const size_t half = order.size() / 2;
      std::nth_element(order.begin(), order.begin() + half, order.end(), [&](int a, int b) {
        const double ua = uv[2 * (size_t)a + ax], ub = uv[2 * (size_t)b + ax];
        return ua < ub || (ua == ub && a < b);
      });
      left.assign(order.begin(), order.begin() + half);
      right.assign(order.begin() + half, order.end());
      std::sort(left.begin(), left.end());
      std::sort(right.begin(), right.end());
      for (int i : left) side[i] = 1;
      for (int i : right) side[i] = 2;
      // cut vertices and their cut degree
      std::vector<int> cutv;
      std::vector<int> deg;
      for (int i : ids) {
        int dg = 0;
        for (int o : (*adj)[i])
          if (side[o] && side[o] != side[i]) dg++;
        if (dg) {
          cutv.push_back(i);
          deg.push_back(dg);
        }
      }
      // greedy vertex cover of the cut edges: highest remaining cut degree first (ties: lower row)
      std::vector<int>& sep = own[t];
      for (;;) {
        int best = -1;
        for (size_t k = 0; k < cutv.size(); k++)
          if (deg[k] > 0 && (best < 0 || deg[k] > deg[best])) best = (int)k;
        if (best < 0) break;
        const int v = cutv[best];
        sep.push_back(v);
        for (int o : (*adj)[v]) {
          if (!side[o] || side[o] == side[v]) continue;
          // edge (v, o) is covered: o loses one cut degree
          for (size_t k = 0; k < cutv.size(); k++)
            if (cutv[k] == o) {
              if (deg[k] > 0) deg[k]--;
              break;
            }
        }
        deg[best] = 0;
        side[v] = 0;  // removed: its remaining edges no longer cross
      }
      std::sort(sep.begin(), sep.end());
      std::vector<int> l2, r2;
      for (int i : left)
        if (side[i] == 1) l2.push_back(i);
      for (int i : right)
        if (side[i] == 2) r2.push_back(i);
      for (int i : ids) side[i] = 0;
      left.swap(l2);
      right.swap(r2);
    } else {
      left = ids;  // 0 or 1 vertices: pushed down the left spine
    }
    split(2 * t, d + 1, left);
    split(2 * t + 1, d + 1, right);
  }
};

}  // namespace direct_detail

// Depth of the dissection for V rows on at most max_ctas CTAs: leaves of >= ~8 vertices, at most 2^7 leaves.
inline int direct_depth(int V, int max_ctas) {
  int d = 0;
  while (d < 7 && (2 << d) <= max_ctas && (V >> (d + 1)) >= 8) d++;
  return d;
}

// Builds the plan. uv: [2V] pixel coordinates (caller rows); pair_i / pair_j: regulariser pairs (caller rows).
// After the call the caller permutes its rows with plan.old_of_new and then calls direct_inc_pos with the new ids.
inline void build_direct_plan(int V, const double* uv, const std::vector<int>& pair_i, const std::vector<int>& pair_j,
                              int depth, DirectPlanHost& pl) {
  pl.V = V;
  pl.depth = depth;
  pl.n_nodes = (2 << depth) - 1;
  pl.G = 1 << depth;
  const int T = pl.n_nodes;
  std::vector<std::vector<int>> adj(V);
  for (size_t e = 0; e < pair_i.size(); e++) {
    adj[pair_i[e]].push_back(pair_j[e]);
    adj[pair_j[e]].push_back(pair_i[e]);
  }
  direct_detail::Builder b;
  b.uv = uv;
  b.adj = &adj;
  b.depth = depth;
  b.own.assign(T + 1, {});
  b.side.assign(V, 0);
  std::vector<int> all(V);
  std::iota(all.begin(), all.end(), 0);
  b.split(1, 0, all);

  // post-order row numbering
  pl.vb.assign(T + 1, 0);
  pl.nv.assign(T + 1, 0);
  pl.old_of_new.clear();
  pl.old_of_new.reserve(V);
  std::vector<int> new_of_old(V, -1);
  {
    // iterative post-order over the complete binary tree
    struct Item { int t; int state; };
    std::vector<Item> st;
    st.push_back({1, 0});
    while (!st.empty()) {
      Item& it = st.back();
      const int t = it.t;
      if (it.state == 0 && 2 * t <= T) {
        it.state = 1;
        st.push_back({2 * t, 0});
      } else if (it.state <= 1 && 2 * t + 1 <= T) {
        it.state = 2;
        st.push_back({2 * t + 1, 0});
      } else {
        pl.vb[t] = (int)pl.old_of_new.size();
        pl.nv[t] = (int)b.own[t].size();
        for (int o : b.own[t]) {
          new_of_old[o] = (int)pl.old_of_new.size();
          pl.old_of_new.push_back(o);
        }
        st.pop_back();
      }
    }
  }
  pl.nv[1] += 2;  // the pose pseudo-vertices V, V+1 follow the root's separator
  pl.owner.assign(V, 0);
  for (int t = 1; t <= T; t++) {
    const int npts = (t == 1) ? pl.nv[t] - 2 : pl.nv[t];
    for (int k = 0; k < npts; k++) pl.owner[pl.vb[t] + k] = t;
  }

  // boundaries, children before parents (descending heap index)
  std::vector<std::vector<int>> bnd(T + 1);
  std::vector<int> mark(V + 3, 0);
  for (int t = T; t >= 1; t--) {
    std::vector<int>& bt = bnd[t];
    const int npts = (t == 1) ? pl.nv[t] - 2 : pl.nv[t];
    const int last_own = pl.vb[t] + pl.nv[t] - 1;  // root: V + 1
    auto add = [&](int v) {
      if (v > last_own && !mark[v]) {
        mark[v] = 1;
        bt.push_back(v);
      }
    };
    if (2 * t <= T)
      for (int c = 2 * t; c <= 2 * t + 1; c++)
        for (int v : bnd[c]) add(v);
    for (int k = 0; k < npts; k++)
      for (int o : adj[pl.old_of_new[pl.vb[t] + k]]) add(new_of_old[o]);
    add(V);
    add(V + 1);
    add(V + 2);
    std::sort(bt.begin(), bt.end());
    for (int v : bt) mark[v] = 0;
  }
  pl.nbv.assign(T + 1, 0);
  pl.bnd_ptr.assign(T + 2, 0);
  pl.bnd.clear();
  for (int t = 1; t <= T; t++) {
    pl.bnd_ptr[t] = (int)pl.bnd.size();
    pl.nbv[t] = (int)bnd[t].size();
    pl.bnd.insert(pl.bnd.end(), bnd[t].begin(), bnd[t].end());
  }
  pl.bnd_ptr[T + 1] = (int)pl.bnd.size();

  // path offsets and boundary -> path-vector index
  pl.path_off.assign(T + 1, 0);
  for (int t = 2; t <= T; t++) pl.path_off[t] = pl.path_off[t / 2] + 3 * pl.nv[t / 2];
  pl.max_path = 0;
  for (int t = 1; t <= T; t++) pl.max_path = std::max(pl.max_path, pl.path_off[t] + 3 * pl.nv[t]);
  pl.bpath.assign(pl.bnd.size(), -1);
  for (int t = 2; t <= T; t++)
    for (int k = 0; k < pl.nbv[t]; k++) {
      const int v = pl.bnd[pl.bnd_ptr[t] + k];
      if (v == V + 2) continue;
      int a = t / 2;
      while (a >= 1 && !(v >= pl.vb[a] && v < pl.vb[a] + pl.nv[a])) a /= 2;
      pl.bpath[pl.bnd_ptr[t] + k] = (a >= 1) ? pl.path_off[a] + 3 * (v - pl.vb[a]) : -1;
    }

  // inverse maps child boundary <- parent front position
  pl.inv_ptr.assign(T + 2, 0);
  pl.inv.clear();
  for (int t = 2; t <= T; t++) {
    const int p = t / 2;
    pl.inv_ptr[t] = (int)pl.inv.size();
    const int nf = pl.nv[p] + pl.nbv[p];
    pl.inv.resize(pl.inv.size() + nf, -1);
    int* iv = pl.inv.data() + pl.inv_ptr[t];
    for (int k = 0; k < pl.nbv[t]; k++) {
      const int v = pl.bnd[pl.bnd_ptr[t] + k];
      int pos;
      if (v >= pl.vb[p] && v < pl.vb[p] + pl.nv[p]) {
        pos = v - pl.vb[p];
      } else {
        const int* bb = pl.bnd.data() + pl.bnd_ptr[p];
        pos = pl.nv[p] + (int)(std::lower_bound(bb, bb + pl.nbv[p], v) - bb);
      }
      iv[pos] = k;
    }
  }
  pl.inv_ptr[T + 1] = (int)pl.inv.size();

  // storage
  pl.p_off.assign(T + 1, 0);
  pl.u_off.assign(T + 1, 0);
  pl.p_total = pl.u_total = 0;
  size_t smem = 0;
  for (int t = 1; t <= T; t++) {
    const long long ns = 3LL * pl.nv[t], nb = 3LL * pl.nbv[t];
    pl.p_off[t] = pl.p_total;
    pl.p_total += (ns + nb) * ns;
    pl.u_off[t] = pl.u_total;
    pl.u_total += nb * nb;
    // shared-memory need of the node's team member with the most rows (see nrs_direct_core.cuh):
    int d = 0;
    while ((2 << d) <= t) d++;  // depth of t
    const int R = pl.G >> d;
    const long long rows_ab = pl.nv[t] + (pl.nbv[t] + R - 1) / R;
    const long long ld = (ns | 1);
    const size_t ab = (size_t)(3 * rows_ab * ld);
    const size_t c = (size_t)(nb * ld);            // stage C: every boundary row of the panel
    const size_t bw = (size_t)(ns * ld);           // backward: L11
    smem = std::max(smem, std::max(ab, std::max(c, bw)));
  }
  pl.smem_doubles = smem;
}

// Front position of the other endpoint of every pair incidence (rows already renumbered): inc_pos[a] >= 0 when the
// block (other, row) belongs to the panel of owner(row), i.e. `other` is eliminated after `row`; -1 otherwise.
inline void direct_inc_pos(const DirectPlanHost& pl, const std::vector<int>& inc_ptr, const std::vector<int>& inc_other,
                           std::vector<int>& inc_pos) {
  inc_pos.assign(inc_other.size(), -1);
  for (int v = 0; v < pl.V; v++) {
    const int t = pl.owner[v];
    const int* bb = pl.bnd.data() + pl.bnd_ptr[t];
    for (int a = inc_ptr[v]; a < inc_ptr[v + 1]; a++) {
      const int o = inc_other[a];
      if (o <= v) continue;
      if (o < pl.vb[t] + pl.nv[t] && o < pl.V)
        inc_pos[a] = o - pl.vb[t];
      else
        inc_pos[a] = pl.nv[t] + (int)(std::lower_bound(bb, bb + pl.nbv[t], o) - bb);
    }
  }
}

}  // namespace nrs
