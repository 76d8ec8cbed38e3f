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
#include <functional>
#include <type_traits>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace nrs {

struct DirectPlanHost {
  int V = 0;        // point rows (deformation vertices)
  int depth = 0;    // leaves at this depth
  int n_nodes = 0;  // 2^(depth+1) - 1, nodes are 1..n_nodes
  int G = 1;        // CTAs = 2^depth
  int np = 2;       // pose pseudo-vertices owned by the root: 2 (pose is an unknown) or 0 (pose fixed)
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
  int max_rows = 0;           // most panel rows (own + boundary share) of any team member
  std::vector<int> owner;     // [V] tree node owning a row
};

// Depth of the dissection for V rows on at most max_ctas CTAs: leaves of >= ~8 vertices, at most 2^7 leaves.
inline int direct_depth(int V, int max_ctas, int min_leaf = 8) {
  int d = 0;
  while (d < 7 && (2 << d) <= max_ctas && (V >> (d + 1)) >= min_leaf) d++;
  return d;
}

// Builds the plan. uv: [2V] pixel coordinates (caller rows); pair_i / pair_j: regulariser pairs (caller rows).
// After the call the caller permutes its rows with plan.old_of_new and then calls direct_inc_pos with the new ids.
#ifndef NRS_DIRECT_SCAN_NODES
#define NRS_DIRECT_SCAN_NODES 3
#endif
constexpr int kScanNodes = NRS_DIRECT_SCAN_NODES;  // heap indices 1..kScanNodes get the cut-position scan (3 = two levels)

inline void build_direct_plan(int V, const double* uv, const std::vector<int>& pair_i, const std::vector<int>& pair_j,
                              int depth, DirectPlanHost& pl, bool with_pose = true) {
  pl.V = V;
  pl.np = with_pose ? 2 : 0;
  const int np = pl.np;
  pl.depth = depth;
  pl.n_nodes = (2 << depth) - 1;
  pl.G = 1 << depth;
  const int T = pl.n_nodes;
  std::vector<int> adj_ptr(V + 1, 0), adj(2 * pair_i.size());
  for (size_t e = 0; e < pair_i.size(); e++) {
    adj_ptr[pair_i[e] + 1]++;
    adj_ptr[pair_j[e] + 1]++;
  }
  for (int i = 0; i < V; i++) adj_ptr[i + 1] += adj_ptr[i];
  {
    std::vector<int> w(adj_ptr.begin(), adj_ptr.end() - 1);
    for (size_t e = 0; e < pair_i.size(); e++) {
      adj[w[pair_i[e]]++] = pair_j[e];
      adj[w[pair_j[e]]++] = pair_i[e];
    }
  }
  // (1) k-d partition of the pixel coordinates: `depth` rounds of median cuts along the wider axis. Keys are
  // (coordinate bits, caller row) packed into 64 bits: a total order, so the result is deterministic.
  const int n_leaf = 1 << depth;
  std::vector<int> ord(V), leaf(V, 0), lo_of(T + 2, 0), hi_of(T + 2, 0);
  std::iota(ord.begin(), ord.end(), 0);
  lo_of[1] = 0;
  hi_of[1] = V;
  {
    std::vector<uint64_t> key(V);
    auto fkey = [](double x) {  // order-preserving map of a float to uint32
      const float f = (float)x;
      uint32_t u;
      memcpy(&u, &f, 4);
      return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    };
    for (int t = 1; t < n_leaf; t++) {
      const int lo = lo_of[t], hi = hi_of[t], mid = lo + (hi - lo) / 2;
      lo_of[2 * t] = lo;
      hi_of[2 * t] = mid;
      lo_of[2 * t + 1] = mid;
      hi_of[2 * t + 1] = hi;
      if (hi - lo < 2) {  // 0 or 1 vertices: pushed down the left spine
        hi_of[2 * t] = hi;
        lo_of[2 * t + 1] = hi;
        continue;
      }
      double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
      for (int k = lo; k < hi; k++)
        for (int a = 0; a < 2; a++) {
          const double c = uv[2 * (size_t)ord[k] + a];
          mn[a] = std::min(mn[a], c);
          mx[a] = std::max(mx[a], c);
        }
      const int ax = (mx[0] - mn[0] >= mx[1] - mn[1]) ? 0 : 1;
      for (int k = lo; k < hi; k++) key[k] = ((uint64_t)fkey(uv[2 * (size_t)ord[k] + ax]) << 32) | (uint32_t)ord[k];
      if (t <= kScanNodes && hi - lo >= 128) {
        // The separators of the top two levels are the largest fronts (the root's must fit the shared memory of one
        // SM, and they sit on every CTA's critical path): instead of the exact median, take the cut position within
        // +-8 % of it whose cut edges have the smallest vertex cover (= the largest matching of the bipartite cut
        // graph). Ties go to the position closest to the median.
        std::sort(key.begin() + lo, key.begin() + hi);
        std::vector<int> rank(V, -1);
        for (int k = lo; k < hi; k++) rank[(int)(uint32_t)key[k]] = k;
        std::vector<int> node_edges;
        for (size_t e = 0; e < pair_i.size(); e++)
          if (rank[pair_i[e]] >= 0 && rank[pair_j[e]] >= 0) node_edges.push_back((int)e);
        const int step = std::max(1, (hi - lo) / 50);
        int best_mid = mid, best_cover = -1;
        std::vector<int> lidx(V, -1), l_ptr, l_adj, mt, stamp;
        for (int c = 0; c < 9; c++) {
          const int off = (c == 0) ? 0 : ((c + 1) / 2) * ((c & 1) ? 1 : -1);  // 0, +1, -1, +2, -2, ...
          const int m = mid + off * step;
          std::vector<std::pair<int, int>> ed;
          int nl = 0, nr = 0;
          std::vector<int> touched;
          for (int e : node_edges) {
            int a = pair_i[e], b2 = pair_j[e];
            if ((rank[a] < m) == (rank[b2] < m)) continue;
            if (rank[a] >= m) std::swap(a, b2);
            if (lidx[a] < 0) { lidx[a] = nl++; touched.push_back(a); }
            if (lidx[b2] < 0) { lidx[b2] = nr++; touched.push_back(b2); }
            ed.emplace_back(lidx[a], lidx[b2]);
          }
          for (int v : touched) lidx[v] = -1;
          l_ptr.assign(nl + 1, 0);
          for (auto& x : ed) l_ptr[x.first + 1]++;
          for (int k = 0; k < nl; k++) l_ptr[k + 1] += l_ptr[k];
          l_adj.resize(ed.size());
          {
            std::vector<int> w(l_ptr.begin(), l_ptr.end() - 1);
            for (auto& x : ed) l_adj[w[x.first]++] = x.second;
          }
          mt.assign(nr, -1);
          stamp.assign(nr, -1);
          std::function<bool(int, int)> aug = [&](int u, int st) -> bool {
            for (int a = l_ptr[u]; a < l_ptr[u + 1]; a++) {
              const int r = l_adj[a];
              if (stamp[r] == st) continue;
              stamp[r] = st;
              if (mt[r] < 0 || aug(mt[r], st)) {
                mt[r] = u;
                return true;
              }
            }
            return false;
          };
          int cover = 0;
          for (int u = 0; u < nl; u++) cover += aug(u, u) ? 1 : 0;
          if (best_cover < 0 || cover < best_cover) {
            best_cover = cover;
            best_mid = m;
          }
        }
        hi_of[2 * t] = best_mid;
        lo_of[2 * t + 1] = best_mid;
      } else {
        std::nth_element(key.begin() + lo, key.begin() + mid, key.begin() + hi);
      }
      for (int k = lo; k < hi; k++) ord[k] = (int)(uint32_t)key[k];
    }
    for (int l = 0; l < n_leaf; l++)
      for (int k = lo_of[n_leaf + l]; k < hi_of[n_leaf + l]; k++) leaf[ord[k]] = l;
  }
  // (2) every pair is cut at the lowest common ancestor of its endpoints' leaves
  auto lca = [&](int a, int b) -> int {  // 0: same leaf
    const unsigned x = (unsigned)(leaf[a] ^ leaf[b]);
    if (!x) return 0;
    const int hb = 31 - __builtin_clz(x);
    return (1 << (depth - 1 - hb)) + (leaf[a] >> (hb + 1));
  };
  std::vector<int> cut_ptr(T + 2, 0), cut_e;
  {
    std::vector<int> node_of(pair_i.size());
    for (size_t e = 0; e < pair_i.size(); e++) {
      node_of[e] = lca(pair_i[e], pair_j[e]);
      cut_ptr[node_of[e] + 1]++;
    }
    for (int t = 0; t <= T; t++) cut_ptr[t + 1] += cut_ptr[t];
    cut_e.resize(pair_i.size());
    std::vector<int> w(cut_ptr.begin(), cut_ptr.end() - 1);
    for (size_t e = 0; e < pair_i.size(); e++) cut_e[w[node_of[e]]++] = (int)e;
  }
  // (3) separators, top down: a MINIMUM vertex cover of the node's cut edges whose endpoints are not already in an
  // ancestor's separator. The cut edges of a node join its left and its right subtree — a bipartite graph — so the
  // minimum cover follows from a maximum matching (Koenig): with Z = the vertices reachable from the unmatched left
  // vertices along alternating paths, cover = (left \ Z) + (right & Z). Augmenting paths in a fixed vertex order keep
  // the plan deterministic. (The first version took vertices greedily by remaining cut degree: covers 10-20 % larger,
  // and the root front of 2 frames in 8 no longer fitted the shared memory of one SM.)
  std::vector<std::vector<int>> own(T + 1);
  std::vector<char> insep(V, 0);
  {
    std::vector<int> lid(V, -1), verts, side, match, eptr, eadj, seen_stamp;
    std::vector<char> inz;
    for (int t = 1; t < n_leaf; t++) {
      int level = 0;
      while ((1 << (level + 1)) <= t) level++;
      const int sbit = depth - 1 - level;  // leaf-index bit that tells the node's left subtree from its right one
      verts.clear();
      std::vector<std::pair<int, int>> edges;  // (left local id, right local id)
      auto local = [&](int v) {
        if (lid[v] < 0) {
          lid[v] = (int)verts.size();
          verts.push_back(v);
        }
        return lid[v];
      };
      for (int c = cut_ptr[t]; c < cut_ptr[t + 1]; c++) {
        int a = pair_i[cut_e[c]], b2 = pair_j[cut_e[c]];
        if (insep[a] || insep[b2]) continue;
        if ((leaf[a] >> sbit) & 1) std::swap(a, b2);  // a: left subtree, b2: right subtree
        const int la = local(a), lb = local(b2);
        edges.emplace_back(la, lb);
      }
      const int nl = (int)verts.size();
      side.assign(nl, 0);
      for (int k = 0; k < nl; k++) side[k] = (leaf[verts[k]] >> sbit) & 1;
      eptr.assign(nl + 1, 0);
      for (auto& ed : edges) eptr[ed.first + 1]++;
      for (int k = 0; k < nl; k++) eptr[k + 1] += eptr[k];
      eadj.resize(edges.size());
      {
        std::vector<int> w(eptr.begin(), eptr.end() - 1);
        for (auto& ed : edges) eadj[w[ed.first]++] = ed.second;
      }
      match.assign(nl, -1);
      seen_stamp.assign(nl, -1);
      // Kuhn's augmenting paths from every left vertex, in local-id (first touched) order
      std::function<bool(int, int)> augment = [&](int u, int stamp) -> bool {
        for (int a = eptr[u]; a < eptr[u + 1]; a++) {
          const int r = eadj[a];
          if (seen_stamp[r] == stamp) continue;
          seen_stamp[r] = stamp;
          if (match[r] < 0 || augment(match[r], stamp)) {
            match[r] = u;
            match[u] = r;
            return true;
          }
        }
        return false;
      };
      for (int u = 0; u < nl; u++)
        if (!side[u] && eptr[u + 1] > eptr[u]) augment(u, u);
      // alternating reachability from the unmatched left vertices
      inz.assign(nl, 0);
      std::vector<int> stack;
      for (int u = 0; u < nl; u++)
        if (!side[u] && match[u] < 0) {
          inz[u] = 1;
          stack.push_back(u);
        }
      while (!stack.empty()) {
        const int u = stack.back();
        stack.pop_back();
        for (int a = eptr[u]; a < eptr[u + 1]; a++) {
          const int r = eadj[a];
          if (match[u] == r || inz[r]) continue;  // unmatched edges left -> right
          inz[r] = 1;
          const int u2 = match[r];                // matched edge right -> left
          if (u2 >= 0 && !inz[u2]) {
            inz[u2] = 1;
            stack.push_back(u2);
          }
        }
      }
      std::vector<int>& sep = own[t];
      for (int k = 0; k < nl; k++) {
        const bool in_cover = side[k] ? (inz[k] != 0) : (inz[k] == 0 && eptr[k + 1] > eptr[k]);
        if (in_cover) {
          sep.push_back(verts[k]);
          insep[verts[k]] = 1;
        }
      }
      for (int v : verts) lid[v] = -1;
      std::sort(sep.begin(), sep.end());
    }
    for (int l = 0; l < n_leaf; l++) {
      std::vector<int>& o = own[n_leaf + l];
      for (int k = lo_of[n_leaf + l]; k < hi_of[n_leaf + l]; k++)
        if (!insep[ord[k]]) o.push_back(ord[k]);
    }
  }

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
        pl.nv[t] = (int)own[t].size();
        for (int o : own[t]) {
          new_of_old[o] = (int)pl.old_of_new.size();
          pl.old_of_new.push_back(o);
        }
        st.pop_back();
      }
    }
  }
  pl.nv[1] += np;  // the pose pseudo-vertices V, V+1 follow the root's separator
  pl.owner.assign(V, 0);
  for (int t = 1; t <= T; t++) {
    const int npts = (t == 1) ? pl.nv[t] - np : pl.nv[t];
    for (int k = 0; k < npts; k++) pl.owner[pl.vb[t] + k] = t;
  }

  // boundaries, children before parents (descending heap index). A boundary vertex is owned by an ancestor, and the
  // ancestors' own ranges ascend towards the root, so scanning them with a mark array emits the list sorted.
  std::vector<int> tmp_ptr(T + 2, 0), tmp;  // lists in processing order (t descending)
  std::vector<char> mark(V + 3, 0);
  pl.nbv.assign(T + 1, 0);
  tmp.reserve(32 * (size_t)T);
  for (int t = T; t >= 1; t--) {
    const int npts = (t == 1) ? pl.nv[t] - np : pl.nv[t];
    const int last_own = pl.vb[t] + pl.nv[t] - 1;  // root: V + 1
    if (2 * t <= T)
      for (int c = 2 * t; c <= 2 * t + 1; c++)
        for (int k = tmp_ptr[c]; k < tmp_ptr[c] + pl.nbv[c]; k++)
          if (tmp[k] > last_own) mark[tmp[k]] = 1;
    for (int k = 0; k < npts; k++) {
      const int o_row = pl.old_of_new[pl.vb[t] + k];
      for (int a = adj_ptr[o_row]; a < adj_ptr[o_row + 1]; a++) {
        const int v = new_of_old[adj[a]];
        if (v > last_own) mark[v] = 1;
      }
    }
    tmp_ptr[t] = (int)tmp.size();
    for (int a = t / 2; a >= 1; a /= 2) {
      const int e = pl.vb[a] + ((a == 1) ? pl.nv[a] - np : pl.nv[a]);
      for (int v = pl.vb[a]; v < e; v++)
        if (mark[v]) {
          mark[v] = 0;
          tmp.push_back(v);
        }
    }
    if (t != 1 && np) {
      tmp.push_back(V);
      tmp.push_back(V + 1);
    }
    tmp.push_back(V + 2);
    pl.nbv[t] = (int)tmp.size() - tmp_ptr[t];
  }
  pl.bnd_ptr.assign(T + 2, 0);
  pl.bnd.clear();
  pl.bnd.reserve(tmp.size());
  for (int t = 1; t <= T; t++) {
    pl.bnd_ptr[t] = (int)pl.bnd.size();
    pl.bnd.insert(pl.bnd.end(), tmp.begin() + tmp_ptr[t], tmp.begin() + tmp_ptr[t] + pl.nbv[t]);
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
  pl.max_rows = 0;
  for (int t = 1; t <= T; t++) {
    const long long ns = 3LL * pl.nv[t], nb = 3LL * pl.nbv[t];
    pl.p_off[t] = pl.p_total;
    pl.p_total += (ns + nb) * ns;
    pl.p_total += pl.p_total & 1;  // panels start 16-byte aligned (bulk copies)
    pl.u_off[t] = pl.u_total;
    pl.u_total += nb * nb;
    // shared-memory need of the node's team member with the most rows (see nrs_direct_core.cuh):
    int d = 0;
    while ((2 << d) <= t) d++;  // depth of t
    const int R = pl.G >> d;
    const long long rows_ab = pl.nv[t] + (pl.nbv[t] + R - 1) / R;
    const long long ld = (ns | 1);
    const size_t ab = (size_t)(3 * rows_ab * ld);
    const size_t c = (size_t)((nb + 3 * ((pl.nbv[t] + R - 1) / R)) * ld);  // stage C: all boundary rows + mine times D
    pl.max_rows = std::max(pl.max_rows, (int)rows_ab);
    const size_t bw = (size_t)(ns * ns) + 2;       // backward: L11 as it lies in global memory
    smem = std::max(smem, std::max(ab, std::max(c, bw)));
  }
  pl.smem_doubles = smem;
}

// Front position of the other endpoint of every pair incidence (rows already renumbered): inc_pos[a] >= 0 when the
// block (other, row) belongs to the panel of owner(row), i.e. `other` is eliminated after `row`; -1 otherwise.
template <typename Pool = void>
inline void direct_inc_pos(const DirectPlanHost& pl, const std::vector<int>& inc_ptr, const std::vector<int>& inc_other,
                           std::vector<int>& inc_pos, Pool* pool = nullptr) {
  inc_pos.assign(inc_other.size(), -1);
  auto rows = [&](int v0, int v1) {
    for (int v = v0; v < v1; v++) {
      const int t = pl.owner[v];
      const int* bb = pl.bnd.data() + pl.bnd_ptr[t];
      for (int a = inc_ptr[v]; a < inc_ptr[v + 1]; a++) {
        const int o = inc_other[a];
        if (o <= v || o >= pl.V) continue;  // earlier in the elimination order, or a fixed row (not an unknown)
        if (o < pl.vb[t] + pl.nv[t] && o < pl.V)
          inc_pos[a] = o - pl.vb[t];
        else
          inc_pos[a] = pl.nv[t] + (int)(std::lower_bound(bb, bb + pl.nbv[t], o) - bb);
      }
    }
  };
  if constexpr (std::is_void<Pool>::value) {
    rows(0, pl.V);
  } else {
    // rows are independent: the staging threads take ranges
    if (pool && pl.V >= 512)
      pool->run([&](int t, int nt) { rows((int)((long long)pl.V * t / nt), (int)((long long)pl.V * (t + 1) / nt)); });
    else
      rows(0, pl.V);
  }
}

}  // namespace nrs
