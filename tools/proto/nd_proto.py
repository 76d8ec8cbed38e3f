"""Prototype: geometric nested dissection of the tracking frame's regulariser graph, multifrontal front sizes."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from nrslam_b200 import synth

def build(n=None, leaf=24, cfg="c2"):
    p = synth.tracking_problem(cfg, n=n)
    g = p["graph"]
    uv = p["uv"].astype(np.float64)
    N = len(uv)
    rowptr, cols = np.asarray(g.rowptr), np.asarray(g.col)
    nv = len(rowptr) - 1
    print("points", N, "graph vertices", nv, "edges", len(cols) // 2)
    return p, uv, rowptr, cols

def nd(ids, uv, adjset, leaf, depth=0):
    """returns tree: dict(own=[...], children=[...])"""
    if len(ids) <= leaf:
        return dict(own=list(ids), ch=[], depth=depth)
    pts = uv[ids]
    ax = 0 if np.ptp(pts[:, 0]) >= np.ptp(pts[:, 1]) else 1
    order = np.argsort(pts[:, ax], kind="stable")
    half = len(ids) // 2
    A = set(int(ids[k]) for k in order[:half]); B = set(int(ids[k]) for k in order[half:])
    # cut edges
    sa = set(a for a in A if any((b in B) for b in adjset[a]))
    sb = set(b for b in B if any((a in A) for a in adjset[b]))
    # greedy vertex cover of bipartite cut graph
    cut = [(a, b) for a in sa for b in adjset[a] if b in sb]
    sep = set()
    deg = {}
    for a, b in cut:
        deg[a] = deg.get(a, 0) + 1; deg[b] = deg.get(b, 0) + 1
    rem = set(cut)
    # max-degree greedy
    import heapq
    while rem:
        v = max(deg, key=lambda k: deg[k])
        if deg[v] == 0: break
        sep.add(v)
        for e in [e for e in rem if v in e]:
            rem.discard(e)
            deg[e[0]] -= 1; deg[e[1]] -= 1
        deg[v] = 0
    A2 = np.array(sorted(A - sep), dtype=np.int64); B2 = np.array(sorted(B - sep), dtype=np.int64)
    return dict(own=sorted(sep), ch=[nd(A2, uv, adjset, leaf, depth + 1), nd(B2, uv, adjset, leaf, depth + 1)], depth=depth)

def analyse(tree, adjset, npose=6):
    # postorder; compute boundary sets
    stats = []
    def rec(t, anc):  # anc: set of ancestor vertices
        anc2 = anc | set(t["own"])
        bnd = set()
        for c in t["ch"]:
            bnd |= rec(c, anc2)
        for v in t["own"]:
            bnd |= (adjset[v] & anc)
        bnd -= set(t["own"])
        bnd &= anc
        ns = 3 * len(t["own"]); nb = 3 * len(bnd) + npose
        t["ns"], t["nb"] = ns, nb
        nf = ns + nb
        fl = sum((nf - k) ** 2 for k in range(ns))  # multiply-adds*... approx flops/2
        t["fl"] = fl
        stats.append((t["depth"], ns, nb, nf, fl))
        return bnd | (set(t["own"]) & anc)  # own are not in anc; boundary passes up
    rec(tree, set())
    return stats

if __name__ == "__main__":
    leaf = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    p, uv, rowptr, cols = build()
    N = len(uv)
    # graph vertex ids -> frame points: use point_vertex mapping
    pv = np.asarray(p["point_vertex"])
    inv = -np.ones(len(rowptr) - 1, np.int64); inv[pv] = np.arange(N)
    adjset = [set() for _ in range(N)]
    for i in range(N):
        v = pv[i]
        for c in cols[rowptr[v]:rowptr[v + 1]]:
            j = inv[c]
            if j >= 0: adjset[i].add(int(j))
    tree = nd(np.arange(N), uv, adjset, leaf)
    st = analyse(tree, adjset)
    st = np.array(st)
    maxd = st[:, 0].max()
    tot = 0; crit = 0
    for d in range(maxd + 1):
        s = st[st[:, 0] == d]
        print("depth %d: %3d fronts  ns max %4d mean %5.0f  nf max %4d mean %5.0f  MFMA max %6.2f sum %7.2f" % (
            d, len(s), s[:, 1].max(), s[:, 1].mean(), s[:, 3].max(), s[:, 3].mean(), s[:, 4].max() / 1e6, s[:, 4].sum() / 1e6))
        tot += s[:, 4].sum(); crit += s[:, 4].max()
    nnz = sum(ns * nb + ns * (ns + 1) // 2 for _, ns, nb, nf, fl in st)
    print("total MFMA %.1f  critical MFMA %.1f  nnz(L) %d" % (tot / 1e6, crit / 1e6, nnz))
