"""Profiling driver for the kernels either side of the LM engines (used under ncu, profiles/README.md): image
pre-processing (gray + CLAHE + mask), Shi-Tomasi, KLT pyramid + reference patches + track, the regularisation-graph
kernels (UpdateVertex loop, GetEdges top-k). Each stage runs twice (first call warms allocations)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
rng = np.random.default_rng(5)
p = synth.klt_pair(seed=41, n_points=2000)
rgb = np.repeat(p["ref"][:, :, None], 3, 2).copy()
pre = api.Pre(core)
shi = api.ShiTomasi(core)
k = api.KLT(core)
t = synth.tracking_problem("c2", seed=1235)
g = t["graph"]
verts = np.unique(t["point_vertex"]).astype(np.int32)
pos = t["last_world_position"] + synth.smooth_field(np.random.default_rng(3), t["last_world_position"], 0.05)
for rep in range(2):
    gray, eq = pre.image(rgb)
    m = pre.mask(None, [("bright", 200), ("border", 6, 6, 10, 8)], shape=gray.shape)
    s = shi.extract(eq, existing=p["pts"][:500])
    k.set_reference(p["ref"], p["pts"])
    r = k.track(p["cur"], p["pts"], p["status"])
    core.graph_update_vertices(g.copy(), verts, pos)
    core.graph_get_edges_batch(g, verts, top_k=32)
print("shi", s["n"], "tracked", r["n_tracked"])
k.close()
shi.close()
pre.close()
core.close()
