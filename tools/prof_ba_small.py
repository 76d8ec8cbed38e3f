"""compute-sanitizer driver: the reference's 5-keyframe BA window (runs as ONE 16-CTA cluster: cluster barrier + DSMEM
halo pushes of the cluster-native CG loop) and a mid-size window on the cooperative grid (atomics grid barrier)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
for cfg, kw in (("c1", {}), ("c3", dict(n=1500, n_kf=10, run=6))):
    q = synth.ba_problem(cfg, **kw)
    b = core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
    print(cfg, {k: b["stats"][k] for k in ("gpu_ms", "lm_iterations", "pcg_iterations", "grid_ctas", "block_threads")})
core.close()
