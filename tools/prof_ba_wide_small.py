"""compute-sanitizer driver for the wide BA variant (nrs_lm_kernel_wide: cooperative grid, two CTAs per SM, segment
reductions, 256-bit record loads): the smallest window that runs on it (> 148 chunks of 128 rows), 2 LM iterations."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
q = synth.ba_problem("c3", n=3000, n_kf=10, run=7)
b = core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"], iterations=2)
print("wide window", len(q["obs_kf"]), "observations",
      {k: b["stats"][k] for k in ("gpu_ms", "lm_iterations", "pcg_iterations", "grid_ctas", "block_threads")})
core.close()
