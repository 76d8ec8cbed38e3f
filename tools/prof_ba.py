"""Profiling target: stage the configs[2] BA window once, then re-run its device program (one nrs_lm_kernel_wide launch
per resolve)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa: F401,E402
from nrslam_b200 import api, synth  # noqa: E402

core = api.Core()
q = synth.ba_problem("c3")
core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
s = core.resolve(2)
print("BA c3 gpu_ms %.3f lm %d pcg %d grid %d x %d" % (s["gpu_ms"], s["lm_iterations"], s["pcg_iterations"],
                                                       s["grid_ctas"], s["block_threads"]))
core.close()
