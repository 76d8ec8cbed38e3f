"""Tracking frame of every bench rank's seed on ONE GPU: which engine ran, solves, device time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
for seed in range(1235, 1243):
    p = synth.tracking_problem("c2", seed=seed)
    for rep in range(2):
        r0, r1 = core.track_pose_and_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                                            p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
    s = r1["stats"]
    print(seed, "gpu_ms %.2f" % s["gpu_ms"], "pose_only %.2f" % r0["stats"]["gpu_ms"], {k: s[k] for k in ("lm_iterations", "lm_trials", "direct_solves", "pcg_iterations", "solve_failures", "grid_ctas", "kernel_launches", "n_pair_edges")}, "lost", len(r1["lost"]), flush=True)
core.close()
