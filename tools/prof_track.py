"""Profiling driver: one configs[1] pose+deformation call (two LM kernel launches: main rounds, lost-point stage),
or one configs[2] BA call with `ba`. Used under ncu (see profiles/README.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
which = sys.argv[1] if len(sys.argv) > 1 else "track"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if which == "track":
    p = synth.tracking_problem("c2", seed=1235)
    for _ in range(reps):
        r0 = core.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
        r = core.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"],
                             p["graph"].copy(), p["scale"], r0["pose"], p["last_world_position"])
    print({k: r["stats"][k] for k in ("gpu_ms", "lm_iterations", "lm_trials", "pcg_iterations", "grid_ctas", "block_threads")})
else:
    q = synth.ba_problem("c3")
    for _ in range(reps):
        b = core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
    print({k: b["stats"][k] for k in ("gpu_ms", "lm_iterations", "lm_trials", "pcg_iterations", "grid_ctas", "block_threads")})
core.close()
