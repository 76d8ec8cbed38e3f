import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import nrslam_b200
from nrslam_b200 import api, synth
import oracle_lib
core = api.Core(); orc = oracle_lib.Oracle()
for n in (40, 150, 700, 3000):
    for seed in (400 + n, 500 + n, 600 + n):
        p = synth.tracking_problem("c2", n=n, seed=seed)
        args = (p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"])
        a = core.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
        os.environ["NRSLAM_B200_DIRECT"] = "0"
        c = core.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
        os.environ.pop("NRSLAM_B200_DIRECT")
        b = orc.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
        ta, tb = np.array(a["stats"]["chi2_trace"]), np.array(b["stats"]["chi2_trace"])
        m = min(len(ta), len(tb))
        print(n, seed, "it", a["stats"]["lm_iterations"], c["stats"]["lm_iterations"], b["stats"]["lm_iterations"], "tr", a["stats"]["lm_trials"], c["stats"]["lm_trials"], b["stats"]["lm_trials"],
              "fail", a["stats"]["solve_failures"], "pose", np.abs(a["pose"]-b["pose"]).max(), "def", np.abs(a["deformation"]-b["deformation"]).max(),
              "trace", np.abs(ta[:m]/tb[:m]-1).max(), "status", np.array_equal(a["status"], b["status"]), "lost", len(a["lost"]))
