"""Measured differences of the exact-solve tracking engine against the oracle on the parity-test problems."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import nrslam_b200
from nrslam_b200 import api, synth
import oracle_lib
core = api.Core(); orc = oracle_lib.Oracle()
for cfg, n, kw in [("c1", None, {}), ("c2", None, {}), ("c2", 700, dict(outlier_frac=0.3)), ("c1", 120, dict(extra_frac=0.0)), ("c1", 400, {}), ("c1", 300, {})]:
    p = synth.tracking_problem(cfg, n=n, **kw)
    args = (p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"])
    a = orc.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
    b = core.pose_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
    ta, tb = np.array(a["stats"]["chi2_trace"]), np.array(b["stats"]["chi2_trace"])
    print(cfg, n, "pose %.2e def %.2e X %.2e last %.2e chi2 %.2e median %.2e trace %.2e" % (
        np.abs(a["pose"] - b["pose"]).max(), np.abs(a["deformation"] - b["deformation"]).max(), np.abs(a["X"] - b["X"]).max(),
        np.abs(a["last_pos"] - b["last_pos"]).max(), np.abs(a["chi2"] - b["chi2"]).max(), abs(a["median"] - b["median"]),
        np.abs(ta / tb - 1).max() if len(ta) == len(tb) else -1), a["stats"]["lm_trials"], b["stats"]["lm_trials"])
