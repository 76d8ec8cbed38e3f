"""Where does the end-to-end time of a tracking frame go? (host staging vs device vs copies)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
p = synth.tracking_problem("c2", seed=1235)
for rep in range(4):
    t0 = time.perf_counter()
    r0 = core.pose_only(p["cam"], p["uv"], p["X_rest"], p["seed_pose"])
    t1 = time.perf_counter()
    g = p["graph"].copy()
    t2 = time.perf_counter()
    r1 = core.pose_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"], g, p["scale"],
                          r0["pose"], p["last_world_position"])
    t3 = time.perf_counter()
    s0, s1 = r0["stats"], r1["stats"]
    print("rep %d pose_only: wall %.2f ms (lib host %.2f, stage %.2f, gpu %.2f) | graph copy %.2f | pose_deform: wall %.2f ms "
          "(lib host %.2f, stage %.2f, gpu %.2f, launches %d, lost %d, h2d %d d2h %d)" % (
              rep, 1e3 * (t1 - t0), s0["host_ms"], s0["stage_ms"], s0["gpu_ms"], 1e3 * (t2 - t1), 1e3 * (t3 - t2),
              s1["host_ms"], s1["stage_ms"], s1["gpu_ms"], s1["kernel_launches"], len(r1["lost"]), s1["h2d_bytes"], s1["d2h_bytes"]))
core.close()
