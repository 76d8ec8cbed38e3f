"""Host-phase timers of the end-to-end tracking frame (track_pose_and_deform), NRSLAM_B200_HOSTPROF=1."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
p = synth.tracking_problem("c2", seed=1235)
im = synth.klt_pair(seed=77, n_points=2000)
klt = api.KLT(core)
klt.set_reference(im["ref"], im["pts"])
for rep in range(8):
    if rep == 6:
        os.environ["NRSLAM_B200_HOSTPROF"] = "1"
    t0 = time.perf_counter()
    klt.track(im["cur"], im["pts"], im["status"])
    t1 = time.perf_counter()
    g = p["graph"].copy()
    t2 = time.perf_counter()
    r0, r1 = core.track_pose_and_deform(p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"], g,
                                        p["scale"], p["seed_pose"], p["last_world_position"])
    t3 = time.perf_counter()
    s0, s1 = r0["stats"], r1["stats"]
    print("rep %d klt %.3f | graph copy %.3f | track: wall %.3f ms (gpu pose_only %.3f + deform %.3f = %.3f; lib host %.3f stage %.3f)" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), s0["gpu_ms"], s1["gpu_ms"], s0["gpu_ms"] + s1["gpu_ms"],
        s1["host_ms"], s1["stage_ms"]), flush=True)
core.close()
