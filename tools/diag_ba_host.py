"""Where does the end-to-end time of a configs[2] BA call go? (host graph build / staging vs device vs copies)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
q = synth.ba_problem(sys.argv[1] if len(sys.argv) > 1 else "c3")
for rep in range(3):
    t0 = time.perf_counter()
    r = core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
    t1 = time.perf_counter()
    s = r["stats"]
    print("rep %d wall %.2f ms | lib host %.2f stage %.2f gpu %.2f h2d %d d2h %d" % (
        rep, 1e3 * (t1 - t0), s["host_ms"], s["stage_ms"], s["gpu_ms"], s["h2d_bytes"], s["d2h_bytes"]))
core.close()
