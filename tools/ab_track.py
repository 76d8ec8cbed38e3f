"""A/B of library builds on the same GPU: device time of the tracking launches over the bench seeds.
Usage: python tools/ab_track.py lib1.so lib2.so ..."""
import os, sys, subprocess, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import nrslam_b200  # noqa
    from nrslam_b200 import api, synth
    api.LIB_PATH = sys.argv[2]
    core = api.Core()
    out = {}
    for seed in range(1235, 1243):
        p = synth.tracking_problem("c2", seed=seed)
        args = (p["cam"], p["uv"], p["X_rest"], p["point_vertex"], p["vertex_frame_status"])
        for rep in range(3):
            r0, r1 = core.track_pose_and_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
        os.environ["NRSLAM_B200_PLAN_CACHE"] = "0"
        t0 = time.perf_counter()
        for rep in range(5):
            core.track_pose_and_deform(*args, p["graph"].copy(), p["scale"], p["seed_pose"], p["last_world_position"])
        cold = (time.perf_counter() - t0) / 5 * 1e3
        os.environ.pop("NRSLAM_B200_PLAN_CACHE")
        out[seed] = [round(r1["stats"]["gpu_ms"], 2), round(cold, 2)]
    print(json.dumps({os.path.basename(sys.argv[2]): out, "mean_gpu": round(sum(v[0] for v in out.values()) / len(out), 3),
                      "mean_cold_e2e": round(sum(v[1] for v in out.values()) / len(out), 3)}))
else:
    for lib in sys.argv[1:]:
        r = subprocess.run([sys.executable, __file__, "--child", os.path.abspath(lib)], capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
