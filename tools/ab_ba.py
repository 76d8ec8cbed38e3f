"""A/B of library builds on the same GPU: device time of the configs[2] / configs[3] BA launch (L2 flushed).
Usage: python tools/ab_ba.py lib1.so lib2.so ..."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import nrslam_b200  # noqa
    from nrslam_b200 import api, synth
    api.LIB_PATH = sys.argv[2]
    core = api.Core()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {}
    for cfg in ("c3", "c4"):
        q = synth.ba_problem(cfg)
        core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
        ms = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            ms.append(core.resolve(2)["gpu_ms"])
        out[cfg] = [round(m, 2) for m in ms]
    print(json.dumps({os.path.basename(sys.argv[2]): out}))
else:
    for rep in range(2):
        for lib in sys.argv[1:]:
            r = subprocess.run([sys.executable, __file__, "--child", os.path.abspath(lib)], capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
