"""A/B of environment switches on the same GPU: device time of the configs[2] / configs[3] BA launch (L2 flushed).
Usage: python tools/ab_env.py "A=1" "A=0 B=2" ... ("-" = no override)"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    import nrslam_b200  # noqa
    from nrslam_b200 import api, synth
    core = api.Core()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {}
    for cfg in ("c3", "c4"):
        q = synth.ba_problem(cfg)
        core.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
        ms = []
        for _ in range(4):
            flush.zero_(); torch.cuda.synchronize()
            ms.append(core.resolve(2)["gpu_ms"])
        out[cfg] = [round(m, 2) for m in ms]
    print(json.dumps(out))
else:
    for rep in range(2):
        for v in sys.argv[1:]:
            env = dict(os.environ)
            if v != "-":
                env.update(dict(kv.split("=") for kv in v.split()))
            r = subprocess.run([sys.executable, __file__, "--child"], capture_output=True, text=True, env=env)
            print(v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
