"""Generates tests/golden/ba_full_{c3,c4}.npz: the ORACLE's result of LocalDeformableBundleAdjustment on the FULL-SIZE
windows of BASELINE.json configs[2] (5000 landmarks / 30 keyframes / 50k observations) and configs[3] (20000 / 100 /
200k, KannalaBrandt8). The oracle's sparse Cholesky (plain minimum degree, scalar up-looking) does not finish such
windows in reasonable time, so the damped systems are solved by its converged-CG path (ORC_SOLVER=cg: fp64, relative
residual 1e-14 — an exact solve up to rounding; tests/test_oracle_drivers.py pins it on the factorisation).
Everything else (edges, LM control flow, graph construction) is the same restatement the small parity tests use.

    python tests/golden/make_ba_full.py c3 [c4]

Stored: keyframe poses (all), every `stride`-th point (stride 8 for c3, 16 for c4), the accepted chi2 trace and the LM
iteration / trial counts. tests/test_gpu_parity.py compares the CUDA path with these on the same seeded problems."""
import os
import sys
import time

os.environ["ORC_SOLVER"] = "cg"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib  # noqa: E402
from nrslam_b200 import synth  # noqa: E402

STRIDE = {"c3": 8, "c4": 16}

for cfg in sys.argv[1:] or ["c3"]:
    q = synth.ba_problem(cfg)
    o = oracle_lib.Oracle()
    t0 = time.time()
    r = o.local_ba(q["cam"], q["kf_pose"], q["obs_kf"], q["obs_vertex"], q["uv"], q["X"], q["graph"], q["scale"])
    dt = time.time() - t0
    st = r["stats"]
    out = os.path.join(ROOT, "tests", "golden", "ba_full_%s.npz" % cfg)
    np.savez_compressed(out, kf_pose=r["kf_pose"].astype(np.float32), X_sub=r["X"][::STRIDE[cfg]].astype(np.float32),
                        stride=np.int32(STRIDE[cfg]), chi2_trace=np.array(st["chi2_trace"], np.float64),
                        lm_iterations=np.int32(st["lm_iterations"]), lm_trials=np.int32(st["lm_trials"]),
                        n_obs=np.int32(len(q["obs_kf"])), seconds=np.float32(dt))
    print(cfg, "obs", len(q["obs_kf"]), "iterations", st["lm_iterations"], "trials", st["lm_trials"], "%.1f s" % dt,
          "->", out, os.path.getsize(out), "bytes")
