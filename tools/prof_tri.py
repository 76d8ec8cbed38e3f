"""Profiling driver: one frame's batch of DeformableTriangulation candidates, run three times (used under ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
b = synth.triangulation_batch(seed=21, n_cand=600, fail_frac=0.1)
t = api.Triangulator(core)
for _ in range(3):
    r = t.run_batch(b)
print("ok", int((r["status"] == 0).sum()), "of", b["n_cand"], "device ms", t.rerun())
t.close()
core.close()
