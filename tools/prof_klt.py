"""Profiling driver: SetReferenceImage + 3 Track calls of 2000 points on a 640x480 pair (used under ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nrslam_b200  # noqa
from nrslam_b200 import api, synth
core = api.Core()
p = synth.klt_pair(seed=41, n_points=2000)
k = api.KLT(core)
k.set_reference(p["ref"], p["pts"])
for _ in range(3):
    r = k.track(p["cur"], p["pts"], p["status"])
print("tracked", r["n_tracked"], "device ms", k.retrack())
k.close()
core.close()
