"""nrslam_b200 — B200-native optimisation core behind NR-SLAM's g2o_optimization.h / LucasKanadeTracker seam.

The product is the C-ABI shared library built from csrc/ (include/nrslam_b200.h); this Python package is only
the ctypes binding used by tests/, bench.py and __graft_entry__.py, plus the synthetic-problem generators.
There is no CPU fallback: `load()` raises if libnrslam_b200.so is missing.
"""
from . import abi  # noqa: F401
