"""The reference-side shim (shim/g2o_optimization_b200.cc, shim/lucas_kanade_tracker_b200.cc) compiles and marshals
correctly. The image has no Eigen / Sophus / OpenCV / abseil headers, so the shim is compiled against the stand-in
declarations of shim/standin/ (same class / method names as modules/map/*.h, g2o_optimization.h:27-40,
lucas_kanade_tracker.h:55-92) and driven with fake Frame / Map / KeyFrame / TemporalBuffer objects:

  * CPU: linked against a recording mock of the C ABI that validates every buffer contract of include/nrslam_b200.h
    (CSR invariants, ascending vertex order, oldest-first keyframes, statuses) and returns recognisable results which
    must land exactly where the reference writes them (frame.cc, keyframe.cc, regularization_graph.cc:107-123, ...);
  * GPU: the same driver linked against libnrslam_b200.so runs a consistent synthetic scene through the CUDA path.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "shim", "g2o_optimization_b200.cc"), os.path.join(ROOT, "shim", "lucas_kanade_tracker_b200.cc"),
       os.path.join(ROOT, "tests", "shim", "standin_impl.cc"), os.path.join(ROOT, "tests", "shim", "driver.cc")]
INC = ["-I", os.path.join(ROOT, "shim", "standin"), "-I", os.path.join(ROOT, "include")]


def _build(out, extra):
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-o", out] + INC + SRC + extra)
    return out


def test_shim_compiles_and_marshals_through_a_mock_abi(tmp_path):
    exe = _build(str(tmp_path / "shim_mock"), [os.path.join(ROOT, "tests", "shim", "mock_abi.cc")])
    r = subprocess.run([exe, "mock"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 driver failures, 0 ABI-contract failures" in r.stdout


@pytest.mark.gpu
def test_shim_drives_the_real_library(tmp_path):
    lib = os.path.join(ROOT, "nr-slam_b200")
    exe = _build(str(tmp_path / "shim_real"), ["-L", lib, "-lnrslam_b200", "-Wl,-rpath," + lib])
    r = subprocess.run([exe, "real"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 driver failures" in r.stdout
