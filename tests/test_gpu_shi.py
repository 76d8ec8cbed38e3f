"""Parity of the CUDA Shi-Tomasi detector against the oracle's clean definition: score map, keypoints, raster order
and class ids are BIT-EXACT (integer tensor sums, one fp32 rounding per operation in the same order)."""
import numpy as np
import pytest

import oracle_lib
from nrslam_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,seed", [((320, 240), 71), ((640, 480), 72), ((333, 257), 73)])
def test_scores_and_keypoints_bit_exact(core, size, seed):
    img = synth.klt_pair(seed=seed, size=size, n_points=1)["ref"]
    a = oracle_lib.OracleShiTomasi()
    b = api.ShiTomasi(core)
    ra = a.extract(img, want_scores=True)
    rb = b.extract(img, want_scores=True)
    assert np.array_equal(ra["scores"], rb["scores"])
    assert ra["n"] == rb["n"] and ra["n"] > 30
    assert np.array_equal(ra["xy"], rb["xy"]) and np.array_equal(ra["ids"], rb["ids"])
    # second call on the same extractor with already-tracked keypoints: marks, exclusion window, running ids
    keep = ra["xy"][::3] + np.float32(0.3)
    ra2 = a.extract(img, existing=keep, want_scores=True)
    rb2 = b.extract(img, existing=keep, want_scores=True)
    assert np.array_equal(ra2["scores"], rb2["scores"]) and (rb2["scores"] == -1).sum() == len(keep)
    assert np.array_equal(ra2["xy"], rb2["xy"]) and np.array_equal(ra2["ids"], rb2["ids"])
    assert rb2["ids"][0] == ra["n"]
    b.close()


def test_flat_image_and_capacity(core):
    b = api.ShiTomasi(core)
    r = b.extract(np.full((240, 320), 127, np.uint8))
    assert r["n"] == 0
    img = synth.klt_pair(seed=74, size=(320, 240), n_points=1)["ref"]
    full = b.extract(img)
    few = api.ShiTomasi(core).extract(img, capacity=5)
    assert few["n"] == full["n"] and len(few["xy"]) == 5 and np.array_equal(few["xy"], full["xy"][:5])
    b.close()


def test_detect_then_track_pipeline(core):
    """KF creation path (tracking.cc:350-392): detect on the reference image, seed the KLT with the detections, track."""
    p = synth.klt_pair(seed=75, n_points=1, shift=(2.2, 1.4))
    det = api.ShiTomasi(core).extract(p["ref"])
    m = (det["xy"][:, 0] > 40) & (det["xy"][:, 0] < 600) & (det["xy"][:, 1] > 40) & (det["xy"][:, 1] < 440)
    pts = det["xy"][m]
    assert len(pts) > 100
    k = api.KLT(core)
    k.set_reference(p["ref"], pts)
    from nrslam_b200 import abi
    r = k.track(p["cur"], pts, np.full(len(pts), abi.TRACKED, np.uint8))
    ok = r["status"] == abi.TRACKED
    assert ok.mean() > 0.9
    assert np.median(np.abs(r["pts"][ok] - (pts[ok] + np.array([2.2, 1.4], np.float32)))) < 0.1
    k.close()
