"""KLT oracle: the pyramid restatement is pinned bit-exactly on cv2.buildOpticalFlowPyramid golden vectors
(tests/golden/klt_pyramid.npz, generator make_klt_golden.py); the tracker (unpinned by the reference, which ships no
tests) is checked through properties: identity tracking, recovery of a known sub-pixel shift under gain/bias change,
status semantics at the image border, patch get / insert round trip."""
import os

import numpy as np
import pytest

import oracle_lib
from nrslam_b200 import abi, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "klt_pyramid.npz")


@pytest.mark.parametrize("name", ["even", "odd"])
def test_pyramid_matches_opencv_golden(name):
    oracle_lib.build()
    g = np.load(GOLD)
    img = g[name + "_image"]
    win = 21
    for lv in range(3):
        gi, gd = g["%s_L%d_img" % (name, lv)], g["%s_L%d_deriv" % (name, lv)]
        oi, od = oracle_lib.klt_pyramid(img, win, 2, lv)
        assert np.array_equal(oi[win:-win, win:-win], gi), "level %d image" % lv
        assert np.array_equal(od[win:-win, win:-win], gd), "level %d Scharr derivative" % lv
        # borders the reference reads through negative offsets: REFLECT_101 image, constant-0 derivative
        assert np.array_equal(oi, np.pad(gi, win, mode="reflect"))
        assert np.array_equal(od, np.pad(gd, ((win, win), (win, win), (0, 0))))


def test_identity_tracking_and_counts():
    oracle_lib.build()
    p = synth.klt_pair(seed=3, n_points=150, shift=(0.0, 0.0), gain=1.0, bias=0.0, noise=0.0)
    k = oracle_lib.OracleKLT()
    assert k.set_reference(p["ref"], p["pts"]) == 0
    r = k.track(p["ref"], p["pts"], p["status"])
    ok = r["status"] == abi.TRACKED
    assert ok.mean() > 0.95
    assert np.abs(r["pts"][ok] - p["pts"][ok]).max() < 1e-3
    assert r["n_tracked"] == int(ok.sum())


def test_recovers_shift_under_gain_and_bias():
    oracle_lib.build()
    p = synth.klt_pair(seed=5, n_points=200, shift=(3.3, -2.4), gain=1.15, bias=-12.0, noise=1.0)
    k = oracle_lib.OracleKLT()
    k.set_reference(p["ref"], p["pts"])
    r = k.track(p["cur"], p["pts"], p["status"])
    ok = r["status"] == abi.TRACKED
    assert ok.mean() > 0.8
    err = np.abs(r["pts"][ok] - p["pts_true"][ok])
    assert np.median(err) < 0.1 and np.percentile(err, 90) < 0.4


def test_border_points_and_unusable_statuses():
    oracle_lib.build()
    p = synth.klt_pair(seed=7, n_points=60)
    pts = p["pts"].copy()
    pts[0] = (1.0, 1.0)            # window mostly outside: no level-0 reference patch
    pts[1] = (p["ref"].shape[1] - 2.0, 50.0)
    st = p["status"].copy()
    st[2] = abi.BAD                # not usable: untouched, position untouched
    k = oracle_lib.OracleKLT()
    k.set_reference(p["ref"], pts)
    r = k.track(p["cur"], pts, st)
    assert r["status"][0] == abi.OUT_IMAGE_BOUNDARIES and r["status"][1] == abi.OUT_IMAGE_BOUNDARIES
    assert r["status"][2] == abi.BAD and np.array_equal(r["pts"][2], pts[2])


def test_patch_get_insert_roundtrip():
    oracle_lib.build()
    p = synth.klt_pair(seed=9, n_points=40)
    a = oracle_lib.OracleKLT()
    a.set_reference(p["ref"], p["pts"])
    b = oracle_lib.OracleKLT()
    b.set_reference(p["ref"], p["pts"][:0])
    for i in range(len(p["pts"])):
        b.insert_patch(p["pts"][i, 0], p["pts"][i, 1], a.get_patch(i))
    assert b.num_points() == a.num_points() == len(p["pts"])
    ra = a.track(p["cur"], p["pts"], p["status"])
    rb = b.track(p["cur"], p["pts"], p["status"])
    assert np.array_equal(ra["pts"], rb["pts"]) and np.array_equal(ra["status"], rb["status"])
    a.clear()
    assert a.num_points() == 0


def test_point_reuse_oracle_bookkeeping_and_gate():
    """oracle/orc_reuse.cc (Tracking::PointReuse, tracking.cc:394-506): points already in the frame are skipped unless
    the optimiser reported them lost, points behind the camera or outside the image never become candidates, a seed
    that the tracker moves by more than sqrt(5.99) px is rejected, the rest is re-found within a fraction of a pixel."""
    from nrslam_b200 import abi, synth
    im = synth.klt_pair(seed=51, n_points=80, shift=(1.2, -0.9))
    k = oracle_lib.OracleKLT()
    k.set_reference(im["ref"], im["pts"])
    patches = [k.get_patch(i) for i in range(80)]
    cam = abi.Camera.pinhole(520.0, 520.0, 320.0, 240.0)
    pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    tg = im["pts_true"].copy()
    tg[5] += np.array([6.0, 0.0], np.float32)          # seed 6 px off: KLT pulls it back -> reprojection gate rejects
    z = 3.0
    X = np.stack([(tg[:, 0] - 320) / 520 * z, (tg[:, 1] - 240) / 520 * z, np.full(len(tg), z)], 1).astype(np.float32)
    X[7, 2] = -1.0                                      # behind the camera
    X[9, 0] = 50.0                                      # projects outside the image
    in_frame = np.zeros(80, bool)
    in_frame[[2, 3]] = True
    forced = np.zeros(80, bool)
    forced[3] = True
    o = oracle_lib.point_reuse(cam, pose, im["cur"], X, patches, in_frame, forced)
    c = list(o["candidates"])
    assert 2 not in c and 3 in c and 7 not in c and 9 not in c and c == sorted(c)
    acc = dict(zip(c, o["accepted"]))
    assert not acc[5]
    ok = o["accepted"]
    assert ok.mean() > 0.8 and o["n_reused"] == int(ok.sum())
    assert np.abs(o["pts"][ok] - im["pts_true"][o["candidates"]][ok]).max() < 0.5
