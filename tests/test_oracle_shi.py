"""Shi-Tomasi oracle: the clean definition (what the CUDA kernel implements) equals the literal simulation of the
reference wherever the reference's gradients are aligned, and both detect the same keypoints away from the border."""
import numpy as np

import oracle_lib
from nrslam_b200 import synth


def test_clean_equals_literal_in_the_interior():
    oracle_lib.build()
    img = synth.klt_pair(seed=61, size=(320, 240), n_points=1)["ref"]
    a = oracle_lib.OracleShiTomasi().extract(img, literal=True, want_scores=True)
    b = oracle_lib.OracleShiTomasi().extract(img, literal=False, want_scores=True)
    h, w = img.shape
    assert np.array_equal(a["scores"][4:h - 4, 1:w - 1], b["scores"][4:h - 4, 1:w - 1])
    assert np.all(b["scores"][:4] == 0) and np.all(b["scores"][h - 4:] == 0)
    assert np.all(b["scores"][:, 0] == 0) and np.all(b["scores"][:, w - 1] == 0)
    assert b["n"] > 50
    # keypoints farther than 4 + 15 px from the border are the same set in the same raster order
    def inner(r):
        m = (r["xy"][:, 0] >= 20) & (r["xy"][:, 0] < w - 20) & (r["xy"][:, 1] >= 20) & (r["xy"][:, 1] < h - 20)
        return r["xy"][m]
    assert np.array_equal(inner(a), inner(b))


def test_existing_keypoints_suppress_and_ids_run_on():
    oracle_lib.build()
    img = synth.klt_pair(seed=62, size=(320, 240), n_points=1)["ref"]
    s = oracle_lib.OracleShiTomasi()
    first = s.extract(img)
    assert np.array_equal(first["ids"], np.arange(first["n"]))
    assert np.all(np.diff(first["xy"][:, 1] * 1000 + first["xy"][:, 0]) > 0)        # raster order
    keep = first["xy"][::2]
    second = s.extract(img, existing=keep)
    assert second["ids"][0] == first["n"]                                            # the counter keeps running
    # no new keypoint within the 15-px exclusion window of an existing one
    d = np.abs(second["xy"][:, None, :] - keep[None, :, :]).max(axis=2)
    assert d.min() > 15
    # every dropped keypoint that is far from all kept ones comes back
    dropped = first["xy"][1::2]
    far = np.abs(dropped[:, None, :] - keep[None, :, :]).max(axis=2).min(axis=1) > 15
    got = {tuple(p) for p in second["xy"]}
    assert all(tuple(p) in got for p in dropped[far])
