"""Golden vectors for the KLT pyramid: cv2.buildOpticalFlowPyramid (OpenCV, the un-vendored third-party routine
lucas_kanade_tracker.cc:50,184 calls) run in the authoring container on seeded images.

    python tests/golden/make_klt_golden.py      ->  tests/golden/klt_pyramid.npz

The oracle's pyramid restatement (oracle/orc_klt.cc) and the CUDA pyramid kernels are checked bit-exactly against
these arrays. cv2 returns border-less views of every level; the winSize border the reference reads through negative
offsets is REFLECT_101 for the image and constant 0 for the derivative (buildOpticalFlowPyramid defaults) and is
checked separately with numpy.pad.
"""
import os

import cv2
import numpy as np


def texture(rng, h, w):
    """Band-limited noise + a few blobs, 8-bit."""
    img = rng.normal(size=(h, w)).astype(np.float32)
    img = cv2.GaussianBlur(img, (0, 0), 1.6) * 3 + cv2.GaussianBlur(img, (0, 0), 5.0) * 8
    img = (img - img.min()) / (img.max() - img.min())
    return np.clip(img * 255, 0, 255).astype(np.uint8)


def main():
    out = {}
    rng = np.random.default_rng(4242)
    for name, (h, w, levels) in {"even": (96, 120, 2), "odd": (101, 127, 2)}.items():
        img = texture(rng, h, w)
        n, pyr = cv2.buildOpticalFlowPyramid(img, (21, 21), levels)
        assert n == levels and len(pyr) == 2 * (levels + 1)
        out[name + "_image"] = img
        for lv in range(levels + 1):
            out["%s_L%d_img" % (name, lv)] = np.ascontiguousarray(pyr[2 * lv])
            out["%s_L%d_deriv" % (name, lv)] = np.ascontiguousarray(pyr[2 * lv + 1])
    out["opencv_version"] = np.array(cv2.__version__)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "klt_pyramid.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes, OpenCV", cv2.__version__)


if __name__ == "__main__":
    main()
