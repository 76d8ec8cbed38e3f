"""Golden vectors for the image pre-processing (System::ImageProcessing, Masker), produced by OpenCV itself
(cv2 4.13, authoring container): python tests/golden/make_preproc_golden.py -> tests/golden/preproc.npz"""
import os

import cv2
import numpy as np

out = {"cv2_version": np.array(cv2.__version__)}
rng = np.random.default_rng(20261017)
for name, (h, w) in (("a", (96, 128)), ("b", (101, 135))):   # divisible / ragged tile grids
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 5)
    smooth = cv2.normalize(smooth, None, 0, 255, cv2.NORM_MINMAX)
    smooth[:6, :, :] = 0                 # a black endoscope border
    smooth[:, -9:, :] = 0
    for tag, rgb in (("noise", noise), ("smooth", smooth)):
        gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
        eq = cv2.createCLAHE(3.0, (8, 8)).apply(gray)
        th = 200
        bright = cv2.threshold(gray, th, 255, cv2.THRESH_BINARY_INV)[1]
        bright = cv2.erode(bright, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (11, 11)))
        bright = cv2.GaussianBlur(bright, (11, 11), 5, 5, borderType=cv2.BORDER_REFLECT_101)
        rb, re, cb, ce = 8, 6, 12, 10
        border = np.zeros((h, w), np.uint8)
        border[rb:h - re, cb:w - ce] = 255
        border[gray == 0] = 0
        border = cv2.erode(border, cv2.getStructuringElement(cv2.MORPH_RECT, (21, 21)))
        glob = cv2.bitwise_and(cv2.bitwise_and(np.full((h, w), 255, np.uint8), bright), border)
        glob = cv2.erode(glob, cv2.getStructuringElement(cv2.MORPH_RECT, (10, 10)))
        k = "%s_%s_" % (name, tag)
        out.update({k + "rgb": rgb, k + "gray": gray, k + "clahe": eq, k + "bright": bright, k + "border": border,
                    k + "global": glob})
out["ellipse11"] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (11, 11))
out["border_params"] = np.array([8, 6, 12, 10])
out["bright_th"] = np.array(200)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "preproc.npz"), **out)
print("wrote preproc.npz", {k: getattr(v, "shape", None) for k, v in out.items()})
