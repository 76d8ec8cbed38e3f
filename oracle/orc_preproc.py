"""oracle/orc_preproc.py — CPU restatement (numpy) of the per-frame image pre-processing. TEST INFRASTRUCTURE ONLY:
imported by tests/ (and nothing else); the product path is nr-slam_b200/csrc/nrs_pre.cu.

What it restates (reference paths relative to /root/reference):
  System::ImageProcessing            modules/SLAM/system.cc:189-201    cv::cvtColor(RGB2GRAY) + cv::CLAHE(3.0, 8x8)
  BrightFilter::generateMask         modules/masking/bright_filter.cc:24-39
  BorderFilter::generateMask         modules/masking/border_filter.cc:24-40
  Masker::mask / GetAllMasks["Global"]  modules/masking/masker.cc:80-92,94-115
The arithmetic lives in OpenCV 4 (un-vendored, unpinned: modules/CMakeLists.txt:3). Restated from OpenCV's published
algorithms and PINNED bit-exactly on golden vectors produced by cv2 4.13 in the authoring container
(tests/golden/preproc.npz, generator tests/golden/make_preproc_golden.py):
  cvtColor 8u RGB2GRAY   fixed point, 15 fractional bits: (9798 R + 19235 G + 3735 B + 2^14) >> 15
  CLAHE                  per-tile histogram, clip at max(1, int(clip * area / 256)), excess redistributed (batch +
                         residual with stride), LUT = round(cdf * 255 / area), fp32 bilinear blend of the 4 tile LUTs
                         ((l11 xa1 + l12 xa) ya1 + (l21 xa1 + l22 xa) ya, no fused multiply-add), round, saturate
  erode                  min over the structuring element, anchor at the centre (size / 2), outside pixels ignored
  GaussianBlur 11x11 s=5 8-bit fixed point: kernel [17 20 24 26 27 28 27 26 24 20 17] / 256 (error-diffused rounding
                         of the normalised Gaussian), REFLECT_101, (sum + 2^15) >> 16 after both passes
"""
import numpy as np

GAUSS11_S5 = np.array([17, 20, 24, 26, 27, 28, 27, 26, 24, 20, 17], np.int64)


def rgb2gray(im):
    r, g, b = (im[..., k].astype(np.int32) for k in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def clahe(src, clip=3.0, tiles=(8, 8)):
    h, w = src.shape
    tx, ty = tiles
    ew = w if w % tx == 0 else w + (tx - w % tx)
    eh = h if h % ty == 0 else h + (ty - h % ty)
    if (ew, eh) != (w, h):  # copyMakeBorder(0, pad_b, 0, pad_r, BORDER_REFLECT_101); both axes padded if either is ragged
        ew, eh = w + (tx - w % tx), h + (ty - h % ty)
        ext = src[_reflect101(np.arange(eh), h)][:, _reflect101(np.arange(ew), w)]
    else:
        ext = src
    tw, th = ew // tx, eh // ty
    area = tw * th
    lut_scale = np.float32(255.0) / np.float32(area)
    clip_limit = 0
    if clip > 0:
        clip_limit = max(int(np.float32(clip) * np.float32(area) / np.float32(256)), 1)
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int32)
            if clip_limit > 0:
                clipped = int(np.maximum(hist - clip_limit, 0).sum())
                hist = np.minimum(hist, clip_limit)
                batch = clipped // 256
                resid = clipped - batch * 256
                hist += batch
                if resid != 0:
                    step = max(256 // resid, 1)
                    k = 0
                    while k < 256 and resid > 0:
                        hist[k] += 1
                        k += step
                        resid -= 1
            luts[j, i] = np.clip(np.rint(np.cumsum(hist).astype(np.float32) * lut_scale), 0, 255).astype(np.uint8)
    inv_tw, inv_th = np.float32(1.0) / np.float32(tw), np.float32(1.0) / np.float32(th)
    xs = np.arange(w, dtype=np.float32) * inv_tw - np.float32(0.5)
    tx1 = np.floor(xs).astype(np.int32)
    xa = (xs - tx1).astype(np.float32)
    xa1 = np.float32(1) - xa
    tx2 = np.minimum(tx1 + 1, tx - 1)
    tx1 = np.maximum(tx1, 0)
    ys = np.arange(h, dtype=np.float32) * inv_th - np.float32(0.5)
    ty1 = np.floor(ys).astype(np.int32)
    ya = (ys - ty1).astype(np.float32)
    ya1 = np.float32(1) - ya
    ty2 = np.minimum(ty1 + 1, ty - 1)
    ty1 = np.maximum(ty1, 0)
    v = src.astype(np.int64)
    Y1, X1 = np.meshgrid(ty1, tx1, indexing="ij")
    Y2, X2 = np.meshgrid(ty2, tx2, indexing="ij")
    l11, l12 = luts[Y1, X1, v].astype(np.float32), luts[Y1, X2, v].astype(np.float32)
    l21, l22 = luts[Y2, X1, v].astype(np.float32), luts[Y2, X2, v].astype(np.float32)
    XA, XA1, YA, YA1 = xa[None, :], xa1[None, :], ya[:, None], ya1[:, None]
    res = (l11 * XA1 + l12 * XA) * YA1 + (l21 * XA1 + l22 * XA) * YA
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def ellipse_half_widths(size):
    """cv::getStructuringElement(MORPH_ELLIPSE, (size, size)): per row the half width dx of the run [c - dx, c + dx]."""
    r = c = size // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    out = []
    for i in range(size):
        dy = i - r
        out.append(int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2))) if abs(dy) <= r else -1)
    return out


def erode(img, element):
    """element: 2-D 0/1 array; anchor at (cols // 2, rows // 2); pixels outside the image do not constrain the min."""
    kh, kw = element.shape
    ay, ax = kh // 2, kw // 2
    h, w = img.shape
    p = np.full((h + kh, w + kw), 255, np.uint8)
    p[ay:ay + h, ax:ax + w] = img
    out = np.full((h, w), 255, np.uint8)
    for dy in range(kh):
        for dx in range(kw):
            if element[dy, dx]:
                out = np.minimum(out, p[dy:dy + h, dx:dx + w])
    return out


def rect(size):
    return np.ones((size, size), np.uint8)


def ellipse(size):
    e = np.zeros((size, size), np.uint8)
    c = size // 2
    for i, dx in enumerate(ellipse_half_widths(size)):
        if dx >= 0:
            e[i, max(c - dx, 0):min(c + dx + 1, size)] = 1
    return e


def gauss11(img):
    h, w = img.shape
    p = img.astype(np.int64)[_reflect101(np.arange(-5, h + 5), h)][:, _reflect101(np.arange(-5, w + 5), w)]
    t = np.zeros((h + 10, w), np.int64)
    for i in range(11):
        t += p[:, i:i + w] * GAUSS11_S5[i]
    v = np.zeros((h, w), np.int64)
    for i in range(11):
        v += t[i:i + h, :] * GAUSS11_S5[i]
    return ((v + (1 << 15)) >> 16).clip(0, 255).astype(np.uint8)


def bright_filter(gray, th):
    m = np.where(gray > th, 0, 255).astype(np.uint8)       # THRESH_BINARY_INV
    return gauss11(erode(m, ellipse(11)))


def border_filter(gray, rb, re, cb, ce):
    h, w = gray.shape
    m = np.zeros((h, w), np.uint8)
    m[rb:h - re, cb:w - ce] = 255
    m[gray == 0] = 0
    return erode(m, rect(21))


def global_mask(gray, filters):
    """filters: list of ("bright", th) / ("border", rb, re, cb, ce) / ("predefined", mask)."""
    m = np.full(gray.shape, 255, np.uint8)
    for f in filters:
        if f[0] == "bright":
            m &= bright_filter(gray, f[1])
        elif f[0] == "border":
            m &= border_filter(gray, *f[1:5])
        else:
            m &= f[1]
    return erode(m, rect(10))
