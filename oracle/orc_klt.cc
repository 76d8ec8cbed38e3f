// ORACLE — TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this file's library; the product path never does.
//
// CPU restatement of NR-SLAM's KLT tracker:
//   LucasKanadeTracker::SetReferenceImage                modules/matching/lucas_kanade_tracker.cc:47-168
//   LucasKanadeTracker::Track (incl. the SSIM gate)      modules/matching/lucas_kanade_tracker.cc:170-596
//   Get/InsertPhotometricInformation, clear              modules/matching/lucas_kanade_tracker.cc:598-631
// and of the third-party routine it calls, which is NOT in /root/reference:
//   cv::buildOpticalFlowPyramid (OpenCV 4, unpinned: modules/CMakeLists.txt:3, README.md:47 "tested 3.2.0, 4.4.0"):
//   pyrDown (5-tap [1 4 6 4 1]/16 separable, (sum + 128) >> 8, BORDER_REFLECT_101), calcSharrDeriv (3-10-3 Scharr,
//   int16 x2 interleaved, REFLECT_101 inside the level), levels stored with a winSize border (image: REFLECT_101,
//   derivative: constant 0).
// Pinning: the pyramid restatement is checked bit-exactly against cv2.buildOpticalFlowPyramid 4.13 run in the
// authoring container (tests/golden/klt_pyramid_*.npz, generator tests/golden/make_klt_golden.py). The tracker
// itself has no reference test or fixture: PARITY UNPINNED for Track / SetReferenceImage.
//
// Deviation, documented: the reference indexes the mask at full-resolution coordinates without a bounds check
// (lucas_kanade_tracker.cc:125-131,521-526: undefined behaviour for windows overlapping the image border); here
// an out-of-range mask coordinate reads as 0 (masked) in SetReferenceImage and as "valid" in the SSIM stage (where
// the value is never used, :521-526).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../include/nrslam_b200.h"

namespace {

inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

inline int cv_floor(float v) { return (int)floorf(v); }
inline int cv_round(float v) { return (int)lrintf(v); }  // round-half-even under the default rounding mode
inline int cv_round_d(double v) { return (int)lrint(v); }
#define ORC_DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

struct Level {
  int w = 0, h = 0, border = 0;
  std::vector<uint8_t> img;    // (h + 2b) x (w + 2b)
  std::vector<int16_t> deriv;  // (h + 2b) x (w + 2b) x 2
  int stride() const { return w + 2 * border; }
  const uint8_t* I(int y) const { return img.data() + (size_t)(y + border) * stride() + border; }
  const int16_t* D(int y) const { return deriv.data() + ((size_t)(y + border) * stride() + border) * 2; }
};

// cv::pyrDown, 8-bit single channel
void pyr_down(const Level& src, Level& dst) {
  const int sw = src.w, sh = src.h, dw = (sw + 1) / 2, dh = (sh + 1) / 2;
  dst.w = dw;
  dst.h = dh;
  std::vector<int> hrow((size_t)sh * dw);
  for (int y = 0; y < sh; y++) {
    const uint8_t* s = src.I(y);
    for (int x = 0; x < dw; x++) {
      const int x0 = reflect101(2 * x - 2, sw), x1 = reflect101(2 * x - 1, sw), x2 = reflect101(2 * x, sw),
                x3 = reflect101(2 * x + 1, sw), x4 = reflect101(2 * x + 2, sw);
      hrow[(size_t)y * dw + x] = s[x0] + 4 * s[x1] + 6 * s[x2] + 4 * s[x3] + s[x4];
    }
  }
  const int b = dst.border;
  dst.img.assign((size_t)(dh + 2 * b) * (dw + 2 * b), 0);
  for (int y = 0; y < dh; y++) {
    const int y0 = reflect101(2 * y - 2, sh), y1 = reflect101(2 * y - 1, sh), y2 = reflect101(2 * y, sh),
              y3 = reflect101(2 * y + 1, sh), y4 = reflect101(2 * y + 2, sh);
    uint8_t* d = dst.img.data() + (size_t)(y + b) * dst.stride() + b;
    for (int x = 0; x < dw; x++) {
      const int v = hrow[(size_t)y0 * dw + x] + 4 * hrow[(size_t)y1 * dw + x] + 6 * hrow[(size_t)y2 * dw + x] +
                    4 * hrow[(size_t)y3 * dw + x] + hrow[(size_t)y4 * dw + x];
      d[x] = (uint8_t)((v + 128) >> 8);
    }
  }
}

void fill_image_border(Level& L) {  // copyMakeBorder(..., BORDER_REFLECT_101 | BORDER_ISOLATED)
  const int b = L.border, st = L.stride();
  for (int y = -b; y < L.h + b; y++) {
    const int sy = reflect101(y, L.h);
    uint8_t* d = L.img.data() + (size_t)(y + b) * st + b;
    const uint8_t* s = L.img.data() + (size_t)(sy + b) * st + b;
    for (int x = -b; x < L.w + b; x++) {
      if (y >= 0 && y < L.h && x >= 0 && x < L.w) continue;
      d[x] = s[reflect101(x, L.w)];
    }
  }
}

// calcSharrDeriv + constant-0 border
void scharr(Level& L) {
  const int b = L.border, st = L.stride(), w = L.w, h = L.h;
  L.deriv.assign((size_t)(h + 2 * b) * st * 2, 0);
  std::vector<int> t0(w + 2), t1(w + 2);
  for (int y = 0; y < h; y++) {
    const uint8_t* r0 = L.I(y > 0 ? y - 1 : (h > 1 ? 1 : 0));
    const uint8_t* r1 = L.I(y);
    const uint8_t* r2 = L.I(y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0));
    for (int x = 0; x < w; x++) {
      t0[x + 1] = (r0[x] + r2[x]) * 3 + r1[x] * 10;
      t1[x + 1] = r2[x] - r0[x];
    }
    const int x0 = (w > 1 ? 1 : 0), x1 = (w > 1 ? w - 2 : 0);
    t0[0] = t0[x0 + 1];
    t0[w + 1] = t0[x1 + 1];
    t1[0] = t1[x0 + 1];
    t1[w + 1] = t1[x1 + 1];
    int16_t* d = L.deriv.data() + ((size_t)(y + b) * st + b) * 2;
    for (int x = 0; x < w; x++) {
      d[2 * x] = (int16_t)(t0[x + 2] - t0[x]);
      d[2 * x + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
    }
  }
}

// cv::buildOpticalFlowPyramid(img, pyr, winSize, maxLevel) with the defaults the reference uses
// (lucas_kanade_tracker.cc:50,184). Returns the number of levels built (the reference ignores the return value and
// would read past the pyramid when a level gets smaller than the window; callers here refuse such sizes).
int build_pyramid(const uint8_t* img, int w, int h, int pitch, int win, int max_level, std::vector<Level>& pyr) {
  pyr.clear();
  pyr.resize(max_level + 1);
  int lw = w, lh = h;
  for (int level = 0; level <= max_level; level++) {
    Level& L = pyr[level];
    L.border = win;
    if (level == 0) {
      L.w = w;
      L.h = h;
      L.img.assign((size_t)(h + 2 * win) * (w + 2 * win), 0);
      for (int y = 0; y < h; y++) memcpy(L.img.data() + (size_t)(y + win) * L.stride() + win, img + (size_t)y * pitch, w);
    } else {
      pyr_down(pyr[level - 1], L);
    }
    fill_image_border(L);
    scharr(L);
    lw = (lw + 1) / 2;
    lh = (lh + 1) / 2;
    if (lw <= win || lh <= win) {
      pyr.resize(level + 1);
      return level + 1;
    }
  }
  return max_level + 1;
}

inline bool is_usable(uint8_t s) {  // utilities/landmark_status.cc:21-23
  return s == NRSLAM_TRACKED_WITH_3D || s == NRSLAM_TRACKED || s == NRSLAM_JUST_TRIANGULATED;
}

struct Patch {
  bool valid = false;
  float mean = -1.f, mean2 = -1.f;
  std::vector<int16_t> gray, grad;  // win*win, win*win*2
};

struct Tracker {
  int win = 21, max_level = 4, max_iters = 10;
  float eps = 1e-4f, min_eig = 1e-4f;
  std::vector<float> prev;                  // 2 per point
  std::vector<std::vector<Patch>> patches;  // [level][point]
};

}  // namespace

extern "C" {

void* orc_klt_create(int win, int max_level, int max_iters, float eps, float min_eig) {
  Tracker* t = new Tracker();
  t->win = win;
  t->max_level = max_level;
  t->max_iters = max_iters;
  t->eps = eps;
  t->min_eig = min_eig;
  t->patches.resize(max_level + 1);
  return t;
}
void orc_klt_destroy(void* p) { delete static_cast<Tracker*>(p); }
int orc_klt_num_points(void* p) { return (int)(static_cast<Tracker*>(p)->prev.size() / 2); }

// Pinning hook: bordered level image and derivative of the restated pyramid.
int orc_klt_pyramid(const uint8_t* img, int w, int h, int pitch, int win, int max_level, int level, uint8_t* img_out,
                    int16_t* deriv_out, int* lw, int* lh) {
  std::vector<Level> pyr;
  const int n = build_pyramid(img, w, h, pitch, win, max_level, pyr);
  if (level >= n) return -1;
  const Level& L = pyr[level];
  *lw = L.w;
  *lh = L.h;
  if (img_out) memcpy(img_out, L.img.data(), L.img.size());
  if (deriv_out) memcpy(deriv_out, L.deriv.data(), L.deriv.size() * sizeof(int16_t));
  return n;
}

// SetReferenceImage — lucas_kanade_tracker.cc:47-168
int orc_klt_set_reference(void* p, const uint8_t* img, int w, int h, int pitch, int n, const float* pts,
                          const uint8_t* mask, int mask_pitch) {
  Tracker& T = *static_cast<Tracker*>(p);
  std::vector<Level> pyr;
  if (build_pyramid(img, w, h, pitch, T.win, T.max_level, pyr) != T.max_level + 1) return NRSLAM_B200_ERR_ARG;
  T.prev.assign(pts, pts + 2 * (size_t)n);
  const int win = T.win;
  const float half = (win - 1) * 0.5f;
  const int borderGap = (int)round((double)(win / 2));  // :58 round(winSize_.width/2), integer division first
  for (int level = T.max_level; level >= 0; level--) {
    T.patches[level].assign(n, Patch());
    const Level& L = pyr[level];
    const int st = L.stride();
    const int scale = 1 << level;
    for (int i = 0; i < n; i++) {
      float px = pts[2 * i] / (float)(1 << level) - half, py = pts[2 * i + 1] / (float)(1 << level) - half;
      const int ix = cv_floor(px), iy = cv_floor(py);
      if (ix < -borderGap || ix >= L.w - borderGap || iy < -borderGap || iy >= L.h - borderGap) continue;
      const float a = px - ix, b = py - iy;
      const int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << 14));
      const int iw01 = cv_round(a * (1.f - b) * (1 << 14));
      const int iw10 = cv_round((1.f - a) * b * (1 << 14));
      const int iw11 = (1 << 14) - iw00 - iw01 - iw10;
      Patch pt;
      pt.gray.resize((size_t)win * win);
      pt.grad.resize((size_t)win * win * 2);
      float meanI = 0.f, meanI2 = 0.f;
      bool valid = true;
      for (int y = 0; y < win && valid; y++) {
        const uint8_t* src = L.I(y + iy) + ix;
        const int16_t* dsrc = L.D(y + iy) + ix * 2;
        for (int x = 0; x < win; x++, dsrc += 2) {
          if (mask) {
            const int mx = (ix + x) * scale, my = (iy + y) * scale;
            const bool in = mx >= 0 && mx < w && my >= 0 && my < h;
            if (!in || mask[(size_t)my * mask_pitch + mx] == 0) {
              valid = false;
              break;
            }
          }
          const int ival = ORC_DESCALE(src[x] * iw00 + src[x + 1] * iw01 + src[x + st] * iw10 + src[x + st + 1] * iw11, 14 - 5);
          const int ixval = ORC_DESCALE(dsrc[0] * iw00 + dsrc[2] * iw01 + dsrc[2 * st] * iw10 + dsrc[2 * st + 2] * iw11, 14);
          const int iyval = ORC_DESCALE(dsrc[1] * iw00 + dsrc[3] * iw01 + dsrc[2 * st + 1] * iw10 + dsrc[2 * st + 3] * iw11, 14);
          pt.gray[(size_t)y * win + x] = (int16_t)ival;
          pt.grad[((size_t)y * win + x) * 2] = (int16_t)ixval;
          pt.grad[((size_t)y * win + x) * 2 + 1] = (int16_t)iyval;
          meanI += (float)ival;
          meanI2 += (float)(ival * ival);
        }
      }
      if (!valid) continue;
      const float FLT_SCALE = 1.f / (1 << 20);
      pt.mean = (meanI * FLT_SCALE) / (win * win);
      pt.mean2 = (meanI2 * FLT_SCALE) / (win * win);
      pt.valid = true;
      T.patches[level][i] = std::move(pt);
    }
  }
  return 0;
}

// Track — lucas_kanade_tracker.cc:170-596
int orc_klt_track(void* p, const uint8_t* img, int w, int h, int pitch, int n, float* pts_io, uint8_t* status_io,
                  int use_initial_flow, float min_ssim, const uint8_t* mask, int mask_pitch, int* n_tracked_out) {
  Tracker& T = *static_cast<Tracker*>(p);
  (void)mask;
  (void)mask_pitch;
  if (n != (int)(T.prev.size() / 2)) return NRSLAM_B200_ERR_ARG;
  std::vector<Level> pyr;
  if (build_pyramid(img, w, h, pitch, T.win, T.max_level, pyr) != T.max_level + 1) return NRSLAM_B200_ERR_ARG;
  const int win = T.win, area = win * win;
  const float half = (win - 1) * 0.5f;
  const int borderGap = (int)round((double)(win / 2)) + 1;  // :186
  const float FLT_SCALE = 1.f / (1 << 20);
  std::vector<int16_t> Jw(area), dJw(2 * (size_t)area);
  for (int level = T.max_level; level >= 0; level--) {
    const Level& L = pyr[level];
    const int st = L.stride();
    for (int i = 0; i < n; i++) {
      if (!is_usable(status_io[i])) continue;
      float prevx = T.prev[2 * i] * (float)(1. / (1 << level)), prevy = T.prev[2 * i + 1] * (float)(1. / (1 << level));
      float nx, ny;
      if (level == T.max_level) {
        if (use_initial_flow) {
          nx = pts_io[2 * i] * (float)(1. / (1 << level));
          ny = pts_io[2 * i + 1] * (float)(1. / (1 << level));
        } else {
          nx = prevx;
          ny = prevy;
        }
      } else {
        nx = pts_io[2 * i] * 2.f;
        ny = pts_io[2 * i + 1] * 2.f;
      }
      pts_io[2 * i] = nx;
      pts_io[2 * i + 1] = ny;
      prevx -= half;
      prevy -= half;
      const int ipx = cv_floor(prevx), ipy = cv_floor(prevy);
      if (ipx < -borderGap || ipx >= L.w - borderGap || ipy < -borderGap || ipy >= L.h - borderGap) {
        if (level == 0) status_io[i] = NRSLAM_OUT_IMAGE_BOUNDARIES;
        continue;
      }
      const Patch& ref = T.patches[level][i];
      if (!ref.valid) {
        if (level == 0) status_io[i] = NRSLAM_OUT_IMAGE_BOUNDARIES;
        continue;
      }
      const float meanI = ref.mean, meanI2 = ref.mean2;
      const float startx = nx, starty = ny;
      float pdx = 0.f, pdy = 0.f;
      nx -= half;
      ny -= half;
      for (int j = 0; j < T.max_iters; j++) {
        const int inx = cv_floor(nx), iny = cv_floor(ny);
        if (inx < -borderGap || inx >= L.w - borderGap || iny < -borderGap || iny >= L.h - borderGap) {
          if (level == 0) status_io[i] = NRSLAM_OUT_IMAGE_BOUNDARIES;
          break;
        }
        const float aJ = nx - inx, bJ = ny - iny;
        const int jw00 = cv_round((1.f - aJ) * (1.f - bJ) * (1 << 14));
        const int jw01 = cv_round(aJ * (1.f - bJ) * (1 << 14));
        const int jw10 = cv_round((1.f - aJ) * bJ * (1 << 14));
        const int jw11 = (1 << 14) - jw00 - jw01 - jw10;
        float meanJ = 0.f, meanJ2 = 0.f;
        for (int y = 0; y < win; y++) {
          const uint8_t* src = L.I(y + iny) + inx;
          const int16_t* dsrc = L.D(y + iny) + inx * 2;
          for (int x = 0; x < win; x++, dsrc += 2) {
            const int jval = ORC_DESCALE(src[x] * jw00 + src[x + 1] * jw01 + src[x + st] * jw10 + src[x + st + 1] * jw11, 14 - 5);
            const int jxval = ORC_DESCALE(dsrc[0] * jw00 + dsrc[2] * jw01 + dsrc[2 * st] * jw10 + dsrc[2 * st + 2] * jw11, 14);
            const int jyval = ORC_DESCALE(dsrc[1] * jw00 + dsrc[3] * jw01 + dsrc[2 * st + 1] * jw10 + dsrc[2 * st + 3] * jw11, 14);
            Jw[(size_t)y * win + x] = (int16_t)jval;
            dJw[((size_t)y * win + x) * 2] = (int16_t)jxval;
            dJw[((size_t)y * win + x) * 2 + 1] = (int16_t)jyval;
            meanJ += (float)jval;
            meanJ2 += (float)(jval * jval);
          }
        }
        meanJ = (meanJ * FLT_SCALE) / area;
        meanJ2 = (meanJ2 * FLT_SCALE) / area;
        const float alpha = sqrtf(meanI2 / meanJ2);
        const float beta = meanI - alpha * meanJ;
        float ib1 = 0, ib2 = 0, iA11 = 0, iA12 = 0, iA22 = 0;
        for (int k = 0; k < area; k++) {
          const int diff = (int)(Jw[k] * alpha - ref.gray[k] - beta);
          const float dx = (float)(ref.grad[2 * k] + dJw[2 * k] * alpha);
          const float dy = (float)(ref.grad[2 * k + 1] + dJw[2 * k + 1] * alpha);
          ib1 += (float)(diff * dx);
          ib2 += (float)(diff * dy);
          iA11 += (float)(dx * dx);
          iA22 += (float)(dy * dy);
          iA12 += (float)(dx * dy);
        }
        const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
        const float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
        if (minEig < T.min_eig || D < 1.1920928955078125e-7f) {
          if (level == 0) status_io[i] = NRSLAM_BAD_FEATURE;
          continue;  // :418-426 re-evaluates the same window until the iterations run out
        }
        D = 1.f / D;
        const float dlx = (float)((A12 * b2 - A22 * b1) * D), dly = (float)((A12 * b1 - A11 * b2) * D);
        nx += dlx;
        ny += dly;
        pts_io[2 * i] = nx + half;
        pts_io[2 * i + 1] = ny + half;
        if (pts_io[2 * i] < borderGap + 1 || pts_io[2 * i] >= L.w - 1 - borderGap || pts_io[2 * i + 1] < borderGap + 1 ||
            pts_io[2 * i + 1] >= L.h - 1 - borderGap) {
          if (level == 0) status_io[i] = NRSLAM_OUT_IMAGE_BOUNDARIES;
          break;
        }
        const double ex = (double)(pts_io[2 * i] - startx), ey = (double)(pts_io[2 * i + 1] - starty);
        if (sqrt(ex * ex + ey * ey) > 10) {
          pts_io[2 * i] = startx;
          pts_io[2 * i + 1] = starty;
          if (level == 0) status_io[i] = NRSLAM_BAD;
          break;
        }
        if ((double)dlx * dlx + (double)dly * dly <= T.eps) break;
        if (j > 0 && fabsf(dlx + pdx) < 0.01 && fabsf(dly + pdy) < 0.01) {
          pts_io[2 * i] -= dlx * 0.5f;
          pts_io[2 * i + 1] -= dly * 0.5f;
          break;
        }
        pdx = dlx;
        pdy = dly;
      }
    }
  }
  // SSIM gate on level 0 — :469-592
  int tracked = 0;
  const Level& L0 = pyr[0];
  const int st0 = L0.stride();
  const float C1 = (float)((0.01 * 255) * (0.01 * 255)), C2 = (float)((0.03 * 255) * (0.03 * 255));
  const float N_inv = 1.f / (float)area, N_inv_1 = 1.f / (float)(area - 1);
  std::vector<float> xr(area), yc(area);
  for (int i = 0; i < n; i++) {
    if (!is_usable(status_io[i])) continue;
    if (isnan(pts_io[2 * i]) || isnan(pts_io[2 * i + 1])) {
      status_io[i] = NRSLAM_OUT_IMAGE_BOUNDARIES;
      continue;
    }
    const float nx = pts_io[2 * i] - half, ny = pts_io[2 * i + 1] - half;
    const int inx = cv_floor(nx), iny = cv_floor(ny);
    const float aJ = nx - inx, bJ = ny - iny;
    const int jw00 = cv_round((1.f - aJ) * (1.f - bJ) * (1 << 14));
    const int jw01 = cv_round(aJ * (1.f - bJ) * (1 << 14));
    const int jw10 = cv_round((1.f - aJ) * bJ * (1 << 14));
    const int jw11 = (1 << 14) - jw00 - jw01 - jw10;
    if (inx < -borderGap || inx >= L0.w - borderGap * 2 || iny < -borderGap || iny >= L0.h - borderGap * 2) {
      status_io[i] = NRSLAM_OUT_IMAGE_BOUNDARIES;
      continue;
    }
    const Patch& ref = T.patches[0][i];
    float mu_x = 0.f, mu_y = 0.f;
    for (int y = 0; y < win; y++) {
      const uint8_t* src = L0.I(y + iny) + inx;
      for (int x = 0; x < win; x++) {
        const int jval = ORC_DESCALE(src[x] * jw00 + src[x + 1] * jw01 + src[x + st0] * jw10 + src[x + st0 + 1] * jw11, 14 - 5);
        // corrected /= 32 on CV_16S (convertTo with scale 1/32: round half to even), then saturate to u8 (:546-551)
        int c = cv_round_d((double)(int16_t)jval * (1. / 32));
        c = std::min(255, std::max(0, c));
        // refWin = Iref / 32 stays 16-bit (:553)
        const int r = ref.valid ? cv_round_d((double)ref.gray[(size_t)y * win + x] * (1. / 32)) : 0;
        xr[(size_t)y * win + x] = (float)r;
        yc[(size_t)y * win + x] = (float)c;
        mu_x += (float)r;
        mu_y += (float)c;
      }
    }
    mu_x *= N_inv;
    mu_y *= N_inv;
    double sxx = 0, syy = 0, sxy = 0;  // Mat::dot accumulates in double (:577-579)
    for (int k = 0; k < area; k++) {
      const float xn = xr[k] - mu_x, yn = yc[k] - mu_y;
      sxx += (double)xn * xn;
      syy += (double)yn * yn;
      sxy += (double)xn * yn;
    }
    const float sigma_x = sqrtf((float)(sxx * N_inv_1)), sigma_y = sqrtf((float)(syy * N_inv_1));
    const float sigma_xy = (float)(sxy * N_inv_1);
    const float ssim = ((2.f * mu_x * mu_y + C1) * (2.f * sigma_xy + C2)) /
                       ((mu_x * mu_x + mu_y * mu_y + C1) * (sigma_x * sigma_x + sigma_y * sigma_y + C2));
    if (ssim < min_ssim)
      status_io[i] = NRSLAM_BAD_FEATURE;
    else
      tracked++;
  }
  if (n_tracked_out) *n_tracked_out = tracked;
  return 0;
}

// GetPhotometricInformationOfPoint — :598-609
int orc_klt_get_patch(void* p, int idx, int16_t* gray, int16_t* grad, float* mean, float* mean2, uint8_t* valid) {
  Tracker& T = *static_cast<Tracker*>(p);
  if (idx < 0 || idx >= (int)(T.prev.size() / 2)) return NRSLAM_B200_ERR_ARG;
  const size_t area = (size_t)T.win * T.win;
  for (int level = 0; level <= T.max_level; level++) {
    const Patch& pt = T.patches[level][idx];
    valid[level] = pt.valid;
    mean[level] = pt.mean;
    mean2[level] = pt.mean2;
    if (pt.valid) {
      memcpy(gray + level * area, pt.gray.data(), area * 2);
      memcpy(grad + level * area * 2, pt.grad.data(), area * 4);
    } else {
      memset(gray + level * area, 0, area * 2);
      memset(grad + level * area * 2, 0, area * 4);
    }
  }
  return 0;
}

// InsertPhotometricInformation — :611-620
int orc_klt_insert_patch(void* p, float x, float y, const int16_t* gray, const int16_t* grad, const float* mean,
                         const float* mean2, const uint8_t* valid) {
  Tracker& T = *static_cast<Tracker*>(p);
  const size_t area = (size_t)T.win * T.win;
  T.prev.push_back(x);
  T.prev.push_back(y);
  for (int level = 0; level <= T.max_level; level++) {
    Patch pt;
    pt.valid = valid[level] != 0;
    pt.mean = mean[level];
    pt.mean2 = mean2[level];
    if (pt.valid) {
      pt.gray.assign(gray + level * area, gray + (level + 1) * area);
      pt.grad.assign(grad + level * area * 2, grad + (level + 1) * area * 2);
    }
    T.patches[level].push_back(std::move(pt));
  }
  return 0;
}

int orc_klt_clear(void* p) {  // :622-631
  Tracker& T = *static_cast<Tracker*>(p);
  for (auto& l : T.patches) l.clear();
  T.prev.clear();
  return 0;
}

}  // extern "C"
