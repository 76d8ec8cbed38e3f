// nrs_klt.cu — LucasKanadeTracker on sm_100a (C ABI: include/nrslam_b200.h, nrslam_b200_klt_*).
//
// What it replaces (reference paths relative to /root/reference):
//   LucasKanadeTracker::SetReferenceImage                modules/matching/lucas_kanade_tracker.cc:47-168
//   LucasKanadeTracker::Track + SSIM gate                modules/matching/lucas_kanade_tracker.cc:170-596
//   Get/InsertPhotometricInformation, clear              modules/matching/lucas_kanade_tracker.cc:598-631
//   cv::buildOpticalFlowPyramid (OpenCV, not vendored): pyrDown 5x5 + Scharr derivative + winSize borders
//
// Design. Integer pixel work, a few hundred KB per frame: bound by latency and L2, not HBM.
//   pyramid      one kernel per level over the BORDERED level (border pixels recompute the value of their
//                REFLECT_101 source, so no second pass), one Scharr kernel per level; int16x2 derivative.
//   patches      one warp per (point, level): 441 fixed-point bilinear samples, integer sums.
//   track        one warp per point runs ALL levels and iterations: the 21x21 reference patch and the current window
//                live in registers (14 pixels per lane), the two reductions per iteration are warp shuffles, the
//                image is read through the read-only path (the whole pyramid is L2 resident). No block barriers.
//   ssim         fused tail of the track kernel.
// Integer results (pyramid, reference patches) are bit-exact with the oracle; the float accumulations of the
// reference run row-major over the window (lucas_kanade_tracker.cc:300-401) while a warp sums per lane then by
// shuffle tree, so positions agree to ~1e-3 px and a status can flip on a threshold tie.
// There is no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "nrs_host.h"

namespace {

constexpr int kWin = 21;             // the reference hard-codes 21x21 (modules/SLAM/system.cc:78-83)
constexpr int kArea = kWin * kWin;   // 441
constexpr int kPerLane = (kArea + 31) / 32;  // 14
constexpr int kMaxLevels = 8;

#define KLT_DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

__host__ __device__ inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

struct LevelDev {
  int w, h, stride;  // stride of the bordered buffers (w + 2 * kWin)
  unsigned char* img;
  short2* deriv;
};

struct PyrDev {
  int n_levels;
  LevelDev lv[kMaxLevels];
};

// level 0: pitched source -> bordered level (REFLECT_101)
__global__ void klt_level0_kernel(const unsigned char* __restrict__ src, int pitch, LevelDev L) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
  if (X >= L.stride || Y >= L.h + 2 * kWin) return;
  const int x = reflect101(X - kWin, L.w), y = reflect101(Y - kWin, L.h);
  L.img[(size_t)Y * L.stride + X] = src[(size_t)y * pitch + x];
}

// cv::pyrDown ([1 4 6 4 1] x [1 4 6 4 1], (sum + 128) >> 8, REFLECT_101 inside the source level) evaluated for every
// pixel of the bordered destination level
__global__ void klt_pyrdown_kernel(LevelDev S, LevelDev D) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
  if (X >= D.stride || Y >= D.h + 2 * kWin) return;
  const int x = reflect101(X - kWin, D.w), y = reflect101(Y - kWin, D.h);
  const int wgt[5] = {1, 4, 6, 4, 1};
  int sum = 0;
#pragma unroll
  for (int dy = 0; dy < 5; dy++) {
    const int sy = reflect101(2 * y - 2 + dy, S.h);
    const unsigned char* row = S.img + (size_t)(sy + kWin) * S.stride + kWin;
    int hs = 0;
#pragma unroll
    for (int dx = 0; dx < 5; dx++) hs += wgt[dx] * row[reflect101(2 * x - 2 + dx, S.w)];
    sum += wgt[dy] * hs;
  }
  D.img[(size_t)Y * D.stride + X] = (unsigned char)((sum + 128) >> 8);
}

// calcSharrDeriv on the level interior (the bordered image already carries the in-level REFLECT_101 neighbours);
// the derivative border stays 0 (BORDER_CONSTANT)
__global__ void klt_scharr_kernel(LevelDev L) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= L.w || y >= L.h) return;
  const unsigned char* c = L.img + (size_t)(y + kWin) * L.stride + (x + kWin);
  const int s = L.stride;
  const int t0m = (c[-s - 1] + c[s - 1]) * 3 + c[-1] * 10, t0p = (c[-s + 1] + c[s + 1]) * 3 + c[1] * 10;
  const int t1m = c[s - 1] - c[-s - 1], t1c = c[s] - c[-s], t1p = c[s + 1] - c[-s + 1];
  L.deriv[(size_t)(y + kWin) * L.stride + (x + kWin)] = make_short2((short)(t0p - t0m), (short)((t1p + t1m) * 3 + t1c * 10));
}

struct PatchStore {
  short* gray;      // [cap][levels][441]
  short2* grad;     // [cap][levels][441]
  float* mean;      // [cap][levels]
  float* mean2;     // [cap][levels]
  unsigned char* valid;  // [cap][levels]
};

__device__ __forceinline__ int cv_floor_dev(float v) { return __float2int_rd(v); }
__device__ __forceinline__ int cv_round_dev(float v) { return __float2int_rn(v); }  // round half to even

// SetReferenceImage: one warp per (point, level)
__global__ void klt_ref_patches_kernel(PyrDev pyr, int n, int n_levels, const float* __restrict__ pts, PatchStore ps,
                                       const unsigned char* __restrict__ mask, int mask_pitch, int img_w, int img_h) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n * n_levels) return;
  const int i = gw / n_levels, level = gw % n_levels;
  const LevelDev L = pyr.lv[level];
  const float half = (kWin - 1) * 0.5f;
  const int borderGap = kWin / 2;  // :58
  const float scale_div = (float)(1 << level);
  const float px = __fdiv_rn(pts[2 * i], scale_div) - half, py = __fdiv_rn(pts[2 * i + 1], scale_div) - half;
  const int ix = cv_floor_dev(px), iy = cv_floor_dev(py);
  const size_t slot = (size_t)i * n_levels + level;
  bool ok = !(ix < -borderGap || ix >= L.w - borderGap || iy < -borderGap || iy >= L.h - borderGap);
  long long sum = 0, sum2 = 0;
  if (ok) {
    const float a = px - ix, b = py - iy;
    const int iw00 = cv_round_dev(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), 16384.f));
    const int iw01 = cv_round_dev(__fmul_rn(__fmul_rn(a, 1.f - b), 16384.f));
    const int iw10 = cv_round_dev(__fmul_rn(__fmul_rn(1.f - a, b), 16384.f));
    const int iw11 = 16384 - iw00 - iw01 - iw10;
    const int s = L.stride;
    const int sc = 1 << level;
    bool masked = false;
    for (int k = lane; k < kArea; k += 32) {
      const int y = k / kWin, x = k - y * kWin;
      if (mask) {
        const int mx = (ix + x) * sc, my = (iy + y) * sc;
        const bool in = mx >= 0 && mx < img_w && my >= 0 && my < img_h;
        if (!in || mask[(size_t)my * mask_pitch + mx] == 0) masked = true;
      }
      const unsigned char* src = L.img + (size_t)(y + iy + kWin) * s + (x + ix + kWin);
      const short2* d = L.deriv + (size_t)(y + iy + kWin) * s + (x + ix + kWin);
      const short2 d00 = d[0], d01 = d[1], d10 = d[s], d11 = d[s + 1];
      const int ival = KLT_DESCALE(src[0] * iw00 + src[1] * iw01 + src[s] * iw10 + src[s + 1] * iw11, 14 - 5);
      const int ixv = KLT_DESCALE(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
      const int iyv = KLT_DESCALE(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
      ps.gray[slot * kArea + k] = (short)ival;
      ps.grad[slot * kArea + k] = make_short2((short)ixv, (short)iyv);
      sum += ival;
      sum2 += (long long)ival * ival;
    }
    ok = !__any_sync(0xffffffffu, masked);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
  }
  if (lane == 0) {
    const float FLT_SCALE = 1.f / (1 << 20);
    ps.valid[slot] = ok ? 1 : 0;
    // exact integer sums rounded once (the reference accumulates the 441 terms in fp32, :149-150,160-161)
    ps.mean[slot] = ok ? __fdiv_rn(__fmul_rn((float)sum, FLT_SCALE), (float)kArea) : -1.f;
    ps.mean2[slot] = ok ? __fdiv_rn(__fmul_rn((float)sum2, FLT_SCALE), (float)kArea) : -1.f;
  }
}

__device__ __forceinline__ bool usable(unsigned char s) {
  return s == NRSLAM_TRACKED_WITH_3D || s == NRSLAM_TRACKED || s == NRSLAM_JUST_TRIANGULATED;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Track: one warp per point, all levels, all iterations, SSIM gate.
__global__ void __launch_bounds__(128) klt_track_kernel(PyrDev pyr, int n, int n_levels, int max_iters, float epsilon,
                                                        float min_eig, const float* __restrict__ prev, PatchStore ps,
                                                        float* pts_io, unsigned char* status_io, int use_initial_flow,
                                                        float min_ssim) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  unsigned char status = status_io[i];
  if (!usable(status)) return;
  const float half = (kWin - 1) * 0.5f;
  const int borderGap = kWin / 2 + 1;  // :186
  const float FLT_SCALE = 1.f / (1 << 20);
  float outx = pts_io[2 * i], outy = pts_io[2 * i + 1];
  const float p0x = prev[2 * i], p0y = prev[2 * i + 1];
  for (int level = n_levels - 1; level >= 0; level--) {
    if (!usable(status)) break;
    const LevelDev L = pyr.lv[level];
    const int s = L.stride;
    const float inv = (float)(1. / (1 << level));
    float prevx = __fmul_rn(p0x, inv), prevy = __fmul_rn(p0y, inv);
    float nx, ny;
    if (level == n_levels - 1) {
      if (use_initial_flow) {
        nx = __fmul_rn(outx, inv);
        ny = __fmul_rn(outy, inv);
      } else {
        nx = prevx;
        ny = prevy;
      }
    } else {
      nx = outx * 2.f;
      ny = outy * 2.f;
    }
    outx = nx;
    outy = ny;
    prevx -= half;
    prevy -= half;
    const int ipx = cv_floor_dev(prevx), ipy = cv_floor_dev(prevy);
    const size_t slot = (size_t)i * n_levels + level;
    if (ipx < -borderGap || ipx >= L.w - borderGap || ipy < -borderGap || ipy >= L.h - borderGap || !ps.valid[slot]) {
      if (level == 0) status = NRSLAM_OUT_IMAGE_BOUNDARIES;
      continue;
    }
    const float meanI = ps.mean[slot], meanI2 = ps.mean2[slot];
    // reference patch of this level in registers
    int Iv[kPerLane], Ixv[kPerLane], Iyv[kPerLane];
#pragma unroll
    for (int t = 0; t < kPerLane; t++) {
      const int k = lane + 32 * t;
      if (k < kArea) {
        Iv[t] = ps.gray[slot * kArea + k];
        const short2 g = ps.grad[slot * kArea + k];
        Ixv[t] = g.x;
        Iyv[t] = g.y;
      } else {
        Iv[t] = Ixv[t] = Iyv[t] = 0;
      }
    }
    const float startx = nx, starty = ny;
    float pdx = 0.f, pdy = 0.f;
    nx -= half;
    ny -= half;
    for (int j = 0; j < max_iters; j++) {
      const int inx = cv_floor_dev(nx), iny = cv_floor_dev(ny);
      if (inx < -borderGap || inx >= L.w - borderGap || iny < -borderGap || iny >= L.h - borderGap) {
        if (level == 0) status = NRSLAM_OUT_IMAGE_BOUNDARIES;
        break;
      }
      const float aJ = nx - inx, bJ = ny - iny;
      const int jw00 = cv_round_dev(__fmul_rn(__fmul_rn(1.f - aJ, 1.f - bJ), 16384.f));
      const int jw01 = cv_round_dev(__fmul_rn(__fmul_rn(aJ, 1.f - bJ), 16384.f));
      const int jw10 = cv_round_dev(__fmul_rn(__fmul_rn(1.f - aJ, bJ), 16384.f));
      const int jw11 = 16384 - jw00 - jw01 - jw10;
      int Jv[kPerLane], Jxv[kPerLane], Jyv[kPerLane];
      long long sumJ = 0, sumJ2 = 0;
      const unsigned char* base_i = L.img + (size_t)(iny + kWin) * s + (inx + kWin);
      const short2* base_d = L.deriv + (size_t)(iny + kWin) * s + (inx + kWin);
#pragma unroll
      for (int t = 0; t < kPerLane; t++) {
        const int k = lane + 32 * t;
        if (k < kArea) {
          const int y = k / kWin, x = k - y * kWin;
          const unsigned char* src = base_i + y * s + x;
          const short2* d = base_d + y * s + x;
          const short2 d00 = __ldg(d), d01 = __ldg(d + 1), d10 = __ldg(d + s), d11 = __ldg(d + s + 1);
          const int jval = KLT_DESCALE(__ldg(src) * jw00 + __ldg(src + 1) * jw01 + __ldg(src + s) * jw10 +
                                           __ldg(src + s + 1) * jw11, 14 - 5);
          Jv[t] = (short)jval;
          Jxv[t] = (short)KLT_DESCALE(d00.x * jw00 + d01.x * jw01 + d10.x * jw10 + d11.x * jw11, 14);
          Jyv[t] = (short)KLT_DESCALE(d00.y * jw00 + d01.y * jw01 + d10.y * jw10 + d11.y * jw11, 14);
          sumJ += jval;
          sumJ2 += (long long)jval * jval;
        } else {
          Jv[t] = Jxv[t] = Jyv[t] = 0;
        }
      }
      sumJ = warp_sum(sumJ);
      sumJ2 = warp_sum(sumJ2);
      const float meanJ = __fdiv_rn(__fmul_rn((float)sumJ, FLT_SCALE), (float)kArea);
      const float meanJ2 = __fdiv_rn(__fmul_rn((float)sumJ2, FLT_SCALE), (float)kArea);
      const float alpha = __fsqrt_rn(__fdiv_rn(meanI2, meanJ2));
      const float beta = __fsub_rn(meanI, __fmul_rn(alpha, meanJ));
      float ib1 = 0, ib2 = 0, iA11 = 0, iA12 = 0, iA22 = 0;
#pragma unroll
      for (int t = 0; t < kPerLane; t++) {
        const int k = lane + 32 * t;
        if (k < kArea) {
          // int diff = Jptr[x] * alpha - Iptr[x] - beta  (truncation toward zero, :392)
          const int diff = (int)__fsub_rn(__fsub_rn(__fmul_rn((float)Jv[t], alpha), (float)Iv[t]), beta);
          const float dx = __fadd_rn((float)Ixv[t], __fmul_rn((float)Jxv[t], alpha));
          const float dy = __fadd_rn((float)Iyv[t], __fmul_rn((float)Jyv[t], alpha));
          ib1 = __fadd_rn(ib1, __fmul_rn((float)diff, dx));
          ib2 = __fadd_rn(ib2, __fmul_rn((float)diff, dy));
          iA11 = __fadd_rn(iA11, __fmul_rn(dx, dx));
          iA22 = __fadd_rn(iA22, __fmul_rn(dy, dy));
          iA12 = __fadd_rn(iA12, __fmul_rn(dx, dy));
        }
      }
      ib1 = warp_sum(ib1);
      ib2 = warp_sum(ib2);
      iA11 = warp_sum(iA11);
      iA12 = warp_sum(iA12);
      iA22 = warp_sum(iA22);
      const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
      const float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
      float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
      const float dA = __fsub_rn(A11, A22);
      const float minEig = __fdiv_rn(
          __fsub_rn(__fadd_rn(A22, A11),
                    __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
          (float)(2 * kWin * kWin));
      if (minEig < min_eig || D < 1.1920928955078125e-7f) {
        if (level == 0) status = NRSLAM_BAD_FEATURE;
        break;  // the reference `continue`s and re-evaluates the same window until the iterations run out (E15)
      }
      D = __fdiv_rn(1.f, D);
      const float dlx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dly = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      nx += dlx;
      ny += dly;
      outx = nx + half;
      outy = ny + half;
      if (outx < borderGap + 1 || outx >= L.w - 1 - borderGap || outy < borderGap + 1 || outy >= L.h - 1 - borderGap) {
        if (level == 0) status = NRSLAM_OUT_IMAGE_BOUNDARIES;
        break;
      }
      const double ex = (double)(outx - startx), ey = (double)(outy - starty);
      if (sqrt(ex * ex + ey * ey) > 10) {
        outx = startx;
        outy = starty;
        if (level == 0) status = NRSLAM_BAD;
        break;
      }
      if ((double)dlx * dlx + (double)dly * dly <= (double)epsilon) break;
      if (j > 0 && fabsf(dlx + pdx) < 0.01 && fabsf(dly + pdy) < 0.01) {
        outx -= dlx * 0.5f;
        outy -= dly * 0.5f;
        break;
      }
      pdx = dlx;
      pdy = dly;
    }
  }
  // ---- SSIM gate on level 0 (:469-592)
  if (usable(status)) {
    if (isnan(outx) || isnan(outy)) {
      status = NRSLAM_OUT_IMAGE_BOUNDARIES;
    } else {
      const LevelDev L = pyr.lv[0];
      const int s = L.stride;
      const float nx = outx - half, ny = outy - half;
      const int inx = cv_floor_dev(nx), iny = cv_floor_dev(ny);
      if (inx < -borderGap || inx >= L.w - borderGap * 2 || iny < -borderGap || iny >= L.h - borderGap * 2) {
        status = NRSLAM_OUT_IMAGE_BOUNDARIES;
      } else {
        const float aJ = nx - inx, bJ = ny - iny;
        const int jw00 = cv_round_dev(__fmul_rn(__fmul_rn(1.f - aJ, 1.f - bJ), 16384.f));
        const int jw01 = cv_round_dev(__fmul_rn(__fmul_rn(aJ, 1.f - bJ), 16384.f));
        const int jw10 = cv_round_dev(__fmul_rn(__fmul_rn(1.f - aJ, bJ), 16384.f));
        const int jw11 = 16384 - jw00 - jw01 - jw10;
        const size_t slot = (size_t)i * n_levels;
        const bool rvalid = ps.valid[slot] != 0;
        int rv[kPerLane], cv[kPerLane];
        int sr = 0, sc = 0;
#pragma unroll
        for (int t = 0; t < kPerLane; t++) {
          const int k = lane + 32 * t;
          rv[t] = cv[t] = 0;
          if (k < kArea) {
            const int y = k / kWin, x = k - y * kWin;
            const unsigned char* src = L.img + (size_t)(y + iny + kWin) * s + (x + inx + kWin);
            const int jval = (short)KLT_DESCALE(__ldg(src) * jw00 + __ldg(src + 1) * jw01 + __ldg(src + s) * jw10 +
                                                    __ldg(src + s + 1) * jw11, 14 - 5);
            // Mat /= 32 on CV_16S: convertTo with scale 1/32, round half to even; then saturate to u8 (:546-551)
            int c = __double2int_rn((double)jval * (1. / 32));
            c = min(255, max(0, c));
            const int r = rvalid ? __double2int_rn((double)ps.gray[slot * kArea + k] * (1. / 32)) : 0;
            rv[t] = r;
            cv[t] = c;
            sr += r;
            sc += c;
          }
        }
        sr = warp_sum(sr);
        sc = warp_sum(sc);
        const float N_inv = 1.f / (float)kArea, N_inv_1 = 1.f / (float)(kArea - 1);
        const float mu_x = __fmul_rn((float)sr, N_inv), mu_y = __fmul_rn((float)sc, N_inv);
        double sxx = 0, syy = 0, sxy = 0;
#pragma unroll
        for (int t = 0; t < kPerLane; t++) {
          const int k = lane + 32 * t;
          if (k < kArea) {
            const float xn = __fsub_rn((float)rv[t], mu_x), yn = __fsub_rn((float)cv[t], mu_y);
            sxx += (double)xn * xn;
            syy += (double)yn * yn;
            sxy += (double)xn * yn;
          }
        }
        sxx = warp_sum(sxx);
        syy = warp_sum(syy);
        sxy = warp_sum(sxy);
        const float C1 = (float)((0.01 * 255) * (0.01 * 255)), C2 = (float)((0.03 * 255) * (0.03 * 255));
        const float sigma_x = __fsqrt_rn((float)(sxx * N_inv_1)), sigma_y = __fsqrt_rn((float)(syy * N_inv_1));
        const float sigma_xy = (float)(sxy * N_inv_1);
        const float num = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.f, mu_x), mu_y), C1),
                                    __fadd_rn(__fmul_rn(2.f, sigma_xy), C2));
        const float den = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(mu_x, mu_x), __fmul_rn(mu_y, mu_y)), C1),
                                    __fadd_rn(__fadd_rn(__fmul_rn(sigma_x, sigma_x), __fmul_rn(sigma_y, sigma_y)), C2));
        if (__fdiv_rn(num, den) < min_ssim) status = NRSLAM_BAD_FEATURE;
      }
    }
  }
  if (lane == 0) {
    pts_io[2 * i] = outx;
    pts_io[2 * i + 1] = outy;
    status_io[i] = status;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct nrslam_b200_klt {
  nrslam_b200_ctx* ctx = nullptr;
  int max_level = 4, max_iters = 10;
  float eps = 1e-4f, min_eig = 1e-4f;
  int n_levels() const { return max_level + 1; }
  // image-size dependent buffers
  int w = 0, h = 0;
  unsigned char* d_src = nullptr;  // pitched upload
  unsigned char* d_mask = nullptr;
  size_t src_cap = 0;
  unsigned char* h_src = nullptr;  // pinned staging
  PyrDev ref{}, cur{};
  // points
  int n = 0, cap = 0;
  PatchStore ps{};
  float* d_prev = nullptr;
  float* d_pts = nullptr;
  unsigned char* d_status = nullptr;
  float* h_pts = nullptr;          // pinned
  unsigned char* h_status = nullptr;
  // last track call (retrack hook)
  bool have_track = false;
  int last_use_flow = 0;
  float last_min_ssim = 0.7f;
  std::vector<float> last_pts;
  std::vector<unsigned char> last_status;
};

namespace {

int kfail(nrslam_b200_klt* k, int code, const std::string& msg) {
  if (k && k->ctx) k->ctx->err = msg;
  return code;
}
#define KLT_CUDA(k, call)                                                                             \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) return kfail(k, NRSLAM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

void free_pyr(PyrDev& p) {
  for (int l = 0; l < kMaxLevels; l++) {
    if (p.lv[l].img) cudaFree(p.lv[l].img);
    if (p.lv[l].deriv) cudaFree(p.lv[l].deriv);
    p.lv[l].img = nullptr;
    p.lv[l].deriv = nullptr;
  }
  p.n_levels = 0;
}

int alloc_pyr(nrslam_b200_klt* k, PyrDev& p, int w, int h) {
  free_pyr(p);
  int lw = w, lh = h;
  p.n_levels = k->n_levels();
  for (int l = 0; l < p.n_levels; l++) {
    if (l > 0 && (lw <= kWin || lh <= kWin))
      return kfail(k, NRSLAM_B200_ERR_ARG, "image too small for the pyramid (a level would be smaller than the window)");
    LevelDev& L = p.lv[l];
    L.w = lw;
    L.h = lh;
    L.stride = lw + 2 * kWin;
    const size_t px = (size_t)(lh + 2 * kWin) * L.stride;
    KLT_CUDA(k, cudaMalloc(&L.img, px));
    KLT_CUDA(k, cudaMalloc(&L.deriv, px * sizeof(short2)));
    KLT_CUDA(k, cudaMemsetAsync(L.deriv, 0, px * sizeof(short2), k->ctx->stream));  // constant-0 derivative border
    lw = (lw + 1) / 2;
    lh = (lh + 1) / 2;
  }
  return 0;
}

int ensure_image(nrslam_b200_klt* k, int w, int h) {
  if (w == k->w && h == k->h) return 0;
  if (w <= kWin || h <= kWin) return kfail(k, NRSLAM_B200_ERR_ARG, "image smaller than the window");
  int rc = alloc_pyr(k, k->ref, w, h);
  if (rc) return rc;
  rc = alloc_pyr(k, k->cur, w, h);
  if (rc) return rc;
  if (k->d_src) cudaFree(k->d_src);
  if (k->d_mask) cudaFree(k->d_mask);
  if (k->h_src) cudaFreeHost(k->h_src);
  KLT_CUDA(k, cudaMalloc(&k->d_src, (size_t)w * h));
  KLT_CUDA(k, cudaMalloc(&k->d_mask, (size_t)w * h));
  KLT_CUDA(k, cudaMallocHost(&k->h_src, (size_t)w * h));
  k->w = w;
  k->h = h;
  return 0;
}

int ensure_points(nrslam_b200_klt* k, int n) {
  if (n <= k->cap) return 0;
  const int cap = std::max(n + n / 2, 256);
  const int nl = k->n_levels();
  PatchStore ns{};
  float *prev = nullptr, *pts = nullptr, *hpts = nullptr;
  unsigned char *st = nullptr, *hst = nullptr;
  KLT_CUDA(k, cudaMalloc(&ns.gray, (size_t)cap * nl * kArea * sizeof(short)));
  KLT_CUDA(k, cudaMalloc(&ns.grad, (size_t)cap * nl * kArea * sizeof(short2)));
  KLT_CUDA(k, cudaMalloc(&ns.mean, (size_t)cap * nl * sizeof(float)));
  KLT_CUDA(k, cudaMalloc(&ns.mean2, (size_t)cap * nl * sizeof(float)));
  KLT_CUDA(k, cudaMalloc(&ns.valid, (size_t)cap * nl));
  KLT_CUDA(k, cudaMalloc(&prev, (size_t)cap * 2 * sizeof(float)));
  KLT_CUDA(k, cudaMalloc(&pts, (size_t)cap * 2 * sizeof(float)));
  KLT_CUDA(k, cudaMalloc(&st, (size_t)cap));
  KLT_CUDA(k, cudaMallocHost(&hpts, (size_t)cap * 2 * sizeof(float)));
  KLT_CUDA(k, cudaMallocHost(&hst, (size_t)cap));
  if (k->n > 0) {  // keep the existing points (InsertPhotometricInformation appends)
    cudaStream_t s = k->ctx->stream;
    const size_t m = (size_t)k->n * nl;
    KLT_CUDA(k, cudaMemcpyAsync(ns.gray, k->ps.gray, m * kArea * sizeof(short), cudaMemcpyDeviceToDevice, s));
    KLT_CUDA(k, cudaMemcpyAsync(ns.grad, k->ps.grad, m * kArea * sizeof(short2), cudaMemcpyDeviceToDevice, s));
    KLT_CUDA(k, cudaMemcpyAsync(ns.mean, k->ps.mean, m * sizeof(float), cudaMemcpyDeviceToDevice, s));
    KLT_CUDA(k, cudaMemcpyAsync(ns.mean2, k->ps.mean2, m * sizeof(float), cudaMemcpyDeviceToDevice, s));
    KLT_CUDA(k, cudaMemcpyAsync(ns.valid, k->ps.valid, m, cudaMemcpyDeviceToDevice, s));
    KLT_CUDA(k, cudaMemcpyAsync(prev, k->d_prev, (size_t)k->n * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    KLT_CUDA(k, cudaStreamSynchronize(s));
  }
  if (k->ps.gray) cudaFree(k->ps.gray);
  if (k->ps.grad) cudaFree(k->ps.grad);
  if (k->ps.mean) cudaFree(k->ps.mean);
  if (k->ps.mean2) cudaFree(k->ps.mean2);
  if (k->ps.valid) cudaFree(k->ps.valid);
  if (k->d_prev) cudaFree(k->d_prev);
  if (k->d_pts) cudaFree(k->d_pts);
  if (k->d_status) cudaFree(k->d_status);
  if (k->h_pts) cudaFreeHost(k->h_pts);
  if (k->h_status) cudaFreeHost(k->h_status);
  k->ps = ns;
  k->d_prev = prev;
  k->d_pts = pts;
  k->d_status = st;
  k->h_pts = hpts;
  k->h_status = hst;
  k->cap = cap;
  return 0;
}

// upload a pitched host image and build the pyramid
int build_pyramid(nrslam_b200_klt* k, PyrDev& p, const uint8_t* image, int pitch) {
  cudaStream_t s = k->ctx->stream;
  for (int y = 0; y < k->h; y++) memcpy(k->h_src + (size_t)y * k->w, image + (size_t)y * pitch, k->w);
  KLT_CUDA(k, cudaMemcpyAsync(k->d_src, k->h_src, (size_t)k->w * k->h, cudaMemcpyHostToDevice, s));
  const dim3 blk(32, 8);
  {
    const LevelDev& L = p.lv[0];
    const dim3 grd((L.stride + 31) / 32, (L.h + 2 * kWin + 7) / 8);
    klt_level0_kernel<<<grd, blk, 0, s>>>(k->d_src, k->w, L);
  }
  for (int l = 0; l < p.n_levels; l++) {
    const LevelDev& L = p.lv[l];
    if (l > 0) {
      const dim3 grd((L.stride + 31) / 32, (L.h + 2 * kWin + 7) / 8);
      klt_pyrdown_kernel<<<grd, blk, 0, s>>>(p.lv[l - 1], L);
    }
    const dim3 grd2((L.w + 31) / 32, (L.h + 7) / 8);
    klt_scharr_kernel<<<grd2, blk, 0, s>>>(L);
  }
  KLT_CUDA(k, cudaGetLastError());
  return 0;
}

int launch_track(nrslam_b200_klt* k, int n, int use_flow, float min_ssim) {
  const int threads = 128, warps_per_block = threads / 32;
  const int blocks = (n + warps_per_block - 1) / warps_per_block;
  klt_track_kernel<<<blocks, threads, 0, k->ctx->stream>>>(k->cur, n, k->n_levels(), k->max_iters, k->eps, k->min_eig,
                                                           k->d_prev, k->ps, k->d_pts, k->d_status, use_flow, min_ssim);
  KLT_CUDA(k, cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

int nrslam_b200_klt_create(nrslam_b200_ctx* ctx, int32_t win_size, int32_t max_level, int32_t max_iters, float epsilon,
                           float min_eig_threshold, nrslam_b200_klt** out) {
  if (out) *out = nullptr;
  if (!ctx || !out) return NRSLAM_B200_ERR_ARG;
  if (win_size != kWin) {
    ctx->err = "klt: only the reference's 21x21 window is built (modules/SLAM/system.cc:78-83)";
    return NRSLAM_B200_ERR_ARG;
  }
  if (max_level < 0 || max_level >= kMaxLevels || max_iters < 1) {
    ctx->err = "klt: bad pyramid depth / iteration count";
    return NRSLAM_B200_ERR_ARG;
  }
  nrslam_b200_klt* k = new nrslam_b200_klt();
  k->ctx = ctx;
  k->max_level = max_level;
  k->max_iters = max_iters;
  k->eps = epsilon;
  k->min_eig = min_eig_threshold;
  *out = k;
  return 0;
}

void nrslam_b200_klt_destroy(nrslam_b200_klt* k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);
  free_pyr(k->ref);
  free_pyr(k->cur);
  if (k->d_src) cudaFree(k->d_src);
  if (k->d_mask) cudaFree(k->d_mask);
  if (k->h_src) cudaFreeHost(k->h_src);
  if (k->ps.gray) cudaFree(k->ps.gray);
  if (k->ps.grad) cudaFree(k->ps.grad);
  if (k->ps.mean) cudaFree(k->ps.mean);
  if (k->ps.mean2) cudaFree(k->ps.mean2);
  if (k->ps.valid) cudaFree(k->ps.valid);
  if (k->d_prev) cudaFree(k->d_prev);
  if (k->d_pts) cudaFree(k->d_pts);
  if (k->d_status) cudaFree(k->d_status);
  if (k->h_pts) cudaFreeHost(k->h_pts);
  if (k->h_status) cudaFreeHost(k->h_status);
  delete k;
}

int nrslam_b200_klt_set_reference(nrslam_b200_klt* k, const uint8_t* image, int32_t width, int32_t height, int32_t pitch,
                                  int32_t n_points, const float* pts_xy, const uint8_t* mask, int32_t mask_pitch) {
  if (!k || !image || n_points < 0 || (n_points > 0 && !pts_xy) || pitch < width)
    return kfail(k, NRSLAM_B200_ERR_ARG, "klt_set_reference: bad argument");
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  int rc = ensure_image(k, width, height);
  if (rc) return rc;
  k->n = 0;  // prevPts_ = refPts (:53)
  rc = ensure_points(k, n_points);
  if (rc) return rc;
  rc = build_pyramid(k, k->ref, image, pitch);
  if (rc) return rc;
  cudaStream_t s = k->ctx->stream;
  k->n = n_points;
  k->have_track = false;
  if (n_points == 0) {
    KLT_CUDA(k, cudaStreamSynchronize(s));
    return 0;
  }
  memcpy(k->h_pts, pts_xy, (size_t)n_points * 2 * sizeof(float));
  KLT_CUDA(k, cudaMemcpyAsync(k->d_prev, k->h_pts, (size_t)n_points * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  const unsigned char* dmask = nullptr;
  if (mask) {
    KLT_CUDA(k, cudaStreamSynchronize(s));  // the pinned staging buffer is still the source of the image upload
    for (int y = 0; y < height; y++) memcpy(k->h_src + (size_t)y * width, mask + (size_t)y * mask_pitch, width);
    KLT_CUDA(k, cudaMemcpyAsync(k->d_mask, k->h_src, (size_t)width * height, cudaMemcpyHostToDevice, s));
    dmask = k->d_mask;
  }
  const int total_warps = n_points * k->n_levels();
  const int threads = 128;
  const int blocks = (total_warps * 32 + threads - 1) / threads;
  klt_ref_patches_kernel<<<blocks, threads, 0, s>>>(k->ref, n_points, k->n_levels(), k->d_prev, k->ps, dmask, width, width,
                                                    height);
  KLT_CUDA(k, cudaGetLastError());
  KLT_CUDA(k, cudaStreamSynchronize(s));
  return 0;
}

int nrslam_b200_klt_track(nrslam_b200_klt* k, const uint8_t* image, int32_t width, int32_t height, int32_t pitch,
                          int32_t n_points, float* pts_io, uint8_t* status_io, int32_t use_initial_flow, float min_ssim,
                          const uint8_t* mask, int32_t mask_pitch, int32_t* n_tracked_out) {
  (void)mask;  // the reference's Track never reads the mask (`if(false && ...)`, :321; Jvalid is unused in the SSIM stage)
  (void)mask_pitch;
  if (!k || !image || !pts_io || !status_io || pitch < width)
    return kfail(k, NRSLAM_B200_ERR_ARG, "klt_track: bad argument");
  if (n_points != k->n) return kfail(k, NRSLAM_B200_ERR_ARG, "klt_track: point count differs from the reference set");
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  if (k->w == 0) {  // a tracker fed only through InsertPhotometricInformation (Tracking::PointReuse, tracking.cc:421-447)
    const int rc0 = ensure_image(k, width, height);
    if (rc0) return rc0;
  }
  if (width != k->w || height != k->h) return kfail(k, NRSLAM_B200_ERR_ARG, "klt_track: image size differs from the reference image");
  if (n_tracked_out) *n_tracked_out = 0;
  int rc = build_pyramid(k, k->cur, image, pitch);
  if (rc) return rc;
  cudaStream_t s = k->ctx->stream;
  if (n_points == 0) {
    KLT_CUDA(k, cudaStreamSynchronize(s));
    return 0;
  }
  memcpy(k->h_pts, pts_io, (size_t)n_points * 2 * sizeof(float));
  memcpy(k->h_status, status_io, n_points);
  k->last_pts.assign(pts_io, pts_io + 2 * (size_t)n_points);
  k->last_status.assign(status_io, status_io + n_points);
  k->last_use_flow = use_initial_flow;
  k->last_min_ssim = min_ssim;
  KLT_CUDA(k, cudaMemcpyAsync(k->d_pts, k->h_pts, (size_t)n_points * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->d_status, k->h_status, n_points, cudaMemcpyHostToDevice, s));
  rc = launch_track(k, n_points, use_initial_flow, min_ssim);
  if (rc) return rc;
  KLT_CUDA(k, cudaMemcpyAsync(k->h_pts, k->d_pts, (size_t)n_points * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->h_status, k->d_status, n_points, cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaStreamSynchronize(s));
  // unusable points keep their caller-side position (the kernel does not touch them)
  int tracked = 0;
  for (int i = 0; i < n_points; i++) {
    const uint8_t before = status_io[i];
    const bool was_usable = before == NRSLAM_TRACKED_WITH_3D || before == NRSLAM_TRACKED || before == NRSLAM_JUST_TRIANGULATED;
    if (!was_usable) continue;
    pts_io[2 * i] = k->h_pts[2 * i];
    pts_io[2 * i + 1] = k->h_pts[2 * i + 1];
    status_io[i] = k->h_status[i];
    const uint8_t a = status_io[i];
    if (a == NRSLAM_TRACKED_WITH_3D || a == NRSLAM_TRACKED || a == NRSLAM_JUST_TRIANGULATED) tracked++;
  }
  if (n_tracked_out) *n_tracked_out = tracked;
  k->have_track = true;
  return 0;
}

int nrslam_b200_klt_retrack(nrslam_b200_klt* k, float* gpu_ms_out) {
  if (!k || !k->have_track) return kfail(k, NRSLAM_B200_ERR_ARG, "klt_retrack: no previous Track call");
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  cudaStream_t s = k->ctx->stream;
  const int n = k->n;
  memcpy(k->h_pts, k->last_pts.data(), (size_t)n * 2 * sizeof(float));
  memcpy(k->h_status, k->last_status.data(), n);
  KLT_CUDA(k, cudaMemcpyAsync(k->d_pts, k->h_pts, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->d_status, k->h_status, n, cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaEventRecord(k->ctx->ev0, s));
  // pyramid of the resident current image + track
  const dim3 blk(32, 8);
  {
    const LevelDev& L = k->cur.lv[0];
    const dim3 grd((L.stride + 31) / 32, (L.h + 2 * kWin + 7) / 8);
    klt_level0_kernel<<<grd, blk, 0, s>>>(k->d_src, k->w, L);
  }
  for (int l = 0; l < k->cur.n_levels; l++) {
    const LevelDev& L = k->cur.lv[l];
    if (l > 0) {
      const dim3 grd((L.stride + 31) / 32, (L.h + 2 * kWin + 7) / 8);
      klt_pyrdown_kernel<<<grd, blk, 0, s>>>(k->cur.lv[l - 1], L);
    }
    const dim3 grd2((L.w + 31) / 32, (L.h + 7) / 8);
    klt_scharr_kernel<<<grd2, blk, 0, s>>>(L);
  }
  int rc = launch_track(k, n, k->last_use_flow, k->last_min_ssim);
  if (rc) return rc;
  KLT_CUDA(k, cudaEventRecord(k->ctx->ev1, s));
  KLT_CUDA(k, cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, k->ctx->ev0, k->ctx->ev1);
  if (gpu_ms_out) *gpu_ms_out = ms;
  return 0;
}

int nrslam_b200_klt_get_patch(nrslam_b200_klt* k, int32_t idx, int16_t* gray_out, int16_t* grad_out, float* mean_out,
                              float* mean2_out, uint8_t* valid_out) {
  if (!k || idx < 0 || idx >= k->n || !gray_out || !grad_out || !mean_out || !mean2_out || !valid_out)
    return kfail(k, NRSLAM_B200_ERR_ARG, "klt_get_patch: bad argument");
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  const int nl = k->n_levels();
  const size_t slot = (size_t)idx * nl;
  cudaStream_t s = k->ctx->stream;
  KLT_CUDA(k, cudaMemcpyAsync(gray_out, k->ps.gray + slot * kArea, (size_t)nl * kArea * sizeof(short), cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaMemcpyAsync(grad_out, k->ps.grad + slot * kArea, (size_t)nl * kArea * sizeof(short2), cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaMemcpyAsync(mean_out, k->ps.mean + slot, nl * sizeof(float), cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaMemcpyAsync(mean2_out, k->ps.mean2 + slot, nl * sizeof(float), cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaMemcpyAsync(valid_out, k->ps.valid + slot, nl, cudaMemcpyDeviceToHost, s));
  KLT_CUDA(k, cudaStreamSynchronize(s));
  for (int l = 0; l < nl; l++)
    if (!valid_out[l]) {  // the reference holds an empty Mat
      memset(gray_out + (size_t)l * kArea, 0, kArea * sizeof(short));
      memset(grad_out + (size_t)l * kArea * 2, 0, kArea * 2 * sizeof(short));
    }
  return 0;
}

int nrslam_b200_klt_insert_patch(nrslam_b200_klt* k, float x, float y, const int16_t* gray, const int16_t* grad,
                                 const float* mean, const float* mean2, const uint8_t* valid) {
  if (!k || !gray || !grad || !mean || !mean2 || !valid) return kfail(k, NRSLAM_B200_ERR_ARG, "klt_insert_patch: bad argument");
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  const int rc = ensure_points(k, k->n + 1);
  if (rc) return rc;
  const int nl = k->n_levels();
  const size_t slot = (size_t)k->n * nl;
  cudaStream_t s = k->ctx->stream;
  const float xy[2] = {x, y};
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.gray + slot * kArea, gray, (size_t)nl * kArea * sizeof(short), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.grad + slot * kArea, grad, (size_t)nl * kArea * sizeof(short2), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.mean + slot, mean, nl * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.mean2 + slot, mean2, nl * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.valid + slot, valid, nl, cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->d_prev + 2 * (size_t)k->n, xy, 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaStreamSynchronize(s));
  k->n++;
  k->have_track = false;
  return 0;
}

// Batch form of InsertPhotometricInformation (lucas_kanade_tracker.cc:610-620): n points in one set of copies.
int nrslam_b200_klt_insert_patches(nrslam_b200_klt* k, int32_t n, const float* xy, const int16_t* gray,
                                   const int16_t* grad, const float* mean, const float* mean2, const uint8_t* valid) {
  if (!k || n < 0 || (n > 0 && (!xy || !gray || !grad || !mean || !mean2 || !valid)))
    return kfail(k, NRSLAM_B200_ERR_ARG, "klt_insert_patches: bad argument");
  if (n == 0) return 0;
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  const int rc = ensure_points(k, k->n + n);
  if (rc) return rc;
  const int nl = k->n_levels();
  const size_t slot = (size_t)k->n * nl, cnt = (size_t)n * nl;
  cudaStream_t s = k->ctx->stream;
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.gray + slot * kArea, gray, cnt * kArea * sizeof(short), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.grad + slot * kArea, grad, cnt * kArea * sizeof(short2), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.mean + slot, mean, cnt * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.mean2 + slot, mean2, cnt * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->ps.valid + slot, valid, cnt, cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaMemcpyAsync(k->d_prev + 2 * (size_t)k->n, xy, 2 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
  KLT_CUDA(k, cudaStreamSynchronize(s));
  k->n += n;
  k->have_track = false;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Tracking::PointReuse — modules/tracking/tracking.cc:394-506
// Candidates are the map points that are not in the frame with a 3-D position (in_frame[i] == 0) or that the optimiser
// reported lost (forced[i] != 0), whose projection through the frame pose has non-negative depth and lies inside the
// image (:397-414,431-440). They are tracked by a fresh 2-level tracker fed from their stored patches with the
// projection as initial flow and SSIM 0.75 (:421-458), and accepted when still TRACKED_WITH_3D and within 5.99 px^2 of
// the projection (:474-479). The reference walks an absl::flat_hash_set (unspecified order); candidates are taken in
// ascending index here. fp32 throughout like Sophus::SE3f * Eigen::Vector3f and CameraModel::Project.
// ---------------------------------------------------------------------------------------------------------------
int nrslam_b200_point_reuse(nrslam_b200_ctx* ctx, const nrslam_b200_camera* cam, const float* pose, const uint8_t* image,
                            int32_t width, int32_t height, int32_t pitch, const uint8_t* mask, int32_t mask_pitch,
                            int32_t n, const float* X_world, const uint8_t* in_frame, const uint8_t* forced,
                            int32_t max_iters, float epsilon, float min_eig_threshold, const int16_t* gray,
                            const int16_t* grad, const float* mean, const float* mean2, const uint8_t* valid,
                            int32_t* cand_out, float* seed_out, float* uv_out, uint8_t* status_out,
                            uint8_t* accepted_out, int32_t* n_cand_out, int32_t* n_reused_out) {
  if (n_cand_out) *n_cand_out = 0;
  if (n_reused_out) *n_reused_out = 0;
  if (!ctx || !cam || !pose || !image || n < 0 || (n > 0 && (!X_world || !in_frame || !gray || !grad || !mean || !mean2 ||
      !valid || !cand_out || !uv_out || !accepted_out)))
    return NRSLAM_B200_ERR_ARG;
  nrs::Cam c;
  c.model = cam->model;
  for (int i = 0; i < 8; i++) c.p[i] = cam->params[i];
  // Eigen::Quaternion::_transformVector: v + w t + q x t with t = 2 q x v (the form Sophus::SE3f * point evaluates)
  const float qx = pose[0], qy = pose[1], qz = pose[2], qw = pose[3];
  std::vector<int32_t> cand;
  std::vector<float> seeds;
  for (int i = 0; i < n; i++) {
    if (in_frame[i] && !(forced && forced[i])) continue;
    const float x = X_world[3 * (size_t)i], y = X_world[3 * (size_t)i + 1], z = X_world[3 * (size_t)i + 2];
    float tx = NRS_FS(NRS_FM(qy, z), NRS_FM(qz, y)), ty = NRS_FS(NRS_FM(qz, x), NRS_FM(qx, z)),
          tz = NRS_FS(NRS_FM(qx, y), NRS_FM(qy, x));
    tx = NRS_FA(tx, tx); ty = NRS_FA(ty, ty); tz = NRS_FA(tz, tz);
    const float cx_ = NRS_FS(NRS_FM(qy, tz), NRS_FM(qz, ty)), cy_ = NRS_FS(NRS_FM(qz, tx), NRS_FM(qx, tz)),
                cz_ = NRS_FS(NRS_FM(qx, ty), NRS_FM(qy, tx));
    const float px = NRS_FA(NRS_FA(NRS_FA(x, NRS_FM(qw, tx)), cx_), pose[4]);
    const float py = NRS_FA(NRS_FA(NRS_FA(y, NRS_FM(qw, ty)), cy_), pose[5]);
    const float pz = NRS_FA(NRS_FA(NRS_FA(z, NRS_FM(qw, tz)), cz_), pose[6]);
    if (pz < 0) continue;  // :403-405
    float u, v;
    nrs::project_f(c, px, py, pz, u, v);
    if (!(u >= 0 && u < (float)width && v >= 0 && v < (float)height)) continue;  // :409-412
    cand.push_back(i);
    seeds.push_back(u);
    seeds.push_back(v);
  }
  const int m = (int)cand.size();
  if (n_cand_out) *n_cand_out = m;
  if (m == 0) return 0;  // :416-418
  nrslam_b200_klt* k = nullptr;
  int rc = nrslam_b200_klt_create(ctx, kWin, 1, max_iters, epsilon, min_eig_threshold, &k);  // :424-426
  if (rc) return rc;
  // gather the candidates' two-level patches
  const size_t A = kArea;
  std::vector<int16_t> g2((size_t)m * 2 * A), d2((size_t)m * 2 * A * 2);
  std::vector<float> m1((size_t)m * 2), m2((size_t)m * 2);
  std::vector<uint8_t> va((size_t)m * 2);
  for (int j = 0; j < m; j++) {
    const size_t i = cand[j];
    memcpy(&g2[(size_t)j * 2 * A], gray + i * 2 * A, 2 * A * sizeof(int16_t));
    memcpy(&d2[(size_t)j * 4 * A], grad + i * 4 * A, 4 * A * sizeof(int16_t));
    m1[2 * j] = mean[2 * i]; m1[2 * j + 1] = mean[2 * i + 1];
    m2[2 * j] = mean2[2 * i]; m2[2 * j + 1] = mean2[2 * i + 1];
    va[2 * j] = valid[2 * i]; va[2 * j + 1] = valid[2 * i + 1];
  }
  rc = nrslam_b200_klt_insert_patches(k, m, seeds.data(), g2.data(), d2.data(), m1.data(), m2.data(), va.data());
  std::vector<float> pts(seeds);
  std::vector<uint8_t> st(m, NRSLAM_TRACKED_WITH_3D);
  int32_t n_tracked = 0;
  if (!rc)
    rc = nrslam_b200_klt_track(k, image, width, height, pitch, m, pts.data(), st.data(), 1, 0.75f, mask, mask_pitch,
                               &n_tracked);  // :456-458
  nrslam_b200_klt_destroy(k);
  if (rc) return rc;
  int reused = 0;
  for (int j = 0; j < m; j++) {
    cand_out[j] = cand[j];
    if (seed_out) { seed_out[2 * j] = seeds[2 * j]; seed_out[2 * j + 1] = seeds[2 * j + 1]; }
    uv_out[2 * j] = pts[2 * j];
    uv_out[2 * j + 1] = pts[2 * j + 1];
    if (status_out) status_out[j] = st[j];
    const float dx = NRS_FS(seeds[2 * j], pts[2 * j]), dy = NRS_FS(seeds[2 * j + 1], pts[2 * j + 1]);
    const float err2 = NRS_FA(NRS_FM(dx, dx), NRS_FM(dy, dy));  // SquaredReprojectionError, geometry_toolbox.cc
    accepted_out[j] = (st[j] == NRSLAM_TRACKED_WITH_3D && !(err2 > 5.99f)) ? 1 : 0;  // :474-479
    reused += accepted_out[j];
  }
  if (n_reused_out) *n_reused_out = reused;
  return 0;
}

int nrslam_b200_klt_debug_level(nrslam_b200_klt* k, int32_t which, int32_t level, uint8_t* img_out, int16_t* deriv_out,
                                int32_t* w_out, int32_t* h_out) {
  if (!k || level < 0 || level >= k->n_levels() || k->w == 0) return kfail(k, NRSLAM_B200_ERR_ARG, "klt_debug_level: bad argument");
  KLT_CUDA(k, cudaSetDevice(k->ctx->device));
  const LevelDev& L = (which ? k->cur : k->ref).lv[level];
  const size_t px = (size_t)(L.h + 2 * kWin) * L.stride;
  if (w_out) *w_out = L.w;
  if (h_out) *h_out = L.h;
  if (img_out) KLT_CUDA(k, cudaMemcpyAsync(img_out, L.img, px, cudaMemcpyDeviceToHost, k->ctx->stream));
  if (deriv_out) KLT_CUDA(k, cudaMemcpyAsync(deriv_out, L.deriv, px * sizeof(short2), cudaMemcpyDeviceToHost, k->ctx->stream));
  KLT_CUDA(k, cudaStreamSynchronize(k->ctx->stream));
  return 0;
}

int nrslam_b200_klt_clear(nrslam_b200_klt* k) {
  if (!k) return NRSLAM_B200_ERR_ARG;
  k->n = 0;
  k->have_track = false;
  return 0;
}

int32_t nrslam_b200_klt_num_points(const nrslam_b200_klt* k) { return k ? k->n : 0; }

}  // extern "C"
