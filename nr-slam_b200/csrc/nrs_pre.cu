// nrs_pre.cu — per-frame image pre-processing on sm_100a (C ABI: nrslam_b200_pre_*), so that a frame never has to be
// touched by the CPU between capture and the KLT / Shi-Tomasi kernels.
//
// What it replaces (reference paths relative to /root/reference):
//   System::ImageProcessing                      modules/SLAM/system.cc:189-201   cvtColor(RGB2GRAY) + CLAHE(3.0, 8x8)
//   BrightFilter / BorderFilter::generateMask    modules/masking/bright_filter.cc:24-39, border_filter.cc:24-40
//   Masker::mask, GetAllMasks()["Global"]        modules/masking/masker.cc:80-92,94-115
// The arithmetic is OpenCV's (un-vendored). Every kernel is integer / byte work restated from OpenCV's published
// algorithms and bit-exact with oracle/orc_preproc.py, which is pinned on cv2 golden vectors:
//   gray      (9798 R + 19235 G + 3735 B + 2^14) >> 15
//   CLAHE     one CTA per tile: shared-memory histogram, clip + redistribution in closed form, 256-wide scan, LUT;
//             then one thread per pixel blends the 4 neighbouring tile LUTs in fp32 with OpenCV's operation order —
//             explicit __fmul_rn / __fadd_rn: a fused multiply-add would change the rounding of ~0.1 % of the pixels
//   erode     rectangles separably (row minimum, column minimum), the 11x11 ellipse directly
//   Gaussian  11x11 sigma 5 in 8.8 fixed point, separable, REFLECT_101
// All of it is bound by launch latency and L2 (a 640x480 frame is 0.3 MB): ~15 small launches, no host round trip.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "nrs_host.h"

struct nrslam_b200_pre {
  nrslam_b200_ctx* ctx = nullptr;
  int max_w = 0, max_h = 0;
  unsigned char *d_rgb = nullptr, *d_gray = nullptr, *d_eq = nullptr, *d_luts = nullptr;
  unsigned char* d_m[4] = {nullptr, nullptr, nullptr, nullptr};
  unsigned short* d_t16 = nullptr;
  unsigned char *h_in = nullptr, *h_out = nullptr;  // pinned staging
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float last_ms = 0.f;
  int launches = 0;
};

namespace {

constexpr int kMaxTiles = 16;  // per axis

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void pre_gray_kernel(const unsigned char* __restrict__ rgb, int pitch, int w, int h, unsigned char* gray) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const unsigned char* p = rgb + (size_t)y * pitch + 3 * x;
  const int r = __ldg(p), g = __ldg(p + 1), b = __ldg(p + 2);
  gray[(size_t)y * w + x] = (unsigned char)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
}

// One CTA (256 threads) per tile: histogram of the (REFLECT_101-extended) tile, clip, redistribute, cdf -> LUT.
__global__ void __launch_bounds__(256) pre_clahe_lut_kernel(const unsigned char* __restrict__ gray, int w, int h,
                                                            int tiles_x, int tw, int th, int clip_limit,
                                                            float lut_scale, unsigned char* luts) {
  __shared__ int hist[256];
  __shared__ int scan[256];
  __shared__ int s_clipped;
  const int t = threadIdx.x;
  const int tile = blockIdx.x, ti = tile % tiles_x, tj = tile / tiles_x;
  hist[t] = 0;
  if (t == 0) s_clipped = 0;
  __syncthreads();
  const int n = tw * th;
  for (int k = t; k < n; k += 256) {
    const int ex = ti * tw + k % tw, ey = tj * th + k / tw;
    atomicAdd(&hist[__ldg(gray + (size_t)reflect101(ey, h) * w + reflect101(ex, w))], 1);
  }
  __syncthreads();
  int v = hist[t];
  if (clip_limit > 0) {
    const int over = max(v - clip_limit, 0);
    if (over) atomicAdd(&s_clipped, over);
    v = min(v, clip_limit);
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    const int resid = clipped - batch * 256;
    v += batch;
    if (resid != 0) {  // bins 0, step, 2 step, ... take one more until the residual is used up
      const int step = max(256 / resid, 1);
      if (t % step == 0 && t / step < resid) v += 1;
    }
  }
  scan[t] = v;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {  // inclusive scan
    const int add = (t >= off) ? scan[t - off] : 0;
    __syncthreads();
    scan[t] += add;
    __syncthreads();
  }
  const int r = __float2int_rn(__fmul_rn((float)scan[t], lut_scale));
  luts[(size_t)tile * 256 + t] = (unsigned char)min(max(r, 0), 255);
}

__global__ void pre_clahe_apply_kernel(const unsigned char* __restrict__ gray, const unsigned char* __restrict__ luts,
                                       int w, int h, int tiles_x, int tiles_y, float inv_tw, float inv_th,
                                       unsigned char* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
  int tx1 = (int)floorf(txf);
  const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
  const int tx2 = min(tx1 + 1, tiles_x - 1);
  tx1 = max(tx1, 0);
  const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
  int ty1 = (int)floorf(tyf);
  const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
  const int ty2 = min(ty1 + 1, tiles_y - 1);
  ty1 = max(ty1, 0);
  const int v = gray[(size_t)y * w + x];
  const float l11 = (float)__ldg(luts + ((size_t)(ty1 * tiles_x + tx1)) * 256 + v);
  const float l12 = (float)__ldg(luts + ((size_t)(ty1 * tiles_x + tx2)) * 256 + v);
  const float l21 = (float)__ldg(luts + ((size_t)(ty2 * tiles_x + tx1)) * 256 + v);
  const float l22 = (float)__ldg(luts + ((size_t)(ty2 * tiles_x + tx2)) * 256 + v);
  const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
  const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
  const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
  const int r = __float2int_rn(res);
  out[(size_t)y * w + x] = (unsigned char)min(max(r, 0), 255);
}

__global__ void pre_threshold_inv_kernel(const unsigned char* __restrict__ gray, int n, int th, unsigned char* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = gray[i] > th ? 0 : 255;
}

__global__ void pre_border_kernel(const unsigned char* __restrict__ gray, int w, int h, int rb, int re, int cb, int ce,
                                  unsigned char* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const bool in = y >= rb && y < h - re && x >= cb && x < w - ce;
  out[(size_t)y * w + x] = (in && gray[(size_t)y * w + x] != 0) ? 255 : 0;
}

// min over [i - anchor, i - anchor + k) along x (dir 0) or y (dir 1); samples outside the image do not constrain
__global__ void pre_erode_line_kernel(const unsigned char* __restrict__ in, int w, int h, int k, int anchor, int dir,
                                      unsigned char* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  int m = 255;
  if (dir == 0) {
    const int a = max(x - anchor, 0), b = min(x - anchor + k, w);
    for (int c = a; c < b; c++) m = min(m, (int)__ldg(in + (size_t)y * w + c));
  } else {
    const int a = max(y - anchor, 0), b = min(y - anchor + k, h);
    for (int r = a; r < b; r++) m = min(m, (int)__ldg(in + (size_t)r * w + x));
  }
  out[(size_t)y * w + x] = (unsigned char)m;
}

struct EllipseRows {
  int size;
  int half[32];  // per row of the element: half width of its run around the centre column, -1: empty
};

__global__ void pre_erode_ellipse_kernel(const unsigned char* __restrict__ in, int w, int h, EllipseRows el,
                                         unsigned char* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int c = el.size / 2;
  int m = 255;
  for (int i = 0; i < el.size; i++) {
    const int r = y + i - c;
    if (r < 0 || r >= h || el.half[i] < 0) continue;
    const int a = max(x - el.half[i], 0), b = min(x + el.half[i] + 1, w);
    for (int cc = a; cc < b; cc++) m = min(m, (int)__ldg(in + (size_t)r * w + cc));
  }
  out[(size_t)y * w + x] = (unsigned char)m;
}

__constant__ int c_gauss11[11] = {17, 20, 24, 26, 27, 28, 27, 26, 24, 20, 17};

__global__ void pre_gauss_h_kernel(const unsigned char* __restrict__ in, int w, int h, unsigned short* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  int s = 0;
#pragma unroll
  for (int i = 0; i < 11; i++) s += c_gauss11[i] * (int)__ldg(in + (size_t)y * w + reflect101(x + i - 5, w));
  out[(size_t)y * w + x] = (unsigned short)s;  // 8.8 fixed point, <= 255 * 256
}

__global__ void pre_gauss_v_kernel(const unsigned short* __restrict__ in, int w, int h, unsigned char* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  unsigned int s = 0;
#pragma unroll
  for (int i = 0; i < 11; i++) s += (unsigned)c_gauss11[i] * (unsigned)__ldg(in + (size_t)reflect101(y + i - 5, h) * w + x);
  out[(size_t)y * w + x] = (unsigned char)min((s + (1u << 15)) >> 16, 255u);
}

__global__ void pre_and_kernel(const unsigned char* a, const unsigned char* b, int n, unsigned char* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] & b[i];
}

__global__ void pre_fill_kernel(unsigned char* out, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (unsigned char)v;
}

int pfail(nrslam_b200_pre* p, int code, const std::string& msg) {
  if (p && p->ctx) p->ctx->err = msg;
  return code;
}

#define PRE_CUDA(p, call)                                                                           \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return pfail(p, NRSLAM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

inline dim3 grid2(int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); }
const dim3 kBlk2(32, 8);

EllipseRows ellipse_rows(int size) {  // cv::getStructuringElement(MORPH_ELLIPSE)
  EllipseRows el;
  el.size = size;
  const int r = size / 2, c = size / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  for (int i = 0; i < 32; i++) el.half[i] = -1;
  for (int i = 0; i < size; i++) {
    const int dy = i - r;
    if (std::abs(dy) <= r) el.half[i] = (int)std::lrint(c * std::sqrt((r * r - dy * dy) * inv_r2));
  }
  return el;
}

void erode_rect(nrslam_b200_pre* p, cudaStream_t st, unsigned char* buf, unsigned char* tmp, int w, int h, int k) {
  pre_erode_line_kernel<<<grid2(w, h), kBlk2, 0, st>>>(buf, w, h, k, k / 2, 0, tmp);
  pre_erode_line_kernel<<<grid2(w, h), kBlk2, 0, st>>>(tmp, w, h, k, k / 2, 1, buf);
  p->launches += 2;
}

}  // namespace

extern "C" {

int nrslam_b200_pre_create(nrslam_b200_ctx* ctx, int32_t max_width, int32_t max_height, nrslam_b200_pre** out) {
  if (out) *out = nullptr;
  if (!ctx || !out || max_width < 16 || max_height < 16) return NRSLAM_B200_ERR_ARG;
  nrslam_b200_pre* p = new nrslam_b200_pre();
  p->ctx = ctx;
  p->max_w = max_width;
  p->max_h = max_height;
  const size_t n = (size_t)max_width * max_height;
  cudaSetDevice(ctx->device);
  bool ok = cudaMalloc(&p->d_rgb, 3 * n) == cudaSuccess && cudaMalloc(&p->d_gray, n) == cudaSuccess &&
            cudaMalloc(&p->d_eq, n) == cudaSuccess &&
            cudaMalloc(&p->d_luts, (size_t)kMaxTiles * kMaxTiles * 256) == cudaSuccess &&
            cudaMalloc(&p->d_t16, n * sizeof(unsigned short)) == cudaSuccess &&
            cudaMallocHost(&p->h_in, 3 * n) == cudaSuccess && cudaMallocHost(&p->h_out, 2 * n) == cudaSuccess &&
            cudaEventCreate(&p->e0) == cudaSuccess && cudaEventCreate(&p->e1) == cudaSuccess;
  for (int i = 0; i < 4 && ok; i++) ok = cudaMalloc(&p->d_m[i], n) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    nrslam_b200_pre_destroy(p);
    ctx->err = "pre_create: allocation failed";
    return NRSLAM_B200_ERR_ALLOC;
  }
  *out = p;
  return 0;
}

void nrslam_b200_pre_destroy(nrslam_b200_pre* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  cudaFree(p->d_rgb);
  cudaFree(p->d_gray);
  cudaFree(p->d_eq);
  cudaFree(p->d_luts);
  cudaFree(p->d_t16);
  for (int i = 0; i < 4; i++) cudaFree(p->d_m[i]);
  if (p->h_in) cudaFreeHost(p->h_in);
  if (p->h_out) cudaFreeHost(p->h_out);
  if (p->e0) cudaEventDestroy(p->e0);
  if (p->e1) cudaEventDestroy(p->e1);
  delete p;
}

int nrslam_b200_pre_image(nrslam_b200_pre* p, const uint8_t* rgb, int32_t width, int32_t height, int32_t pitch,
                          float clip_limit, int32_t tiles_x, int32_t tiles_y, uint8_t* gray_out, uint8_t* clahe_out) {
  if (!p || !rgb || width < 16 || height < 16 || width > p->max_w || height > p->max_h || pitch < 3 * width ||
      tiles_x < 1 || tiles_y < 1 || tiles_x > kMaxTiles || tiles_y > kMaxTiles)
    return pfail(p, NRSLAM_B200_ERR_ARG, "pre_image: bad argument");
  PRE_CUDA(p, cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int w = width, h = height;
  const size_t n = (size_t)w * h;
  for (int y = 0; y < h; y++) memcpy(p->h_in + (size_t)y * 3 * w, rgb + (size_t)y * pitch, 3 * (size_t)w);
  PRE_CUDA(p, cudaMemcpyAsync(p->d_rgb, p->h_in, 3 * n, cudaMemcpyHostToDevice, st));
  PRE_CUDA(p, cudaEventRecord(p->e0, st));
  p->launches = 0;
  pre_gray_kernel<<<grid2(w, h), kBlk2, 0, st>>>(p->d_rgb, 3 * w, w, h, p->d_gray);
  // CLAHE geometry (cv::CLAHE_Impl::apply): a ragged grid extends the image to the right / bottom by REFLECT_101
  int ew = w, eh = h;
  if (w % tiles_x != 0 || h % tiles_y != 0) {
    ew = w + (tiles_x - w % tiles_x);
    eh = h + (tiles_y - h % tiles_y);
  }
  const int tw = ew / tiles_x, th = eh / tiles_y;
  const int area = tw * th;
  const float lut_scale = 255.0f / (float)area;
  int clip = 0;
  if (clip_limit > 0.f) clip = std::max((int)((double)clip_limit * area / 256), 1);  // double, as cv::CLAHE_Impl
  pre_clahe_lut_kernel<<<tiles_x * tiles_y, 256, 0, st>>>(p->d_gray, w, h, tiles_x, tw, th, clip, lut_scale, p->d_luts);
  pre_clahe_apply_kernel<<<grid2(w, h), kBlk2, 0, st>>>(p->d_gray, p->d_luts, w, h, tiles_x, tiles_y, 1.0f / (float)tw,
                                                          1.0f / (float)th, p->d_eq);
  p->launches += 3;
  PRE_CUDA(p, cudaEventRecord(p->e1, st));
  PRE_CUDA(p, cudaGetLastError());
  if (gray_out) PRE_CUDA(p, cudaMemcpyAsync(p->h_out, p->d_gray, n, cudaMemcpyDeviceToHost, st));
  if (clahe_out) PRE_CUDA(p, cudaMemcpyAsync(p->h_out + n, p->d_eq, n, cudaMemcpyDeviceToHost, st));
  PRE_CUDA(p, cudaStreamSynchronize(st));
  if (gray_out) memcpy(gray_out, p->h_out, n);
  if (clahe_out) memcpy(clahe_out, p->h_out + n, n);
  cudaEventElapsedTime(&p->last_ms, p->e0, p->e1);
  return 0;
}

int nrslam_b200_pre_mask(nrslam_b200_pre* p, const uint8_t* gray, int32_t width, int32_t height,
                         const nrslam_b200_mask_filter* filters, int32_t n_filters, uint8_t* mask_out) {
  if (!p || !mask_out || width < 16 || height < 16 || width > p->max_w || height > p->max_h || n_filters < 0 ||
      (n_filters > 0 && !filters))
    return pfail(p, NRSLAM_B200_ERR_ARG, "pre_mask: bad argument");
  PRE_CUDA(p, cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int w = width, h = height;
  const int n = w * h;
  if (gray) {  // else: the gray image of the last pre_image call, still resident
    memcpy(p->h_in, gray, n);
    PRE_CUDA(p, cudaMemcpyAsync(p->d_gray, p->h_in, n, cudaMemcpyHostToDevice, st));
  }
  PRE_CUDA(p, cudaEventRecord(p->e0, st));
  p->launches = 0;
  unsigned char *acc = p->d_m[0], *cur = p->d_m[1], *tmp = p->d_m[2];
  const int g1 = (n + 255) / 256;
  pre_fill_kernel<<<g1, 256, 0, st>>>(acc, n, 255);
  p->launches++;
  for (int f = 0; f < n_filters; f++) {
    const nrslam_b200_mask_filter& F = filters[f];
    if (F.kind == NRSLAM_B200_FILTER_BRIGHT) {  // bright_filter.cc:24-39
      pre_threshold_inv_kernel<<<g1, 256, 0, st>>>(p->d_gray, n, F.th, cur);
      pre_erode_ellipse_kernel<<<grid2(w, h), kBlk2, 0, st>>>(cur, w, h, ellipse_rows(11), tmp);
      pre_gauss_h_kernel<<<grid2(w, h), kBlk2, 0, st>>>(tmp, w, h, p->d_t16);
      pre_gauss_v_kernel<<<grid2(w, h), kBlk2, 0, st>>>(p->d_t16, w, h, cur);
      p->launches += 4;
    } else if (F.kind == NRSLAM_B200_FILTER_BORDER) {  // border_filter.cc:24-40
      if (F.rb < 0 || F.re < 0 || F.cb < 0 || F.ce < 0 || F.rb + F.re > h || F.cb + F.ce > w)
        return pfail(p, NRSLAM_B200_ERR_ARG, "pre_mask: border larger than the image");
      pre_border_kernel<<<grid2(w, h), kBlk2, 0, st>>>(p->d_gray, w, h, F.rb, F.re, F.cb, F.ce, cur);
      p->launches++;
      erode_rect(p, st, cur, tmp, w, h, 21);
    } else if (F.kind == NRSLAM_B200_FILTER_PREDEFINED) {  // predefined_filter.cc:39-41 (mask prepared by the caller)
      if (!F.mask) return pfail(p, NRSLAM_B200_ERR_ARG, "pre_mask: predefined filter without a mask");
      memcpy(p->h_in + n, F.mask, n);
      PRE_CUDA(p, cudaMemcpyAsync(cur, p->h_in + n, n, cudaMemcpyHostToDevice, st));
      PRE_CUDA(p, cudaStreamSynchronize(st));  // the staging slot is reused by the next predefined filter
    } else {
      return pfail(p, NRSLAM_B200_ERR_ARG, "pre_mask: unknown filter kind");
    }
    pre_and_kernel<<<g1, 256, 0, st>>>(acc, cur, n, acc);  // masker.cc:85-87
    p->launches++;
  }
  erode_rect(p, st, acc, tmp, w, h, 10);  // masker.cc:89-90
  PRE_CUDA(p, cudaEventRecord(p->e1, st));
  PRE_CUDA(p, cudaGetLastError());
  PRE_CUDA(p, cudaMemcpyAsync(p->h_out, acc, n, cudaMemcpyDeviceToHost, st));
  PRE_CUDA(p, cudaStreamSynchronize(st));
  memcpy(mask_out, p->h_out, n);
  cudaEventElapsedTime(&p->last_ms, p->e0, p->e1);
  return 0;
}

float nrslam_b200_pre_last_ms(const nrslam_b200_pre* p) { return p ? p->last_ms : 0.f; }
int32_t nrslam_b200_pre_last_launches(const nrslam_b200_pre* p) { return p ? p->launches : 0; }

}  // extern "C"
