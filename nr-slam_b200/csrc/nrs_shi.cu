// nrs_shi.cu — Shi-Tomasi corner detector on sm_100a (C ABI: nrslam_b200_shi_*).
//
// What it replaces (reference paths relative to /root/reference):
//   ShiTomasi::Extract / GetKeyPoints / IsLocalMaximum      modules/features/shi_tomasi.cc:38-54,75-160
//   ShiTomasi::FastSobelXYandScore / DetectCorner           modules/features/shi_tomasi.cc:163-409
//
// Definition implemented (the "clean" mode of oracle/orc_shi.cc, bit-exact with it): 3x3 Sobel gradients (int16),
// 3x3 structure tensor as exact integer sums scaled once by 1/9 in fp32, min eigenvalue in fp32 with the reference's
// operation order; score rows 4 .. rows-5 and columns 1 .. cols-2 are the ones the reference computes from aligned
// gradients, everything outside is 0 (the reference leaves artefacts of its row-pointer rotation there, SURVEY App.
// E16). Already-tracked keypoints mark their pixel -1; non-maximum suppression with the reference's windows (inner
// +-N, exclusion +-15, score >= 80); keypoints come out in raster order and take consecutive class ids from the
// extractor's running counter (shi_tomasi.cc:81), which is the index bookkeeping the tracker relies on.
//
// Kernels: score map (one thread per pixel, 5x5 neighbourhood through the read-only path), mark, NMS flag map,
// per-row count + ordered compaction. Byte / integer work bound by L2 latency: 307k pixels, < 0.1 ms.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "nrs_host.h"

namespace {

__device__ __forceinline__ int px(const unsigned char* __restrict__ im, int pitch, int r, int c) {
  return (int)__ldg(im + (size_t)r * pitch + c);
}

__global__ void shi_score_kernel(const unsigned char* __restrict__ im, int pitch, int rows, int cols, float* score) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= cols || s >= rows) return;
  float v = 0.f;
  if (s >= 4 && s <= rows - 5 && c >= 1 && c <= cols - 2) {
    int g11 = 0, g12 = 0, g22 = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++) {
      const int r = s + dr;
#pragma unroll
      for (int dc = -1; dc <= 1; dc++) {
        const int cc = c + dc;
        int gx = 0;
        if (cc > 0 && cc < cols - 1)
          gx = (px(im, pitch, r - 1, cc + 1) + 2 * px(im, pitch, r, cc + 1) + px(im, pitch, r + 1, cc + 1)) -
               (px(im, pitch, r - 1, cc - 1) + 2 * px(im, pitch, r, cc - 1) + px(im, pitch, r + 1, cc - 1));
        int up, dn;
        if (cc == 0) {
          up = 2 * px(im, pitch, r - 1, 0) + 2 * px(im, pitch, r - 1, 1);
          dn = 2 * px(im, pitch, r + 1, 0) + 2 * px(im, pitch, r + 1, 1);
        } else if (cc == cols - 1) {
          up = 2 * px(im, pitch, r - 1, cols - 1) + 2 * px(im, pitch, r - 1, cols - 2);
          dn = 2 * px(im, pitch, r + 1, cols - 1) + 2 * px(im, pitch, r + 1, cols - 2);
        } else {
          up = px(im, pitch, r - 1, cc - 1) + 2 * px(im, pitch, r - 1, cc) + px(im, pitch, r - 1, cc + 1);
          dn = px(im, pitch, r + 1, cc - 1) + 2 * px(im, pitch, r + 1, cc) + px(im, pitch, r + 1, cc + 1);
        }
        const int gy = dn - up;
        g11 += gx * gx;
        g12 += gx * gy;
        g22 += gy * gy;
      }
    }
    const float inv_size = 1.f / 9.f;
    const float t0 = __fmul_rn((float)g11, inv_size), t1 = __fmul_rn((float)g12, inv_size),
                t2 = __fmul_rn((float)g22, inv_size);
    const float tr = __fadd_rn(t0, t2);
    const float det = __fsub_rn(__fmul_rn(t0, t2), __fmul_rn(t1, t1));
    const float root = __fsub_rn(__fmul_rn(tr, tr), __fmul_rn(4.f, det));
    v = __fmul_rn(__fsub_rn(tr, __fsqrt_rn(root)), 0.5f);
  }
  score[(size_t)s * cols + c] = v;
}

__global__ void shi_mark_kernel(const float* __restrict__ pts, int n, int rows, int cols, float* score) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // round(): half away from zero on the double value (shi_tomasi.cc:93-95)
  const int r = (int)round((double)pts[2 * i + 1]), c = (int)round((double)pts[2 * i]);
  if (r >= 0 && r < rows && c >= 0 && c < cols) score[(size_t)r * cols + c] = -1.f;
}

// IsLocalMaximum (:123-160): flag map
__global__ void shi_nms_kernel(const float* __restrict__ score, int rows, int cols, int nms, unsigned char* flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= cols || r >= rows) return;
  const float cur = score[(size_t)r * cols + c];
  bool ok = !(cur == -1.f) && !(cur < 80.f);
  if (ok) {
    const int NPrev = 15;
    const int r0 = max(0, r - NPrev), r1 = min(rows - 1, r + NPrev), c0 = max(0, c - NPrev), c1 = min(cols - 1, c + NPrev);
    const int ri0 = max(0, r - nms), ri1 = min(rows - 1, r + nms), ci0 = max(0, c - nms), ci1 = min(cols - 1, c + nms);
    for (int i = r0; i <= r1 && ok; i++) {
      const float* row = score + (size_t)i * cols;
      const bool inner_row = i >= ri0 && i <= ri1;
      for (int j = c0; j <= c1; j++) {
        const float v = __ldg(row + j);
        if (v == -1.f || (inner_row && j >= ci0 && j <= ci1 && v > cur)) {
          ok = false;
          break;
        }
      }
    }
  }
  flag[(size_t)r * cols + c] = ok ? 1 : 0;
}

// one warp per row: number of flagged pixels
__global__ void shi_row_count_kernel(const unsigned char* __restrict__ flag, int rows, int cols, int* row_count) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  int n = 0;
  for (int c = lane; c < cols; c += 32) n += flag[(size_t)r * cols + c];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
  if (lane == 0) row_count[r] = n;
}

// exclusive scan of the row counts (single block) -> row offsets, total
__global__ void shi_scan_kernel(const int* __restrict__ row_count, int rows, int* row_offset, int* total) {
  __shared__ int part[1024];
  const int t = threadIdx.x, per = (rows + blockDim.x - 1) / blockDim.x;
  int s = 0;
  for (int k = 0; k < per; k++) {
    const int r = t * per + k;
    if (r < rows) s += row_count[r];
  }
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    int acc = 0;
    for (int i = 0; i < (int)blockDim.x; i++) {
      const int v = part[i];
      part[i] = acc;
      acc += v;
    }
    *total = acc;
  }
  __syncthreads();
  int acc = part[t];
  for (int k = 0; k < per; k++) {
    const int r = t * per + k;
    if (r < rows) {
      row_offset[r] = acc;
      acc += row_count[r];
    }
  }
}

// one warp per row: ordered write of the keypoints of the row
__global__ void shi_compact_kernel(const unsigned char* __restrict__ flag, int rows, int cols,
                                   const int* __restrict__ row_offset, int capacity, float* out_xy) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  int base = row_offset[r];
  for (int c0 = 0; c0 < cols; c0 += 32) {
    const int c = c0 + lane;
    const bool f = c < cols && flag[(size_t)r * cols + c];
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (f) {
      const int idx = base + __popc(m & ((1u << lane) - 1));
      if (idx < capacity) {
        out_xy[2 * idx] = (float)c;
        out_xy[2 * idx + 1] = (float)r;
      }
    }
    base += __popc(m);
  }
}

}  // namespace

struct nrslam_b200_shi {
  nrslam_b200_ctx* ctx = nullptr;
  int nms = 7;
  unsigned next_id = 0;
  int w = 0, h = 0, cap = 0;
  unsigned char *d_im = nullptr, *d_flag = nullptr, *h_im = nullptr;
  float *d_score = nullptr, *d_pts = nullptr, *d_out = nullptr, *h_out = nullptr, *h_pts = nullptr;
  int *d_row_count = nullptr, *d_row_offset = nullptr, *d_total = nullptr, *h_total = nullptr;
  int pts_cap = 0;
};

namespace {
int sfail(nrslam_b200_shi* s, int code, const std::string& msg) {
  if (s && s->ctx) s->ctx->err = msg;
  return code;
}
#define SHI_CUDA(s, call)                                                                              \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) return sfail(s, NRSLAM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

void shi_free(nrslam_b200_shi* s) {
  if (s->d_im) cudaFree(s->d_im);
  if (s->d_flag) cudaFree(s->d_flag);
  if (s->h_im) cudaFreeHost(s->h_im);
  if (s->d_score) cudaFree(s->d_score);
  if (s->d_out) cudaFree(s->d_out);
  if (s->h_out) cudaFreeHost(s->h_out);
  if (s->d_row_count) cudaFree(s->d_row_count);
  if (s->d_row_offset) cudaFree(s->d_row_offset);
  if (s->d_total) cudaFree(s->d_total);
  if (s->h_total) cudaFreeHost(s->h_total);
  s->d_im = s->d_flag = s->h_im = nullptr;
  s->d_score = s->d_out = s->h_out = nullptr;
  s->d_row_count = s->d_row_offset = s->d_total = s->h_total = nullptr;
}
}  // namespace

extern "C" {

int nrslam_b200_shi_create(nrslam_b200_ctx* ctx, int32_t nms_window, nrslam_b200_shi** out) {
  if (out) *out = nullptr;
  if (!ctx || !out || nms_window < 0 || nms_window > 15) return NRSLAM_B200_ERR_ARG;
  nrslam_b200_shi* s = new nrslam_b200_shi();
  s->ctx = ctx;
  s->nms = nms_window;
  *out = s;
  return 0;
}

void nrslam_b200_shi_destroy(nrslam_b200_shi* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  shi_free(s);
  if (s->d_pts) cudaFree(s->d_pts);
  if (s->h_pts) cudaFreeHost(s->h_pts);
  delete s;
}

int nrslam_b200_shi_extract(nrslam_b200_shi* s, const uint8_t* image, int32_t width, int32_t height, int32_t pitch,
                            const float* existing_xy, int32_t n_existing, float* out_xy, int32_t* out_class_id,
                            int32_t capacity, int32_t* n_out) {
  if (!s || !image || !n_out || width < 8 || height < 8 || pitch < width || n_existing < 0 || capacity < 0 ||
      (n_existing > 0 && !existing_xy) || (capacity > 0 && (!out_xy || !out_class_id)))
    return sfail(s, NRSLAM_B200_ERR_ARG, "shi_extract: bad argument");
  SHI_CUDA(s, cudaSetDevice(s->ctx->device));
  cudaStream_t st = s->ctx->stream;
  const size_t npx = (size_t)width * height;
  if (width != s->w || height != s->h || capacity > s->cap) {  // ResizeBuffers (:56-67)
    shi_free(s);
    s->w = width;
    s->h = height;
    s->cap = std::max(capacity, 4096);
    SHI_CUDA(s, cudaMalloc(&s->d_im, npx));
    SHI_CUDA(s, cudaMalloc(&s->d_flag, npx));
    SHI_CUDA(s, cudaMallocHost(&s->h_im, npx));
    SHI_CUDA(s, cudaMalloc(&s->d_score, npx * sizeof(float)));
    SHI_CUDA(s, cudaMalloc(&s->d_out, (size_t)s->cap * 2 * sizeof(float)));
    SHI_CUDA(s, cudaMallocHost(&s->h_out, (size_t)s->cap * 2 * sizeof(float)));
    SHI_CUDA(s, cudaMalloc(&s->d_row_count, height * sizeof(int)));
    SHI_CUDA(s, cudaMalloc(&s->d_row_offset, height * sizeof(int)));
    SHI_CUDA(s, cudaMalloc(&s->d_total, sizeof(int)));
    SHI_CUDA(s, cudaMallocHost(&s->h_total, sizeof(int)));
  }
  if (n_existing > s->pts_cap) {
    if (s->d_pts) cudaFree(s->d_pts);
    if (s->h_pts) cudaFreeHost(s->h_pts);
    s->pts_cap = n_existing + n_existing / 2 + 256;
    SHI_CUDA(s, cudaMalloc(&s->d_pts, (size_t)s->pts_cap * 2 * sizeof(float)));
    SHI_CUDA(s, cudaMallocHost(&s->h_pts, (size_t)s->pts_cap * 2 * sizeof(float)));
  }
  for (int y = 0; y < height; y++) memcpy(s->h_im + (size_t)y * width, image + (size_t)y * pitch, width);
  SHI_CUDA(s, cudaMemcpyAsync(s->d_im, s->h_im, npx, cudaMemcpyHostToDevice, st));
  const dim3 blk(32, 8), grd((width + 31) / 32, (height + 7) / 8);
  shi_score_kernel<<<grd, blk, 0, st>>>(s->d_im, width, height, width, s->d_score);
  if (n_existing > 0) {
    memcpy(s->h_pts, existing_xy, (size_t)n_existing * 2 * sizeof(float));
    SHI_CUDA(s, cudaMemcpyAsync(s->d_pts, s->h_pts, (size_t)n_existing * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    shi_mark_kernel<<<(n_existing + 127) / 128, 128, 0, st>>>(s->d_pts, n_existing, height, width, s->d_score);
  }
  shi_nms_kernel<<<grd, blk, 0, st>>>(s->d_score, height, width, s->nms, s->d_flag);
  const int row_blocks = (height * 32 + 127) / 128;
  shi_row_count_kernel<<<row_blocks, 128, 0, st>>>(s->d_flag, height, width, s->d_row_count);
  shi_scan_kernel<<<1, 256, 0, st>>>(s->d_row_count, height, s->d_row_offset, s->d_total);
  shi_compact_kernel<<<row_blocks, 128, 0, st>>>(s->d_flag, height, width, s->d_row_offset, s->cap, s->d_out);
  SHI_CUDA(s, cudaGetLastError());
  SHI_CUDA(s, cudaMemcpyAsync(s->h_total, s->d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
  SHI_CUDA(s, cudaMemcpyAsync(s->h_out, s->d_out, (size_t)s->cap * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  SHI_CUDA(s, cudaStreamSynchronize(st));
  const int total = *s->h_total;
  const int n_copy = std::min(std::min(total, capacity), s->cap);
  for (int i = 0; i < n_copy; i++) {
    out_xy[2 * i] = s->h_out[2 * i];
    out_xy[2 * i + 1] = s->h_out[2 * i + 1];
    out_class_id[i] = (int32_t)(s->next_id + (unsigned)i);  // kp.class_id = next_feature_id_++ in raster order (:81)
  }
  s->next_id += (unsigned)total;
  *n_out = total;
  return 0;
}

int nrslam_b200_shi_debug_scores(nrslam_b200_shi* s, float* scores_out) {
  if (!s || !scores_out || s->w == 0) return sfail(s, NRSLAM_B200_ERR_ARG, "shi_debug_scores: bad argument");
  SHI_CUDA(s, cudaSetDevice(s->ctx->device));
  SHI_CUDA(s, cudaMemcpy(scores_out, s->d_score, (size_t)s->w * s->h * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
