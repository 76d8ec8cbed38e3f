// Micro-benchmark: FP64 FMA latency / throughput per SM, reciprocal / rsqrt / division latency, shared-memory load
// latency on the device at hand (one CTA). nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_fma(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); a = fma(a, b, 1e-9);
    a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); a = fma(a, b, 1e-9); a = fma(a, b, 1e-9);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[2 + threadIdx.x] = a;
}
template <int ILP>
__global__ void thr_fma(double* out, long long* cyc, int n) {
  double a[ILP], b = out[1];
#pragma unroll
  for (int k = 0; k < ILP; k++) a[k] = out[0] + k;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) a[k] = fma(a[k], b, 1e-9);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += a[k];
  out[2 + threadIdx.x] = s;
}
__global__ void lat_ops(double* out, long long* cyc, int n) {
  double a = out[0] + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = 1.0 / a + 1.5;
  long long t1 = clock64();
  for (int i = 0; i < n; i++) a = rsqrt(a) + 1.5;
  long long t2 = clock64();
  for (int i = 0; i < n; i++) a = sqrt(a) + 1.5;
  long long t3 = clock64();
  for (int i = 0; i < n; i++) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); r = fma(r, e, r);
    a = r + 1.5;
  }
  long long t4 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
  out[2 + threadIdx.x] = a;
}
__global__ void lat_smem(double* out, long long* cyc, int n) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) p = nxt[p];
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[2 + threadIdx.x] = p;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 64);
  double h[2] = {1.0000001, 0.9999999};
  cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
  long long c[4];
  const int n = 4096;
  lat_fma<<<1, 32>>>(out, cyc, n); cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.2f cycles\n", (double)c[0] / (8.0 * n));
  for (int threads : {32, 64, 128, 256, 512, 1024}) {
    thr_fma<8><<<1, threads>>>(out, cyc, n); cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA throughput %4d threads x ILP 8: %.2f FMA/clk/SM\n", threads, (double)threads * 8 * n / c[0]);
  }
  lat_ops<<<1, 32>>>(out, cyc, n); cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
  printf("latency: 1/x+add %.1f, rsqrt+add %.1f, sqrt+add %.1f, rcp.approx+3NR+add %.1f cycles\n", (double)c[0] / n,
         (double)c[1] / n, (double)c[2] / n, (double)c[3] / n);
  lat_smem<<<1, 32>>>(out, cyc, n); cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("shared-memory dependent load latency: %.1f cycles\n", (double)c[0] / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
