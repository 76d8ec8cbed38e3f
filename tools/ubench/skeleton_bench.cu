// Micro-benchmark: cost of the synchronisation SKELETON of one cluster-native CG iteration (nrs_engine.cu: pcg_cluster)
// with no arithmetic: 5 CTA barriers + 2 cluster barriers + 2 pushes of reduction values into every CTA's gather
// buffer (remote shared-memory stores) + local reads. Variants drop pieces to attribute the cost.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void cbar() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mode bits: 1 = CTA barriers, 2 = cluster barriers, 4 = push + local read, 8 = warp-shuffle reductions like the real loop
__global__ void skel(int iters, int mode, long long* cycles, double* sink) {
  __shared__ double gather[2][16][8];
  __shared__ double slot[8];
  __shared__ double red[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double acc = tid;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    for (int half = 0; half < 2; half++) {
      double v = acc * 1e-9;
      if (mode & 8) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp] = v;
      }
      if (mode & 1) __syncthreads();  // S1 / S4
      if (warp == 0) {
        double t = 0;
        if (mode & 8) {
          for (int w = lane; w < nw; w += 32) t += red[w];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        }
        if (lane == 0) slot[half ? 7 : 0] = t;
        __syncwarp();
        if ((mode & 4) && lane < G) {
          double* dst = cluster.map_shared_rank(&gather[0][0][0], lane) + (size_t)(par * 16 + blockIdx.x) * 8;
          const int k0 = half ? 7 : 0, nk = half ? 1 : 7;
          for (int k = 0; k < nk; k++) dst[k0 + k] = slot[half ? 7 : 0];
        }
      }
      if (mode & 2) cbar();  // B1 / B2
      if (warp == 0) {
        double t = 0;
        if ((mode & 4) && lane < G) t = gather[par][lane][half ? 7 : 0];
        if (mode & 8) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        }
        if (lane == 0) slot[1] = t;
      }
      if (mode & 1) __syncthreads();  // S2 / S5
      acc += slot[1];
      if (half == 0 && (mode & 1)) __syncthreads();  // S3
    }
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  sink[blockIdx.x * blockDim.x + tid] = acc;
  cbar();
}

int main() {
  long long* cyc;
  double* sink;
  cudaMalloc(&cyc, 8);
  cudaMalloc(&sink, 16 * 1024 * 8);
  const int iters = 2000;
  cudaFuncSetAttribute(skel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int block : {128, 256, 512}) {
    for (int mode : {1, 2, 3, 7, 15}) {
      for (int rep = 0; rep < 2; rep++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16);
        cfg.blockDim = dim3(block);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int it = iters, m = mode;
        void* args[] = {&it, &m, &cyc, &sink};
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)skel, args);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) {
          printf("launch error %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
          return 1;
        }
      }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("block %3d mode %2d (%s%s%s%s): %6lld cycles / iteration\n", block, mode, (mode & 1) ? "5xCTA-bar " : "",
             (mode & 2) ? "2xcluster-bar " : "", (mode & 4) ? "push+read " : "", (mode & 8) ? "shuffle-reductions" : "",
             c / iters);
    }
  }
  return 0;
}
