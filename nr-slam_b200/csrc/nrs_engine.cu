// nrs_engine.cu — persistent Levenberg–Marquardt kernel for sm_100a.
//
// What it replaces (reference paths relative to /root/reference):
//   SparseOptimizer::optimize / computeActiveErrors / update / push / pop   third_party/g2o/g2o/core/sparse_optimizer.cpp:62-114,392-470
//   OptimizationAlgorithmLevenberg::solve (lambda control, gain ratio)      third_party/g2o/g2o/core/optimization_algorithm_levenberg.cpp:57-174
//   BlockSolver::buildSystem / setLambda / solve                           third_party/g2o/g2o/core/block_solver.hpp:329-341,495-603
//   BaseFixedSizedEdge::constructQuadraticForm                             third_party/g2o/g2o/core/base_fixed_sized_edge.hpp:49-133
//   the ten edge types of modules/optimization/*.cc (cited at each formula)
//   the round / re-levelling logic of modules/optimization/g2o_optimization.cc:100-140,338-395
//
// Design. One launch runs a whole driver program. Each CTA owns "chunks" of point rows; a row is a
// point vertex with its reprojection edge and the regulariser edges incident to it, so the normal equations are
// applied matrix-free, row by row, without atomics and in a fixed summation order. The reference factorises
// H + lambda*I exactly (sparse LL^T); here the damped system is solved by preconditioned CG. The path is bound by
// synchronisation latency, not bandwidth, so the kernel is organised around the cost of a sync:
//   - small problems (a tracking frame, the reference's 5-keyframe BA window) run as ONE thread-block cluster of up
//     to 16 CTAs and synchronise with the hardware cluster barrier (0.35 us measured) instead of a global-memory
//     barrier (0.9-1.3 us, profiles/r01_barrier_latency.txt);
//   - in "resident" mode every CTA keeps the Jacobians, the regulariser coefficients, the CG vectors and a dense
//     16-row block-Jacobi preconditioner of its chunk in shared memory for the duration of a solve; the only
//     vector that crosses CTAs (z = M^-1 r) is exchanged through L2;
//   - one CG iteration costs two synchronisations and one batched gather;
//   - grid reductions go through per-CTA slots summed in a fixed order, so every CTA derives bit-identical scalars
//     and takes the same branches; the 6-dof pose blocks are replicated in every CTA's shared memory.
#include <cooperative_groups.h>
#include <float.h>
#include <stdio.h>

#include <algorithm>

#include "nrs_engine.cuh"

namespace nrs {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// system scope: flags of the multi-GPU exchange live in peer-mapped memory
__device__ __forceinline__ unsigned long long atom_acqrel_add_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long old;
  asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct V3 {
  double x, y, z;
};
// data written by other CTAs during the launch: L2-coherent loads
__device__ __forceinline__ V3 ld3(const double* base, int i) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(base + 4 * (size_t)i));
  const double b = __ldcg(base + 4 * (size_t)i + 2);
  return V3{a.x, a.y, b};
}
__device__ __forceinline__ void st3(double* base, int i, const V3& v) {
  *reinterpret_cast<double2*>(base + 4 * (size_t)i) = make_double2(v.x, v.y);
  base[4 * (size_t)i + 2] = v.z;
}
// read-only inputs
__device__ __forceinline__ V3 ld3c(const double* base, int i) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(base + 4 * (size_t)i));
  const double b = __ldg(base + 4 * (size_t)i + 2);
  return V3{a.x, a.y, b};
}
// plain (generic) access: CTA-private data in shared memory or in this CTA's own rows of a global array
__device__ __forceinline__ V3 ld3p(const double* base, int i) {
  const double2 a = *reinterpret_cast<const double2*>(base + 4 * (size_t)i);
  return V3{a.x, a.y, base[4 * (size_t)i + 2]};
}

// One 4-double record (32 bytes, 32-byte aligned) with ONE 256-bit access (sm_100: LDG.256 / STG.256). The wide CG
// loop is bound by L1TEX wavefronts (distinct 128-byte lines per load instruction): half the instructions of the
// double2 + double pair per gathered row.
struct __align__(32) D4 {
  double x, y, z, w;
};
__device__ __forceinline__ D4 ld4(const double* base, size_t i) {
  D4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(base + 4 * i)
               : "memory");
  return r;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st4(double* base, size_t i, double x, double y, double z, double w = 0.0) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(base + 4 * i), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
}

// 3-double packed rows (shared-memory copies of z): a 32-byte row stride touches only half of the banks when lanes
// gather random rows; 24-byte rows spread over all of them
__device__ __forceinline__ V3 ld3s(const double* base, int i) {
  const double* p = base + 3 * (size_t)i;
  return V3{p[0], p[1], p[2]};
}
__device__ __forceinline__ void st3s(double* base, int i, const V3& v) {
  double* p = base + 3 * (size_t)i;
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}

__device__ __forceinline__ int sym6(int a, int c) { return a * 6 - (a * (a - 1)) / 2 + (c - a); }  // a <= c

// Inverse of the SPD 6x6 (upper-packed H + lambda I) by Cholesky; out = full 36. Returns false if not SPD.
// Fully unrolled (everything stays in registers) with one reciprocal per pivot.
__device__ bool invert6(const double* Hu, double lambda, double* out) {
  double L[6][6], id[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double s = Hu[sym6(j, j)] + lambda;
#pragma unroll
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (!(s > 0)) ok = false;
    const double r = rsqrt(s);
    id[j] = r;  // 1 / L_jj
    L[j][j] = s * r;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double t = Hu[sym6(j, i)];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t * r;
    }
  }
  if (!ok) return false;
  // W = L^-1 (lower), then H^-1 = W^T W
  double W[6][6];
#pragma unroll
  for (int c = 0; c < 6; c++) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (i < c) {
        W[i][c] = 0;
      } else {
        double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = c; k < i; k++) s -= L[i][k] * W[k][c];
        W[i][c] = s * id[i];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int c = a; c < 6; c++) {
      double s = 0;
#pragma unroll
      for (int k = c; k < 6; k++) s += W[k][a] * W[k][c];
      out[a * 6 + c] = s;
      out[c * 6 + a] = s;
    }
  return true;
}


#define NRS_SHARED(p) __builtin_assume(__isShared(p))
constexpr int kPB = kPrecBlock;      // rows per dense preconditioner block
constexpr int kPN = 3 * kPrecBlock;  // its dimension
constexpr int kJS = 22;              // shared-memory stride of a Jacobian row (20 doubles + 2: spreads rows over the banks)
constexpr int kPS = kPN + 4;         // padded row stride (floats): 16-byte aligned rows, conflict-free float4 row reads

// Sum of value v over the chunk partials of pose slot k in chunk order; the loads of a batch are issued together
// (a plain `s += load` loop serialises one L2 round trip per chunk).
__device__ __forceinline__ double sum_chunk_partials_of(const Params& P, int k, int v, int par) {
  const int c1 = P.kf_chunk_ptr[k + 1];
  double s = 0;
  for (int c0 = P.kf_chunk_ptr[k]; c0 < c1; c0 += 8) {
    double t[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int c = min(c0 + u, c1 - 1);
      t[u] = __ldcg(P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals + v);
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (c0 + u < c1) s += t[u];
  }
  return s;
}

// The same over the segment partials of the wide CG loop (Engine<true>::pcg_wide).
__device__ __forceinline__ double sum_wseg_range(const Params& P, int cb, int c1, int v, int par) {
  double s = 0;
  for (int c0 = cb; c0 < c1; c0 += 8) {
    double t[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int c = min(c0 + u, c1 - 1);
      t[u] = __ldcg(P.wseg_part + ((size_t)par * P.n_wseg + c) * 8 + v);
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (c0 + u < c1) s += t[u];
  }
  return s;
}
__device__ __forceinline__ double sum_wseg_partials_of(const Params& P, int k, int v, int par) {
  return sum_wseg_range(P, P.kf_wseg_ptr[k], P.kf_wseg_ptr[k + 1], v, par);
}

struct XRet {
  unsigned long long xe;
  size_t xcur;
  unsigned gen;
  int dead;
};
// ---- fused grid reduction + multi-GPU exchange (landmark-sharded BA, cooperative-grid mode): ONE synchronisation
// instead of grid barrier -> exchange by CTA 0 -> grid barrier. Every CTA has written its partials to its slot and
// arrives on the grid counter; the LAST CTA to arrive sums the slots (fixed order), adds this rank's pose partials
// (kind 1: 6 F CG values from the segment / chunk partials, kind 2: 27 F linearisation blocks), pushes the record into
// every rank's buffer and releases this rank's flag on every rank at system scope. ALL CTAs then acquire-poll the
// flags of all ranks in their own buffer — their own rank's flag doubles as the local grid barrier — and combine the
// records in rank order: every CTA of every GPU derives bit-identical values. Halo rows pushed before the call are
// covered: CTA stores -> acq_rel arrive -> last CTA's system fence + flag -> acquire at the readers.
__device__ __forceinline__ XRet xreduce_impl(const Params& P, double* s_scal, int* s_flag, int n, unsigned maxmask,
                                             int kind, int par, int xpar, unsigned long long xe, unsigned gen, int dead) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  xe++;
  gen++;
  const unsigned long long epoch = P.xepoch0 + xe;
  const int W = P.world;
  const size_t half = (size_t)(epoch & 1) * W * P.xstride;
  const size_t rec = half + (size_t)P.rank * P.xstride;
  __syncthreads();  // the slot of this CTA (and everything else it wrote) precedes the arrive
  const unsigned long long ta = global_timer_ns();
  if (tid == 0) s_flag[1] = (atom_acqrel_add_u64(P.bar, 1ULL) + 1 == (unsigned long long)gen * gridDim.x) ? 1 : 0;
  __syncthreads();
  if (s_flag[1]) {
    const unsigned long long tl0 = global_timer_ns();
    const int G = (int)gridDim.x;
    if (tid < 32 * n) {
      const int k = tid >> 5, lane = tid & 31;
      const bool mx = (maxmask >> k) & 1;
      double s = mx ? -DBL_MAX : 0.0;
      for (int c0 = lane; c0 < G; c0 += 8 * 32) {
        double o[8];
#pragma unroll
        for (int u = 0; u < 8; u++) o[u] = __ldcg(P.slots + ((size_t)par * G + min(c0 + 32 * u, G - 1)) * kSlotVals + k);
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (c0 + 32 * u < G) s = mx ? fmax(s, o[u]) : s + o[u];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, s, off);
        s = mx ? fmax(s, o) : s + o;
      }
      if (lane == 0)
        for (int r = 0; r < W; r++) P.xred[r][rec + k] = s;
    }
    const int per = (kind == 1) ? 6 : 27;
    const int nx = (kind == 0) ? 0 : per * P.F;
    const bool segs = kind == 1 && P.wide;
    const int* rptr = segs ? P.kf_wseg_ptr : P.kf_chunk_ptr;
    const double* part = segs ? P.wseg_part + (size_t)xpar * P.n_wseg * 8 : P.chunk_part + (size_t)xpar * P.n_chunks * kChunkVals;
    const int pstride = segs ? 8 : kChunkVals;
    // up to four values per thread and round: first the ranges, then the first eight partials of each value (all in
    // flight together: this CTA is the critical path of every GPU), then the rare longer tails
    for (int t0 = tid; t0 < nx; t0 += 4 * nthr) {
      int cb[4], ce[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int t = t0 + u * nthr;
        cb[u] = t < nx ? rptr[t / per] : 0;
        ce[u] = t < nx ? rptr[t / per + 1] : 0;
      }
      double o[4][8];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v = min(t0 + u * nthr, nx - 1) % per;
#pragma unroll
        for (int q = 0; q < 8; q++)
          o[u][q] = (cb[u] + q < ce[u]) ? __ldcg(part + (size_t)(cb[u] + q) * pstride + v) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int t = t0 + u * nthr;
        if (t >= nx) continue;
        double sum = 0;
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (cb[u] + q < ce[u]) sum += o[u][q];
        for (int c = cb[u] + 8; c < ce[u]; c++) sum += __ldcg(part + (size_t)c * pstride + t % per);
        for (int r = 0; r < W; r++) P.xred[r][rec + 8 + t] = sum;
      }
    }
    __syncthreads();
    const unsigned long long tl1 = global_timer_ns();
    if (tid < W) st_release_sys_u64(P.xflag[tid] + P.rank, epoch);  // release: cumulative over the CTA's stores above
    if (tid == 0) {  // diagnostics (ns): the last CTA's sums + record stores, its fence + flag stores
      atomicAdd(reinterpret_cast<unsigned long long*>(&P.stats->prof[8]), tl1 - tl0);
      atomicAdd(reinterpret_cast<unsigned long long*>(&P.stats->prof[9]), global_timer_ns() - tl1);
    }
  }
  if (tid < W && (!dead || tid == P.rank)) {  // after a time-out only the local barrier is kept
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys_u64(P.xflag[P.rank] + tid) < epoch) {
      if ((++spins & 63u) == 0 && tid != P.rank) {
        if (__ldcg(P.xabort) || global_timer_ns() - t0 > P.xtimeout_ns) {
          *P.xabort = 1;
          break;
        }
      }
    }
    if (blockIdx.x == 0 && tid == P.rank)  // diagnostics (ns): CTA 0 from its arrival to its own rank's flag ...
      atomicAdd(reinterpret_cast<unsigned long long*>(&P.stats->prof[10]), global_timer_ns() - ta);
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0)  // ... and to the flags of all ranks
    atomicAdd(reinterpret_cast<unsigned long long*>(&P.stats->prof[11]), global_timer_ns() - ta);
  if (__ldcg(P.xabort)) dead = 1;
  if (tid < n) {
    const bool mx = (maxmask >> tid) & 1;
    double s = mx ? -DBL_MAX : 0.0;
    double o[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; r++) o[r] = r < W ? __ldcg(P.xred[P.rank] + half + (size_t)r * P.xstride + tid) : 0.0;
#pragma unroll
    for (int r = 0; r < kMaxWorld; r++)
      if (r < W) s = mx ? fmax(s, o[r]) : s + o[r];
    s_scal[tid] = s;
  }
  __syncthreads();
  return XRet{xe, half, gen, dead};
}

// ---- multi-GPU exchange of the landmark-sharded BA (cooperative-grid mode only). Called by every CTA with this
// GPU's totals in s_scal[0..n): CTA 0 pushes them (and, kind 1 / 2, this rank's 6 F CG pose partials / 27 F
// linearisation pose blocks summed over its chunks) into the record [parity][rank] of EVERY rank's reduction buffer,
// signals every rank and waits for every rank's signal; a second grid barrier releases the other CTAs. All ranks
// then combine the records in rank order, so every CTA of every GPU derives bit-identical values and takes the same
// branches. Halo rows pushed before the call are covered by the same signal (CTA stores -> grid barrier ->
// system fence -> flag).
__device__ __noinline__ XRet xexchange_impl(const Params& P, double* s_scal, int n, unsigned maxmask, int kind,
                                            int par, unsigned long long xe, unsigned gen, int dead) {
  // par: bit 0 = parity of the grid-reduction slots, bit 1 = buffer of the pose partials (CG loops)
  if (P.xfused)
    return xreduce_impl(P, s_scal, reinterpret_cast<int*>(s_scal + 28), n, maxmask, kind, par & 1, par >> 1, xe, gen, dead);
  par >>= 1;
  const int tid = threadIdx.x, nthr = blockDim.x;
  xe++;
  const unsigned long long epoch = P.xepoch0 + xe;
  const int W = P.world;
  const size_t half = (size_t)(epoch & 1) * W * P.xstride;
  const size_t rec = half + (size_t)P.rank * P.xstride;
  if (blockIdx.x == 0) {
    for (int t = tid; t < n * W; t += nthr) P.xred[t / n][rec + t % n] = s_scal[t % n];
    const int per = (kind == 1) ? 6 : 27;
    const int nx = (kind == 0) ? 0 : per * P.F;
    for (int t = tid; t < nx; t += nthr) {
      const double s = (kind == 1 && P.wide) ? sum_wseg_partials_of(P, t / per, t % per, par)
                                             : sum_chunk_partials_of(P, t / per, t % per, par);
      for (int r = 0; r < W; r++) P.xred[r][rec + 8 + t] = s;
    }
    __syncthreads();
    if (tid < W) {
      __threadfence_system();
      st_release_sys_u64(P.xflag[tid] + P.rank, epoch);
      if (!dead) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys_u64(P.xflag[P.rank] + tid) < epoch) {
          if (global_timer_ns() - t0 > P.xtimeout_ns) {
            *P.xabort = 1;
            break;
          }
        }
      }
    }
  }
  // grid barrier (release-add / acquire-poll, as Engine::barrier in grid mode)
  gen++;
  __syncthreads();
  if (tid == 0) {
    red_release_add_u64(P.bar, 1ULL);
    const unsigned long long target = (unsigned long long)gen * gridDim.x;
    while (ld_acquire_u64(P.bar) < target) {
    }
  }
  __syncthreads();
  if (__ldcg(P.xabort)) dead = 1;  // uniform over the grid: written before the barrier
  if (tid < n) {
    const bool mx = (maxmask >> tid) & 1;
    double s = mx ? -DBL_MAX : 0.0;
    double o[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; r++) o[r] = r < W ? __ldcg(P.xred[P.rank] + half + (size_t)r * P.xstride + tid) : 0.0;
#pragma unroll
    for (int r = 0; r < kMaxWorld; r++)
      if (r < W) s = mx ? fmax(s, o[r]) : s + o[r];
    s_scal[tid] = s;
  }
  __syncthreads();
  return XRet{xe, half, gen, dead};
}

// Value t of the pose records of the last exchange, summed over the ranks in rank order. Every rank's load is issued
// before the first add (a plain accumulate loop pays one L2 round trip per rank); a real call keeps its eight values
// in flight out of the single-GPU loops' register budget.
__device__ __noinline__ double xextra_of(const Params& P, size_t xcur, int t) {
  const double* base = P.xred[P.rank] + xcur + 8 + t;
  double o[kMaxWorld];
#pragma unroll
  for (int r = 0; r < kMaxWorld; r++) o[r] = r < P.world ? __ldcg(base + (size_t)r * P.xstride) : 0.0;
  double s = 0;
#pragma unroll
  for (int r = 0; r < kMaxWorld; r++)
    if (r < P.world) s += o[r];
  return s;
}

static_assert((kPN / 4) % kTPR == 0, "the lanes of a row split the float4 columns of the block preconditioner evenly");

// WIDE: the variant for large windows (cooperative grid, vectors in global memory): compiled without the
// shared-memory-resident and cluster-native paths so that it fits 128 registers and TWO CTAs share an SM — the big
// BA matvec is a chain of dependent L2 gathers per row, and twice the warps hide twice the latency.
// SHARD: the variant a landmark-sharded rank launches (world > 1). The single-GPU kernels are compiled without any
// of the exchange paths: their mere presence (call sites of the exchange functions inside the CG loops) cost the
// single-GPU BA 8 % through register allocation.
template <bool WIDE, bool SHARD>
struct Engine {
  const Params& P;
  // shared memory
  double *s_pose, *s_pose_bak, *s_H, *s_M, *s_bp, *s_xp, *s_rp, *s_zp, *s_pp, *s_qp, *s_red, *s_scal;
  double *s_jac, *s_x, *s_r, *s_p, *s_q, *s_z, *s_minv, *s_coef;  // resident chunk state
  double* s_pr;             // per-row pose partials of the CG matvec (6 per row)
  double* s_gather;         // [2][16][8] reduction values PUSHED here by every CTA of the cluster (cluster-native CG)
  double* s_rowA;           // per-row pose-block records of the linearisation (16 per row)
  const double** s_zptr;    // per incidence: where the neighbour's z lives (shared memory or L2)
  float* s_rf;              // residual as fp32 for the block preconditioner
  double* s_halo;           // z of the out-of-chunk neighbours, pushed by their owners (3 doubles per halo row)
  // coarse level of the cluster-native loop
  unsigned char *s_cid, *s_hcid;  // aggregate (= chunk) of every incidence's neighbour / of every halo row (255: fixed row)
  float *s_Aall, *s_Ac, *s_rcv;   // published row blocks [16][kRowBlk], coarse matrix -> -inverse [54][kCoarseS], residual
  double* s_y;                    // coarse correction [kCoarseN]
  float* s_binv;  // dense block inverses (fp32, symmetric): a preconditioner need not be exact
  int* s_oth;
  int* s_flag;
  unsigned gen;
  unsigned long long xe;    // multi-GPU: exchanges done by this launch
  size_t xcur;              // record base (parity) of the last exchange inside this rank's reduction buffer
  bool xdead;               // a peer timed out: stop waiting, finish the program, report
  int tid, nthr;
  // LM state (uniform)
  double lambda, ni;
  int lm_iters, lm_trials, pcg_iters, n_sweeps, n_chi2, n_trace, pcg_fail;
  long long prof[16];

  __device__ Engine(const Params& p, double* sm) : P(p) {
    tid = threadIdx.x;
    nthr = blockDim.x;
    const int F = p.F;
    s_pose = sm;            sm += 7 * F;
    s_pose_bak = sm;        sm += 7 * F;
    s_H = sm;               if (!WIDE) sm += 21 * F;
    s_M = sm;               sm += 36 * F;
    s_bp = sm;              sm += 6 * F;
    s_xp = sm;              sm += 6 * F;
    s_rp = sm;              sm += 6 * F;
    s_zp = sm;              sm += 6 * F;
    s_pp = sm;              sm += 6 * F;
    s_qp = sm;              sm += 6 * F;
    s_red = sm;             sm += 32 * kChunkVals;
    s_scal = sm;            sm += 32;
    s_pr = sm;              if (!WIDE) sm += 6 * kMaxRows;
    s_gather = sm;          if (!WIDE) sm += 2 * 16 * kGatherVals;
    s_flag = reinterpret_cast<int*>(sm);  sm += 2;
    sm = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm) + 15) & ~uintptr_t(15));  // double2 accesses below
    // the pose-block records of the linearisation and the Jacobian cache of the CG loop are never live together
    s_rowA = sm;
    // wide variant (engine_smem_bytes_wide): the CG loop keeps its pose partials in registers (no s_pr), there is no
    // cluster (no s_gather) and H_pp — filled after the linearisation's row records were reduced, read until the
    // next linearisation — lives on top of those records: two CTAs per SM fit up to ~110 poses
    if (WIDE) s_H = sm;
    if (WIDE) sm += (21 * F > 16 * kMaxRows) ? 21 * F : 16 * kMaxRows;
    else if (!p.resident) sm += 16 * kMaxRows;
    if (p.resident) {
      const int R = p.res_rows, CI = p.res_inc;
      s_jac = sm;           sm += (kJS * (size_t)R > 16 * (size_t)kMaxRows) ? kJS * (size_t)R : 16 * (size_t)kMaxRows;
      s_x = sm;             sm += 4 * (size_t)R;
      s_r = sm;             sm += 4 * (size_t)R;
      s_p = sm;             sm += 4 * (size_t)R;
      s_q = sm;             sm += 4 * (size_t)R;
      s_z = sm;             sm += (3 * (size_t)R + 1) & ~(size_t)1;
      s_minv = sm;          if (!p.block_prec) sm += 8 * (size_t)R;
      s_coef = sm;          sm += 4 * (size_t)CI;
      s_zptr = reinterpret_cast<const double**>(sm);  sm += CI;
      s_oth = nullptr;
      sm = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm) + 15) & ~uintptr_t(15));
      s_rf = reinterpret_cast<float*>(sm);  sm += (3 * (size_t)R + 1) / 2;
      sm = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm) + 15) & ~uintptr_t(15));
      s_halo = sm;          sm += (3 * (size_t)p.halo_rows + 1) & ~(size_t)1;
      s_binv = reinterpret_cast<float*>(sm);
      s_cid = s_hcid = nullptr;
      s_Aall = s_Ac = s_rcv = nullptr;
      s_y = nullptr;
      if (p.coarse) {  // behind the block inverses
        float* f = s_binv + (size_t)(R / kPB) * kPN * kPS;
        s_Aall = f;         f += 16 * kRowBlk;
        s_Ac = f;           f += kCoarseN * kCoarseS;
        s_rcv = f;          f += kCoarseS;
        s_y = reinterpret_cast<double*>(f);  f += 2 * kCoarseS;
        s_cid = reinterpret_cast<unsigned char*>(f);
        s_hcid = s_cid + ((CI + 15) & ~15);
      }
    } else {
      s_jac = s_x = s_r = s_p = s_q = s_z = s_minv = s_coef = nullptr;
      s_zptr = nullptr;
      s_oth = nullptr;
      s_rf = nullptr;
      s_halo = nullptr;
      s_binv = nullptr;
      s_cid = s_hcid = nullptr;
      s_Aall = s_Ac = s_rcv = nullptr;
      s_y = nullptr;
    }
    gen = 0;
    xe = 0;
    xcur = 0;
    xdead = false;
    lambda = -1;
    ni = 2;
    lm_iters = lm_trials = pcg_iters = n_sweeps = n_chi2 = n_trace = pcg_fail = 0;
    for (int i = 0; i < 16; i++) prof[i] = 0;
  }

  // ---- grid-wide barrier. Cluster mode: hardware barrier with release/acquire semantics at cluster scope.
  // Grid mode: release-add on one counter, acquire-poll until every CTA of this generation arrived; bar.sync on both
  // sides extends the ordering to every thread of the CTA (no separate fences: profiles/r01_barrier_latency.txt).
  __device__ __forceinline__ void barrier() {
    gen++;
    if (P.cluster_mode) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      const long long tb0 = clock64();
      __syncthreads();
      if (tid == 0) {
        red_release_add_u64(P.bar, 1ULL);
        const unsigned long long target = (unsigned long long)gen * gridDim.x;
        while (ld_acquire_u64(P.bar) < target) {
        }
      }
      __syncthreads();
      prof[12] += clock64() - tb0;  // grid barriers (arrival skew + latency)
    }
  }

  // ---- dot product of two replicated pose vectors in shared memory, identical in every thread (and every CTA and
  // rank: fixed order). Called by all threads after the vectors are complete. Long vectors (BA windows of 30-100
  // keyframes) are summed by warp 0 — the serial loop cost 6 F dependent DFMAs per thread, twice per CG iteration.
  __device__ __forceinline__ double pose_dot(const double* a, const double* b, int n) {
    if (n <= 32) {
      double s = 0;
      for (int t = 0; t < n; t++) s += a[t] * b[t];
      return s;
    }
    if (tid < 32) {
      double s = 0;
      for (int t = tid; t < n; t += 32) s += a[t] * b[t];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (tid == 0) s_scal[30] = s;
    }
    __syncthreads();
    return s_scal[30];
  }

  // ---- block reduction of NV per-thread values (fixed order); result in dst[0..NV)
  template <int NV>
  __device__ __forceinline__ void block_reduce(double (&v)[NV], double* dst) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    }
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) s_red[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (tid < NV) {
      double s = 0;
      for (int w = 0; w < nw; w++) s += s_red[w * NV + tid];
      dst[tid] = s;
    }
    __syncthreads();
  }

  // ---- grid reduction of n (<= 4) values: v[k] are per-thread partials. maxmask bit k: max instead of sum.
  // Result (identical in every CTA) lands in s_scal[0..n). Includes one grid barrier.
  template <int N>
  __device__ __forceinline__ void grid_reduce(double (&v)[N], unsigned maxmask, int xkind = 0, int xpar = -1) {
    const int par = gen & 1;
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v[k], off);
        v[k] = ((maxmask >> k) & 1) ? fmax(v[k], o) : v[k] + o;
      }
    }
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < N; k++) s_red[warp * N + k] = v[k];
    }
    __syncthreads();
    if (tid < N) {
      double s = s_red[tid];
      for (int w = 1; w < nw; w++) s = ((maxmask >> tid) & 1) ? fmax(s, s_red[w * N + tid]) : s + s_red[w * N + tid];
      P.slots[((size_t)par * gridDim.x + blockIdx.x) * kSlotVals + tid] = s;
    }
    // sharded runs with the fused exchange synchronise the grid and the ranks in ONE step inside xexchange()
    const bool fused = (SHARD && P.world > 1) && P.xfused;
    if (!fused) barrier();
    if (!fused && tid < 32 * N) {
      const int k = tid >> 5;
      const bool mx = (maxmask >> k) & 1;
      double s = mx ? -DBL_MAX : 0.0;
      // batches of 8 slots per lane: every load of a batch is issued before the first use (a plain accumulate loop
      // pays one L2 round trip per 32 CTAs)
      for (int c0 = lane; c0 < (int)gridDim.x; c0 += 8 * 32) {
        double o[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int c = min(c0 + 32 * u, (int)gridDim.x - 1);
          o[u] = __ldcg(P.slots + ((size_t)par * gridDim.x + c) * kSlotVals + k);
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (c0 + 32 * u < (int)gridDim.x) s = mx ? fmax(s, o[u]) : s + o[u];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, s, off);
        s = mx ? fmax(s, o) : s + o;
      }
      if (lane == 0) s_scal[k] = s;
    }
    __syncthreads();
    if ((SHARD && P.world > 1)) xexchange(N, maxmask, xkind, par + 2 * (xpar >= 0 ? xpar : par));
  }

  // ---- multi-GPU (landmark-sharded BA): see xexchange_impl. The exchange is a free function that gets the engine
  // state it needs by value — a non-inlined MEMBER would make `this` escape and push every member of the engine
  // (shared-memory pointers, lambda, ...) from registers into local memory, also in the single-GPU hot loops.
  __device__ __forceinline__ void xexchange(int n, unsigned maxmask, int kind, int par) {
    const XRet r = xexchange_impl(P, s_scal, n, maxmask, kind, par, xe, gen, xdead ? 1 : 0);
    xe = r.xe;
    xcur = r.xcur;
    gen = r.gen;
    xdead = r.dead != 0;
  }
  // Value t of the pose records of the last exchange, summed over the ranks in rank order.
  __device__ __forceinline__ double xextra(int t) { return xextra_of(P, xcur, t); }
  // Refresh the halo copies of owned row i on the ranks that read it.
  __device__ __forceinline__ void xpush3(double* const* base, int i, const V3& v) {
    for (int a = P.xp_ptr[i]; a < P.xp_ptr[i + 1]; a++) {
      const int d = P.xp_dst[a];
      st3(base[d >> 26], d & 0x3ffffff, v);
    }
  }
  // Grid barrier that also waits for the halo rows pushed by the other ranks.
  __device__ __forceinline__ void xsync() {
    double d[1] = {0};
    grid_reduce<1>(d, 0);
  }

  // ================================================================================================
  // Edge-parallel pass over pair and damper edges: chi2 (and, when LIN, the linearised coefficients).
  //   spatial  : SpatialRegularizerWithDeformation   optimization/spatial_regularizer_with_deformation.cc:36-49
  //   spring   : PositionRegularizerWithDeformation  optimization/position_regularizer_with_deformation.cc:31-57
  //              PositionRegularizer (quirk E1)      optimization/position_regularizer.cc:32-61
  //   damper   : SpatialRegularizer                  optimization/spatial_regularizer.cc:32-59
  // ================================================================================================
  template <bool LIN>
  __device__ void edges_pass(double& chi) {
    const int gsz = gridDim.x * nthr;
    for (int e = blockIdx.x * nthr + tid; e < P.P; e += gsz) {
      const int i = P.pair_i[e], j = P.pair_j[e];
      const V3 xi = ld3(P.x, i), xj = ld3(P.x, j);
      double s = 0, u0 = 0, u1 = 0, u2 = 0, c = 0;
      const double w = P.pair_w[e];
      // an edge whose vertices are all fixed is not part of the active set (sparse_optimizer.cpp:232-246)
      const bool live = !(P.pt_fixed && P.pt_fixed[i] && P.pt_fixed[j]);
      const bool cnt = !P.pair_cnt || P.pair_cnt[e];  // sharded BA: an edge shared with another rank is counted once
      if (live && w >= 0 && P.sp_level[e] == 0) {
        const double e0 = w * (xi.x - xj.x), e1 = w * (xi.y - xj.y), e2 = w * (xi.z - xj.z);
        const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
        double rho, drho;
        huber(c2, P.delta_spatial, rho, drho);
        if (cnt) chi += rho;
        s = drho * P.info_spatial * w * w;
      }
      if (live && P.spring_kind != SPRING_NONE) {
        const V3 ri = ld3c(P.rest, i), rj = ld3c(P.rest, j);
        const double c1x = ri.x + xi.x, c1y = ri.y + xi.y, c1z = ri.z + xi.z;
        const double c2x = rj.x + xj.x, c2y = rj.y + xj.y, c2z = rj.z + xj.z;
        const double dx = c1x - c2x, dy = c1y - c2y, dz = c1z - c2z;
        const double dist = sqrt(dx * dx + dy * dy + dz * dz);
        const double d0 = P.pair_d0[e];
        const double err = P.spring_k * (dist - d0) / d0;
        const double ch = err * err * P.info_spring;
        double rho, drho;
        huber(ch, P.delta_spring, rho, drho);
        if (cnt) chi += rho;
        if (LIN) {
          double j0, j1, j2;
          if (P.spring_kind == SPRING_DEFORM) {
            const double aa = P.spring_k / (2 * d0 * dist);
            j0 = aa * (2 * c1x - 2 * c2x);
            j1 = aa * (2 * c1y - 2 * c2y);
            j2 = aa * (2 * c1z - 2 * c2z);
          } else {
            const double kd = P.spring_k / d0, dcs = 1.0 / sqrt(dist);
            j0 = kd * dcs * (2.0 * dx);
            j1 = kd * dcs * (2.0 * dy);
            j2 = kd * dcs * (2.0 * dz);
          }
          const double sw = sqrt(drho * P.info_spring);
          u0 = sw * j0;
          u1 = sw * j1;
          u2 = sw * j2;
          c = sw * err;
        }
      }
      if (LIN) {
        double2* o = reinterpret_cast<double2*>(P.pc + 4 * (size_t)e);
        o[0] = make_double2(s, u0);
        o[1] = make_double2(u1, u2);
        P.pcc[e] = c;
      }
    }
    for (int e = blockIdx.x * nthr + tid; e < P.D; e += gsz) {
      const int4 v = *reinterpret_cast<const int4*>(P.dmp_v + 4 * (size_t)e);
      const V3 a = ld3(P.x, v.x), b = ld3(P.x, v.y), an = ld3(P.x, v.z), bn = ld3(P.x, v.w);
      const double w = P.dmp_w[e];
      const double e0 = w * ((an.x - a.x) - (bn.x - b.x));
      const double e1 = w * ((an.y - a.y) - (bn.y - b.y));
      const double e2 = w * ((an.z - a.z) - (bn.z - b.z));
      const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
      double rho, drho;
      huber(c2, P.delta_spatial, rho, drho);
      if (!P.dmp_cnt || P.dmp_cnt[e]) chi += rho;
      if (LIN) {
        const double g = drho * P.info_spatial * w;
        double2* o = reinterpret_cast<double2*>(P.dc + 4 * (size_t)e);
        o[0] = make_double2(g * w, g * e0);
        o[1] = make_double2(g * e1, g * e2);
      }
    }
  }

  // Fixed-order reduction of the chunk's pose-block records (s_rowA: A 2x6, omega, weighted error) into the 21
  // upper-triangular entries of H_pp and the 6 of b_p: 27 x 8 threads sum strided row groups, 27 threads finish.
  __device__ __forceinline__ void reduce_pose_block(int nrows, double* dst) {
    __syncthreads();
    for (int tt = tid; tt < 27 * 8; tt += nthr) {
      const int v = tt % 27, g = tt / 27;
      int a = 0, c = 0;
      if (v < 21) {
        int t = v;
        while (t >= 6 - a) {
          t -= 6 - a;
          a++;
        }
        c = a + t;
      } else {
        a = v - 21;
      }
      double s = 0;
      for (int r = g; r < nrows; r += 8) {
        const double* rec = s_rowA + 16 * (size_t)r;
        if (v < 21)
          s += rec[12] * (rec[a] * rec[c] + rec[6 + a] * rec[6 + c]);
        else
          s += rec[a] * rec[14] + rec[6 + a] * rec[15];
      }
      s_red[tt] = s;
    }
    __syncthreads();
    if (tid < 27) {
      double s = 0;
#pragma unroll
      for (int g = 0; g < 8; g++) s += s_red[27 * g + tid];
      dst[tid] = s;
    }
    __syncthreads();
  }

  // Reprojection error of point row i at the current estimate; returns chi2 (info * |e|^2).
  //   ReprojectionErrorWithDeformation::computeError  optimization/reprojection_error_with_deformation.cc:37-50
  //   ReprojectionError::computeError                 optimization/reprojection_error.cc:32-44
  //   ReprojectionErrorOnlyPose::computeError         optimization/reprojection_error_only_pose.cc:50-58
  __device__ __forceinline__ double reproj_error(int i, int kf, const V3& xi, double pc[3], double err[2]) {
    const V3 r = ld3c(P.rest, i);
    const double Xw[3] = {xi.x + r.x, xi.y + r.y, xi.z + r.z};
    pose_map(s_pose + 7 * kf, Xw, pc);
    float u, v;
    project_f(P.cam, (float)pc[0], (float)pc[1], (float)pc[2], u, v);
    const double2 z = __ldg(reinterpret_cast<const double2*>(P.uv) + i);
    err[0] = z.x - (double)u;
    err[1] = z.y - (double)v;
    return (err[0] * err[0] + err[1] * err[1]) * P.info_reproj;
  }

  // ================================================================================================
  // Row pass: reprojection edge of each point + gather of the incident regulariser coefficients.
  // LIN: stores Jacobians, diagonal blocks, gradient; reduces the pose blocks per chunk. Always: chi2.
  // ================================================================================================
  template <bool LIN>
  __device__ void rows_pass(double& chi, double& maxd, int par) {
    for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
      const int i = P.chunk_begin[c] + tid;
      const bool valid = i < P.chunk_end[c];
      bool pose_rec = false;  // this row stored a pose-block record (A, omega, weighted error)
      if (valid) {
        const V3 xi = ld3(P.x, i);
        const int kf = P.pt_kf[i];
        double D[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        double A[12], B[6], omega = 0;
        const bool var = !P.points_fixed && !(P.pt_fixed && P.pt_fixed[i]);  // this point is an unknown
        if (kf >= 0 && P.rp_level[i] == 0 && (var || !P.poses_fixed)) {
          double pc[3], err[2];
          const double c2 = reproj_error(i, kf, xi, pc, err);
          double rho, drho;
          huber(c2, P.delta_reproj, rho, drho);
          chi += rho;
          P.rp_chi2[i] = c2;
          if (LIN) {
            // linearizeOplus: J_pose = -J_pi * [ -[p]x | I ],  J_point = -J_pi * R
            //   optimization/reprojection_error_with_deformation.cc:52-68, reprojection_error.cc:46-64,
            //   reprojection_error_only_pose.cc:60-76
            float Jf[6];
            projection_jacobian_f(P.cam, (float)pc[0], (float)pc[1], (float)pc[2], Jf);
            double Jp[6];
#pragma unroll
            for (int k = 0; k < 6; k++) Jp[k] = -(double)Jf[k];
            const double x = pc[0], y = pc[1], z = pc[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
              const double a = Jp[r * 3], bb = Jp[r * 3 + 1], cc = Jp[r * 3 + 2];
              A[r * 6 + 0] = bb * (-z) + cc * y;
              A[r * 6 + 1] = a * z + cc * (-x);
              A[r * 6 + 2] = a * (-y) + bb * x;
              A[r * 6 + 3] = a;
              A[r * 6 + 4] = bb;
              A[r * 6 + 5] = cc;
            }
            omega = drho * P.info_reproj;
            const double we0 = -omega * err[0], we1 = -omega * err[1];
            if (!P.poses_fixed) {
              // pose block H_pp += omega A^T A, b_p += A^T we: the row record goes to shared memory and is reduced
              // over the chunk in a fixed order below (keeps 27 accumulators out of the registers)
              double2* rec = reinterpret_cast<double2*>(s_rowA + 16 * (size_t)tid);
#pragma unroll
              for (int k = 0; k < 6; k++) rec[k] = make_double2(A[2 * k], A[2 * k + 1]);
              rec[6] = make_double2(omega, 0.0);
              rec[7] = make_double2(we0, we1);
              pose_rec = true;
            }
            if (var) {
              double R[9];
              quat_to_R(s_pose + 7 * kf, R);
#pragma unroll
              for (int r = 0; r < 2; r++)
#pragma unroll
                for (int cc = 0; cc < 3; cc++)
                  B[r * 3 + cc] = Jp[r * 3] * R[cc] + Jp[r * 3 + 1] * R[3 + cc] + Jp[r * 3 + 2] * R[6 + cc];
              D[0] = omega * (B[0] * B[0] + B[3] * B[3]);
              D[1] = omega * (B[0] * B[1] + B[3] * B[4]);
              D[2] = omega * (B[0] * B[2] + B[3] * B[5]);
              D[3] = omega * (B[1] * B[1] + B[4] * B[4]);
              D[4] = omega * (B[1] * B[2] + B[4] * B[5]);
              D[5] = omega * (B[2] * B[2] + B[5] * B[5]);
              b[0] = B[0] * we0 + B[3] * we1;
              b[1] = B[1] * we0 + B[4] * we1;
              b[2] = B[2] * we0 + B[5] * we1;
            } else {
#pragma unroll
              for (int k = 0; k < 6; k++) B[k] = 0;
            }
          }
        }
        if (LIN) {
          double* jo = P.jac + 20 * (size_t)i;
          if (omega == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) A[k] = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) B[k] = 0;
          }
          if constexpr (WIDE) {
            // component-major (10 double2 planes of V rows): the wide CG loop reads it one thread per row, coalesced
            double2* jt = reinterpret_cast<double2*>(P.jac) + i;
            const size_t V = (size_t)P.V;
#pragma unroll
            for (int k = 0; k < 6; k++) jt[k * V] = make_double2(A[2 * k], A[2 * k + 1]);
#pragma unroll
            for (int k = 0; k < 3; k++) jt[(6 + k) * V] = make_double2(B[2 * k], B[2 * k + 1]);
            jt[9 * V] = make_double2(omega, 0.0);
          } else {
#pragma unroll
          for (int k = 0; k < 6; k++) reinterpret_cast<double2*>(jo)[k] = make_double2(A[2 * k], A[2 * k + 1]);
#pragma unroll
          for (int k = 0; k < 3; k++) reinterpret_cast<double2*>(jo)[6 + k] = make_double2(B[2 * k], B[2 * k + 1]);
          reinterpret_cast<double2*>(jo)[9] = make_double2(omega, 0.0);
          }
        }
        if (!var && !P.points_fixed && LIN) {
          double2* d = reinterpret_cast<double2*>(P.dg + 8 * (size_t)i);
          d[0] = d[1] = d[2] = d[3] = make_double2(0.0, 0.0);
          st3(P.bvec, i, V3{0, 0, 0});
        }
        if (var) {
          double su = 0;  // unary diagonal weight
          if (P.unary_on) {
            // SpatialRegularizerFixed  optimization/spatial_regularizer_fixed.cc:32-43 — the reference value is
            // read live from another vertex and carries no Jacobian
            for (int a = P.un_ptr[i]; a < P.un_ptr[i + 1]; a++) {
              const double w = P.un_w[a];
              const V3 rf = ld3(P.x, P.un_ref[a]);
              const double d0 = xi.x - rf.x, d1 = xi.y - rf.y, d2 = xi.z - rf.z;
              const double c2 = w * w * (d0 * d0 + d1 * d1 + d2 * d2) * P.info_spatial;
              double rho, drho;
              huber(c2, P.delta_spatial, rho, drho);
              chi += rho;
              if (LIN) {
                const double s = drho * P.info_spatial * w * w;
                su += s;
                b[0] -= s * d0;
                b[1] -= s * d1;
                b[2] -= s * d2;
              }
            }
          }
          if (LIN) {
            const int a1 = P.inc_ptr[i + 1];
            for (int a0 = P.inc_ptr[i]; a0 < a1; a0 += 4) {
              // batch of up to 4 incidences: issue every load before the first use
              int other[4], ent[4];
              double2 c0[4], c1[4];
              double cc[4];
              V3 xo[4];
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int a = min(a0 + k, a1 - 1);
                other[k] = P.inc_other[a];
                ent[k] = P.inc_ent[a];
              }
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const double2* cf = reinterpret_cast<const double2*>(P.pc + 4 * (size_t)(ent[k] >> 1));
                c0[k] = __ldcg(cf);
                c1[k] = __ldcg(cf + 1);
                cc[k] = __ldcg(P.pcc + (ent[k] >> 1));
                xo[k] = ld3(P.x, other[k]);
              }
#pragma unroll
              for (int k = 0; k < 4; k++) {
                if (a0 + k < a1) {
                  const double s = c0[k].x, u0 = c0[k].y, u1 = c1[k].x, u2 = c1[k].y;
                  D[0] += s + u0 * u0;
                  D[1] += u0 * u1;
                  D[2] += u0 * u2;
                  D[3] += s + u1 * u1;
                  D[4] += u1 * u2;
                  D[5] += s + u2 * u2;
                  const double sg = (ent[k] & 1) ? -cc[k] : cc[k];
                  b[0] -= s * (xi.x - xo[k].x) + sg * u0;
                  b[1] -= s * (xi.y - xo[k].y) + sg * u1;
                  b[2] -= s * (xi.z - xo[k].z) + sg * u2;
                }
              }
            }
            if (P.D > 0) {
              for (int a = P.dinc_ptr[i]; a < P.dinc_ptr[i + 1]; a++) {
                const int ent = P.dinc_ent[a];
                const double2* cf = reinterpret_cast<const double2*>(P.dc + 4 * (size_t)(ent >> 2));
                const double2 c0 = __ldcg(cf), c1 = __ldcg(cf + 1);
                const int role = ent & 3;
                const double sg = (role == 1 || role == 2) ? 1.0 : -1.0;  // J = (-w, +w, +w, -w) I
                D[0] += c0.x;
                D[3] += c0.x;
                D[5] += c0.x;
                b[0] -= sg * c0.y;
                b[1] -= sg * c1.x;
                b[2] -= sg * c1.y;
              }
            }
            D[0] += su;
            D[3] += su;
            D[5] += su;
            double2* d = reinterpret_cast<double2*>(P.dg + 8 * (size_t)i);
            d[0] = make_double2(D[0], D[1]);
            d[1] = make_double2(D[2], D[3]);
            d[2] = make_double2(D[4], D[5]);
            d[3] = make_double2(su, 0.0);
            st3(P.bvec, i, V3{b[0], b[1], b[2]});
            maxd = fmax(maxd, fmax(fabs(D[0]), fmax(fabs(D[3]), fabs(D[5]))));
          }
        }
      }
      if (LIN && !P.poses_fixed) {
        if (tid < kMaxRows && !pose_rec) {
          double2* rec = reinterpret_cast<double2*>(s_rowA + 16 * (size_t)tid);
#pragma unroll
          for (int k = 0; k < 8; k++) rec[k] = make_double2(0.0, 0.0);
        }
        reduce_pose_block(P.chunk_end[c] - P.chunk_begin[c],
                          P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals);
      }
    }
  }

  __device__ __forceinline__ double sum_chunk_partials(int k, int v, int par) {
    return sum_chunk_partials_of(P, k, v, par);
  }

  // own-row iteration helper
  template <typename Fn>
  __device__ __forceinline__ void for_rows(Fn fn) {
    for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
      const int i = P.chunk_begin[c] + tid;
      if (i < P.chunk_end[c]) fn(i);
    }
  }

  // ================================================================================================
  // Dense block-Jacobi preconditioner (resident mode): blocks of kPB consecutive rows of the chunk. The block of
  // H + lambda I (3x3 diagonal blocks + the pair couplings inside the block) is assembled in fp32 and inverted in
  // place with the symmetric sweep operator; it is applied as a dense 48x48 product in the CG loop.
  // ================================================================================================
  __device__ void build_block_prec(int rb, int re, int ab) {
    NRS_SHARED(s_binv);
    NRS_SHARED(s_coef);
    const int nrows = re - rb;
    const int nblk = (nrows + kPB - 1) / kPB;
    const int total = nblk * kPN * kPS;
    for (int t = tid; t < total; t += nthr) s_binv[t] = 0.f;
    __syncthreads();
    // diagonal 3x3 blocks (+ lambda); rows without unknowns and the padding of the last block get the identity
    for (int lr = tid; lr < nblk * kPB; lr += nthr) {
      const int i = rb + lr;
      float* M = s_binv + (size_t)(lr / kPB) * kPN * kPS;
      const int o = 3 * (lr % kPB);
      const bool live = lr < nrows && !(P.pt_fixed && P.pt_fixed[i]);
      if (live) {
        const double2* d = reinterpret_cast<const double2*>(P.dg + 8 * (size_t)i);
        const double2 d0 = d[0], d1 = d[1], d2 = d[2];
        M[(o + 0) * kPS + o + 0] = (float)(d0.x + lambda);
        M[(o + 0) * kPS + o + 1] = M[(o + 1) * kPS + o + 0] = (float)d0.y;
        M[(o + 0) * kPS + o + 2] = M[(o + 2) * kPS + o + 0] = (float)d1.x;
        M[(o + 1) * kPS + o + 1] = (float)(d1.y + lambda);
        M[(o + 1) * kPS + o + 2] = M[(o + 2) * kPS + o + 1] = (float)d2.x;
        M[(o + 2) * kPS + o + 2] = (float)(d2.y + lambda);
      } else {
        M[(o + 0) * kPS + o + 0] = M[(o + 1) * kPS + o + 1] = M[(o + 2) * kPS + o + 2] = 1.f;
      }
    }
    __syncthreads();
    // pair couplings inside a block: -(s I + u u^T), written once per incidence (both incidences of a pair write
    // the two mirrored 3x3 blocks, each from its own side)
    for (int lr = tid; lr < nrows; lr += nthr) {
      const int i = rb + lr;
      if (P.pt_fixed && P.pt_fixed[i]) continue;
      float* M = s_binv + (size_t)(lr / kPB) * kPN * kPS;
      const int oi = 3 * (lr % kPB);
      for (int a = P.inc_ptr[i]; a < P.inc_ptr[i + 1]; a++) {
        const int other = P.inc_other[a];
        const int lo = other - rb;
        if (lo < 0 || lo >= nrows || lo / kPB != lr / kPB) continue;
        if (P.pt_fixed && P.pt_fixed[other]) continue;
        const double* cf = s_coef + 4 * (size_t)(a - ab);
        const float s = (float)cf[0], u0 = (float)cf[1], u1 = (float)cf[2], u2 = (float)cf[3];
        const int oj = 3 * (lo % kPB);
        const float u[3] = {u0, u1, u2};
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) M[(oi + r) * kPS + oj + c] = -((r == c ? s : 0.f) + u[r] * u[c]);
      }
    }
    __syncthreads();
    // symmetric sweep of every pivot: A -> -A^-1. One warp per block, lanes own columns.
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    for (int bk = warp; bk < nblk; bk += nw) {
      float* M = s_binv + (size_t)bk * kPN * kPS;
      for (int j = 0; j < kPN; j++) {
        const float d = M[j * kPS + j];
        const float id = 1.f / d;
        // row j (pivot row) values for the columns of this lane
        const int k0 = lane, k1 = lane + 32;
        const bool has1 = k1 < kPN;
        const float pj0 = (k0 == j) ? 0.f : M[j * kPS + k0];        // column j itself is rewritten below
        const float pj1 = (has1 && k1 != j) ? M[j * kPS + k1] : 0.f;
        __syncwarp();
        // eliminate in groups of 8 rows: all loads of a group first (independent chains for the scheduler); the
        // pivot row takes f = 0 and is rewritten afterwards
#pragma unroll 1
        for (int i0 = 0; i0 < kPN; i0 += 8) {
          float f[8], a0[8], a1[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int i = i0 + u;
            f[u] = (i == j) ? 0.f : M[i * kPS + j] * id;  // broadcast read
            a0[u] = M[i * kPS + k0];
            a1[u] = has1 ? M[i * kPS + k1] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int i = i0 + u;
            if (k0 != j) M[i * kPS + k0] = fmaf(-f[u], pj0, a0[u]);
            if (has1 && k1 != j) M[i * kPS + k1] = fmaf(-f[u], pj1, a1[u]);
          }
        }
        __syncwarp();
        // pivot column and row: a_ij / d ; pivot: -1/d
        for (int i = lane; i < kPN; i += 32) {
          if (i != j) {
            const float v = M[i * kPS + j] * id;
            M[i * kPS + j] = v;
            M[j * kPS + i] = v;
          }
        }
        if (lane == 0) M[j * kPS + j] = -id;
        __syncwarp();
      }
      // the two triangles differ by rounding ((a/d) b vs (b/d) a): mirror the lower one so M is exactly symmetric
      for (int t = lane; t < kPN * kPN; t += 32) {
        const int i = t / kPN, k = t % kPN;
        if (k > i) M[i * kPS + k] = M[k * kPS + i];
      }
      __syncwarp();
    }
    __syncthreads();
  }

  // Reduce the per-row pose partials (6 doubles per row, written to s_pr by lane 0 of every quad) of a chunk in a
  // fixed order: 48 threads sum strided row groups, 6 threads finish. dst: chunk partial (global).
  __device__ __forceinline__ void reduce_pose_partials(int nrows, double* dst) {
    __syncthreads();
    for (int tt = tid; tt < 48; tt += nthr) {
      const int a = tt % 6, g = tt / 6;
      double s = 0;
      for (int r = g; r < nrows; r += 8) s += s_pr[6 * r + a];
      s_red[tt] = s;
    }
    __syncthreads();
    if (tid < 6) {
      double s = 0;
#pragma unroll
      for (int g = 0; g < 8; g++) s += s_red[6 * g + tid];
      dst[tid] = s;
    }
    __syncthreads();
  }

  // z = Minv_block r for the resident chunk: lane l < 3 of the row's quad produces component l as a 48-term fp32 dot
  // product (r was stored as fp32 in s_rf by the caller, followed by __syncthreads). A preconditioner only has to be
  // a fixed SPD operator, so fp32 is enough; everything that defines the solution (r, x, A) stays fp64.
  // Publishes z to shared and global memory; returns this thread's r_l z_l.
  __device__ __forceinline__ double prec_apply_quads(int rb, int re, bool publish = true) {
    NRS_SHARED(s_binv);
    NRS_SHARED(s_rf);
    NRS_SHARED(s_z);
    NRS_SHARED(s_r);
    const int lr = tid / kTPR, l = tid % kTPR;
    const bool act = lr < re - rb && !(P.pt_fixed && P.pt_fixed[rb + lr]);
    // the lanes of a row split the 48 columns; every lane accumulates its slice of all three components (the residual
    // slice is loaded once for the three rows of M), then the lane group adds up by shuffle
    float acc[3] = {0.f, 0.f, 0.f};
    if (act) {
      const int bk = lr / kPB;
      constexpr int kQ = kPN / 4 / kTPR;  // float4 columns per lane
      const float4* rf = reinterpret_cast<const float4*>(s_rf + (size_t)bk * kPN) + l * kQ;
      const float4* M0 =
          reinterpret_cast<const float4*>(s_binv + (size_t)bk * kPN * kPS + (size_t)(3 * (lr % kPB)) * kPS) + l * kQ;
      float4 rv[kQ];
#pragma unroll
      for (int j = 0; j < kQ; j++) rv[j] = rf[j];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4* M = M0 + (size_t)c * (kPS / 4);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < kQ; j++) {
          const float4 m = M[j];
          s0 = fmaf(m.x, rv[j].x, s0);
          s1 = fmaf(m.y, rv[j].y, s1);
          s0 = fmaf(m.z, rv[j].z, s0);
          s1 = fmaf(m.w, rv[j].w, s1);
        }
        acc[c] = s0 + s1;
      }
    }
#pragma unroll
    for (int o = 1; o < kTPR; o <<= 1) {
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
      acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o);
      acc[2] += __shfl_xor_sync(0xffffffffu, acc[2], o);
    }
    double part = 0;
    if (act) {
      for (int c = l; c < 3; c += kTPR) {
        const double z = -(double)acc[c];  // the sweep leaves -A^-1
        s_z[3 * (size_t)lr + c] = z;
        if (publish) P.zvec[4 * (size_t)(rb + lr) + c] = z;
        part += s_r[4 * (size_t)lr + c] * z;
      }
    }
    return part;
  }

  // ================================================================================================
  // Preconditioned CG on (H + lambda I) delta = b.  Result: xcg rows (global), s_xp poses.
  // The matvec is applied to z = M^-1 r and p, q = A p follow by recurrence (q = A z + beta q), so one iteration is
  //   [gather z of the neighbours, w = A z, p, q, partial p.q] sync [alpha; x, r, z = M^-1 r, partial r.z] sync.
  // Thread organisation inside the loop: a quad of kTPR = 4 consecutive lanes per point row. Phase 1: the quad
  // splits the row's regulariser incidences, reduces the three sums with two shuffle steps, lane 0 adds the
  // reprojection part and runs the recurrences. Phase 2: lanes 0..2 own one component each. The loop is bound by
  // dependent-issue latency, not by bandwidth (profiles/r01_*), hence 4 threads per row.
  // Returns false on breakdown (treated like g2o's failed linear solve,
  // optimization_algorithm_levenberg.cpp:102-121).
  // ================================================================================================
  __device__ bool pcg() {
    const int F = P.F;
    const bool pts = !P.points_fixed, pos = !P.poses_fixed;
    const bool res = !WIDE && P.resident != 0;
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    if (pos) {
      for (int k = tid; k < F; k += nthr)
        if (!invert6(s_H + 21 * k, lambda, s_M + 36 * k)) *s_flag = 1;
    }
    __syncthreads();
    if (*s_flag) return false;  // uniform: every CTA inverts the same blocks
    if (!pts) {
      // pose-only system: F independent 6x6 solves (LinearSolverDense, solvers/dense/linear_solver_dense.h:56-104)
      for (int t = tid; t < 6 * F; t += nthr) {
        const int k = t / 6, a = t % 6;
        double s = 0;
        for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_bp[6 * k + c];
        s_xp[t] = s;
      }
      __syncthreads();
      return true;
    }
    // ---- chunk-resident state
    const int c0 = blockIdx.x;  // resident: exactly one chunk per CTA (or none)
    const bool has = res && c0 < P.n_chunks;
    const int rb = has ? P.chunk_begin[c0] : 0, re = has ? P.chunk_end[c0] : 0;
    const int ab = has ? P.inc_ptr[rb] : 0, ae = has ? P.inc_ptr[re] : 0;
    const bool bprec = res && P.block_prec;
    if (has) {
      for (int t = tid; t < 10 * (re - rb); t += nthr)
        reinterpret_cast<double2*>(s_jac + kJS * (size_t)(t / 10))[t % 10] =
            reinterpret_cast<const double2*>(P.jac + 20 * (size_t)rb)[t];
      for (int a = ab + tid; a < ae; a += nthr) {
        const int ent = P.inc_ent[a];
        const double2* cf = reinterpret_cast<const double2*>(P.pc + 4 * (size_t)(ent >> 1));
        reinterpret_cast<double2*>(s_coef)[2 * (a - ab)] = cf[0];
        reinterpret_cast<double2*>(s_coef)[2 * (a - ab) + 1] = cf[1];
        const int other = P.inc_other[a];
        // where the neighbour's z lives: this chunk's shared memory, or L2
        s_zptr[a - ab] = (other >= rb && other < re) ? s_z + 3 * (size_t)(other - rb) : P.zvec + 4 * (size_t)other;
      }
      __syncthreads();
      if (bprec) build_block_prec(rb, re, ab);
    }
    // vectors: shared memory (indexed from the chunk's first row) or global
    double* X = res ? s_x : P.xcg;
    double* R = res ? s_r : P.rvec;
    double* Pv = res ? s_p : P.pvec;
    double* Q = res ? s_q : P.qvec;
    double* MI = res ? s_minv : P.minv;
    const int vb = res ? rb : 0;

    if (bprec && has) {  // rows of the last block past the chunk end: zero residual
      const int padded = ((re - rb + kPB - 1) / kPB) * kPB;
      for (int lr = re - rb + tid; lr < padded; lr += nthr) {
        st3(s_r, lr, V3{0, 0, 0});
        s_rf[3 * lr] = s_rf[3 * lr + 1] = s_rf[3 * lr + 2] = 0.f;
      }
    }
    // ---- initial residual, z = M^-1 r (one thread per row)
    double rz_part[1] = {0};
    for_rows([&](int i) {
      const int li = i - vb;
      if (P.pt_fixed && P.pt_fixed[i]) {  // no unknowns: keep z = 0 for this row
        if (!bprec) {
          double2* mo = reinterpret_cast<double2*>(MI + 8 * (size_t)li);
          mo[0] = mo[1] = mo[2] = make_double2(0.0, 0.0);
        }
        st3(R, li, V3{0, 0, 0});
        st3(X, li, V3{0, 0, 0});
        st3(P.zvec, i, V3{0, 0, 0});
        if (res) st3s(s_z, li, V3{0, 0, 0});
        if (bprec) s_rf[3 * li] = s_rf[3 * li + 1] = s_rf[3 * li + 2] = 0.f;
        return;
      }
      const V3 r = ld3p(P.bvec, i);
      st3(R, li, r);
      st3(X, li, V3{0, 0, 0});
      if (bprec) {
        s_rf[3 * li] = (float)r.x;
        s_rf[3 * li + 1] = (float)r.y;
        s_rf[3 * li + 2] = (float)r.z;
        return;
      }
      const double2* d = reinterpret_cast<const double2*>(P.dg + 8 * (size_t)i);
      const double2 d0 = d[0], d1 = d[1], d2 = d[2];
      const double a = d0.x + lambda, b = d0.y, c = d1.x, e = d1.y + lambda, f = d2.x, g = d2.y + lambda;
      // symmetric 3x3 inverse by cofactors
      const double C00 = e * g - f * f, C01 = c * f - b * g, C02 = b * f - c * e;
      const double det = a * C00 + b * C01 + c * C02;
      const double id = 1.0 / det;
      const double m00 = C00 * id, m01 = C01 * id, m02 = C02 * id;
      const double m11 = (a * g - c * c) * id, m12 = (b * c - a * f) * id, m22 = (a * e - b * b) * id;
      double2* mo = reinterpret_cast<double2*>(MI + 8 * (size_t)li);
      mo[0] = make_double2(m00, m01);
      mo[1] = make_double2(m02, m11);
      mo[2] = make_double2(m12, m22);
      const V3 z{m00 * r.x + m01 * r.y + m02 * r.z, m01 * r.x + m11 * r.y + m12 * r.z,
                 m02 * r.x + m12 * r.y + m22 * r.z};
      st3(P.zvec, i, z);
      if (res) st3s(s_z, li, z);
      if ((SHARD && P.world > 1)) xpush3(P.xz, i, z);
      rz_part[0] += r.x * z.x + r.y * z.y + r.z * z.z;
    });
    if (bprec) {
      __syncthreads();
      if (has) rz_part[0] += prec_apply_quads(rb, re);
    }
    double rz_pose = 0;
    if (pos) {
      for (int t = tid; t < 6 * F; t += nthr) {
        s_rp[t] = s_bp[t];
        s_xp[t] = 0;
        s_pp[t] = 0;
        s_qp[t] = 0;
      }
      __syncthreads();
      for (int t = tid; t < 6 * F; t += nthr) {
        const int k = t / 6, a = t % 6;
        double s = 0;
        for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_rp[6 * k + c];
        s_zp[t] = s;
      }
      __syncthreads();
      rz_pose = pose_dot(s_rp, s_zp, 6 * F);
    }
    grid_reduce<1>(rz_part, 0);
    double rz = s_scal[0] + rz_pose;
    const double rz0 = rz;
    if (!(rz0 > 0)) {  // b == 0: delta = 0
      if (res) for_rows([&](int i) { st3(P.xcg, i, V3{0, 0, 0}); });
      return isfinite(rz0);
    }
    const double stop = P.pcg_tol * P.pcg_tol * rz0;
    double beta = 0;
    bool ok = true;
    int it = 0;
    // quad mapping of the loop and the per-row constants, kept in registers when this CTA owns at most one chunk
    const int qr = tid / kTPR, ql = tid % kTPR;
    const bool single = !WIDE && P.n_chunks <= (int)gridDim.x && (!SHARD || P.world == 1);
    bool my_fixed = false;
    int my_kf = -1, my_a0 = 0, my_a1 = 0, my_cb = 0, my_ce = 0, my_kc0 = 0, my_kc1 = 0;
    double my_su = 0;
    if (single && tid >= 32 && tid - 32 < 6 * F) {
      my_kc0 = P.kf_chunk_ptr[(tid - 32) / 6];
      my_kc1 = P.kf_chunk_ptr[(tid - 32) / 6 + 1];
    }
    if (single && (int)blockIdx.x < P.n_chunks) {
      my_cb = P.chunk_begin[blockIdx.x];
      my_ce = P.chunk_end[blockIdx.x];
      const int i = my_cb + qr;
      if (i < my_ce) {
        my_fixed = P.pt_fixed && P.pt_fixed[i];
        my_kf = P.pt_kf[i];
        my_a0 = P.inc_ptr[i];
        my_a1 = P.inc_ptr[i + 1];
        my_su = P.dg[8 * (size_t)i + 6];
      }
    }
    for (; it < P.pcg_max_iter; it++) {
      const bool first = (it == 0);
      const int par = gen & 1;
      // ---- phase 1: w = (H + lambda I) z ; p = z + beta p ; q = w + beta q ; partial p.q
      const long long tm0 = clock64();
      double pq_part[1] = {0};
      for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
        const int cb = single ? my_cb : P.chunk_begin[c], ce = single ? my_ce : P.chunk_end[c];
        const int i = cb + qr;
        const bool valid = i < ce;
        const int li = i - vb;
        double w0 = 0, w1 = 0, w2 = 0;
        V3 zi{0, 0, 0};
        bool fixed = true;
        int kf = -1;
        double su = 0;
        if (valid) {
          fixed = single ? my_fixed : (P.pt_fixed && P.pt_fixed[i]);
          kf = single ? my_kf : P.pt_kf[i];
          su = single ? my_su : P.dg[8 * (size_t)i + 6];
          zi = res ? ld3s(s_z, li) : ld3p(P.zvec, i);
          if (!fixed) {
            const int a_beg = single ? my_a0 : P.inc_ptr[i];
            const int a1 = single ? my_a1 : P.inc_ptr[i + 1];
            // lane l takes incidences a_beg + l, + 4, + 8 (one batch covers 12 incidences of the row): addresses
            // first, then every load, then the arithmetic
            for (int a0 = a_beg + ql; a0 < a1; a0 += 3 * kTPR) {
              const double* zp[3];
              const double2* cp[3];
#pragma unroll
              for (int k = 0; k < 3; k++) {
                const int a = min(a0 + kTPR * k, a1 - 1);
                if (res) {
                  cp[k] = reinterpret_cast<const double2*>(s_coef) + 2 * (a - ab);
                  zp[k] = s_zptr[a - ab];
                } else {
                  cp[k] = reinterpret_cast<const double2*>(P.pc + 4 * (size_t)(P.inc_ent[a] >> 1));
                  zp[k] = P.zvec + 4 * (size_t)P.inc_other[a];
                }
              }
              double2 c0v[3], c1v[3];
              double zx[3], zy[3], zb[3];
#pragma unroll
              for (int k = 0; k < 3; k++) {
                c0v[k] = cp[k][0];
                c1v[k] = cp[k][1];
                zx[k] = zp[k][0];
                zy[k] = zp[k][1];
                zb[k] = zp[k][2];
              }
#pragma unroll
              for (int k = 0; k < 3; k++) {
                if (a0 + kTPR * k < a1) {
                  const double dx = zi.x - zx[k], dy = zi.y - zy[k], dz = zi.z - zb[k];
                  const double ud = c0v[k].y * dx + c1v[k].x * dy + c1v[k].y * dz;
                  w0 += c0v[k].x * dx + c0v[k].y * ud;
                  w1 += c0v[k].x * dy + c1v[k].x * ud;
                  w2 += c0v[k].x * dz + c1v[k].y * ud;
                }
              }
            }
            if (P.D > 0) {
              // dampers: a row sits in ~16 of them and every incidence is a chain of dependent loads (entry ->
              // vertices + coefficient -> four z rows). Batches of kDB incidences per lane issue each level of the
              // chain for the whole batch before the first use.
              constexpr int kDB = 2;
              const int d1 = P.dinc_ptr[i + 1];
              for (int a0 = P.dinc_ptr[i] + ql; a0 < d1; a0 += kDB * kTPR) {
                int ent[kDB];
#pragma unroll
                for (int k = 0; k < kDB; k++) ent[k] = __ldg(P.dinc_ent + min(a0 + kTPR * k, d1 - 1));
                int4 vv[kDB];
                double sc[kDB];
#pragma unroll
                for (int k = 0; k < kDB; k++) {
                  vv[k] = __ldg(reinterpret_cast<const int4*>(P.dmp_v) + (ent[k] >> 2));
                  sc[k] = P.dc[4 * (size_t)(ent[k] >> 2)];
                }
                double ax[kDB], ay[kDB], az[kDB];
#pragma unroll
                for (int k = 0; k < kDB; k++) {
                  const V3 z0 = ld3p(P.zvec, vv[k].x), z1 = ld3p(P.zvec, vv[k].y), z2 = ld3p(P.zvec, vv[k].z),
                           z3 = ld3p(P.zvec, vv[k].w);
                  ax[k] = -z0.x + z1.x + z2.x - z3.x;
                  ay[k] = -z0.y + z1.y + z2.y - z3.y;
                  az[k] = -z0.z + z1.z + z2.z - z3.z;
                }
#pragma unroll
                for (int k = 0; k < kDB; k++) {
                  if (a0 + kTPR * k < d1) {
                    const int role = ent[k] & 3;
                    const double sg = ((role == 1 || role == 2) ? 1.0 : -1.0) * sc[k];
                    w0 += sg * ax[k];
                    w1 += sg * ay[k];
                    w2 += sg * az[k];
                  }
                }
              }
            }
          }
        }
        // quad reduction (whole warp participates; lanes of invalid rows carry zeros)
        #pragma unroll
        for (int o = 1; o < kTPR; o <<= 1) {
          w0 += __shfl_xor_sync(0xffffffffu, w0, o);
          w1 += __shfl_xor_sync(0xffffffffu, w1, o);
          w2 += __shfl_xor_sync(0xffffffffu, w2, o);
        }
        if (valid && ql == 0) {
          double red[6] = {0, 0, 0, 0, 0, 0};
          double q0 = w0 + (lambda + su) * zi.x, q1 = w1 + (lambda + su) * zi.y, q2 = w2 + (lambda + su) * zi.z;
          if (kf >= 0) {
            const double2* jo =
                reinterpret_cast<const double2*>(res ? s_jac + kJS * (size_t)li : P.jac + 20 * (size_t)i);
            const double omega = jo[9].x;
            if (omega != 0) {
              double A[12], B[6];
#pragma unroll
              for (int k = 0; k < 6; k++) {
                const double2 t = jo[k];
                A[2 * k] = t.x;
                A[2 * k + 1] = t.y;
              }
#pragma unroll
              for (int k = 0; k < 3; k++) {
                const double2 t = jo[6 + k];
                B[2 * k] = t.x;
                B[2 * k + 1] = t.y;
              }
              double jp0 = 0, jp1 = 0;
              if (pos) {
                const double* pk = s_zp + 6 * kf;
#pragma unroll
                for (int a = 0; a < 6; a++) {
                  jp0 += A[a] * pk[a];
                  jp1 += A[6 + a] * pk[a];
                }
              }
              const double t0 = omega * (jp0 + B[0] * zi.x + B[1] * zi.y + B[2] * zi.z);
              const double t1 = omega * (jp1 + B[3] * zi.x + B[4] * zi.y + B[5] * zi.z);
              if (!fixed) {
                q0 += B[0] * t0 + B[3] * t1;
                q1 += B[1] * t0 + B[4] * t1;
                q2 += B[2] * t0 + B[5] * t1;
              }
              if (pos) {
#pragma unroll
                for (int a = 0; a < 6; a++) red[a] = A[a] * t0 + A[6 + a] * t1;
              }
            }
          }
          if (!fixed) {
            V3 pn = zi, qn{q0, q1, q2};
            if (!first) {
              const V3 po = ld3p(Pv, li), qo = ld3p(Q, li);
              pn = V3{zi.x + beta * po.x, zi.y + beta * po.y, zi.z + beta * po.z};
              qn = V3{q0 + beta * qo.x, q1 + beta * qo.y, q2 + beta * qo.z};
            }
            st3(Pv, li, pn);
            st3(Q, li, qn);
            pq_part[0] += pn.x * qn.x + pn.y * qn.y + pn.z * qn.z;
          } else {
            st3(Pv, li, V3{0, 0, 0});
            st3(Q, li, V3{0, 0, 0});
          }
          if (pos) {
#pragma unroll
            for (int a = 0; a < 6; a++) s_pr[6 * qr + a] = red[a];
          }
        }
        // single-chunk CTAs send the pose partials through their reduction slot (one exchange with p.q)
        if (pos)
          reduce_pose_partials(ce - cb, single ? s_scal + 8 : P.chunk_part + ((size_t)par * P.n_chunks + c) * kChunkVals);
      }
      const long long tm1 = clock64();
      double pq;
      if (pos && single) {
        // block-level p.q, then slot = [p.q, pose partials]; after the barrier warp 0 sums p.q over the CTAs while
        // 6 F threads of the other warps sum the pose partials of their pose slot: one L2 round trip in total
        double v = pq_part[0];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        double* slot = P.slots + ((size_t)par * gridDim.x + blockIdx.x) * kSlotVals;
        if (tid == 0) {
          double t = s_red[0];
          for (int w = 1; w < nw; w++) t += s_red[w];
          slot[0] = t;
        } else if (tid >= 32 && tid < 38) {
          slot[1 + tid - 32] = ((int)blockIdx.x < P.n_chunks) ? s_scal[8 + tid - 32] : 0.0;
        }
        barrier();
        if (tid < 32) {
          double t = 0;
          for (int c = lane; c < (int)gridDim.x; c += 32) t += __ldcg(P.slots + ((size_t)par * gridDim.x + c) * kSlotVals);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
          if (lane == 0) s_scal[0] = t;
        } else {
         for (int t = tid - 32; t < 6 * F; t += nthr - 32) {
          const int a = t % 6;
          const bool mine = t == tid - 32;
          const int c1 = mine ? my_kc1 : P.kf_chunk_ptr[t / 6 + 1];
          double sum = 0;
          for (int cc = mine ? my_kc0 : P.kf_chunk_ptr[t / 6]; cc < c1; cc += 8) {
            double tv[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
              tv[u] = __ldcg(P.slots + ((size_t)par * gridDim.x + min(cc + u, c1 - 1)) * kSlotVals + 1 + a);
#pragma unroll
            for (int u = 0; u < 8; u++)
              if (cc + u < c1) sum += tv[u];
          }
          const double w = lambda * s_zp[t] + sum;
          s_pp[t] = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
          s_qp[t] = first ? w : w + beta * s_qp[t];
         }
        }
        __syncthreads();
        pq = s_scal[0];
        pq += pose_dot(s_pp, s_qp, 6 * F);
      } else {
        grid_reduce<1>(pq_part, 0, pos ? 1 : 0);
        pq = s_scal[0];
        const long long tp0 = clock64();
        if (pos) {
          // pose rows: w_p = lambda z_p + sum_i A_i^T t_i ; p_p, q_p by the same recurrences (replicated per CTA)
          for (int t = tid; t < 6 * F; t += nthr) {
            const int k = t / 6, a = t % 6;
            const double w = lambda * s_zp[t] + ((SHARD && P.world > 1) ? xextra(t) : sum_chunk_partials(k, a, par));
            s_pp[t] = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
            s_qp[t] = first ? w : w + beta * s_qp[t];
          }
          __syncthreads();
          pq += pose_dot(s_pp, s_qp, 6 * F);
        }
        prof[13] += clock64() - tp0;  // pose rows of the matvec (chunk-partial sums, recurrences, serial dot)
      }
      const long long tm2 = clock64();
      if (!(pq > 0) || !isfinite(pq)) {
        ok = false;
        break;
      }
      const double alpha = rz / pq;
      // ---- phase 2: x += alpha p ; r -= alpha q ; z = M^-1 r ; partial r.z   (lane l < 3 of a quad: component l)
      double rzn_part[1] = {0};
      for (int c = blockIdx.x; c < P.n_chunks; c += gridDim.x) {
        const int i = (single ? my_cb : P.chunk_begin[c]) + qr;
        const bool valid = i < (single ? my_ce : P.chunk_end[c]);
        const int li = i - vb;
        const bool act = valid && !(single ? my_fixed : (P.pt_fixed && P.pt_fixed[i]));
        if (act) {
          for (int cmp = ql; cmp < 3; cmp += kTPR) {
            const size_t o = 4 * (size_t)li + cmp;
            X[o] += alpha * Pv[o];
            const double rc = R[o] - alpha * Q[o];
            R[o] = rc;
            if (bprec) s_rf[3 * li + cmp] = (float)rc;
          }
        }
        if (!bprec) {
          __syncwarp();  // the lanes of the quad wrote one component each
          if (act && ql == 0) {
            const V3 ri = ld3p(R, li);
            const double2* mo = reinterpret_cast<const double2*>(MI + 8 * (size_t)li);
            const double2 m0 = mo[0], m1 = mo[1], m2 = mo[2];
            const V3 z{m0.x * ri.x + m0.y * ri.y + m1.x * ri.z, m0.y * ri.x + m1.y * ri.y + m2.x * ri.z,
                       m1.x * ri.x + m2.x * ri.y + m2.y * ri.z};
            st3(P.zvec, i, z);
            if (res) st3s(s_z, li, z);
            if ((SHARD && P.world > 1)) xpush3(P.xz, i, z);
            rzn_part[0] += ri.x * z.x + ri.y * z.y + ri.z * z.z;
          }
        }
      }
      if (bprec) {
        __syncthreads();
        if (has) rzn_part[0] += prec_apply_quads(rb, re);
      }
      double rzn_pose = 0;
      if (pos) {
        for (int t = tid; t < 6 * F; t += nthr) {
          s_xp[t] += alpha * s_pp[t];
          s_rp[t] -= alpha * s_qp[t];
        }
        __syncthreads();
        for (int t = tid; t < 6 * F; t += nthr) {
          const int k = t / 6, a = t % 6;
          double s = 0;
          for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_rp[6 * k + c];
          s_zp[t] = s;
        }
        __syncthreads();
        rzn_pose = pose_dot(s_rp, s_zp, 6 * F);
      }
      const long long tm3 = clock64();
      grid_reduce<1>(rzn_part, 0);
      const long long tm4 = clock64();
      prof[1] += tm1 - tm0;  // matvec pass
      prof[2] += tm2 - tm1;  // pq reduce (incl. barrier)
      prof[3] += tm3 - tm2;  // update pass
      prof[4] += tm4 - tm3;  // rz reduce (incl. barrier)
      const double rzn = s_scal[0] + rzn_pose;
      if (!isfinite(rzn)) {
        ok = false;
        it++;
        break;
      }
      beta = rzn / rz;
      rz = rzn;
      if (rz <= stop) {
        it++;
        break;
      }
    }
    pcg_iters += it;
    if (res) for_rows([&](int i) { st3(P.xcg, i, ld3p(s_x, i - vb)); });
    return ok;
  }

  // ================================================================================================
  // CG loop of the WIDE variant (large BA windows, cooperative grid, vectors in L2 / HBM). Same recurrences and the
  // same 3x3 block-Jacobi preconditioner as pcg(), organised for throughput instead of lock-step chunk passes:
  //  * every CTA owns ONE contiguous row range of equal length (host: wseg_*), cut into segments at the pose-slot
  //    boundaries — no chunk-count quantisation between CTAs;
  //  * a row group (kTPR lanes) keeps the pose partial of its rows in REGISTERS for a whole segment: one block
  //    reduction per segment (usually one or two per CTA and iteration) instead of three block syncs per 128-row
  //    chunk pass, so the warps of a CTA run through their rows independently and hide each other's gather chains;
  //  * the pair coefficients are expanded into incidence order once per solve (wrec) and a damper incidence is one
  //    record (the three other vertices ordered (+, -, -) and the coefficient): c (z_i + z_+ - z_- - z_-), one
  //    dependent level and one gathered row less than id -> vertices -> four rows;
  //  * the update pass runs one thread per row.
  // All sums keep a fixed order: results are reproducible run to run.
  // ================================================================================================
  __device__ __forceinline__ double block_sum(double v, int buf) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    double* b = s_scal + 8 + 8 * buf;  // two buffers: the next call may start before every thread has read this one
    if (lane == 0) b[warp] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < nw; w++) s += b[w];
    return s;
  }

  // Row-local part of the NEXT matvec, evaluated where z_i is produced (initial residual, update pass):
  // wl_i = (lambda + s_unary) z_i + B_i^T t_i with t_i = omega_i (A_i z_p + B_i z_i), and the row's pose partial
  // A_i^T t_i added to red. A fixed row (z_i = 0) still feeds the pose rows. The Jacobian is component-major; a row
  // without a reprojection edge stores omega = 0 and zero blocks, so nothing is branched on and the ten loads can be
  // issued together with the caller's (j: loaded by the caller before it needs z).
  __device__ __forceinline__ void wide_load_jac(int i, double2 (&j)[10]) {
    const double2* jt = reinterpret_cast<const double2*>(P.jac) + i;
    const size_t V = (size_t)P.V;
#pragma unroll
    for (int k = 0; k < 10; k++) j[k] = jt[k * V];
  }
  __device__ __forceinline__ void wide_row_local(int i, int kf, bool fixed, double su, const V3& z, bool pos,
                                                 const double2 (&j)[10], double (&red)[6]) {
    V3 wl{(lambda + su) * z.x, (lambda + su) * z.y, (lambda + su) * z.z};
    const double omega = j[9].x;
    double jp0 = 0, jp1 = 0;
    if (pos) {
      const double* pk = s_zp + 6 * kf;
      jp0 = j[0].x * pk[0] + j[0].y * pk[1] + j[1].x * pk[2] + j[1].y * pk[3] + j[2].x * pk[4] + j[2].y * pk[5];
      jp1 = j[3].x * pk[0] + j[3].y * pk[1] + j[4].x * pk[2] + j[4].y * pk[3] + j[5].x * pk[4] + j[5].y * pk[5];
    }
    // B = [j6.x j6.y j7.x ; j7.y j8.x j8.y]
    const double t0 = omega * (jp0 + j[6].x * z.x + j[6].y * z.y + j[7].x * z.z);
    const double t1 = omega * (jp1 + j[7].y * z.x + j[8].x * z.y + j[8].y * z.z);
    wl.x += j[6].x * t0 + j[7].y * t1;
    wl.y += j[6].y * t0 + j[8].x * t1;
    wl.z += j[7].x * t0 + j[8].y * t1;
    if (pos) {
      red[0] += j[0].x * t0 + j[3].x * t1;
      red[1] += j[0].y * t0 + j[3].y * t1;
      red[2] += j[1].x * t0 + j[4].x * t1;
      red[3] += j[1].y * t0 + j[4].y * t1;
      red[4] += j[2].x * t0 + j[5].x * t1;
      red[5] += j[2].y * t0 + j[5].y * t1;
    }
    if (!fixed) st4(P.wvec, i, wl.x, wl.y, wl.z);
  }

  // One thread per row over this CTA's segments; fn(i, slot, red) adds the row's pose partial to red. The partials of a
  // segment are summed in a fixed order (lanes, warps) and stored to wseg_part[buf][segment].
  template <typename Fn>
  __device__ __forceinline__ void wide_segments(int sg0, int sg1, bool pos, int buf, Fn fn) {
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    int slot = 0;
    for (int sg = sg0; sg < sg1; sg++) {
      const int se = P.wseg_end[sg], kf = P.wseg_kf[sg];
      double red[6] = {0, 0, 0, 0, 0, 0};
      for (int i = P.wseg_begin[sg] + tid; i < se; i += nthr) fn(i, kf, red);
      if (pos) {
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) red[a] += __shfl_xor_sync(0xffffffffu, red[a], off);
        }
        if (lane == 0) {
#pragma unroll
          for (int a = 0; a < 6; a++) s_red[(slot * nw + warp) * 6 + a] = red[a];
        }
        slot++;
        if (slot == 8 || sg == sg1 - 1) {  // the warps of up to 8 segments are combined per flush
          __syncthreads();
          for (int tt = tid; tt < slot * 6; tt += nthr) {
            const int sl = tt / 6, a = tt % 6;
            double sum = 0;
            for (int w = 0; w < nw; w++) sum += s_red[(sl * nw + w) * 6 + a];
            P.wseg_part[((size_t)buf * P.n_wseg + (sg - slot + 1 + sl)) * 8 + a] = sum;
          }
          __syncthreads();
          slot = 0;
        }
      }
    }
  }

  __device__ bool pcg_wide() {
    const int F = P.F;
    const bool pos = !P.poses_fixed;
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    if (pos) {
      for (int k = tid; k < F; k += nthr)
        if (!invert6(s_H + 21 * k, lambda, s_M + 36 * k)) *s_flag = 1;
    }
    __syncthreads();
    if (*s_flag) return false;  // uniform: every CTA inverts the same blocks
    const int sg0 = P.wseg_ptr[blockIdx.x], sg1 = P.wseg_ptr[blockIdx.x + 1];
    const int r0 = sg0 < sg1 ? P.wseg_begin[sg0] : 0, r1 = sg0 < sg1 ? P.wseg_end[sg1 - 1] : 0;  // this CTA's rows
    double* const minvA = P.minv;                      // m00 m01 m02 m11
    double* const minvB = P.minv + 4 * (size_t)P.V;    // m12 m22 s_unary -
    // ---- incidence records of this CTA's rows
    {
      const int ab = P.inc_ptr[r0], ae = P.inc_ptr[r1];
      for (int a = ab + tid; a < ae; a += nthr) {
        const double2* cf = reinterpret_cast<const double2*>(P.pc + 4 * (size_t)(P.inc_ent[a] >> 1));
        const double2 c0 = __ldcg(cf), c1 = __ldcg(cf + 1);
        st4(P.wrec, a, c0.x, c0.y, c1.x, c1.y);
      }
      if (P.D > 0) {
        const int db = P.dinc_ptr[r0], de = P.dinc_ptr[r1];
        for (int a = db + tid; a < de; a += nthr) {
          const int ent = P.dinc_ent[a], role = ent & 3;
          const int4 v = __ldg(reinterpret_cast<const int4*>(P.dmp_v) + (ent >> 2));
          // J = (-w, +w, +w, -w) I over (x, y, z, w): H_rk = c sigma_r sigma_k; seen from role r exactly one of the
          // three other vertices enters with +c
          int o0, o1, o2;
          if (role == 0) { o0 = v.w; o1 = v.y; o2 = v.z; }
          else if (role == 1) { o0 = v.z; o1 = v.x; o2 = v.w; }
          else if (role == 2) { o0 = v.y; o1 = v.x; o2 = v.w; }
          else { o0 = v.x; o1 = v.y; o2 = v.z; }
          st4(P.wdrec, a, __hiloint2double(o1, o0), __hiloint2double(0, o2), __ldcg(P.dc + 4 * (size_t)(ent >> 2)), 0.0);
        }
      }
    }
    // ---- pose part of the initial residual (replicated per CTA): z_p is read by the rows below
    double rz_pose = 0;
    if (pos) {
      for (int t = tid; t < 6 * F; t += nthr) {
        s_rp[t] = s_bp[t];
        s_xp[t] = 0;
        s_pp[t] = 0;
        s_qp[t] = 0;
      }
      __syncthreads();
      double v = 0;
      for (int t = tid; t < 6 * F; t += nthr) {
        const int k = t / 6, a = t % 6;
        double s = 0;
        for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_rp[6 * k + c];
        s_zp[t] = s;
        v += s_rp[t] * s;
      }
      rz_pose = block_sum(v, 0);
    }
    // ---- initial residual, z = M^-1 r, row-local part of the first matvec (one thread per row)
    double rz_part[1] = {0};
    wide_segments(sg0, sg1, pos, 0, [&](int i, int kf, double (&red)[6]) {
      double2 j[10];
      wide_load_jac(i, j);
      if (P.pt_fixed && P.pt_fixed[i]) {  // no unknowns: z = 0 for this row
        st4(minvA, i, 0, 0, 0, 0);
        st4(minvB, i, 0, 0, 0, 0);
        st4(P.rvec, i, 0, 0, 0);
        st4(P.xcg, i, 0, 0, 0);
        st4(P.zvec, i, 0, 0, 0);
        st4(P.pvec, i, 0, 0, 0);
        st4(P.qvec, i, 0, 0, 0);
        wide_row_local(i, kf, true, 0.0, V3{0, 0, 0}, pos, j, red);
        return;
      }
      const D4 r = ld4(P.bvec, i);
      st4(P.rvec, i, r.x, r.y, r.z);
      st4(P.xcg, i, 0, 0, 0);
      const D4 d0 = ld4(P.dg, 2 * (size_t)i), d1 = ld4(P.dg, 2 * (size_t)i + 1);  // 00 01 02 11 | 12 22 su -
      const double a = d0.x + lambda, b = d0.y, c = d0.z, e = d0.w + lambda, f = d1.x, g = d1.y + lambda;
      const double C00 = e * g - f * f, C01 = c * f - b * g, C02 = b * f - c * e;  // symmetric 3x3 inverse by cofactors
      const double det = a * C00 + b * C01 + c * C02;
      const double id = 1.0 / det;
      const double m00 = C00 * id, m01 = C01 * id, m02 = C02 * id;
      const double m11 = (a * g - c * c) * id, m12 = (b * c - a * f) * id, m22 = (a * e - b * b) * id;
      st4(minvA, i, m00, m01, m02, m11);
      st4(minvB, i, m12, m22, d1.z, 0.0);
      const V3 z{m00 * r.x + m01 * r.y + m02 * r.z, m01 * r.x + m11 * r.y + m12 * r.z,
                 m02 * r.x + m12 * r.y + m22 * r.z};
      st4(P.zvec, i, z.x, z.y, z.z);
      if ((SHARD && P.world > 1)) xpush3(P.xz, i, z);
      rz_part[0] += r.x * z.x + r.y * z.y + r.z * z.z;
      wide_row_local(i, kf, false, d1.z, z, pos, j, red);
    });
    grid_reduce<1>(rz_part, 0);
    double rz = s_scal[0] + rz_pose;
    const double rz0 = rz;
    if (!(rz0 > 0)) return isfinite(rz0);  // b == 0: delta = 0 (xcg is zero already)
    const double stop = P.pcg_tol * P.pcg_tol * rz0;
    double beta = 0;
    bool ok = true;
    int it = 0;
    const int qr = tid / kTPR, ql = tid % kTPR;
    const int RPP = nthr / kTPR;  // row groups per pass
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    // segment range of the first pose value this thread sums after the p.q barrier (fixed over the loop)
    int my_c0 = 0, my_c1 = 0;
    if (tid >= 32 && tid - 32 < 6 * F) {
      my_c0 = P.kf_wseg_ptr[(tid - 32) / 6];
      my_c1 = P.kf_wseg_ptr[(tid - 32) / 6 + 1];
    }
    for (; it < P.pcg_max_iter; it++) {
      const bool first = (it == 0);
      const int wbuf = it & 1;  // wseg_part buffer holding this iteration's pose partials
      // ---- phase 1: w = (H + lambda I) z ; p = z + beta p ; q = w + beta q ; partial p.q
      //      (regulariser gathers only: the row-local part came with z). No block syncs: warps run independently.
      const long long tm0 = clock64();
      double pq_part[1] = {0};
      for (int base = r0; base < r1; base += RPP) {
        const int i = base + qr;
        const bool valid = i < r1 && !(P.pt_fixed && P.pt_fixed[i]);
        double w0 = 0, w1 = 0, w2 = 0;
        D4 zi{0, 0, 0, 0}, wl{0, 0, 0, 0}, po{0, 0, 0, 0}, qo{0, 0, 0, 0};
        if (valid) {
          const int a_beg = P.inc_ptr[i], a1 = P.inc_ptr[i + 1];
          const int d_beg = P.D > 0 ? P.dinc_ptr[i] : 0, d1 = P.D > 0 ? P.dinc_ptr[i + 1] : 0;
          zi = ld4(P.zvec, i);
          if (ql == 0) {
            wl = ld4(P.wvec, i);
            if (!first) {
              po = ld4(P.pvec, i);
              qo = ld4(P.qvec, i);
            }
          }
          // the damper records of this lane are fetched into L1 while the spring batch walks its two dependent
          // levels (record -> z): their loads below are L1 hits instead of one more L2 round trip per batch
          if (P.wide_prefetch) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (d_beg + ql + kTPR * k < d1) prefetch_l1(P.wdrec + 4 * (size_t)(d_beg + ql + kTPR * k));
          }
          // lane l takes incidences a_beg + l, + kTPR, ...: batches of 3 per lane, every load of a level first
          for (int a0 = a_beg + ql; a0 < a1; a0 += 3 * kTPR) {
            int oth[3];
            D4 cf[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const int a = min(a0 + kTPR * k, a1 - 1);
              oth[k] = P.inc_other[a];
              cf[k] = ld4(P.wrec, a);
            }
            D4 zo[3];
#pragma unroll
            for (int k = 0; k < 3; k++) zo[k] = ld4(P.zvec, oth[k]);
#pragma unroll
            for (int k = 0; k < 3; k++) {
              if (a0 + kTPR * k < a1) {
                const double dx = zi.x - zo[k].x, dy = zi.y - zo[k].y, dz = zi.z - zo[k].z;
                const double ud = cf[k].y * dx + cf[k].z * dy + cf[k].w * dz;
                w0 += cf[k].x * dx + cf[k].y * ud;
                w1 += cf[k].x * dy + cf[k].z * ud;
                w2 += cf[k].x * dz + cf[k].w * ud;
              }
            }
          }
          constexpr int kDB = 2;
          for (int a0 = d_beg + ql; a0 < d1; a0 += kDB * kTPR) {
            D4 rec[kDB];
#pragma unroll
            for (int k = 0; k < kDB; k++) rec[k] = ld4(P.wdrec, min(a0 + kTPR * k, d1 - 1));
            D4 zp[kDB], zm[kDB], zn[kDB];
#pragma unroll
            for (int k = 0; k < kDB; k++) {
              zp[k] = ld4(P.zvec, __double2loint(rec[k].x));
              zm[k] = ld4(P.zvec, __double2hiint(rec[k].x));
              zn[k] = ld4(P.zvec, __double2loint(rec[k].y));
            }
#pragma unroll
            for (int k = 0; k < kDB; k++) {
              if (a0 + kTPR * k < d1) {
                w0 += rec[k].z * ((zi.x + zp[k].x) - (zm[k].x + zn[k].x));
                w1 += rec[k].z * ((zi.y + zp[k].y) - (zm[k].y + zn[k].y));
                w2 += rec[k].z * ((zi.z + zp[k].z) - (zm[k].z + zn[k].z));
              }
            }
          }
        }
        // reduction over the lanes of the row (whole warp participates; lanes of invalid rows carry zeros)
#pragma unroll
        for (int o = 1; o < kTPR; o <<= 1) {
          w0 += __shfl_xor_sync(0xffffffffu, w0, o);
          w1 += __shfl_xor_sync(0xffffffffu, w1, o);
          w2 += __shfl_xor_sync(0xffffffffu, w2, o);
        }
        if (valid && ql == 0) {
          const double p0 = zi.x + beta * po.x, p1 = zi.y + beta * po.y, p2 = zi.z + beta * po.z;
          const double q0 = (w0 + wl.x) + beta * qo.x, q1 = (w1 + wl.y) + beta * qo.y, q2 = (w2 + wl.z) + beta * qo.z;
          st4(P.pvec, i, p0, p1, p2);
          st4(P.qvec, i, q0, q1, q2);
          pq_part[0] += p0 * q0 + p1 * q1 + p2 * q2;
        }
      }
      const long long tm1 = clock64();
      double pq;
      long long tp0;
      if (pos && (!SHARD || P.world == 1)) {
        // p.q over the grid and the pose rows of the matvec in ONE round trip: after the barrier warp 0 sums the
        // CTAs' slots while the other warps sum the segment partials of the pose slots and run their recurrences
        double v = pq_part[0];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        const int par = gen & 1;
        if (tid == 0) {
          double t = s_red[0];
          for (int w = 1; w < nw; w++) t += s_red[w];
          P.slots[((size_t)par * gridDim.x + blockIdx.x) * kSlotVals] = t;
        }
        barrier();
        tp0 = clock64();
        double pv = 0;
        if (tid < 32) {
          double t = 0;
          for (int c0 = lane; c0 < (int)gridDim.x; c0 += 8 * 32) {
            double o[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
              o[u] = __ldcg(P.slots + ((size_t)par * gridDim.x + min(c0 + 32 * u, (int)gridDim.x - 1)) * kSlotVals);
#pragma unroll
            for (int u = 0; u < 8; u++)
              if (c0 + 32 * u < (int)gridDim.x) t += o[u];
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
          if (lane == 0) s_scal[0] = t;
        } else {
          // three pose values per thread and round: the segment ranges first, then the first twelve partials of each
          // value (all in flight together), then the rare longer tails — one L2 round trip instead of two per value
          const int pstep = nthr - 32;
          const double* part = P.wseg_part + (size_t)wbuf * P.n_wseg * 8;
          if (6 * F <= pstep) {  // one value per thread (<= 37 poses): nothing to batch
            const int t = tid - 32;
            if (t < 6 * F) {
              const double w = lambda * s_zp[t] + sum_wseg_range(P, my_c0, my_c1, t % 6, wbuf);
              const double pp = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
              const double qp = first ? w : w + beta * s_qp[t];
              s_pp[t] = pp;
              s_qp[t] = qp;
              pv += pp * qp;
            }
          } else
          for (int t0 = tid - 32; t0 < 6 * F; t0 += 3 * pstep) {
            int cb[3], ce[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const int t = t0 + u * pstep, k = min(t, 6 * F - 1) / 6;
              const bool mine = (u == 0 && t0 == tid - 32);
              cb[u] = mine ? my_c0 : (t < 6 * F ? P.kf_wseg_ptr[k] : 0);
              ce[u] = mine ? my_c1 : (t < 6 * F ? P.kf_wseg_ptr[k + 1] : 0);
            }
            double o[3][12];
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const int t = t0 + u * pstep, a = min(t, 6 * F - 1) % 6;
              // predicated, not clamped: clamped duplicates of every thread of every CTA hammer ONE L2 line
#pragma unroll
              for (int q = 0; q < 12; q++)
                o[u][q] = (t < 6 * F && cb[u] + q < ce[u]) ? __ldcg(part + (size_t)(cb[u] + q) * 8 + a) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const int t = t0 + u * pstep;
              if (t >= 6 * F) continue;
              double sum = 0;
#pragma unroll
              for (int q = 0; q < 12; q++)
                if (cb[u] + q < ce[u]) sum += o[u][q];
              for (int c = cb[u] + 12; c < ce[u]; c++) sum += __ldcg(part + (size_t)c * 8 + t % 6);
              const double w = lambda * s_zp[t] + sum;
              const double pp = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
              const double qp = first ? w : w + beta * s_qp[t];
              s_pp[t] = pp;
              s_qp[t] = qp;
              pv += pp * qp;
            }
          }
        }
        const double ps = block_sum(pv, 1);  // its block sync publishes s_scal[0]
        pq = s_scal[0] + ps;
      } else {
        grid_reduce<1>(pq_part, 0, pos ? 1 : 0, wbuf);
        pq = s_scal[0];
        tp0 = clock64();
        if (pos) {
          // pose rows: w_p = lambda z_p + sum_i A_i^T t_i ; p_p, q_p by the same recurrences (replicated per CTA)
          double v = 0;
          if (SHARD && P.world > 1) {
            // three values per thread and round, the records of every rank in flight together (one L2 round trip)
            const double* base = P.xred[P.rank] + xcur + 8;
            for (int t0 = tid; t0 < 6 * F; t0 += 3 * nthr) {
              double o[3][kMaxWorld];
#pragma unroll
              for (int u = 0; u < 3; u++) {
                const int t = t0 + u * nthr;
#pragma unroll
                for (int r = 0; r < kMaxWorld; r++)
                  o[u][r] = (t < 6 * F && r < P.world) ? __ldcg(base + t + (size_t)r * P.xstride) : 0.0;
              }
#pragma unroll
              for (int u = 0; u < 3; u++) {
                const int t = t0 + u * nthr;
                if (t >= 6 * F) continue;
                double xs = 0;
#pragma unroll
                for (int r = 0; r < kMaxWorld; r++)
                  if (r < P.world) xs += o[u][r];
                const double w = lambda * s_zp[t] + xs;
                const double pp = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
                const double qp = first ? w : w + beta * s_qp[t];
                s_pp[t] = pp;
                s_qp[t] = qp;
                v += pp * qp;
              }
            }
          } else {
            for (int t = tid; t < 6 * F; t += nthr) {
              const double w = lambda * s_zp[t] + sum_wseg_partials_of(P, t / 6, t % 6, wbuf);
              const double pp = first ? s_zp[t] : s_zp[t] + beta * s_pp[t];
              const double qp = first ? w : w + beta * s_qp[t];
              s_pp[t] = pp;
              s_qp[t] = qp;
              v += pp * qp;
            }
          }
          pq += block_sum(v, 1);
        }
      }
      prof[13] += clock64() - tp0;  // pose rows of the matvec
      const long long tm2 = clock64();
      if (!(pq > 0) || !isfinite(pq)) {
        ok = false;
        break;
      }
      const double alpha = rz / pq;
      // ---- phase 2: x += alpha p ; r -= alpha q ; z = M^-1 r ; partial r.z ; row-local part of the next matvec
      double rzn_pose = 0;
      if (pos) {
        for (int t = tid; t < 6 * F; t += nthr) {
          s_xp[t] += alpha * s_pp[t];
          s_rp[t] -= alpha * s_qp[t];
        }
        __syncthreads();
        double v = 0;
        for (int t = tid; t < 6 * F; t += nthr) {
          const int k = t / 6, a = t % 6;
          double s = 0;
          for (int c = 0; c < 6; c++) s += s_M[36 * k + a * 6 + c] * s_rp[6 * k + c];
          s_zp[t] = s;
          v += s_rp[t] * s;
        }
        rzn_pose = block_sum(v, 0);  // (its block sync also publishes z_p to the rows below)
      }
      double rzn_part[1] = {0};
      wide_segments(sg0, sg1, pos, wbuf ^ 1, [&](int i, int kf, double (&red)[6]) {
        // one level of loads: the vectors, the preconditioner block and the Jacobian of the row
        const D4 ro = ld4(P.rvec, i), qo = ld4(P.qvec, i), mA = ld4(minvA, i), mB = ld4(minvB, i);
        double2 j[10];
        wide_load_jac(i, j);
        if (P.pt_fixed && P.pt_fixed[i]) {
          wide_row_local(i, kf, true, 0.0, V3{0, 0, 0}, pos, j, red);
          return;
        }
        const D4 xo = ld4(P.xcg, i), po = ld4(P.pvec, i);
        const V3 ri{ro.x - alpha * qo.x, ro.y - alpha * qo.y, ro.z - alpha * qo.z};
        const V3 z{mA.x * ri.x + mA.y * ri.y + mA.z * ri.z, mA.y * ri.x + mA.w * ri.y + mB.x * ri.z,
                   mA.z * ri.x + mB.x * ri.y + mB.y * ri.z};
        st4(P.rvec, i, ri.x, ri.y, ri.z);
        st4(P.zvec, i, z.x, z.y, z.z);
        if ((SHARD && P.world > 1)) xpush3(P.xz, i, z);
        rzn_part[0] += ri.x * z.x + ri.y * z.y + ri.z * z.z;
        wide_row_local(i, kf, false, mB.z, z, pos, j, red);
        st4(P.xcg, i, xo.x + alpha * po.x, xo.y + alpha * po.y, xo.z + alpha * po.z);
      });
      const long long tm3 = clock64();
      grid_reduce<1>(rzn_part, 0);
      const long long tm4 = clock64();
      prof[1] += tm1 - tm0;  // matvec pass
      prof[2] += tm2 - tm1;  // pq reduce (incl. barrier)
      prof[3] += tm3 - tm2;  // update pass
      prof[4] += tm4 - tm3;  // rz reduce (incl. barrier)
      const double rzn = s_scal[0] + rzn_pose;
      if (!isfinite(rzn)) {
        ok = false;
        it++;
        break;
      }
      beta = rzn / rz;
      rz = rzn;
      if (rz <= stop) {
        it++;
        break;
      }
    }
    pcg_iters += it;
    return ok;
  }

  // ================================================================================================
  // Cluster-native CG loop for a tracking frame (cluster mode, resident, one chunk per CTA, one pose, no dampers).
  // Nothing in the loop touches global memory: z of the out-of-chunk neighbours and the reduction values are PUSHED
  // into the consumers' shared memory (remote stores before the cluster barrier, local loads after it).
  // Preconditioner: dense 16-row blocks + (P.coarse) an additive coarse level with one 3-dof aggregate per CTA and the
  // 6 pose unknowns — the pose / common-mode deformation gauge direction is only held by lambda, and a 54 x 54 Galerkin
  // system solved redundantly by every CTA removes it from the CG spectrum (about 1.6x fewer iterations).
  // Per iteration: 5-6 CTA barriers + 2 cluster barriers; warp 0 does the small serial parts while the rest wait.
  // ================================================================================================
  __device__ bool pcg_cluster() {
    NRS_SHARED(s_jac); NRS_SHARED(s_x); NRS_SHARED(s_r); NRS_SHARED(s_p); NRS_SHARED(s_q); NRS_SHARED(s_z);
    NRS_SHARED(s_minv); NRS_SHARED(s_coef); NRS_SHARED(s_zptr); NRS_SHARED(s_rf); NRS_SHARED(s_pr);
    NRS_SHARED(s_red); NRS_SHARED(s_scal); NRS_SHARED(s_zp); NRS_SHARED(s_pp); NRS_SHARED(s_qp); NRS_SHARED(s_rp);
    NRS_SHARED(s_xp); NRS_SHARED(s_M); NRS_SHARED(s_bp); NRS_SHARED(s_halo); NRS_SHARED(s_gather);
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const bool pos = !P.poses_fixed;
    const bool coarse = P.coarse != 0 && (blockDim.x >> 5) >= 8;  // the sweep splits the 54 rows over >= 8 warps
    bool use_coarse = coarse;
    const int G = gridDim.x;
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    if (pos && !coarse && tid == 0)
      if (!invert6(s_H, lambda, s_M)) *s_flag = 1;
    __syncthreads();
    if (*s_flag) return false;  // uniform: every CTA inverts the same block
    const long long ts0 = clock64();
    const int c0 = blockIdx.x;
    const int rb = P.chunk_begin[c0], re = P.chunk_end[c0], nrows = re - rb;
    const int ab = P.inc_ptr[rb], ae = P.inc_ptr[re];
    const bool bprec = P.block_prec != 0;
    for (int t = tid; t < 10 * nrows; t += nthr)
      reinterpret_cast<double2*>(s_jac + kJS * (size_t)(t / 10))[t % 10] =
          reinterpret_cast<const double2*>(P.jac + 20 * (size_t)rb)[t];
    if (coarse) {
      for (int t = tid; t < P.halo_rows; t += nthr) s_hcid[t] = 255;  // unused halo slots: never corrected
      __syncthreads();
    }
    for (int a = ab + tid; a < ae; a += nthr) {
      const int ent = P.inc_ent[a];
      const double2* cf = reinterpret_cast<const double2*>(P.pc + 4 * (size_t)(ent >> 1));
      reinterpret_cast<double2*>(s_coef)[2 * (a - ab)] = cf[0];
      reinterpret_cast<double2*>(s_coef)[2 * (a - ab) + 1] = cf[1];
      const int other = P.inc_other[a];
      const bool in = other >= rb && other < re;
      if (in) {
        s_zptr[a - ab] = s_z + 3 * (size_t)(other - rb);
      } else {
        s_zptr[a - ab] = s_halo + 3 * (size_t)P.inc_halo[a];  // pushed by the owner after every z update
      }
      if (coarse) {
        int cid = c0;
        if (!in) {  // owning chunk == aggregate: binary search over the chunk starts
          int lo = 0, hi = P.n_chunks - 1;
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.chunk_begin[mid] <= other) lo = mid; else hi = mid - 1;
          }
          cid = lo;
        }
        const bool ofixed = P.pt_fixed && P.pt_fixed[other];
        s_cid[a - ab] = ofixed ? 255 : (unsigned char)cid;
        if (!in) s_hcid[P.inc_halo[a]] = ofixed ? 255 : (unsigned char)cid;
      }
    }
    __syncthreads();
    const long long ts1 = clock64();
    prof[8] += ts1 - ts0;   // setup: copies, pointers
    if (bprec) build_block_prec(rb, re, ab);
    const long long ts2 = clock64();
    prof[9] += ts2 - ts1;   // setup: block preconditioner
    if (bprec) {  // rows of the last block past the chunk end: zero residual
      const int padded = ((nrows + kPB - 1) / kPB) * kPB;
      for (int lr = nrows + tid; lr < padded; lr += nthr) {
        st3(s_r, lr, V3{0, 0, 0});
        s_rf[3 * lr] = s_rf[3 * lr + 1] = s_rf[3 * lr + 2] = 0.f;
      }
    }
    double* s_slot = s_scal + 8;   // 16 doubles: staging of this CTA's reduction values
    double* s_bc = s_scal + 24;    // broadcast scalars
    const int my_rank = blockIdx.x;
    // Exchange by PUSH: warp 0 stores this CTA's values into every CTA's gather buffer (remote stores are fire and
    // forget); after the cluster barrier everybody sums its LOCAL copy.
    auto push = [&](int buf, int k0, int nk) {  // called by warp 0 after its values sit in s_slot[k0 .. k0 + nk)
      __syncwarp();
      const int target = lane & 15, half = lane >> 4;  // two lanes per target CTA share the values
      if (target < G) {
        double* dst = cluster.map_shared_rank(s_gather, target) + (size_t)(buf * 16 + my_rank) * kGatherVals;
        for (int k = half; k < nk; k += 2) dst[k0 + k] = s_slot[k0 + k];
      }
    };
    // halo push: every thread takes entries of this chunk's push list (row -> target chunk, slot)
    const int hp0 = P.push_ptr[c0], hp1 = P.push_ptr[c0 + 1];
    auto push_halo = [&]() {  // after a CTA barrier that follows the z update
      for (int e = hp0 + tid; e < hp1; e += nthr) {
        const int row = P.push_row[e], dst = P.push_dst[e];
        double* h = cluster.map_shared_rank(s_halo, dst >> 16) + 3 * (size_t)(dst & 65535);
        const V3 z = ld3s(s_z, row - rb);
        h[0] = z.x;
        h[1] = z.y;
        h[2] = z.z;
      }
    };

    // ---- coarse level: Galerkin matrix Z^T (H + lambda I) Z over [pose ; one 3-dof aggregate per chunk]
    const long long ts3 = clock64();
    if (coarse) {
      NRS_SHARED(s_cid); NRS_SHARED(s_Aall); NRS_SHARED(s_Ac);
      float* stage = s_Ac;  // the coarse matrix is assembled only after the row blocks have been exchanged
      // this CTA's row block: 16 x 6 (symmetric 3x3 per target aggregate) + 18 (pose coupling) + [n free rows].
      // (1) one thread per row: the row's share of the OWN aggregate block (diagonal block + lambda minus the
      //     in-chunk couplings), of the pose coupling and of the free-row count; fixed-order block reduction.
      {
        double v[25];
#pragma unroll
        for (int k = 0; k < 25; k++) v[k] = 0;
        if (tid < nrows && !(P.pt_fixed && P.pt_fixed[rb + tid])) {
          const int i = rb + tid;
          const double* d = P.dg + 8 * (size_t)i;  // 00 01 02 11 12 22
#pragma unroll
          for (int k = 0; k < 6; k++) v[k] = d[k];
          v[0] += lambda;
          v[3] += lambda;
          v[5] += lambda;
          for (int a = P.inc_ptr[i]; a < P.inc_ptr[i + 1]; a++) {
            if (s_cid[a - ab] != c0) continue;
            const double* cf = s_coef + 4 * (size_t)(a - ab);
            v[0] -= cf[0] + cf[1] * cf[1];
            v[1] -= cf[1] * cf[2];
            v[2] -= cf[1] * cf[3];
            v[3] -= cf[0] + cf[2] * cf[2];
            v[4] -= cf[2] * cf[3];
            v[5] -= cf[0] + cf[3] * cf[3];
          }
          if (pos && P.pt_kf[i] >= 0) {
            const double* jo = s_jac + kJS * (size_t)tid;  // A(2x6) B(2x3) omega
#pragma unroll
            for (int a6 = 0; a6 < 6; a6++)
#pragma unroll
              for (int b3 = 0; b3 < 3; b3++) v[6 + a6 * 3 + b3] = jo[18] * (jo[a6] * jo[12 + b3] + jo[6 + a6] * jo[15 + b3]);
          }
          v[24] = 1.0;
        }
#pragma unroll
        for (int k = 0; k < 25; k++) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 25; k++) s_red[128 + 25 * warp + k] = v[k];
        }
      }
      __syncthreads();
      for (int t = tid; t < kRowBlk; t += nthr) stage[t] = 0.f;
      __syncthreads();
      if (tid < 25) {
        double s = 0;
        for (int w = 0; w < nw; w++) s += s_red[128 + 25 * w + tid];
        if (tid < 6) stage[c0 * 6 + tid] = (float)s;
        else if (tid < 24) stage[96 + tid - 6] = (float)s;
        else stage[114] = (float)s;
      }
      // (2) couplings to OTHER aggregates: (aggregate, component) threads scan the chunk's cross-chunk incidences
      if (tid >= 32 && tid < 32 + 96) {
        const int ct = (tid - 32) / 6, comp = (tid - 32) % 6;
        if (ct != c0 && ct < G) {
          const int r = comp < 3 ? 0 : (comp < 5 ? 1 : 2), c = comp < 3 ? comp : (comp < 5 ? comp - 2 : 2);
          double s = 0;
          for (int e = P.xinc_ptr[c0]; e < P.xinc_ptr[c0 + 1]; e++) {
            const int a = P.xinc_idx[e];
            if (s_cid[a - ab] != ct) continue;
            if (P.pt_fixed && P.pt_fixed[P.inc_row[a]]) continue;  // rows without unknowns are not in the aggregate
            const double* cf = s_coef + 4 * (size_t)(a - ab);
            s -= (r == c ? cf[0] : 0.0) + cf[1 + r] * cf[1 + c];
          }
          stage[ct * 6 + comp] = (float)s;
        }
      }
      __syncthreads();
      for (int t = tid; t < 16 * kRowBlk; t += nthr) {
        const int target = t / kRowBlk, k = t % kRowBlk;
        if (target < G) cluster.map_shared_rank(s_Aall, target)[(size_t)my_rank * kRowBlk + k] = stage[k];
      }
      const long long ts4 = clock64();
      prof[10] += ts4 - ts3;  // setup: coarse row block + push
      barrier();
      const long long ts5 = clock64();
      prof[11] += ts5 - ts4;  // setup: coarse exchange barrier
      // assemble (upper triangle mirrored so that the matrix is exactly symmetric)
      for (int t = tid; t < kCoarseN * (kCoarseS - kCoarseN); t += nthr)
        s_Ac[(size_t)(t / (kCoarseS - kCoarseN)) * kCoarseS + kCoarseN + t % (kCoarseS - kCoarseN)] = 0.f;
      for (int t = tid; t < kCoarseN * kCoarseN; t += nthr) {
        int I = t / kCoarseN, J = t % kCoarseN;
        if (I > J) { const int q = I; I = J; J = q; }
        float v;
        if (J < 6) {
          v = pos ? (float)(s_H[sym6(I, J)] + (I == J ? lambda : 0.0)) : (I == J ? 1.f : 0.f);
        } else {
          const int cj = (J - 6) / 3, bj = (J - 6) % 3;
          const bool live_j = cj < G && s_Aall[(size_t)cj * kRowBlk + 114] > 0.5f;
          if (I < 6) {
            v = (pos && live_j) ? s_Aall[(size_t)cj * kRowBlk + 96 + I * 3 + bj] : 0.f;
          } else {
            const int ci = (I - 6) / 3, bi = (I - 6) % 3;
            const bool live_i = ci < G && s_Aall[(size_t)ci * kRowBlk + 114] > 0.5f;
            if (live_i && live_j) {
              const int r = bi < bj ? bi : bj, c = bi < bj ? bj : bi;
              const int comp = r == 0 ? c : (r == 1 ? 2 + c : 5);
              v = s_Aall[(size_t)ci * kRowBlk + cj * 6 + comp];
            } else {
              v = (I == J) ? 1.f : 0.f;  // aggregate without unknowns (or beyond the cluster): identity row
            }
          }
        }
        s_Ac[(size_t)(t / kCoarseN) * kCoarseS + (t % kCoarseN)] = v;
      }
      __syncthreads();
      // symmetric diagonal scaling D^-1/2 A D^-1/2 (the pose block is ~1e9, the aggregates ~1e7, lambda ~1e2: without
      // it the fp32 sweep loses the small pivots); the scales go where the row blocks were (no longer needed)
      float* s_ds = s_Aall;
      if (tid < kCoarseN) {
        const float d = s_Ac[tid * kCoarseS + tid];
        s_ds[64 + tid] = d > 0.f ? rsqrtf(d) : 0.f;
      }
      __syncthreads();
      if (tid < kCoarseN) s_ds[tid] = s_ds[64 + tid];
      for (int t = tid; t < kCoarseN * kCoarseN; t += nthr) {
        const int I = t / kCoarseN, J = t % kCoarseN;
        s_Ac[I * kCoarseS + J] *= s_ds[64 + I] * s_ds[64 + J];
      }
      __syncthreads();
      const long long ts6 = clock64();
      prof[12] += ts6 - ts5;  // setup: coarse assembly + scaling
      // sweep every pivot: A -> -A^-1. Rows are split over the warps, lanes own columns; a pivot that is not safely
      // positive disables the coarse level for this solve (identical decision in every CTA: identical inputs)
      if (tid == 0) *s_flag = 0;
      __syncthreads();
      {
        float* M = s_Ac;
        for (int j = 0; j < kCoarseN; j++) {
          const float d = M[j * kCoarseS + j];
          if (!(d > 1e-5f)) {
            if (tid == 0) *s_flag = 1;
            break;  // uniform: every thread reads the same pivot
          }
          const float id = 1.f / d;
          const int k0 = lane, k1 = lane + 32;
          const float pj0 = M[j * kCoarseS + k0];
          const float pj1 = (k1 < kCoarseN) ? M[j * kCoarseS + k1] : 0.f;
          float fcol[(kCoarseN + 7) / 8];  // this warp's rows: the pivot-column entries, read before anything changes
#pragma unroll
          for (int q = 0; q < (kCoarseN + 7) / 8; q++) {
            const int i = warp + nw * q;
            fcol[q] = (i < kCoarseN) ? M[i * kCoarseS + j] * id : 0.f;
          }
          __syncthreads();
#pragma unroll
          for (int q = 0; q < (kCoarseN + 7) / 8; q++) {
            const int i = warp + nw * q;
            if (i < kCoarseN && i != j) {
              if (k0 != j) M[i * kCoarseS + k0] -= fcol[q] * pj0;
              if (k1 < kCoarseN && k1 != j) M[i * kCoarseS + k1] -= fcol[q] * pj1;
              if (lane == 0) M[i * kCoarseS + j] = fcol[q];  // pivot column: a_ij / d
            }
          }
          if (warp == 0) {  // pivot row: a_ji / d ; pivot: -1 / d   (pj0 / pj1 hold the old row)
            if (k0 != j) M[j * kCoarseS + k0] = pj0 * id;
            if (k1 < kCoarseN && k1 != j) M[j * kCoarseS + k1] = pj1 * id;
            if (lane == 0) M[j * kCoarseS + j] = -id;
          }
          __syncthreads();
        }
        for (int t = tid; t < kCoarseN * kCoarseN; t += nthr) {
          const int i = t / kCoarseN, k = t % kCoarseN;
          if (k > i) M[i * kCoarseS + k] = M[k * kCoarseS + i];
        }
      }
      __syncthreads();
      prof[13] += clock64() - ts6;  // setup: coarse sweep
      use_coarse = *s_flag == 0;
      if (!use_coarse && pos) {  // fall back to the exact 6x6 pose block for this solve
        __syncthreads();
        if (tid == 0) {
          *s_flag = 0;
          if (!invert6(s_H, lambda, s_M)) *s_flag = 1;
        }
        __syncthreads();
        if (*s_flag) {
          barrier();
          return false;
        }
      }
    }

    // ---- z = M^-1 r is finished in two places (initial residual, every iteration): block reduction of r.t and of the
    // aggregate's residual sum, halo + value push, cluster barrier, coarse solve (warp 0), correction of z.
    // Returns r.z. rsum: this thread's sum of residual components of its (non-fixed) rows.
    auto finish_z = [&](int buf, double rz_part, double rs0, double rs1, double rs2) -> double {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) rz_part += __shfl_xor_sync(0xffffffffu, rz_part, off);
      if (use_coarse) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          rs0 += __shfl_xor_sync(0xffffffffu, rs0, off);
          rs1 += __shfl_xor_sync(0xffffffffu, rs1, off);
          rs2 += __shfl_xor_sync(0xffffffffu, rs2, off);
        }
      }
      if (lane == 0) {
        s_red[warp] = rz_part;
        if (use_coarse) {
          s_red[96 + 3 * warp] = rs0;
          s_red[96 + 3 * warp + 1] = rs1;
          s_red[96 + 3 * warp + 2] = rs2;
        }
      }
      __syncthreads();  // S4
      push_halo();
      if (warp == 0) {
        double t = 0;
        for (int w = lane; w < nw; w += 32) t += s_red[w];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) s_slot[7] = t;
        if (use_coarse && lane < 3) {
          double v = 0;
          for (int w = 0; w < nw; w++) v += s_red[96 + 3 * w + lane];
          s_slot[8 + lane] = v;
        }
        push(buf, 7, use_coarse ? 4 : 1);
      }
      barrier();  // B2
      if (warp == 0) {
        double t = 0;
        if (lane < G) t = s_gather[(size_t)(buf * 16 + lane) * kGatherVals + 7];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (use_coarse) {
          // coarse residual [pose ; aggregate sums] (scaled), y = D^-1/2 (D^-1/2 A_c D^-1/2)^-1 D^-1/2 r_c; the sweep
          // left minus the inverse
          const float* s_ds = s_Aall;
          double rfull[2] = {0, 0};
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int j = lane + 32 * q;
            if (j < kCoarseN) {
              if (j < 6)
                rfull[q] = pos ? s_rp[j] : 0.0;
              else if ((j - 6) / 3 < G)
                rfull[q] = s_gather[(size_t)(buf * 16 + (j - 6) / 3) * kGatherVals + 8 + (j - 6) % 3];
              s_rcv[j] = (float)rfull[q] * s_ds[j];
            } else if (j < kCoarseS) {
              s_rcv[j] = 0.f;
            }
          }
          __syncwarp();
          double dotp = 0;
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int j = lane + 32 * q;
            if (j < kCoarseN) {
              const float4* Mr = reinterpret_cast<const float4*>(s_Ac + (size_t)j * kCoarseS);
              const float4* rv = reinterpret_cast<const float4*>(s_rcv);
              float a0 = 0.f, a1 = 0.f;
#pragma unroll
              for (int k = 0; k < (kCoarseN + 3) / 4; k++) {
                const float4 m = Mr[k], r = rv[k];
                a0 = fmaf(m.x, r.x, a0);
                a1 = fmaf(m.y, r.y, a1);
                a0 = fmaf(m.z, r.z, a0);
                a1 = fmaf(m.w, r.w, a1);
              }
              const double y = -(double)((a0 + a1) * s_ds[j]);
              s_y[j] = y;
              if (pos && j < 6) s_zp[j] = y;
              dotp += rfull[q] * y;  // r.z gains (aggregate residual).(correction)
            }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) dotp += __shfl_xor_sync(0xffffffffu, dotp, off);
          t += dotp;
        } else if (pos) {
          double rzp = 0;
          for (int a = 0; a < 6; a++) rzp += s_rp[a] * s_zp[a];
          t += rzp;
        }
        if (lane == 0) s_bc[1] = t;
      }
      __syncthreads();  // S5
      if (use_coarse) {
        // z += Z y on the own rows and on the halo copies (fixed rows keep z = 0)
        for (int t = tid; t < 3 * nrows; t += nthr) {
          const int lr = t / 3, c = t % 3;
          if (!(P.pt_fixed && P.pt_fixed[rb + lr])) s_z[t] += s_y[6 + 3 * c0 + c];
        }
        for (int t = tid; t < 3 * P.halo_rows; t += nthr) {
          const int cid = s_hcid[t / 3];
          if (cid != 255) s_halo[t] += s_y[6 + 3 * cid + t % 3];
        }
        __syncthreads();  // S6
      }
      return s_bc[1];
    };

    // ---- initial residual, t = M_B^-1 r (one thread per row)
    double rz_part = 0, rs0 = 0, rs1 = 0, rs2 = 0;
    if (tid < nrows) {
      const int li = tid, i = rb + tid;
      if (P.pt_fixed && P.pt_fixed[i]) {
        if (!bprec) {
          double2* mo = reinterpret_cast<double2*>(s_minv + 8 * (size_t)li);
          mo[0] = mo[1] = mo[2] = make_double2(0.0, 0.0);
        }
        st3(s_r, li, V3{0, 0, 0});
        st3(s_x, li, V3{0, 0, 0});
        st3s(s_z, li, V3{0, 0, 0});
        if (bprec) s_rf[3 * li] = s_rf[3 * li + 1] = s_rf[3 * li + 2] = 0.f;
      } else {
        const V3 r = ld3p(P.bvec, i);
        st3(s_r, li, r);
        st3(s_x, li, V3{0, 0, 0});
        rs0 = r.x;
        rs1 = r.y;
        rs2 = r.z;
        if (bprec) {
          s_rf[3 * li] = (float)r.x;
          s_rf[3 * li + 1] = (float)r.y;
          s_rf[3 * li + 2] = (float)r.z;
        } else {
          const double2* d = reinterpret_cast<const double2*>(P.dg + 8 * (size_t)i);
          const double2 d0 = d[0], d1 = d[1], d2 = d[2];
          const double a = d0.x + lambda, b = d0.y, c = d1.x, e = d1.y + lambda, f = d2.x, g = d2.y + lambda;
          const double C00 = e * g - f * f, C01 = c * f - b * g, C02 = b * f - c * e;
          const double det = a * C00 + b * C01 + c * C02;
          const double id = 1.0 / det;
          const double m00 = C00 * id, m01 = C01 * id, m02 = C02 * id;
          const double m11 = (a * g - c * c) * id, m12 = (b * c - a * f) * id, m22 = (a * e - b * b) * id;
          double2* mo = reinterpret_cast<double2*>(s_minv + 8 * (size_t)li);
          mo[0] = make_double2(m00, m01);
          mo[1] = make_double2(m02, m11);
          mo[2] = make_double2(m12, m22);
          const V3 z{m00 * r.x + m01 * r.y + m02 * r.z, m01 * r.x + m11 * r.y + m12 * r.z,
                     m02 * r.x + m12 * r.y + m22 * r.z};
          st3s(s_z, li, z);
          rz_part += r.x * z.x + r.y * z.y + r.z * z.z;
        }
      }
    }
    if (bprec) {
      __syncthreads();
      rz_part += prec_apply_quads(rb, re, false);
    }
    if (pos && tid < 6) {
      s_rp[tid] = s_bp[tid];
      s_xp[tid] = 0;
      s_pp[tid] = 0;
      s_qp[tid] = 0;
      if (!use_coarse) {
        double s = 0;
        for (int c = 0; c < 6; c++) s += s_M[tid * 6 + c] * s_bp[c];
        s_zp[tid] = s;
      }
    }
    const long long ts7 = clock64();
    double rz = finish_z(1, rz_part, rs0, rs1, rs2);
    prof[14] += clock64() - ts7;  // setup: first z exchange
    const double rz0 = rz;
    if (!(rz0 > 0)) {  // b == 0: delta = 0
      if (tid < nrows) st3(P.xcg, rb + tid, V3{0, 0, 0});
      if (pos && tid < 6) s_xp[tid] = 0;
      barrier();  // nobody leaves while its buffers may still be written
      return isfinite(rz0);
    }
    const double stop = P.pcg_tol * P.pcg_tol * rz0;
    double beta = 0;
    bool ok = true;
    int it = 0;
    const int qr = tid / kTPR, ql = tid % kTPR;
    const bool valid = qr < nrows;
    const int i = rb + qr, li = qr;
    bool fixed = true;
    int kf = -1, a_beg = 0, a1 = 0;
    double su = 0;
    if (valid) {
      fixed = P.pt_fixed && P.pt_fixed[i];
      kf = P.pt_kf[i];
      a_beg = P.inc_ptr[i];
      a1 = P.inc_ptr[i + 1];
      su = P.dg[8 * (size_t)i + 6];
    }
    for (; it < P.pcg_max_iter; it++) {
      const bool first = (it == 0);
      // two gather buffers (iteration parity): entries [0..6] carry phase 1, [7..10] phase 2; a buffer is rewritten
      // only after two further cluster barriers, when every reader has moved on
      const int par = it & 1;
      // ---- phase 1: w = (H + lambda I) z ; p = z + beta p ; q = w + beta q ; partial p.q
      const long long tm0 = clock64();
      double pq_part = 0;
      double w0 = 0, w1 = 0, w2 = 0;
      double redw[6] = {0, 0, 0, 0, 0, 0};  // this row's pose partial (lane 0 of the row's lane group)
      V3 zi{0, 0, 0};
      if (valid) {
        zi = ld3s(s_z, li);
        if (!fixed) {
          for (int a0 = a_beg + ql; a0 < a1; a0 += 3 * kTPR) {
            const double* zp[3];
            const double2* cp[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const int a = min(a0 + kTPR * k, a1 - 1);
              cp[k] = reinterpret_cast<const double2*>(s_coef) + 2 * (a - ab);
              zp[k] = s_zptr[a - ab];
              __builtin_assume(__isShared(zp[k]));  // own rows or the pushed halo: never a remote address
            }
            double2 c0v[3], c1v[3];
            double zx[3], zy[3], zb[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
              c0v[k] = cp[k][0];
              c1v[k] = cp[k][1];
              zx[k] = zp[k][0];
              zy[k] = zp[k][1];
              zb[k] = zp[k][2];
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
              if (a0 + kTPR * k < a1) {
                const double dx = zi.x - zx[k], dy = zi.y - zy[k], dz = zi.z - zb[k];
                const double ud = c0v[k].y * dx + c1v[k].x * dy + c1v[k].y * dz;
                w0 += c0v[k].x * dx + c0v[k].y * ud;
                w1 += c0v[k].x * dy + c1v[k].x * ud;
                w2 += c0v[k].x * dz + c1v[k].y * ud;
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = 1; o < kTPR; o <<= 1) {
        w0 += __shfl_xor_sync(0xffffffffu, w0, o);
        w1 += __shfl_xor_sync(0xffffffffu, w1, o);
        w2 += __shfl_xor_sync(0xffffffffu, w2, o);
      }
      if (valid && ql == 0) {
        double q0 = w0 + (lambda + su) * zi.x, q1 = w1 + (lambda + su) * zi.y, q2 = w2 + (lambda + su) * zi.z;
        if (kf >= 0) {
          const double2* jo = reinterpret_cast<const double2*>(s_jac + kJS * (size_t)li);
          const double omega = jo[9].x;
          if (omega != 0) {
            double A[12], B[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
              const double2 t = jo[k];
              A[2 * k] = t.x;
              A[2 * k + 1] = t.y;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const double2 t = jo[6 + k];
              B[2 * k] = t.x;
              B[2 * k + 1] = t.y;
            }
            double jp0 = 0, jp1 = 0;
            if (pos) {
#pragma unroll
              for (int a = 0; a < 6; a++) {
                jp0 += A[a] * s_zp[a];
                jp1 += A[6 + a] * s_zp[a];
              }
            }
            const double t0 = omega * (jp0 + B[0] * zi.x + B[1] * zi.y + B[2] * zi.z);
            const double t1 = omega * (jp1 + B[3] * zi.x + B[4] * zi.y + B[5] * zi.z);
            if (!fixed) {
              q0 += B[0] * t0 + B[3] * t1;
              q1 += B[1] * t0 + B[4] * t1;
              q2 += B[2] * t0 + B[5] * t1;
            }
            if (pos) {
#pragma unroll
              for (int a = 0; a < 6; a++) redw[a] = A[a] * t0 + A[6 + a] * t1;
            }
          }
        }
        if (!fixed) {
          V3 pn = zi, qn{q0, q1, q2};
          if (!first) {
            const V3 po = ld3p(s_p, li), qo = ld3p(s_q, li);
            pn = V3{zi.x + beta * po.x, zi.y + beta * po.y, zi.z + beta * po.z};
            qn = V3{q0 + beta * qo.x, q1 + beta * qo.y, q2 + beta * qo.z};
          }
          st3(s_p, li, pn);
          st3(s_q, li, qn);
          pq_part = pn.x * qn.x + pn.y * qn.y + pn.z * qn.z;
        }
      }
      if (pos) {
        // pose partials: every warp reduces its own rows by shuffle (lanes that own no row carry zeros); warp 0 only
        // adds the per-warp sums after the barrier — keeps the serial section in front of the cluster barrier short
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
          for (int off = 16; off >= kTPR; off >>= 1) redw[a] += __shfl_xor_sync(0xffffffffu, redw[a], off);
        }
        if (lane == 0) {
#pragma unroll
          for (int a = 0; a < 6; a++) s_red[32 + 6 * warp + a] = redw[a];
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) pq_part += __shfl_xor_sync(0xffffffffu, pq_part, off);
      if (lane == 0) s_red[warp] = pq_part;
      __syncthreads();  // S1
      if (warp == 0) {
        double t = 0;
        for (int w = lane; w < nw; w += 32) t += s_red[w];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) s_slot[0] = t;
        if (pos && lane < 6) {
          double v = 0;
          for (int w = 0; w < nw; w++) v += s_red[32 + 6 * w + lane];
          s_slot[1 + lane] = v;
        }
        push(par, 0, pos ? 7 : 1);
      }
      const long long tm1 = clock64();
      barrier();  // B1
      if (warp == 0) {
        const double* rs = s_gather + (size_t)(par * 16 + (lane < G ? lane : 0)) * kGatherVals;
        double t[7];
#pragma unroll
        for (int k = 0; k < 7; k++) t[k] = (lane < G && (k == 0 || pos)) ? rs[k] : 0.0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], off);
        }
        double pq = t[0];
        if (pos) {
          // pose rows: w_p = lambda z_p + sum_i A_i^T t_i ; p_p, q_p by the same recurrences (lanes 0..5)
          double ppv = 0, qpv = 0;
          if (lane < 6) {
            double wsum = t[1];
#pragma unroll
            for (int k = 2; k < 7; k++)
              if (lane == k - 1) wsum = t[k];
            double w = lambda * s_zp[lane] + wsum;
            // H_pp z_p is part of the row sums already (every row adds A^T omega A z_p)
            ppv = first ? s_zp[lane] : s_zp[lane] + beta * s_pp[lane];
            qpv = first ? w : w + beta * s_qp[lane];
            s_pp[lane] = ppv;
            s_qp[lane] = qpv;
          }
          double d = ppv * qpv;
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          pq += __shfl_sync(0xffffffffu, d, 0);
        }
        if (lane == 0) s_bc[0] = pq;
      }
      __syncthreads();  // S2
      const long long tm2 = clock64();
      const double pq = s_bc[0];
      if (!(pq > 0) || !isfinite(pq)) {
        ok = false;
        break;
      }
      const double alpha = rz / pq;
      // ---- phase 2: x += alpha p ; r -= alpha q ; t = M_B^-1 r ; partial r.t and aggregate residual sums
      double rzn_part = 0;
      rs0 = rs1 = rs2 = 0;
      const bool act = valid && !fixed;
      if (act) {
        for (int cmp = ql; cmp < 3; cmp += kTPR) {
          const size_t o = 4 * (size_t)li + cmp;
          s_x[o] += alpha * s_p[o];
          const double rc = s_r[o] - alpha * s_q[o];
          s_r[o] = rc;
          if (bprec) s_rf[3 * li + cmp] = (float)rc;
          if (cmp == 0) rs0 = rc;
          else if (cmp == 1) rs1 = rc;
          else rs2 = rc;
        }
      }
      if (pos && tid < 6) {  // pose rows (replicated): x, r
        s_xp[tid] += alpha * s_pp[tid];
        s_rp[tid] -= alpha * s_qp[tid];
      }
      if (!bprec) {
        __syncwarp();  // the lanes of the quad wrote one component each
        if (act && ql == 0) {
          const V3 ri = ld3p(s_r, li);
          const double2* mo = reinterpret_cast<const double2*>(s_minv + 8 * (size_t)li);
          const double2 m0 = mo[0], m1 = mo[1], m2 = mo[2];
          const V3 z{m0.x * ri.x + m0.y * ri.y + m1.x * ri.z, m0.y * ri.x + m1.y * ri.y + m2.x * ri.z,
                     m1.x * ri.x + m2.x * ri.y + m2.y * ri.z};
          st3s(s_z, li, z);
          rzn_part = ri.x * z.x + ri.y * z.y + ri.z * z.z;
        }
        __syncthreads();  // S3 (pose r complete)
      } else {
        __syncthreads();  // S3 (rf and pose r complete)
        rzn_part = prec_apply_quads(rb, re, false);
      }
      if (pos && !use_coarse && tid < 6) {
        double s = 0;
#pragma unroll
        for (int c = 0; c < 6; c++) s += s_M[tid * 6 + c] * s_rp[c];
        s_zp[tid] = s;
      }
      const long long tm3 = clock64();
      const double rzn = finish_z(par, rzn_part, rs0, rs1, rs2);
      const long long tm4 = clock64();
      prof[1] += tm1 - tm0;  // matvec pass
      prof[2] += tm2 - tm1;  // pq exchange
      prof[3] += tm3 - tm2;  // update pass
      prof[4] += tm4 - tm3;  // z exchange (+ coarse level)
      if (!isfinite(rzn)) {
        ok = false;
        it++;
        break;
      }
      beta = rzn / rz;
      rz = rzn;
      if (rz <= stop) {
        it++;
        break;
      }
    }
    pcg_iters += it;
    if (tid < nrows) st3(P.xcg, rb + tid, ld3p(s_x, tid));
    barrier();  // every remote write into this CTA's shared memory has landed; s_xp is complete
    return ok;
  }

  // ================================================================================================
  // One LM iteration. Returns true when the optimisation must terminate (g2o "Terminate").
  // ================================================================================================
  __device__ bool lm_iteration(int iteration) {
    const int F = P.F;
    const bool pts = !P.points_fixed, pos = !P.poses_fixed;
    const long long tl0 = clock64();
    if (pts) {  // estimates written by other CTAs / ranks (restore, reset) are visible
      if ((SHARD && P.world > 1)) xsync(); else barrier();
    }
    double acc[2] = {0, 0};  // chi2, max diagonal
    if (P.P > 0 || P.D > 0) {
      edges_pass<true>(acc[0]);
      barrier();
    }
    const int par = gen & 1;
    rows_pass<true>(acc[0], acc[1], par);
    grid_reduce<2>(acc, 2u, pos ? 2 : 0);
    double currentChi = s_scal[0];
    double maxDiag = s_scal[1];
    n_sweeps++;
    if (pos) {
      for (int t = tid; t < 27 * F; t += nthr) {
        const int k = t / 27, v = t % 27;
        const double s = ((SHARD && P.world > 1)) ? xextra(t) : sum_chunk_partials(k, v, par);
        if (v < 21)
          s_H[21 * k + v] = s;
        else
          s_bp[6 * k + v - 21] = s;
      }
      __syncthreads();
      if (iteration == 0)
        for (int k = 0; k < F; k++)
          for (int a = 0; a < 6; a++) maxDiag = fmax(maxDiag, fabs(s_H[21 * k + sym6(a, a)]));
    }
    if (iteration == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:153-165
      lambda = P.lm_tau * maxDiag;
      ni = 2;
    }
    prof[5] += clock64() - tl0;  // linearisation
    double rho = 0;
    int qmax = 0;
    do {
      const long long ts0 = clock64();
      bool solved;
      if constexpr (WIDE) {
        solved = pcg_wide();
      } else {
        const bool native = P.cluster_mode && P.resident && P.F == 1 && P.D == 0 && !P.points_fixed &&
                            P.n_chunks == (int)gridDim.x && !P.no_dsmem && P.push_ptr != nullptr;
        solved = native ? pcg_cluster() : pcg();
      }
      const long long ts1 = clock64();
      prof[6] += ts1 - ts0;  // solve
      lm_trials++;
      if (!solved) pcg_fail++;
      // push + update (sparse_optimizer.cpp:457-470), scale = delta^T (lambda delta + b)
      double acc2[2] = {0, 0};
      if (solved) {
        if (pts) {
          for_rows([&](int i) {
            const V3 d = ld3p(P.xcg, i), bb = ld3p(P.bvec, i);
            V3 x = ld3p(P.x, i);
            st3(P.x_bak, i, x);
            x.x += d.x; x.y += d.y; x.z += d.z;
            st3(P.x, i, x);
            if ((SHARD && P.world > 1)) xpush3(P.xx, i, x);
            acc2[1] += d.x * (lambda * d.x + bb.x) + d.y * (lambda * d.y + bb.y) + d.z * (lambda * d.z + bb.z);
          });
        }
        if (pos) {
          for (int t = tid; t < 7 * F; t += nthr) s_pose_bak[t] = s_pose[t];
          __syncthreads();
          for (int k = tid; k < F; k += nthr) pose_oplus(s_pose + 7 * k, s_xp + 6 * k);
          __syncthreads();
        }
        if (pts) {  // updated point estimates are read across CTAs (and ranks) by the regulariser edges
          if ((SHARD && P.world > 1)) xsync(); else barrier();
        }
        if (P.P > 0 || P.D > 0) edges_pass<false>(acc2[0]);
        double dummy = 0;
        rows_pass<false>(acc2[0], dummy, 0);
        n_chi2++;
      }
      grid_reduce<2>(acc2, 0);
      double tempChi = solved ? s_scal[0] : DBL_MAX;
      double scale = s_scal[1];
      if (pos && solved)
        for (int t = 0; t < 6 * F; t++) scale += s_xp[t] * (lambda * s_xp[t] + s_bp[t]);
      scale += 1e-3;
      rho = (currentChi - tempChi) / scale;
      if (rho > 0 && isfinite(tempChi)) {
        const double t3 = 2 * rho - 1;
        double alpha = 1. - t3 * t3 * t3;
        alpha = fmin(alpha, 2. / 3.);
        const double scaleFactor = fmax(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        if (solved) {  // pop
          if (pts)
            for_rows([&](int i) {
              const V3 xb = ld3p(P.x_bak, i);
              st3(P.x, i, xb);
              if ((SHARD && P.world > 1)) xpush3(P.xx, i, xb);
            });
          if (pos) {
            __syncthreads();
            for (int t = tid; t < 7 * F; t += nthr) s_pose[t] = s_pose_bak[t];
            __syncthreads();
          }
        }
        if (!isfinite(lambda)) break;
      }
      prof[7] += clock64() - ts1;  // update + chi2
      qmax++;
    } while (rho < 0 && qmax < P.lm_max_trials);
    lm_iters++;
    if (blockIdx.x == 0 && tid == 0 && n_trace < kTrace) P.stats->chi2_trace[n_trace] = currentChi;
    n_trace++;
    return (qmax == P.lm_max_trials || rho == 0 || !isfinite(lambda));
  }

  __device__ void run() {
    const int F = P.F;
    const long long trun0 = clock64();
    for (int t = tid; t < 7 * F; t += nthr) s_pose[t] = P.pose[t];
    __syncthreads();
    for (int o = 0; o < P.n_ops; o++) {
      const int op = P.op[o], arg = P.op_arg[o];
      switch (op) {
        case OP_RESET: {
          barrier();
          for_rows([&](int i) { st3(P.x, i, ld3c(P.x_seed, i)); });
          __syncthreads();
          if (P.seed_via_f32) {
            if (tid == 0) {  // the seed is an earlier launch's fp64 result crossing the Frame as Sophus::SE3f
              for (int k = 0; k < F; k++) {
                double sd[7];
                float sf[7];
                for (int t = 0; t < 7; t++) sd[t] = __ldcg(P.pose_seed + 7 * k + t);
                pose_to_f7(sd, sf);
                pose_from_f7(sf, s_pose + 7 * k);
              }
            }
          } else {
            for (int t = tid; t < 7 * F; t += nthr) s_pose[t] = P.pose_seed[t];
          }
          __syncthreads();
        } break;
        case OP_CLEAR_LEVELS: {
          barrier();
          const int gsz = gridDim.x * nthr;
          for (int i = blockIdx.x * nthr + tid; i < P.V; i += gsz) P.rp_level[i] = 0;
          for (int e = blockIdx.x * nthr + tid; e < P.P; e += gsz) P.sp_level[e] = 0;
        } break;
        case OP_OPTIMIZE: {
          for (int it = 0; it < arg; it++)
            if (lm_iteration(it)) break;
        } break;
        case OP_RELEVEL_POSE: {
          // g2o_optimization.cc:113-134 — inlier edges keep the error of the last evaluated trial (stale after a
          // rejected step), outlier edges are re-evaluated at the final pose.
          barrier();
          for_rows([&](int i) {
            const int kf = P.pt_kf[i];
            if (kf < 0) return;
            double c2;
            if (P.rp_level[i] == 0) {
              c2 = P.rp_chi2[i];
            } else {
              double pc[3], err[2];
              c2 = reproj_error(i, kf, ld3(P.x, i), pc, err);
              P.rp_chi2[i] = c2;
            }
            P.rp_level[i] = ((float)c2 > P.th2f) ? 1 : 0;
          });
        } break;
        case OP_RELEVEL_DEFORM: {
          // g2o_optimization.cc:352-394 — reprojection edges by chi2 > 5.99; every spatial edge ends on its own
          // chi2 > 0.584 test.
          barrier();
          for_rows([&](int i) {
            const int kf = P.pt_kf[i];
            if (kf < 0) return;
            double pc[3], err[2];
            const double c2 = reproj_error(i, kf, ld3(P.x, i), pc, err);
            P.rp_chi2[i] = c2;
            P.rp_level[i] = ((float)c2 > P.th2f) ? 1 : 0;
          });
          const int gsz = gridDim.x * nthr;
          for (int e = blockIdx.x * nthr + tid; e < P.P; e += gsz) {
            const double w = P.pair_w[e];
            if (w < 0) continue;
            const V3 xi = ld3(P.x, P.pair_i[e]), xj = ld3(P.x, P.pair_j[e]);
            const double e0 = w * (xi.x - xj.x), e1 = w * (xi.y - xj.y), e2 = w * (xi.z - xj.z);
            const double c2 = (e0 * e0 + e1 * e1 + e2 * e2) * P.info_spatial;
            P.sp_level[e] = (c2 > (double)P.th3f) ? 1 : 0;
          }
        } break;
        case OP_FINAL_CHI2: {
          barrier();
          for_rows([&](int i) {
            const int kf = P.pt_kf[i];
            if (kf < 0) return;
            double pc[3], err[2];
            P.rp_chi2[i] = reproj_error(i, kf, ld3(P.x, i), pc, err);
          });
        } break;
        default: break;
      }
    }
    __syncthreads();
    if ((SHARD && P.world > 1)) xsync();  // no push of this launch is in flight when any rank's kernel ends
    if (blockIdx.x == 0) {
      for (int t = tid; t < 7 * F; t += nthr) P.pose[t] = s_pose[t];
      if (tid == 0) {
        EngineStats* st = P.stats;
        st->lm_iterations = lm_iters;
        st->lm_trials = lm_trials;
        st->pcg_iterations = pcg_iters;
        st->n_sweeps = n_sweeps;
        st->n_chi2_passes = n_chi2;
        st->n_trace = n_trace < kTrace ? n_trace : kTrace;
        st->pcg_fail = pcg_fail;
        st->barriers = (int)gen;
        st->xepochs = (int)xe;
        st->xfail = xdead ? 1 : 0;
        st->lambda_final = lambda;
        prof[15] = clock64() - trun0;
        for (int i = 0; i < 16; i++)
          if (!(WIDE || SHARD) || i < 8 || i > 11) st->prof[i] = prof[i];
      }
    }
    if (WIDE && !SHARD && tid == 0) {  // slowest / fastest CTA in the matvec pass, slowest in the update pass (diagnostics)
      atomicMax(reinterpret_cast<long long*>(&P.stats->prof[8]), prof[1]);
      atomicMax(reinterpret_cast<long long*>(&P.stats->prof[9]), prof[3]);
      atomicMax(reinterpret_cast<long long*>(&P.stats->prof[10]), -prof[1]);
    }
    // a cluster must not retire CTAs while others may still arrive at the hardware barrier
    if (P.cluster_mode) barrier();
  }
};

__global__ void __launch_bounds__(kMaxBlock, 1) nrs_lm_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) double nrs_smem[];
  Engine<false, false> eng(p, nrs_smem);
  eng.run();
}

__global__ void __launch_bounds__(kMaxBlock, 2) nrs_lm_kernel_wide(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) double nrs_smem[];
  Engine<true, false> eng(p, nrs_smem);
  eng.run();
}

// the same two for a rank of a landmark-sharded BA
__global__ void __launch_bounds__(kMaxBlock, 1) nrs_lm_kernel_shard(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) double nrs_smem[];
  Engine<false, true> eng(p, nrs_smem);
  eng.run();
}

__global__ void __launch_bounds__(kMaxBlock, 2) nrs_lm_kernel_wide_shard(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) double nrs_smem[];
  Engine<true, true> eng(p, nrs_smem);
  eng.run();
}

}  // namespace

size_t engine_smem_bytes(int F, int res_rows, int res_inc, int block_prec) {
  size_t d = (size_t)F * (7 + 7 + 21 + 36 + 6 * 6) + 32 * kChunkVals + 32 + 6 * kMaxRows + 2 * 16 * kGatherVals + 2 + 6;  // + alignment slack
  if (res_rows == 0) d += 16 * kMaxRows;
  if (res_rows > 0) {
    d += std::max((size_t)kJS * res_rows, 16 * (size_t)kMaxRows) + (size_t)res_rows * (4 * 4 + 3 + (block_prec ? 0 : 8)) +
         5 * (size_t)res_inc + (3 * (size_t)res_rows + 1) / 2 + 6;
    size_t bytes = d * sizeof(double);
    if (block_prec) bytes += (size_t)((res_rows + kPrecBlock - 1) / kPrecBlock) * kPN * kPS * sizeof(float);
    return bytes + 16;
  }
  return d * sizeof(double) + 16;
}

size_t engine_smem_bytes_wide(int F) {
  const size_t d = (size_t)F * (7 + 7 + 36 + 6 * 6) + 32 * kChunkVals + 32 + 2 + 6 +
                   std::max((size_t)21 * F, (size_t)16 * kMaxRows);
  return d * sizeof(double) + 16;
}

size_t engine_smem_extra(int res_inc, int halo_rows, int coarse) {
  size_t b = sizeof(double) * ((3 * (size_t)halo_rows + 1) & ~(size_t)1);
  if (coarse)
    b += sizeof(float) * (16 * kRowBlk + kCoarseN * kCoarseS + kCoarseS + 2 * kCoarseS) + (((size_t)res_inc + 15) & ~(size_t)15) +
         (((size_t)halo_rows + 15) & ~(size_t)15) + 64;
  return b;
}

static const void* kernel_of(int wide, int shard = 0) {
  if (shard) return wide ? (const void*)nrs_lm_kernel_wide_shard : (const void*)nrs_lm_kernel_shard;
  return wide ? (const void*)nrs_lm_kernel_wide : (const void*)nrs_lm_kernel;
}

// The attribute belongs to the function in the CURRENT device's context, and several contexts on different devices
// may live in one process (opt.device): it is set on every call instead of being cached process-wide (a cache made a
// second device miss it; the call costs microseconds and is thread safe).
static bool set_smem(size_t smem, int wide = 0, int shard = 0) {
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kernel_of(wide, shard), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
  }
  return true;
}

int engine_max_grid(int block, size_t smem, int wide) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!set_smem(smem, wide)) return 0;
  const cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrs_lm_kernel_wide, block, smem)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nrs_lm_kernel, block, smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return sms * per_sm;
}

int engine_max_cluster(int block, size_t smem) {
  if (!set_smem(smem)) return 0;
  cudaFuncSetAttribute(nrs_lm_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaGetLastError();
  for (int cs = 16; cs >= 2; cs >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, nrs_lm_kernel, &cfg) == cudaSuccess && n >= 1) return cs;
    cudaGetLastError();
  }
  return 0;
}

int launch_engine(const Params& p, int grid, int block, size_t smem, cudaStream_t stream) {
  const int wide = (p.wide && !p.cluster_mode) ? 1 : 0;
  const int shard = (p.world > 1 && !p.cluster_mode) ? 1 : 0;  // a sharded rank never runs in cluster mode
  if (!set_smem(smem, wide, shard)) return (int)cudaErrorInvalidValue;
  void* args[] = {const_cast<Params*>(&p)};
  if (p.cluster_mode) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = grid;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelExC(&cfg, (const void*)nrs_lm_kernel, args);
  }
  cudaError_t e = cudaMemsetAsync(p.bar, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaLaunchCooperativeKernel(kernel_of(wide, shard), dim3(grid), dim3(block), args, smem, stream);
  return (int)e;
}

}  // namespace nrs
