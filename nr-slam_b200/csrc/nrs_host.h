// nrs_host.h — host-side plumbing of libnrslam_b200: context, arenas (one pinned host block mirrored by one
// device block so a call costs ONE host->device and ONE device->host copy), staged problems.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nrslam_b200.h"
#include "nrs_direct.cuh"
#include "nrs_direct_plan.h"
#include "nrs_engine.cuh"

namespace nrs {

// Bump allocator over a pinned host block and a device block of identical layout.
class Arena {
 public:
  ~Arena() { release(); }
  void release() {
    if (host_) cudaFreeHost(host_);
    if (dev_) cudaFree(dev_);
    host_ = nullptr;
    dev_ = nullptr;
    cap_ = 0;
  }
  // Reserve capacity (bytes) — reallocates when too small. Returns false on allocation failure.
  bool reserve(size_t bytes, bool need_host) {
    if (bytes <= cap_ && (!need_host || host_)) {
      used_ = 0;
      return true;
    }
    release();
    size_t cap = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&dev_, cap) != cudaSuccess) return false;
    if (need_host && cudaMallocHost(&host_, cap) != cudaSuccess) return false;
    cap_ = cap;
    used_ = 0;
    return true;
  }
  template <typename T>
  size_t take(size_t n) {
    used_ = (used_ + 255) & ~size_t(255);
    const size_t off = used_;
    used_ += n * sizeof(T);
    return off;
  }
  template <typename T>
  T* h(size_t off) const { return reinterpret_cast<T*>(static_cast<char*>(host_) + off); }
  template <typename T>
  T* d(size_t off) const { return reinterpret_cast<T*>(static_cast<char*>(dev_) + off); }
  size_t used() const { return used_; }
  size_t capacity() const { return cap_; }
  void* host() const { return host_; }
  void* dev() const { return dev_; }

 private:
  void* host_ = nullptr;
  void* dev_ = nullptr;
  size_t cap_ = 0, used_ = 0;
};

// A problem staged in HBM: inputs (arena `in`), device-only work arrays (`work`), results (`out`).
struct Staged {
  Arena in, work, out;
  Params p;
  int grid = 0, block = 0;
  size_t smem = 0;
  // second launch plan over all rows (tracking: the lost-point stage adds rows behind the optimised ones)
  bool has_plan2 = false;
  Params p2;
  int grid2 = 0, block2 = 0;
  size_t smem2 = 0;
  bool valid = false;
  // offsets of the results inside `out`
  size_t o_pose = 0, o_x = 0, o_chi2 = 0, o_rp_level = 0, o_sp_level = 0, o_stats = 0;
  size_t o_fixed = 0;  // per-vertex fixed flags inside `in` (lost-point stage)
  size_t h2d_bytes = 0, d2h_bytes = 0;
  // rows are re-ordered along a space-filling curve before staging (locality for the resident chunks and the
  // dense block preconditioner): row_of[caller row] = engine row
  std::vector<int> row_of;
  Arena xin;  // landmark-sharded BA: push lists and edge-count flags of this rank
  // exact-solve tracking engine (nrs_direct.cu): used for the first launch plan when the problem qualifies
  bool use_direct = false;
  DirectParams dq;
  int dgrid = 0;
  long long dfactor_doubles = 0, dupdate_doubles = 0;
  bool plan_reused = false;  // the symbolic analysis of the previous frame was re-used
  size_t dsmem = 0;
};

// Landmark-sharded BA: this rank's exchange buffer and the peer mappings of the other ranks' buffers.
// Layout (identical on every rank): flags [kMaxWorld] u64 | abort int | reduction records [2][world][xstride] |
// z [4 max_rows] | x [4 max_rows].
struct Shard {
  int rank = 0, world = 1, max_rows = 0, max_poses = 0, xstride = 0;
  void* local = nullptr;
  void* peer[kMaxWorld] = {};
  size_t bytes = 0, off_abort = 0, off_red = 0, off_z = 0, off_x = 0;
  unsigned long long epoch = 0;  // exchanges completed so far (identical on every rank)
  bool attached = false, broken = false;
};

}  // namespace nrs

namespace nrs {
// A few persistent host threads for the per-frame staging loops (thousands of independent ~0.3 us items: a thread
// spawn per loop costs as much as the loop). Workers sleep on a condition variable between calls; between begin() and
// end() they spin, so a run() inside that window dispatches in about a microsecond.
class HostPool {
 public:
  ~HostPool() { stop(); }
  void start(int workers) {
    if (!th_.empty() || workers < 1) return;
    for (int t = 0; t < workers; t++) th_.emplace_back([this, t] { worker(t + 1); });
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
    th_.clear();
  }
  int threads() const { return (int)th_.size() + 1; }
  void begin() {
    {
      std::lock_guard<std::mutex> lk(m_);
      spin_.store(true, std::memory_order_release);
    }
    cv_.notify_all();
  }
  void end() { spin_.store(false, std::memory_order_release); }
  // fn(t, n_threads) on every thread of the pool (the caller is thread 0); returns when all are done.
  void run(const std::function<void(int, int)>& fn) {
    if (th_.empty()) {
      fn(0, 1);
      return;
    }
    job_ = &fn;
    pending_.store((int)th_.size(), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(m_);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    fn(0, threads());
    while (pending_.load(std::memory_order_acquire) > 0) {
    }
  }

 private:
  void worker(int t) {
    unsigned long long seen = 0;
    for (;;) {
      // spin while the owner is inside a begin()/end() window, otherwise sleep
      while (spin_.load(std::memory_order_acquire) && gen_.load(std::memory_order_acquire) == seen && !stop_) {
      }
      if (gen_.load(std::memory_order_acquire) == seen) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || gen_.load(std::memory_order_acquire) != seen || spin_.load(std::memory_order_acquire); });
        if (stop_) return;
        if (gen_.load(std::memory_order_acquire) == seen) continue;  // woken into a spin window
      }
      if (stop_) return;
      seen = gen_.load(std::memory_order_acquire);
      (*job_)(t, threads());
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::atomic<unsigned long long> gen_{0};
  std::atomic<int> pending_{0};
  std::atomic<bool> spin_{false};
  bool stop_ = false;
  const std::function<void(int, int)>* job_ = nullptr;
};

// Symbolic analysis of the exact solve kept across frames: a tracking frame whose optimised points are
// the same map points in the same order as the previous frame's, and whose regulariser pairs all lie inside the
// adjacency the plan was built from, re-uses the plan (a missing pair is a zero block of the same front).
struct PlanCache {
  DirectPlanHost plan;
  bool valid = false;
  int depth = -1, np = -1;
  std::vector<int32_t> key;          // point_vertex of the frame the plan was built for
  std::vector<uint64_t> pairs;       // sorted (min << 32 | max) of the pairs it was built from (caller rows)
  std::vector<int> pair_i, pair_j;   // the same pairs in their original order (exact-match fast path)
  long long builds = 0, reuses = 0;
};
}  // namespace nrs

struct nrslam_b200_ctx {
  nrslam_b200_options opt;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev2 = nullptr, ev3 = nullptr;  // asynchronous launches (nrslam_b200_track_pose_and_deform)
  std::string err;
  nrs::Staged staged[4];  // 0 pose_only, 1 pose_deform, 2 local_ba, 3 lost-point stage
  unsigned long long* bar = nullptr;
  int max_cluster = -1;   // largest schedulable thread-block cluster of the LM kernel (queried lazily)
  nrs::Shard shard;
  nrs::PlanCache plan_cache;       // main rounds of pose_deform
  nrs::HostPool pool;              // staging threads of the tracking path (started lazily)
  nrs::Arena graph_in, graph_out;  // nrslam_b200_graph_update_vertices staging (nrs_tri.cu)
};
