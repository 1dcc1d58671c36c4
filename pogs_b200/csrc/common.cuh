// Host-side plumbing shared by the solver classes: error handling, RAII device
// buffers, launch-shape planning for the two streaming products.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace pogs_b200 {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define POGS_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw ::pogs_b200::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " \
                               __FILE__ ":" + std::to_string(__LINE__) + " (" #expr ")");     \
  } while (0)

inline size_t round_up(size_t x, size_t q) { return (x + q - 1) / q * q; }

// Function attributes (dynamic shared memory opt-in) are per device: the "already set" caches
// of the launch helpers are indexed by the calling thread's current device.
constexpr int kMaxDevices = 64;
inline int current_device_index() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
  return d < 0 ? 0 : (d >= kMaxDevices ? kMaxDevices - 1 : d);
}

// POGS_B200_TRACE=1: host wall-clock timeline of a solver's life on stderr (each mark
// synchronises the stream first, so the trace itself perturbs the overlap it reports).
struct Trace {
  bool on = false;
  std::chrono::steady_clock::time_point t0, last;
  Trace() {
    const char* e = getenv("POGS_B200_TRACE");
    on = e != nullptr && e[0] == '1';
    t0 = last = std::chrono::steady_clock::now();
  }
  void mark(const char* what, cudaStream_t s = nullptr) {
    if (!on) return;
    if (s != nullptr) cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "pogs_b200 trace: %-28s +%9.3f ms   t = %9.3f ms\n", what,
            std::chrono::duration<double, std::milli>(now - last).count(),
            std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
  }
};

// Kernel launches issued by this library (graph replays count their kernel nodes).
inline std::atomic<unsigned long long>& launch_counter() {
  static std::atomic<unsigned long long> c{0};
  return c;
}
inline void count_launch(unsigned long long n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

// ---- device memory --------------------------------------------------------------------------
// All solver buffers come from one library-owned stream-ordered pool per device
// (cudaMemPoolCreate + cudaMallocFromPoolAsync).  Freed blocks go back to the pool instead of
// the driver: releasing a solver (4+ GB for the BASELINE problem) cost ~100 ms of unmapping per
// one-shot PogsS call with cudaFree, and allocating the next one paid the mapping again.  The
// pool keeps what it has (release threshold = max) until pogs_b200_trim_memory() or process
// exit; POGS_B200_POOL=0 selects plain cudaMalloc / cudaFree.
//
// Ordering: blocks are allocated, cleared and released on the legacy default stream.  alloc()
// synchronises that stream, so a new block is usable on any stream; every owner synchronises
// its own (non-blocking) streams before its buffers are destroyed, and release() falls back to
// a device-wide synchronise when it runs during stack unwinding.
struct MemPool {
  bool enabled = true;
  std::mutex mu;
  std::vector<cudaMemPool_t> pools;   // indexed by device ordinal
  MemPool() {
    const char* e = getenv("POGS_B200_POOL");
    enabled = !(e != nullptr && e[0] == '0');
  }
  cudaMemPool_t get(int dev) {
    std::lock_guard<std::mutex> lock(mu);
    if (static_cast<size_t>(dev) >= pools.size()) pools.resize(dev + 1, nullptr);
    if (pools[dev] == nullptr) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      cudaMemPool_t pool = nullptr;
      if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
        cudaGetLastError();
        enabled = false;
        return nullptr;
      }
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      pools[dev] = pool;
    }
    return pools[dev];
  }
  // give cached blocks back to the driver
  void trim() {
    std::lock_guard<std::mutex> lock(mu);
    for (cudaMemPool_t pool : pools)
      if (pool != nullptr) cudaMemPoolTrimTo(pool, 0);
  }
};
inline MemPool& mem_pool() {
  static MemPool* p = new MemPool();   // leaked on purpose: buffers may outlive static destructors
  return *p;
}

// Zero-initialised device array, padded to a multiple of 32 elements so that
// 16 B vector reads past the logical end stay in bounds and read zeros.
template <typename T>
class DevBuf {
 public:
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  // `slack` extra elements stay allocated (and zero) behind the padded end.
  void alloc(size_t n, size_t slack = 0) {
    release();
    n_ = n;
    cap_ = round_up(n > 0 ? n : 1, 32) + slack;
    MemPool& mp = mem_pool();
    cudaMemPool_t pool = nullptr;
    if (mp.enabled) {
      int dev = 0;
      POGS_CUDA(cudaGetDevice(&dev));
      pool = mp.get(dev);
    }
    if (pool != nullptr) {
      POGS_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&p_), cap_ * sizeof(T), pool, 0));
      pooled_ = true;
    } else {
      POGS_CUDA(cudaMalloc(&p_, cap_ * sizeof(T)));
      pooled_ = false;
    }
    // The clear runs on the legacy default stream, which the solver's non-blocking streams
    // do not wait for: finish it here so that no later kernel can be overtaken by it.
    POGS_CUDA(cudaMemsetAsync(p_, 0, cap_ * sizeof(T), 0));
    POGS_CUDA(cudaStreamSynchronize(0));
  }
  void release() {
    if (p_ != nullptr) {
      if (pooled_) {
        if (std::uncaught_exceptions() > 0) cudaDeviceSynchronize();   // error path: work may be in flight
        cudaFreeAsync(p_, 0);
      } else {
        cudaFree(p_);
      }
    }
    p_ = nullptr; n_ = cap_ = 0;
  }
  T* get() const { return p_; }
  size_t size() const { return n_; }
  operator T*() const { return p_; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  bool pooled_ = false;
};

struct DeviceInfo {
  int device = 0;
  int sm_count = 148;
};

inline DeviceInfo query_device() {
  DeviceInfo d;
  POGS_CUDA(cudaGetDevice(&d.device));
  POGS_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device));
  return d;
}

// Launch shape of k_rowdot / k_colacc for an R x C row-major operand.
struct RowdotPlan {
  unsigned grid = 1;          // CTAs == rows of the partials array
  bool cta_per_row = false;   // a CTA (not a warp) owns a row: used when there are few, long rows
};
struct ColaccPlan {
  unsigned tiles = 1;         // column tiles == rows of the partials array
  unsigned chunks = 1;        // row chunks
  size_t rows_per_chunk = 1;
};

template <typename K>
inline int occupancy_of(K kernel) {
  int nb = 0;
  POGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, 0));
  return nb > 0 ? nb : 1;
}

inline RowdotPlan plan_rowdot(size_t R, int sm_count, int occ, size_t C = 0) {
  RowdotPlan p;
  const size_t need = (R + kWarps - 1) / kWarps;
  const size_t wave = static_cast<size_t>(sm_count) * occ;
  p.grid = static_cast<unsigned>(need < wave ? (need > 0 ? need : 1) : wave);
  // With fewer than ~8 rows per resident warp the one-row granularity leaves a long tail
  // (10000 rows on 4736 warps: 3 vs 2.1 rows); rows of >= 2048 entries are long enough to
  // keep a whole CTA busy, so hand out rows per CTA instead.
  const char* cr = getenv("POGS_B200_CTAROW");   // measured: no gain on B200 (79.7 vs 73.7 us for the 10k x 10k factor); off by default
  if (cr != nullptr && cr[0] == '1' && C >= 2048 && R < 8 * wave * kWarps && R >= wave) {
    p.cta_per_row = true;
    p.grid = static_cast<unsigned>(wave);
  }
  return p;
}

template <typename T>
inline ColaccPlan plan_colacc(size_t R, size_t ld, int sm_count, int occ) {
  ColaccPlan p;
  const size_t cols_per_cta = static_cast<size_t>(kThreads) * V16<T>::N;
  p.tiles = static_cast<unsigned>((ld + cols_per_cta - 1) / cols_per_cta);
  const size_t wave = static_cast<size_t>(sm_count) * occ;
  size_t chunks = wave / p.tiles;
  if (chunks < 1) chunks = 1;
  const size_t max_chunks = (R + 31) / 32;      // at least 32 rows per chunk
  if (chunks > max_chunks) chunks = max_chunks > 0 ? max_chunks : 1;
  if (chunks > 65535) chunks = 65535;
  p.rows_per_chunk = (R + chunks - 1) / chunks;
  if (p.rows_per_chunk < 1) p.rows_per_chunk = 1;
  p.chunks = static_cast<unsigned>((R + p.rows_per_chunk - 1) / p.rows_per_chunk);
  if (p.chunks < 1) p.chunks = 1;
  return p;
}

}  // namespace pogs_b200
