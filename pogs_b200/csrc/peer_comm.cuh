// Cross-GPU exchange over NVLink peer memory for the row-block multi-GPU solver.
//
// One process per GPU (torchrun); every rank allocates one "symmetric" region
// with cudaMalloc, exports it with cudaIpcGetMemHandle, and maps the regions of
// all peers (the 64-byte handles travel through torch.distributed, which is only
// plumbing here).  Kernels then read and write peer memory directly:
//
//   region layout   data[2][cap]  double-buffered contribution of this rank
//                   scal[2][8]    double-buffered scalar contributions
//                   flags[channels][world]   written BY peers, polled locally
//                   seq[channels]            local sequence counters
//                   err                      set when a wait timed out
//
// One exchange ("one-shot all-reduce"): write own contribution to data[s&1],
// __threadfence_system, store the sequence number s into the flag slot
// flags[ch][rank] of every peer, wait until every local flag slot reached s,
// read the contributions of all ranks in rank order (same order everywhere =>
// bit-identical sums on every GPU), continue.  Sequence numbers only grow, and
// the double buffer makes the slot written at s+2 safe because a peer publishes
// s+1 only after it finished reading s.  The n-vector of the ADMM iteration is
// 40 KB, so the exchange is latency-bound: two NVLink round trips (~4-5 us),
// fused into the tail of the A^T product (k_colacc) tile by tile instead of a
// separate collective launch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pogs_b200 {

constexpr int kMaxPeers = 8;
// Cross-GPU waits give up after this many SM cycles (~20 s) and raise the error flag.  Long on
// purpose: the ranks of a job are separate processes, and one of them may spend many seconds in a
// one-time host-side stall (first page-in of a CUDA library on a fresh machine) while its peers
// already wait inside a kernel.
constexpr long long kPeerTimeoutCycles = 40000000000LL;
constexpr int kScalSlots = 8;
constexpr int kMaxTileChannels = 4096;           // per-tile channels of the vector exchange
constexpr int kScalChannel = kMaxTileChannels;   // scalar exchange
constexpr int kGatherChannel = kMaxTileChannels + 1;   // all-gather of the sharded factor apply
constexpr int kNumChannels = kMaxTileChannels + 2;

struct PeerView {
  int rank = 0, world = 1;
  char* base[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap_bytes = 0;     // bytes of one data slot
  size_t data_off[2] = {0, 0};
  size_t scal_off[2] = {0, 0};
  size_t scal2_off[2] = {0, 0};  // y-side sums carried by the fold exchange of the one-launch iteration (own sequence)
  size_t gath_off[2] = {0, 0};   // double-buffered slice of x owned by this rank (cap_bytes each)
  size_t spec_off[2] = {0, 0};   // double-buffered share of the fused pass's A^T t_y' (cap_bytes each)
  // push exchange of the one-launch iteration (admm_pass.cuh): every rank WRITES its contribution into a
  // slot of every peer, [kind 0: phase B, 1: phase E][parity][source rank], so that after the flags have
  // arrived the sum reads local memory only (the pull exchange above pays a remote read round trip)
  size_t push_off[2][2] = {{0, 0}, {0, 0}};
  size_t push_stride = 0;          // bytes per source rank
  size_t scal2p_off[2] = {0, 0};   // [parity][source rank][8 doubles]
  size_t flag_off = 0, seq_off = 0, err_off = 0;

  __host__ __device__ bool active() const { return world > 1; }
  __device__ char* data(int r, unsigned s) const { return base[r] + data_off[s & 1u]; }
  __device__ char* gath(int r, unsigned s) const { return base[r] + gath_off[s & 1u]; }
  __device__ char* spec(int r, unsigned s) const { return base[r] + spec_off[s & 1u]; }
  __device__ double* scal(int r, unsigned s) const { return reinterpret_cast<double*>(base[r] + scal_off[s & 1u]); }
  __device__ double* scal2(int r, unsigned s) const { return reinterpret_cast<double*>(base[r] + scal2_off[s & 1u]); }
  __device__ char* push(int kind, int dest, int src, unsigned s) const {
    return base[dest] + push_off[kind][s & 1u] + static_cast<size_t>(src) * push_stride;
  }
  __device__ double* scal2p(int dest, int src, unsigned s) const {
    return reinterpret_cast<double*>(base[dest] + scal2p_off[s & 1u]) + src * 8;
  }
  __device__ unsigned* flag(int r, int ch, int from) const {
    return reinterpret_cast<unsigned*>(base[r] + flag_off) + static_cast<size_t>(ch) * kMaxPeers + from;
  }
  __device__ unsigned* seq(int ch) const { return reinterpret_cast<unsigned*>(base[rank] + seq_off) + ch; }
  __device__ int* err() const { return reinterpret_cast<int*>(base[rank] + err_off); }
};

__device__ __forceinline__ void st_sys(unsigned* p, unsigned v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Peer data is read with volatile loads (never served from a stale L1 line, never elided);
// they carry no memory clobber so that a batch of them can be in flight at once -- the
// ordering against the flag wait comes from the barrier + fence in peer_signal_wait.
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_peer(const double2* p) {
  double2 r;
  asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_peer(const float* p) {
  float r;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_peer(const double* p) {
  double r;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

// Publish sequence s on channel ch to every peer and wait for all of them.
// Called by all threads of a CTA after the contribution has been written;
// contains the fences and barriers.  Returns false (and raises err) on timeout.
__device__ __forceinline__ bool peer_signal_wait(const PeerView& pv, int ch, unsigned s) {
  __threadfence_system();
  __syncthreads();
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < static_cast<unsigned>(pv.world) && static_cast<int>(threadIdx.x) != pv.rank) {
    const int q = threadIdx.x;
    st_sys(pv.flag(q, ch, pv.rank), s);
    const unsigned* mine = pv.flag(pv.rank, ch, q);
    const long long t0 = clock64();
    // sequence numbers only grow; (int) difference handles wrap-around
    while (static_cast<int>(ld_sys(mine) - s) < 0) {
      if (clock64() - t0 > kPeerTimeoutCycles) {   // a peer died or the ranks desynchronised
        *pv.err() = 1;
        s_ok = 0;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  return s_ok != 0;
}

// Same for the push exchange: the threads that wrote peer memory have fenced (system scope) themselves;
// this only synchronises the CTA, raises the flags and waits.  Called by all threads of the CTA.
__device__ __forceinline__ bool peer_flags_wait(const PeerView& pv, int ch, unsigned s) {
  __syncthreads();
  __shared__ int s_ok2;
  if (threadIdx.x == 0) s_ok2 = 1;
  __syncthreads();
  if (threadIdx.x < static_cast<unsigned>(pv.world) && static_cast<int>(threadIdx.x) != pv.rank) {
    const int q = threadIdx.x;
    st_sys(pv.flag(q, ch, pv.rank), s);
    const unsigned* mine = pv.flag(pv.rank, ch, q);
    const long long t0 = clock64();
    while (static_cast<int>(ld_sys(mine) - s) < 0) {
      if (clock64() - t0 > kPeerTimeoutCycles) { *pv.err() = 1; s_ok2 = 0; break; }
    }
  }
  __syncthreads();
  return s_ok2 != 0;
}

// Sum K doubles over all ranks (rank order).  All threads of ONE CTA call it
// with the same vals; every thread gets the result.
template <int K>
__device__ __forceinline__ void peer_sum_scalars(const PeerView& pv, double (&vals)[K]) {
  static_assert(K <= kScalSlots, "too many scalars");
  if (!pv.active()) return;
  __shared__ double s_res[kScalSlots];
  const unsigned s = *pv.seq(kScalChannel) + 1u;
  __syncthreads();
  if (threadIdx.x < K) pv.scal(pv.rank, s)[threadIdx.x] = vals[threadIdx.x];
  peer_signal_wait(pv, kScalChannel, s);
  if (threadIdx.x < K) {
    double part[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) part[r] = r < pv.world ? ld_peer(pv.scal(r, s) + threadIdx.x) : 0.0;
    double acc = 0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r) acc += part[r];   // rank order; absent ranks add +0.0
    s_res[threadIdx.x] = acc;
  }
  if (threadIdx.x == 0) *pv.seq(kScalChannel) = s;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) vals[k] = s_res[k];
  __syncthreads();
}

}  // namespace pogs_b200
