// One kernel launch per ADMM iteration (dense, row-major, m > n, direct projector).
//
// k_fused_pass (fused_pass.cuh) already turns the two passes over A of the reference iteration
// (projector_direct_dense.cpp:122-127) into one.  What was left around it -- the factor apply
// x = (I + A^T A)^-1 u (two launches at 3.4 TB/s), the controller (one launch), two gated-off
// launches and the gaps between all of them -- cost 16 % of a BASELINE iteration on one GPU and
// more than half of it on eight.  k_admm_pass runs the whole committed
// iteration in ONE persistent launch of one CTA per SM:
//
//   A  stream the rows of A_g through the TMA ring: y = A x, finish iteration k for the rows,
//      speculative first half-step of iteration k+1 for the rows, column sums of A^T t_y'
//      (unchanged: the k_fused_pass pipeline, pogs.cpp:254-278, 397-399)
//      -- grid barrier --
//   B  fold the column sums over the CTAs [and over the ranks, NVLink peer memory], speculative
//      x half-step, u' = t_x' + A^T t_y'; CTA 0 also folds [and exchanges] the five y-side sums
//      -- grid barrier --
//   C  controller (pogs.cpp:342-469): every CTA evaluates the stopping rule and the rho update
//      redundantly from the same partial sums (bit-identical decisions, no third barrier);
//      an idle map warp of CTA 0 publishes the result (controller state, mapped host progress word)
//   D  if the speculation is committed: x'' = M u' for the NEXT iteration.  M is symmetric and
//      kept as a packed lower triangle (diagonal halved), streamed through the same TMA ring
//      exactly once: row i gives the dot product sum_{j<=i} M_ij u_j AND, in the same sweep over
//      shared memory, the column update sum_{i>=j} M_ij u_i -- n^2/2 elements instead of n^2,
//      at the streaming rate of the A pass.  Row blocks: the rows of the triangle are dealt to
//      the CTAs of ALL ranks, each rank streams 1/G of it.
//      -- grid barrier --
//   E  fold row dots + column sums over the CTAs [and ranks], x half-step of the next iteration
//      (z <- x'', z~ <- t - x'', the two residual terms; pogs.cpp:296, 342-348, 397-399).
//
// Discarded speculation (rho moved) or a pending exact-residual decision skip D and E and leave a flag in
// the controller; the launches that follow in the captured run return at once, and the gated service kernels
// at the head of the next round take over: the exact-residual pass, k_prox, k_colacc and this kernel in mode 1
// (phases D and E only).  One iteration per launch, deliberately (see the comment in the kernel).
// Per-iteration cross-GPU traffic is two push exchanges (the n-vector of phase B carrying the y-side scalars,
// the n-vector of phase E) instead of three pull exchanges behind three launch boundaries.
#pragma once

#include "fused_pass.cuh"

namespace pogs_b200 {

constexpr int kPassEChannel = kMaxTileChannels / 2;   // fold channels of phase E: [2048, 4096)

// ---- packed lower triangle of the symmetric factor ------------------------------------------------
// Row i holds M_i0 .. M_i,i-1, M_ii / 2 and zeros up to the next 16 B boundary; rows are
// contiguous.  With the diagonal halved the row sweep (dot with u) and the column sweep
// (+= row * u_i) can treat every stored entry alike and still count the diagonal once.
template <typename T> __host__ __device__ inline size_t sym_row_len(size_t i) {
  constexpr size_t V = V16<T>::N;
  return (i / V + 1) * V;
}
template <typename T> __host__ __device__ inline size_t sym_row_off(size_t i) {
  constexpr size_t V = V16<T>::N;
  const size_t q = i / V, r = i % V;
  return V * (q + 1) * (V * q / 2 + r);
}
template <typename T>
__global__ void __launch_bounds__(kThreads) k_pack_sym(size_t n, const T* __restrict__ M, size_t ldm, T* __restrict__ P) {
  const size_t i = blockIdx.x;
  if (i >= n) return;
  T* dst = P + sym_row_off<T>(i);
  const size_t len = sym_row_len<T>(i);
  for (size_t j = threadIdx.x; j < len; j += kThreads)
    dst[j] = j < i ? M[i * ldm + j] : (j == i ? M[i * ldm + i] * T(0.5) : T(0));
}

// ---- row sources of the streaming pipeline ------------------------------------------------------------
template <typename T>
struct DenseRows {                 // rows r0 .. r0+nrows of a row-major array
  const T* A; size_t ld, r0; unsigned nvec;
  __device__ __forceinline__ const T* ptr(unsigned r) const { return A + (r0 + r) * ld; }
  __device__ __forceinline__ unsigned vecs(unsigned) const { return nvec; }
  __device__ __forceinline__ size_t index(unsigned r) const { return r0 + r; }
};
template <typename T>
struct SymRows {                   // rows of the packed triangle dealt boustrophedon to `wtot` workers
  const T* P; unsigned w, wtot, nk;   // nk = rows of this worker
  // The worker's rows k = 0 .. nk-1 (row k * wtot +- w: lengths grow with k) are visited from both ends
  // alternately -- shortest, longest, second shortest, ... -- so that the ring always holds a mix of short
  // and long rows: visited in order of length, every CTA starts with rows of a few KB at the same time and
  // the whole GPU is latency-bound until the rows are long enough to fill the pipe.
  __device__ __forceinline__ unsigned row(unsigned r) const {
    const unsigned k = (r & 1u) ? (nk - 1u - (r >> 1)) : (r >> 1);
    return k * wtot + ((k & 1u) ? (wtot - 1u - w) : w);
  }
  __device__ __forceinline__ const T* ptr(unsigned r) const { return P + sym_row_off<T>(row(r)); }
  __device__ __forceinline__ unsigned vecs(unsigned r) const { return row(r) / V16<T>::N + 1u; }
  __device__ __forceinline__ size_t index(unsigned r) const { return row(r); }
};
__host__ __device__ inline unsigned sym_rows_of_worker(unsigned n, unsigned w, unsigned wtot) {
  const unsigned full = n / wtot, rem = n % wtot;
  const unsigned pos = (full & 1u) ? (wtot - 1u - w) : w;
  return full + (pos < rem ? 1u : 0u);
}
__host__ __device__ inline unsigned sym_owner_worker(unsigned i, unsigned wtot) {
  const unsigned k = i / wtot, pos = i % wtot;
  return (k & 1u) ? (wtot - 1u - pos) : pos;
}

// Row functor of phase D: keeps the row dot, hands u_i to nobody (the main warps fetch it early).
template <typename T>
struct SymRowOp {
  static constexpr int NRED = 1;
  T* xrow;
  struct State {};
  __device__ __forceinline__ void load(size_t, State&) const {}
  __device__ __forceinline__ T apply(size_t i, const State&, T dot, T, double (&)[NRED]) const {
    xrow[i] = dot;
    return T(0);
  }
  __device__ __forceinline__ void store(unsigned, unsigned, const double*) const {}
};
// Column functor of phase E: the x half-step (EpiState arithmetic).
template <typename T>
struct XStateColOp {
  static constexpr int NRED = 2;
  EpiState<T> epi;
  double* xs_part;                 // [nfold][2]
  __device__ __forceinline__ void apply(size_t j, T total, T, double (&red)[NRED]) const { epi(j, total, red); }
  __device__ __forceinline__ void store(unsigned cta, const double* red) const {
    xs_part[static_cast<size_t>(cta) * 2 + 0] = red[0];
    xs_part[static_cast<size_t>(cta) * 2 + 1] = red[1];
  }
};

// ---- shared-memory bookkeeping of one streaming phase -----------------------------------------------------
template <typename T, int B, int RN>
struct PassSmem {
  uint64_t full[32];                                    // copy engine -> main warps, one per ring slot
  uint64_t dots[kFusedMapWarps], coefr[kFusedMapWarps], free_[kFusedMapWarps];   // index = batch mod W
  T dot[kFusedMapWarps][kFusedWarps][B];
  T coef[kFusedMapWarps][B];
  double red[kFusedMapWarps][RN];
};
template <typename SM>
__device__ __forceinline__ void pass_smem_init(SM& sh, unsigned nslots) {
  for (unsigned s = 0; s < nslots; ++s) mbar_init(&sh.full[s], 1);
  for (int p = 0; p < kFusedMapWarps; ++p) {
    mbar_init(&sh.dots[p], kFusedWarps);
    mbar_init(&sh.coefr[p], 1);
    mbar_init(&sh.free_[p], kFusedWarps);
  }
}

// The streaming pipeline of k_fused_pass as a device function (see the comment there): the 16 main
// warps and W map warps of one CTA push `nrows` rows through the ring.  Leaves the column sums in
// acc (main threads) and the per-map-warp reduction terms in sh.red.
// EARLY: the column coefficient of a row does not depend on its dot product (phase D: coef_i =
// u_i): the main warps fetch it themselves and do the dot product and the column update in ONE
// sweep over the row in shared memory; the map warps only finish the dots and refill the ring
// (barriers: dots = "batch done", coefr = "dots consumed"; free_ unused).
// The column sums are written to colpart_out by the main threads at the end; nothing is handed back in
// registers, and the kernel keeps no per-thread state alive across a call (a first version with
// loop-carried state around two inlined copies made the register allocator spill inside the streaming
// loops: 755 instead of 549 us per pass over A; a not-inlined version serialised the row loads).
template <typename T, bool SQ, int NV, int B, bool EARLY, typename Rows, typename RowOp, typename SM>
__device__ __forceinline__ void stream_rows(SM& sh, unsigned char* ring, const Rows src, unsigned nrw, unsigned slot_bytes,
                                            unsigned nslots, unsigned W, const T* __restrict__ xin, size_t nvec_x,
                                            const T* __restrict__ coef_early, const RowOp& rop_ref, T rho,
                                            T* __restrict__ colpart_out, unsigned prefilled, int tid) {
  using VT = typename V16<T>::type;
  constexpr int RN = RowOp::NRED;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned nbt = (nrw + B - 1) / B;
  if (tid < kFusedThreads) {
    VT acc[NV];   // column accumulators
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = zerov(static_cast<VT*>(nullptr));
    // ================= main warps =================
    VT xv[NV];   // this thread's slice of the multiplied vector (coherent loads: it may have been
                 // written earlier in this launch by other CTAs)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
      xv[k] = jv < nvec_x ? ld_cg(reinterpret_cast<const VT*>(xin) + jv) : zerov(static_cast<VT*>(nullptr));
    }
    unsigned slot = 0, phase = 0;        // ring slot / mbarrier parity of the first row of batch bi
    unsigned wb = 0;                     // bi % W
    if constexpr (EARLY) {
      unsigned parc = 0;                 // parity of the "dots consumed" barrier for the batch W back
      // column coefficients are fetched one batch ahead: an L2 round trip per row in front of the sweep
      // cost more than the sweep itself
      T cf_next[B];
#pragma unroll
      for (int b = 0; b < B; ++b) cf_next[b] = static_cast<unsigned>(b) < nrw ? __ldcg(coef_early + src.index(b)) : T(0);
      for (unsigned bi = 0; bi < nbt; ++bi) {
        const unsigned left = nrw - bi * B;
        const int nb = static_cast<int>(left < static_cast<unsigned>(B) ? left : B);
        T cf[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
          cf[b] = cf_next[b];
          const unsigned rn = (bi + 1) * B + b;
          cf_next[b] = rn < nrw ? __ldcg(coef_early + src.index(rn)) : T(0);
        }
        T d[B];
        unsigned s = slot, ph = phase;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          d[b] = 0;
          if (b < nb) {
            const unsigned nv = src.vecs(bi * B + b);
            mbar_wait(&sh.full[s], ph);
            const VT* rowp = reinterpret_cast<const VT*>(ring + static_cast<size_t>(s) * slot_bytes);
            // all loads of the row first (independent, in flight together), then the arithmetic;
            // vectors past the end of the row are zeros and add nothing
            VT v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const unsigned jv = static_cast<unsigned>(tid) + static_cast<unsigned>(k) * kFusedThreads;
              v[k] = jv < nv ? rowp[jv] : zerov(static_cast<VT*>(nullptr));
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              d[b] += dotv<SQ>(v[k], xv[k]);
              fmav<SQ>(acc[k], v[k], cf[b]);
            }
            if (++s == nslots) { s = 0; ph ^= 1u; }
          }
        }
        // the dots of batch bi-W must have been read before their slot is overwritten
        if (bi >= W) mbar_wait(&sh.coefr[wb], parc);
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const T dd = warp_sum(d[b]);
          if (lane == 0) sh.dot[wb][warp][b] = dd;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.dots[wb]);
        slot = s; phase = ph;
        if (++wb == W) { wb = 0; if (bi >= W) parc ^= 1u; }
      }
    } else {
      unsigned uslot = 0, wp = 0, parp = 0, pb = 0;   // update side: slot, pb % W, (pb / W) & 1, batch index
      for (unsigned bi = 0; bi + 1 < nbt + W; ++bi) {
        if (bi < nbt) {
          // ---- partial dot products of batch bi ---------------------------------------------------
          const unsigned left = nrw - bi * B;
          const int nb = static_cast<int>(left < static_cast<unsigned>(B) ? left : B);
          T d[B];
          unsigned s = slot, ph = phase;
#pragma unroll
          for (int b = 0; b < B; ++b) {
            d[b] = 0;
            if (b < nb) {
              const unsigned nv = src.vecs(bi * B + b);
              mbar_wait(&sh.full[s], ph);
              const VT* rowp = reinterpret_cast<const VT*>(ring + static_cast<size_t>(s) * slot_bytes);
              VT v[NV];
#pragma unroll
              for (int k = 0; k < NV; ++k) {
                const unsigned jv = static_cast<unsigned>(tid) + static_cast<unsigned>(k) * kFusedThreads;
                v[k] = jv < nv ? rowp[jv] : zerov(static_cast<VT*>(nullptr));
              }
#pragma unroll
              for (int k = 0; k < NV; ++k) d[b] += dotv<SQ>(v[k], xv[k]);
              if (++s == nslots) { s = 0; ph ^= 1u; }
            }
          }
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const T dd = warp_sum(d[b]);
            if (lane == 0) sh.dot[wb][warp][b] = dd;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.dots[wb]);
          slot = s; phase = ph;
          if (++wb == W) wb = 0;
        }
        if (bi + 1 >= W && pb < nbt) {
          // ---- column update of batch pb = bi-W+1 from the rows still in shared memory ----------------
          const unsigned left = nrw - pb * B;
          const int nbp = static_cast<int>(left < static_cast<unsigned>(B) ? left : B);
          mbar_wait(&sh.coefr[wp], parp);
          unsigned s = uslot;
#pragma unroll
          for (int b = 0; b < B; ++b) {
            if (b < nbp) {
              const unsigned nv = src.vecs(pb * B + b);
              const T c = sh.coef[wp][b];
              const VT* rowp = reinterpret_cast<const VT*>(ring + static_cast<size_t>(s) * slot_bytes);
              VT v[NV];
#pragma unroll
              for (int k = 0; k < NV; ++k) {
                const unsigned jv = static_cast<unsigned>(tid) + static_cast<unsigned>(k) * kFusedThreads;
                v[k] = jv < nv ? rowp[jv] : zerov(static_cast<VT*>(nullptr));
              }
#pragma unroll
              for (int k = 0; k < NV; ++k) fmav<SQ>(acc[k], v[k], c);
              if (++s == nslots) s = 0;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.free_[wp]);
          uslot = s;
          ++pb;
          if (++wp == W) { wp = 0; parp ^= 1u; }
        }
      }
    }
    // column sums of this CTA
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
      if (jv < nvec_x) reinterpret_cast<VT*>(colpart_out)[jv] = acc[k];
    }
  } else {
    // ================= map warps: warp j < W owns the batches b with b % W == j =================
    const int j = warp - kFusedWarps;
    const RowOp rop = rop_ref;   // private copy: through the reference (shared memory) every store forces a reload
    double red[RN];   // held by lanes < B
#pragma unroll
    for (int k = 0; k < RN; ++k) red[k] = 0;
    // first fill of the ring (map warp 0); rows below `prefilled` were issued ahead by ring_prefill
    if (j == 0 && lane == 0) {
      for (unsigned r = prefilled; r < nrw && r < nslots; ++r) {
        const unsigned bytes = src.vecs(r) * 16u;
        mbar_expect_tx(&sh.full[r], bytes);
        bulk_g2s(ring + static_cast<size_t>(r) * slot_bytes, src.ptr(r), bytes, &sh.full[r]);
      }
    }
    const unsigned step = W * B;                       // rows between two batches of this warp (<= nslots)
    unsigned par = 0;                                  // (bi / W) & 1
    unsigned mslot = (static_cast<unsigned>(j) * B) % nslots;   // ring slot of the first row of batch bi
    for (unsigned bi = j; static_cast<unsigned>(j) < W && bi < nbt; bi += W) {
      const unsigned row = bi * B, left = nrw - row;
      const int nb = static_cast<int>(left < static_cast<unsigned>(B) ? left : B);
      typename RowOp::State rs{};
      if (lane < nb) rop.load(src.index(row + lane), rs);
      mbar_wait(&sh.dots[j], par);
      if (lane < nb) {
        double tot = 0;
#pragma unroll
        for (int w = 0; w < kFusedWarps; ++w) tot += static_cast<double>(sh.dot[j][w][lane]);
        const T c = rop.apply(src.index(row + lane), rs, static_cast<T>(tot), rho, red);
        if constexpr (!EARLY) sh.coef[j][lane] = c;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.coefr[j]);        // !EARLY: coefficients ready; EARLY: dots consumed
      // the batch has been read (twice) once every main warp has updated its columns: refill its slots
      if constexpr (!EARLY) mbar_wait(&sh.free_[j], par);
      if (lane == 0) {
        unsigned sl = mslot;
        for (int b = 0; b < nb; ++b) {
          const unsigned r = row + b + nslots;       // row that takes over the slot of row bi*B + b
          if (r < nrw) {
            const unsigned bytes = src.vecs(r) * 16u;
            mbar_expect_tx(&sh.full[sl], bytes);
            bulk_g2s(ring + static_cast<size_t>(sl) * slot_bytes, src.ptr(r), bytes, &sh.full[sl]);
          }
          if (++sl == nslots) sl = 0;
        }
      }
      par ^= 1u;
      mslot += step;
      if (mslot >= nslots) mslot -= nslots;
    }
    // per-warp sums of the reduction terms, lanes folded in fixed order
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < RN; ++k) sh.red[j][k] = 0;
    }
    __syncwarp();
    for (int b = 0; b < B; ++b) {
      if (lane == b) {
#pragma unroll
        for (int k = 0; k < RN; ++k) sh.red[j][k] += red[k];
      }
      __syncwarp();
    }
  }
}

// Fold phase shared by B and E.  CTA b < nfold finishes FV 16 B column vectors: sums the per-CTA
// column sums (fixed order), adds `extra` where this rank owns the entry (phase E: the row dots),
// exchanges the slice with the other ranks and runs the column functor, one column per thread.
// All threads of the CTA call it.  kind 0: spec slots (phase B), 1: gath slots (phase E).
template <typename T, typename ColOp>
__device__ __forceinline__ void fold_columns(const T* __restrict__ colpart, size_t ld, size_t nvec, size_t n,
                                             unsigned nparts, unsigned FV, const T* __restrict__ extra, unsigned wtot,
                                             unsigned grid, const ColOp& cop, T rho, const PeerView& pv, int channel,
                                             int kind, typename V16<T>::type* s_fold, T* s_tot /*[kFusedThreads]*/,
                                             double* s_rx /*[kFusedWarps][kMaxRed]*/,
                                             double* s_yscal /*[5] or null: CTA 0 of phase B*/, int tid, unsigned bid) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  constexpr int CN = ColOp::NRED;
  const int lane = tid & 31, warp = tid >> 5;
  const bool main_thr = tid < kFusedThreads;
  const unsigned NG = kFusedThreads / FV;   // NG groups of partials x FV vectors
  const unsigned v16 = tid & (FV - 1), grp = tid / FV;
  const size_t jv = static_cast<size_t>(bid) * FV + v16;
  VT part = zerov(static_cast<VT*>(nullptr));
  if (main_thr && jv < nvec) {
    unsigned p = grp;
    for (; p + 3 * NG < nparts; p += 4 * NG) {   // four independent loads in flight
      const VT a0 = ld_cg(reinterpret_cast<const VT*>(colpart + static_cast<size_t>(p) * ld) + jv);
      const VT a1 = ld_cg(reinterpret_cast<const VT*>(colpart + static_cast<size_t>(p + NG) * ld) + jv);
      const VT a2 = ld_cg(reinterpret_cast<const VT*>(colpart + static_cast<size_t>(p + 2 * NG) * ld) + jv);
      const VT a3 = ld_cg(reinterpret_cast<const VT*>(colpart + static_cast<size_t>(p + 3 * NG) * ld) + jv);
      addv(part, a0); addv(part, a1); addv(part, a2); addv(part, a3);
    }
    for (; p < nparts; p += NG) addv(part, ld_cg(reinterpret_cast<const VT*>(colpart + static_cast<size_t>(p) * ld) + jv));
  }
  if (main_thr) s_fold[grp * FV + v16] = part;
  __syncthreads();
  VT total = zerov(static_cast<VT*>(nullptr));
  const bool fin = static_cast<unsigned>(tid) < FV && jv < nvec;   // threads that finish a column vector
  if (fin) {
    for (unsigned q = 0; q < NG; ++q) addv(total, s_fold[q * FV + tid]);   // fixed order
    if (extra != nullptr) {
      T* te = reinterpret_cast<T*>(&total);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const size_t j = jv * VEC + e;
        if (j < n && sym_owner_worker(static_cast<unsigned>(j), wtot) / grid == static_cast<unsigned>(pv.rank))
          te[e] += __ldcg(extra + j);
      }
    }
  }
  if (pv.active()) {
    // row blocks: `total` is one rank's share.  Push it into the slot this rank owns on EVERY rank (itself
    // included), fence, raise the flags; once the peers' flags are in, the sum reads local memory only, in
    // rank order (same bits on every rank).  One-way NVLink latencies instead of a remote read round trip.
    const unsigned seq = *pv.seq(channel) + 1u;
    const bool yw = s_yscal != nullptr && tid < 5;
    if (fin) {
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) reinterpret_cast<VT*>(pv.push(kind, r, pv.rank, seq))[jv] = total;
    }
    if (yw) {
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) pv.scal2p(r, pv.rank, seq)[tid] = s_yscal[tid];
    }
    if (fin || yw) __threadfence_system();
    peer_flags_wait(pv, channel, seq);
    if (fin) {
      VT share[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) share[r] = ld_peer(reinterpret_cast<const VT*>(pv.push(kind, pv.rank, r, seq)) + jv);
      total = zerov(static_cast<VT*>(nullptr));
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) addv(total, share[r]);
    }
    if (yw) {
      double sh5[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) sh5[r] = r < pv.world ? ld_peer(pv.scal2p(pv.rank, r, seq) + tid) : 0.0;
      double acc5 = 0;
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) acc5 += sh5[r];
      s_yscal[tid] = acc5;
    }
    if (tid == 0) *pv.seq(channel) = seq;
  }
  // one column per thread for the column functor (an iterative prox there would otherwise run
  // four columns in sequence on FV threads)
  if (static_cast<unsigned>(tid) < FV) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) s_tot[tid * VEC + e] = elemv(total, e);
  }
  __syncthreads();
  double rx[CN];
#pragma unroll
  for (int k = 0; k < CN; ++k) rx[k] = 0;
  const unsigned ne = FV * VEC;   // <= kFusedThreads
  if (static_cast<unsigned>(tid) < ne) {
    const size_t j = static_cast<size_t>(bid) * ne + tid;
    if (j < n) cop.apply(j, s_tot[tid], rho, rx);
  }
  if (main_thr) {
#pragma unroll
    for (int k = 0; k < CN; ++k) {
      const double v = warp_sum(rx[k]);
      if (lane == 0) s_rx[warp * kMaxRed + k] = v;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double t[CN];
#pragma unroll
    for (int k = 0; k < CN; ++k) t[k] = 0;
    const unsigned nw = (ne + 31) / 32;
    for (unsigned q = 0; q < nw; ++q) {   // fixed order
#pragma unroll
      for (int k = 0; k < CN; ++k) t[k] += s_rx[q * kMaxRed + k];
    }
    cop.store(bid, t);
  }
}

// Sum of column k of an [nb][stride] array of partials by ONE warp (lane-strided, then the
// shuffle tree): the same result in every CTA and on every rank.
__device__ __forceinline__ double warp_fold(const double* p, unsigned nb, int stride, int k, int lane) {
  double s = 0;
  for (unsigned b = lane; b < nb; b += 32) s += __ldcg(p + static_cast<size_t>(b) * stride + k);
  return warp_sum(s);
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Issues the first fill of a phase's ring ahead of time (one thread): the rows of A and of the packed
// factor never change, so the copies can start as soon as the previous phase has left the ring.
// Returns the number of rows issued; stream_rows is told to skip them.
template <typename Rows, typename SM>
__device__ __forceinline__ unsigned ring_prefill(SM& sh, unsigned char* ring, const Rows src, unsigned nrw,
                                                 unsigned slot_bytes, unsigned nslots) {
  unsigned r = 0;
  for (; r < nrw && r < nslots; ++r) {
    const unsigned bytes = src.vecs(r) * 16u;
    mbar_expect_tx(&sh.full[r], bytes);
    bulk_g2s(ring + static_cast<size_t>(r) * slot_bytes, src.ptr(r), bytes, &sh.full[r]);
  }
  return r;
}
// Re-arm a phase's barriers for its next use (all of its waits have completed; one thread).
template <typename SM>
__device__ __forceinline__ void pass_smem_reinit(SM& sh, unsigned nslots) {
  for (unsigned s = 0; s < nslots; ++s) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.full[s])) : "memory");
  for (int p = 0; p < kFusedMapWarps; ++p) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.dots[p])) : "memory");
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.coefr[p])) : "memory");
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.free_[p])) : "memory");
  }
  pass_smem_init(sh, nslots);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Buffers of one iteration parity p (iteration k uses p = k & 1).
template <typename T>
struct ParityArgs {
  AdmmRowOp<T> rop;            // phase A of iteration p
  AdmmColOp<T> cop;            // phase B of iteration p
  const T* x;                  // multiplied vector of phase A: x^{k+1} = x_[1-p]
  EpiState<T> xnext;           // x half-step the tail of iteration p runs, i.e. that of iteration 1-p
  const double* first_x_spec;  // [nfold][3] x-side first-half-step sums of iteration p (committed speculation)
  const double* spec_y_next;   // [grid][3] written by phase A: y rows of the next iteration's speculation
  double* ysum_cur;            // [8]: 0..2 first-half-step y sums of iteration p (filled earlier), 3..4 written here
  double* ysum_next;           // [8]: 0..2 written here for iteration 1-p
};

template <typename T>
struct PassArgs {
  // operator (local row block) and pass bookkeeping
  const T* A; size_t m, n, ld;
  T* colpart; unsigned* bar;
  unsigned nfold, fold_vecs, nstages, nmap;
  // factor apply (phases D, E)
  const T* Mlow;               // packed lower triangle, diagonal halved
  const T* u;                  // phase D input (written by phase B's column functor / by k_colacc)
  T* xrow;                     // [n] row dots
  double* xs_part;             // [nfold][2] x half-step sums (written by phase E, read by the next phase C)
  // controller
  Ctrl<T>* ctrl;
  const double* first_x_prox;  // [prox_gx][3] x-side first-half-step sums from k_prox (speculation discarded)
  unsigned prox_gx;
  const double* ys_part;       // [grid][2] written by phase A
  volatile unsigned* host_progress;   // mapped host memory: {iterations done, done flag, rounds done (k_service_done)}
  unsigned long long* phase_ns;   // [9] accumulated phase times of CTA 0, may be null
  int mode;                    // 0: one whole iteration, 1: factor apply only (phases D, E)
};

// Controller arithmetic of one iteration on a private copy of the state (pogs.cpp:268-273, 342-352):
// xs = {<w,z12>, |w|^2, |z12|^2, |xprev-x|^2, |x12-x|^2} of the x part, ys likewise of the y part.
template <typename T>
__device__ __forceinline__ void control_step(Ctrl<T>* c, const double* xs, const double* ys, volatile unsigned* host_progress) {
  const T rho_c = c->rho;
  c->gap = m_abs(static_cast<T>(xs[0] + ys[0]));
  c->eps_gap = c->sqrtmn_atol + c->rel_tol * static_cast<T>(sqrt(xs[1] + ys[1])) * static_cast<T>(sqrt(xs[2] + ys[2]));
  c->eps_pri = c->sqrtm_atol + c->rel_tol * static_cast<T>(sqrt(ys[2]));
  c->eps_dua = rho_c * (c->sqrtn_atol + c->rel_tol * static_cast<T>(sqrt(xs[1])));
  c->nrm_s = rho_c * (c->nrmA * static_cast<T>(sqrt(ys[3])) + static_cast<T>(sqrt(xs[3])));
  c->nrm_r = c->nrmA * static_cast<T>(sqrt(xs[4])) + static_cast<T>(sqrt(ys[4]));
  const bool need = c->nrm_r < T(10) * c->eps_pri && c->nrm_s < T(10) * c->eps_dua;
  c->need_exact = need ? 1 : 0;
  if (!need) finish_iteration(c, false, host_progress, /*tail_follows=*/true);
  else c->need_solve = 1;   // the exact branch decides; the factor apply then runs in the rare path
}

template <typename T, int NV, int B>
__global__ void __launch_bounds__(kFusedCta, 1)
k_admm_pass(PassArgs<T> a, ParityArgs<T> par0, ParityArgs<T> par1, Gate gate, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  // Programmatic dependent launch (the runs of identical launches in the captured loop carry the attribute):
  // this grid may have been scheduled while the previous one was still finishing; nothing of the previous
  // grid's results is touched before the wait, and the next grid is allowed to be scheduled right away.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (gate_closed(gate)) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ PassSmem<T, B, AdmmRowOp<T>::NRED> shA;
  __shared__ PassSmem<T, B, 1> shD;
  __shared__ Ctrl<T> s_ctrl;
  __shared__ ParityArgs<T> s_par[2];
  __shared__ VT s_fold[kFusedThreads];
  __shared__ T s_tot[kFusedThreads];
  __shared__ double s_rx[kFusedWarps * kMaxRed];
  __shared__ double s_c[16];
  __shared__ double s_y[8];
  __shared__ unsigned long long s_tprev;

  // ONE iteration per launch, on purpose: with a loop over iterations around the phases the register
  // allocator spilled and serialised the row loads of the streaming loops (587 instead of 549 us per pass
  // over A).  The captured graph is a run of identical launches of this kernel; the parity of the
  // iteration in hand is read from the controller, and a launch that finds a rare event pending (exact
  // residuals due, speculation discarded, factor apply missing) returns at once -- the service kernels that
  // follow the run in the graph take care of it.
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned bid = blockIdx.x;
  const size_t ld = a.ld, nvec = ld / VEC;
  const unsigned row_bytes = static_cast<unsigned>(ld * sizeof(T));
  const unsigned nslots = a.nstages, W = a.nmap;
  const bool timing = a.phase_ns != nullptr && bid == 0 && tid == 0;
  auto lap = [&](int k) {
    if (timing) { const unsigned long long t = global_ns(); a.phase_ns[k] += t - s_tprev; s_tprev = t; }
  };
  const bool prefiller = tid == kFusedThreads;                                          // first map warp, lane 0
  const bool publisher = bid == 0 && tid == kFusedThreads + 32 * (kFusedMapWarps - 1);   // last map warp of CTA 0

  if (tid == 0) {
    s_ctrl = *a.ctrl;   // before anything in this launch changes it (CTA 0 publishes it after phase C)
    s_par[0] = par0; s_par[1] = par1;
    if (a.mode == 0) {
      // rho-action prediction: speculate on what the controller does if it repeats its last action
      // (finish_iteration, kernels.cuh: the same operations on the same operands give the same bits)
      T srho = s_ctrl.rho, ssc = T(1);
      if (s_ctrl.adaptive_rho && s_ctrl.pred_act > 0 && srho < T(1e4)) { ssc = 1 / s_ctrl.delta; srho *= s_ctrl.delta; }
      else if (s_ctrl.adaptive_rho && s_ctrl.pred_act < 0 && srho > T(1e-4)) { ssc = s_ctrl.delta; srho /= s_ctrl.delta; }
      s_ctrl.spec_pred = 1; s_ctrl.spec_rho = srho; s_ctrl.spec_scale = ssc;
      const int q = static_cast<int>(s_ctrl.k & 1u);
      s_par[q].rop.zsc = ssc; s_par[q].cop.zsc = ssc;
    }
    pass_smem_init(shA, nslots);
    pass_smem_init(shD, nslots);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (timing) s_tprev = global_ns();
  }
  __syncthreads();
  if (s_ctrl.done) return;
  if (a.mode == 0 && (s_ctrl.need_exact || s_ctrl.need_solve)) return;   // a rare event waits for the service kernels

  const unsigned wtot = static_cast<unsigned>(pv.world) * gridDim.x;
  const unsigned wme = static_cast<unsigned>(pv.rank) * gridDim.x + bid;
  const unsigned nrowsM = sym_rows_of_worker(static_cast<unsigned>(a.n), wme, wtot);
  const SymRows<T> rowsM{a.Mlow, wme, wtot, nrowsM};
  T* const my_colpart = a.colpart + static_cast<size_t>(bid) * ld;
  const int p = static_cast<int>(s_ctrl.k & 1u);   // parity of the iteration in hand
  unsigned preD = 0;

  if (a.mode == 0) {
    const ParityArgs<T>& pa = s_par[p];
    const T rho = s_ctrl.spec_rho;   // rho is only used by the speculative half-steps of phases A and B
    // ================= phase A: one pass over the local rows of A =================
    {
      const size_t rows_per_cta = (a.m + gridDim.x - 1) / gridDim.x;
      const size_t r0 = static_cast<size_t>(bid) * rows_per_cta;
      const size_t r1 = r0 + rows_per_cta < a.m ? r0 + rows_per_cta : a.m;
      const unsigned nrowsA = r1 > r0 ? static_cast<unsigned>(r1 - r0) : 0u;
      const DenseRows<T> rowsA{a.A, ld, r0, static_cast<unsigned>(nvec)};
      stream_rows<T, false, NV, B, false>(shA, smem_raw, rowsA, nrowsA, row_bytes, nslots, W, pa.x, nvec,
                                          static_cast<const T*>(nullptr), pa.rop, rho, my_colpart, 0u, tid);
    }
    __syncthreads();
    // the ring is free: start the first rows of the factor for phase D (speculatively: D runs unless the
    // controller decides otherwise; drained below if it does not)
    if (prefiller) preD = ring_prefill(shD, smem_raw, rowsM, nrowsM, row_bytes, nslots);
    if (tid == 0) {
      double tot[AdmmRowOp<T>::NRED];
#pragma unroll
      for (int k = 0; k < AdmmRowOp<T>::NRED; ++k) {
        tot[k] = 0;
#pragma unroll
        for (int w = 0; w < kFusedMapWarps; ++w) tot[k] += shA.red[w][k];   // fixed order
      }
      pa.rop.store(bid, a.nfold, tot);
    }
    lap(0);
    if (!grid_barrier(a.bar, gridDim.x)) return;
    lap(1);

    // ================= phase B: fold A^T t_y' over CTAs [and ranks], speculative x half-step =================
    if (bid == 0 && warp == kFusedWarps) {
      // CTA 0, first map warp: the five y-side sums of this rank (two of this iteration, three of the
      // next iteration's speculation); exchanged with the column slice below
      const double v0 = warp_fold(a.ys_part, gridDim.x, 2, 0, lane), v1 = warp_fold(a.ys_part, gridDim.x, 2, 1, lane);
      const double v2 = warp_fold(pa.spec_y_next, gridDim.x, 3, 0, lane), v3 = warp_fold(pa.spec_y_next, gridDim.x, 3, 1, lane);
      const double v4 = warp_fold(pa.spec_y_next, gridDim.x, 3, 2, lane);
      if (lane == 0) { s_y[0] = v0; s_y[1] = v1; s_y[2] = v2; s_y[3] = v3; s_y[4] = v4; }
    }
    if (bid < a.nfold) {
      fold_columns<T>(a.colpart, ld, nvec, a.n, gridDim.x, a.fold_vecs, static_cast<const T*>(nullptr), 1u, 1u, pa.cop, rho,
                      pv, static_cast<int>(bid), 0, s_fold, s_tot, s_rx, bid == 0 ? s_y : nullptr, tid, bid);
      if (bid == 0 && tid < 5) {
        // (fold_columns ends behind a __syncthreads that follows the last write of s_y)
        if (tid < 2) pa.ysum_cur[3 + tid] = s_y[tid]; else pa.ysum_next[tid - 2] = s_y[tid];
      }
    }
    lap(2);
    if (!grid_barrier(a.bar, gridDim.x)) return;
    lap(3);

    // ================= phase C: controller, evaluated by every CTA on its own copy =================
    {
      const bool spec = s_ctrl.spec_miss == 0;
      const double* fx = spec ? pa.first_x_spec : a.first_x_prox;
      const unsigned fxn = spec ? a.nfold : a.prox_gx;
      if (warp < 3) { const double v = warp_fold(fx, fxn, 3, warp, lane); if (lane == 0) s_c[warp] = v; }
      else if (warp < 5) { const double v = warp_fold(a.xs_part, a.nfold, 2, warp - 3, lane); if (lane == 0) s_c[warp] = v; }
      else if (warp == 5 && lane < 5) s_c[5 + lane] = __ldcg(pa.ysum_cur + lane);
      __syncthreads();
      if (tid == 0) {
        control_step(&s_ctrl, s_c, s_c + 5, nullptr);
        if (!s_ctrl.done && (s_ctrl.need_exact || s_ctrl.need_solve)) s_ctrl.rare_count += 1;
      }
      __syncthreads();
      if (publisher) {
        // CTA 0 publishes the decision off the critical path (the mapped-host-memory write and its fence
        // cost microseconds): controller state, progress word
        *a.ctrl = s_ctrl;
        if (a.host_progress != nullptr) {
          a.host_progress[0] = s_ctrl.done ? s_ctrl.final_iter + 1u : s_ctrl.k;
          a.host_progress[1] = static_cast<unsigned>(s_ctrl.done);
          __threadfence_system();
        }
      }
    }
    lap(4);
    if (s_ctrl.done || s_ctrl.need_exact || s_ctrl.need_solve) {
      // no tail: wait for the copies that were started for phase D before leaving
      if (prefiller) for (unsigned s = 0; s < preD; ++s) mbar_wait(&shD.full[s], 0u);
      return;
    }
  }

  // ================= phase D: x'' = M u from the packed lower triangle, streamed once =================
  // The x half-step that follows belongs to the iteration AFTER the one phase C has just finished (mode 0;
  // the controller has already advanced k) or to the iteration in hand (mode 1): in both cases it is the
  // `xnext` of the parity the iteration in hand had at the start of the launch in mode 0, of the other one
  // in mode 1.
  {
    const T rho = s_ctrl.rho;
    const ParityArgs<T>& pt = s_par[a.mode == 0 ? p : 1 - p];
    const SymRowOp<T> sop{a.xrow};
    // (preD is known to the prefiller thread only: every thread derives it)
    const unsigned pre = a.mode == 0 ? (nrowsM < nslots ? nrowsM : nslots) : 0u;
    stream_rows<T, false, NV, B, true>(shD, smem_raw, rowsM, nrowsM, row_bytes, nslots, W, a.u, nvec, a.u, sop, rho, my_colpart,
                                       pre, tid);
    lap(5);
    if (!grid_barrier(a.bar, gridDim.x)) return;
    lap(6);
    // ================= phase E: fold, [exchange,] x half-step =================
    if (bid < a.nfold) {
      const XStateColOp<T> xop{pt.xnext, a.xs_part};
      fold_columns<T>(a.colpart, ld, nvec, a.n, gridDim.x, a.fold_vecs, a.xrow, wtot, gridDim.x, xop, rho, pv,
                      kPassEChannel + static_cast<int>(bid), 1, s_fold, s_tot, s_rx, static_cast<double*>(nullptr), tid, bid);
    }
    lap(7);
  }
}

// First-half-step y sums on the rare path (speculation discarded: k_prox recomputed the half-step):
// fold the y blocks of k_prox's partials, sum over the ranks, leave them where phase C reads them.
static __global__ void __launch_bounds__(kThreads)
k_ysum_first(const double* __restrict__ prox_part, unsigned gx, unsigned gy, double* __restrict__ ysum_cur, Gate gate,
             PeerView pv) {
  if (gate_closed(gate)) return;
  double v[3];
  fold_partials_multi<3>(prox_part + static_cast<size_t>(gx) * 3, gy, 3, v);
  peer_sum_scalars<3>(pv, v);
  if (threadIdx.x < 3) ysum_cur[threadIdx.x] = v[threadIdx.x];
}

// Defined in admm_pass_inst.cu (explicit instantiations for float and double).
template <typename T>
void launch_admm_pass(int nv, int batch, unsigned grid, size_t smem, cudaStream_t st, const PassArgs<T>& a,
                      const ParityArgs<T>& par0, const ParityArgs<T>& par1, Gate gate, const PeerView& pv, bool pdl);

// End of the service kernels of a round: the factor apply the pass was waiting for has run (or was not
// needed); count the round for the host, which keeps a few rounds queued.
template <typename T>
__global__ void k_service_done(Ctrl<T>* __restrict__ c, volatile unsigned* host_progress) {
  if (threadIdx.x != 0) return;
  if (!c->done && !c->need_exact) c->need_solve = 0;
  c->rounds += 1;
  if (host_progress != nullptr) { host_progress[2] = c->rounds; __threadfence_system(); }
}

}  // namespace pogs_b200
