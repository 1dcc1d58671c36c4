// Single pass over A per ADMM iteration (dense, row-major, m > n, direct projector).
//
// The reference iteration touches A twice: u = t_x + A^T t_y, then y = A x
// (projector_direct_dense.cpp:122-127).  t_y of the NEXT iteration depends on y
// only row by row -- y_i -> dual update -> prox_f -> over-relaxation are all
// separable (pogs.cpp:254-278, 397-399) -- *given* rho and the rescaling factor
// of z~, which the controller fixes after it has seen the norms of the whole
// iteration (pogs.cpp:402-466).  In the large majority of iterations neither
// changes.  So this kernel, while row i of A is on chip for y_i = A_i x, also
//   * finishes iteration k for that row   (y, z~_y, the two residual terms),
//   * speculatively runs iteration k+1's first half-step for it, assuming
//     "rho unchanged, not converged": y12', t_y', q_y' and the prox reductions,
//   * accumulates t_y' * A_i into per-thread column sums,
// and after a grid barrier folds the column sums, adds the (equally speculative)
// x half-step t_x' and leaves u' = t_x' + A^T t_y' ready for the next factor
// apply.  The controller then either commits the speculation (the next
// iteration skips k_prox and the A^T pass: ONE pass over A instead of two) or
// discards it and the regular two-pass kernels recompute the same quantities.
// Either way every iterate equals the reference sequence up to rounding.
//
// Data movement: each CTA (512 threads, one per SM) streams its block of rows
// through a ring of shared-memory stages filled by 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx, the TMA engine), 3-7 stages in
// flight; a thread copies its columns of the staged rows to registers once and
// uses them for both the dot product and the column update, so A is read from
// HBM once and from shared memory once.
#pragma once

#include "kernels.cuh"

namespace pogs_b200 {

constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;

template <typename T>
struct FusedArgs {
  const T* A; size_t m, n, ld;        // local row block
  const T* xnew;                      // x^{k+1} (zero-padded to ld)
  const T* yprev; const T* y12; const T* ty;      // iteration k, y side
  T* ynew; T* yt_next;                            // y^{k+1}, z~_y^{k+1} (unscaled)
  Desc<T> f;
  T* y12n; T* tyn; T* qyn;                        // speculative iteration k+1, y side
  Desc<T> g;
  const T* xt_next;                               // z~_x^{k+1} (unscaled), written by the factor apply
  T* x12n; T* txn; T* qxn;                        // speculative iteration k+1, x side
  T* u_out;                                       // u' = t_x' + A^T t_y'
  T alpha;
  T* colpart;                                     // [gridDim.x][ld] column sums per CTA
  unsigned* bar;                                  // grid barrier counter (monotone)
  double* ys_part;                                // [gridDim.x][2]
  double* spec_part;                              // [nfold + gridDim.x][3]: x rows then y rows
  unsigned nfold;                                 // CTAs taking part in the fold phase
  unsigned fold_vecs;                             // 16 B column vectors per fold CTA (power of two, <= 128)
  unsigned nstages;                               // ring slots (one row each), <= 32
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// All CTAs of the (co-resident) grid meet here.  Monotone ticket counter, no reset.
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned nblocks) {
  __shared__ int s_bar_ok;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(bar, 1u);
    const unsigned target = (ticket / nblocks + 1u) * nblocks;
    int ok = 1;
    const long long t0 = clock64();
    while (static_cast<int>(ld_sys(bar) - target) < 0) {
      if (clock64() - t0 > 4000000000LL) { ok = 0; break; }
    }
    s_bar_ok = ok;
  }
  __syncthreads();
  __threadfence();
  return s_bar_ok != 0;
}

// Rows are processed in batches of B: while a batch sits in shared memory the CTA (1) forms
// the B dot products (one block reduction for all of them), (2) runs the B row-local maps in
// the lanes of one warp, (3) re-reads the batch from shared memory for the column update, and
// (4) hands the B slots back to the copy engine.  The serial part is paid once per batch, not
// once per row, and the remaining ring slots keep nslots-B rows in flight meanwhile.
// NV = 16 B column vectors per thread per row, B = rows per batch.
template <typename T, int NV, int B>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_fused_pass(FusedArgs<T> a, const Ctrl<T>* __restrict__ ctrl, Gate gate, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  if (gate_closed(gate)) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_full[32];
  __shared__ T s_dot[kFusedWarps][B];
  __shared__ T s_coef[B];
  __shared__ double s_red[5];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t ld = a.ld, nvec = ld / VEC;
  const unsigned row_bytes = static_cast<unsigned>(ld * sizeof(T));
  const unsigned nslots = a.nstages;              // ring slots, one row each
  const T rho = ctrl->rho;

  const size_t rows_per_cta = (a.m + gridDim.x - 1) / gridDim.x;
  const size_t r0 = static_cast<size_t>(blockIdx.x) * rows_per_cta;
  const size_t r1 = r0 + rows_per_cta < a.m ? r0 + rows_per_cta : a.m;
  const size_t nrows = r1 > r0 ? r1 - r0 : 0;

  // this thread's slice of x and its column accumulators
  VT xv[NV], acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
    xv[k] = jv < nvec ? __ldg(reinterpret_cast<const VT*>(a.xnew) + jv) : zerov(static_cast<VT*>(nullptr));
    acc[k] = zerov(static_cast<VT*>(nullptr));
  }

  if (tid == 0) {
    for (unsigned s = 0; s < nslots; ++s) mbar_init(&s_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // ring bookkeeping (uniform across the CTA): row `issued` goes to slot `slot_in`
  size_t issued = 0;
  unsigned slot_in = 0;
  if (tid == 0) {
    for (; issued < nrows && issued < nslots; ++issued) {
      mbar_expect_tx(&s_full[slot_in], row_bytes);
      bulk_g2s(smem_raw + static_cast<size_t>(slot_in) * row_bytes, a.A + (r0 + issued) * ld, row_bytes, &s_full[slot_in]);
      slot_in = slot_in + 1 == nslots ? 0 : slot_in + 1;
    }
  }

  double red_s = 0, red_r = 0, red_wz = 0, red_ww = 0, red_zz = 0;   // held by lanes < B of warp 0
  unsigned slot = 0, phase = 0;   // slot / mbarrier parity of the first row of the current batch

  for (size_t done = 0; done < nrows; done += B) {
    const int nb = static_cast<int>(nrows - done < static_cast<size_t>(B) ? nrows - done : B);
    // row state for the lanes that run the row-local maps (issued early: hidden behind the waits)
    T zp = 0, zh = 0, ti = 0, fa = 1, fb = 0, fc = 0, fd = 0, fe = 0;
    int fh = kZero;
    if (tid < nb) {
      const size_t i = r0 + done + tid;
      zp = a.yprev[i]; zh = a.y12[i]; ti = a.ty[i];
      fh = a.f.h[i]; fa = a.f.a[i]; fb = a.f.b[i]; fc = a.f.c[i]; fd = a.f.d[i]; fe = a.f.e[i];
    }
    // ---- (1) dot products of the batch --------------------------------------------------------
    T d[B];
    {
      unsigned s = slot, ph = phase;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        d[b] = 0;
        if (b < nb) {
          mbar_wait(&s_full[s], ph);
          const VT* rowp = reinterpret_cast<const VT*>(smem_raw + static_cast<size_t>(s) * row_bytes);
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
            if (jv < nvec) d[b] += dotv<false>(rowp[jv], xv[k]);
          }
          if (++s == nslots) { s = 0; ph ^= 1u; }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const T dd = warp_sum(d[b]);
      if (lane == 0) s_dot[warp][b] = dd;
    }
    __syncthreads();
    // ---- (2) row-local maps, one lane per row --------------------------------------------------
    if (tid < nb) {
      double tot = 0;
#pragma unroll
      for (int w = 0; w < kFusedWarps; ++w) tot += static_cast<double>(s_dot[w][tid]);
      const size_t i = r0 + done + tid;
      // iteration k, second half-step for row i (EpiState)
      const T yn = static_cast<T>(tot);
      const T ztn = ti - yn;
      a.ynew[i] = yn;
      a.yt_next[i] = ztn;
      const double ds = static_cast<double>(zp) - static_cast<double>(yn);
      const double dr = static_cast<double>(zh) - static_cast<double>(yn);
      red_s += ds * ds;
      red_r += dr * dr;
      // iteration k+1, first half-step for row i, assuming rho and the z~ scale unchanged
      const T v = yn - ztn;
      const T zh2 = prox_eval<T>(fh, fa, fb, fc, fd, fe, v, rho);
      const T w = v - zh2;
      T t2 = ztn + a.alpha * zh2;
      t2 += (T(1) - a.alpha) * yn;
      a.y12n[i] = zh2;
      a.tyn[i] = t2;
      a.qyn[i] = (zh2 + ztn) - yn;
      const double wd = w, zd = zh2;
      red_wz += wd * zd;
      red_ww += wd * wd;
      red_zz += zd * zd;
      s_coef[tid] = t2;
    }
    __syncthreads();
    // ---- (3) column update from the rows still in shared memory -------------------------------------
    {
      unsigned s = slot;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        if (b < nb) {
          const T c = s_coef[b];
          const VT* rowp = reinterpret_cast<const VT*>(smem_raw + static_cast<size_t>(s) * row_bytes);
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
            if (jv < nvec) fmav<false>(acc[k], rowp[jv], c);
          }
          if (++s == nslots) s = 0;
        }
      }
    }
    __syncthreads();   // (4) the batch has been read twice: its slots may be refilled
    if (tid == 0) {
      for (int b = 0; b < nb && issued < nrows; ++b, ++issued) {
        mbar_expect_tx(&s_full[slot_in], row_bytes);
        bulk_g2s(smem_raw + static_cast<size_t>(slot_in) * row_bytes, a.A + (r0 + issued) * ld, row_bytes, &s_full[slot_in]);
        slot_in = slot_in + 1 == nslots ? 0 : slot_in + 1;
      }
    }
    for (int b = 0; b < nb; ++b) {
      if (++slot == nslots) { slot = 0; phase ^= 1u; }
    }
  }

  // column sums of this CTA
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
    if (jv < nvec) reinterpret_cast<VT*>(a.colpart + static_cast<size_t>(blockIdx.x) * ld)[jv] = acc[k];
  }
  // per-CTA reductions (lanes < B of warp 0 hold them), folded in fixed order
  if (tid == 0) { s_red[0] = 0; s_red[1] = 0; s_red[2] = 0; s_red[3] = 0; s_red[4] = 0; }
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    if (tid == b) { s_red[0] += red_s; s_red[1] += red_r; s_red[2] += red_wz; s_red[3] += red_ww; s_red[4] += red_zz; }
    __syncwarp();
  }
  if (tid == 0) {
    a.ys_part[static_cast<size_t>(blockIdx.x) * 2 + 0] = s_red[0];
    a.ys_part[static_cast<size_t>(blockIdx.x) * 2 + 1] = s_red[1];
    double* sp = a.spec_part + (static_cast<size_t>(a.nfold) + blockIdx.x) * 3;
    sp[0] = s_red[2]; sp[1] = s_red[3]; sp[2] = s_red[4];
  }
  const unsigned nparts = gridDim.x;   // rows of colpart

  // ---- second phase: fold the column sums over the CTAs, add the speculative x half-step -------------
  if (!grid_barrier(a.bar, gridDim.x)) return;
  if (blockIdx.x >= a.nfold) return;
  __shared__ VT s_fold[kFusedThreads];
  __shared__ double s_rx[128][3];
  const unsigned FV = a.fold_vecs, NG = kFusedThreads / FV;   // NG groups of partials x FV vectors
  const unsigned v16 = tid & (FV - 1), grp = tid / FV;
  const size_t jv = static_cast<size_t>(blockIdx.x) * FV + v16;
  VT part = zerov(static_cast<VT*>(nullptr));
  if (jv < nvec) {
    for (unsigned p = grp; p < nparts; p += NG)
      addv(part, ld_cg(reinterpret_cast<const VT*>(a.colpart + static_cast<size_t>(p) * ld) + jv));
  }
  s_fold[grp * FV + v16] = part;
  __syncthreads();
  VT total = zerov(static_cast<VT*>(nullptr));
  const bool fin = static_cast<unsigned>(tid) < FV && jv < nvec;   // threads that finish a column vector
  if (fin) {
    for (unsigned q = 0; q < NG; ++q) addv(total, s_fold[q * FV + tid]);   // fixed order
  }
  if (pv.active()) {
    // row blocks: this is one rank's share of A^T t_y'; sum the shares over NVLink peer memory
    const unsigned seq = *pv.seq(blockIdx.x) + 1u;
    if (fin) reinterpret_cast<VT*>(pv.spec(pv.rank, seq))[jv] = total;
    peer_signal_wait(pv, blockIdx.x, seq);
    if (fin) {
      VT share[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) share[r] = ld_peer(reinterpret_cast<const VT*>(pv.spec(r, seq)) + jv);
      total = zerov(static_cast<VT*>(nullptr));
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) addv(total, share[r]);
    }
    if (tid == 0) *pv.seq(blockIdx.x) = seq;
  }
  double rx[3] = {0, 0, 0};
  if (fin) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const size_t j = jv * VEC + e;
      if (j < a.n) {
        const T xk = a.xnew[j];
        const T zt = a.xt_next[j];
        const T v = xk - zt;
        const T zh2 = prox_eval<T>(a.g.h[j], a.g.a[j], a.g.b[j], a.g.c[j], a.g.d[j], a.g.e[j], v, rho);
        const T w = v - zh2;
        T t2 = zt + a.alpha * zh2;
        t2 += (T(1) - a.alpha) * xk;
        a.x12n[j] = zh2;
        a.txn[j] = t2;
        a.qxn[j] = (zh2 + zt) - xk;
        a.u_out[j] = t2 + elemv(total, e);
        const double wd = w, zd = zh2;
        rx[0] += wd * zd; rx[1] += wd * wd; rx[2] += zd * zd;
      }
    }
  }
  if (static_cast<unsigned>(tid) < FV) { s_rx[tid][0] = rx[0]; s_rx[tid][1] = rx[1]; s_rx[tid][2] = rx[2]; }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (unsigned q = 0; q < FV; ++q) { t0 += s_rx[q][0]; t1 += s_rx[q][1]; t2 += s_rx[q][2]; }   // fixed order
    double* sp = a.spec_part + static_cast<size_t>(blockIdx.x) * 3;
    sp[0] = t0; sp[1] = t1; sp[2] = t2;
  }
}

}  // namespace pogs_b200
