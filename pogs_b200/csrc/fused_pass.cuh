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
  T* colpart;                                     // [2*gridDim.x][ld] column sums per half CTA
  unsigned* bar;                                  // grid barrier counter (monotone)
  double* ys_part;                                // [2*gridDim.x][2]
  double* spec_part;                              // [nfold + 2*gridDim.x][3]: x rows then y rows
  unsigned nfold;                                 // CTAs taking part in the fold phase
  unsigned fold_vecs;                             // 16 B column vectors per fold CTA (power of two, <= 128)
  unsigned nstages;                               // stages per half CTA
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

// Named barrier over one half of the CTA (256 threads); ids 1 and 2 (0 is __syncthreads).
__device__ __forceinline__ void half_sync(int half) {
  asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "r"(kFusedThreads / 2) : "memory");
}

// The CTA is split into two independent halves of 256 threads; each half streams its own
// row groups (even / odd) through its own ring of stages, so that while one half sits in the
// short serial part of a group (reduce the dot product -> row-local map -> broadcast the
// coefficient) the other half is loading, multiplying or accumulating.  NV = 16 B column
// vectors per thread per row (a half covers a whole row), RS = rows per group.
template <typename T, int NV, int RS>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_fused_pass(FusedArgs<T> a, const Ctrl<T>* __restrict__ ctrl, Gate gate, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  constexpr int HT = kFusedThreads / 2, HW = HT / 32;
  if (gate_closed(gate)) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_full[2][8];
  __shared__ T s_dot[2][2][HW][RS];
  __shared__ T s_coef[2][RS];
  __shared__ double s_red[2][5];

  const int tid = threadIdx.x, half = tid / HT, htid = tid % HT, lane = tid & 31, hw = htid >> 5;
  const size_t ld = a.ld, nvec = ld / VEC;
  const unsigned row_bytes = static_cast<unsigned>(ld * sizeof(T));
  const unsigned stage_bytes = row_bytes * RS;
  const unsigned nst = a.nstages;                 // stages per half
  const T rho = ctrl->rho;
  // shared memory: [x copy (row_bytes)] [half 0 stages] [half 1 stages]
  VT* xs = reinterpret_cast<VT*>(smem_raw);
  unsigned char* ring = smem_raw + row_bytes + static_cast<size_t>(half) * nst * stage_bytes;

  // rows of this CTA, in groups of RS; half h owns groups h, h+2, ...
  const size_t rows_per_cta = (a.m + gridDim.x - 1) / gridDim.x;
  const size_t r0 = static_cast<size_t>(blockIdx.x) * rows_per_cta;
  const size_t r1 = r0 + rows_per_cta < a.m ? r0 + rows_per_cta : a.m;
  const size_t nrows = r1 > r0 ? r1 - r0 : 0;
  const size_t ngroups = (nrows + RS - 1) / RS;
  const size_t my_groups = ngroups > static_cast<size_t>(half) ? (ngroups - half + 1) / 2 : 0;

  for (size_t jv = tid; jv < nvec; jv += kFusedThreads) xs[jv] = __ldg(reinterpret_cast<const VT*>(a.xnew) + jv);
  VT acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = zerov(static_cast<VT*>(nullptr));

  auto issue = [&](size_t j) {   // thread 0 of the half: fill the stage of its j-th group
    const unsigned s = static_cast<unsigned>(j % nst);
    const size_t row = r0 + (2 * j + half) * RS;
    const unsigned rows_here = static_cast<unsigned>(row + RS <= r1 ? RS : r1 - row);
    mbar_expect_tx(&s_full[half][s], rows_here * row_bytes);
    bulk_g2s(ring + static_cast<size_t>(s) * stage_bytes, a.A + row * ld, rows_here * row_bytes, &s_full[half][s]);
  };

  if (htid == 0) {
    for (unsigned s = 0; s < nst; ++s) mbar_init(&s_full[half][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (htid == 0) {
    for (size_t j = 0; j < my_groups && j < nst; ++j) issue(j);
  }

  double red_s = 0, red_r = 0, red_wz = 0, red_ww = 0, red_zz = 0;   // held by threads htid < RS

  for (size_t j = 0; j < my_groups; ++j) {
    const unsigned s = static_cast<unsigned>(j % nst);
    const unsigned parity = static_cast<unsigned>((j / nst) & 1u);
    const size_t row = r0 + (2 * j + half) * RS;
    const int rows_here = static_cast<int>(row + RS <= r1 ? RS : r1 - row);
    // row state for the threads that run the row-local map (issued early: hidden behind the wait)
    T zp = 0, zh = 0, ti = 0, fa = 1, fb = 0, fc = 0, fd = 0, fe = 0;
    int fh = kZero;
    if (htid < rows_here) {
      const size_t i = row + htid;
      zp = a.yprev[i]; zh = a.y12[i]; ti = a.ty[i];
      fh = a.f.h[i]; fa = a.f.a[i]; fb = a.f.b[i]; fc = a.f.c[i]; fd = a.f.d[i]; fe = a.f.e[i];
    }
    mbar_wait(&s_full[half][s], parity);
    const unsigned char* stage = ring + static_cast<size_t>(s) * stage_bytes;
    VT av[RS][NV];
    T d[RS];
#pragma unroll
    for (int r = 0; r < RS; ++r) d[r] = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const size_t jv = static_cast<size_t>(htid) + static_cast<size_t>(k) * HT;
      const bool in = jv < nvec;
      const VT x = in ? xs[jv] : zerov(static_cast<VT*>(nullptr));
#pragma unroll
      for (int r = 0; r < RS; ++r) {
        av[r][k] = (in && r < rows_here)
                       ? *reinterpret_cast<const VT*>(stage + static_cast<size_t>(r) * row_bytes + jv * sizeof(VT))
                       : zerov(static_cast<VT*>(nullptr));
        d[r] += dotv<false>(av[r][k], x);
      }
    }
#pragma unroll
    for (int r = 0; r < RS; ++r) {
      const T dd = warp_sum(d[r]);
      if (lane == 0) s_dot[half][j & 1][hw][r] = dd;
    }
    half_sync(half);   // stage is in registers (hand it back to the copy engine) + partial dots visible
    if (htid == 0 && j + nst < my_groups) issue(j + nst);
    if (htid < rows_here) {
      double tot = 0;
#pragma unroll
      for (int w = 0; w < HW; ++w) tot += static_cast<double>(s_dot[half][j & 1][w][htid]);
      const size_t i = row + htid;
      // ---- iteration k, second half-step for row i (EpiState) --------------------------------------
      const T yn = static_cast<T>(tot);
      const T ztn = ti - yn;
      a.ynew[i] = yn;
      a.yt_next[i] = ztn;
      const double ds = static_cast<double>(zp) - static_cast<double>(yn);
      const double dr = static_cast<double>(zh) - static_cast<double>(yn);
      red_s += ds * ds;
      red_r += dr * dr;
      // ---- iteration k+1, first half-step for row i, assuming rho and z~ scale unchanged ----------
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
      s_coef[half][htid] = t2;
    }
    half_sync(half);
#pragma unroll
    for (int r = 0; r < RS; ++r) {
      if (r < rows_here) {
        const T c = s_coef[half][r];
#pragma unroll
        for (int k = 0; k < NV; ++k) fmav<false>(acc[k], av[r][k], c);
      }
    }
    // s_coef is rewritten only after the next half_sync, s_dot[j&1] two groups later: no extra barrier
  }

  // column sums of this half
  const size_t prow = static_cast<size_t>(blockIdx.x) * 2 + half;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const size_t jv = static_cast<size_t>(htid) + static_cast<size_t>(k) * HT;
    if (jv < nvec) reinterpret_cast<VT*>(a.colpart + prow * ld)[jv] = acc[k];
  }
  // per-half reductions (threads htid < RS hold them), folded in fixed order
  if (htid == 0) { s_red[half][0] = 0; s_red[half][1] = 0; s_red[half][2] = 0; s_red[half][3] = 0; s_red[half][4] = 0; }
  half_sync(half);
  for (int r = 0; r < RS; ++r) {
    if (htid == r) {
      s_red[half][0] += red_s; s_red[half][1] += red_r; s_red[half][2] += red_wz; s_red[half][3] += red_ww; s_red[half][4] += red_zz;
    }
    half_sync(half);
  }
  if (htid == 0) {
    a.ys_part[prow * 2 + 0] = s_red[half][0];
    a.ys_part[prow * 2 + 1] = s_red[half][1];
    double* sp = a.spec_part + (static_cast<size_t>(a.nfold) + prow) * 3;
    sp[0] = s_red[half][2]; sp[1] = s_red[half][3]; sp[2] = s_red[half][4];
  }
  const unsigned nparts = gridDim.x * 2;   // rows of colpart

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
