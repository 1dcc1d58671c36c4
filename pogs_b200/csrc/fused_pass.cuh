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
// Data movement: each CTA (one per SM: 16 streaming warps + 2 map warps) streams its block of
// rows through a ring of shared-memory slots filled by 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx, the TMA engine); a row is read from HBM once and
// from shared memory twice (dot product, column update).  See k_fused_pass for the pipeline.
#pragma once

#include "kernels.cuh"

namespace pogs_b200 {

constexpr int kFusedThreads = 512;                  // threads that stream the rows ("main" warps)
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMapWarps = 4;                   // warps that can run the row-local maps and refill the ring
constexpr int kFusedCta = kFusedThreads + 32 * kFusedMapWarps;

// Launch-independent arguments of the pass.
template <typename T>
struct OnePassArgs {
  const T* A; size_t m, n, ld;        // local row block
  const T* x;                         // multiplied vector (zero-padded to ld)
  T* colpart;                         // [gridDim.x][ld] column sums per CTA
  unsigned* bar;                      // grid barrier counter (monotone)
  unsigned nfold;                     // CTAs taking part in the fold phase
  unsigned fold_vecs;                 // 16 B column vectors per fold CTA (power of two, <= 128)
  unsigned nstages;                   // ring slots (one row each), <= 32
  unsigned nmap;                      // active map warps W (1..kFusedMapWarps): W batches are in the pipeline
};

// What happens to a row once its dot product d_i = A_i . x is known, and to a column once its
// sum c_j = sum_i coef_i * A_ij is known, is supplied by two functors:
//
//   RowOp:  struct State;  load(i, State&)            row-local inputs, issued before the waits
//           T apply(i, State, d_i, rho, red[NRED])    row-local map; returns coef_i
//           store(cta, nfold, red[NRED])              per-CTA sums of the NRED reduction terms
//   ColOp:  apply(j, c_j, rho, red[NRED])             column-local map
//           store(cta, red[NRED])
//
// The ADMM iteration (below), the Sinkhorn-Knopp sweep and the power iteration of the setup
// (mat_algos.cuh) are three such pairs; with SQ the entries are squared in registers.

// ADMM, y side: finishes iteration k for row i (EpiState arithmetic) and runs iteration k+1's
// first half-step for it under the assumption "rho and the z~ scale unchanged".
template <typename T>
struct AdmmRowOp {
  static constexpr int NRED = 5;      // |yprev-y|^2, |y12-y|^2, <w,z12>, |w|^2, |z12|^2
  const T* yprev; const T* y12; const T* ty;      // iteration k
  T* ynew; T* yt_next;                            // y^{k+1}, z~_y^{k+1} (unscaled)
  Desc<T> f;
  T* y12n; T* tyn; T* qyn;                        // speculative iteration k+1
  T alpha;
  double* ys_part;                                // [gridDim.x][2]
  double* spec_part;                              // [nfold + gridDim.x][3]: x rows then y rows
  T zsc = T(1);                                   // z~ scale the speculation assumes (1: "rho unchanged")
  struct State { T zp, zh, ti, fa, fb, fc, fd, fe; int fh; };
  __device__ __forceinline__ void load(size_t i, State& s) const {
    s.zp = yprev[i]; s.zh = y12[i]; s.ti = ty[i];
    s.fh = f.h[i]; s.fa = f.a[i]; s.fb = f.b[i]; s.fc = f.c[i]; s.fd = f.d[i]; s.fe = f.e[i];
  }
  __device__ __forceinline__ T apply(size_t i, const State& s, T yn, T rho, double (&red)[NRED]) const {
    const T ztn = s.ti - yn;
    ynew[i] = yn;
    yt_next[i] = ztn;
    const double ds = static_cast<double>(s.zp) - static_cast<double>(yn);
    const double dr = static_cast<double>(s.zh) - static_cast<double>(yn);
    red[0] += ds * ds;
    red[1] += dr * dr;
    const T zts = zsc * ztn;                      // == k_prox: z~ = zt_scale * stored z~
    const T v = yn - zts;
    const T zh2 = prox_eval<T>(s.fh, s.fa, s.fb, s.fc, s.fd, s.fe, v, rho);
    const T w = v - zh2;
    T t2 = zts + alpha * zh2;
    t2 += (T(1) - alpha) * yn;
    y12n[i] = zh2;
    tyn[i] = t2;
    qyn[i] = (zh2 + zts) - yn;
    const double wd = w, zd = zh2;
    red[2] += wd * zd;
    red[3] += wd * wd;
    red[4] += zd * zd;
    return t2;
  }
  __device__ __forceinline__ void store(unsigned cta, unsigned nfold, const double* red) const {
    ys_part[static_cast<size_t>(cta) * 2 + 0] = red[0];
    ys_part[static_cast<size_t>(cta) * 2 + 1] = red[1];
    double* sp = spec_part + (static_cast<size_t>(nfold) + cta) * 3;
    sp[0] = red[2]; sp[1] = red[3]; sp[2] = red[4];
  }
};

// ADMM, x side of the speculative half-step: u' = t_x' + A^T t_y'.
template <typename T>
struct AdmmColOp {
  static constexpr int NRED = 3;
  const T* xnew;                                  // x^{k+1}
  const T* xt_next;                               // z~_x^{k+1} (unscaled), written by the factor apply
  Desc<T> g;
  T* x12n; T* txn; T* qxn; T* u_out;
  T alpha;
  double* spec_part;
  T zsc = T(1);                                   // z~ scale the speculation assumes
  __device__ __forceinline__ void apply(size_t j, T total, T rho, double (&red)[NRED]) const {
    const T xk = xnew[j];
    const T zt = zsc * xt_next[j];
    const T v = xk - zt;
    const T zh2 = prox_eval<T>(g.h[j], g.a[j], g.b[j], g.c[j], g.d[j], g.e[j], v, rho);
    const T w = v - zh2;
    T t2 = zt + alpha * zh2;
    t2 += (T(1) - alpha) * xk;
    x12n[j] = zh2;
    txn[j] = t2;
    qxn[j] = (zh2 + zt) - xk;
    u_out[j] = t2 + total;
    const double wd = w, zd = zh2;
    red[0] += wd * zd; red[1] += wd * wd; red[2] += zd * zd;
  }
  __device__ __forceinline__ void store(unsigned cta, const double* red) const {
    double* sp = spec_part + static_cast<size_t>(cta) * 3;
    sp[0] = red[0]; sp[1] = red[1]; sp[2] = red[2];
  }
};

// Exact residuals (pogs.cpp:353-376) in one pass over A instead of two products:
//   rows:    r_i = A_i . x12 - y12_i          -> |r|^2,      and q_y,i as the column coefficient,
//   columns: s_j = q_x,j + (A^T q_y)_j        -> |s|^2.
template <typename T>
struct ExactRowOp {
  static constexpr int NRED = 1;
  const T* y12; const T* qy; double* er_part;     // [gridDim.x]
  struct State { T yh, q; };
  __device__ __forceinline__ void load(size_t i, State& s) const { s.yh = y12[i]; s.q = qy[i]; }
  __device__ __forceinline__ T apply(size_t, const State& s, T dot, T, double (&red)[NRED]) const {
    const double r = static_cast<double>(dot) - static_cast<double>(s.yh);
    red[0] += r * r;
    return s.q;
  }
  __device__ __forceinline__ void store(unsigned cta, unsigned, const double* red) const { er_part[cta] = red[0]; }
};
template <typename T>
struct ExactColOp {
  static constexpr int NRED = 1;
  const T* qx; double* es_part;                   // [nfold]
  __device__ __forceinline__ void apply(size_t j, T total, T, double (&red)[NRED]) const {
    const double v = static_cast<double>(qx[j]) + static_cast<double>(total);
    red[0] += v * v;
  }
  __device__ __forceinline__ void store(unsigned cta, const double* red) const { es_part[cta] = red[0]; }
};

// Sinkhorn-Knopp (equil_helper.h:149-163) with B = A.^2:  d_i = nd / (B_i . e + cd) for the rows,
// then e_j = ne / ((B^T d)_j + ce) for the columns: the d-update of one sweep and the e-update of
// the next in one pass over A.
template <typename T>
struct SinkhornRowOp {
  static constexpr int NRED = 1;
  T num, cst; T* d;
  struct State {};
  __device__ __forceinline__ void load(size_t, State&) const {}
  __device__ __forceinline__ T apply(size_t i, const State&, T dot, T, double (&)[NRED]) const {
    const T di = num / (dot + cst);
    d[i] = di;
    return di;
  }
  __device__ __forceinline__ void store(unsigned, unsigned, const double*) const {}
};
template <typename T>
struct SinkhornColOp {
  static constexpr int NRED = 1;
  T num, cst; T* e;
  __device__ __forceinline__ void apply(size_t j, T total, T, double (&)[NRED]) const { e[j] = num / (total + cst); }
  __device__ __forceinline__ void store(unsigned, const double*) const {}
};

// Power iteration on A^T A (equil_helper.h:119-131): Sx = A (x * inv), x' = A^T Sx, with the
// norms |Sx|^2 (rows) and |x'|^2 (columns); `inv` = 1/|x| of the previous sweep lives on the device.
template <typename T>
struct PowerRowOp {
  static constexpr int NRED = 1;
  const T* inv; double* part;         // [gridDim.x]
  struct State {};
  __device__ __forceinline__ void load(size_t, State&) const {}
  __device__ __forceinline__ T apply(size_t, const State&, T dot, T, double (&red)[NRED]) const {
    const T sx = dot * (*inv);
    red[0] += static_cast<double>(sx) * static_cast<double>(sx);
    return sx;
  }
  __device__ __forceinline__ void store(unsigned cta, unsigned, const double* red) const { part[cta] = red[0]; }
};
template <typename T>
struct PowerColOp {
  static constexpr int NRED = 1;
  T* xn; double* part;                // [nfold]
  __device__ __forceinline__ void apply(size_t j, T total, T, double (&red)[NRED]) const {
    xn[j] = total;
    red[0] += static_cast<double>(total) * static_cast<double>(total);
  }
  __device__ __forceinline__ void store(unsigned cta, const double* red) const { part[cta] = red[0]; }
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Waits are bounded (~2 s): a protocol bug must end in a failed launch, not in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
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

// Rows are processed in batches of B rows.  The 16 main warps never meet at a CTA-wide barrier
// inside the loop; they and the W map warps hand work to each other through mbarriers (index =
// batch mod W):
//
//   main warps, batch b :  wait for the rows (TMA mbarrier) -> partial dot products -> s_dot
//                          -> arrive dots[b%W];   then, for batch b-W+1: wait coef[..] ->
//                          column update from the rows still in shared memory -> arrive free[..]
//   map warp b%W, batch b: wait dots[b%W] -> finish the dot products, run the B row-local maps in
//                          its lanes (prox, reductions, stores) -> s_coef -> arrive coef[b%W];
//                          wait free[b%W] -> hand the batch's ring slots back to the copy engine
//
// so the serial row-local map of one batch runs while the main warps are already streaming the
// next W-1 (it used to cost ~900 cycles per batch with every other warp parked at a barrier:
// 40 % of all stall samples, and far more with an iterative prox such as the logistic one, which
// needs several map warps to keep up with HBM).  W*B ring slots are held by the batches in the
// pipeline, the rest stays in flight.
// NV = 16 B column vectors per thread per row, B = rows per batch.
template <typename T, bool SQ, int NV, int B, typename RowOp, typename ColOp>
__global__ void __launch_bounds__(kFusedCta, 1)
k_fused_pass(OnePassArgs<T> a, RowOp rop, ColOp cop, const Ctrl<T>* __restrict__ ctrl, Gate gate, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  if (gate_closed(gate)) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_full[32];                       // copy engine -> main warps, one per ring slot
  __shared__ uint64_t s_dots[kFusedMapWarps], s_coefr[kFusedMapWarps], s_free[kFusedMapWarps];  // index = batch mod W
  __shared__ T s_dot[kFusedMapWarps][kFusedWarps][B];
  __shared__ T s_coef[kFusedMapWarps][B];
  constexpr int RN = RowOp::NRED, CN = ColOp::NRED;
  __shared__ double s_red[kFusedMapWarps][RN];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool main_thr = tid < kFusedThreads;
  const size_t ld = a.ld, nvec = ld / VEC;
  const unsigned row_bytes = static_cast<unsigned>(ld * sizeof(T));
  const unsigned nslots = a.nstages;              // ring slots, one row each
  const T rho = ctrl != nullptr ? ctrl->rho : T(0);

  const size_t rows_per_cta = (a.m + gridDim.x - 1) / gridDim.x;
  const size_t r0 = static_cast<size_t>(blockIdx.x) * rows_per_cta;
  const size_t r1 = r0 + rows_per_cta < a.m ? r0 + rows_per_cta : a.m;
  const size_t nrows = r1 > r0 ? r1 - r0 : 0;
  const size_t nbatch = (nrows + B - 1) / B;
  const unsigned W = a.nmap;

  if (tid == 0) {
    for (unsigned s = 0; s < nslots; ++s) mbar_init(&s_full[s], 1);
    for (int p = 0; p < kFusedMapWarps; ++p) {
      mbar_init(&s_dots[p], kFusedWarps);
      mbar_init(&s_coefr[p], 1);
      mbar_init(&s_free[p], kFusedWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  VT acc[NV];   // main threads: column accumulators
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = zerov(static_cast<VT*>(nullptr));

  if (main_thr) {
    // ================= main warps =================
    VT xv[NV];   // this thread's slice of x
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
      xv[k] = jv < nvec ? __ldg(reinterpret_cast<const VT*>(a.x) + jv) : zerov(static_cast<VT*>(nullptr));
    }
    // All indices are 32-bit and advanced incrementally: an integer division per batch and warp
    // costs more than the batch's arithmetic (measured: 0.58 -> 0.66 ms per pass with three
    // 64-bit divisions in this loop).
    const unsigned nbt = static_cast<unsigned>(nbatch), nrw = static_cast<unsigned>(nrows);
    unsigned slot = 0, phase = 0;        // ring slot / mbarrier parity of the first row of batch bi
    unsigned wb = 0;                     // bi % W
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
            mbar_wait(&s_full[s], ph);
            const VT* rowp = reinterpret_cast<const VT*>(smem_raw + static_cast<size_t>(s) * row_bytes);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
              if (jv < nvec) d[b] += dotv<SQ>(rowp[jv], xv[k]);
            }
            if (++s == nslots) { s = 0; ph ^= 1u; }
          }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const T dd = warp_sum(d[b]);
          if (lane == 0) s_dot[wb][warp][b] = dd;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_dots[wb]);
        slot = s; phase = ph;
        if (++wb == W) wb = 0;
      }
      if (bi + 1 >= W && pb < nbt) {
        // ---- column update of batch pb = bi-W+1 from the rows still in shared memory ----------------
        const unsigned left = nrw - pb * B;
        const int nbp = static_cast<int>(left < static_cast<unsigned>(B) ? left : B);
        mbar_wait(&s_coefr[wp], parp);
        unsigned s = uslot;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          if (b < nbp) {
            const T c = s_coef[wp][b];
            const VT* rowp = reinterpret_cast<const VT*>(smem_raw + static_cast<size_t>(s) * row_bytes);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
              if (jv < nvec) fmav<SQ>(acc[k], rowp[jv], c);
            }
            if (++s == nslots) s = 0;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wp]);
        uslot = s;
        ++pb;
        if (++wp == W) { wp = 0; parp ^= 1u; }
      }
    }
  } else {
    // ================= map warps: warp j < W owns the batches b with b % W == j =================
    const int j = warp - kFusedWarps;
    double red[RN];   // held by lanes < B
#pragma unroll
    for (int k = 0; k < RN; ++k) red[k] = 0;
    // first fill of the ring (map warp 0)
    if (j == 0 && lane == 0) {
      for (size_t r = 0; r < nrows && r < nslots; ++r) {
        mbar_expect_tx(&s_full[r], row_bytes);
        bulk_g2s(smem_raw + r * row_bytes, a.A + (r0 + r) * ld, row_bytes, &s_full[r]);
      }
    }
    const unsigned nbt = static_cast<unsigned>(nbatch), nrw = static_cast<unsigned>(nrows);
    const unsigned step = W * B;                       // rows between two batches of this warp (<= nslots)
    unsigned par = 0;                                  // (bi / W) & 1
    unsigned mslot = (static_cast<unsigned>(j) * B) % nslots;   // ring slot of the first row of batch bi (one division per launch)
    for (unsigned bi = j; static_cast<unsigned>(j) < W && bi < nbt; bi += W) {
      const unsigned row = bi * B, left = nrw - row;
      const int nb = static_cast<int>(left < static_cast<unsigned>(B) ? left : B);
      // row state for the lanes that run the row-local maps (issued early: hidden behind the wait)
      typename RowOp::State rs{};
      if (lane < nb) rop.load(r0 + row + lane, rs);
      mbar_wait(&s_dots[j], par);
      if (lane < nb) {
        double tot = 0;
#pragma unroll
        for (int w = 0; w < kFusedWarps; ++w) tot += static_cast<double>(s_dot[j][w][lane]);
        s_coef[j][lane] = rop.apply(r0 + row + lane, rs, static_cast<T>(tot), rho, red);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_coefr[j]);
      // the batch has been read twice once every main warp has updated its columns: refill its slots
      mbar_wait(&s_free[j], par);
      if (lane == 0) {
        unsigned sl = mslot;
        for (int b = 0; b < nb; ++b) {
          const unsigned r = row + b + nslots;       // row that takes over the slot of row bi*B + b
          if (r < nrw) {
            mbar_expect_tx(&s_full[sl], row_bytes);
            bulk_g2s(smem_raw + static_cast<size_t>(sl) * row_bytes, a.A + (r0 + r) * ld, row_bytes, &s_full[sl]);
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
      for (int k = 0; k < RN; ++k) s_red[j][k] = 0;
    }
    __syncwarp();
    for (int b = 0; b < B; ++b) {
      if (lane == b) {
#pragma unroll
        for (int k = 0; k < RN; ++k) s_red[j][k] += red[k];
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // column sums of this CTA
  if (main_thr) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const size_t jv = static_cast<size_t>(tid) + static_cast<size_t>(k) * kFusedThreads;
      if (jv < nvec) reinterpret_cast<VT*>(a.colpart + static_cast<size_t>(blockIdx.x) * ld)[jv] = acc[k];
    }
  }
  if (tid == 0) {
    double tot[RN];
#pragma unroll
    for (int k = 0; k < RN; ++k) {
      tot[k] = 0;
#pragma unroll
      for (int w = 0; w < kFusedMapWarps; ++w) tot[k] += s_red[w][k];   // fixed order
    }
    rop.store(blockIdx.x, a.nfold, tot);
  }
  const unsigned nparts = gridDim.x;   // rows of colpart

  // ---- second phase: fold the column sums over the CTAs, add the speculative x half-step -------------
  if (!grid_barrier(a.bar, gridDim.x)) return;
  if (blockIdx.x >= a.nfold) return;
  __shared__ VT s_fold[kFusedThreads];
  __shared__ double s_rx[128][CN];
  const unsigned FV = a.fold_vecs, NG = kFusedThreads / FV;   // NG groups of partials x FV vectors
  const unsigned v16 = tid & (FV - 1), grp = tid / FV;
  const size_t jv = static_cast<size_t>(blockIdx.x) * FV + v16;
  VT part = zerov(static_cast<VT*>(nullptr));
  if (main_thr && jv < nvec) {
    for (unsigned p = grp; p < nparts; p += NG)
      addv(part, ld_cg(reinterpret_cast<const VT*>(a.colpart + static_cast<size_t>(p) * ld) + jv));
  }
  if (main_thr) s_fold[grp * FV + v16] = part;
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
  double rx[CN];
#pragma unroll
  for (int k = 0; k < CN; ++k) rx[k] = 0;
  if (fin) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const size_t j = jv * VEC + e;
      if (j < a.n) cop.apply(j, elemv(total, e), rho, rx);
    }
  }
  if (static_cast<unsigned>(tid) < FV) {
#pragma unroll
    for (int k = 0; k < CN; ++k) s_rx[tid][k] = rx[k];
  }
  __syncthreads();
  if (tid == 0) {
    double t[CN];
#pragma unroll
    for (int k = 0; k < CN; ++k) t[k] = 0;
    for (unsigned q = 0; q < FV; ++q) {   // fixed order
#pragma unroll
      for (int k = 0; k < CN; ++k) t[k] += s_rx[q][k];
    }
    cop.store(blockIdx.x, t);
  }
}

}  // namespace pogs_b200
