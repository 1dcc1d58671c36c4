// sm_100a kernels of the graph-form ADMM hot path.
//
// Everything per-iteration is HBM-bound (arithmetic intensity ~0.5 flop/B), so
// the design rules are: stream the matrix exactly once per product with 128-bit
// loads that bypass L1 (the re-used vector stays L1-resident), keep >= 8
// independent 16 B loads in flight per thread, size grids as one full wave of
// the 148 SMs, fuse every elementwise update and every norm/dot of the
// reference loop (src/cpu/pogs.cpp:253-470) into the epilogue of the product
// that makes its input, and reduce deterministically (fixed-order partials, no
// floating-point atomics) so that two runs are bit-identical.
//
//   k_rowdot  : out[r] = sum_c M[r][c]*v[c]    (row-major M; one warp per row)
//   k_colacc  : out[c] = sum_r M[r][c]*w[r]    (same storage; column tiles x row
//               chunks, the last chunk CTA of a tile folds the partials)
//   k_prox    : first ADMM half-step: prox of f and g, over-relaxation input,
//               the five reductions of pogs.cpp:267-273
//   k_control : device-side stopping rule + adaptive rho (pogs.cpp:342-469)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "peer_comm.cuh"
#include "prox.cuh"

namespace pogs_b200 {

constexpr int kThreads = 256;          // CTA size of every streaming kernel
constexpr int kWarps = kThreads / 32;
constexpr int kMaxRed = 6;             // max fused reductions per kernel

// ---- 16-byte vector access ------------------------------------------------
template <typename T> struct V16;
template <> struct V16<float> { using type = float4; static constexpr int N = 4; };
template <> struct V16<double> { using type = double2; static constexpr int N = 2; };

// Streaming load of matrix data: read-only path, no L1 allocation, so that the
// multiplied vector (re-read by every row) is not evicted from L1.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
// L2-coherent load (cross-CTA partials written earlier in the same launch).
__device__ __forceinline__ float4 ld_cg(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ld_cg(const double2* p) { return __ldcg(p); }

template <bool SQ> __device__ __forceinline__ float dotv(float4 a, float4 x) {
  if (SQ) return a.x * a.x * x.x + a.y * a.y * x.y + a.z * a.z * x.z + a.w * a.w * x.w;
  return a.x * x.x + a.y * x.y + a.z * x.z + a.w * x.w;
}
template <bool SQ> __device__ __forceinline__ double dotv(double2 a, double2 x) {
  if (SQ) return a.x * a.x * x.x + a.y * a.y * x.y;
  return a.x * x.x + a.y * x.y;
}
template <bool SQ> __device__ __forceinline__ void fmav(float4& acc, float4 a, float w) {
  if (SQ) { acc.x += a.x * a.x * w; acc.y += a.y * a.y * w; acc.z += a.z * a.z * w; acc.w += a.w * a.w * w; }
  else    { acc.x += a.x * w; acc.y += a.y * w; acc.z += a.z * w; acc.w += a.w * w; }
}
template <bool SQ> __device__ __forceinline__ void fmav(double2& acc, double2 a, double w) {
  if (SQ) { acc.x += a.x * a.x * w; acc.y += a.y * a.y * w; }
  else    { acc.x += a.x * w; acc.y += a.y * w; }
}
__device__ __forceinline__ void addv(float4& s, float4 a) { s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
__device__ __forceinline__ void addv(double2& s, double2 a) { s.x += a.x; s.y += a.y; }
__device__ __forceinline__ float4 zerov(float4*) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ double2 zerov(double2*) { return make_double2(0., 0.); }
__device__ __forceinline__ float elemv(const float4& a, int i) { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : a.w; }
__device__ __forceinline__ double elemv(const double2& a, int i) { return i == 0 ? a.x : a.y; }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- device-resident controller state ---------------------------------------
// Scalars of PogsImplementation::Solve (pogs.cpp:198-251) kept in HBM so that
// the loop never synchronises with the host.
template <typename T>
struct Ctrl {
  // parameters
  T abs_tol, rel_tol, nrmA;
  T sqrtn_atol, sqrtm_atol, sqrtmn_atol;
  unsigned max_iter;
  int adaptive_rho, gap_stop;
  // persistent across solves
  T rho;
  // per-solve
  T delta, xi, prev_nrm_r, zt_scale;
  unsigned k, kd, ku;
  int done, converged, need_exact;
  int fused_enabled;   // the single-pass kernel runs every iteration and speculates the next half-step
  int spec_miss;       // 1: this iteration must run k_prox and the A^T pass itself (speculation absent / discarded)
  int need_solve;      // one-launch iteration (admm_pass.cuh): 1 = the factor apply of the coming iteration has not
                       // run in the tail of the previous pass (speculation discarded, exact-residual branch taken,
                       // first iteration): the rare path runs it before the pass
  unsigned rare_count; // times the rare path was armed
  unsigned rounds;     // launches of the multi-iteration kernel that have finished their iterations
  unsigned spec_hits;
  // rho-action prediction of the one-launch kernel: the speculative half-step of the next iteration is computed
  // for "the same rho action as last time" (the update rule increases rho in streaks of consecutive
  // iterations: without the prediction every iteration of a streak discards its speculation)
  int pred_act;        // action the last finished iteration took: +1 rho *= delta, -1 rho /= delta, 0 none / other
  int spec_pred;       // 1: the pass in hand speculated on (spec_rho, spec_scale) instead of "unchanged"
  T spec_rho, spec_scale;
  unsigned pred_hits;  // committed speculations whose rho had moved
  // indirect projector: which form of y = A x the coming projection uses (graph_solver.cuh: cgls_epilogue)
  int y_refresh;       // 1: the product
  int y_rec;           // 1: t_y - r from the CGLS residual recurrence
  unsigned final_iter, exact_count;
  T nrm_r, nrm_s, eps_pri, eps_dua, gap, eps_gap;
  // norm-estimate scratch (setup)
  T est, est_last;
  int est_done;
  unsigned est_iters;
};

// Early-exit gate evaluated by every kernel that sits inside the captured loop.
struct Gate {
  const int* stop;        // skip the launch when *stop != 0   (may be null)
  const int* need;        // skip the launch when *need == 0   (may be null)
  const unsigned* kpar;   // skip the launch when (*kpar & 1) != par: kernels captured for one iteration parity
  unsigned par;           //   (kpar may be null)
};
__device__ __forceinline__ bool gate_closed(const Gate& g) {
  if (g.stop != nullptr && *reinterpret_cast<const volatile int*>(g.stop) != 0) return true;
  if (g.need != nullptr && *reinterpret_cast<const volatile int*>(g.need) == 0) return true;
  if (g.kpar != nullptr && (*reinterpret_cast<const volatile unsigned*>(g.kpar) & 1u) != g.par) return true;
  return false;
}

// Block-level fold of per-thread reduction terms; thread t<NRED writes term t.
template <int NRED>
__device__ __forceinline__ void block_fold(double (&red)[NRED], double* out) {
  __shared__ double s_red[kWarps][kMaxRed];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    const double v = warp_sum(red[k]);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += s_red[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// ---- epilogues ----------------------------------------------------------------
// An epilogue consumes one finished product entry (index i, value val) and may
// add to NRED reduction terms.

// res = alpha*val + beta*add[i]; optional store; optional ||res||^2.
template <typename T>
struct EpiAffine {
  static constexpr int NRED = 1;
  T alpha, beta;
  const T* add;   // may be null
  T* out;         // may be null
  T* out2 = nullptr;   // optional second copy of the result
  __device__ __forceinline__ void operator()(size_t i, T val, double* red) const {
    T res = alpha * val;
    if (add != nullptr) res += beta * add[i];
    if (out != nullptr) out[i] = res;
    if (out2 != nullptr) out2[i] = res;
    red[0] += static_cast<double>(res) * static_cast<double>(res);
  }
};

// Sinkhorn-Knopp sweep: out = num / (val + cst)   (equil_helper.h:149-163)
template <typename T>
struct EpiSinkhorn {
  static constexpr int NRED = 1;
  T num, cst;
  T* out;
  __device__ __forceinline__ void operator()(size_t i, T val, double* red) const {
    out[i] = num / (val + cst);
  }
};

// Frobenius norm of D*A*E from one squared product: sum_i s_i * val_i with
// s = d^2 (matrix_dense.cpp:174-182 without materialising the scaled matrix).
template <typename T>
struct EpiWeightedSum {
  static constexpr int NRED = 1;
  const T* scale;
  __device__ __forceinline__ void operator()(size_t i, T val, double* red) const {
    const double s = static_cast<double>(scale[i]);
    red[0] += s * s * static_cast<double>(val);
  }
};

// Second ADMM half-step for one part (x or y) of z, fused behind the product
// that yields the projected value (pogs.cpp:296, 342-348, 397-399):
//   znew = alpha*val + add[i];  zt_next = t - znew;
//   red0 += (zprev - znew)^2 ;  red1 += (z12 - znew)^2
template <typename T>
struct EpiState {
  static constexpr int NRED = 2;
  T alpha;
  const T* add;     // may be null
  const T* zprev;   // previous projection (this part)
  const T* z12;     // prox output
  const T* t;       // over-relaxed point (ztemp)
  T* znew;          // new projection
  T* zt_next;       // dual variable after the update (unscaled)
  T* aux;           // optional copy of the raw product (wide case), may be null
  __device__ __forceinline__ void operator()(size_t i, T val, double* red) const {
    T zn = alpha * val;
    if (add != nullptr) zn += add[i];
    const T zp = zprev[i], zh = z12[i], ti = t[i];
    znew[i] = zn;
    zt_next[i] = ti - zn;
    if (aux != nullptr) aux[i] = val;
    const double ds = static_cast<double>(zp) - static_cast<double>(zn);
    const double dr = static_cast<double>(zh) - static_cast<double>(zn);
    red[0] += ds * ds;
    red[1] += dr * dr;
  }
};

// ---- first half-step -----------------------------------------------------------------
// Descriptor arrays of one separable function, already rescaled by the
// equilibration (pogs.cpp:608-617).
template <typename T>
struct Desc {
  const int* h;
  const T *a, *b, *c, *d, *e;
};

template <typename T>
struct ProxArgs {
  size_t n, m;               // |x|, |y| (local)
  Desc<T> g, f;              // g on x, f on y
  const T *x, *y;            // previous projection z^k
  const T *xt, *yt;          // stored dual z~ (to be multiplied by ctrl->zt_scale)
  T *x12, *y12;              // prox outputs z^{k+1/2}
  T *tx, *ty;                // over-relaxed point handed to the projection
  T *qx, *qy;                // z12 + z~ - z^k  (exact dual residual input / lambda, mu)
  T alpha;
};

// v = z - z~ ; z12 = prox(v) ; w = v - z12 ; t = z~ + alpha z12 + (1-alpha) z
// Blocks [0, gx) own the x part, blocks [gx, gridDim.x) the y part, so that the
// x-side sums do not depend on the (rank-local) length of y.  Per block three
// reductions: <w,z12>, |w|^2, |z12|^2                        (pogs.cpp:254-278)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_prox(ProxArgs<T> p, unsigned gx, const Ctrl<T>* __restrict__ ctrl, double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  const T rho = ctrl->rho, sc = ctrl->zt_scale;
  const bool isx = blockIdx.x < gx;
  const size_t len = isx ? p.n : p.m;
  const unsigned b0 = isx ? blockIdx.x : blockIdx.x - gx;
  const unsigned nb = isx ? gx : gridDim.x - gx;
  const Desc<T> D = isx ? p.g : p.f;
  const T* __restrict__ zin = isx ? p.x : p.y;
  const T* __restrict__ ztin = isx ? p.xt : p.yt;
  T* __restrict__ z12 = isx ? p.x12 : p.y12;
  T* __restrict__ tt = isx ? p.tx : p.ty;
  T* __restrict__ qq = isx ? p.qx : p.qy;
  double red[3] = {0, 0, 0};
  for (size_t j = static_cast<size_t>(b0) * kThreads + threadIdx.x; j < len; j += static_cast<size_t>(nb) * kThreads) {
    const T zk = zin[j];
    const T zt = sc * ztin[j];
    const T v = zk - zt;
    const T zh = prox_eval<T>(D.h[j], D.a[j], D.b[j], D.c[j], D.d[j], D.e[j], v, rho);
    const T w = v - zh;
    T t = zt + p.alpha * zh;
    t += (T(1) - p.alpha) * zk;
    z12[j] = zh;
    tt[j] = t;
    qq[j] = (zh + zt) - zk;
    const double wd = w, zd = zh;
    red[0] += wd * zd;
    red[1] += wd * wd;
    red[2] += zd * zd;
  }
  block_fold<3>(red, partials + static_cast<size_t>(blockIdx.x) * 3);
}

// ---- controller -----------------------------------------------------------------------
// Deterministic fold of `nb` per-block partial rows of width `stride` (column k).
__device__ __forceinline__ double fold_partials(const double* p, unsigned nb, int stride, int k) {
  __shared__ double s_part[kThreads];
  double s = 0;
  for (unsigned b = threadIdx.x; b < nb; b += kThreads) s += __ldcg(p + static_cast<size_t>(b) * stride + k);
  s_part[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
    __syncthreads();
  }
  const double r = s_part[0];
  __syncthreads();
  return r;
}

// Fold K columns of an [nb][stride] partials array in one pass (fixed order).
template <int K>
__device__ __forceinline__ void fold_partials_multi(const double* p, unsigned nb, int stride, double* out) {
  __shared__ double s_multi[kWarps][kMaxRed];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0;
  for (unsigned b = threadIdx.x; b < nb; b += kThreads) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += __ldcg(p + static_cast<size_t>(b) * stride + k);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_multi[warp][k] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += s_multi[w][k];
    out[k] = t;
  }
  __syncthreads();
}

struct CtrlIn {
  const double* prox_part;  unsigned prox_gx, prox_gy;   // [gx + gy][3]: x blocks then y blocks
  const double* spec_part;  unsigned spec_gx, spec_gy;   // same terms from the fused pass of the previous iteration
  const double* xs_part;    unsigned xs_nb;     // [nb][2]  x half-step
  const double* ys_part;    unsigned ys_nb;     // [nb][2]  y half-step
  const double* er_part;    unsigned er_nb;     // [nb][1]  exact primal residual
  const double* es_part;    unsigned es_nb;     // [nb][1]  exact dual residual
  volatile unsigned* host_progress;             // mapped host memory: {iterations done, done flag}
  PeerView pv;                                  // row-block multi-GPU: y-side sums are per rank
};

constexpr unsigned kYRefresh = 16;   // indirect projector: iterations between two products y = A x

// End of an iteration: stopping rule and adaptive rho (pogs.cpp:379-469).
template <typename T>
__device__ void finish_iteration(Ctrl<T>* c, bool exact, volatile unsigned* host_progress, bool tail_follows = false) {
  const T nrm_r = c->nrm_r, nrm_s = c->nrm_s, eps_pri = c->eps_pri, eps_dua = c->eps_dua;
  const unsigned k = c->k;
  const bool converged = exact && nrm_r < eps_pri && nrm_s < eps_dua && (!c->gap_stop || c->gap < c->eps_gap);
  if (exact) c->exact_count += 1;
  if (converged || k == c->max_iter - 1) {
    c->final_iter = k;
    c->converged = converged ? 1 : 0;
    c->done = 1;
  } else {
    T scale = 1;
    const T rho_before = c->rho;
    int act = 0;
    if (c->adaptive_rho) {
      const T kDeltaMin = T(1.05), kGamma = T(1.01), kTau = T(0.8), kRhoMin = T(1e-4), kRhoMax = T(1e4),
              kKappa = T(0.9);
      T rho = c->rho, delta = c->delta, xi = c->xi;
      if (k > 0 && k % 50u == 0 && eps_pri > 0 && eps_dua > 0) {
        const T pn = nrm_r / eps_pri, dn = nrm_s / eps_dua;
        if (pn > 0 && dn > 0) {
          const T imb = pn / dn;
          if (imb > T(10) || imb < T(1) / T(10)) {
            T ratio = m_sqrt(imb);
            ratio = m_max(T(0.67), m_min(T(1.5), ratio));
            T rho_new = rho * ratio;
            rho_new = m_max(kRhoMin, m_min(kRhoMax, rho_new));
            if (m_abs(rho_new - rho) / rho > T(0.05)) {
              scale = rho / rho_new;
              rho = rho_new;
            }
          }
        }
      } else if (nrm_s < xi * eps_dua && nrm_r > xi * eps_pri && kTau * static_cast<T>(k) > static_cast<T>(c->kd)) {
        if (rho < kRhoMax) { rho *= delta; scale = 1 / delta; delta = kGamma * delta; c->ku = k; act = 1; }
      } else if (nrm_s > xi * eps_dua && nrm_r < xi * eps_pri && kTau * static_cast<T>(k) > static_cast<T>(c->ku)) {
        if (rho > kRhoMin) { rho /= delta; scale = delta; delta = kGamma * delta; c->kd = k; act = -1; }
      } else if (nrm_s < xi * eps_dua && nrm_r < xi * eps_pri) {
        xi *= kKappa;
      } else {
        delta = kDeltaMin;
      }
      c->rho = rho; c->delta = delta; c->xi = xi;
    }
    c->zt_scale = scale;
    c->prev_nrm_r = nrm_r;
    c->k = k + 1;
    c->y_refresh = ((k + 1) % kYRefresh == 0) ? 1 : 0;
    c->y_rec = 1 - c->y_refresh;
    // the speculative half-step of the next iteration assumed rho and the z~ scale unchanged -- or, in the
    // one-launch kernel, the outcome of the predicted action (same operands, same operations: same bits)
    const T srho = c->spec_pred ? c->spec_rho : rho_before;
    const T ssc = c->spec_pred ? c->spec_scale : T(1);
    const int miss = (c->fused_enabled && c->rho == srho && scale == ssc) ? 0 : 1;
    c->spec_pred = 0;
    c->pred_act = act;
    c->spec_miss = miss;
    c->need_solve = (miss || !tail_follows) ? 1 : 0;
    if (!miss) { c->spec_hits += 1; if (scale != T(1)) c->pred_hits += 1; }
  }
  if (host_progress != nullptr) {
    host_progress[0] = k + 1;
    host_progress[1] = static_cast<unsigned>(c->done);
    __threadfence_system();
  }
}

// Optional device-side switch of a CUDA-graph IF node that holds the exact-residual
// branch (two extra products + phase 1): the controller arms it only when needed, so a
// normal iteration launches nothing for that branch.
struct CondSwitch {
  cudaGraphConditionalHandle handle;
  int enabled;
};
__device__ __forceinline__ void cond_set(const CondSwitch& cs, bool on) {
  if (cs.enabled) cudaGraphSetConditional(cs.handle, on ? 1u : 0u);
}

// phase 0: after the projection -- tolerances, approximate residuals, decide
//          whether the exact residuals are needed (pogs.cpp:268-273, 342-352).
// Called by all threads of one CTA.
template <typename T>
__device__ __forceinline__ void control_phase0(Ctrl<T>* c, const CtrlIn& in, const CondSwitch& cs) {
  double xs[5], ys[5];
  // first half-step sums: from k_prox, or from the committed speculation of the previous pass
  const bool spec = in.spec_part != nullptr && c->spec_miss == 0;
  const double* pp = spec ? in.spec_part : in.prox_part;
  const unsigned pgx = spec ? in.spec_gx : in.prox_gx, pgy = spec ? in.spec_gy : in.prox_gy;
  fold_partials_multi<3>(pp, pgx, 3, xs);
  fold_partials_multi<2>(in.xs_part, in.xs_nb, 2, xs + 3);
  fold_partials_multi<3>(pp + static_cast<size_t>(pgx) * 3, pgy, 3, ys);
  fold_partials_multi<2>(in.ys_part, in.ys_nb, 2, ys + 3);
  peer_sum_scalars<5>(in.pv, ys);   // y lives row-sharded across the ranks
  const double dxs = xs[3], dxr = xs[4], dys = ys[3], dyr = ys[4];
  if (threadIdx.x == 0) {
    const T rho = c->rho;
    c->gap = m_abs(static_cast<T>(xs[0] + ys[0]));
    c->eps_gap = c->sqrtmn_atol +
                 c->rel_tol * static_cast<T>(sqrt(xs[1] + ys[1])) * static_cast<T>(sqrt(xs[2] + ys[2]));
    c->eps_pri = c->sqrtm_atol + c->rel_tol * static_cast<T>(sqrt(ys[2]));
    c->eps_dua = rho * (c->sqrtn_atol + c->rel_tol * static_cast<T>(sqrt(xs[1])));
    c->nrm_s = rho * (c->nrmA * static_cast<T>(sqrt(dys)) + static_cast<T>(sqrt(dxs)));
    c->nrm_r = c->nrmA * static_cast<T>(sqrt(dxr)) + static_cast<T>(sqrt(dyr));
    const bool need = c->nrm_r < T(10) * c->eps_pri && c->nrm_s < T(10) * c->eps_dua;
    c->need_exact = need ? 1 : 0;
    cond_set(cs, need);
    if (!need) finish_iteration(c, false, in.host_progress);
  }
}

// phase 1: after the two extra products -- exact residuals (pogs.cpp:353-376).
template <typename T>
__global__ void __launch_bounds__(kThreads) k_control(Ctrl<T>* c, CtrlIn in, int phase, CondSwitch cs, int parity = -1) {
  if (c->done) return;
  if (parity >= 0 && static_cast<int>(c->k & 1u) != parity) return;   // captured for the other iteration parity
  if (phase == 0) {
    control_phase0(c, in, cs);
  } else {
    if (!c->need_exact) return;
    double erv[1] = {fold_partials(in.er_part, in.er_nb, 1, 0)};
    peer_sum_scalars<1>(in.pv, erv);
    const double er = erv[0];
    const double es = fold_partials(in.es_part, in.es_nb, 1, 0);
    if (threadIdx.x == 0) {
      c->nrm_r = static_cast<T>(sqrt(er));
      c->nrm_s = c->rho * static_cast<T>(sqrt(es));
      c->need_exact = 0;
      finish_iteration(c, true, in.host_progress);
    }
  }
}

// Controller fused behind the last product of the projection: the last CTA of
// k_rowdot to finish (ticket) runs phase 0, saving a launch per iteration.
template <typename T>
struct TailCtrl {
  Ctrl<T>* c;          // null: no tail
  CtrlIn in;
  unsigned* ticket;
  CondSwitch cs;
};

// ---- row-dot product ------------------------------------------------------------
// out[r] = sum_c f(M[r][c]) * v[c], f = identity or square.  One warp owns a row
// at a time and walks it with 16 B loads, UNROLL independent loads in flight per
// lane; rows are dealt round-robin to the resident warps of the whole grid.
// Requires ld % VEC == 0 and 16 B-aligned M, v (v zero-padded to ld).
template <typename T, bool SQ, int UNROLL, typename Epi, bool CTAROW = false>
__global__ void __launch_bounds__(kThreads, 4)
k_rowdot(const T* __restrict__ M, size_t R, size_t C, size_t ld, const T* __restrict__ v, Epi epi,
         double* __restrict__ partials, Gate gate, TailCtrl<T> tail) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  if (gate_closed(gate)) return;
  const int lane = threadIdx.x & 31;
  const size_t nvec = (C + VEC - 1) / VEC;   // vectors per row (pad inside ld is zero)
  const VT* __restrict__ vv = reinterpret_cast<const VT*>(v);
  double red[Epi::NRED];
#pragma unroll
  for (int k = 0; k < Epi::NRED; ++k) red[k] = 0.0;

  if (CTAROW) {
    // Few rows per warp (e.g. the n x n factor): a whole CTA walks one row, so that the
    // work divides evenly over the resident CTAs instead of leaving a one-row tail.
    __shared__ double s_rowpart[kWarps];
    for (size_t r = blockIdx.x; r < R; r += gridDim.x) {
      const VT* __restrict__ row = reinterpret_cast<const VT*>(M + r * ld);
      T acc0 = 0, acc1 = 0;
      size_t j = threadIdx.x;
      for (; j + static_cast<size_t>(kThreads) * (UNROLL - 1) < nvec; j += static_cast<size_t>(kThreads) * UNROLL) {
        VT a[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) a[u] = ld_stream(row + j + static_cast<size_t>(kThreads) * u);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const VT x = __ldg(vv + j + static_cast<size_t>(kThreads) * u);
          if (u & 1) acc1 += dotv<SQ>(a[u], x); else acc0 += dotv<SQ>(a[u], x);
        }
      }
      {
        // remainder: up to UNROLL-1 more vectors per thread, still issued together
        VT a[UNROLL];
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < UNROLL - 1; ++u)
          if (j + static_cast<size_t>(kThreads) * u < nvec) { a[u] = ld_stream(row + j + static_cast<size_t>(kThreads) * u); cnt = u + 1; }
#pragma unroll
        for (int u = 0; u < UNROLL - 1; ++u)
          if (u < cnt) acc0 += dotv<SQ>(a[u], __ldg(vv + j + static_cast<size_t>(kThreads) * u));
      }
      const double ws = warp_sum(static_cast<double>(acc0 + acc1));
      if (lane == 0) s_rowpart[threadIdx.x >> 5] = ws;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_rowpart[w];
        epi(r, static_cast<T>(t), red);
      }
      __syncthreads();
    }
  } else {
    const size_t gwarp = static_cast<size_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    const size_t nwarps = static_cast<size_t>(gridDim.x) * kWarps;
    for (size_t r = gwarp; r < R; r += nwarps) {
      const VT* __restrict__ row = reinterpret_cast<const VT*>(M + r * ld);
      T acc0 = 0, acc1 = 0;
      size_t j = lane;
      for (; j + 32 * (UNROLL - 1) < nvec; j += 32 * UNROLL) {
        VT a[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) a[u] = ld_stream(row + j + 32 * u);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const VT x = __ldg(vv + j + 32 * u);
          if (u & 1) acc1 += dotv<SQ>(a[u], x); else acc0 += dotv<SQ>(a[u], x);
        }
      }
      for (; j < nvec; j += 32) {
        const VT a = ld_stream(row + j);
        const VT x = __ldg(vv + j);
        acc0 += dotv<SQ>(a, x);
      }
      const T sum = warp_sum(acc0 + acc1);
      if (lane == 0) epi(r, sum, red);
    }
  }
  if (partials != nullptr) block_fold<Epi::NRED>(red, partials + static_cast<size_t>(blockIdx.x) * Epi::NRED);
  if (tail.c != nullptr) {
    __shared__ int s_tail_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_tail_last = (atomicAdd(tail.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_tail_last) {
      __threadfence();
      control_phase0(tail.c, tail.in, tail.cs);
      if (threadIdx.x == 0) *tail.ticket = 0u;
    }
  }
}

// ---- symmetric factor apply: x = M u from the lower triangle only ---------------------------------
// M = (I + A^T A)^-1 is symmetric, so every off-diagonal entry M_ij (j < i) serves two outputs:
// x_i += M_ij u_j and x_j += M_ij u_i.  Reading only the lower triangle halves the bytes of the
// factor apply (0.4 GB -> 0.2 GB for n = 10000; it is 10 % of a BASELINE iteration).
// Strips of kSymStrip rows x (kThreads * VEC) columns that touch the triangle are dealt
// round-robin to a persistent grid; a thread owns VEC adjacent columns of the strip (16 B loads,
// coalesced per row) and keeps their column sums in registers for the whole strip; the strip is
// walked in sub-blocks of kSymRows rows whose per-row partial dots are reduced over the CTA.
// A strip leaves
//   rowpart[column block][i]   (row dots over its columns) and
//   colpart[strip][j]          (column sums over its rows)
// in fixed slots; k_symv_fold adds them up in index order (deterministic) and runs the epilogue.
constexpr int kSymRows = 16;     // rows per sub-block (per-row partials held in registers)
constexpr int kSymStrip = 64;    // rows per strip

struct SymTile { int rb, cb; };  // strip index, column block index

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
k_symv_tiles(const T* __restrict__ M, size_t n, size_t ld, const T* __restrict__ u, const SymTile* __restrict__ tiles,
             unsigned ntiles, T* __restrict__ rowpart, T* __restrict__ colpart, Gate gate) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  constexpr int kTileCols = kThreads * VEC;
  if (gate_closed(gate)) return;
  __shared__ T s_rp[kWarps][kSymRows];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int sb = tiles[t].rb, cb = tiles[t].cb;
    const size_t s0 = static_cast<size_t>(sb) * kSymStrip;
    const size_t c0 = static_cast<size_t>(cb) * kTileCols + static_cast<size_t>(threadIdx.x) * VEC;
    const bool active = c0 < ld;
    const bool strict = static_cast<size_t>(cb + 1) * kTileCols <= s0;   // every column of the strip is left of every row
    VT uc = zerov(static_cast<VT*>(nullptr));
    if (active) uc = __ldg(reinterpret_cast<const VT*>(u + c0));
    VT acc = zerov(static_cast<VT*>(nullptr));
    for (int blk = 0; blk < kSymStrip / kSymRows; ++blk) {
      const size_t i0 = s0 + static_cast<size_t>(blk) * kSymRows;
      if (i0 >= n) break;   // uniform over the CTA
      T rp[kSymRows];
#pragma unroll
      for (int r0 = 0; r0 < kSymRows; r0 += 8) {
        VT a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const size_t i = i0 + r0 + q;
          a[q] = (active && i < n) ? ld_stream(reinterpret_cast<const VT*>(M + i * ld + c0)) : zerov(static_cast<VT*>(nullptr));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const size_t i = i0 + r0 + q;
          VT m = a[q];
          if (!strict) {
            // keep j <= i for the row part
            T* me = reinterpret_cast<T*>(&m);
#pragma unroll
            for (int e = 0; e < VEC; ++e)
              if (c0 + e > i) me[e] = T(0);
          }
          rp[r0 + q] = dotv<false>(m, uc);
          if (!strict) {
            // the diagonal entry must not feed the column part
            T* me = reinterpret_cast<T*>(&m);
#pragma unroll
            for (int e = 0; e < VEC; ++e)
              if (c0 + e == i) me[e] = T(0);
          }
          const T ui = i < n ? __ldg(u + i) : T(0);
          fmav<false>(acc, m, ui);
        }
      }
      // row dots over this strip's columns: reduce over the CTA
#pragma unroll
      for (int r = 0; r < kSymRows; ++r) {
        const T v = warp_sum(rp[r]);
        if (lane == 0) s_rp[warp][r] = v;
      }
      __syncthreads();
      if (threadIdx.x < kSymRows) {
        T tsum = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tsum += s_rp[w][threadIdx.x];
        const size_t i = i0 + threadIdx.x;
        if (i < n) rowpart[static_cast<size_t>(cb) * ld + i] = tsum;
      }
      __syncthreads();
    }
    // column sums over this strip's rows
    if (active) *reinterpret_cast<VT*>(colpart + static_cast<size_t>(sb) * ld + c0) = acc;
  }
}

// x_i = sum over column blocks <= the diagonal's of rowpart[cb][i]
//     + sum over the strips whose tiles reach column i of colpart[sb][i];   then the epilogue.
// A CTA finishes 32 consecutive outputs: its 8 warps split the strips (coalesced 128 B loads,
// four independent partial sums each), shared-memory fold in warp order, warp 0 runs the epilogue.
template <typename T, typename Epi>
__global__ void __launch_bounds__(kThreads)
k_symv_fold(size_t n, size_t ld, unsigned nsb, const T* __restrict__ rowpart, const T* __restrict__ colpart, Epi epi,
            double* __restrict__ partials, Gate gate) {
  constexpr int kTileCols = kThreads * V16<T>::N;
  if (gate_closed(gate)) return;
  __shared__ T s_part[kWarps][32];
  double red[Epi::NRED];
#pragma unroll
  for (int k = 0; k < Epi::NRED; ++k) red[k] = 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t i = static_cast<size_t>(blockIdx.x) * 32 + lane;
  const unsigned cbd = static_cast<unsigned>((static_cast<size_t>(blockIdx.x) * 32) / kTileCols);   // same for all 32 outputs
  const unsigned sb_lo = cbd * (kTileCols / kSymStrip);   // strips (sb, cbd) exist for sb >= sb_lo
  T c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  if (i < n) {
    unsigned sb = sb_lo + warp;
    for (; sb + 3 * kWarps < nsb; sb += 4 * kWarps) {
      c0 += __ldcg(colpart + static_cast<size_t>(sb) * ld + i);
      c1 += __ldcg(colpart + static_cast<size_t>(sb + kWarps) * ld + i);
      c2 += __ldcg(colpart + static_cast<size_t>(sb + 2 * kWarps) * ld + i);
      c3 += __ldcg(colpart + static_cast<size_t>(sb + 3 * kWarps) * ld + i);
    }
    for (; sb < nsb; sb += kWarps) c0 += __ldcg(colpart + static_cast<size_t>(sb) * ld + i);
    if (warp == 0) {
      for (unsigned cb = 0; cb <= cbd; ++cb) c1 += __ldcg(rowpart + static_cast<size_t>(cb) * ld + i);
    }
  }
  s_part[warp][lane] = (c0 + c1) + (c2 + c3);
  __syncthreads();
  if (warp == 0 && i < n) {
    T s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += s_part[w][lane];   // fixed order
    epi(i, s, red);
  }
  if (partials != nullptr) block_fold<Epi::NRED>(red, partials + static_cast<size_t>(blockIdx.x) * Epi::NRED);
}

// ---- row-sharded factor apply with fused all-gather (row-block multi-GPU) -------------------------
// Every rank holds the whole n x n inverse but applies only its slice of rows
// [row0,row1): x_i = M[i,:] . u.  The slices are written to a peer-visible slot; the last
// CTA to finish (ticket) publishes the slice to every rank (itself included).  EVERY CTA then
// waits for the world's flags -- they live in local memory -- and runs the x half-step epilogue
// for its share of the full vector, reading the slices in rank order of ownership, so every
// rank ends up with a bit-identical x without a separate collective or a second launch.
// (The first version left the whole epilogue to the last CTA: 10000 elements behind NVLink loads
// on 256 threads cost ~44 us per iteration at every world size.)  The grid is one co-resident
// wave, so the in-kernel wait cannot starve a CTA that still has to run.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(kThreads, 4)
k_solve_shard(const T* __restrict__ M, size_t row0, size_t row1, size_t n, size_t ld, size_t slice,
              const T* __restrict__ v, EpiState<T> epi, double* __restrict__ partials,
              unsigned* __restrict__ ticket, Gate gate, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  if (gate_closed(gate)) return;
  __shared__ int s_last;
  const int lane = threadIdx.x & 31;
  const size_t gwarp = static_cast<size_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const size_t nwarps = static_cast<size_t>(gridDim.x) * kWarps;
  const size_t nvec = (n + VEC - 1) / VEC;
  const VT* __restrict__ vv = reinterpret_cast<const VT*>(v);
  const unsigned seq = *pv.seq(kGatherChannel) + 1u;   // read by every CTA before it takes its ticket
  T* mine = reinterpret_cast<T*>(pv.gath(pv.rank, seq));
  for (size_t r = row0 + gwarp; r < row1; r += nwarps) {
    const VT* __restrict__ row = reinterpret_cast<const VT*>(M + r * ld);
    T acc0 = 0, acc1 = 0;
    size_t j = lane;
    for (; j + 32 * (UNROLL - 1) < nvec; j += 32 * UNROLL) {
      VT a[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) a[u] = ld_stream(row + j + 32 * u);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const VT x = __ldg(vv + j + 32 * u);
        if (u & 1) acc1 += dotv<false>(a[u], x); else acc0 += dotv<false>(a[u], x);
      }
    }
    for (; j < nvec; j += 32) acc0 += dotv<false>(ld_stream(row + j), __ldg(vv + j));
    const T sum = warp_sum(acc0 + acc1);
    if (lane == 0) mine[r] = sum;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    // the whole slice of this rank is in place: tell every rank, ourselves included
    __threadfence_system();
    if (threadIdx.x < static_cast<unsigned>(pv.world)) st_sys(pv.flag(threadIdx.x, kGatherChannel, pv.rank), seq);
    if (threadIdx.x == 0) { *pv.seq(kGatherChannel) = seq; *ticket = 0u; }
  }
  // all slices present?  (flags are written by the owners into OUR memory)
  if (threadIdx.x < static_cast<unsigned>(pv.world)) {
    const unsigned* f = pv.flag(pv.rank, kGatherChannel, threadIdx.x);
    const long long t0 = clock64();
    while (static_cast<int>(ld_sys(f) - seq) < 0) {
      if (clock64() - t0 > kPeerTimeoutCycles) { *pv.err() = 1; break; }
    }
  }
  __syncthreads();
  __threadfence_system();
  // this CTA's share of the x half-step
  double red[2] = {0.0, 0.0};
  const size_t per = (nvec + gridDim.x - 1) / gridDim.x;
  const size_t v0 = static_cast<size_t>(blockIdx.x) * per, v1 = v0 + per < nvec ? v0 + per : nvec;
  for (size_t jv = v0 + threadIdx.x; jv < v1; jv += kThreads) {
    const int owner = static_cast<int>((jv * VEC) / slice);
    const VT got = ld_peer(reinterpret_cast<const VT*>(pv.gath(owner, seq)) + jv);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (jv * VEC + e < n) epi(jv * VEC + e, elemv(got, e), red);
  }
  block_fold<2>(red, partials + static_cast<size_t>(blockIdx.x) * 2);
}

// ---- column accumulation ----------------------------------------------------------
// out[c] = sum_r f(M[r][c]) * w[r].  grid = (column tiles, row chunks); a thread
// owns VEC adjacent columns and streams its chunk of rows with UNROLL loads in
// flight; the chunk result goes to part[chunk][c]; the last CTA to finish a
// column tile (ticket counter) folds the chunks in fixed order and runs the
// epilogue -- no second launch and no floating-point atomics.
template <typename T, bool SQ, int UNROLL, typename Epi>
__global__ void __launch_bounds__(kThreads, 4)
k_colacc(const T* __restrict__ M, size_t R, size_t C, size_t ld, const T* __restrict__ w,
         size_t rows_per_chunk, T* __restrict__ part, unsigned* __restrict__ tickets, Epi epi,
         double* __restrict__ partials, Gate gate, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  if (gate_closed(gate)) return;
  __shared__ int s_last;
  const size_t c0 = (static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x) * VEC;
  const bool active = c0 < ld;
  const size_t r0 = static_cast<size_t>(blockIdx.y) * rows_per_chunk;
  size_t r1 = r0 + rows_per_chunk;
  if (r1 > R) r1 = R;
  VT acc = zerov(static_cast<VT*>(nullptr));
  if (active && r0 < r1) {
    const T* __restrict__ base = M + c0;
    size_t r = r0;
    for (; r + UNROLL <= r1; r += UNROLL) {
      VT a[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) a[u] = ld_stream(reinterpret_cast<const VT*>(base + (r + u) * ld));
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) fmav<SQ>(acc, a[u], __ldg(w + r + u));
    }
    for (; r < r1; ++r) {
      const VT a = ld_stream(reinterpret_cast<const VT*>(base + r * ld));
      fmav<SQ>(acc, a, __ldg(w + r));
    }
  }
  if (active) *reinterpret_cast<VT*>(part + static_cast<size_t>(blockIdx.y) * ld + c0) = acc;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(tickets + blockIdx.x, 1u);
    s_last = (prev == gridDim.y - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double red[Epi::NRED];
#pragma unroll
  for (int k = 0; k < Epi::NRED; ++k) red[k] = 0.0;
  VT sum = zerov(static_cast<VT*>(nullptr));
  if (active) {
    const unsigned nch = gridDim.y;
    unsigned ch = 0;
    for (; ch + 4 <= nch; ch += 4) {
      VT p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = ld_cg(reinterpret_cast<const VT*>(part + static_cast<size_t>(ch + u) * ld + c0));
#pragma unroll
      for (int u = 0; u < 4; ++u) addv(sum, p[u]);
    }
    for (; ch < nch; ++ch) addv(sum, ld_cg(reinterpret_cast<const VT*>(part + static_cast<size_t>(ch) * ld + c0)));
  }
  if (pv.active()) {
    // Row-block multi-GPU: this tile of the local A_g^T w is one rank's share of
    // A^T w.  Exchange it over NVLink peer memory right here, in the tail of the
    // product, and sum the shares in rank order (see peer_comm.cuh).
    const unsigned seq = *pv.seq(blockIdx.x) + 1u;
    if (active) *reinterpret_cast<VT*>(pv.data(pv.rank, seq) + c0 * sizeof(T)) = sum;
    peer_signal_wait(pv, blockIdx.x, seq);
    if (active) {
      VT share[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)   // all loads in flight together: one NVLink round trip
        if (r < pv.world) share[r] = ld_peer(reinterpret_cast<const VT*>(pv.data(r, seq) + c0 * sizeof(T)));
      sum = zerov(static_cast<VT*>(nullptr));
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < pv.world) addv(sum, share[r]);
    }
    if (threadIdx.x == 0) *pv.seq(blockIdx.x) = seq;
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (c0 + e < C) epi(c0 + e, elemv(sum, e), red);
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;   // re-arm for the next launch
  if (partials != nullptr) block_fold<Epi::NRED>(red, partials + static_cast<size_t>(blockIdx.x) * Epi::NRED);
}

// ---- small elementwise kernels -------------------------------------------------------------
// Apply the equilibration to the descriptors (pogs.cpp:608-617) and clamp c,e>=0
// (FunctionObj::CheckConsts, prox_lib.h:62-69).  mode 0: divide by s (f with d),
// mode 1: multiply by s (g with e).
template <typename T>
__global__ void k_scale_desc(size_t n, const T* __restrict__ s, int mode, const T* a_in, const T* c_in,
                             const T* d_in, const T* e_in, T* a, T* c, T* d, T* e) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T si = s[i];
  const T cc = m_max(c_in[i], T(0)), ee = m_max(e_in[i], T(0));
  c[i] = cc;
  if (mode == 0) { a[i] = a_in[i] / si; d[i] = d_in[i] / si; e[i] = ee / (si * si); }
  else           { a[i] = a_in[i] * si; d[i] = d_in[i] * si; e[i] = ee * (si * si); }
}

// Objective sum_j g_j(x_j) [blocks < gx] and sum_i f_i(y_i) [other blocks]
// (pogs.cpp:473, prox_lib.h:521-529).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_objective(size_t n, size_t m, unsigned gx, Desc<T> g, Desc<T> f, const T* __restrict__ x,
            const T* __restrict__ y, double* __restrict__ partials) {
  double red[1] = {0};
  const bool isx = blockIdx.x < gx;
  const size_t len = isx ? n : m;
  const unsigned b0 = isx ? blockIdx.x : blockIdx.x - gx;
  const unsigned nb = isx ? gx : gridDim.x - gx;
  const Desc<T> D = isx ? g : f;
  const T* __restrict__ v = isx ? x : y;
  for (size_t j = static_cast<size_t>(b0) * kThreads + threadIdx.x; j < len; j += static_cast<size_t>(nb) * kThreads)
    red[0] += static_cast<double>(func_eval<T>(D.h[j], D.a[j], D.b[j], D.c[j], D.d[j], D.e[j], v[j]));
  block_fold<1>(red, partials + blockIdx.x);
}

// g(x) + sum over ranks of the local f(y).
static __global__ void __launch_bounds__(kThreads)
k_fold_objective(const double* part, unsigned gx, unsigned gy, PeerView pv, double* out) {
  const double gsum = fold_partials(part, gx, 1, 0);
  double fv[1] = {fold_partials(part + gx, gy, 1, 0)};
  peer_sum_scalars<1>(pv, fv);
  if (threadIdx.x == 0) *out = gsum + fv[0];
}

// Un-scale the outputs (pogs.cpp:510-518): x = x12*e, y = y12/d,
// mu = -rho*q_x/e, lambda = -rho*q_y*d.
template <typename T>
__global__ void k_outputs(size_t n, size_t m, const T* __restrict__ d, const T* __restrict__ e,
                          const T* __restrict__ x12, const T* __restrict__ y12, const T* __restrict__ qx,
                          const T* __restrict__ qy, const Ctrl<T>* __restrict__ ctrl, T* xo, T* yo, T* mu,
                          T* lambda) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const T nrho = -ctrl->rho;
  if (i < n) {
    xo[i] = x12[i] * e[i];
    mu[i] = nrho * qx[i] / e[i];
  } else if (i < n + m) {
    const size_t j = i - n;
    yo[j] = y12[j] / d[j];
    lambda[j] = nrho * qy[j] * d[j];
  }
}

// out = s * in   (s read from device memory when sp != null)
template <typename T>
__global__ void k_scale_copy(size_t n, const T* __restrict__ in, T s, const T* __restrict__ sp, T* out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (sp != nullptr ? *sp : s) * in[i];
}
template <typename T>
__global__ void k_div(size_t n, const T* __restrict__ a, const T* __restrict__ b, T* out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] / b[i];
}
template <typename T>
__global__ void k_fill(size_t n, T v, T* out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
template <typename T>
__global__ void k_sqrt_inplace(size_t n, T* v) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) v[i] = m_sqrt(v[i]);
}
template <typename T>
__global__ void k_square(size_t n, const T* __restrict__ in, T* out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * in[i];
}

// A := diag(d) * A * diag(e) * s  on the row-major storage (rows_are_d: storage
// rows carry d, columns carry e; otherwise the transposed roles).  Vectorised
// 16 B read-modify-write (matrix_dense.cpp:182-189, 227-246).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_scale_matrix(T* __restrict__ M, size_t R, size_t C, size_t ld, const T* __restrict__ rs,
               const T* __restrict__ cs, const T* __restrict__ s_ptr) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  const T s = *s_ptr;
  const size_t nvec_row = ld / VEC;
  const size_t total = R * nvec_row;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kThreads) {
    const size_t r = idx / nvec_row, cv = idx - r * nvec_row;
    const size_t c = cv * VEC;
    VT* p = reinterpret_cast<VT*>(M + r * ld + c);
    VT a = *p;
    const T rr = rs[r] * s;
    T* ae = reinterpret_cast<T*>(&a);
#pragma unroll
    for (int e = 0; e < VEC; ++e) ae[e] = (c + e < C) ? ae[e] * (rr * cs[c + e]) : T(0);
    *p = a;
  }
}

// Norm-estimate bookkeeping (equil_helper.h:119-131): est = |x|/|Sx|,
// x /= |x|, stop when the estimate stalls.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_normest_step(Ctrl<T>* c, const double* nx_part, unsigned nx_nb, const double* nsx_part, unsigned nsx_nb,
               T* inv_out, PeerView pv) {
  if (c->est_done) return;
  const double nx2 = fold_partials(nx_part, nx_nb, 1, 0);
  double sv[1] = {fold_partials(nsx_part, nsx_nb, 1, 0)};
  peer_sum_scalars<1>(pv, sv);
  const double nsx2 = sv[0];
  if (threadIdx.x == 0) {
    const T normx = static_cast<T>(sqrt(nx2)), normSx = static_cast<T>(sqrt(nsx2));
    const T last = c->est;
    const T est = normx / normSx;
    c->est_last = last;
    c->est = est;
    c->est_iters += 1;
    *inv_out = 1 / normx;
    if (m_abs(last - est) < T(1e-4) * est) c->est_done = 1;
  }
}

// Scalar glue after the Frobenius reduction: normA = sqrt(sum)/sqrt(min(m,n));
// outputs 1/normA and 1/sqrt(normA).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_fro_finish(const double* part, unsigned nb, double min_dim, T* inv_norm, T* inv_sqrt_norm, PeerView pv) {
  double fv[1] = {fold_partials(part, nb, 1, 0)};
  peer_sum_scalars<1>(pv, fv);
  const double s = fv[0];
  if (threadIdx.x == 0) {
    const T normA = static_cast<T>(sqrt(s) / sqrt(min_dim));
    *inv_norm = 1 / normA;
    *inv_sqrt_norm = 1 / m_sqrt(normA);
  }
}

// Fold objective partials to one double.
static __global__ void __launch_bounds__(kThreads) k_fold1(const double* part, unsigned nb, double* out) {
  const double s = fold_partials(part, nb, 1, 0);
  if (threadIdx.x == 0) *out = s;
}

// In-place one-shot all-reduce of buf[0..len) over the ranks: CTA b owns 16 B
// vectors [b*kThreads, (b+1)*kThreads) and exchanges them on tile channel b.
// len * sizeof(T) must fit one data slot and gridDim.x <= kMaxTileChannels.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_peer_allreduce(T* __restrict__ buf, size_t len, PeerView pv) {
  using VT = typename V16<T>::type;
  constexpr int VEC = V16<T>::N;
  if (!pv.active()) return;
  const size_t c0 = (static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x) * VEC;
  const bool active = c0 < len;   // len is a multiple of VEC (padded buffers)
  const unsigned seq = *pv.seq(blockIdx.x) + 1u;
  if (active) *reinterpret_cast<VT*>(pv.data(pv.rank, seq) + c0 * sizeof(T)) = *reinterpret_cast<const VT*>(buf + c0);
  peer_signal_wait(pv, blockIdx.x, seq);
  if (active) {
    VT share[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < pv.world) share[r] = ld_peer(reinterpret_cast<const VT*>(pv.data(r, seq) + c0 * sizeof(T)));
    VT sum = zerov(static_cast<VT*>(nullptr));
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < pv.world) addv(sum, share[r]);
    *reinterpret_cast<VT*>(buf + c0) = sum;
  }
  if (threadIdx.x == 0) *pv.seq(blockIdx.x) = seq;
}

// Symmetrise the lower/upper triangle returned by potri and convert from the working precision
// W of the factorisation to T with a padded leading dimension.
template <typename T, typename W>
__global__ void k_sym_cast(size_t k, const W* __restrict__ src, size_t lds, T* __restrict__ dst, size_t ldd,
                           int src_lower_rowmajor) {
  const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t i = blockIdx.y;
  if (i >= k || j >= ldd) return;
  T v = T(0);
  if (j < k) {
    size_t a = i, b = j;
    const bool in_lower = a >= b;
    if (in_lower != (src_lower_rowmajor != 0)) { const size_t t = a; a = b; b = t; }
    v = static_cast<T>(src[a * lds + b]);
  }
  dst[i * ldd + j] = v;
}
// ones on the diagonal of a zero-initialised k x k array
static __global__ void k_set_identity(size_t k, float* __restrict__ dst, size_t ld) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < k) dst[i * ld + i] = 1.0f;
}
// dst (working precision W) = src + diag * I
template <typename T, typename W>
__global__ void k_widen_add_diag(size_t k, const T* __restrict__ src, size_t lds, W* __restrict__ dst,
                                 size_t ldd, W diag) {
  const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t i = blockIdx.y;
  if (i >= k || j >= k) return;
  dst[i * ldd + j] = static_cast<W>(src[i * lds + j]) + (i == j ? diag : W(0));
}

}  // namespace pogs_b200
