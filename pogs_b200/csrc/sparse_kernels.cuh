// Kernels of the indirect (CGLS) projector and of the sparse operator.
//
//   k_spmv         : out[r] = sum_k f(val[k]) * v[ind[k]] over one compressed copy
//                    (CSR for A v, CSC-as-CSR for A^T w) -- the row-gather form of
//                    the reference (src/cpu/include/gsl/gsl_spblas.h:10-40), one
//                    sub-warp of G lanes per row, 4 independent gathers in flight
//                    per lane, same fused epilogues as the dense products.
//   k_cgls_*       : the vector updates and scalar recurrences of cgls::Solve
//                    (src/cpu/include/cgls.h:222-306) with the scalars (gamma,
//                    alpha, beta, norms) kept in double on the device, like the
//                    reference keeps them in double on the host.
#pragma once

#include "kernels.cuh"

namespace pogs_b200 {

// Scalar streaming loads of the compressed arrays: read-only path without L1 allocation, so that
// L1 is left to the gathered vector (ncu, C5: 100 M gathers = 4 GB of 32 B sector traffic against
// 0.8 GB of streamed matrix; the products are bound by that gather traffic, not by HBM).
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream1(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// ---- sparse row-gather product ---------------------------------------------------------------
template <typename T, bool SQ, typename Epi>
__global__ void __launch_bounds__(kThreads)
k_spmv(const T* __restrict__ val, const int* __restrict__ ind, const int* __restrict__ ptr, size_t rows,
       int lg_group, const T* __restrict__ v, Epi epi, double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  const int G = 1 << lg_group;
  const int lane_g = threadIdx.x & (G - 1);
  const size_t group = (static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x) >> lg_group;
  const size_t ngroups = (static_cast<size_t>(gridDim.x) * kThreads) >> lg_group;
  double red[Epi::NRED];
#pragma unroll
  for (int k = 0; k < Epi::NRED; ++k) red[k] = 0.0;
  // all lanes of a warp run the same number of trips (rows padded up per warp) so that
  // the shuffles below are always executed by the full warp
  const size_t rows_pad = (rows + ngroups - 1) / ngroups * ngroups;
  for (size_t r = group; r < rows_pad; r += ngroups) {
    T acc = 0;
    if (r < rows) {
      const int k0 = ptr[r], k1 = ptr[r + 1];
      int k = k0 + lane_g;
      for (; k + 3 * G < k1; k += 4 * G) {
        const int i0 = ld_stream1(ind + k), i1 = ld_stream1(ind + k + G), i2 = ld_stream1(ind + k + 2 * G), i3 = ld_stream1(ind + k + 3 * G);
        const T a0 = ld_stream1(val + k), a1 = ld_stream1(val + k + G), a2 = ld_stream1(val + k + 2 * G), a3 = ld_stream1(val + k + 3 * G);
        const T x0 = __ldg(v + i0), x1 = __ldg(v + i1), x2 = __ldg(v + i2), x3 = __ldg(v + i3);
        if (SQ) acc += a0 * a0 * x0 + a1 * a1 * x1 + a2 * a2 * x2 + a3 * a3 * x3;
        else    acc += a0 * x0 + a1 * x1 + a2 * x2 + a3 * x3;
      }
      for (; k < k1; k += G) {
        const T a = ld_stream1(val + k);
        const T x = __ldg(v + ld_stream1(ind + k));
        acc += SQ ? a * a * x : a * x;
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, G);
    if (lane_g == 0 && r < rows) epi(r, acc, red);
  }
  if (partials != nullptr) block_fold<Epi::NRED>(red, partials + static_cast<size_t>(blockIdx.x) * Epi::NRED);
}

// val[k] *= rs[row] * cs[ind[k]] * (*s)   (A := D A E / normA on one compressed copy,
// matrix_sparse.cpp:268-304)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_spscale(T* __restrict__ val, const int* __restrict__ ind, const int* __restrict__ ptr, size_t rows, int lg_group,
          const T* __restrict__ rs, const T* __restrict__ cs, const T* __restrict__ s_ptr) {
  const int G = 1 << lg_group;
  const int lane_g = threadIdx.x & (G - 1);
  const size_t group = (static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x) >> lg_group;
  const size_t ngroups = (static_cast<size_t>(gridDim.x) * kThreads) >> lg_group;
  const T s = *s_ptr;
  for (size_t r = group; r < rows; r += ngroups) {
    const T rr = rs[r] * s;
    for (int k = ptr[r] + lane_g; k < ptr[r + 1]; k += G) val[k] *= rr * cs[ind[k]];
  }
}

// out_i = epilogue(a_i - b_i): the epilogue of a product applied to a vector that is already known
// (y = A x = t_y - r from the CGLS residual recurrence); same number of partial blocks as the product.
template <typename T, typename Epi>
__global__ void __launch_bounds__(kThreads)
k_epi_diff(size_t n, const T* __restrict__ a, const T* __restrict__ b, Epi epi, double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  double red[Epi::NRED];
#pragma unroll
  for (int k = 0; k < Epi::NRED; ++k) red[k] = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * kThreads)
    epi(i, a[i] - b[i], red);
  if (partials != nullptr) block_fold<Epi::NRED>(red, partials + static_cast<size_t>(blockIdx.x) * Epi::NRED);
}

// ---- CGLS state ------------------------------------------------------------------------------------
// The inner loop runs either from the host in small batches (test hook, verbose tables) or,
// inside the captured ADMM iteration, as the body of a CUDA-graph WHILE node whose condition
// k_cgls_start arms and k_cgls_beta updates on the device (`loop` below) -- no host round trip.
struct CglsState {
  double gamma, norms0, norms, normx, xmax, pnorm2, tol, shift;
  int done, flag, indefinite;
  unsigned iters, maxit;
  unsigned long long total_iters;
};

// dx = x_warm - x0 ; red0 = |dx|^2 ; r = y0 - A x_warm   (projector_cgls.cpp:60-64 + cgls.h:228-236)
// The reference forms the start residual with two products, r = (y0 - A x0) - A dx.  Because
// x0 + dx = x_warm is the previous projection and the loop keeps y_warm = A x_warm (the
// projection epilogue recomputes y = A x, projector_cgls.cpp:78), the same vector is
// y0 - y_warm: no pass over A at all.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cgls_delta(size_t n, size_t m, const T* __restrict__ xw, const T* __restrict__ x0, T* __restrict__ dx,
             const T* __restrict__ y0, const T* __restrict__ yw, T* __restrict__ r,
             double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  double red[1] = {0};
  const size_t N = n + m;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < N;
       i += static_cast<size_t>(gridDim.x) * kThreads) {
    if (i < n) {
      const T v = xw[i] - x0[i];
      dx[i] = v;
      red[0] += static_cast<double>(v) * static_cast<double>(v);
    } else {
      const size_t j = i - n;
      r[j] = y0[j] - yw[j];
    }
  }
  block_fold<1>(red, partials + blockIdx.x);
}

// After s = A^T r - shift*dx has been formed (with |s|^2 partials): start-up scalars
// of cgls::Solve (cgls.h:240-250) and the projection tolerance of pogs.cpp:287-290.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cgls_start(CglsState* st, const Ctrl<T>* ctrl, double fixed_tol, const double* s_part, unsigned s_nb,
             const double* dx_part, unsigned dx_nb, unsigned maxit, Gate gate, CondSwitch loop) {
  if (gate_closed(gate)) return;   // the WHILE node keeps its default (0): no inner iterations
  const double s2 = fold_partials(s_part, s_nb, 1, 0);
  const double x2 = fold_partials(dx_part, dx_nb, 1, 0);
  if (threadIdx.x == 0) {
    double tol = fixed_tol;
    if (ctrl != nullptr) {
      const T kMin = T(1e-2), kMax = T(1e-8);
      T t = kMin * m_pow(m_min(ctrl->prev_nrm_r, T(1)), T(0.5));
      t = m_max(t, kMax);
      tol = static_cast<double>(t);
    }
    const double norms = static_cast<double>(static_cast<T>(sqrt(s2)));
    st->tol = tol; st->shift = 1.0;
    st->norms = norms; st->norms0 = norms; st->gamma = norms * norms; st->pnorm2 = s2;
    st->normx = static_cast<double>(static_cast<T>(sqrt(x2))); st->xmax = st->normx;
    st->iters = 0; st->maxit = maxit; st->indefinite = 0; st->flag = 0;
    const double eps = sizeof(T) == 4 ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    st->done = 0;
    if (norms < eps) { st->flag = 1; st->done = 1; }
    if (maxit == 0) st->done = 1;
    cond_set(loop, st->done == 0);   // graph mode: arm the WHILE node that holds the inner iteration
  }
}

// alpha = gamma / (|q|^2 + shift |p|^2);  dx += alpha p ; r -= alpha q ; red0 = |dx|^2
// (cgls.h:263-279).  Every block folds the |q|^2 partials itself (same order, same
// value) so no separate scalar launch is needed.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cgls_update1(size_t n, size_t m, const CglsState* __restrict__ st, const double* q_part, unsigned q_nb,
               const T* __restrict__ p, const T* __restrict__ q, T* __restrict__ dx, T* __restrict__ r,
               double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  const double q2 = fold_partials(q_part, q_nb, 1, 0);
  const double normq = static_cast<double>(static_cast<T>(sqrt(q2)));
  const double normp = static_cast<double>(static_cast<T>(sqrt(st->pnorm2)));
  double delta = normq * normq + st->shift * normp * normp;
  const double eps = sizeof(T) == 4 ? 1.1920928955078125e-07 : 2.220446049250313e-16;
  if (delta == 0.) delta = eps;
  const T alpha = static_cast<T>(st->gamma / delta);
  const T nalpha = static_cast<T>(-st->gamma / delta);
  double red[1] = {0};
  const size_t N = n + m;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < N;
       i += static_cast<size_t>(gridDim.x) * kThreads) {
    if (i < n) {
      const T v = dx[i] + alpha * p[i];
      dx[i] = v;
      red[0] += static_cast<double>(v) * static_cast<double>(v);
    } else {
      const size_t j = i - n;
      r[j] += nalpha * q[j];
    }
  }
  block_fold<1>(red, partials + blockIdx.x);
}

// beta, convergence test, iteration count (cgls.h:288-304).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cgls_beta(CglsState* st, const double* s_part, unsigned s_nb, const double* dx_part, unsigned dx_nb,
            const double* q_part, unsigned q_nb, Gate gate, CondSwitch loop) {
  if (gate_closed(gate)) return;
  const double s2 = fold_partials(s_part, s_nb, 1, 0);
  const double x2 = fold_partials(dx_part, dx_nb, 1, 0);
  const double q2 = fold_partials(q_part, q_nb, 1, 0);
  if (threadIdx.x == 0) {
    const double normq = static_cast<double>(static_cast<T>(sqrt(q2)));
    const double normp = static_cast<double>(static_cast<T>(sqrt(st->pnorm2)));
    if (normq * normq + st->shift * normp * normp <= 0.) st->indefinite = 1;
    const double norms = static_cast<double>(static_cast<T>(sqrt(s2)));
    const double gamma1 = st->gamma;
    st->norms = norms;
    st->gamma = norms * norms;
    // beta is consumed by k_cgls_update2 through gamma/gamma1: keep gamma1 in xmax's neighbour
    st->normx = static_cast<double>(static_cast<T>(sqrt(x2)));
    if (st->normx > st->xmax) st->xmax = st->normx;
    st->iters += 1;
    st->total_iters += 1;
    // stash beta in pnorm2's place holder until update2 recomputes |p|^2
    st->pnorm2 = st->gamma / gamma1;   // == beta (double); update2 overwrites with |p|^2 via partials
    const bool converged = (norms <= st->norms0 * st->tol) || (st->normx * st->tol >= 1.);
    if (converged || st->iters >= st->maxit) st->done = 1;
    cond_set(loop, st->done == 0);   // graph mode: another trip of the WHILE body or leave it
  }
}

// p = s + beta p ; red0 = |p|^2   (cgls.h:294-295).  Runs even when the test above
// fired (the reference also updates p before it leaves the loop); harmless.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cgls_update2(size_t n, const CglsState* __restrict__ st, const T* __restrict__ s, T* __restrict__ p,
               double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  const T beta = static_cast<T>(st->pnorm2);
  double red[1] = {0};
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * kThreads) {
    const T v = s[i] + beta * p[i];
    p[i] = v;
    red[0] += static_cast<double>(v) * static_cast<double>(v);
  }
  block_fold<1>(red, partials + blockIdx.x);
}

// |p|^2 from the partials of k_cgls_update2 into the state (single block).
static __global__ void __launch_bounds__(kThreads)
k_cgls_pnorm(CglsState* st, const double* p_part, unsigned p_nb, Gate gate) {
  if (gate_closed(gate)) return;
  const double p2 = fold_partials(p_part, p_nb, 1, 0);
  if (threadIdx.x == 0) st->pnorm2 = p2;
}

// x = x0 + dx with the x half of the second ADMM half-step (same arithmetic as
// EpiState): znew = x0 + dx ; zt_next = t - znew ; red0 = |zprev - znew|^2 ;
// red1 = |z12 - znew|^2.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_cgls_finish_x(size_t n, const T* __restrict__ x0, const T* __restrict__ dx, const T* __restrict__ zprev,
                const T* __restrict__ z12, const T* __restrict__ t, T* __restrict__ znew, T* __restrict__ zt_next,
                double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  double red[2] = {0, 0};
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * kThreads) {
    const T zn = x0[i] + dx[i];
    const T zp = zprev[i], zh = z12[i], ti = t[i];
    znew[i] = zn;
    zt_next[i] = ti - zn;
    const double ds = static_cast<double>(zp) - static_cast<double>(zn);
    const double dr = static_cast<double>(zh) - static_cast<double>(zn);
    red[0] += ds * ds;
    red[1] += dr * dr;
  }
  block_fold<2>(red, partials + static_cast<size_t>(blockIdx.x) * 2);
}

}  // namespace pogs_b200
