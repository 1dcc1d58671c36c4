// Setup algorithms shared by the dense and the sparse operator, written against
// the two products only (CRTP: Derived provides mul_n / mul_t / apply_scaling /
// nb_n / nb_t / nb_max): modified Sinkhorn-Knopp equilibration with Frobenius
// normalisation (reference src/cpu/matrix/matrix_dense.cpp:116-200,
// matrix_sparse.cpp:158-242, src/cpu/include/equil_helper.h:141-164) and the
// power-iteration estimate of ||A^||_2 (equil_helper.h:108-135).
#pragma once

#include <random>

#include "common.cuh"
#include "fused_pass.cuh"

namespace pogs_b200 {

template <typename Derived, typename T>
class MatAlgos {
 public:
  MatAlgos(size_t m, size_t n, cudaStream_t stream, size_t m_global = 0, const PeerView& pv = PeerView())
      : m_(m), n_(n), mg_(m_global ? m_global : m), stream_(stream), pv_(pv) {
    dev_ = query_device();
  }
  size_t rows() const { return m_; }
  size_t rows_global() const { return mg_; }
  const PeerView& peers() const { return pv_; }
  size_t cols() const { return n_; }
  const DeviceInfo& device() const { return dev_; }
  void set_stream(cudaStream_t s) { stream_ = s; }
  cudaStream_t stream() const { return stream_; }

  // Modified Sinkhorn-Knopp on A.^2 (squares formed in registers, so neither a
  // squared copy nor the reference's sign bit-vector exists), Frobenius
  // normalisation, in-place A := D A E / normA.  d (m) and e (n) are outputs.
  void equilibrate(T* d, T* e) {
    const size_t m = m_, n = n_;
    Derived& A = derived();
    const unsigned tb = 256;
    k_fill<T><<<(unsigned)((m + tb - 1) / tb), tb, 0, stream_>>>(m, T(1), d);
    k_fill<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, T(1), e);
    const size_t mg = mg_;   // the constants use the global row count
    const T ce = T(1e-4) * static_cast<T>(mg + n) / static_cast<T>(mg);
    const T cd = T(1e-4) * static_cast<T>(mg + n) / static_cast<T>(n);
    bool swept = false;
    if constexpr (Derived::kDense) {
      if (A.one_pass_ok() && !no_fused_setup()) {
        // One pass per sweep instead of two: while row i is on chip for d_i = n / (B_i . e + cd)
        // it is also added, times d_i, to the column sums that give the NEXT sweep's e.  One
        // plain column pass starts the chain (e_1 from d_0 = 1); the last pass's column sums are
        // not used.  Same arithmetic as the two-kernel sweep, 51 passes over A instead of 100.
        DevBuf<T> e_alt(n);
        A.template mul_t<true>(d, EpiSinkhorn<T>{static_cast<T>(mg), ce, e}, nullptr);
        T* e_cur = e;
        T* e_nxt = e_alt.get();
        for (int k = 0; k < 50; ++k) {
          A.template one_pass<true>(e_cur, SinkhornRowOp<T>{static_cast<T>(n), cd, d},
                                    SinkhornColOp<T>{static_cast<T>(mg), ce, e_nxt}, nullptr);
          if (k + 1 < 50) { T* t = e_cur; e_cur = e_nxt; e_nxt = t; }
        }
        if (e_cur != e) POGS_CUDA(cudaMemcpyAsync(e, e_cur, n * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
        POGS_CUDA(cudaStreamSynchronize(stream_));   // e_alt goes out of scope
        swept = true;
      }
    }
    if (!swept) {
      for (int k = 0; k < 50; ++k) {
        A.template mul_t<true>(d, EpiSinkhorn<T>{static_cast<T>(mg), ce, e}, nullptr);
        A.template mul_n<true>(e, EpiSinkhorn<T>{static_cast<T>(n), cd, d}, nullptr);
      }
    }
    k_sqrt_inplace<T><<<(unsigned)((m + tb - 1) / tb), tb, 0, stream_>>>(m, d);
    k_sqrt_inplace<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, e);
    // ||D A E||_F^2 = sum_i d_i^2 (A.^2 e.^2)_i
    DevBuf<T> e2(n), scal(2);
    DevBuf<double> fpart(A.nb_max());
    k_square<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, e, e2.get());
    A.template mul_n<true>(e2.get(), EpiWeightedSum<T>{d}, fpart.get());
    const double min_dim = static_cast<double>(mg < n ? mg : n);
    k_fro_finish<T><<<1, kThreads, 0, stream_>>>(fpart.get(), A.nb_n(), min_dim, scal.get(), scal.get() + 1, pv_);
    A.apply_scaling(d, e, scal.get());
    k_scale_copy<T><<<(unsigned)((m + tb - 1) / tb), tb, 0, stream_>>>(m, d, T(0), scal.get() + 1, d);
    k_scale_copy<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, e, T(0), scal.get() + 1, e);
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaStreamSynchronize(stream_));   // temporaries go out of scope
  }

  // Power iteration on A^T A from the reference's fixed start vector
  // (gsl_rand.h:9-16: default-seeded std::default_random_engine), <= 50 sweeps,
  // relative stall tolerance 1e-4; runs without host synchronisation.
  T norm2est(Ctrl<T>* ctrl) {
    const size_t m = m_, n = n_;
    Derived& A = derived();
    std::vector<T> x0(n);
    {
      std::default_random_engine gen;
      std::uniform_real_distribution<T> dist(static_cast<T>(0), static_cast<T>(1));
      for (size_t i = 0; i < n; ++i) x0[i] = dist(gen);
    }
    DevBuf<T> x(n), xn(n), Sx(m), inv(1);
    DevBuf<double> p_sx(A.nb_max()), p_x(A.nb_max());
    POGS_CUDA(cudaMemcpyAsync(x.get(), x0.data(), n * sizeof(T), cudaMemcpyHostToDevice, stream_));
    Ctrl<T> hc;
    POGS_CUDA(cudaMemcpyAsync(&hc, ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    hc.est = 0; hc.est_last = 0; hc.est_done = 0; hc.est_iters = 0;
    POGS_CUDA(cudaMemcpyAsync(ctrl, &hc, sizeof(hc), cudaMemcpyHostToDevice, stream_));
    Gate gate{&ctrl->est_done, nullptr};
    const unsigned tb = 256;
    if constexpr (Derived::kDense) {
      if (A.one_pass_ok() && !no_fused_setup()) {
        // One pass per sweep: Sx_i = A_i . x / |x| is row-local, so x' = A^T Sx accumulates while
        // the row is on chip.  x stays unnormalised; 1/|x| of the previous sweep (on the device)
        // is folded into the row map.
        const auto& pl = A.one_pass_plan();
        DevBuf<double> q_sx(pl.grid), q_x(pl.nfold);
        k_fill<T><<<1, 32, 0, stream_>>>(1, T(1), inv.get());
        T* x_cur = x.get();
        T* x_nxt = xn.get();
        for (int i = 0; i < 50; ++i) {
          A.template one_pass<false>(x_cur, PowerRowOp<T>{inv.get(), q_sx.get()}, PowerColOp<T>{x_nxt, q_x.get()},
                                     nullptr, gate);
          k_normest_step<T><<<1, kThreads, 0, stream_>>>(ctrl, q_x.get(), pl.nfold, q_sx.get(), pl.grid, inv.get(), pv_);
          T* t = x_cur; x_cur = x_nxt; x_nxt = t;
        }
        POGS_CUDA(cudaGetLastError());
        POGS_CUDA(cudaMemcpyAsync(&hc, ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream_));
        POGS_CUDA(cudaStreamSynchronize(stream_));
        normest_iters_ = hc.est_iters;
        return hc.est;
      }
    }
    for (int i = 0; i < 50; ++i) {
      A.template mul_n<false>(x.get(), EpiAffine<T>{T(1), T(0), nullptr, Sx.get()}, p_sx.get(), gate);
      A.template mul_t<false>(Sx.get(), EpiAffine<T>{T(1), T(0), nullptr, xn.get()}, p_x.get(), gate);
      k_normest_step<T><<<1, kThreads, 0, stream_>>>(ctrl, p_x.get(), A.nb_t(), p_sx.get(), A.nb_n(), inv.get(), pv_);
      k_scale_copy<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, xn.get(), T(0), inv.get(), x.get());
    }
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaMemcpyAsync(&hc, ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    normest_iters_ = hc.est_iters;
    return hc.est;
  }
  unsigned normest_iters() const { return normest_iters_; }

 protected:
  Derived& derived() { return static_cast<Derived&>(*this); }
  static bool no_fused_setup() {
    const char* e = getenv("POGS_B200_NO_FUSE_SETUP");
    return e != nullptr && e[0] == '1';
  }
  size_t m_, n_, mg_;
  cudaStream_t stream_;
  PeerView pv_;
  DeviceInfo dev_;
  unsigned normest_iters_ = 0;
};

}  // namespace pogs_b200
