// Setup algorithms shared by the dense and the sparse operator, written against
// the two products only (CRTP: Derived provides mul_n / mul_t / apply_scaling /
// nb_n / nb_t / nb_max): modified Sinkhorn-Knopp equilibration with Frobenius
// normalisation (reference src/cpu/matrix/matrix_dense.cpp:116-200,
// matrix_sparse.cpp:158-242, src/cpu/include/equil_helper.h:141-164) and the
// power-iteration estimate of ||A^||_2 (equil_helper.h:108-135).
#pragma once

#include <random>

#include "common.cuh"

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
    for (int k = 0; k < 50; ++k) {
      A.template mul_t<true>(d, EpiSinkhorn<T>{static_cast<T>(mg), ce, e}, nullptr);
      A.template mul_n<true>(e, EpiSinkhorn<T>{static_cast<T>(n), cd, d}, nullptr);
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
  size_t m_, n_, mg_;
  cudaStream_t stream_;
  PeerView pv_;
  DeviceInfo dev_;
  unsigned normest_iters_ = 0;
};

}  // namespace pogs_b200
