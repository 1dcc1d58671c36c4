// Device-resident dense operator: owns the (equilibrated) copy of A in HBM and
// exposes the two streaming products with fused epilogues.
//
// Replaces MatrixDense<T>::{Init,Mul,Equil} of the reference
// (/root/reference/src/cpu/matrix/matrix_dense.cpp:76-200) plus Norm2Est
// (src/cpu/include/equil_helper.h:108-135).  Layout: one row-major R x C array
// with leading dimension ld (multiple of the 16 B vector width).  A row-major
// input is stored as is (R=m, C=n); a column-major input is stored as the
// row-major array of A^T (R=n, C=m) and the two products swap kernels, so no
// transpose copy is ever made.
#pragma once

#include <random>

#include "common.cuh"

namespace pogs_b200 {

template <typename T, bool SQ, typename Epi>
inline void launch_rowdot(cudaStream_t s, const RowdotPlan& pl, const T* M, size_t R, size_t C, size_t ld,
                          const T* v, const Epi& epi, double* partials, Gate gate) {
  k_rowdot<T, SQ, 8, Epi><<<pl.grid, kThreads, 0, s>>>(M, R, C, ld, v, epi, partials, gate);
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

template <typename T, bool SQ, typename Epi>
inline void launch_colacc(cudaStream_t s, const ColaccPlan& pl, const T* M, size_t R, size_t C, size_t ld,
                          const T* w, T* part, unsigned* tickets, const Epi& epi, double* partials, Gate gate) {
  dim3 grid(pl.tiles, pl.chunks);
  k_colacc<T, SQ, 8, Epi><<<grid, kThreads, 0, s>>>(M, R, C, ld, w, pl.rows_per_chunk, part, tickets, epi,
                                                    partials, gate);
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

constexpr int kPlanOcc = 4;   // CTAs per SM the streaming kernels are built for (__launch_bounds__)

template <typename T>
class DenseMat {
 public:
  // `A` is m x n, row-major (rowmaj=true) or column-major.  on_device: A is a
  // device pointer (copied device-to-device), else a host pointer.
  DenseMat(bool rowmaj, size_t m, size_t n, const T* A, bool on_device, cudaStream_t stream)
      : m_(m), n_(n), tstore_(!rowmaj), stream_(stream) {
    dev_ = query_device();
    R_ = tstore_ ? n : m;
    C_ = tstore_ ? m : n;
    constexpr size_t VEC = V16<T>::N;
    ld_ = round_up(C_, VEC);
    data_.alloc(R_ * ld_);
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (ld_ == C_) {
      POGS_CUDA(cudaMemcpyAsync(data_.get(), A, R_ * C_ * sizeof(T), kind, stream_));
    } else {
      POGS_CUDA(cudaMemcpy2DAsync(data_.get(), ld_ * sizeof(T), A, C_ * sizeof(T), C_ * sizeof(T), R_, kind,
                                  stream_));
    }
    rd_plan_ = plan_rowdot(R_, dev_.sm_count, kPlanOcc);
    ca_plan_ = plan_colacc<T>(R_, ld_, dev_.sm_count, kPlanOcc);
    part_.alloc(static_cast<size_t>(ca_plan_.chunks) * ld_);
    tickets_.alloc(ca_plan_.tiles);
  }

  size_t rows() const { return m_; }
  size_t cols() const { return n_; }
  bool transposed_storage() const { return tstore_; }
  T* data() { return data_.get(); }
  size_t R() const { return R_; }
  size_t C() const { return C_; }
  size_t ld() const { return ld_; }
  const DeviceInfo& device() const { return dev_; }

  // Rows of the partials array written by mul_n / mul_t.
  unsigned nb_n() const { return tstore_ ? ca_plan_.tiles : rd_plan_.grid; }
  unsigned nb_t() const { return tstore_ ? rd_plan_.grid : ca_plan_.tiles; }
  unsigned nb_max() const { return rd_plan_.grid > ca_plan_.tiles ? rd_plan_.grid : ca_plan_.tiles; }

  // out(m) <- epi(A v),  v of length n (zero-padded buffer).
  template <bool SQ, typename Epi>
  void mul_n(const T* v, const Epi& epi, double* partials, Gate gate = Gate{nullptr, nullptr}) {
    if (!tstore_) launch_rowdot<T, SQ>(stream_, rd_plan_, data_.get(), R_, C_, ld_, v, epi, partials, gate);
    else launch_colacc<T, SQ>(stream_, ca_plan_, data_.get(), R_, C_, ld_, v, part_.get(), tickets_.get(), epi,
                              partials, gate);
  }
  // out(n) <- epi(A^T w),  w of length m.
  template <bool SQ, typename Epi>
  void mul_t(const T* w, const Epi& epi, double* partials, Gate gate = Gate{nullptr, nullptr}) {
    if (!tstore_) launch_colacc<T, SQ>(stream_, ca_plan_, data_.get(), R_, C_, ld_, w, part_.get(), tickets_.get(),
                                       epi, partials, gate);
    else launch_rowdot<T, SQ>(stream_, rd_plan_, data_.get(), R_, C_, ld_, w, epi, partials, gate);
  }

  // Modified Sinkhorn-Knopp on A.^2 (squares formed in registers, so neither a
  // squared copy nor the reference's sign bit-vector exists), Frobenius
  // normalisation, in-place A := D A E / normA.  d (m) and e (n) are outputs.
  void equilibrate(T* d, T* e) {
    const size_t m = m_, n = n_;
    const unsigned tb = 256;
    k_fill<T><<<(unsigned)((m + tb - 1) / tb), tb, 0, stream_>>>(m, T(1), d);
    k_fill<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, T(1), e);
    const T ce = T(1e-4) * static_cast<T>(m + n) / static_cast<T>(m);
    const T cd = T(1e-4) * static_cast<T>(m + n) / static_cast<T>(n);
    for (int k = 0; k < 50; ++k) {
      mul_t<true>(d, EpiSinkhorn<T>{static_cast<T>(m), ce, e}, nullptr);
      mul_n<true>(e, EpiSinkhorn<T>{static_cast<T>(n), cd, d}, nullptr);
    }
    k_sqrt_inplace<T><<<(unsigned)((m + tb - 1) / tb), tb, 0, stream_>>>(m, d);
    k_sqrt_inplace<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, e);
    // ||D A E||_F^2 = sum_i d_i^2 (A.^2 e.^2)_i
    DevBuf<T> e2(n), scal(2);
    DevBuf<double> fpart(nb_max());
    k_square<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, e, e2.get());
    mul_n<true>(e2.get(), EpiWeightedSum<T>{d}, fpart.get());
    const double min_dim = static_cast<double>(m < n ? m : n);
    k_fro_finish<T><<<1, kThreads, 0, stream_>>>(fpart.get(), nb_n(), min_dim, scal.get(), scal.get() + 1);
    const T* rs = tstore_ ? e : d;
    const T* cs = tstore_ ? d : e;
    k_scale_matrix<T><<<dev_.sm_count * 8, kThreads, 0, stream_>>>(data_.get(), R_, C_, ld_, rs, cs, scal.get());
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
    std::vector<T> x0(n);
    {
      std::default_random_engine gen;
      std::uniform_real_distribution<T> dist(static_cast<T>(0), static_cast<T>(1));
      for (size_t i = 0; i < n; ++i) x0[i] = dist(gen);
    }
    DevBuf<T> x(n), xn(n), Sx(m), inv(1);
    DevBuf<double> p_sx(nb_max()), p_x(nb_max());
    POGS_CUDA(cudaMemcpyAsync(x.get(), x0.data(), n * sizeof(T), cudaMemcpyHostToDevice, stream_));
    Ctrl<T> hc;
    POGS_CUDA(cudaMemcpyAsync(&hc, ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    hc.est = 0; hc.est_last = 0; hc.est_done = 0; hc.est_iters = 0;
    POGS_CUDA(cudaMemcpyAsync(ctrl, &hc, sizeof(hc), cudaMemcpyHostToDevice, stream_));
    Gate gate{&ctrl->est_done, nullptr};
    const unsigned tb = 256;
    for (int i = 0; i < 50; ++i) {
      mul_n<false>(x.get(), EpiAffine<T>{T(1), T(0), nullptr, Sx.get()}, p_sx.get(), gate);
      mul_t<false>(Sx.get(), EpiAffine<T>{T(1), T(0), nullptr, xn.get()}, p_x.get(), gate);
      k_normest_step<T><<<1, kThreads, 0, stream_>>>(ctrl, p_x.get(), nb_t(), p_sx.get(), nb_n(), inv.get());
      k_scale_copy<T><<<(unsigned)((n + tb - 1) / tb), tb, 0, stream_>>>(n, xn.get(), T(0), inv.get(), x.get());
    }
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaMemcpyAsync(&hc, ctrl, sizeof(hc), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    normest_iters_ = hc.est_iters;
    return hc.est;
  }
  unsigned normest_iters() const { return normest_iters_; }

 private:
  size_t m_, n_;
  bool tstore_;
  cudaStream_t stream_;
  DeviceInfo dev_;
  size_t R_, C_, ld_;
  DevBuf<T> data_, part_;
  DevBuf<unsigned> tickets_;
  RowdotPlan rd_plan_;
  ColaccPlan ca_plan_;
  unsigned normest_iters_ = 0;
};

}  // namespace pogs_b200
