// Device-resident sparse operator: two compressed copies of A in HBM -- CSR for
// A v and CSC (== CSR of A^T) for A^T w -- so that both products are row-gather
// SpMVs, as in the reference's MatrixSparse (src/cpu/matrix/matrix_sparse.cpp:97-155,
// src/cpu/include/gsl/gsl_spmat.h:32-98: 2*nnz values and indices, m+n+2 pointers,
// int32 indices).  The transposed copy is built once on the device
// (sparse_transpose.cuh).  Equilibration / norm estimate come from MatAlgos.
#pragma once

#include <cmath>
#include <cub/device/device_scan.cuh>

#include "mat_algos.cuh"
#include "sparse_kernels.cuh"
#include "sparse_tiled.cuh"
#include "sparse_transpose.cuh"

namespace pogs_b200 {

template <typename T>
class SparseMat : public MatAlgos<SparseMat<T>, T> {
 public:
  static constexpr bool kDense = false;

  // rowmaj: (val, ptr[m+1], ind) is CSR; else CSC (ptr[n+1]).  Host pointers.
  SparseMat(bool rowmaj, size_t m, size_t n, size_t nnz, const T* val, const int* ptr, const int* ind,
            cudaStream_t stream)
      : MatAlgos<SparseMat<T>, T>(m, n, stream), nnz_(nnz) {
    if (nnz > 0x7fffffffULL || m > 0x7fffffffULL || n > 0x7fffffffULL)
      throw Error("sparse dimensions / nnz must fit int32 (POGS_INT, matrix_sparse.h:10)");
    // copy 0 = rows of A (CSR), copy 1 = rows of A^T (CSC of A)
    const int given = rowmaj ? 0 : 1, other = 1 - given;
    const size_t len_given = (rowmaj ? m : n) + 1, len_other = (rowmaj ? n : m) + 1;
    val_[given].alloc(nnz); ind_[given].alloc(nnz); ptr_[given].alloc(len_given);
    val_[other].alloc(nnz); ind_[other].alloc(nnz); ptr_[other].alloc(len_other);
    POGS_CUDA(cudaMemcpyAsync(val_[given].get(), val, nnz * sizeof(T), cudaMemcpyHostToDevice, stream));
    POGS_CUDA(cudaMemcpyAsync(ind_[given].get(), ind, nnz * sizeof(int), cudaMemcpyHostToDevice, stream));
    POGS_CUDA(cudaMemcpyAsync(ptr_[given].get(), ptr, len_given * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (nnz > 0) {
      const size_t rows_g = rowmaj ? m : n, cols_g = rowmaj ? n : m;
      csr_transpose<T>(ptr_[given].get(), ind_[given].get(), val_[given].get(), rows_g, cols_g, nnz, ptr_[other].get(),
                       ind_[other].get(), val_[other].get(), this->dev_.sm_count, stream);
    }
    rows_[0] = m; rows_[1] = n;
    cols_[0] = n; cols_[1] = m;
    for (int c = 0; c < 2; ++c) {
      const double avg = rows_[c] > 0 ? static_cast<double>(nnz) / rows_[c] : 0.0;
      int lg = 0;
      while (lg < 5 && (1 << lg) * 4 < avg) ++lg;    // ~4+ entries per lane before widening the group
      lg_[c] = lg;
      const size_t threads = rows_[c] << lg;
      const size_t need = (threads + kThreads - 1) / kThreads;
      const size_t cap = static_cast<size_t>(this->dev_.sm_count) * 8;
      grid_[c] = static_cast<unsigned>(need < cap ? (need > 0 ? need : 1) : cap);
    }
    // 2-D tiled re-layout (sparse_tiled.cuh) for matrices that do not live in L2 anyway (C5, 1e8 entries:
    // 565 it/s on the plain copies -> 1690 it/s); POGS_B200_SPMV=tiled|plain forces the choice.
    const char* sel = getenv("POGS_B200_SPMV");
    bool want_tiled = nnz >= (size_t(1) << 22);
    if (sel != nullptr && sel[0] == 't') want_tiled = true;
    if (sel != nullptr && sel[0] == 'p') want_tiled = false;
    if (want_tiled && nnz > 0) {
      for (int c = 0; c < 2; ++c) {
        if (tiled_[c].build(ptr_[c].get(), ind_[c].get(), val_[c].get(), rows_[c], cols_[c], nnz_,
                            static_cast<unsigned>(this->dev_.sm_count), stream)) {
          const unsigned sms = static_cast<unsigned>(this->dev_.sm_count);
          tiled_grid_[c] = tiled_[c].sh.ntiles < sms ? tiled_[c].sh.ntiles : sms;
          // test knob: fewer CTAs than tiles, so that a CTA walks several tiles (on one GPU generation the planner
          // nearly always ends at one tile per CTA)
          const char* tg = getenv("POGS_B200_TL_GRID");
          if (tg != nullptr && atoi(tg) > 0 && static_cast<unsigned>(atoi(tg)) < tiled_grid_[c])
            tiled_grid_[c] = static_cast<unsigned>(atoi(tg));
          grid_[c] = tiled_[c].fold_grid;
          val_[c].release(); ind_[c].release();
        }
      }
    }
  }

  size_t nnz() const { return nnz_; }
  unsigned nb_n() const { return grid_[0]; }
  unsigned nb_t() const { return grid_[1]; }
  unsigned nb_max() const { return grid_[0] > grid_[1] ? grid_[0] : grid_[1]; }

  template <bool SQ, typename Epi>
  void mul_n(const T* v, const Epi& epi, double* partials, Gate gate = Gate{nullptr, nullptr}) {
    launch<SQ>(0, v, epi, partials, gate);
  }
  template <bool SQ, typename Epi>
  void mul_t(const T* w, const Epi& epi, double* partials, Gate gate = Gate{nullptr, nullptr}) {
    launch<SQ>(1, w, epi, partials, gate);
  }

  template <bool SQ, typename Epi>
  bool mul_n_tail(const T* v, const Epi& epi, double* partials, Gate gate, const TailCtrl<T>&) {
    mul_n<SQ>(v, epi, partials, gate);
    return false;
  }

  // both copies: val *= d[row of A] * e[col of A] * (*s)   (matrix_sparse.cpp:293-304)
  void apply_scaling(const T* d, const T* e, const T* s_ptr) {
    for (int c = 0; c < 2; ++c) {
      const T* rs = c == 0 ? d : e;
      const T* cs = c == 0 ? e : d;
      if (tiled_[c].ok) {
        TiledCopy<T>& tc = tiled_[c];
        const size_t nsl = static_cast<size_t>(tc.sh.ntiles) * tc.sh.ns;
        const unsigned g = static_cast<unsigned>(std::min<size_t>((nsl + 7) / 8, 148u * 64u));
        k_spscale_tiled<T><<<g, 256, 0, this->stream_>>>(tc.stream.get(), tc.soff.get(), tc.rowid.get(), tc.sh, rs, cs, s_ptr);
      } else {
        k_spscale<T><<<grid_[c], kThreads, 0, this->stream_>>>(val_[c].get(), ind_[c].get(), ptr_[c].get(), rows_[c],
                                                              lg_[c], rs, cs, s_ptr);
      }
    }
    POGS_CUDA(cudaGetLastError());
    count_launch(2);
  }
  // kernels per product: the tiled product is followed by the kernel that folds the partial sums
  unsigned launches_per_product() const { return tiled_[0].ok || tiled_[1].ok ? 2u : 1u; }

 private:
  template <bool SQ, typename Epi>
  void launch(int c, const T* v, const Epi& epi, double* partials, Gate gate) {
    if (tiled_[c].ok) {
      TiledCopy<T>& tc = tiled_[c];
      tl_launch<T, SQ>(tiled_grid_[c], tc.smem, this->stream_, tc.stream.get(), tc.soff.get(), tc.rowid.get(),
                       tc.wsplit.get(), rows_[c], cols_[c], tc.sh, v, tc.part.get(), gate);
      k_tl_fold<T, Epi><<<tc.fold_grid, kThreads, 0, this->stream_>>>(tc.part.get(), rows_[c], tc.sh.Q, epi, partials, gate);
      count_launch();
    } else {
      k_spmv<T, SQ, Epi><<<grid_[c], kThreads, 0, this->stream_>>>(val_[c].get(), ind_[c].get(), ptr_[c].get(),
                                                                   rows_[c], lg_[c], v, epi, partials, gate);
    }
    POGS_CUDA(cudaGetLastError());
    count_launch();
  }

  size_t nnz_;
  DevBuf<T> val_[2];
  DevBuf<int> ind_[2], ptr_[2];
  size_t rows_[2], cols_[2];
  int lg_[2];
  unsigned grid_[2];
  // 2-D tiled layout
  TiledCopy<T> tiled_[2];
  unsigned tiled_grid_[2] = {1, 1};
};

}  // namespace pogs_b200
