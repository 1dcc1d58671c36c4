// Device-resident sparse operator: two compressed copies of A in HBM -- CSR for
// A v and CSC (== CSR of A^T) for A^T w -- so that both products are row-gather
// SpMVs, as in the reference's MatrixSparse (src/cpu/matrix/matrix_sparse.cpp:97-155,
// src/cpu/include/gsl/gsl_spmat.h:32-98: 2*nnz values and indices, m+n+2 pointers,
// int32 indices).  The transposed copy is built once on the device (cuSPARSE
// csr2csc, a one-time layout conversion).  Equilibration / norm estimate come
// from MatAlgos.
#pragma once

#include <cusparse.h>

#include <cmath>
#include <cub/device/device_scan.cuh>

#include "mat_algos.cuh"
#include "sparse_kernels.cuh"

namespace pogs_b200 {

#define POGS_CUSPARSE(expr)                                                                 \
  do {                                                                                      \
    cusparseStatus_t _s = (expr);                                                           \
    if (_s != CUSPARSE_STATUS_SUCCESS)                                                      \
      throw ::pogs_b200::Error(std::string("cuSPARSE error ") + std::to_string((int)_s) +   \
                               " at " __FILE__ ":" + std::to_string(__LINE__));             \
  } while (0)

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
      cusparseHandle_t h;
      POGS_CUSPARSE(cusparseCreate(&h));
      POGS_CUSPARSE(cusparseSetStream(h, stream));
      const cudaDataType dt = sizeof(T) == 4 ? CUDA_R_32F : CUDA_R_64F;
      const int rows_g = static_cast<int>(rowmaj ? m : n), cols_g = static_cast<int>(rowmaj ? n : m);
      size_t ws = 0;
      POGS_CUSPARSE(cusparseCsr2cscEx2_bufferSize(h, rows_g, cols_g, static_cast<int>(nnz), val_[given].get(),
                                                  ptr_[given].get(), ind_[given].get(), val_[other].get(),
                                                  ptr_[other].get(), ind_[other].get(), dt, CUSPARSE_ACTION_NUMERIC,
                                                  CUSPARSE_INDEX_BASE_ZERO, CUSPARSE_CSR2CSC_ALG1, &ws));
      DevBuf<char> work(ws);
      POGS_CUSPARSE(cusparseCsr2cscEx2(h, rows_g, cols_g, static_cast<int>(nnz), val_[given].get(),
                                       ptr_[given].get(), ind_[given].get(), val_[other].get(), ptr_[other].get(),
                                       ind_[other].get(), dt, CUSPARSE_ACTION_NUMERIC, CUSPARSE_INDEX_BASE_ZERO,
                                       CUSPARSE_CSR2CSC_ALG1, work.get()));
      POGS_CUDA(cudaStreamSynchronize(stream));
      cusparseDestroy(h);
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
    // Column-blocked re-layout (sparse_kernels.cuh) for matrices that do not live in L2 anyway
    // (C5, 1e8 entries: 565 -> 911 ADMM iterations/s); POGS_B200_SPMV=blocked|plain forces the choice.
    const char* sel = getenv("POGS_B200_SPMV");
    bool want_blocked = nnz >= (size_t(1) << 22);
    if (sel != nullptr && sel[0] == 'b') want_blocked = true;
    if (sel != nullptr && sel[0] == 'p') want_blocked = false;
    if (want_blocked && nnz > 0) {
      for (int c = 0; c < 2; ++c) build_blocked(c, stream);
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
      if (blocked_[c]) {
        k_spscale_blocked<T><<<shape_[c].ncta, kSpThreads, 0, this->stream_>>>(bval_[c].get(), bind_[c].get(), seg_[c].get(),
                                                                               rows_[c], shape_[c], rs, cs, s_ptr);
      } else {
        k_spscale<T><<<grid_[c], kThreads, 0, this->stream_>>>(val_[c].get(), ind_[c].get(), ptr_[c].get(), rows_[c],
                                                              lg_[c], rs, cs, s_ptr);
      }
    }
    POGS_CUDA(cudaGetLastError());
    count_launch(2);
  }
  bool blocked() const { return blocked_[0] && blocked_[1]; }

 private:
  template <bool SQ, typename Epi>
  void launch(int c, const T* v, const Epi& epi, double* partials, Gate gate) {
    if (blocked_[c]) {
      auto kernel = k_spmv_blocked<T, SQ, Epi>;
      static size_t attr_smem_dev[kMaxDevices] = {};   // per instantiation and device
      size_t& attr_smem = attr_smem_dev[current_device_index()];
      if (attr_smem < bsmem_[c]) {
        POGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bsmem_[c])));
        attr_smem = bsmem_[c];
      }
      kernel<<<shape_[c].ncta, kSpThreads, bsmem_[c], this->stream_>>>(bval_[c].get(), bind_[c].get(), seg_[c].get(),
                                                                        rows_[c], cols_[c], shape_[c], blg_[c], v, epi,
                                                                        partials, gate);
    } else {
      k_spmv<T, SQ, Epi><<<grid_[c], kThreads, 0, this->stream_>>>(val_[c].get(), ind_[c].get(), ptr_[c].get(),
                                                                   rows_[c], lg_[c], v, epi, partials, gate);
    }
    POGS_CUDA(cudaGetLastError());
    count_launch();
  }

  // Re-lay copy c into (row range) x (column block) x (row) order with 16-bit local column
  // indices; frees the plain copy.  Not applicable (plain copy kept) when the row sums of one
  // range do not fit next to a useful column block in shared memory.
  void build_blocked(int c, cudaStream_t stream) {
    const size_t rows = rows_[c], cols = cols_[c];
    BlockedShape sh;
    sh.ncta = static_cast<unsigned>(this->dev_.sm_count);
    sh.rpc = static_cast<unsigned>((rows + sh.ncta - 1) / sh.ncta);
    const size_t budget = 200u * 1024u;
    const size_t acc_bytes = round_up(static_cast<size_t>(sh.rpc) * sizeof(T), 16);
    if (acc_bytes + 4096 * sizeof(T) > budget) return;
    size_t bc = (budget - acc_bytes) / sizeof(T);
    if (bc > 49152) bc = 49152;
    bc = bc / 32 * 32;
    size_t nblk = (cols + bc - 1) / bc;
    if (nblk > 128) return;
    bc = round_up((cols + nblk - 1) / nblk, 32);
    sh.nblk = static_cast<unsigned>(nblk);
    sh.blk_cols = static_cast<unsigned>(bc);
    const size_t L = static_cast<size_t>(sh.ncta) * sh.nblk * (sh.rpc + 1);
    if (L > 0x7fffffffULL) return;
    DevBuf<int> cnt(L);
    seg_[c].alloc(L);
    const unsigned tb = 256, gb = static_cast<unsigned>((rows + tb - 1) / tb);
    k_blk_count<<<gb, tb, 0, stream>>>(ptr_[c].get(), ind_[c].get(), rows, sh, cnt.get());
    size_t tmp_bytes = 0;
    POGS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.get(), seg_[c].get(), static_cast<int>(L), stream));
    DevBuf<char> tmp(tmp_bytes);
    POGS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, cnt.get(), seg_[c].get(), static_cast<int>(L), stream));
    // padded size = end of the last list (the padding entries keep the zero the buffers are created with)
    int last_seg = 0, last_cnt = 0;
    POGS_CUDA(cudaMemcpyAsync(&last_seg, seg_[c].get() + (L - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    POGS_CUDA(cudaMemcpyAsync(&last_cnt, cnt.get() + (L - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
    POGS_CUDA(cudaStreamSynchronize(stream));
    const size_t padded = static_cast<size_t>(last_seg) + static_cast<size_t>(last_cnt);
    if (padded > 0x7fffffffULL - 8) return;
    bnnz_[c] = padded;
    bval_[c].alloc(padded + 8); bind_[c].alloc(padded + 8);
    k_blk_scatter<T><<<gb, tb, 0, stream>>>(ptr_[c].get(), ind_[c].get(), val_[c].get(), rows, sh, seg_[c].get(),
                                           bval_[c].get(), bind_[c].get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaStreamSynchronize(stream));
    shape_[c] = sh;
    bsmem_[c] = static_cast<size_t>(sh.blk_cols) * sizeof(T) + acc_bytes;
    // lanes per row segment: a trip covers two 4-entry vectors per lane; wide enough for the average
    // segment (plus its spread) to finish in one trip
    const double avg_seg = rows > 0 ? static_cast<double>(nnz_) / rows / sh.nblk : 0.0;
    const double want_vec = (avg_seg + 1.3 * std::sqrt(avg_seg > 0 ? avg_seg : 0.0)) / 4.0 + 1.0;
    int lg = 0;
    while (lg < 5 && (1 << lg) * 2 < want_vec) ++lg;
    blg_[c] = lg;
    grid_[c] = sh.ncta;
    blocked_[c] = true;
    val_[c].release(); ind_[c].release();
  }

  size_t nnz_;
  DevBuf<T> val_[2];
  DevBuf<int> ind_[2], ptr_[2];
  size_t rows_[2], cols_[2];
  int lg_[2];
  unsigned grid_[2];
  // column-blocked layout
  bool blocked_[2] = {false, false};
  BlockedShape shape_[2];
  DevBuf<T> bval_[2];
  DevBuf<unsigned short> bind_[2];
  DevBuf<int> seg_[2];
  size_t bsmem_[2] = {0, 0};
  size_t bnnz_[2] = {0, 0};   // entries of the blocked copy incl. the padding of the row segments
  int blg_[2] = {0, 0};
};

}  // namespace pogs_b200
