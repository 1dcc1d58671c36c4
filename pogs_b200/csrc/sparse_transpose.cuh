// Compressed-row copy of the transpose (CSR of A -> CSR of A^T == CSC of A), on the device:
// what gsl::spmat's csr2csc does on the host (src/cpu/include/gsl/gsl_spmat.h:32-58) -- a counting
// sort of the entries by column that keeps the rows of a column in ascending order.
//
//   pack     : (row, value) of every entry; histogram of the columns (integer atomics: exact)
//   scan     : column pointers
//   sort     : stable LSD radix sort of the entries by column (cub::DeviceRadixSort, log2(cols) bits)
//   unpack   : values / row indices of the transposed copy
//
// Replaces cusparseCsr2cscEx2: the library was the last one on the link line, and the first use of
// it in a process cost seconds of page-in on a fresh machine (5.1 s of a 5.4 s first PogsSparseS call).
#pragma once

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace pogs_b200 {

template <typename T> struct TrEntry { int row; T val; };

template <typename T>
__global__ void __launch_bounds__(256)
k_tr_pack(const int* __restrict__ ptr, const int* __restrict__ ind, const T* __restrict__ val, size_t rows,
          TrEntry<T>* __restrict__ pk, int* __restrict__ colcnt) {
  const int lane = threadIdx.x & 31;
  const size_t w = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const size_t nw = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
  for (size_t r = w; r < rows; r += nw) {
    for (int k = ptr[r] + lane; k < ptr[r + 1]; k += 32) {
      pk[k] = TrEntry<T>{static_cast<int>(r), val[k]};
      atomicAdd(colcnt + ind[k], 1);
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
k_tr_unpack(const TrEntry<T>* __restrict__ pk, size_t nnz, T* __restrict__ val_t, int* __restrict__ ind_t) {
  for (size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz; k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const TrEntry<T> e = pk[k];
    val_t[k] = e.val;
    ind_t[k] = e.row;
  }
}

// (ptr[rows+1], ind, val) -> (ptr_t[cols+1], ind_t, val_t); all device pointers, ptr_t zero-initialised.
template <typename T>
void csr_transpose(const int* ptr, const int* ind, const T* val, size_t rows, size_t cols, size_t nnz, int* ptr_t,
                   int* ind_t, T* val_t, int sm_count, cudaStream_t stream) {
  if (nnz == 0) return;
  DevBuf<TrEntry<T>> pk(nnz), pk_s(nnz);
  DevBuf<int> key_s(nnz), colcnt(cols + 1);
  const unsigned grid = static_cast<unsigned>(sm_count) * 16;
  k_tr_pack<T><<<grid, 256, 0, stream>>>(ptr, ind, val, rows, pk.get(), colcnt.get());
  POGS_CUDA(cudaGetLastError());
  size_t tmp_bytes = 0;
  POGS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, colcnt.get(), ptr_t, static_cast<int>(cols + 1), stream));
  {
    DevBuf<char> tmp(tmp_bytes);
    POGS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, colcnt.get(), ptr_t, static_cast<int>(cols + 1), stream));
    POGS_CUDA(cudaStreamSynchronize(stream));
  }
  int bits = 1;
  while ((size_t(1) << bits) < cols) ++bits;
  POGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ind, key_s.get(), pk.get(), pk_s.get(), static_cast<int>(nnz), 0,
                                            bits, stream));
  {
    DevBuf<char> tmp(tmp_bytes);
    POGS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, ind, key_s.get(), pk.get(), pk_s.get(), static_cast<int>(nnz),
                                              0, bits, stream));
    POGS_CUDA(cudaStreamSynchronize(stream));
  }
  k_tr_unpack<T><<<grid, 256, 0, stream>>>(pk_s.get(), nnz, val_t, ind_t);
  POGS_CUDA(cudaGetLastError());
  POGS_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace pogs_b200
