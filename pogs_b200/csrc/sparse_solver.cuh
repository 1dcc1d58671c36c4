// Sparse graph-form solver (CSR + CSC copies in HBM, CGLS projector).
// Placeholder until the SpMV / CGLS kernels land: construction reports an error.
#pragma once

#include "dense_solver.cuh"

namespace pogs_b200 {

template <typename T>
class SparseSolver : public DenseSolver<T> {
 public:
  SparseSolver(bool, size_t, size_t, size_t, const T*, const int*, const int*)
      : DenseSolver<T>(true, 1, 1, nullptr, false) {
    throw Error("sparse path not built yet");
  }
};

}  // namespace pogs_b200
