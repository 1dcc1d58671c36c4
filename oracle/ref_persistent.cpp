// TEST INFRASTRUCTURE ONLY -- thin extern "C" handle around the *reference's own*
// persistent C++ solver objects (pogs::PogsDirect / pogs::PogsIndirect,
// /root/reference/src/include/pogs.h:55-131,155-158), compiled together with the
// unmodified reference sources by oracle/build_ref.sh.  It exists because the
// reference's C ABI is one-shot (src/interface_c/pogs_c.cpp:19-54) while the
// warm-start / lambda-path behaviour (examples/cpp/lasso_path.cpp:75-107) lives
// only in the C++ object.  This file contains no reference code; it only calls
// the reference's public methods.
#include <vector>

#include "matrix/matrix_dense.h"
#include "matrix/matrix_sparse.h"
#include "pogs.h"

namespace {

template <typename T>
struct DenseHandle {
  pogs::MatrixDense<T> A;
  pogs::PogsDirect<T, pogs::MatrixDense<T>> solver;
  size_t m, n;
  DenseHandle(char ord, size_t m, size_t n, const T* data) : A(ord, m, n, data), solver(A), m(m), n(n) {}
};

template <typename T>
std::vector<FunctionObj<T>> make(size_t len, const int* h, const T* a, const T* b, const T* c, const T* d,
                                 const T* e) {
  std::vector<FunctionObj<T>> v;
  v.reserve(len);
  for (size_t i = 0; i < len; ++i) v.emplace_back(static_cast<Function>(h[i]), a[i], b[i], c[i], d[i], e[i]);
  return v;
}

template <typename T>
int solve(DenseHandle<T>* H, const int* fh, const T* fa, const T* fb, const T* fc, const T* fd, const T* fe,
          const int* gh, const T* ga, const T* gb, const T* gc, const T* gd, const T* ge, T rho, int set_rho,
          T abs_tol, T rel_tol, unsigned max_iter, int adaptive_rho, int gap_stop, T* x, T* y, T* l, T* mu,
          T* optval, unsigned* final_iter, T* rho_out) {
  auto f = make<T>(H->m, fh, fa, fb, fc, fd, fe);
  auto g = make<T>(H->n, gh, ga, gb, gc, gd, ge);
  if (set_rho) H->solver.SetRho(rho);
  H->solver.SetAbsTol(abs_tol);
  H->solver.SetRelTol(rel_tol);
  H->solver.SetMaxIter(max_iter);
  H->solver.SetVerbose(0);
  H->solver.SetAdaptiveRho(adaptive_rho != 0);
  H->solver.SetGapStop(gap_stop != 0);
  int st = H->solver.Solve(f, g);
  for (size_t j = 0; j < H->n; ++j) { x[j] = H->solver.GetX()[j]; mu[j] = H->solver.GetMu()[j]; }
  for (size_t i = 0; i < H->m; ++i) { y[i] = H->solver.GetY()[i]; l[i] = H->solver.GetLambda()[i]; }
  *optval = H->solver.GetOptval();
  *final_iter = H->solver.GetFinalIter();
  *rho_out = H->solver.GetRho();
  return st;
}

}  // namespace

extern "C" {
// `data` must stay alive until the first solve (the reference copies lazily in Init).
void* refp_create_dense_d(int rowmaj, size_t m, size_t n, const double* data) {
  return new DenseHandle<double>(rowmaj ? 'r' : 'c', m, n, data);
}
void* refp_create_dense_s(int rowmaj, size_t m, size_t n, const float* data) {
  return new DenseHandle<float>(rowmaj ? 'r' : 'c', m, n, data);
}
void refp_destroy_d(void* h) { delete static_cast<DenseHandle<double>*>(h); }
void refp_destroy_s(void* h) { delete static_cast<DenseHandle<float>*>(h); }
void refp_set_init_d(void* h, const double* x, const double* l) {
  auto* H = static_cast<DenseHandle<double>*>(h);
  if (x) H->solver.SetInitX(x);
  if (l) H->solver.SetInitLambda(l);
}
void refp_set_init_s(void* h, const float* x, const float* l) {
  auto* H = static_cast<DenseHandle<float>*>(h);
  if (x) H->solver.SetInitX(x);
  if (l) H->solver.SetInitLambda(l);
}
int refp_solve_d(void* h, const int* fh, const double* fa, const double* fb, const double* fc, const double* fd,
                 const double* fe, const int* gh, const double* ga, const double* gb, const double* gc,
                 const double* gd, const double* ge, double rho, int set_rho, double abs_tol, double rel_tol,
                 unsigned max_iter, int adaptive_rho, int gap_stop, double* x, double* y, double* l, double* mu,
                 double* optval, unsigned* final_iter, double* rho_out) {
  return solve<double>(static_cast<DenseHandle<double>*>(h), fh, fa, fb, fc, fd, fe, gh, ga, gb, gc, gd, ge, rho,
                       set_rho, abs_tol, rel_tol, max_iter, adaptive_rho, gap_stop, x, y, l, mu, optval,
                       final_iter, rho_out);
}
int refp_solve_s(void* h, const int* fh, const float* fa, const float* fb, const float* fc, const float* fd,
                 const float* fe, const int* gh, const float* ga, const float* gb, const float* gc,
                 const float* gd, const float* ge, float rho, int set_rho, float abs_tol, float rel_tol,
                 unsigned max_iter, int adaptive_rho, int gap_stop, float* x, float* y, float* l, float* mu,
                 float* optval, unsigned* final_iter, float* rho_out) {
  return solve<float>(static_cast<DenseHandle<float>*>(h), fh, fa, fb, fc, fd, fe, gh, ga, gb, gc, gd, ge, rho,
                      set_rho, abs_tol, rel_tol, max_iter, adaptive_rho, gap_stop, x, y, l, mu, optval, final_iter,
                      rho_out);
}
}
