// extern "C" boundary of libpogs_b200.so -- see include/pogs_b200.h.
//
// Part 1 mirrors the reference's src/interface_c/pogs_c.cpp:8-203 (construct,
// set parameters, solve once, copy x / y / lambda / optval / final_iter out,
// throw everything away).  No C++ exception crosses the boundary: failures are
// reported as POGS_ERROR with the message kept for pogs_b200_last_error().
#include "../../include/pogs_b200.h"

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <condition_variable>

#include "graph_solver.cuh"

using namespace pogs_b200;

namespace {

thread_local std::string g_last_error;

// One mutex per device: calls that use the same device are serialised (they share its
// memory pool and the per-kernel attribute caches), calls on
// different devices run concurrently.
std::mutex& device_mutex(int dev) {
  static std::mutex table_mu;
  static std::vector<std::mutex*>* table = new std::vector<std::mutex*>();   // never destroyed
  std::lock_guard<std::mutex> lock(table_mu);
  if (dev < 0) dev = 0;
  while (static_cast<size_t>(dev) >= table->size()) table->push_back(new std::mutex());
  return *(*table)[dev];
}

// Locks the device's mutex, makes the device current for the calling thread and restores the
// caller's current device on exit.  want < 0: use the calling thread's current device (what the
// one-shot entry points and the create functions do); handles remember the device they live on.
struct DeviceScope {
  int prev = -1, dev = 0;
  std::unique_lock<std::mutex> lock;
  explicit DeviceScope(int want = -1) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    dev = want >= 0 ? want : (prev >= 0 ? prev : 0);
    lock = std::unique_lock<std::mutex>(device_mutex(dev));
    if (want >= 0 && want != prev) cudaSetDevice(want);
  }
  ~DeviceScope() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};

int fail(const std::exception& e) {
  g_last_error = e.what();
  fprintf(stderr, "pogs_b200: %s\n", e.what());
  return POGS_ERROR;
}

void require_device() {
  int count = 0;
  const cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    throw Error("no usable CUDA device (this library has no CPU fallback)");
  }
}

}  // namespace

struct pogs_b200_comm {
  int device = 0;
  PeerComm* c = nullptr;
  ~pogs_b200_comm() { delete c; }
};

struct pogs_b200_handle {
  int device = 0;
  int is_double = 0;
  SolverBase<float>* s = nullptr;
  SolverBase<double>* d = nullptr;
  ~pogs_b200_handle() { delete s; delete d; }
};

namespace {

template <typename T> SolverBase<T>* impl(pogs_b200_handle* h);
template <> SolverBase<float>* impl<float>(pogs_b200_handle* h) {
  if (h == nullptr || h->s == nullptr) throw Error("handle is null or not single precision");
  return h->s;
}
template <> SolverBase<double>* impl<double>(pogs_b200_handle* h) {
  if (h == nullptr || h->d == nullptr) throw Error("handle is null or not double precision");
  return h->d;
}

template <typename T>
void set_params(SolverBase<T>* s, T rho, T abs_tol, T rel_tol, unsigned max_iter, unsigned verbose,
                bool adaptive_rho, bool gap_stop) {
  s->SetRho(rho); s->SetAbsTol(abs_tol); s->SetRelTol(rel_tol); s->SetMaxIter(max_iter);
  s->SetVerbose(verbose); s->SetAdaptiveRho(adaptive_rho); s->SetGapStop(gap_stop);
}

template <typename T>
int one_shot(SolverBase<T>& solver, size_t m, size_t n, const T* f_a, const T* f_b, const T* f_c, const T* f_d,
             const T* f_e, const FUNCTION* f_h, const T* g_a, const T* g_b, const T* g_c, const T* g_d,
             const T* g_e, const FUNCTION* g_h, T rho, T abs_tol, T rel_tol, unsigned max_iter,
             unsigned verbose, int adaptive_rho, int gap_stop, T* x, T* y, T* l, T* optval,
             unsigned* final_iter) {
  static_assert(sizeof(FUNCTION) == sizeof(int), "enum FUNCTION must be int-sized");
  set_params<T>(&solver, rho, abs_tol, rel_tol, max_iter, verbose, adaptive_rho != 0, gap_stop != 0);
  const int status = solver.Solve(f_a, f_b, f_c, f_d, f_e, reinterpret_cast<const int*>(f_h), g_a, g_b, g_c,
                                  g_d, g_e, reinterpret_cast<const int*>(g_h));
  *optval = solver.GetOptval();
  *final_iter = solver.GetFinalIter();
  std::memcpy(x, solver.GetX(), n * sizeof(T));
  std::memcpy(y, solver.GetY(), m * sizeof(T));
  std::memcpy(l, solver.GetLambda(), m * sizeof(T));
  return status;
}

// ---- one call, several GPUs (SURVEY 8e process model) ---------------------------------------------------------
// The reference's C ABI is one blocking call from one host process.  With POGS_B200_GPUS=G (2..8) in the
// environment, PogsS / PogsD split a row-major tall matrix into G row blocks and drive the row-block solver
// (the one `torchrun` launches one process per GPU for) from G host threads of THIS process: thread g owns
// device g, uploads its block from the caller's host array, and the threads meet only inside the kernels
// (NVLink peer memory).  x comes from rank 0, y and lambda are written block by block.
struct HostBarrier {
  std::mutex mu; std::condition_variable cv; int count = 0, gen = 0, n;
  explicit HostBarrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    const int g = gen;
    if (++count == n) { count = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

int requested_gpus() {
  const char* e = getenv("POGS_B200_GPUS");
  if (e == nullptr) return 1;
  int g = atoi(e), have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess) { cudaGetLastError(); return 1; }
  if (g > have) g = have;
  if (g > kMaxPeers) g = kMaxPeers;
  return g < 1 ? 1 : g;
}

template <typename T>
int dense_entry_multi(int G, size_t m, size_t n, const T* A, const T* f_a, const T* f_b, const T* f_c, const T* f_d,
                      const T* f_e, const FUNCTION* f_h, const T* g_a, const T* g_b, const T* g_c, const T* g_d,
                      const T* g_e, const FUNCTION* g_h, T rho, T abs_tol, T rel_tol, unsigned max_iter, unsigned verbose,
                      int adaptive_rho, int gap_stop, T* x, T* y, T* l, T* optval, unsigned* final_iter) {
  // contiguous balanced row blocks, multiples of 4 rows (== pogs_b200/dist.py: row_partition)
  const size_t base = (m / G) / 4 * 4;
  std::vector<size_t> r0(G + 1);
  for (int g = 0; g < G; ++g) r0[g] = g * base;
  r0[G] = m;
  std::vector<PeerComm*> comms(G, nullptr);
  std::vector<void*> bases(G, nullptr);
  std::vector<int> status(G, POGS_ERROR);
  std::vector<std::string> errors(G);
  HostBarrier bar(G);
  std::atomic<int> failed{0};
  auto worker = [&](int g) {
    DeviceScope scope(g);
    bool in_step = true;   // while true this thread still owes its peers the two barrier visits
    try {
      for (int q = 0; q < G; ++q) {
        if (q == g) continue;
        const cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) POGS_CUDA(e);
        cudaGetLastError();
      }
      comms[g] = new PeerComm(g, G, 8 * (n + 64) < (size_t(1) << 22) ? (size_t(1) << 22) : 8 * (n + 64));
      bases[g] = comms[g]->local_base();
    } catch (const std::exception& e) { errors[g] = e.what(); failed.fetch_add(1); }
    bar.wait();   // every region exists
    try {
      if (failed.load() == 0) comms[g]->open_peers_in_process(bases.data());
    } catch (const std::exception& e) { errors[g] = e.what(); failed.fetch_add(1); }
    bar.wait();   // every rank sees every region
    in_step = false;
    if (failed.load() != 0) return;
    try {
      const size_t a = r0[g], rows = r0[g + 1] - r0[g];
      DenseSolver<T> solver(true, rows, n, A + a * n, false, true, m, comms[g]);
      set_params<T>(&solver, rho, abs_tol, rel_tol, max_iter, g == 0 ? verbose : 0u, adaptive_rho != 0, gap_stop != 0);
      status[g] = solver.Solve(f_a + a, f_b + a, f_c + a, f_d + a, f_e + a, reinterpret_cast<const int*>(f_h) + a, g_a, g_b, g_c,
                               g_d, g_e, reinterpret_cast<const int*>(g_h));
      std::memcpy(y + a, solver.GetY(), rows * sizeof(T));
      std::memcpy(l + a, solver.GetLambda(), rows * sizeof(T));
      if (g == 0) {
        std::memcpy(x, solver.GetX(), n * sizeof(T));
        *optval = solver.GetOptval();
        *final_iter = solver.GetFinalIter();
      }
    } catch (const std::exception& e) { errors[g] = e.what(); failed.fetch_add(1); status[g] = POGS_ERROR; }
    (void)in_step;
  };
  std::vector<std::thread> threads;
  for (int g = 1; g < G; ++g) threads.emplace_back(worker, g);
  worker(0);
  for (auto& t : threads) t.join();
  for (int g = 0; g < G; ++g) {
    if (comms[g] != nullptr) { DeviceScope scope(g); delete comms[g]; }
  }
  for (int g = 0; g < G; ++g)
    if (!errors[g].empty()) throw Error("GPU " + std::to_string(g) + ": " + errors[g]);
  return status[0];
}

template <typename T>
int dense_entry(ORD ord, size_t m, size_t n, const T* A, const T* f_a, const T* f_b, const T* f_c, const T* f_d,
                const T* f_e, const FUNCTION* f_h, const T* g_a, const T* g_b, const T* g_c, const T* g_d,
                const T* g_e, const FUNCTION* g_h, T rho, T abs_tol, T rel_tol, unsigned max_iter,
                unsigned verbose, int adaptive_rho, int gap_stop, T* x, T* y, T* l, T* optval,
                unsigned* final_iter) {
  {
    // several GPUs behind the one call (row-major, tall, enough rows per GPU): see dense_entry_multi
    int G = 1;
    try { require_device(); G = requested_gpus(); } catch (const std::exception& e) { return fail(e); }
    if (G > 1 && ord == ROW_MAJ && m > n && m / G >= 1024) {
      int prev = 0;
      cudaGetDevice(&prev);
      try {
        const int st = dense_entry_multi<T>(G, m, n, A, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h, rho, abs_tol,
                                            rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
        cudaSetDevice(prev);
        return st;
      } catch (const std::exception& e) {
        cudaSetDevice(prev);
        return fail(e);
      }
    }
  }
  DeviceScope scope;
  try {
    require_device();
    Trace tr;
    int status;
    {
      DenseSolver<T> solver(ord == ROW_MAJ, m, n, A, false);
      status = one_shot<T>(solver, m, n, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h, rho, abs_tol,
                           rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
      tr.mark("solve + copy out");
    }
    tr.mark("teardown");
    return status;
  } catch (const std::exception& e) {
    return fail(e);
  }
}

template <typename T>
int sparse_entry(ORD ord, size_t m, size_t n, size_t nnz, const T* data, const int* ptr, const int* ind,
                 const T* f_a, const T* f_b, const T* f_c, const T* f_d, const T* f_e, const FUNCTION* f_h,
                 const T* g_a, const T* g_b, const T* g_c, const T* g_d, const T* g_e, const FUNCTION* g_h,
                 T rho, T abs_tol, T rel_tol, unsigned max_iter, unsigned verbose, int adaptive_rho,
                 int gap_stop, T* x, T* y, T* l, T* optval, unsigned* final_iter) {
  DeviceScope scope;
  try {
    require_device();
    SparseSolver<T> solver(ord == ROW_MAJ, m, n, nnz, data, ptr, ind);
    return one_shot<T>(solver, m, n, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h, rho, abs_tol,
                       rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
  } catch (const std::exception& e) {
    return fail(e);
  }
}

template <typename T>
int get_solution(pogs_b200_handle* h, T* x, T* y, T* lambda, T* mu, T* optval, unsigned* final_iter, T* rho) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    SolverBase<T>* s = impl<T>(h);
    if (x) std::memcpy(x, s->GetX(), s->Cols() * sizeof(T));
    if (y) std::memcpy(y, s->GetY(), s->Rows() * sizeof(T));
    if (lambda) std::memcpy(lambda, s->GetLambda(), s->Rows() * sizeof(T));
    if (mu) std::memcpy(mu, s->GetMu(), s->Cols() * sizeof(T));
    if (optval) *optval = s->GetOptval();
    if (final_iter) *final_iter = s->GetFinalIter();
    if (rho) *rho = s->GetRho();
    return 0;
  } catch (const std::exception& e) {
    return fail(e);
  }
}

// ---- unit-level hooks ------------------------------------------------------------------------
template <typename T>
__global__ void k_prox_only(size_t n, Desc<T> D, T rho, const T* __restrict__ in, T* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = prox_eval<T>(D.h[i], D.a[i], D.b[i], D.c[i], D.d[i], D.e[i], in[i], rho);
}

template <typename T>
struct DescUpload {
  DevBuf<int> h;
  DevBuf<T> a, b, c, d, e, in;
  DescUpload(size_t n, const int* hh, const T* aa, const T* bb, const T* cc, const T* dd, const T* ee,
             const T* v)
      : h(n), a(n), b(n), c(n), d(n), e(n), in(n) {
    POGS_CUDA(cudaMemcpy(h.get(), hh, n * sizeof(int), cudaMemcpyHostToDevice));
    POGS_CUDA(cudaMemcpy(a.get(), aa, n * sizeof(T), cudaMemcpyHostToDevice));
    POGS_CUDA(cudaMemcpy(b.get(), bb, n * sizeof(T), cudaMemcpyHostToDevice));
    POGS_CUDA(cudaMemcpy(c.get(), cc, n * sizeof(T), cudaMemcpyHostToDevice));
    POGS_CUDA(cudaMemcpy(d.get(), dd, n * sizeof(T), cudaMemcpyHostToDevice));
    POGS_CUDA(cudaMemcpy(e.get(), ee, n * sizeof(T), cudaMemcpyHostToDevice));
    POGS_CUDA(cudaMemcpy(in.get(), v, n * sizeof(T), cudaMemcpyHostToDevice));
  }
  Desc<T> desc() const { return Desc<T>{h.get(), a.get(), b.get(), c.get(), d.get(), e.get()}; }
};

template <typename T>
int prox_hook(size_t n, const int* h, const T* a, const T* b, const T* c, const T* d, const T* e, T rho,
              const T* in, T* out) {
  DeviceScope scope;
  try {
    require_device();
    if (n == 0) return 0;
    DescUpload<T> up(n, h, a, b, c, d, e, in);
    DevBuf<T> o(n);
    k_prox_only<T><<<(unsigned)((n + 255) / 256), 256>>>(n, up.desc(), rho, up.in.get(), o.get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaMemcpy(out, o.get(), n * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
  } catch (const std::exception& ex) {
    return fail(ex);
  }
}

template <typename T>
int func_hook(size_t n, const int* h, const T* a, const T* b, const T* c, const T* d, const T* e, const T* in,
              double* sum) {
  DeviceScope scope;
  try {
    require_device();
    *sum = 0;
    if (n == 0) return 0;
    DescUpload<T> up(n, h, a, b, c, d, e, in);
    const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + kThreads - 1) / kThreads, 1024));
    DevBuf<double> part(grid), o(1);
    // all n entries are treated as the "y" part (empty x part)
    k_objective<T><<<grid, kThreads>>>(0, n, 0u, up.desc(), up.desc(), up.in.get(), up.in.get(), part.get());
    k_fold_objective<<<1, kThreads>>>(part.get(), 0u, grid, PeerView(), o.get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaMemcpy(sum, o.get(), sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
  } catch (const std::exception& ex) {
    return fail(ex);
  }
}

template <typename T>
int gemv_hook(ORD ord, size_t m, size_t n, const T* A, int trans, int square, const T* v, T* out) {
  DeviceScope scope;
  try {
    require_device();
    cudaStream_t st;
    POGS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
      DenseMat<T> M(ord == ROW_MAJ, m, n, A, false, st);
      const size_t in_len = trans ? m : n, out_len = trans ? n : m;
      DevBuf<T> dv(in_len), dout(out_len);
      POGS_CUDA(cudaMemcpyAsync(dv.get(), v, in_len * sizeof(T), cudaMemcpyHostToDevice, st));
      EpiAffine<T> epi{T(1), T(0), nullptr, dout.get()};
      if (!trans) { if (square) M.template mul_n<true>(dv.get(), epi, nullptr); else M.template mul_n<false>(dv.get(), epi, nullptr); }
      else        { if (square) M.template mul_t<true>(dv.get(), epi, nullptr); else M.template mul_t<false>(dv.get(), epi, nullptr); }
      POGS_CUDA(cudaMemcpyAsync(out, dout.get(), out_len * sizeof(T), cudaMemcpyDeviceToHost, st));
      POGS_CUDA(cudaStreamSynchronize(st));
    }
    cudaStreamDestroy(st);
    return 0;
  } catch (const std::exception& ex) {
    return fail(ex);
  }
}

}  // namespace

// ================================================================================================
// Gram matrix hook: G (n x n, row-major, symmetric) = A^T A for a row-major m x n fp32 host
// array, on the tensor-core kernel (gram_tc.cuh) or, with use_tc == 0, on the CUDA-core product of
// dense_factor.cuh (lower triangle computed, then mirrored).
int gram_hook(size_t m, size_t n, const float* A, float* G, int use_tc, float* dbg_smem = nullptr,
              float* dbg_acc = nullptr) {
  DeviceScope scope;
  try {
    require_device();
    if (m == 0 || n == 0) throw Error("empty matrix");
    const DeviceInfo dev = query_device();
    cudaStream_t st;
    POGS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
      const size_t ld = round_up(n, 4);
      DevBuf<float> dA, dG(n * n);
      dA.alloc(m * ld, kGramSlackFloats);
      POGS_CUDA(cudaMemcpy2DAsync(dA.get(), ld * sizeof(float), A, n * sizeof(float), n * sizeof(float), m,
                                  cudaMemcpyHostToDevice, st));
      DevBuf<float> d_smem(kGramStageBytes / 4), d_acc(static_cast<size_t>(kGramBM) * kGramBN);
      if (use_tc) {
        gram_tf32x3(st, dA.get(), m, n, ld, dG.get(), n, dev.sm_count, dbg_smem != nullptr ? d_smem.get() : nullptr,
                    dbg_acc != nullptr ? d_acc.get() : nullptr);
        if (dbg_smem != nullptr) POGS_CUDA(cudaMemcpyAsync(dbg_smem, d_smem.get(), kGramStageBytes, cudaMemcpyDeviceToHost, st));
        if (dbg_acc != nullptr) POGS_CUDA(cudaMemcpyAsync(dbg_acc, d_acc.get(), sizeof(float) * kGramBM * kGramBN, cudaMemcpyDeviceToHost, st));
      } else {
        // CUDA-core product (dense_factor.cuh): the lower triangle of the row-major G, mirrored
        gemm<float, true, false>(st, static_cast<int>(n), static_cast<int>(n), static_cast<int>(m), 1.0f, dA.get(), ld, dA.get(), ld,
                                 0.0f, dG.get(), n, kTriLower);
        dim3 mg(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(n));
        k_mirror_lower<float><<<mg, 256, 0, st>>>(n, dG.get(), n);
        POGS_CUDA(cudaGetLastError());
      }
      POGS_CUDA(cudaMemcpyAsync(G, dG.get(), n * n * sizeof(float), cudaMemcpyDeviceToHost, st));
      POGS_CUDA(cudaStreamSynchronize(st));
    }
    cudaStreamDestroy(st);
    return 0;
  } catch (const std::exception& ex) {
    return fail(ex);
  }
}

extern "C" {

int PogsD(enum ORD ord, size_t m, size_t n, const double* A, const double* f_a, const double* f_b,
          const double* f_c, const double* f_d, const double* f_e, const enum FUNCTION* f_h, const double* g_a,
          const double* g_b, const double* g_c, const double* g_d, const double* g_e, const enum FUNCTION* g_h,
          double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
          int adaptive_rho, int gap_stop, double* x, double* y, double* l, double* optval,
          unsigned int* final_iter) {
  return dense_entry<double>(ord, m, n, A, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h, rho,
                             abs_tol, rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval,
                             final_iter);
}

int PogsS(enum ORD ord, size_t m, size_t n, const float* A, const float* f_a, const float* f_b, const float* f_c,
          const float* f_d, const float* f_e, const enum FUNCTION* f_h, const float* g_a, const float* g_b,
          const float* g_c, const float* g_d, const float* g_e, const enum FUNCTION* g_h, float rho,
          float abs_tol, float rel_tol, unsigned int max_iter, unsigned int verbose, int adaptive_rho,
          int gap_stop, float* x, float* y, float* l, float* optval, unsigned int* final_iter) {
  return dense_entry<float>(ord, m, n, A, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h, rho,
                            abs_tol, rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval,
                            final_iter);
}

int PogsSparseD(enum ORD ord, size_t m, size_t n, size_t nnz, const double* data, const int* ptr, const int* ind,
                const double* f_a, const double* f_b, const double* f_c, const double* f_d, const double* f_e,
                const enum FUNCTION* f_h, const double* g_a, const double* g_b, const double* g_c,
                const double* g_d, const double* g_e, const enum FUNCTION* g_h, double rho, double abs_tol,
                double rel_tol, unsigned int max_iter, unsigned int verbose, int adaptive_rho, int gap_stop,
                double* x, double* y, double* l, double* optval, unsigned int* final_iter) {
  return sparse_entry<double>(ord, m, n, nnz, data, ptr, ind, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d,
                              g_e, g_h, rho, abs_tol, rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l,
                              optval, final_iter);
}

int PogsSparseS(enum ORD ord, size_t m, size_t n, size_t nnz, const float* data, const int* ptr, const int* ind,
                const float* f_a, const float* f_b, const float* f_c, const float* f_d, const float* f_e,
                const enum FUNCTION* f_h, const float* g_a, const float* g_b, const float* g_c, const float* g_d,
                const float* g_e, const enum FUNCTION* g_h, float rho, float abs_tol, float rel_tol,
                unsigned int max_iter, unsigned int verbose, int adaptive_rho, int gap_stop, float* x, float* y,
                float* l, float* optval, unsigned int* final_iter) {
  return sparse_entry<float>(ord, m, n, nnz, data, ptr, ind, f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d,
                             g_e, g_h, rho, abs_tol, rel_tol, max_iter, verbose, adaptive_rho, gap_stop, x, y, l,
                             optval, final_iter);
}

// ---- cone form, first slice: separable cones (linear programs) on the graph-form device path ----------
}  // extern "C" (templates below)

namespace {

template <typename T>
int cone_entry(bool direct, ORD ord, size_t m, size_t n, const T* A, const T* b, const T* c,
               const ConeConstraintC* kx, size_t nkx, const ConeConstraintC* ky, size_t nky, T rho, T abs_tol, T rel_tol,
               unsigned max_iter, unsigned verbose, int adaptive_rho, int gap_stop, T* x, T* y, T* l, T* optval,
               unsigned* final_iter) {
  DeviceScope scope;
  try {
    require_device();
    // cone of every index (-1: free); == ValidCone (prox_lib_cone.h): indices in range, no index in two cones
    auto tags = [](const ConeConstraintC* K, size_t nk, size_t dim, std::vector<int>* tag) -> int {
      tag->assign(dim, -1);
      for (size_t q = 0; q < nk; ++q) {
        const int cone = static_cast<int>(K[q].cone);
        if (cone != CONE_ZERO && cone != CONE_NON_NEG && cone != CONE_NON_POS) return 2;   // not in this slice
        for (unsigned t = 0; t < K[q].size; ++t) {
          const unsigned i = K[q].indices[t];
          if (i >= dim || (*tag)[i] != -1) return 1;
          (*tag)[i] = cone;
        }
      }
      return 0;
    };
    std::vector<int> tx, ty;
    const int ex = tags(kx, nkx, n, &tx), ey = tags(ky, nky, m, &ty);
    if (ex == 1 || ey == 1) { g_last_error = "invalid cone: index out of range or in two cones"; return 5; /* POGS_INVALID_CONE */ }
    if (ex == 2 || ey == 2) throw Error("cone form: only the separable cones (zero, non-negative, non-positive) are implemented");
    // graph-form encoding (== examples/cpp/lp_eq.cpp, lp_ineq.cpp):
    //   f_i = I{b_i - y_i in K}: K = {0}: IndEq0(y - b_i); K = R+: IndLe0(y - b_i); K = R-: IndGe0(y - b_i)
    //   g_j = c_j x_j + I{x_j in K}
    const int kZeroFn = 15, kEq0 = 5, kGe0 = 6, kLe0 = 7;
    std::vector<T> fa(m, T(1)), fb(b, b + m), fc(m, T(1)), fd(m, T(0)), fe(m, T(0));
    std::vector<T> ga(n, T(1)), gb(n, T(0)), gc(n, T(1)), gd(n, T(0)), ge(n, T(0));
    std::vector<int> fh(m), gh(n);
    for (size_t i = 0; i < m; ++i) fh[i] = ty[i] == CONE_ZERO ? kEq0 : ty[i] == CONE_NON_NEG ? kLe0 : ty[i] == CONE_NON_POS ? kGe0 : kZeroFn;
    for (size_t j = 0; j < n; ++j) gh[j] = tx[j] == CONE_ZERO ? kEq0 : tx[j] == CONE_NON_NEG ? kGe0 : tx[j] == CONE_NON_POS ? kLe0 : kZeroFn;
    DenseSolver<T> solver(ord == ROW_MAJ, m, n, A, false, direct);
    // objective normalisation of the reference (PogsObjectiveCone::scale, pogs.cpp:737-754): after the
    // equilibration c o e is scaled to unit norm; optval and the dual are scaled back
    std::vector<T> dsc(m), esc(n);
    solver.GetEquil(dsc.data(), esc.data());
    double ss = 0;
    for (size_t j = 0; j < n; ++j) { const double v = static_cast<double>(c[j]) * esc[j]; ss += v * v; }
    const T c_scale = ss > 0 ? static_cast<T>(1.0 / std::sqrt(ss)) : T(1);
    for (size_t j = 0; j < n; ++j) gd[j] = c[j] * c_scale;
    set_params<T>(&solver, rho, abs_tol, rel_tol, max_iter, verbose, adaptive_rho != 0, gap_stop != 0);
    const int status = solver.Solve(fa.data(), fb.data(), fc.data(), fd.data(), fe.data(), fh.data(), ga.data(), gb.data(),
                                    gc.data(), gd.data(), ge.data(), gh.data());
    *optval = solver.GetOptval() / c_scale;
    *final_iter = solver.GetFinalIter();
    std::memcpy(x, solver.GetX(), n * sizeof(T));
    std::memcpy(y, solver.GetY(), m * sizeof(T));
    const T* lam = solver.GetLambda();
    for (size_t i = 0; i < m; ++i) l[i] = lam[i] / c_scale;
    return status;
  } catch (const std::exception& e) {
    return fail(e);
  }
}

int cone_q_unsupported() {
  g_last_error = "cone form with a quadratic objective is not implemented";
  fprintf(stderr, "pogs_b200: %s\n", g_last_error.c_str());
  return POGS_ERROR;
}

}  // namespace

extern "C" {

int PogsConeD(enum ORD ord, size_t m, size_t n, const double* A, const double* b, const double* c,
              const struct ConeConstraintC* cones_x, size_t num_cones_x, const struct ConeConstraintC* cones_y,
              size_t num_cones_y, double rho, double abs_tol, double rel_tol, unsigned int max_iter, unsigned int verbose,
              int adaptive_rho, int gap_stop, double* x, double* y, double* l, double* optval, unsigned int* final_iter) {
  return cone_entry<double>(false, ord, m, n, A, b, c, cones_x, num_cones_x, cones_y, num_cones_y, rho, abs_tol, rel_tol,
                            max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
}
int PogsConeS(enum ORD ord, size_t m, size_t n, const float* A, const float* b, const float* c,
              const struct ConeConstraintC* cones_x, size_t num_cones_x, const struct ConeConstraintC* cones_y,
              size_t num_cones_y, float rho, float abs_tol, float rel_tol, unsigned int max_iter, unsigned int verbose,
              int adaptive_rho, int gap_stop, float* x, float* y, float* l, float* optval, unsigned int* final_iter) {
  return cone_entry<float>(false, ord, m, n, A, b, c, cones_x, num_cones_x, cones_y, num_cones_y, rho, abs_tol, rel_tol,
                           max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
}
int PogsConeDirectD(enum ORD ord, size_t m, size_t n, const double* A, const double* b, const double* c,
                    const struct ConeConstraintC* cones_x, size_t num_cones_x, const struct ConeConstraintC* cones_y,
                    size_t num_cones_y, double rho, double abs_tol, double rel_tol, unsigned int max_iter,
                    unsigned int verbose, int adaptive_rho, int gap_stop, double* x, double* y, double* l, double* optval,
                    unsigned int* final_iter) {
  return cone_entry<double>(true, ord, m, n, A, b, c, cones_x, num_cones_x, cones_y, num_cones_y, rho, abs_tol, rel_tol,
                            max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
}
int PogsConeDirectS(enum ORD ord, size_t m, size_t n, const float* A, const float* b, const float* c,
                    const struct ConeConstraintC* cones_x, size_t num_cones_x, const struct ConeConstraintC* cones_y,
                    size_t num_cones_y, float rho, float abs_tol, float rel_tol, unsigned int max_iter,
                    unsigned int verbose, int adaptive_rho, int gap_stop, float* x, float* y, float* l, float* optval,
                    unsigned int* final_iter) {
  return cone_entry<float>(true, ord, m, n, A, b, c, cones_x, num_cones_x, cones_y, num_cones_y, rho, abs_tol, rel_tol,
                           max_iter, verbose, adaptive_rho, gap_stop, x, y, l, optval, final_iter);
}
int PogsConeQD(enum ORD, size_t, size_t, const double*, const double*, const double*, const double*,
               const struct ConeConstraintC*, size_t, const struct ConeConstraintC*, size_t, double, double, double,
               unsigned int, unsigned int, int, int, double*, double*, double*, double*, unsigned int*) {
  return cone_q_unsupported();
}
int PogsConeDirectQD(enum ORD, size_t, size_t, const double*, const double*, const double*, const double*,
                     const struct ConeConstraintC*, size_t, const struct ConeConstraintC*, size_t, double, double, double,
                     unsigned int, unsigned int, int, int, double*, double*, double*, double*, unsigned int*) {
  return cone_q_unsupported();
}

// ---- handle API ---------------------------------------------------------------------------------
pogs_b200_handle* pogs_b200_create_dense_s(enum ORD ord, size_t m, size_t n, const float* A, int a_on_device) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->s = new DenseSolver<float>(ord == ROW_MAJ, m, n, A, a_on_device != 0);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
pogs_b200_handle* pogs_b200_create_dense_d(enum ORD ord, size_t m, size_t n, const double* A, int a_on_device) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->is_double = 1;
    h->d = new DenseSolver<double>(ord == ROW_MAJ, m, n, A, a_on_device != 0);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
pogs_b200_handle* pogs_b200_create_dense_indirect_s(enum ORD ord, size_t m, size_t n, const float* A, int a_on_device) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->s = new DenseSolver<float>(ord == ROW_MAJ, m, n, A, a_on_device != 0, /*direct=*/false);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
pogs_b200_handle* pogs_b200_create_dense_indirect_d(enum ORD ord, size_t m, size_t n, const double* A, int a_on_device) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->is_double = 1;
    h->d = new DenseSolver<double>(ord == ROW_MAJ, m, n, A, a_on_device != 0, /*direct=*/false);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
pogs_b200_handle* pogs_b200_create_sparse_s(enum ORD ord, size_t m, size_t n, size_t nnz, const float* data,
                                            const int* ptr, const int* ind) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->s = new SparseSolver<float>(ord == ROW_MAJ, m, n, nnz, data, ptr, ind);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
pogs_b200_handle* pogs_b200_create_sparse_d(enum ORD ord, size_t m, size_t n, size_t nnz, const double* data,
                                            const int* ptr, const int* ind) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->is_double = 1;
    h->d = new SparseSolver<double>(ord == ROW_MAJ, m, n, nnz, data, ptr, ind);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}

// ---- row-block multi-GPU ---------------------------------------------------------------------
pogs_b200_comm* pogs_b200_comm_create(int rank, int world, size_t slot_bytes) {
  DeviceScope scope;
  try {
    require_device();
    pogs_b200_comm* c = new pogs_b200_comm();
    c->device = scope.dev;
    c->c = new PeerComm(rank, world, slot_bytes);
    return c;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
int pogs_b200_comm_handle(pogs_b200_comm* c, void* out64) {
  try {
    if (c == nullptr || c->c == nullptr) throw Error("null communicator");
    c->c->get_handle(out64);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
int pogs_b200_comm_open(pogs_b200_comm* c, const void* handles) {
  DeviceScope scope(c != nullptr ? c->device : -1);
  try {
    if (c == nullptr || c->c == nullptr) throw Error("null communicator");
    c->c->open_peers(handles);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
void pogs_b200_comm_destroy(pogs_b200_comm* c) {
  DeviceScope scope(c != nullptr ? c->device : -1);
  delete c;
}
int pogs_b200_comm_allreduce_s(pogs_b200_comm* c, float* dev_buf, size_t len) {
  DeviceScope scope(c != nullptr ? c->device : -1);
  try {
    if (c == nullptr || c->c == nullptr) throw Error("null communicator");
    c->c->allreduce<float>(dev_buf, len, 0);
    POGS_CUDA(cudaStreamSynchronize(0));
    if (c->c->error_raised()) throw Error("peer exchange timed out");
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
int pogs_b200_comm_allreduce_d(pogs_b200_comm* c, double* dev_buf, size_t len) {
  DeviceScope scope(c != nullptr ? c->device : -1);
  try {
    if (c == nullptr || c->c == nullptr) throw Error("null communicator");
    c->c->allreduce<double>(dev_buf, len, 0);
    POGS_CUDA(cudaStreamSynchronize(0));
    if (c->c->error_raised()) throw Error("peer exchange timed out");
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
pogs_b200_handle* pogs_b200_create_dense_rowblock_s(size_t m_local, size_t n, size_t m_global, const float* A_local,
                                                    int a_on_device, pogs_b200_comm* comm) {
  DeviceScope scope;
  try {
    require_device();
    if (comm == nullptr || comm->c == nullptr) throw Error("null communicator");
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->s = new DenseSolver<float>(true, m_local, n, A_local, a_on_device != 0, true, m_global, comm->c);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
pogs_b200_handle* pogs_b200_create_dense_rowblock_d(size_t m_local, size_t n, size_t m_global, const double* A_local,
                                                    int a_on_device, pogs_b200_comm* comm) {
  DeviceScope scope;
  try {
    require_device();
    if (comm == nullptr || comm->c == nullptr) throw Error("null communicator");
    pogs_b200_handle* h = new pogs_b200_handle();
    h->device = scope.dev;
    h->is_double = 1;
    h->d = new DenseSolver<double>(true, m_local, n, A_local, a_on_device != 0, true, m_global, comm->c);
    return h;
  } catch (const std::exception& e) { fail(e); return nullptr; }
}

void pogs_b200_destroy(pogs_b200_handle* h) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  delete h;
}

int pogs_b200_set_params(pogs_b200_handle* h, double rho, double abs_tol, double rel_tol, unsigned int max_iter,
                         unsigned int verbose, int adaptive_rho, int gap_stop) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (h == nullptr) throw Error("null handle");
    if (h->is_double) set_params<double>(impl<double>(h), rho, abs_tol, rel_tol, max_iter, verbose,
                                         adaptive_rho != 0, gap_stop != 0);
    else set_params<float>(impl<float>(h), (float)rho, (float)abs_tol, (float)rel_tol, max_iter, verbose,
                           adaptive_rho != 0, gap_stop != 0);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_set_rho(pogs_b200_handle* h, double rho) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (h == nullptr) throw Error("null handle");
    if (h->is_double) impl<double>(h)->SetRho(rho); else impl<float>(h)->SetRho((float)rho);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_set_init_s(pogs_b200_handle* h, const float* x, const float* lambda) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (x) impl<float>(h)->SetInitX(x);
    if (lambda) impl<float>(h)->SetInitLambda(lambda);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
int pogs_b200_set_init_d(pogs_b200_handle* h, const double* x, const double* lambda) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (x) impl<double>(h)->SetInitX(x);
    if (lambda) impl<double>(h)->SetInitLambda(lambda);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_set_profile(pogs_b200_handle* h, int on) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (h == nullptr) throw Error("null handle");
    if (h->is_double) impl<double>(h)->SetProfile(on != 0); else impl<float>(h)->SetProfile(on != 0);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_solve_s(pogs_b200_handle* h, const float* f_a, const float* f_b, const float* f_c, const float* f_d,
                      const float* f_e, const int* f_h, const float* g_a, const float* g_b, const float* g_c,
                      const float* g_d, const float* g_e, const int* g_h) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    return impl<float>(h)->Solve(f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h);
  } catch (const std::exception& e) { return fail(e); }
}
int pogs_b200_solve_d(pogs_b200_handle* h, const double* f_a, const double* f_b, const double* f_c,
                      const double* f_d, const double* f_e, const int* f_h, const double* g_a, const double* g_b,
                      const double* g_c, const double* g_d, const double* g_e, const int* g_h) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    return impl<double>(h)->Solve(f_a, f_b, f_c, f_d, f_e, f_h, g_a, g_b, g_c, g_d, g_e, g_h);
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_get_solution_s(pogs_b200_handle* h, float* x, float* y, float* lambda, float* mu, float* optval,
                             unsigned int* final_iter, float* rho) {
  return get_solution<float>(h, x, y, lambda, mu, optval, final_iter, rho);
}
int pogs_b200_get_solution_d(pogs_b200_handle* h, double* x, double* y, double* lambda, double* mu,
                             double* optval, unsigned int* final_iter, double* rho) {
  return get_solution<double>(h, x, y, lambda, mu, optval, final_iter, rho);
}

int pogs_b200_get_timing(pogs_b200_handle* h, double out[16]) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (h == nullptr) throw Error("null handle");
    const Timing& t = h->is_double ? impl<double>(h)->GetTiming() : impl<float>(h)->GetTiming();
    for (int i = 0; i < 16; ++i) out[i] = 0;
    out[0] = t.h2d_ms; out[1] = t.setup_ms; out[2] = t.loop_ms; out[3] = t.total_ms;
    out[4] = t.iterations; out[5] = t.exact_iterations;
    out[6] = t.prox_ms; out[7] = t.gemvt_ms; out[8] = t.solve_ms; out[9] = t.gemv_ms; out[10] = t.ctrl_ms;
    out[11] = t.profiled_iterations; out[12] = static_cast<double>(t.cgls_iterations);
    out[13] = t.equil_ms; out[14] = t.normest_ms; out[15] = t.gram_ms;
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_get_stats(pogs_b200_handle* h, double out[8]) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (h == nullptr) throw Error("null handle");
    const Timing& t = h->is_double ? impl<double>(h)->GetTiming() : impl<float>(h)->GetTiming();
    for (int i = 0; i < 8; ++i) out[i] = 0;
    out[0] = t.spec_hits; out[1] = t.normest_iterations; out[2] = t.factor_ms; out[3] = t.rare_paths; out[4] = t.one_launch;
    out[5] = t.pred_hits;
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

int pogs_b200_get_pass_phases(pogs_b200_handle* h, double out[16]) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    if (h == nullptr) throw Error("null handle");
    const Timing& t = h->is_double ? impl<double>(h)->GetTiming() : impl<float>(h)->GetTiming();
    for (int i = 0; i < 16; ++i) out[i] = t.pass_phase_us[i];
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

void pogs_b200_trim_memory(void) {
  DeviceScope scope;
  cudaDeviceSynchronize();
  mem_pool().trim();
}

const char* pogs_b200_last_error(void) { return g_last_error.c_str(); }
unsigned long long pogs_b200_launch_count(void) { return launch_counter().load(); }

// ---- test hooks ------------------------------------------------------------------------------------
int pogs_b200_plan_sparse_tiles(size_t rows, size_t cols, size_t nnz, unsigned sms, size_t elem, unsigned long long out[8]) {
  const size_t ring = elem == 8 ? tl_ring_bytes<double>(kTlWarpsDef, kTlChunkDef, kTlStagesDef)
                                : tl_ring_bytes<float>(kTlWarpsDef, kTlChunkDef, kTlStagesDef);
  TiledShape sh;
  for (int i = 0; i < 8; ++i) out[i] = 0;
  if (!plan_tiled(rows, cols, nnz, sms, elem, ring, &sh)) return 1;
  out[0] = sh.P; out[1] = sh.Q; out[2] = sh.tr; out[3] = sh.tc; out[4] = sh.ns; out[5] = sh.ntiles;
  out[6] = ring + static_cast<size_t>(sh.tc) * elem; out[7] = kTlSmemBytes;
  return 0;
}

int pogs_b200_prox_eval_s(size_t n, const int* h, const float* a, const float* b, const float* c, const float* d,
                          const float* e, float rho, const float* in, float* out) {
  return prox_hook<float>(n, h, a, b, c, d, e, rho, in, out);
}
int pogs_b200_prox_eval_d(size_t n, const int* h, const double* a, const double* b, const double* c,
                          const double* d, const double* e, double rho, const double* in, double* out) {
  return prox_hook<double>(n, h, a, b, c, d, e, rho, in, out);
}
int pogs_b200_func_eval_s(size_t n, const int* h, const float* a, const float* b, const float* c, const float* d,
                          const float* e, const float* in, double* sum) {
  return func_hook<float>(n, h, a, b, c, d, e, in, sum);
}
int pogs_b200_func_eval_d(size_t n, const int* h, const double* a, const double* b, const double* c,
                          const double* d, const double* e, const double* in, double* sum) {
  return func_hook<double>(n, h, a, b, c, d, e, in, sum);
}
int pogs_b200_gemv_s(enum ORD ord, size_t m, size_t n, const float* A, int trans, int square, const float* v,
                     float* out) {
  return gemv_hook<float>(ord, m, n, A, trans, square, v, out);
}
int pogs_b200_gemv_d(enum ORD ord, size_t m, size_t n, const double* A, int trans, int square, const double* v,
                     double* out) {
  return gemv_hook<double>(ord, m, n, A, trans, square, v, out);
}

int pogs_b200_gram_s(size_t m, size_t n, const float* A, float* G, int use_tc) { return gram_hook(m, n, A, G, use_tc); }
int pogs_b200_gram_debug_s(size_t m, size_t n, const float* A, float* G, float* stage0, float* acc0) {
  return gram_hook(m, n, A, G, 1, stage0, acc0);
}

int pogs_b200_get_equil_s(pogs_b200_handle* h, float* d, float* e, float* nrmA) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    impl<float>(h)->GetEquil(d, e);
    if (nrmA) *nrmA = impl<float>(h)->GetNormA();
    return 0;
  } catch (const std::exception& ex) { return fail(ex); }
}
int pogs_b200_get_equil_d(pogs_b200_handle* h, double* d, double* e, double* nrmA) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try {
    impl<double>(h)->GetEquil(d, e);
    if (nrmA) *nrmA = impl<double>(h)->GetNormA();
    return 0;
  } catch (const std::exception& ex) { return fail(ex); }
}
int pogs_b200_project_s(pogs_b200_handle* h, const float* x0, const float* y0, float* x, float* y) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try { impl<float>(h)->Project(x0, y0, x, y); return 0; } catch (const std::exception& ex) { return fail(ex); }
}
int pogs_b200_project_d(pogs_b200_handle* h, const double* x0, const double* y0, double* x, double* y) {
  DeviceScope scope(h != nullptr ? h->device : -1);
  try { impl<double>(h)->Project(x0, y0, x, y); return 0; } catch (const std::exception& ex) { return fail(ex); }
}

}  // extern "C"
