// Graph-form ADMM solver resident on one B200, generic over the operator
// (DenseMat / SparseMat) and the projector (direct: cached inverse; indirect: CGLS).
//
// Drop-in for pogs::PogsDirect<T, MatrixDense<T>> / pogs::PogsIndirect<T, M> of the reference
// (/root/reference/src/include/pogs.h:55-131, src/cpu/pogs.cpp:31-637,
// src/cpu/projector/projector_direct_dense.cpp): same lazy setup on the first
// solve (equilibrate, norm estimate, Gram matrix), same cached factor, same
// persistent (z, z~, rho) between solves, same stopping rule and adaptive-rho
// schedule -- but the whole iteration lives on the device.  Three forms of the loop:
//
//   one-launch iteration (dense, row-major, m > n, rows >= 6 KB; the BASELINE shapes): k_admm_pass
//       (admm_pass.cuh) runs a committed iteration in one launch; the captured graph is a round of gated
//       service kernels (exact-residual pass, k_prox, k_colacc, factor apply) followed by 16 launches of it;
//   two-pass iteration (short rows, wide or column-major matrices):
//       k_prox                      prox_f, prox_g, over-relaxation, 5 reductions
//       k_colacc(A^, t_y) [+t_x]    u  = t_x + A^T t_y      (one pass over A)
//       k_rowdot / k_symv (M, u)    x  = (I + A^T A)^-1 u   (cached explicit inverse, symmetric; replaces the
//                                   two dependent TRSVs) + x half of the dual update and residual norms
//       k_rowdot(A^, x)             y  = A x                (second pass over A) + y half-step, controller
//       [exact residuals in the body of a graph IF node]
//   CGLS iteration (sparse matrices, dense-indirect): the inner loop is the body of a graph WHILE node.
//
// The loop is captured once into a CUDA graph and replayed; the host never synchronises inside the loop,
// it only watches a progress word in mapped host memory to know when to stop feeding.
#pragma once

#include <chrono>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>
#include <memory>
#include <thread>

#include <functional>
#include <type_traits>

#include "dense_factor.cuh"
#include "dense_mat.cuh"
#include "fused_pass.cuh"
#include "gram_tc.cuh"
#include "peer_comm_host.cuh"
#include "sparse_mat.cuh"

namespace pogs_b200 {

enum Status { kSuccess = 0, kInfeasible = 1, kUnbounded = 2, kMaxIter = 3, kNanFound = 4, kInvalidCone = 5,
              kError = 6 };   // == PogsStatus, src/include/pogs.h:31-37

struct Timing {
  double setup_ms = 0;      // equilibration + norm estimate + Gram + factor (first solve only)
  double loop_ms = 0;       // ADMM iterations (device time, CUDA events)
  double total_ms = 0;      // whole solve() call, host wall clock
  double h2d_ms = 0;        // upload of A (constructor)
  // per-phase device time, filled in profile mode only
  double prox_ms = 0, gemvt_ms = 0, solve_ms = 0, gemv_ms = 0, ctrl_ms = 0;
  unsigned iterations = 0, exact_iterations = 0, profiled_iterations = 0;
  unsigned long long cgls_iterations = 0;   // CGLS inner iterations (indirect projector)
  double equil_ms = 0, normest_ms = 0, gram_ms = 0, factor_ms = 0;   // parts of setup_ms
  unsigned normest_iterations = 0;
  unsigned spec_hits = 0;   // iterations that ran on one pass over A (committed speculation)
  unsigned rare_paths = 0;  // one-launch iteration: times the rare path (two-pass kernels / standalone factor apply) ran
  unsigned one_launch = 0;  // 1: the iterations ran on the one-launch kernel (admm_pass.cuh)
  unsigned pred_hits = 0;   // committed speculations across a rho change (rho-action prediction)
  double pass_phase_us[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // mean time of the phases of k_admm_pass on CTA 0 (POGS_B200_PASS_TIMING=1)
};

// Precision-specific interface the C ABI talks to (dense-direct, dense-CGLS and
// sparse-CGLS solvers all implement it).
template <typename T>
class SolverBase {
 public:
  virtual ~SolverBase() {}
  virtual void SetRho(T v) = 0;
  virtual void SetAbsTol(T v) = 0;
  virtual void SetRelTol(T v) = 0;
  virtual void SetMaxIter(unsigned v) = 0;
  virtual void SetVerbose(unsigned v) = 0;
  virtual void SetAdaptiveRho(bool v) = 0;
  virtual void SetGapStop(bool v) = 0;
  virtual void SetInitX(const T* x) = 0;
  virtual void SetInitLambda(const T* l) = 0;
  virtual void SetProfile(bool v) = 0;
  virtual const T* GetX() const = 0;
  virtual const T* GetY() const = 0;
  virtual const T* GetLambda() const = 0;
  virtual const T* GetMu() const = 0;
  virtual T GetOptval() const = 0;
  virtual unsigned GetFinalIter() const = 0;
  virtual T GetRho() const = 0;
  virtual T GetNormA() const = 0;
  virtual const Timing& GetTiming() const = 0;
  virtual size_t Rows() const = 0;
  virtual size_t Cols() const = 0;
  virtual void Setup() = 0;
  virtual void GetEquil(T* d, T* e) = 0;
  virtual void Project(const T* x0, const T* y0, T* x, T* y) = 0;
  virtual int Solve(const T* f_a, const T* f_b, const T* f_c, const T* f_d, const T* f_e, const int* f_h,
                    const T* g_a, const T* g_b, const T* g_c, const T* g_d, const T* g_e, const int* g_h) = 0;
};

template <typename T, typename Mat>
class GraphSolver : public SolverBase<T> {
 public:
  // make_mat builds the operator on the solver's stream (uploads the matrix).
  // Row-block multi-GPU: m = local rows, m_global = rows of the whole matrix, comm = the
  // NVLink peer communicator shared by the ranks (null / m_global = 0 on a single GPU).
  GraphSolver(size_t m, size_t n, bool direct, const std::function<Mat*(cudaStream_t)>& make_mat,
              size_t m_global = 0, PeerComm* comm = nullptr)
      : m_(m), n_(n), mg_(m_global ? m_global : m), tall_(mg_ > n), kdim_(mg_ > n ? n : mg_), direct_(direct),
        comm_(comm) {
    if (m == 0 || n == 0) throw Error("empty matrix");
    if (direct && !Mat::kDense) throw Error("the direct projector needs a dense matrix");
    if (comm_ != nullptr && comm_->world() > 1) {
      pv_ = comm_->view();
      if (!direct || !tall_) throw Error("row-block multi-GPU supports the direct projector with m > n only");
    }
    POGS_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    trace_.mark("context + stream");
    auto t0 = std::chrono::steady_clock::now();
    A_.reset(make_mat(stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    timing_.h2d_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    trace_.mark("alloc + upload A");
    dev_ = A_->device();
    d_.alloc(m); e_.alloc(n);
    for (int p = 0; p < 2; ++p) { x_[p].alloc(n); y_[p].alloc(m); xt_[p].alloc(n); yt_[p].alloc(m); }
    for (int p = 0; p < 2; ++p) {
      x12_[p].alloc(n); y12_[p].alloc(m); tx_[p].alloc(n); ty_[p].alloc(m); qx_[p].alloc(n); qy_[p].alloc(m);
    }
    u_.alloc(kdim_); aux_.alloc(kdim_);
    gh_.alloc(n); ga_.alloc(n); gb_.alloc(n); gc_.alloc(n); gd_.alloc(n); ge_.alloc(n);
    fh_.alloc(m); fa_.alloc(m); fb_.alloc(m); fc_.alloc(m); fd_.alloc(m); fe_.alloc(m);
    stage_.alloc(4 * (m > n ? m : n));
    xo_.alloc(n); yo_.alloc(m); muo_.alloc(n); lo_.alloc(m);
    ctrl_.alloc(1);
    solve_ticket_.alloc(1); tail_ticket_.alloc(1);
    const size_t cap = static_cast<size_t>(dev_.sm_count) * 4;
    prox_gx_ = static_cast<unsigned>(std::min<size_t>((n + kThreads - 1) / kThreads, cap));
    prox_gy_ = static_cast<unsigned>(std::min<size_t>((m + kThreads - 1) / kThreads, cap));
    prox_grid_ = prox_gx_ + prox_gy_;
    prox_part_.alloc(static_cast<size_t>(prox_grid_) * 3);
    const unsigned nbmax = std::max(std::max(A_->nb_max(), plan_rowdot(kdim_, dev_.sm_count, kPlanOcc, kdim_).grid),
                                    2u * static_cast<unsigned>(dev_.sm_count));
    xs_part_.alloc(static_cast<size_t>(nbmax) * 2);
    ys_part_.alloc(static_cast<size_t>(nbmax) * 2);
    er_part_.alloc(nbmax); es_part_.alloc(nbmax); misc_part_.alloc(std::max(nbmax, prox_grid_));
    obj_.alloc(1);
    plan_fused();
    if (!direct_) {
      dx_.alloc(n); s_.alloc(n); p_.alloc(n); r_.alloc(m); q_.alloc(m);
      cgls_.alloc(1);
      cg_dx_part_.alloc(prox_grid_); cg_p_part_.alloc(prox_grid_);
      cg_s_part_.alloc(nbmax); cg_q_part_.alloc(nbmax);
    }
    void* hp = nullptr;
    POGS_CUDA(cudaHostAlloc(&hp, 4 * sizeof(unsigned), cudaHostAllocMapped));
    host_prog_ = static_cast<volatile unsigned*>(hp);
    host_prog_[0] = host_prog_[1] = host_prog_[2] = host_prog_[3] = 0;
    void* dp = nullptr;
    POGS_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
    dev_prog_ = static_cast<unsigned*>(dp);
    x_out_.assign(n, T(0)); y_out_.assign(m, T(0)); mu_out_.assign(n, T(0)); lambda_out_.assign(m, T(0));
    const char* nsh = getenv("POGS_B200_NO_SHARD");
    shard_solve_ = !(nsh != nullptr && nsh[0] == '1');
    const char* pt = getenv("POGS_B200_PASS_TIMING");
    pass_timing_ = pt != nullptr && pt[0] == '1';
    const char* eo = getenv("POGS_B200_EXACT_ONE_PASS");
    exact_one_pass_ = !(eo != nullptr && eo[0] == '0');
    const char* pd = getenv("POGS_B200_PDL");
    use_pdl_ = !(pd != nullptr && pd[0] == '0');
    const char* ng = getenv("POGS_B200_NO_GRAPH");
    use_graph_ = !(ng != nullptr && ng[0] == '1');
    const char* yr = getenv("POGS_B200_Y_REC");
    y_recurrence_ = !(yr != nullptr && yr[0] == '0');
    trace_.mark("state buffers");
  }

  ~GraphSolver() override {
    trace_.mark("(idle until destructor)");
    // the buffers go back to the pool in legacy-stream order: nothing of ours may still be running
    if (stream_ != nullptr) cudaStreamSynchronize(stream_);
    if (body_stream_ != nullptr) cudaStreamSynchronize(body_stream_);
    if (graph_exec_ != nullptr) cudaGraphExecDestroy(graph_exec_);
    if (body_stream_ != nullptr) cudaStreamDestroy(body_stream_);
    if (host_prog_ != nullptr) cudaFreeHost(const_cast<unsigned*>(host_prog_));
    for (cudaEvent_t ev : events_) cudaEventDestroy(ev);
    if (stream_ != nullptr) cudaStreamDestroy(stream_);
  }

  // ---- parameters (names follow pogs.h:87-119) -----------------------------------------
  void SetRho(T v) override { rho_ = v; }
  void SetAbsTol(T v) override { abs_tol_ = v; }
  void SetRelTol(T v) override { rel_tol_ = v; }
  void SetMaxIter(unsigned v) override { max_iter_ = v; }
  void SetVerbose(unsigned v) override { verbose_ = v; }
  void SetAdaptiveRho(bool v) override { adaptive_rho_ = v; }
  void SetGapStop(bool v) override { gap_stop_ = v; }
  void SetInitX(const T* x) override { init_x_.assign(x, x + n_); has_init_x_ = true; }
  void SetInitLambda(const T* l) override { init_l_.assign(l, l + m_); has_init_l_ = true; }
  void SetProfile(bool v) override { profile_ = v; }
  const T* GetX() const override { return x_out_.data(); }
  const T* GetY() const override { return y_out_.data(); }
  const T* GetLambda() const override { return lambda_out_.data(); }
  const T* GetMu() const override { return mu_out_.data(); }
  T GetOptval() const override { return optval_; }
  unsigned GetFinalIter() const override { return final_iter_; }
  T GetRho() const override { return rho_; }
  T GetNormA() const override { return nrmA_; }
  const Timing& GetTiming() const override { return timing_; }
  size_t Rows() const override { return m_; }
  size_t Cols() const override { return n_; }

  // ---- setup: _Init of the reference (pogs.cpp:59-88) + the factor that the
  //      reference builds inside its first Project (projector_direct_dense.cpp:116-121)
  void Setup() override {
    if (done_init_) return;
    // A is equilibrated in place: a setup that failed half-way (out of memory, a factorisation
    // that broke down) leaves a matrix that must not be equilibrated a second time
    if (setup_failed_) throw Error("this solver's setup failed earlier; create a new one");
    setup_failed_ = true;   // cleared at the end
    cudaEvent_t e0 = event(), e1 = event(), e2 = event(), e3 = event();
    POGS_CUDA(cudaEventRecord(e0, stream_));
    A_->equilibrate(d_.get(), e_.get());
    POGS_CUDA(cudaEventRecord(e1, stream_));
    trace_.mark("equilibrate", stream_);
    nrmA_ = A_->norm2est(ctrl_.get());
    timing_.normest_iterations = A_->normest_iters();
    POGS_CUDA(cudaEventRecord(e2, stream_));
    trace_.mark("norm estimate", stream_);
    if (direct_) build_inverse();
    POGS_CUDA(cudaEventRecord(e3, stream_));
    POGS_CUDA(cudaEventSynchronize(e3));
    float ms = 0;
    POGS_CUDA(cudaEventElapsedTime(&ms, e0, e3)); timing_.setup_ms = ms;
    POGS_CUDA(cudaEventElapsedTime(&ms, e0, e1)); timing_.equil_ms = ms;
    POGS_CUDA(cudaEventElapsedTime(&ms, e1, e2)); timing_.normest_ms = ms;
    done_init_ = true;
    setup_failed_ = false;
  }

  void GetEquil(T* d, T* e) override {
    Setup();
    POGS_CUDA(cudaMemcpy(d, d_.get(), m_ * sizeof(T), cudaMemcpyDeviceToHost));
    POGS_CUDA(cudaMemcpy(e, e_.get(), n_ * sizeof(T), cudaMemcpyDeviceToHost));
  }

  // Projection of (x0, y0) onto {y = A^ x} in the equilibrated space (test hook;
  // == ProjectorDirect::Project, projector_direct_dense.cpp:87-175).
  void Project(const T* x0, const T* y0, T* x, T* y) override {
    Setup();
    // the projection reads the parity-0 set: select it BEFORE the inputs are copied in (a previous
    // solve that ended on an odd iteration leaves hp_ == 1)
    tail_ok_ = false;   // projection only: no controller behind the last product
    fused_now_ = false; hp_ = 0;
    POGS_CUDA(cudaMemcpyAsync(tx_[0].get(), x0, n_ * sizeof(T), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaMemcpyAsync(ty_[0].get(), y0, m_ * sizeof(T), cudaMemcpyHostToDevice, stream_));
    if (direct_) enqueue_projection(0, Gate{nullptr, nullptr});
    else project_cgls(0, false, 1e-8);
    POGS_CUDA(cudaMemcpyAsync(x, x_[1].get(), n_ * sizeof(T), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaMemcpyAsync(y, y_[1].get(), m_ * sizeof(T), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
  }

  // ---- Solve (pogs.cpp:91-581 with the separable objective of :591-621) -----------------
  // All descriptor arrays are host pointers of length m (f) / n (g).
  int Solve(const T* f_a, const T* f_b, const T* f_c, const T* f_d, const T* f_e, const int* f_h,
            const T* g_a, const T* g_b, const T* g_c, const T* g_d, const T* g_e, const int* g_h) override {
    auto t_begin = std::chrono::steady_clock::now();
    if (max_iter_ == 0) throw Error("max_iter must be >= 1");
    if (has_init_x_ != has_init_l_) {
      // the reference hits ASSERT(false) -> exit(1) here (pogs.cpp:159-179)
      has_init_x_ = has_init_l_ = false;   // the handle stays usable
      throw Error("warm start needs both SetInitX and SetInitLambda (or neither)");
    }
    Setup();
    upload_desc(f_a, f_b, f_c, f_d, f_e, f_h, m_, d_.get(), 0, fa_, fb_, fc_, fd_, fe_, fh_);
    upload_desc(g_a, g_b, g_c, g_d, g_e, g_h, n_, e_.get(), 1, ga_, gb_, gc_, gd_, ge_, gh_);
    if (has_init_x_) apply_warm_start();
    has_init_x_ = has_init_l_ = false;
    trace_.mark("descriptors", stream_);

    // controller reset (pogs.cpp:198-251)
    Ctrl<T> hc;
    std::memset(&hc, 0, sizeof(hc));
    hc.abs_tol = abs_tol_; hc.rel_tol = rel_tol_; hc.nrmA = nrmA_;
    hc.sqrtn_atol = std::sqrt(static_cast<T>(n_)) * abs_tol_;
    hc.sqrtm_atol = std::sqrt(static_cast<T>(mg_)) * abs_tol_;
    hc.sqrtmn_atol = std::sqrt(static_cast<T>(mg_ + n_)) * abs_tol_;
    hc.max_iter = max_iter_; hc.adaptive_rho = adaptive_rho_ ? 1 : 0; hc.gap_stop = gap_stop_ ? 1 : 0;
    hc.rho = rho_; hc.delta = T(1.05); hc.xi = T(1); hc.prev_nrm_r = std::numeric_limits<T>::max();
    hc.zt_scale = T(1);
    hc.fused_enabled = fused_ok_ ? 1 : 0;
    hc.y_refresh = 1; hc.y_rec = 0;   // indirect projector: the first iteration forms y = A x by a product
    hc.spec_miss = 1;   // nothing speculated yet
    hc.need_solve = 1;  // ... and no factor apply in the tail of a previous pass
    POGS_CUDA(cudaMemcpyAsync(ctrl_.get(), &hc, sizeof(hc), cudaMemcpyHostToDevice, stream_));
    if (mega_ok_) POGS_CUDA(cudaMemsetAsync(phase_ns_.get(), 0, 16 * sizeof(unsigned long long), stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    host_prog_[0] = 0; host_prog_[1] = 0; host_prog_[2] = 0;
    graph_used_ = false;

    if (verbose_ > 0) print_banner();
    cudaEvent_t e0 = event(), e1 = event();
    POGS_CUDA(cudaEventRecord(e0, stream_));
    bool cgls_graph = false;
    if (!direct_) {
      POGS_CUDA(cudaMemsetAsync(cgls_.get(), 0, sizeof(CglsState), stream_));
      // whole iteration in one graph, inner CGLS loop in a WHILE node; host-driven loop for the
      // verbose tables or when conditional nodes are unavailable
      cgls_graph = verbose_ <= 1 && use_graph_ && cgls_graph_ok_ && !profile_;
      if (cgls_graph) {
        try {
          build_graph();
        } catch (const Error& e) {
          if (verbose_ > 0) fprintf(stderr, "pogs_b200: captured CGLS loop unavailable (%s); host-driven loop\n", e.what());
          cgls_graph_ok_ = false;
          cgls_graph = false;
        }
      }
      if (cgls_graph) run_loop_async(); else run_loop_stepwise();
    } else if (verbose_ > 1) {
      run_loop_stepwise();
    } else {
      run_loop_async();
    }
    POGS_CUDA(cudaEventRecord(e1, stream_));
    POGS_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    POGS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    timing_.loop_ms = ms;
    trace_.mark("loop (incl. graph build)", stream_);

    // results
    POGS_CUDA(cudaMemcpy(&hc, ctrl_.get(), sizeof(hc), cudaMemcpyDeviceToHost));
    if (!hc.done) throw Error("iteration loop ended without a decision");
    if (comm_ != nullptr && comm_->error_raised()) throw Error("peer exchange timed out (a rank died or desynchronised)");
    final_iter_ = hc.final_iter;
    rho_ = hc.rho;
    timing_.iterations = hc.final_iter + 1;
    timing_.exact_iterations = hc.exact_count;
    timing_.spec_hits = hc.spec_hits;
    timing_.rare_paths = hc.rare_count;
    timing_.pred_hits = hc.pred_hits;
    timing_.one_launch = (mega_ok_ && direct_ && tall_) ? 1u : 0u;
    if (graph_used_ && cond_active_ && !graph_rounds_)   // kernels inside IF bodies: counted when taken
      count_launch(static_cast<unsigned long long>(exact_launches()) * hc.exact_count);
    if (pass_timing_ && mega_ok_) {
      unsigned long long ns[16];
      POGS_CUDA(cudaMemcpy(ns, phase_ns_.get(), sizeof(ns), cudaMemcpyDeviceToHost));
      for (int i = 0; i < 16; ++i) timing_.pass_phase_us[i] = ns[i] * 1e-3 / std::max(1u, hc.final_iter + 1);
    }
    hp_ = static_cast<int>(hc.final_iter & 1u);
    if (!direct_) {
      CglsState cs;
      POGS_CUDA(cudaMemcpy(&cs, cgls_.get(), sizeof(cs), cudaMemcpyDeviceToHost));
      timing_.cgls_iterations = cs.total_iters;
      if (cgls_graph) count_launch(cgls_inner_launches() * cs.total_iters);   // trips of the WHILE bodies
    }
    const int p = static_cast<int>(hc.final_iter & 1u);
    optval_ = static_cast<T>(objective());
    const unsigned tb = 256;
    const size_t N = m_ + n_;
    k_outputs<T><<<(unsigned)((N + tb - 1) / tb), tb, 0, stream_>>>(
        n_, m_, d_.get(), e_.get(), x12_[hp_].get(), y12_[hp_].get(), qx_[hp_].get(), qy_[hp_].get(), ctrl_.get(), xo_.get(),
        yo_.get(), muo_.get(), lo_.get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaMemcpyAsync(x_out_.data(), xo_.get(), n_ * sizeof(T), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaMemcpyAsync(y_out_.data(), yo_.get(), m_ * sizeof(T), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaMemcpyAsync(mu_out_.data(), muo_.get(), n_ * sizeof(T), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaMemcpyAsync(lambda_out_.data(), lo_.get(), m_ * sizeof(T), cudaMemcpyDeviceToHost, stream_));
    // keep (z^k, z~^k) of the last iteration as the implicit warm start of the
    // next solve (pogs.cpp:572-573): buffers of parity 0 are the entry state.
    if (p == 1) {
      POGS_CUDA(cudaMemcpyAsync(x_[0].get(), x_[1].get(), n_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
      POGS_CUDA(cudaMemcpyAsync(y_[0].get(), y_[1].get(), m_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
    }
    k_scale_copy<T><<<(unsigned)((n_ + tb - 1) / tb), tb, 0, stream_>>>(n_, xt_[p].get(), hc.zt_scale, nullptr,
                                                                      xt_[0].get());
    k_scale_copy<T><<<(unsigned)((m_ + tb - 1) / tb), tb, 0, stream_>>>(m_, yt_[p].get(), hc.zt_scale, nullptr,
                                                                      yt_[0].get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaStreamSynchronize(stream_));

    const Status status = hc.converged ? kSuccess : kMaxIter;
    timing_.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    if (verbose_ > 0) print_summary(status, hc);
    trace_.mark("outputs", stream_);
    return status;
  }

 private:
  // ---- pieces of one iteration ---------------------------------------------------------------
  ProxArgs<T> prox_args(int p) {
    ProxArgs<T> a;
    a.n = n_; a.m = m_;
    a.g = Desc<T>{gh_.get(), ga_.get(), gb_.get(), gc_.get(), gd_.get(), ge_.get()};
    a.f = Desc<T>{fh_.get(), fa_.get(), fb_.get(), fc_.get(), fd_.get(), fe_.get()};
    a.x = x_[p].get(); a.y = y_[p].get(); a.xt = xt_[p].get(); a.yt = yt_[p].get();
    a.x12 = x12_[hp_].get(); a.y12 = y12_[hp_].get(); a.tx = tx_[hp_].get(); a.ty = ty_[hp_].get();
    a.qx = qx_[hp_].get(); a.qy = qy_[hp_].get();
    a.alpha = T(1.7);   // kAlpha, graph form (pogs.cpp:109-110)
    return a;
  }

  EpiState<T> x_state(int p, T alpha, const T* add, T* aux) {
    return EpiState<T>{alpha, add, x_[p].get(), x12_[hp_].get(), tx_[hp_].get(), x_[1 - p].get(), xt_[1 - p].get(), aux};
  }
  EpiState<T> y_state(int p, T alpha, const T* add, T* aux) {
    return EpiState<T>{alpha, add, y_[p].get(), y12_[hp_].get(), ty_[hp_].get(), y_[1 - p].get(), yt_[1 - p].get(), aux};
  }

  // (x,y) = Pi(t_x, t_y)  (projector_direct_dense.cpp:122-135) with the second
  // half-step fused into the epilogues.  Reads buffers of parity p, writes 1-p.
  void enqueue_projection(int p, Gate gate) {
    if constexpr (Mat::kDense) enqueue_projection_direct(p, gate);
  }
  void enqueue_projection_direct(int p, Gate gate) {
    const RowdotPlan mp = plan_rowdot(kdim_, dev_.sm_count, kPlanOcc, kdim_);
    if (tall_) {
      const Gate at_gate = fused_now_ ? Gate{gate.stop, &ctrl_.get()->spec_miss} : gate;
      A_->template mul_t<false>(ty_[hp_].get(), EpiAffine<T>{T(1), T(1), tx_[hp_].get(), u_.get()}, nullptr, at_gate);
      mark(1);
      if (pv_.active() && shard_solve_) {
        // sharded rows of M + fused all-gather: the replicated n^2 pass is the Amdahl term of
        // row-block scaling, so each rank applies 1/G of it
        const size_t slice = round_up((kdim_ + pv_.world - 1) / pv_.world, 32);
        const size_t row0 = std::min(kdim_, slice * pv_.rank), row1 = std::min(kdim_, row0 + slice);
        const RowdotPlan sp = plan_rowdot(row1 - row0, dev_.sm_count, kPlanOcc);
        k_solve_shard<T, 8><<<sp.grid, kThreads, 0, stream_>>>(Minv_.get(), row0, row1, kdim_, ldk_, slice, u_.get(),
                                                               x_state(p, T(1), nullptr, nullptr), xs_part_.get(),
                                                               solve_ticket_.get(), gate, pv_);
        POGS_CUDA(cudaGetLastError());
        count_launch();
        xs_nb_ = sp.grid;   // every CTA runs a share of the x half-step
      } else if (symv_ok_) {
        // lower triangle only: half the bytes of the full product (kernels.cuh, k_symv_*)
        k_symv_tiles<T><<<symv_grid_, kThreads, 0, stream_>>>(Minv_.get(), kdim_, ldk_, u_.get(), sym_tiles_.get(),
                                                            sym_ntiles_, sym_rowpart_.get(), sym_colpart_.get(), gate);
        k_symv_fold<T, EpiState<T>><<<symv_fold_grid_, kThreads, 0, stream_>>>(
            kdim_, ldk_, sym_nrb_, sym_rowpart_.get(), sym_colpart_.get(), x_state(p, T(1), nullptr, nullptr),
            xs_part_.get(), gate);
        POGS_CUDA(cudaGetLastError());
        count_launch(2);
        xs_nb_ = symv_fold_grid_;
      } else {
        launch_rowdot<T, false>(stream_, mp, Minv_.get(), kdim_, kdim_, ldk_, u_.get(),
                                x_state(p, T(1), nullptr, nullptr), xs_part_.get(), gate);
        xs_nb_ = mp.grid;
      }
      mark(2);
      ys_nb_ = A_->nb_n();
      if (fused_now_) {
        launch_fused(p, gate);
        ys_nb_ = fused_grid_;
        tail_fused_ = false;
      } else if (tail_ok_) {
        TailCtrl<T> tail{ctrl_.get(), ctrl_in(), tail_ticket_.get(), cond_switch(p)};
        tail_fused_ = A_->template mul_n_tail<false>(x_[1 - p].get(), y_state(p, T(1), nullptr, nullptr),
                                                     ys_part_.get(), gate, tail);
      } else {
        A_->template mul_n<false>(x_[1 - p].get(), y_state(p, T(1), nullptr, nullptr), ys_part_.get(), gate);
        tail_fused_ = false;
      }
      mark(3);
    } else {
      A_->template mul_n<false>(tx_[hp_].get(), EpiAffine<T>{T(1), T(-1), ty_[hp_].get(), u_.get()}, nullptr, gate);
      mark(1);
      launch_rowdot<T, false>(stream_, mp, Minv_.get(), kdim_, kdim_, ldk_, u_.get(),
                              y_state(p, T(1), ty_[hp_].get(), aux_.get()), ys_part_.get(), gate);
      mark(2);
      A_->template mul_t<false>(aux_.get(), x_state(p, T(-1), tx_[hp_].get(), nullptr), xs_part_.get(), gate);
      mark(3);
      ys_nb_ = mp.grid; xs_nb_ = A_->nb_t();
      tail_fused_ = false;
    }
  }

  CondSwitch cond_switch(int p) const {
    CondSwitch cs;
    cs.handle = cond_[p];
    cs.enabled = cond_active_ ? 1 : 0;
    return cs;
  }

  // Indirect projection (ProjectorCgls::Project, projector_cgls.cpp:52-88 around
  // cgls::Solve, cgls.h:201-323): warm start from the previous x, shift 1,
  // tolerance tied to the previous primal residual (pogs.cpp:287-290), at most
  // 500 inner iterations.  Three pieces -- start-up, one inner iteration, finish -- that are
  // either fed from the host in small batches (project_cgls: test hook, verbose tables; a
  // batch that runs past convergence is gated off on the device) or captured into the
  // iteration graph with the inner iteration as the body of a WHILE node (cgls_captured).
  void cgls_prologue(int p, bool ctrl_tol, double fixed_tol, Gate gate, CondSwitch loop) {
    CglsState* st = cgls_.get();
    const unsigned eg = prox_grid_;
    Mat& A = *A_;
    // dx = x_prev - t_x and r = t_y - A x_prev = t_y - y_prev in one vector kernel (the loop
    // invariant y_prev = A x_prev saves the reference's two start-up products)
    k_cgls_delta<T><<<eg, kThreads, 0, stream_>>>(n_, m_, x_[p].get(), tx_[hp_].get(), dx_.get(), ty_[hp_].get(),
                                                  y_[p].get(), r_.get(), cg_dx_part_.get(), gate);
    // s = A^T r - dx, and the first search direction p = s written by the same epilogue (cgls.h:248-251)
    A.template mul_t<false>(r_.get(), EpiAffine<T>{T(1), T(-1), dx_.get(), s_.get(), p_.get()}, cg_s_part_.get(), gate);
    k_cgls_start<T><<<1, kThreads, 0, stream_>>>(st, ctrl_tol ? ctrl_.get() : nullptr, fixed_tol, cg_s_part_.get(),
                                                 A.nb_t(), cg_dx_part_.get(), eg, 500u, gate, loop);
    POGS_CUDA(cudaGetLastError());
    count_launch(2);
  }
  // kernels of one trip of the inner CGLS loop: two products + four vector / scalar kernels
  unsigned cgls_inner_launches() const { return 4 + 2 * A_->launches_per_product(); }
  void cgls_inner(CondSwitch loop) {
    CglsState* st = cgls_.get();
    const unsigned eg = prox_grid_;
    Mat& A = *A_;
    const Gate run{&st->done, nullptr};
    A.template mul_n<false>(p_.get(), EpiAffine<T>{T(1), T(0), nullptr, q_.get()}, cg_q_part_.get(), run);
    k_cgls_update1<T><<<eg, kThreads, 0, stream_>>>(n_, m_, st, cg_q_part_.get(), A.nb_n(), p_.get(), q_.get(),
                                                    dx_.get(), r_.get(), cg_dx_part_.get(), run);
    A.template mul_t<false>(r_.get(), EpiAffine<T>{T(1), T(-1), dx_.get(), s_.get()}, cg_s_part_.get(), run);
    k_cgls_beta<T><<<1, kThreads, 0, stream_>>>(st, cg_s_part_.get(), A.nb_t(), cg_dx_part_.get(), eg,
                                                cg_q_part_.get(), A.nb_n(), run, loop);
    k_cgls_update2<T><<<eg, kThreads, 0, stream_>>>(n_, st, s_.get(), p_.get(), cg_p_part_.get(), run);
    k_cgls_pnorm<<<1, kThreads, 0, stream_>>>(st, cg_p_part_.get(), eg, run);
    POGS_CUDA(cudaGetLastError());
    count_launch(4);
  }
  // y = A x of the projection (projector_cgls.cpp:78).  CGLS keeps r = t_y - A x up to date by recurrence
  // (cgls.h:275: r -= alpha q), so A x = t_y - r costs no product; inside the ADMM loop (`in_loop`) the product
  // itself runs only every kYRefresh-th iteration (and in the first one) to stop the rounding of the recurrence
  // from accumulating through y_prev -- the controller keeps the two device flags that gate the variants.
  // (C5: one of the four products per ADMM iteration and its fold.)
  void cgls_epilogue(int p, Gate gate, bool in_loop) {
    const unsigned eg = prox_grid_;
    k_cgls_finish_x<T><<<eg, kThreads, 0, stream_>>>(n_, tx_[hp_].get(), dx_.get(), x_[p].get(), x12_[hp_].get(), tx_[hp_].get(),
                                                     x_[1 - p].get(), xt_[1 - p].get(), xs_part_.get(), gate);
    count_launch();
    Gate g_mul = gate, g_rec = gate;
    if (in_loop && y_recurrence_) {
      Ctrl<T>* c = ctrl_.get();
      g_mul.need = &c->y_refresh;
      g_rec.need = &c->y_rec;
      k_epi_diff<T, EpiState<T>><<<A_->nb_n(), kThreads, 0, stream_>>>(m_, ty_[hp_].get(), r_.get(),
                                                                      y_state(p, T(1), nullptr, nullptr), ys_part_.get(), g_rec);
      count_launch();
    }
    A_->template mul_n<false>(x_[1 - p].get(), y_state(p, T(1), nullptr, nullptr), ys_part_.get(), g_mul);
    POGS_CUDA(cudaGetLastError());
    xs_nb_ = eg; ys_nb_ = A_->nb_n();
  }

  void project_cgls(int p, bool ctrl_tol, double fixed_tol) {
    const Gate none{nullptr, nullptr};
    const CondSwitch off{0, 0};
    cgls_prologue(p, ctrl_tol, fixed_tol, none, off);
    int batch = 2, h_done = 0;
    for (unsigned launched = 0; launched < 500u + 8u;) {
      for (int b = 0; b < batch; ++b) cgls_inner(off);
      launched += batch;
      POGS_CUDA(cudaMemcpyAsync(&h_done, &cgls_.get()->done, sizeof(int), cudaMemcpyDeviceToHost, stream_));
      POGS_CUDA(cudaStreamSynchronize(stream_));
      if (h_done) break;
      if (batch < 8) batch *= 2;
    }
    if (!h_done) throw Error("CGLS did not terminate");
    cgls_epilogue(p, none, ctrl_tol);   // ctrl_tol: called from the ADMM loop (not from the Project() hook)
  }

  // The same projection inside the captured iteration: start-up kernels, WHILE(loop_[p]) { one
  // inner iteration }, finish.  Everything outside the body is gated on the ADMM `done` flag
  // (iterations fed past convergence must not touch the state); with the start-up gated off the
  // WHILE condition keeps its per-launch default of 0.
  void cgls_captured(int p) {
    const Gate run{&ctrl_.get()->done, nullptr};
    CondSwitch loop;
    loop.handle = loop_[p];
    loop.enabled = 1;
    cgls_prologue(p, true, 0.0, run, loop);
    capture_conditional(capture_graph_, loop_[p], cudaGraphCondTypeWhile, [&]() { cgls_inner(loop); });
    cgls_epilogue(p, run, true);
  }

  // ---- single-pass kernel: eligibility for the iteration, launch ------------------------------------
  // The launch shape belongs to the operator (DenseMat::one_pass_plan); the loop uses it when the
  // direct projector runs on a tall row-major matrix.
  void plan_fused() {
    fused_ok_ = false;
    if constexpr (Mat::kDense) {
      const char* nf = getenv("POGS_B200_NO_FUSE");
      if (nf != nullptr && nf[0] == '1') return;
      if (!direct_ || !tall_) return;
      const OnePassPlan& pl = A_->one_pass_plan();
      if (!pl.ok) return;
      fused_grid_ = pl.grid; fused_nfold_ = pl.nfold;
      for (int p = 0; p < 2; ++p) spec_part_[p].alloc(static_cast<size_t>(fused_nfold_ + fused_grid_) * 3);
      fused_ok_ = true;
      // one launch per iteration (admm_pass.cuh): controller and factor apply in the tail of the pass
      const char* mg = getenv("POGS_B200_MEGA");
      mega_ok_ = !(mg != nullptr && mg[0] == '0') && fused_nfold_ <= static_cast<unsigned>(kPassEChannel);
      if (pv_.active() && A_->ld() * sizeof(T) > pv_.push_stride) mega_ok_ = false;   // (same on every rank)
      if (mega_ok_) {
        xrow_.alloc(n_); ysum_.alloc(16); phase_ns_.alloc(16);
        if (xs_part_.size() < static_cast<size_t>(fused_nfold_) * 2) mega_ok_ = false;
      }
    }
  }

  void launch_fused(int p, Gate gate) {
    if constexpr (Mat::kDense) {
      AdmmRowOp<T> rop;
      rop.yprev = y_[p].get(); rop.y12 = y12_[p].get(); rop.ty = ty_[p].get();
      rop.ynew = y_[1 - p].get(); rop.yt_next = yt_[1 - p].get();
      rop.f = Desc<T>{fh_.get(), fa_.get(), fb_.get(), fc_.get(), fd_.get(), fe_.get()};
      rop.y12n = y12_[1 - p].get(); rop.tyn = ty_[1 - p].get(); rop.qyn = qy_[1 - p].get();
      rop.alpha = T(1.7);
      rop.ys_part = ys_part_.get(); rop.spec_part = spec_part_[1 - p].get();
      AdmmColOp<T> cop;
      cop.xnew = x_[1 - p].get(); cop.xt_next = xt_[1 - p].get();
      cop.g = Desc<T>{gh_.get(), ga_.get(), gb_.get(), gc_.get(), gd_.get(), ge_.get()};
      cop.x12n = x12_[1 - p].get(); cop.txn = tx_[1 - p].get(); cop.qxn = qx_[1 - p].get();
      cop.u_out = u_.get();
      cop.alpha = T(1.7);
      cop.spec_part = spec_part_[1 - p].get();
      A_->template one_pass<false>(x_[1 - p].get(), rop, cop, ctrl_.get(), gate);
    }
  }

  // ---- one-launch iteration (admm_pass.cuh) -----------------------------------------------------------------
  ParityArgs<T> parity_args(int p) {
    ParityArgs<T> q;
    std::memset(&q, 0, sizeof(q));
    if constexpr (Mat::kDense) {
      AdmmRowOp<T>& rop = q.rop;
      rop.yprev = y_[p].get(); rop.y12 = y12_[p].get(); rop.ty = ty_[p].get();
      rop.ynew = y_[1 - p].get(); rop.yt_next = yt_[1 - p].get();
      rop.f = Desc<T>{fh_.get(), fa_.get(), fb_.get(), fc_.get(), fd_.get(), fe_.get()};
      rop.y12n = y12_[1 - p].get(); rop.tyn = ty_[1 - p].get(); rop.qyn = qy_[1 - p].get();
      rop.alpha = T(1.7);
      rop.ys_part = ys_part_.get(); rop.spec_part = spec_part_[1 - p].get();
      AdmmColOp<T>& cop = q.cop;
      cop.xnew = x_[1 - p].get(); cop.xt_next = xt_[1 - p].get();
      cop.g = Desc<T>{gh_.get(), ga_.get(), gb_.get(), gc_.get(), gd_.get(), ge_.get()};
      cop.x12n = x12_[1 - p].get(); cop.txn = tx_[1 - p].get(); cop.qxn = qx_[1 - p].get();
      cop.u_out = u_.get();
      cop.alpha = T(1.7);
      cop.spec_part = spec_part_[1 - p].get();
      q.x = x_[1 - p].get();
      const int n1 = 1 - p;   // the tail of iteration p runs the x half-step of iteration 1-p
      q.xnext = EpiState<T>{T(1), nullptr, x_[n1].get(), x12_[n1].get(), tx_[n1].get(), x_[1 - n1].get(), xt_[1 - n1].get(), nullptr};
      q.first_x_spec = spec_part_[p].get();
      q.spec_y_next = spec_part_[1 - p].get() + static_cast<size_t>(fused_nfold_) * 3;
      q.ysum_cur = ysum_.get() + 8 * p; q.ysum_next = ysum_.get() + 8 * (1 - p);
    }
    return q;
  }

  // mode 0: the iteration in hand (its parity is read from the controller on the device); returns at once when a
  //         rare event is pending;   mode 1: only the factor apply of the iteration in hand (service path).
  void launch_mega(int mode, Gate gate, bool pdl = false) {
    if constexpr (Mat::kDense) {
      PassArgs<T> a;
      std::memset(&a, 0, sizeof(a));
      a.Mlow = Mlow_.get(); a.u = u_.get(); a.xrow = xrow_.get();
      a.xs_part = xs_part_.get();
      a.ctrl = ctrl_.get();
      a.first_x_prox = prox_part_.get(); a.prox_gx = prox_gx_;
      a.ys_part = ys_part_.get();
      a.host_progress = dev_prog_;
      a.phase_ns = pass_timing_ ? phase_ns_.get() : nullptr;
      a.mode = mode;
      A_->admm_pass(a, parity_args(0), parity_args(1), gate, pdl);
    }
  }

  // Service kernels for iteration parity p, each gated on the device: the exact-residual branch of an iteration
  // of parity p that asked for it, then -- for the iteration that follows -- the first half-step and the A^T
  // pass when the speculation was discarded (or does not exist yet) and the factor apply that the tail of the
  // previous pass did not run.  In the captured loop both parities are enqueued (parity gates); host-driven
  // loops know the parity.
  void enqueue_exact_for(int p, bool parity_gate) {
    hp_ = p;
    enqueue_exact_branch(parity_gate ? p : -1);
  }
  void enqueue_rare(int p, bool parity_gate) {
    if constexpr (Mat::kDense) {
      Ctrl<T>* c = ctrl_.get();
      hp_ = p;
      const unsigned* kp = parity_gate ? &c->k : nullptr;
      const Gate miss{&c->done, &c->spec_miss, kp, static_cast<unsigned>(p)};
      k_prox<T><<<prox_grid_, kThreads, 0, stream_>>>(prox_args(p), prox_gx_, c, prox_part_.get(), miss);
      POGS_CUDA(cudaGetLastError());
      count_launch();
      A_->template mul_t<false>(ty_[p].get(), EpiAffine<T>{T(1), T(1), tx_[p].get(), u_.get()}, nullptr, miss);
      k_ysum_first<<<1, kThreads, 0, stream_>>>(prox_part_.get(), prox_gx_, prox_gy_, ysum_.get() + 8 * p, miss, pv_);
      POGS_CUDA(cudaGetLastError());
      count_launch();
      launch_mega(1, Gate{&c->done, &c->need_solve, kp, static_cast<unsigned>(p)});
    }
  }
  void enqueue_service_done() {
    k_service_done<T><<<1, 32, 0, stream_>>>(ctrl_.get(), dev_prog_);
    POGS_CUDA(cudaGetLastError());
    count_launch();
  }
  static constexpr unsigned kIterPerRound = 16;   // launches of k_admm_pass between two runs of the service kernels

  // Host-driven form (profile mode, verbose tables, no graph): one iteration of known parity per call.
  void enqueue_iteration_mega(int p, bool with_exact) {
    Ctrl<T>* c = ctrl_.get();
    hp_ = p;
    fused_now_ = true;
    mark(-1);
    enqueue_rare(p, false);
    enqueue_service_done();
    mark(2);
    launch_mega(0, Gate{&c->done, nullptr});
    mark(3);
    xs_nb_ = fused_nfold_; ys_nb_ = fused_grid_;
    tail_fused_ = true;
    if (with_exact) enqueue_exact_for(p, false);
    mark(4);
  }

  // One round of the captured loop: the service kernels (all gated; they do nothing when no rare event is
  // pending), then a run of identical launches of the one-launch iteration kernel.  A launch that meets a rare
  // event leaves it to the next round's service kernels; the launches behind it return at once.
  void capture_round_mega() {
    Ctrl<T>* c = ctrl_.get();
    fused_now_ = true;
    for (int p = 0; p < 2; ++p) enqueue_exact_for(p, true);
    for (int p = 0; p < 2; ++p) enqueue_rare(p, true);
    enqueue_service_done();
    // launches 2..n of the run depend on their predecessor programmatically (their scheduling overlaps its tail)
    for (unsigned i = 0; i < kIterPerRound; ++i) launch_mega(0, Gate{&c->done, nullptr}, use_pdl_ && i > 0);
    xs_nb_ = fused_nfold_; ys_nb_ = fused_grid_;
  }

  CtrlIn ctrl_in() {
    CtrlIn in;
    in.prox_part = prox_part_.get(); in.prox_gx = prox_gx_; in.prox_gy = prox_gy_;
    in.spec_part = fused_now_ ? spec_part_[hp_].get() : nullptr;
    in.spec_gx = fused_nfold_; in.spec_gy = fused_grid_;
    in.xs_part = xs_part_.get(); in.xs_nb = xs_nb_;
    in.ys_part = ys_part_.get(); in.ys_nb = ys_nb_;
    in.er_part = er_part_.get(); in.er_nb = A_->nb_n();
    in.es_part = es_part_.get(); in.es_nb = A_->nb_t();
    in.host_progress = dev_prog_;
    in.pv = pv_;
    return in;
  }

  // Launches of one iteration.  With `cond_body` == false the exact-residual branch is
  // enqueued as device-gated kernels right behind the controller; when the graph uses an
  // IF node for that branch (build_graph) the caller captures enqueue_exact_branch into the
  // node's body instead.
  // kernels of the exact-residual branch: two products + the controller
  unsigned exact_launches() const { return 1 + 2 * A_->launches_per_product(); }
  void enqueue_iteration(int p, bool with_exact = true) {
    if (mega_ok_ && direct_ && tall_) { enqueue_iteration_mega(p, with_exact); return; }
    Ctrl<T>* c = ctrl_.get();
    const Gate run{&c->done, nullptr};
    hp_ = p;
    fused_now_ = fused_ok_;
    // with the single-pass kernel, the first half-step and the A^T pass only run when the
    // speculation of the previous pass was discarded (or does not exist yet)
    const Gate first_half = fused_now_ ? Gate{&c->done, &c->spec_miss} : run;
    mark(-1);
    k_prox<T><<<prox_grid_, kThreads, 0, stream_>>>(prox_args(p), prox_gx_, c, prox_part_.get(), first_half);
    POGS_CUDA(cudaGetLastError());
    count_launch();
    mark(0);
    tail_fused_ = false;
    tail_ok_ = direct_ && tall_;
    if (direct_) enqueue_projection(p, run);
    else if (capture_graph_ != nullptr) cgls_captured(p);
    else project_cgls(p, true, 0.0);
    if (!tail_fused_) {
      k_control<T><<<1, kThreads, 0, stream_>>>(c, ctrl_in(), 0, cond_switch(p));
      POGS_CUDA(cudaGetLastError());
      count_launch();
    }
    if (with_exact) enqueue_exact_branch();
    mark(4);
  }

  // exact residuals (pogs.cpp:353-376): |A^ x12 - y12| and |q_x + A^T q_y|, then phase 1
  void enqueue_exact_branch(int parity = -1) {
    Ctrl<T>* c = ctrl_.get();
    const Gate exact{&c->done, &c->need_exact, parity >= 0 ? &c->k : nullptr, static_cast<unsigned>(parity >= 0 ? parity : 0)};
    CtrlIn in = ctrl_in();
    bool one_pass = false;
    if constexpr (Mat::kDense) {
      if (fused_ok_ && exact_one_pass_) {
        // both residuals from ONE pass over A (fused_pass.cuh, ExactRowOp / ExactColOp)
        A_->template one_pass<false>(x12_[hp_].get(), ExactRowOp<T>{y12_[hp_].get(), qy_[hp_].get(), er_part_.get()},
                                     ExactColOp<T>{qx_[hp_].get(), es_part_.get()}, c, exact);
        count_launch();   // (kept equal to the two-product form: exact_launches())
        in.er_nb = fused_grid_; in.es_nb = fused_nfold_;
        one_pass = true;
      }
    }
    if (!one_pass) {
      A_->template mul_n<false>(x12_[hp_].get(), EpiAffine<T>{T(1), T(-1), y12_[hp_].get(), nullptr}, er_part_.get(), exact);
      A_->template mul_t<false>(qy_[hp_].get(), EpiAffine<T>{T(1), T(1), qx_[hp_].get(), nullptr}, es_part_.get(), exact);
    }
    k_control<T><<<1, kThreads, 0, stream_>>>(c, in, 1, CondSwitch{0, 0}, parity);
    POGS_CUDA(cudaGetLastError());
    count_launch();
  }

  // Two iterations (even / odd buffer parity) captured into one graph.  The
  // exact-residual branch of each iteration sits in the body of an IF node that the
  // controller arms from the device (cudaGraphSetConditional); if conditional nodes are
  // unavailable the branch is captured as device-gated kernels instead.
  void build_graph() {
    if (graph_exec_ != nullptr) return;
    marking_ = false;
    const char* nc = getenv("POGS_B200_NO_COND");
    bool want_cond = !(nc != nullptr && nc[0] == '1');
    for (int attempt = 0; attempt < 2 && graph_exec_ == nullptr; ++attempt) {
      const unsigned long long before = launch_counter().load();
      cudaGraph_t graph = nullptr;
      cond_active_ = want_cond;
      try {
        POGS_CUDA(cudaGraphCreate(&graph, 0));
        if (!direct_ && !cond_active_) throw Error("the captured CGLS loop needs CUDA-graph conditional nodes");
        const bool mega = mega_ok_ && direct_ && tall_;
        if (mega) cond_active_ = false;   // the one-launch loop uses gated kernels only, no conditional nodes
        if (cond_active_) {
          for (int p = 0; p < 2; ++p) {
            POGS_CUDA(cudaGraphConditionalHandleCreate(&cond_[p], graph, 0, cudaGraphCondAssignDefault));
            if (!direct_) POGS_CUDA(cudaGraphConditionalHandleCreate(&loop_[p], graph, 0, cudaGraphCondAssignDefault));
          }
        }
        POGS_CUDA(cudaStreamBeginCaptureToGraph(stream_, graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        capture_graph_ = graph;
        graph_rounds_ = mega;
        if (graph_rounds_) {
          capture_round_mega();
        } else {
          for (int p = 0; p < 2; ++p) {
            enqueue_iteration(p, /*with_exact=*/!cond_active_);
            if (cond_active_) capture_conditional(graph, cond_[p], cudaGraphCondTypeIf, [&]() { enqueue_exact_branch(); });
          }
        }
        capture_graph_ = nullptr;
        cudaGraph_t out = nullptr;
        POGS_CUDA(cudaStreamEndCapture(stream_, &out));
        graph_nodes_ = launch_counter().load() - before;
        // kernels inside the two IF bodies and the two WHILE bodies: counted when taken, not per replay
        exact_nodes_ = (cond_active_ ? exact_launches() * 2 : 0) + (direct_ ? 0 : 2 * cgls_inner_launches());
        if (graph_rounds_) exact_nodes_ = 0;   // no conditional nodes: every captured kernel is launched (most return at their gate)
        launch_counter().store(before);            // captured, not launched
        POGS_CUDA(cudaGraphInstantiate(&graph_exec_, graph, 0));
        POGS_CUDA(cudaGraphDestroy(graph));
      } catch (const Error& e) {
        capture_graph_ = nullptr;
        cudaGraph_t dummy = nullptr;
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream_, &st);
        if (st != cudaStreamCaptureStatusNone) cudaStreamEndCapture(stream_, &dummy);
        if (graph != nullptr) cudaGraphDestroy(graph);
        cudaGetLastError();
        launch_counter().store(before);
        graph_exec_ = nullptr;
        if (!want_cond || !direct_) throw;
        if (verbose_ > 0) fprintf(stderr, "pogs_b200: conditional graph nodes unavailable (%s); using gated launches\n", e.what());
        want_cond = false;
      }
    }
    if (graph_exec_ == nullptr) throw Error("could not build the iteration graph");
  }

  // Adds a conditional node (IF or WHILE on `handle`) behind what has been captured so far and
  // captures `body()` -- launches on stream_ -- into its body graph.
  void capture_conditional(cudaGraph_t graph, cudaGraphConditionalHandle handle, cudaGraphConditionalNodeType type,
                           const std::function<void()>& body_fn) {
    cudaStreamCaptureStatus st;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    cudaGraph_t g = nullptr;
    unsigned long long id = 0;
    POGS_CUDA(cudaStreamGetCaptureInfo_v2(stream_, &st, &id, &g, &deps, &ndeps));
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = type;
    cp.conditional.size = 1;
    cudaGraphNode_t cnode;
    POGS_CUDA(cudaGraphAddNode(&cnode, graph, deps, ndeps, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    if (body_stream_ == nullptr) POGS_CUDA(cudaStreamCreateWithFlags(&body_stream_, cudaStreamNonBlocking));
    cudaStream_t main = stream_;
    POGS_CUDA(cudaStreamBeginCaptureToGraph(body_stream_, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    stream_ = body_stream_;
    A_->set_stream(body_stream_);
    try {
      body_fn();
    } catch (...) {
      stream_ = main; A_->set_stream(main);
      cudaGraph_t dummy = nullptr;
      cudaStreamEndCapture(body_stream_, &dummy);
      throw;
    }
    stream_ = main;
    A_->set_stream(main);
    cudaGraph_t done_body = nullptr;
    POGS_CUDA(cudaStreamEndCapture(body_stream_, &done_body));
    POGS_CUDA(cudaStreamUpdateCaptureDependencies(stream_, &cnode, 1, cudaStreamSetCaptureDependencies));
  }

  // Feed the device two iterations at a time, at most kLookahead iterations
  // ahead of the progress word the controller writes to mapped host memory.
  void run_loop_async() {
    constexpr unsigned kLookahead = 16;
    // profile mode: plain launches with an event between the phases of every
    // iteration (recorded, not waited for), resolved after the loop
    marking_ = profile_;
    const bool graph = use_graph_ && !profile_;
    if (graph) { build_graph(); trace_.mark("graph build"); graph_used_ = true; }
    unsigned launched = 0;
    auto last_progress = std::chrono::steady_clock::now();
    unsigned last_seen = 0;
    if (graph && graph_rounds_) {
      // rounds of a device-decided number of iterations: keep a few rounds queued behind the one that runs
      unsigned rounds = 0;
      for (;;) {
        const unsigned prog = host_prog_[0];
        if (host_prog_[1] != 0) break;
        if (prog != last_seen) { last_seen = prog; last_progress = std::chrono::steady_clock::now(); }
        if (rounds - host_prog_[2] < 3 && rounds < 2 * (max_iter_ + 2)) {
          POGS_CUDA(cudaGraphLaunch(graph_exec_, stream_));
          count_launch(graph_nodes_ - exact_nodes_);
          rounds += 1;
          continue;
        }
        const cudaError_t q = cudaStreamQuery(stream_);
        if (q != cudaSuccess && q != cudaErrorNotReady) POGS_CUDA(q);
        if (q == cudaSuccess && host_prog_[1] == 0 && rounds >= 2 * (max_iter_ + 2))
          throw Error("loop drained without reaching max_iter");
        const double idle =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - last_progress).count();
        if (idle > 120.0) throw Error("no progress from the device for 120 s");
        std::this_thread::yield();
      }
      POGS_CUDA(cudaStreamSynchronize(stream_));
      marking_ = false;
      return;
    }
    for (;;) {
      const unsigned prog = host_prog_[0];
      if (host_prog_[1] != 0) break;
      if (prog != last_seen) { last_seen = prog; last_progress = std::chrono::steady_clock::now(); }
      if (launched - prog < kLookahead && launched < max_iter_ + 1) {
        if (graph) {
          POGS_CUDA(cudaGraphLaunch(graph_exec_, stream_));
          count_launch(graph_nodes_ - exact_nodes_);   // IF bodies counted only when taken (not tracked)
        } else {
          cond_active_ = false;
          enqueue_iteration(0);
          enqueue_iteration(1);
        }
        launched += 2;
        continue;
      }
      const cudaError_t q = cudaStreamQuery(stream_);
      if (q != cudaSuccess && q != cudaErrorNotReady) POGS_CUDA(q);
      if (q == cudaSuccess && host_prog_[1] == 0 && host_prog_[0] == launched && launched >= max_iter_ + 1)
        throw Error("loop drained without reaching max_iter");
      const double idle =
          std::chrono::duration<double>(std::chrono::steady_clock::now() - last_progress).count();
      if (idle > 120.0) throw Error("no progress from the device for 120 s");
      std::this_thread::yield();
    }
    POGS_CUDA(cudaStreamSynchronize(stream_));
    if (marking_) collect_marks();
    marking_ = false;
  }

  // One iteration per launch with a host sync in between: used for verbose
  // tables and by the indirect projector (whose inner CGLS loop is host-driven).
  void run_loop_stepwise() {
    marking_ = false;
    Ctrl<T> hc;
    cond_active_ = false;
    for (unsigned it = 0;; ++it) {
      enqueue_iteration(static_cast<int>(it & 1u));
      POGS_CUDA(cudaStreamSynchronize(stream_));
      POGS_CUDA(cudaMemcpy(&hc, ctrl_.get(), sizeof(hc), cudaMemcpyDeviceToHost));
      const unsigned k = hc.done ? hc.final_iter : hc.k - 1;
      if ((verbose_ > 2 && k % 10 == 0) || (verbose_ > 1 && k % 100 == 0) || (verbose_ > 1 && hc.done && hc.converged)) {
        const double ov = objective();
        printf("%5u : %.2e  %.2e  %.2e  %.2e  %.2e  %.2e % .2e\n", k, (double)hc.nrm_r, (double)hc.eps_pri,
               (double)hc.nrm_s, (double)hc.eps_dua, (double)hc.gap, (double)hc.eps_gap, ov);
      }
      if (hc.done) break;
    }
    marking_ = false;
  }

  // ---- profile marks: events between the phases of an iteration ---------------------------------
  void mark(int id) {
    if (!marking_) return;
    cudaEvent_t ev = event();
    POGS_CUDA(cudaEventRecord(ev, stream_));
    marks_.push_back({id, ev});
  }
  void collect_marks() {
    const unsigned executed = host_prog_[0];   // iterations that really ran (later ones were gated off)
    unsigned seen = 0;
    timing_.prox_ms = timing_.gemvt_ms = timing_.solve_ms = timing_.gemv_ms = timing_.ctrl_ms = 0;
    timing_.profiled_iterations = 0;
    for (size_t i = 1; i < marks_.size(); ++i) {
      const int id = marks_[i].first;
      if (id == -1) { if (++seen >= executed) break; continue; }   // boundary between iterations
      float ms = 0;
      POGS_CUDA(cudaEventElapsedTime(&ms, marks_[i - 1].second, marks_[i].second));
      const bool tall = tall_;
      if (id == 0) timing_.prox_ms += ms;
      else if (id == 1) (tall ? timing_.gemvt_ms : timing_.gemv_ms) += ms;
      else if (id == 2) timing_.solve_ms += ms;
      else if (id == 3) (tall ? timing_.gemv_ms : timing_.gemvt_ms) += ms;
      else if (id == 4) timing_.ctrl_ms += ms;
    }
    timing_.profiled_iterations = executed;
    for (auto& mk : marks_) free_events_.push_back(mk.second);
    marks_.clear();
  }
  cudaEvent_t event() {
    if (!free_events_.empty()) {
      cudaEvent_t ev = free_events_.back();
      free_events_.pop_back();
      return ev;
    }
    cudaEvent_t ev;
    POGS_CUDA(cudaEventCreate(&ev));
    events_.push_back(ev);
    return ev;
  }

  // ---- objective f(y12) + g(x12) in the scaled space (pogs.cpp:473) ------------------------------
  double objective() {
    Desc<T> g{gh_.get(), ga_.get(), gb_.get(), gc_.get(), gd_.get(), ge_.get()};
    Desc<T> f{fh_.get(), fa_.get(), fb_.get(), fc_.get(), fd_.get(), fe_.get()};
    k_objective<T><<<prox_grid_, kThreads, 0, stream_>>>(n_, m_, prox_gx_, g, f, x12_[hp_].get(), y12_[hp_].get(),
                                                         misc_part_.get());
    k_fold_objective<<<1, kThreads, 0, stream_>>>(misc_part_.get(), prox_gx_, prox_gy_, pv_, obj_.get());
    POGS_CUDA(cudaGetLastError());
    double v = 0;
    POGS_CUDA(cudaMemcpyAsync(&v, obj_.get(), sizeof(double), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    return v;
  }

  // ---- descriptors ----------------------------------------------------------------------------------
  void upload_desc(const T* a, const T* b, const T* c, const T* d, const T* e, const int* h, size_t len,
                   const T* scale, int mode, DevBuf<T>& da, DevBuf<T>& db, DevBuf<T>& dc, DevBuf<T>& dd,
                   DevBuf<T>& de, DevBuf<int>& dh) {
    T* st = stage_.get();
    const size_t q = stage_.size() / 4;
    POGS_CUDA(cudaMemcpyAsync(st, a, len * sizeof(T), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaMemcpyAsync(st + q, c, len * sizeof(T), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaMemcpyAsync(st + 2 * q, d, len * sizeof(T), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaMemcpyAsync(st + 3 * q, e, len * sizeof(T), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaMemcpyAsync(db.get(), b, len * sizeof(T), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaMemcpyAsync(dh.get(), h, len * sizeof(int), cudaMemcpyHostToDevice, stream_));
    const unsigned tb = 256;
    k_scale_desc<T><<<(unsigned)((len + tb - 1) / tb), tb, 0, stream_>>>(len, scale, mode, st, st + q, st + 2 * q,
                                                                        st + 3 * q, da.get(), dc.get(), dd.get(),
                                                                        de.get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaStreamSynchronize(stream_));   // staging buffer is reused by the next call
  }

  // Explicit warm start from (x0, lambda0) (pogs.cpp:144-156):
  //   z = [x0/e ; A^(x0/e)],  z~ = [A^T(l0/d)/rho ; -(l0/d)/rho]
  void apply_warm_start() {
    const unsigned tb = 256;
    T* st = stage_.get();
    POGS_CUDA(cudaMemcpyAsync(st, init_x_.data(), n_ * sizeof(T), cudaMemcpyHostToDevice, stream_));
    k_div<T><<<(unsigned)((n_ + tb - 1) / tb), tb, 0, stream_>>>(n_, st, e_.get(), x_[0].get());
    A_->template mul_n<false>(x_[0].get(), EpiAffine<T>{T(1), T(0), nullptr, y_[0].get()}, nullptr);
    POGS_CUDA(cudaStreamSynchronize(stream_));
    POGS_CUDA(cudaMemcpyAsync(st, init_l_.data(), m_ * sizeof(T), cudaMemcpyHostToDevice, stream_));
    k_div<T><<<(unsigned)((m_ + tb - 1) / tb), tb, 0, stream_>>>(m_, st, d_.get(), ty_[hp_].get());
    A_->template mul_t<false>(ty_[hp_].get(), EpiAffine<T>{T(1) / rho_, T(0), nullptr, xt_[0].get()}, nullptr);
    k_scale_copy<T><<<(unsigned)((m_ + tb - 1) / tb), tb, 0, stream_>>>(m_, ty_[hp_].get(), T(-1) / rho_, nullptr,
                                                                      yt_[0].get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaStreamSynchronize(stream_));
  }

  // ---- cached (I + A^T A)^-1  (or (I + A A^T)^-1 when m <= n) -----------------------------------------
  // Gram matrix on cuBLAS (one-time, compute-bound), Cholesky factor and
  // inverse in fp64 on cuSOLVER, narrowed to T.  The reference factors the same
  // matrix once (projector_direct_dense.cpp:116-121, gsl_linalg.h:37-55) and
  // applies it with two dependent triangular solves per iteration; a dense
  // symmetric product is the bandwidth-optimal way to apply it on the device
  // and is safe because the equilibrated Gram matrix is well conditioned.
  void build_inverse() {
    if constexpr (Mat::kDense) build_inverse_dense();
  }
  template <typename M = Mat>
  typename std::enable_if<M::kDense>::type build_inverse_dense() {
    const size_t k = kdim_;
    ldk_ = round_up(k, V16<T>::N);
    const size_t R = A_->R(), C = A_->C(), ld = A_->ld();
    // Gram over storage columns (C x C) = S^T S ; over storage rows (R x R) = S S^T  (S = the R x C row-major store)
    const bool over_cols = (tall_ != A_->transposed_storage());
    if (k != (over_cols ? C : R)) throw Error("internal: Gram dimension mismatch");
    DevBuf<T> G(k * k);
    cudaEvent_t g0 = event(), g1 = event(), g2 = event();
    POGS_CUDA(cudaEventRecord(g0, stream_));
    bool on_tensor_cores = false;
    const char* gsel = getenv("POGS_B200_GRAM");
    const bool want_plain = gsel != nullptr && (gsel[0] == 'c' || gsel[0] == 'p');   // "cuda-core" / "plain" (was: cublas)
    if constexpr (std::is_same<T, float>::value) {
      // row-major tall fp32 operator: hand-written tcgen05 3xTF32 kernel (gram_tc.cuh); other layouts and
      // fp64 use the CUDA-core product of dense_factor.cuh.  POGS_B200_GRAM=plain forces the latter.
      // (thresholds on global sizes: every rank of a row-block solve must take the same branch,
      // or the replicas of the factor would differ in the last bits)
      if (!want_plain && over_cols && !A_->transposed_storage() && k >= 256 && mg_ >= 256 && R >= 1) {
        gram_tf32x3(stream_, A_->data(), R, C, ld, G.get(), k, dev_.sm_count);
        on_tensor_cores = true;
      }
    }
    if (!on_tensor_cores) {
      // lower triangle of the Gram matrix (all the factorisation reads)
      if (over_cols) gemm<T, true, false>(stream_, static_cast<int>(k), static_cast<int>(k), static_cast<int>(R), T(1), A_->data(), ld,
                                          A_->data(), ld, T(0), G.get(), k, kTriLower);
      else gemm<T, false, true>(stream_, static_cast<int>(k), static_cast<int>(k), static_cast<int>(C), T(1), A_->data(), ld,
                                A_->data(), ld, T(0), G.get(), k, kTriLower);
    }
    gram_on_tensor_cores_ = on_tensor_cores;
    // row blocks: A^T A = sum over ranks of A_g^T A_g (one-time, summed in rank order so
    // that every rank factors the same bits)
    if (comm_ != nullptr) comm_->allreduce(G.get(), round_up(k * k, V16<T>::N), stream_);
    POGS_CUDA(cudaEventRecord(g1, stream_));
    trace_.mark("Gram", stream_);
    // Working precision of the factorisation.  kappa(I + G) <= 1 + |A^|_2^2, and the norm estimate
    // is already known: for fp32 data a well-conditioned system (the usual case after
    // equilibration: ~3 for the BASELINE matrices) is factored and inverted in fp32, like the
    // reference's float path does (gsl_linalg.h:37-55 on float), everything else in fp64.
    // POGS_B200_FACTOR=fp32|fp64 forces the choice.
    bool fp32_factor = false;
    if constexpr (std::is_same<T, float>::value) {
      const double kappa_bound = 1.0 + static_cast<double>(nrmA_) * static_cast<double>(nrmA_);
      fp32_factor = kappa_bound <= 64.0;
      if (const char* e = getenv("POGS_B200_FACTOR")) {
        if (strcmp(e, "fp32") == 0) fp32_factor = true;
        if (strcmp(e, "fp64") == 0) fp32_factor = false;
      }
    }
    Minv_.alloc(k * ldk_);
    if (fp32_factor) factor_and_invert<float>(G.get(), k, !want_plain);
    else factor_and_invert<double>(G.get(), k, false);
    plan_symv();
    if (mega_ok_ && tall_) {
      // packed lower triangle (diagonal halved) for the streamed factor apply of k_admm_pass
      Mlow_.alloc(sym_row_off<T>(k) + 64);
      k_pack_sym<T><<<static_cast<unsigned>(k), kThreads, 0, stream_>>>(k, Minv_.get(), ldk_, Mlow_.get());
      POGS_CUDA(cudaGetLastError());
      count_launch();
    }
    POGS_CUDA(cudaEventRecord(g2, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    float gms = 0;
    POGS_CUDA(cudaEventElapsedTime(&gms, g0, g1)); timing_.gram_ms = gms;
    POGS_CUDA(cudaEventElapsedTime(&gms, g1, g2)); timing_.factor_ms = gms;
    trace_.mark("factor + inverse", stream_);
  }

  // Tile list and partial buffers of the symmetric factor apply (tall case, single GPU or
  // unsharded): tiles (row block, column block) that touch the lower triangle.
  void plan_symv() {
    symv_ok_ = false;
    // default for n >= 4096 (C2: 60 us against 73 us for the full product; a first version with
    // 32-row tiles at 128 registers and a 40-CTA fold took 98 us); POGS_B200_SYMV=1 also enables it
    // for smaller systems (tests), =0 disables it
    const char* e = getenv("POGS_B200_SYMV");
    if (e != nullptr && e[0] == '0') return;
    const bool forced = e != nullptr && e[0] == '1';
    if (!forced && kdim_ < 4096) return;
    if (!tall_ || kdim_ < 512) return;
    const size_t tile_cols = static_cast<size_t>(kThreads) * V16<T>::N;
    const size_t ncb = (ldk_ + tile_cols - 1) / tile_cols, nrb = (kdim_ + kSymStrip - 1) / kSymStrip;
    std::vector<SymTile> tiles;
    for (size_t rb = 0; rb < nrb; ++rb)
      for (size_t cb = 0; cb < ncb && cb * tile_cols <= rb * kSymStrip + kSymStrip - 1; ++cb)
        tiles.push_back(SymTile{static_cast<int>(rb), static_cast<int>(cb)});
    symv_fold_grid_ = static_cast<unsigned>((kdim_ + 31) / 32);
    if (tiles.empty() || static_cast<size_t>(symv_fold_grid_) * 2 > xs_part_.size()) return;
    sym_ntiles_ = static_cast<unsigned>(tiles.size());
    sym_nrb_ = static_cast<unsigned>(nrb);
    symv_grid_ = std::min<unsigned>(sym_ntiles_, 3u * static_cast<unsigned>(dev_.sm_count));
    sym_tiles_.alloc(tiles.size());
    POGS_CUDA(cudaMemcpyAsync(sym_tiles_.get(), tiles.data(), tiles.size() * sizeof(SymTile), cudaMemcpyHostToDevice, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));   // `tiles` is a local
    sym_rowpart_.alloc(ncb * ldk_);
    sym_colpart_.alloc(nrb * ldk_);
    symv_ok_ = true;
  }

  // Minv_ = (G + I)^-1 in working precision W with the library's own kernels (dense_factor.cuh):
  //   G + I = L L^T (blocked right-looking Cholesky, == gsl_linalg.h:37-55),  X = L^-1 (blocked triangular
  //   inverse),  (G + I)^-1 = X^T X -- on the tensor-core Gram kernel for fp32 (the n x n x n product is 2/3 of
  //   the flops of the inversion), on the CUDA-core product otherwise.  Only the lower triangle of G is read.
  template <typename W>
  void factor_and_invert(const T* G, size_t k, bool allow_tensor_cores) {
    const size_t ldx = round_up(k, 4);
    DevBuf<W> Gw(k * k), work(factor_work_elems(k)), X;
    X.alloc(k * ldx, kGramSlackFloats);   // zero-initialised: the strict upper triangle of L^-1 stays zero
    DevBuf<int> info(1);
    dim3 grid(static_cast<unsigned>((k + 255) / 256), static_cast<unsigned>(k));
    k_widen_add_diag<T, W><<<grid, 256, 0, stream_>>>(k, G, k, Gw.get(), k, W(1));
    POGS_CUDA(cudaGetLastError());
    const int ki = static_cast<int>(k);
    trace_.mark("factor buffers", stream_);
    // ~630 short, dependent launches (diagonal block, panel, updates per 64 columns).  The host enqueues them in
    // 1.8 ms; replaying the chain from a captured graph was measured and changed nothing (50 ms either way).
    chol_lower<W>(stream_, ki, Gw.get(), k, work.get(), info.get());
    int h_info = 0;
    POGS_CUDA(cudaMemcpyAsync(&h_info, info.get(), sizeof(int), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    if (h_info != 0) throw Error("Cholesky factorisation of I + A^T A failed (info=" + std::to_string(h_info) + ")");
    trace_.mark(sizeof(W) == 4 ? "Cholesky (fp32)" : "Cholesky (fp64)", stream_);
    {
      DevBuf<W> tmp(factor_tmp_elems(k));
      tri_inverse_lower<W>(stream_, ki, Gw.get(), k, X.get(), ldx, work.get(), tmp.get());
      POGS_CUDA(cudaStreamSynchronize(stream_));   // tmp goes out of scope
    }
    trace_.mark(sizeof(W) == 4 ? "L^-1 (fp32)" : "L^-1 (fp64)", stream_);
    bool done = false;
    if constexpr (std::is_same<W, float>::value && std::is_same<T, float>::value) {
      if (allow_tensor_cores && k >= 256) {
        gram_tf32x3(stream_, X.get(), k, k, ldx, Minv_.get(), ldk_, dev_.sm_count);
        trace_.mark("X^T X (tcgen05)", stream_);
        done = true;
      }
    }
    if (!done) {
      // lower triangle of X^T X in W, mirrored and narrowed to T
      gemm<W, true, false>(stream_, ki, ki, ki, W(1), X.get(), ldx, X.get(), ldx, W(0), Gw.get(), k, kTriLower);
      dim3 grid2(static_cast<unsigned>((ldk_ + 255) / 256), static_cast<unsigned>(k));
      k_sym_cast<T, W><<<grid2, 256, 0, stream_>>>(k, Gw.get(), k, Minv_.get(), ldk_, 1);
      POGS_CUDA(cudaGetLastError());
      trace_.mark("X^T X (CUDA cores)", stream_);
    }
    POGS_CUDA(cudaStreamSynchronize(stream_));   // Gw, work, X go out of scope
  }

  // ---- console output (pogs.cpp:186-196, 485-507) -----------------------------------------------------
  static const char* hbar() {
    return "----------------------------------------------------------------------------\n";
  }
  void print_banner() {
    printf("%s           POGS-B200 - Proximal Graph Solver (sm_100a)\n", hbar());
    if (verbose_ > 1)
      printf("%s Iter | pri res | pri tol | dua res | dua tol |   gap   | eps gap | pri obj\n%s", hbar(), hbar());
  }
  void print_summary(Status st, const Ctrl<T>& hc) {
    const char* s = st == kSuccess ? "Solved" : st == kMaxIter ? "Reached max iter" : "Error";
    printf("%sStatus: %s\nTiming: Total = %3.2e s, Init = %3.2e s\nIter  : %u\n", hbar(), s,
           timing_.total_ms * 1e-3, timing_.setup_ms * 1e-3, hc.final_iter);
    printf("%sError Metrics:\nPri: |Ax - y|    / (abs_tol sqrt(m)     / rel_tol + |y|)          = %.2e\n"
           "Dua: |A'l + u|   / (abs_tol sqrt(n)     / rel_tol + |u|)          = %.2e\n"
           "Gap: |x'u + y'l| / (abs_tol sqrt(m + n) / rel_tol + |x,u| |y,l|)  = %.2e\n%s",
           hbar(), (double)(rel_tol_ * hc.nrm_r / hc.eps_pri), (double)(rel_tol_ * hc.nrm_s / hc.eps_dua),
           (double)(rel_tol_ * hc.gap / hc.eps_gap), hbar());
    fflush(stdout);
  }

  // ---- data -----------------------------------------------------------------------------------------------
  Trace trace_;
  size_t m_, n_, mg_;
  bool tall_;
  size_t kdim_, ldk_ = 0;
  cudaStream_t stream_ = nullptr;
  bool direct_;
  PeerComm* comm_ = nullptr;   // not owned
  PeerView pv_;
  std::unique_ptr<Mat> A_;
  DeviceInfo dev_;
  DevBuf<T> Minv_, d_, e_;
  DevBuf<T> x_[2], y_[2], xt_[2], yt_[2];
  // first-half-step results, one set per iteration parity: iteration k uses set k&1 while the
  // fused pass already writes the speculative set (k+1)&1
  DevBuf<T> x12_[2], y12_[2], tx_[2], ty_[2], qx_[2], qy_[2], u_, aux_;
  int hp_ = 0;
  // single-pass kernel (fused_pass.cuh)
  bool fused_ok_ = false, fused_now_ = false;
  // one-launch iteration (admm_pass.cuh)
  bool mega_ok_ = false, use_pdl_ = true, exact_one_pass_ = true, graph_used_ = false, pass_timing_ = false, graph_rounds_ = false;
  DevBuf<T> Mlow_, xrow_;
  DevBuf<double> ysum_;
  DevBuf<unsigned long long> phase_ns_;
  unsigned fused_grid_ = 0, fused_nfold_ = 0;
  DevBuf<double> spec_part_[2];
  DevBuf<int> gh_, fh_;
  DevBuf<T> ga_, gb_, gc_, gd_, ge_, fa_, fb_, fc_, fd_, fe_, stage_;
  DevBuf<T> xo_, yo_, muo_, lo_;
  DevBuf<Ctrl<T>> ctrl_;
  DevBuf<unsigned> solve_ticket_, tail_ticket_;
  cudaGraphConditionalHandle cond_[2] = {0, 0};   // IF: exact-residual branch of iteration parity p
  cudaGraphConditionalHandle loop_[2] = {0, 0};   // WHILE: inner CGLS iteration of parity p
  cudaGraph_t capture_graph_ = nullptr;           // non-null while build_graph is capturing
  bool cgls_graph_ok_ = true;
  bool gram_on_tensor_cores_ = false;
  // symmetric factor apply
  bool symv_ok_ = false;
  unsigned symv_grid_ = 0, symv_fold_grid_ = 0, sym_ntiles_ = 0, sym_nrb_ = 0;
  DevBuf<SymTile> sym_tiles_;
  DevBuf<T> sym_rowpart_, sym_colpart_;
  bool cond_active_ = false, tail_ok_ = false, tail_fused_ = false, shard_solve_ = true;
  cudaStream_t body_stream_ = nullptr;
  unsigned long long exact_nodes_ = 0;
  // indirect projector (CGLS) work space
  DevBuf<T> dx_, s_, p_, r_, q_;
  DevBuf<CglsState> cgls_;
  DevBuf<double> cg_dx_part_, cg_p_part_, cg_s_part_, cg_q_part_;
  DevBuf<double> prox_part_, xs_part_, ys_part_, er_part_, es_part_, misc_part_, obj_;
  unsigned prox_grid_ = 2, prox_gx_ = 1, prox_gy_ = 1, xs_nb_ = 1, ys_nb_ = 1;
  unsigned long long graph_nodes_ = 0;
  volatile unsigned* host_prog_ = nullptr;
  unsigned* dev_prog_ = nullptr;
  cudaGraphExec_t graph_exec_ = nullptr;
  bool use_graph_ = true, done_init_ = false, setup_failed_ = false, profile_ = false, marking_ = false;
  bool y_recurrence_ = true;   // indirect projector: y = t_y - r between two refreshing products (POGS_B200_Y_REC=0: always the product)
  std::vector<cudaEvent_t> events_, free_events_;
  std::vector<std::pair<int, cudaEvent_t>> marks_;
  // parameters (defaults of pogs.h:20-28)
  T rho_ = T(1), abs_tol_ = T(1e-4), rel_tol_ = T(1e-3), nrmA_ = T(0);
  unsigned max_iter_ = 2500, verbose_ = 2;
  bool adaptive_rho_ = true, gap_stop_ = false;
  bool has_init_x_ = false, has_init_l_ = false;
  std::vector<T> init_x_, init_l_;
  // results
  std::vector<T> x_out_, y_out_, mu_out_, lambda_out_;
  T optval_ = T(0);
  unsigned final_iter_ = 0;
  Timing timing_;
};

// PogsDirect<T, MatrixDense<T>>: what PogsD / PogsS construct (pogs_c.cpp:19-20).
template <typename T>
class DenseSolver : public GraphSolver<T, DenseMat<T>> {
 public:
  DenseSolver(bool rowmaj, size_t m, size_t n, const T* A, bool A_on_device, bool direct = true,
              size_t m_global = 0, PeerComm* comm = nullptr)
      : GraphSolver<T, DenseMat<T>>(m, n, direct, [=](cudaStream_t st) {
          return new DenseMat<T>(rowmaj, m, n, A, A_on_device, st, m_global,
                                 comm != nullptr ? comm->view() : PeerView());
        }, m_global, comm) {}
};

// PogsIndirect<T, MatrixSparse<T>>: what PogsSparseD / PogsSparseS construct (pogs_c.cpp:69-73).
template <typename T>
class SparseSolver : public GraphSolver<T, SparseMat<T>> {
 public:
  SparseSolver(bool rowmaj, size_t m, size_t n, size_t nnz, const T* val, const int* ptr, const int* ind)
      : GraphSolver<T, SparseMat<T>>(m, n, false, [=](cudaStream_t st) {
          return new SparseMat<T>(rowmaj, m, n, nnz, val, ptr, ind, st);
        }) {}
};

}  // namespace pogs_b200
