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

#include "admm_pass.cuh"
#include "fused_pass.cuh"
#include "mat_algos.cuh"

namespace pogs_b200 {

template <typename T, bool SQ, typename Epi>
inline void launch_rowdot(cudaStream_t s, const RowdotPlan& pl, const T* M, size_t R, size_t C, size_t ld,
                          const T* v, const Epi& epi, double* partials, Gate gate,
                          const TailCtrl<T>& tail = TailCtrl<T>{nullptr, CtrlIn(), nullptr, CondSwitch{0, 0}}) {
  if (pl.cta_per_row) k_rowdot<T, SQ, 8, Epi, true><<<pl.grid, kThreads, 0, s>>>(M, R, C, ld, v, epi, partials, gate, tail);
  else k_rowdot<T, SQ, 8, Epi, false><<<pl.grid, kThreads, 0, s>>>(M, R, C, ld, v, epi, partials, gate, tail);
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

template <typename T, bool SQ, typename Epi>
inline void launch_colacc(cudaStream_t s, const ColaccPlan& pl, const T* M, size_t R, size_t C, size_t ld,
                          const T* w, T* part, unsigned* tickets, const Epi& epi, double* partials, Gate gate,
                          const PeerView& pv = PeerView()) {
  dim3 grid(pl.tiles, pl.chunks);
  k_colacc<T, SQ, 8, Epi><<<grid, kThreads, 0, s>>>(M, R, C, ld, w, pl.rows_per_chunk, part, tickets, epi,
                                                    partials, gate, pv);
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

constexpr int kPlanOcc = 4;   // CTAs per SM the streaming kernels are built for (__launch_bounds__)

// Launch shape of the single-pass kernel (fused_pass.cuh) for one operator.
struct OnePassPlan {
  bool ok = false;
  int nv = 0, batch = 0;            // 16 B column vectors per thread, rows per batch
  unsigned grid = 0, nfold = 0, fold_vecs = 0, stages = 0;
  unsigned nmap = 1;                // active map warps
  size_t smem = 0;
};

template <typename T>
class DenseMat : public MatAlgos<DenseMat<T>, T> {
 public:
  static constexpr bool kDense = true;
  // `A` is m x n, row-major (rowmaj=true) or column-major.  on_device: A is a
  // device pointer (copied device-to-device), else a host pointer.
  // Row-block multi-GPU: m is the number of LOCAL rows, m_global the row count of the
  // whole matrix and pv the peer view; A^T w is then summed over the ranks inside
  // k_colacc.  Single GPU: m_global == m and pv inactive.
  DenseMat(bool rowmaj, size_t m, size_t n, const T* A, bool on_device, cudaStream_t stream,
           size_t m_global = 0, const PeerView& pv = PeerView())
      : MatAlgos<DenseMat<T>, T>(m, n, stream, m_global ? m_global : m, pv), tstore_(!rowmaj) {
    if (pv.active() && !rowmaj) throw Error("row-block multi-GPU needs a row-major matrix");
    const DeviceInfo& dev_ = this->dev_;
    cudaStream_t stream_ = stream;
    R_ = tstore_ ? n : m;
    C_ = tstore_ ? m : n;
    constexpr size_t VEC = V16<T>::N;
    ld_ = round_up(C_, VEC);
    data_.alloc(R_ * ld_, 64);   // slack: the Gram kernel's TMA view may read up to 124 B past the last row (gram_tc.cuh)
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
    if (pv.active()) {
      if (ld_ * sizeof(T) > pv.cap_bytes) throw Error("peer communicator slot smaller than one row of A");
      if (ca_plan_.tiles > static_cast<unsigned>(kMaxTileChannels)) throw Error("too many column tiles for the peer exchange");
    }
    plan_one_pass();
    agree_one_pass();
  }

  // ---- single-pass kernel (fused_pass.cuh): one pass over the rows gives both A x -> row map
  //      and A^T coef -> column map.  Needs row-major storage and rows long enough to pay for
  //      the per-row bookkeeping.
  const OnePassPlan& one_pass_plan() const { return op_; }
  bool one_pass_ok() const { return op_.ok; }

  template <bool SQ, typename RowOp, typename ColOp>
  void one_pass(const T* x, const RowOp& rop, const ColOp& cop, const Ctrl<T>* ctrl, Gate gate = Gate{nullptr, nullptr}) {
    if (!op_.ok) throw Error("single-pass kernel is not available for this operator");
    OnePassArgs<T> a;
    a.A = data_.get(); a.m = R_; a.n = C_; a.ld = ld_;
    a.x = x;
    a.colpart = colpart_.get(); a.bar = gbar_.get();
    a.nfold = op_.nfold; a.fold_vecs = op_.fold_vecs; a.nstages = op_.stages; a.nmap = op_.nmap;
#define POGS_OP_CASE(NV, B) \
    case NV * 8 + B: launch_one_pass<SQ, NV, B>(a, rop, cop, ctrl, gate); break;
#define POGS_OP_ROW(NV) POGS_OP_CASE(NV, 1) POGS_OP_CASE(NV, 2) POGS_OP_CASE(NV, 4)
    switch (op_.nv * 8 + op_.batch) {
      POGS_OP_ROW(1) POGS_OP_ROW(2) POGS_OP_ROW(3) POGS_OP_ROW(5) POGS_OP_ROW(8)
      default: throw Error("single-pass kernel: no instantiation");
    }
#undef POGS_OP_ROW
#undef POGS_OP_CASE
    POGS_CUDA(cudaGetLastError());
    count_launch();
  }

  // One-launch ADMM iteration (admm_pass.cuh): fills the operator / launch-shape part of the
  // arguments and launches; the caller provides the iteration's buffers.
  void admm_pass(PassArgs<T> a, const ParityArgs<T>& par0, const ParityArgs<T>& par1, Gate gate, bool pdl = false) {
    if (!op_.ok) throw Error("single-pass kernel is not available for this operator");
    a.A = data_.get(); a.m = R_; a.n = C_; a.ld = ld_;
    a.colpart = colpart_.get(); a.bar = gbar_.get();
    a.nfold = op_.nfold; a.fold_vecs = op_.fold_vecs; a.nstages = op_.stages; a.nmap = op_.nmap;
    launch_admm_pass<T>(op_.nv, op_.batch, op_.grid, op_.smem, this->stream_, a, par0, par1, gate, this->pv_, pdl);
  }

  bool transposed_storage() const { return tstore_; }
  T* data() { return data_.get(); }
  size_t R() const { return R_; }
  size_t C() const { return C_; }
  size_t ld() const { return ld_; }

  // Rows of the partials array written by mul_n / mul_t.
  unsigned nb_n() const { return tstore_ ? ca_plan_.tiles : rd_plan_.grid; }
  unsigned nb_t() const { return tstore_ ? rd_plan_.grid : ca_plan_.tiles; }
  unsigned nb_max() const { return rd_plan_.grid > ca_plan_.tiles ? rd_plan_.grid : ca_plan_.tiles; }
  unsigned launches_per_product() const { return 1u; }

  // out(m) <- epi(A v),  v of length n (zero-padded buffer).
  template <bool SQ, typename Epi>
  void mul_n(const T* v, const Epi& epi, double* partials, Gate gate = Gate{nullptr, nullptr}) {
    if (!tstore_) launch_rowdot<T, SQ>(this->stream_, rd_plan_, data_.get(), R_, C_, ld_, v, epi, partials, gate);
    else launch_colacc<T, SQ>(this->stream_, ca_plan_, data_.get(), R_, C_, ld_, v, part_.get(), tickets_.get(), epi,
                              partials, gate);
  }
  // mul_n with the controller's phase 0 fused behind it when the product runs on k_rowdot
  // (row-major storage); returns false if the caller has to launch k_control itself.
  template <bool SQ, typename Epi>
  bool mul_n_tail(const T* v, const Epi& epi, double* partials, Gate gate, const TailCtrl<T>& tail) {
    if (!tstore_) {
      launch_rowdot<T, SQ>(this->stream_, rd_plan_, data_.get(), R_, C_, ld_, v, epi, partials, gate, tail);
      return true;
    }
    mul_n<SQ>(v, epi, partials, gate);
    return false;
  }
  // out(n) <- epi(A^T w),  w of length m.
  template <bool SQ, typename Epi>
  void mul_t(const T* w, const Epi& epi, double* partials, Gate gate = Gate{nullptr, nullptr}) {
    if (!tstore_) launch_colacc<T, SQ>(this->stream_, ca_plan_, data_.get(), R_, C_, ld_, w, part_.get(),
                                       tickets_.get(), epi, partials, gate, this->pv_);
    else launch_rowdot<T, SQ>(this->stream_, rd_plan_, data_.get(), R_, C_, ld_, w, epi, partials, gate);
  }

  // A := diag(d) A diag(e) * (*s) in place (matrix_dense.cpp:182-189, 227-246).
  void apply_scaling(const T* d, const T* e, const T* s_ptr) {
    const T* rs = tstore_ ? e : d;
    const T* cs = tstore_ ? d : e;
    k_scale_matrix<T><<<this->dev_.sm_count * 8, kThreads, 0, this->stream_>>>(data_.get(), R_, C_, ld_, rs, cs, s_ptr);
    POGS_CUDA(cudaGetLastError());
    count_launch();
  }

 private:
  void plan_one_pass() {
    op_ = OnePassPlan();
    if (tstore_) return;
    constexpr size_t VEC = V16<T>::N;
    const size_t nvec = ld_ / VEC;
    const size_t per_thread = (nvec + kFusedThreads - 1) / kFusedThreads;
    if (per_thread > 8) return;   // column slice no longer fits the register file: two-pass kernels
    op_.nv = per_thread <= 1 ? 1 : per_thread <= 2 ? 2 : per_thread <= 3 ? 3 : per_thread <= 5 ? 5 : 8;
    // ring of whole rows in shared memory
    const size_t row_bytes = ld_ * sizeof(T);
    // short rows leave too few bytes per row for the per-row bookkeeping of this kernel.  Round 1 drew the
    // line at 16 KB (8 KB rows ran slower than the two-pass kernels then); with four rows per batch and the
    // whole iteration in one launch 8 KB rows win clearly (C3, 50000 x 2000: 9338 vs 5869 iterations/s); the line
    // is now at 6 KB
    const char* ff = getenv("POGS_B200_FORCE_FUSE");
    if (row_bytes < 6u * 1024u && !(ff != nullptr && ff[0] == '1')) return;   // (C3's rows are 8000 B)
    size_t slots = (200u * 1024u) / row_bytes;
    if (slots > 32) slots = 32;
    if (slots < 3) return;
    // W batches of B rows are in the pipeline at any time (dot products of the newest, maps of the ones in
    // between, column update of the oldest) and hold W*B slots; the rest of the ring is in flight and must
    // cover the copy latency: at least 48 KB and 2 rows.  A map warp has W*B row times for a batch whose B
    // row maps run in parallel in its lanes, so with a long row map (an iterative prox such as the logistic
    // one: C4 ran at 5.4 TB/s with W = 4, B = 1) two rows per batch buy more than a fourth map warp.
    size_t in_flight = (48u * 1024u + row_bytes - 1) / row_bytes;
    if (in_flight < 2) in_flight = 2;
    const size_t budget = slots > in_flight ? slots - in_flight : 1;
    size_t nmap, batch;
    if (budget >= 16) { batch = 4; nmap = 4; }
    else if (budget >= 6) { batch = 2; nmap = budget / 2 < static_cast<size_t>(kFusedMapWarps) ? budget / 2 : kFusedMapWarps; }
    else { batch = 1; nmap = budget < static_cast<size_t>(kFusedMapWarps) ? budget : kFusedMapWarps; }
    if (const char* e = getenv("POGS_B200_PASS_WB")) {   // "W,B" override for experiments
      unsigned w = 0, b = 0;
      if (sscanf(e, "%u,%u", &w, &b) == 2 && w >= 1 && w <= static_cast<unsigned>(kFusedMapWarps) && (b == 1 || b == 2 || b == 4) &&
          w * b < slots) { nmap = w; batch = b; }
    }
    if (nmap < 1) nmap = 1;
    op_.nmap = static_cast<unsigned>(nmap);
    op_.batch = static_cast<int>(batch);
    op_.stages = static_cast<unsigned>(slots);
    op_.smem = slots * row_bytes;
    op_.grid = static_cast<unsigned>(this->dev_.sm_count);
    if (R_ < op_.grid) return;
    size_t fv = 16;
    while (fv * op_.grid < nvec) fv *= 2;
    if (fv > 128) return;
    op_.fold_vecs = static_cast<unsigned>(fv);
    op_.nfold = static_cast<unsigned>((nvec + fv - 1) / fv);
    if (this->pv_.active() && op_.nfold > static_cast<unsigned>(kMaxTileChannels)) return;
    colpart_.alloc(static_cast<size_t>(op_.grid) * ld_);
    gbar_.alloc(1);
    op_.ok = true;
  }

  // Row blocks: the single-pass and the two-pass paths use different peer channels, so every rank
  // must take the same one.  Eligibility depends on the rank-local row count (R_ >= grid), hence
  // the ranks agree on it here: one tiny all-reduce of "I cannot" flags at construction.
  void agree_one_pass() {
    if (!this->pv_.active()) return;
    DevBuf<float> flag(4);
    const float mine[4] = {op_.ok ? 0.f : 1.f, 0.f, 0.f, 0.f};
    POGS_CUDA(cudaMemcpyAsync(flag.get(), mine, sizeof(mine), cudaMemcpyHostToDevice, this->stream_));
    k_peer_allreduce<float><<<1, kThreads, 0, this->stream_>>>(flag.get(), 4, this->pv_);
    POGS_CUDA(cudaGetLastError());
    count_launch();
    float sum[4] = {0, 0, 0, 0};
    POGS_CUDA(cudaMemcpyAsync(sum, flag.get(), sizeof(sum), cudaMemcpyDeviceToHost, this->stream_));
    POGS_CUDA(cudaStreamSynchronize(this->stream_));
    if (sum[0] != 0.f) op_.ok = false;
  }

  template <bool SQ, int NV, int B, typename RowOp, typename ColOp>
  void launch_one_pass(const OnePassArgs<T>& a, const RowOp& rop, const ColOp& cop, const Ctrl<T>* ctrl, Gate gate) {
    auto kernel = k_fused_pass<T, SQ, NV, B, RowOp, ColOp>;
    static size_t attr_smem_dev[kMaxDevices] = {};   // per instantiation and device
    size_t& attr_smem = attr_smem_dev[current_device_index()];
    if (attr_smem < op_.smem) {
      POGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(op_.smem)));
      int nb = 0;
      POGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kFusedCta, op_.smem));
      if (nb < 1) throw Error("single-pass kernel does not fit an SM");   // the grid barrier needs co-residency
      attr_smem = op_.smem;
    }
    kernel<<<op_.grid, kFusedCta, op_.smem, this->stream_>>>(a, rop, cop, ctrl, gate, this->pv_);
  }

  bool tstore_;
  size_t R_, C_, ld_;
  OnePassPlan op_;
  DevBuf<T> colpart_;
  DevBuf<unsigned> gbar_;
  DevBuf<T> data_, part_;
  DevBuf<unsigned> tickets_;
  RowdotPlan rd_plan_;
  ColaccPlan ca_plan_;
};

}  // namespace pogs_b200
