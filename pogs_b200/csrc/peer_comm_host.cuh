// Host side of the NVLink peer-memory communicator (see peer_comm.cuh).
#pragma once

#include <cstring>
#include <vector>

#include "common.cuh"

namespace pogs_b200 {

class PeerComm {
 public:
  // cap_bytes: size of one data slot (>= the largest vector ever exchanged).
  PeerComm(int rank, int world, size_t cap_bytes) {
    if (world < 1 || world > kMaxPeers) throw Error("world size must be 1..8");
    if (rank < 0 || rank >= world) throw Error("bad rank");
    view_.rank = rank; view_.world = world;
    cap_bytes = round_up(cap_bytes, 256);
    view_.cap_bytes = cap_bytes;
    size_t off = 0;
    view_.data_off[0] = off; off += cap_bytes;
    view_.data_off[1] = off; off += cap_bytes;
    view_.gath_off[0] = off; off += cap_bytes;
    view_.gath_off[1] = off; off += cap_bytes;
    view_.spec_off[0] = off; off += cap_bytes;
    view_.spec_off[1] = off; off += cap_bytes;
    view_.scal_off[0] = off; off += 256;
    view_.scal_off[1] = off; off += 256;
    view_.scal2_off[0] = off; off += 256;
    view_.scal2_off[1] = off; off += 256;
    // push slots: one n-vector (<= 64 KB for the widths the single-pass kernels support) per source rank
    view_.push_stride = cap_bytes < (size_t(1) << 18) ? cap_bytes : (size_t(1) << 18);
    for (int kind = 0; kind < 2; ++kind)
      for (int par = 0; par < 2; ++par) { view_.push_off[kind][par] = off; off += view_.push_stride * kMaxPeers; }
    view_.scal2p_off[0] = off; off += 64 * kMaxPeers;
    view_.scal2p_off[1] = off; off += 64 * kMaxPeers;
    view_.flag_off = off; off += round_up(sizeof(unsigned) * kNumChannels * kMaxPeers, 256);
    view_.seq_off = off; off += round_up(sizeof(unsigned) * kNumChannels, 256);
    view_.err_off = off; off += 256;
    bytes_ = off;
    POGS_CUDA(cudaMalloc(&local_, bytes_));
    POGS_CUDA(cudaMemset(local_, 0, bytes_));
    POGS_CUDA(cudaDeviceSynchronize());
    view_.base[rank] = static_cast<char*>(local_);
    POGS_CUDA(cudaIpcGetMemHandle(&handle_, local_));
    opened_.assign(world, nullptr);
  }
  ~PeerComm() {
    for (void* p : opened_) if (p != nullptr) cudaIpcCloseMemHandle(p);
    if (local_ != nullptr) cudaFree(local_);
  }
  PeerComm(const PeerComm&) = delete;
  PeerComm& operator=(const PeerComm&) = delete;

  // Same process, one host thread per GPU: the regions are ordinary device pointers of the other
  // devices (unified addressing + cudaDeviceEnablePeerAccess), no IPC handles involved.
  // bases: `world` region base pointers, rank-major, from local_base() of every rank's communicator.
  void open_peers_in_process(void* const* bases) {
    for (int r = 0; r < view_.world; ++r) {
      if (r == view_.rank) continue;
      view_.base[r] = static_cast<char*>(bases[r]);
    }
    ready_ = true;
  }
  void* local_base() const { return local_; }

  static constexpr size_t kHandleBytes = sizeof(cudaIpcMemHandle_t);   // 64
  void get_handle(void* out) const { std::memcpy(out, &handle_, kHandleBytes); }

  // handles: world * 64 bytes, rank-major (from an all-gather done by the caller).
  void open_peers(const void* handles) {
    const char* h = static_cast<const char*>(handles);
    for (int r = 0; r < view_.world; ++r) {
      if (r == view_.rank) continue;
      cudaIpcMemHandle_t hd;
      std::memcpy(&hd, h + static_cast<size_t>(r) * kHandleBytes, kHandleBytes);
      void* p = nullptr;
      POGS_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      opened_[r] = p;
      view_.base[r] = static_cast<char*>(p);
    }
    ready_ = true;
  }

  const PeerView& view() const {
    if (view_.world > 1 && !ready_) throw Error("peer communicator: peers not opened");
    return view_;
  }
  int rank() const { return view_.rank; }
  int world() const { return view_.world; }
  size_t cap_bytes() const { return view_.cap_bytes; }

  bool error_raised() const {
    int e = 0;
    cudaMemcpy(&e, static_cast<char*>(local_) + view_.err_off, sizeof(int), cudaMemcpyDeviceToHost);
    return e != 0;
  }

  // In-place sum over the ranks of a device buffer of `len` elements (len padded to the
  // vector width by the caller's allocation), in slot-sized pieces.
  template <typename T>
  void allreduce(T* buf, size_t len, cudaStream_t stream) {
    if (view_.world == 1) return;
    constexpr size_t VEC = V16<T>::N;
    const size_t per_cta = static_cast<size_t>(kThreads) * VEC;
    size_t piece = view_.cap_bytes / sizeof(T);
    // every CTA waits for its twin on the other ranks: keep the grid within one resident wave
    const size_t wave = static_cast<size_t>(query_device().sm_count) * 4;
    const size_t max_piece = (wave < static_cast<size_t>(kMaxTileChannels) ? wave : kMaxTileChannels) * per_cta;
    if (piece > max_piece) piece = max_piece;
    piece = piece / per_cta * per_cta;
    if (piece == 0) throw Error("peer communicator: slot too small");
    for (size_t o = 0; o < len; o += piece) {
      const size_t l = round_up(len - o < piece ? len - o : piece, VEC);
      const unsigned grid = static_cast<unsigned>((l + per_cta - 1) / per_cta);
      k_peer_allreduce<T><<<grid, kThreads, 0, stream>>>(buf + o, l, view());
      POGS_CUDA(cudaGetLastError());
      count_launch();
    }
  }

 private:
  PeerView view_;
  void* local_ = nullptr;
  size_t bytes_ = 0;
  cudaIpcMemHandle_t handle_;
  std::vector<void*> opened_;
  bool ready_ = false;
};

}  // namespace pogs_b200
