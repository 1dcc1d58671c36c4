// Tiled sparse product: out[r] = sum_k f(val[k]) * v[ind[k]]  (the row-gather form of the reference,
// src/cpu/include/gsl/gsl_spblas.h:10-40) on a 2-D tiling of one compressed copy.
//
// Why this layout.  Its predecessor, a column-blocked copy (rounds 1-2), kept the row sums of one CTA in
// shared memory and walked the column blocks, re-staging up to 192 KB of the multiplied vector per block.
// For a short-and-wide operand -- A^T of C5 is 100 000 x 1 000 000: 21 column blocks, 676 rows per CTA -- a
// CTA staged as many bytes as it streamed, in serial phases (ncu: 2.6 TB/s, 31 % DRAM, 38 % issue slots),
// a row segment was reduced over a group of lanes with shuffles and predicated-off slots, and the entries
// were fetched by per-lane loads (tens of KB in flight per SM, 4736 interleaved 256 B streams in DRAM).
//
// Here the operand is cut into P x Q tiles, P*Q a multiple of the SM count, such that the tile's slice of
// v (tc values) fits in shared memory next to a ring of TMA stages: each CTA stages v ONCE per tile and
// streams ~nnz/(P*Q) entries.  Inside a tile the rows are sorted by their number of entries and cut into
// slices of 32 rows; a slice is stored lane-interleaved in PAIRS of entries (sliced ELL, padded to the
// longest row of the slice, which after the sort is within one entry of all the others), so one lane
// owns one row: no shuffles, no predicates, and the entries of a row are summed in their original
// order (deterministic).  The pair-rows (32 pairs, one per lane) of the whole copy form one byte stream
// in chunks of kTlChunkDef pair-rows -- [values of the chunk][16-bit local columns of the chunk], 1.5 KB
// in fp32 -- which every warp fetches for its own contiguous share of the tile (shares balanced when the
// layout is built, `wsplit`) with 1-D bulk copies (cp.async.bulk + mbarrier, the TMA engine) into a
// private two-stage ring: 48 KB of ring per SM, no CTA-wide barrier inside a tile.
// The row sums go straight to a Q x rows array (scattered 4 B stores that L2 merges; C5: 16 MB per
// product, 2.5 % of the streamed bytes); a second small kernel folds the Q partial sums of a row in fixed
// order and runs the fused epilogue of the caller.  6 B per entry + ~2 % padding.
#pragma once

#include <cub/device/device_scan.cuh>

#include "fused_pass.cuh"
#include "sparse_kernels.cuh"

namespace pogs_b200 {

constexpr size_t kTlSmemBytes = 226u * 1024u;     // dynamic shared memory of the product kernel
// Warps of the product kernel, pair-rows per ring stage, ring stages per warp.  Measured on C5 (us per ADMM
// iteration = 4 products + the rest): (16,8,2) 623, (16,4,4) 625, (24,4,3) 622, (32,4,3) 676, (32,4,2) 600-620,
// (16,4,3) 593, (32,2,2) 587, (16,4,2) 590: bytes in flight are not the limiter (the product runs at 0.9 of
// the HBM peak), a small ring leaves the most shared memory to the slice of v.
constexpr int kTlWarpsDef = 16;
constexpr int kTlChunkDef = 4;
constexpr int kTlStagesDef = 2;
constexpr int kTlMaxWarps = 32;
constexpr unsigned kTlNoRow = 0xffffu;
constexpr int kTlMaxQ = 128;

struct TiledShape {
  unsigned P = 0, Q = 0;      // row tiles x column tiles
  unsigned tr = 0, tc = 0;    // rows / columns per tile (tc multiple of 32)
  unsigned ns = 0;            // slices of 32 rows per tile
  unsigned ntiles = 0;
  unsigned chunk = kTlChunkDef;   // pair-rows per chunk of the stream
  unsigned warps = kTlWarpsDef;   // warps of the product kernel (shares of a tile in `wsplit`)
};

template <typename T> struct TlPair;
template <> struct __align__(8) TlPair<float> { float x, y; };
template <> struct __align__(16) TlPair<double> { double x, y; };

// bytes of one chunk: kTlChunk x 32 pairs of values, then kTlChunk x 32 pairs of 16-bit columns
template <typename T> __host__ __device__ constexpr unsigned tl_chunk_bytes(unsigned ch) { return ch * 32 * (2 * sizeof(T) + 4); }
template <typename T> __host__ __device__ constexpr unsigned tl_chunk_ind_off(unsigned ch) { return ch * 32 * 2 * sizeof(T); }
template <typename T> __host__ __device__ constexpr size_t tl_ring_bytes(unsigned warps, unsigned ch, unsigned st) {
  return static_cast<size_t>(warps) * st * tl_chunk_bytes<T>(ch);
}
// byte address of pair-row g, lane l inside the stream
template <typename T> __host__ __device__ __forceinline__ size_t tl_val_off(size_t g, unsigned l, unsigned ch) {
  return (g / ch) * tl_chunk_bytes<T>(ch) + ((g % ch) * 32 + l) * 2 * sizeof(T);
}
template <typename T> __host__ __device__ __forceinline__ size_t tl_ind_off(size_t g, unsigned l, unsigned ch) {
  return (g / ch) * tl_chunk_bytes<T>(ch) + tl_chunk_ind_off<T>(ch) + ((g % ch) * 32 + l) * 4;
}

// stream: see above; soff[t*ns + s] = first pair of slice s of tile t (multiples of 32; soff[ntiles*ns] =
// total); rowid[(t*ns + s)*32 + lane] = row of that lane inside the tile (kTlNoRow: none);
// wsplit[t*(kTlWarps+1) + w] = first slice of warp w in tile t;  part = Q x rows partial sums.
template <typename T, bool SQ, int kTlWarps, int kTlChunk, int kTlStages>
__global__ void __launch_bounds__(kTlWarps * 32, 1)
k_spmv_tiled(const unsigned char* __restrict__ stream, const int* __restrict__ soff,
             const unsigned short* __restrict__ rowid, const unsigned* __restrict__ wsplit, size_t rows, size_t cols,
             TiledShape sh, const T* __restrict__ v, T* __restrict__ part, Gate gate) {
  if (gate_closed(gate)) return;
  constexpr unsigned CB = tl_chunk_bytes<T>(kTlChunk);
  constexpr int kTlThreads = kTlWarps * 32;
  extern __shared__ __align__(128) unsigned char tl_smem[];
  __shared__ __align__(8) uint64_t s_full[kTlWarps][kTlStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* ring = tl_smem + static_cast<size_t>(warp) * kTlStages * CB;
  T* s_v = reinterpret_cast<T*>(tl_smem + tl_ring_bytes<T>(kTlWarps, kTlChunk, kTlStages));
  if (lane == 0) {
    for (int st = 0; st < kTlStages; ++st) mbar_init(&s_full[warp][st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  unsigned it = 0;   // chunks this warp has consumed so far: stage = it % kTlStages, parity = (it / kTlStages) & 1

  for (unsigned t = blockIdx.x; t < sh.ntiles; t += gridDim.x) {
    const unsigned p = t / sh.Q, q = t % sh.Q;
    const size_t c0 = static_cast<size_t>(q) * sh.tc, r0 = static_cast<size_t>(p) * sh.tr;
    const unsigned cn = static_cast<unsigned>(cols - c0 < sh.tc ? cols - c0 : sh.tc);
    // this warp's share of the tile: slices [sA, sB), pair-rows [K0, K1), chunks [ch0, ch1)
    const unsigned sA = __ldg(wsplit + static_cast<size_t>(t) * (kTlWarps + 1) + warp);
    const unsigned sB = __ldg(wsplit + static_cast<size_t>(t) * (kTlWarps + 1) + warp + 1);
    const int* __restrict__ so = soff + static_cast<size_t>(t) * sh.ns;
    const unsigned short* __restrict__ rid_t = rowid + static_cast<size_t>(t) * sh.ns * 32;
    T* __restrict__ po = part + static_cast<size_t>(q) * rows + r0;
    const int K0 = sA < sB ? __ldg(so + sA) >> 5 : 0;
    const int K1 = sA < sB ? __ldg(so + sB) >> 5 : 0;
    const int ch0 = K0 / kTlChunk, ch1 = (K1 + kTlChunk - 1) / kTlChunk;
    if (lane == 0) {   // the first stages are on their way while the slice of v is staged
      for (int st = 0; st < kTlStages; ++st) {
        if (ch0 + st < ch1) {
          const unsigned sl = (it + st) % kTlStages;
          mbar_expect_tx(&s_full[warp][sl], CB);
          bulk_g2s(ring + sl * CB, stream + static_cast<size_t>(ch0 + st) * CB, CB, &s_full[warp][sl]);
        }
      }
    }
    __syncthreads();   // every warp is done with the previous tile's slice of v
    {
      const T* __restrict__ vs = v + c0;
      if ((reinterpret_cast<uintptr_t>(vs) & 15u) == 0) {
        constexpr int kV = 16 / sizeof(T);
        const unsigned nv = cn / kV;
        const int4* __restrict__ v4 = reinterpret_cast<const int4*>(vs);
        int4* s4 = reinterpret_cast<int4*>(s_v);
        for (unsigned i = tid; i < nv; i += kTlThreads) s4[i] = __ldg(v4 + i);
        for (unsigned i = nv * kV + tid; i < cn; i += kTlThreads) s_v[i] = __ldg(vs + i);
      } else {
        for (unsigned i = tid; i < cn; i += kTlThreads) s_v[i] = __ldg(vs + i);
      }
    }
    __syncthreads();
    if (sA >= sB) continue;
    unsigned s = sA;
    int nb = __ldg(so + s + 1) >> 5;                       // end of the current slice (pair-rows)
    unsigned rid = rid_t[static_cast<size_t>(s) * 32 + lane];
    int nb_n = s + 1 < sB ? __ldg(so + s + 2) >> 5 : K1;   // one slice ahead
    unsigned rid_n = s + 1 < sB ? rid_t[static_cast<size_t>(s + 1) * 32 + lane] : kTlNoRow;
    T acc = T(0);
    for (int ch = ch0; ch < ch1; ++ch, ++it) {
      const unsigned sl = it % kTlStages;
      mbar_wait(&s_full[warp][sl], (it / kTlStages) & 1u);
      const TlPair<T>* cv = reinterpret_cast<const TlPair<T>*>(ring + sl * CB);
      const unsigned* ci = reinterpret_cast<const unsigned*>(ring + sl * CB + tl_chunk_ind_off<T>(kTlChunk));
      T pr[kTlChunk];
#pragma unroll
      for (int u = 0; u < kTlChunk; ++u) {
        const TlPair<T> a = cv[u * 32 + lane];
        const unsigned ix = ci[u * 32 + lane];
        const T x0 = s_v[ix & 0xffffu], x1 = s_v[ix >> 16];
        if (SQ) pr[u] = a.x * a.x * x0 + a.y * a.y * x1;
        else    pr[u] = a.x * x0 + a.y * x1;
      }
      __syncwarp();   // every lane has read the stage
      if (lane == 0 && ch + kTlStages < ch1) {
        mbar_expect_tx(&s_full[warp][sl], CB);
        bulk_g2s(ring + sl * CB, stream + static_cast<size_t>(ch + kTlStages) * CB, CB, &s_full[warp][sl]);
      }
#pragma unroll
      for (int u = 0; u < kTlChunk; ++u) {
        const int kk = ch * kTlChunk + u;
        if (kk >= K0 && kk < K1) {   // warp-uniform
          while (kk >= nb) {         // slice finished (empty slices fall through and store their zeros)
            if (rid != kTlNoRow) po[rid] = acc;
            acc = T(0);
            ++s;
            nb = nb_n; rid = rid_n;
            nb_n = s + 1 < sB ? __ldg(so + s + 2) >> 5 : K1;
            rid_n = s + 1 < sB ? rid_t[static_cast<size_t>(s + 1) * 32 + lane] : kTlNoRow;
          }
          acc += pr[u];
        }
      }
    }
    if (rid != kTlNoRow) po[rid] = acc;
    for (++s; s < sB; ++s) {   // trailing slices without entries
      const unsigned r = rid_t[static_cast<size_t>(s) * 32 + lane];
      if (r != kTlNoRow) po[r] = T(0);
    }
  }
}

// The product kernel is shared by every epilogue: one high-water mark per kernel and device.
template <typename T, bool SQ, int W, int CH, int ST>
inline void tl_launch_one(unsigned grid, size_t smem, cudaStream_t st, const unsigned char* stream, const int* soff,
                          const unsigned short* rowid, const unsigned* wsplit, size_t rows, size_t cols, const TiledShape& sh,
                          const T* v, T* part, Gate gate) {
  static size_t mark[kMaxDevices] = {};
  size_t& m = mark[current_device_index()];
  if (m < smem) {
    POGS_CUDA(cudaFuncSetAttribute(k_spmv_tiled<T, SQ, W, CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    m = smem;
  }
  k_spmv_tiled<T, SQ, W, CH, ST><<<grid, W * 32, smem, st>>>(stream, soff, rowid, wsplit, rows, cols, sh, v, part, gate);
}
template <typename T, bool SQ>
inline void tl_launch(unsigned grid, size_t smem, cudaStream_t st, const unsigned char* stream, const int* soff,
                      const unsigned short* rowid, const unsigned* wsplit, size_t rows, size_t cols, const TiledShape& sh,
                      const T* v, T* part, Gate gate) {
  tl_launch_one<T, SQ, kTlWarpsDef, kTlChunkDef, kTlStagesDef>(grid, smem, st, stream, soff, rowid, wsplit, rows, cols, sh, v, part, gate);
}

// out[r] = epilogue(sum over the Q partial sums of row r, in fixed order)
template <typename T, typename Epi>
__global__ void __launch_bounds__(kThreads)
k_tl_fold(const T* __restrict__ part, size_t rows, unsigned Q, Epi epi, double* __restrict__ partials, Gate gate) {
  if (gate_closed(gate)) return;
  double red[Epi::NRED];
#pragma unroll
  for (int k = 0; k < Epi::NRED; ++k) red[k] = 0.0;
  for (size_t r = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; r < rows; r += static_cast<size_t>(gridDim.x) * kThreads) {
    T sum = T(0);
    for (unsigned q0 = 0; q0 < Q; q0 += 8) {   // eight loads in flight, added in the order of q
      T pv[8];
#pragma unroll
      for (unsigned u = 0; u < 8; ++u) pv[u] = q0 + u < Q ? __ldcg(part + static_cast<size_t>(q0 + u) * rows + r) : T(0);
#pragma unroll
      for (unsigned u = 0; u < 8; ++u) if (q0 + u < Q) sum += pv[u];
    }
    epi(r, sum, red);
  }
  if (partials != nullptr) block_fold<Epi::NRED>(red, partials + static_cast<size_t>(blockIdx.x) * Epi::NRED);
}

// val *= rs[row] * cs[col] * (*s) on the tiled layout (matrix_sparse.cpp:268-304); one warp per slice.
template <typename T>
__global__ void __launch_bounds__(256)
k_spscale_tiled(unsigned char* __restrict__ stream, const int* __restrict__ soff, const unsigned short* __restrict__ rowid,
                TiledShape sh, const T* __restrict__ rs, const T* __restrict__ cs, const T* __restrict__ s_ptr) {
  const T s = *s_ptr;
  const unsigned lane = threadIdx.x & 31;
  const size_t nsl = static_cast<size_t>(sh.ntiles) * sh.ns;
  for (size_t sl = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); sl < nsl; sl += static_cast<size_t>(gridDim.x) * 8) {
    const unsigned t = static_cast<unsigned>(sl / sh.ns);
    const unsigned p = t / sh.Q, q = t % sh.Q;
    const unsigned rid = rowid[sl * 32 + lane];
    if (rid == kTlNoRow) continue;
    const T rr = rs[static_cast<size_t>(p) * sh.tr + rid] * s;
    const T* __restrict__ csq = cs + static_cast<size_t>(q) * sh.tc;
    for (int g = soff[sl] >> 5; g < (soff[sl + 1] >> 5); ++g) {
      TlPair<T>* a = reinterpret_cast<TlPair<T>*>(stream + tl_val_off<T>(g, lane, sh.chunk));
      const unsigned ix = *reinterpret_cast<const unsigned*>(stream + tl_ind_off<T>(g, lane, sh.chunk));
      // padding entries are zeros with local column 0, which always exists
      a->x *= rr * csq[ix & 0xffffu];
      a->y *= rr * csq[ix >> 16];
    }
  }
}

// ---- layout conversion (one-time, on the device) ---------------------------------------------------
// step 1: entries per (tile, row): cnt[(p*Q + q)*tr + rl], and the longest row segment
static __global__ void __launch_bounds__(256)
k_tl_count(const int* __restrict__ ptr, const int* __restrict__ ind, size_t rows, TiledShape sh, unsigned* __restrict__ cnt,
           unsigned* __restrict__ longest) {
  const size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const size_t p = r / sh.tr, rl = r % sh.tr;
  for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
    const unsigned q = static_cast<unsigned>(ind[k]) / sh.tc;
    cnt[(p * sh.Q + q) * sh.tr + rl] += 1;   // private to the row's thread
  }
  unsigned mx = 0;
  for (unsigned q = 0; q < sh.Q; ++q) mx = max(mx, cnt[(p * sh.Q + q) * sh.tr + rl]);
  atomicMax(longest, mx);
}
// step 2: counting sort of the rows of every tile by their number of entries, descending.  Which of the
// rows with equal counts comes first is left to the atomics: the placement of a row decides the lane that
// sums it, not the sum.
static __global__ void __launch_bounds__(256)
k_tl_hist(const unsigned* __restrict__ cnt, size_t len, unsigned tr, unsigned nbin, unsigned* __restrict__ hist) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < len) atomicAdd(hist + (i / tr) * nbin + cnt[i], 1u);
}
static __global__ void __launch_bounds__(32)
k_tl_starts(unsigned* __restrict__ hist, unsigned ntiles, unsigned nbin) {   // in place: first position of every count
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  unsigned run = 0;
  for (unsigned c = nbin; c-- > 0;) {
    const unsigned h = hist[static_cast<size_t>(t) * nbin + c];
    hist[static_cast<size_t>(t) * nbin + c] = run;
    run += h;
  }
}
static __global__ void __launch_bounds__(256)
k_tl_place(const unsigned* __restrict__ cnt, size_t rows, TiledShape sh, unsigned nbin, unsigned* __restrict__ next,
           unsigned* __restrict__ pos, unsigned* __restrict__ cnt_sorted, unsigned short* __restrict__ rowid) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(sh.ntiles) * sh.tr) return;
  const size_t t = i / sh.tr, rl = i % sh.tr, p = t / sh.Q;
  const unsigned c = cnt[i];
  const unsigned j = atomicAdd(next + t * nbin + c, 1u);
  pos[i] = j;
  cnt_sorted[t * sh.tr + j] = c;
  if (p * sh.tr + rl < rows) rowid[t * sh.ns * 32 + j] = static_cast<unsigned short>(rl);   // others keep kTlNoRow
}
// step 3: pairs per slice = 32 * ceil(longest row of the slice / 2)
static __global__ void __launch_bounds__(256)
k_tl_slices(const unsigned* __restrict__ cnt_sorted, TiledShape sh, int* __restrict__ spairs) {
  const size_t sl = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (sl >= static_cast<size_t>(sh.ntiles) * sh.ns) return;
  const size_t t = sl / sh.ns, s = sl % sh.ns;
  const unsigned longest = cnt_sorted[t * sh.tr + s * 32];
  spairs[sl] = static_cast<int>(((longest + 1) >> 1) * 32);
}
// step 5: move every entry to its place; the entries of a row keep their order inside a tile.
template <typename T>
__global__ void __launch_bounds__(256)
k_tl_scatter(const int* __restrict__ ptr, const int* __restrict__ ind, const T* __restrict__ val, size_t rows,
             TiledShape sh, const int* __restrict__ soff, const unsigned* __restrict__ pos, unsigned char* __restrict__ stream) {
  const size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const size_t p = r / sh.tr, rl = r % sh.tr;
  int seen[kTlMaxQ];
  for (unsigned q = 0; q < sh.Q; ++q) seen[q] = 0;
  for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
    const unsigned c = static_cast<unsigned>(ind[k]);
    const unsigned q = c / sh.tc;
    const size_t t = p * sh.Q + q;
    const unsigned j = pos[t * sh.tr + rl];
    const int e = seen[q]++;
    const size_t g = (static_cast<size_t>(soff[t * sh.ns + (j >> 5)]) >> 5) + static_cast<size_t>(e >> 1);
    *reinterpret_cast<T*>(stream + tl_val_off<T>(g, j & 31, sh.chunk) + (e & 1) * sizeof(T)) = val[k];
    *reinterpret_cast<unsigned short*>(stream + tl_ind_off<T>(g, j & 31, sh.chunk) + (e & 1) * 2) = static_cast<unsigned short>(c - q * sh.tc);
  }
}
// step 6: per tile, first slice of every warp: shares of (about) equal pair counts
static __global__ void __launch_bounds__(kTlMaxWarps + 1)
k_tl_wsplit(const int* __restrict__ soff, TiledShape sh, unsigned* __restrict__ wsplit) {
  const unsigned t = blockIdx.x, w = threadIdx.x;
  const int* so = soff + static_cast<size_t>(t) * sh.ns;
  const long long lo = so[0], hi = so[sh.ns];
  const unsigned kTlWarps = sh.warps;
  const long long want = lo + (hi - lo) * w / kTlWarps;
  unsigned a = 0, b = sh.ns;   // first slice whose start is >= want
  while (a < b) {
    const unsigned mid = (a + b) >> 1;
    if (so[mid] < want) a = mid + 1; else b = mid;
  }
  wsplit[static_cast<size_t>(t) * (kTlWarps + 1) + w] = w == kTlWarps ? sh.ns : (w == 0 ? 0u : a);
}

// Tile grid for a rows x cols operand with nnz entries on ncta SMs (host).  Cost model in bytes per
// CTA: whole waves of tiles (entries at s+2 bytes, staging of the slice of v, a fixed cost per tile) +
// the partial sums written and read back.
inline bool plan_tiled(size_t rows, size_t cols, size_t nnz, unsigned ncta, size_t elem, size_t ring, TiledShape* out) {
  if (rows == 0 || cols == 0 || nnz == 0 || ring + 1024 >= kTlSmemBytes) return false;
  const size_t tc_max = std::min<size_t>((kTlSmemBytes - ring) / elem / 32 * 32, 65536);
  const size_t tr_max = 65535 / 32 * 32;
  double best = -1;
  TiledShape bs;
  const size_t qmax = std::min<size_t>(kTlMaxQ, (cols + 31) / 32);
  for (size_t Q = 1; Q <= qmax; ++Q) {
    const size_t tc = round_up((cols + Q - 1) / Q, 32);
    if (tc > tc_max) continue;
    if ((cols + tc - 1) / tc != Q) continue;             // this Q leaves an empty column tile
    const size_t pmin = (rows + tr_max - 1) / tr_max;
    for (size_t P = pmin; P <= pmin + 2 * ncta && P <= rows; ++P) {
      const size_t tr = (rows + P - 1) / P;
      const size_t Pe = (rows + tr - 1) / tr;
      const size_t tiles = Pe * Q;
      const double waves = static_cast<double>((tiles + ncta - 1) / ncta);
      const double tile_bytes = static_cast<double>(nnz) / tiles * (elem + 2) * 1.03 + tc * elem + 1.5e5;
      const double cost = waves * tile_bytes + 2.0 * Q * rows * elem / ncta;
      if (best < 0 || cost < best) {
        best = cost;
        bs.P = static_cast<unsigned>(Pe); bs.Q = static_cast<unsigned>(Q);
        bs.tr = static_cast<unsigned>(tr); bs.tc = static_cast<unsigned>(tc);
        bs.ns = static_cast<unsigned>((tr + 31) / 32);
        bs.ntiles = static_cast<unsigned>(tiles);
      }
    }
  }
  if (best < 0) return false;
  *out = bs;
  return true;
}

// One compressed copy in the tiled layout.
template <typename T>
struct TiledCopy {
  TiledShape sh;
  DevBuf<unsigned char> stream;   // chunks of pair-rows
  DevBuf<int> soff;
  DevBuf<unsigned short> rowid;
  DevBuf<unsigned> wsplit;
  DevBuf<T> part;                 // Q x rows partial sums
  size_t pairs = 0;               // incl. padding
  size_t smem = 0;
  unsigned fold_grid = 1;
  bool ok = false;

  // (ptr, ind, val): plain CSR of the copy on the device.
  bool build(const int* ptr, const int* indp, const T* valp, size_t rows, size_t cols, size_t nnz, unsigned ncta,
             cudaStream_t stream_) {
    Trace tr_;
    const unsigned ch = kTlChunkDef, nst = kTlStagesDef, nw = kTlWarpsDef;
    const size_t ring = tl_ring_bytes<T>(nw, ch, nst);
    if (!plan_tiled(rows, cols, nnz, ncta, sizeof(T), ring, &sh)) return false;
    sh.chunk = ch; sh.warps = nw;
    const size_t L = static_cast<size_t>(sh.ntiles) * sh.tr;
    const size_t nsl = static_cast<size_t>(sh.ntiles) * sh.ns;
    if (L > 0x7fffffffULL || nsl * 32 > 0x7fffffffULL) return false;
    const unsigned tb = 256;
    const unsigned gr = static_cast<unsigned>((rows + tb - 1) / tb), gl = static_cast<unsigned>((L + tb - 1) / tb);
    DevBuf<unsigned> cnt(L), pos(L), cnt_s(L), longest(1);
    k_tl_count<<<gr, tb, 0, stream_>>>(ptr, indp, rows, sh, cnt.get(), longest.get());
    unsigned nbin = 0;
    POGS_CUDA(cudaMemcpyAsync(&nbin, longest.get(), sizeof(unsigned), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    nbin += 1;
    tr_.mark("tiled: count", stream_);
    rowid.alloc(nsl * 32);
    POGS_CUDA(cudaMemsetAsync(rowid.get(), 0xff, nsl * 32 * sizeof(unsigned short), stream_));
    {
      DevBuf<unsigned> hist(static_cast<size_t>(sh.ntiles) * nbin);
      k_tl_hist<<<gl, tb, 0, stream_>>>(cnt.get(), L, sh.tr, nbin, hist.get());
      k_tl_starts<<<(sh.ntiles + 31) / 32, 32, 0, stream_>>>(hist.get(), sh.ntiles, nbin);
      k_tl_place<<<gl, tb, 0, stream_>>>(cnt.get(), rows, sh, nbin, hist.get(), pos.get(), cnt_s.get(), rowid.get());
      POGS_CUDA(cudaGetLastError());
      POGS_CUDA(cudaStreamSynchronize(stream_));
    }
    tr_.mark("tiled: sort rows", stream_);
    {
      DevBuf<int> spairs(nsl + 1);
      soff.alloc(nsl + 1);
      k_tl_slices<<<static_cast<unsigned>((nsl + tb - 1) / tb), tb, 0, stream_>>>(cnt_s.get(), sh, spairs.get());
      size_t tmp_bytes = 0;
      POGS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, spairs.get(), soff.get(), static_cast<int>(nsl + 1), stream_));
      DevBuf<char> tmp(tmp_bytes);
      POGS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, spairs.get(), soff.get(), static_cast<int>(nsl + 1), stream_));
      POGS_CUDA(cudaStreamSynchronize(stream_));
    }
    int total = 0;
    POGS_CUDA(cudaMemcpyAsync(&total, soff.get() + nsl, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    POGS_CUDA(cudaStreamSynchronize(stream_));
    tr_.mark("tiled: slice offsets", stream_);
    // a wrapped 32-bit sum shows as a total below half the entry count
    if (total < 0 || static_cast<size_t>(total) * 2 < nnz || static_cast<size_t>(total) > 0x3fffffffULL) return false;
    pairs = static_cast<size_t>(total);
    const size_t chunks = (pairs / 32 + ch - 1) / ch + nst;   // whole chunks (+ slack for the last prefetch)
    stream.alloc(chunks * tl_chunk_bytes<T>(ch));
    tr_.mark("tiled: alloc stream", stream_);
    k_tl_scatter<T><<<gr, tb, 0, stream_>>>(ptr, indp, valp, rows, sh, soff.get(), pos.get(), stream.get());
    wsplit.alloc(static_cast<size_t>(sh.ntiles) * (nw + 1));
    k_tl_wsplit<<<sh.ntiles, nw + 1, 0, stream_>>>(soff.get(), sh, wsplit.get());
    POGS_CUDA(cudaGetLastError());
    POGS_CUDA(cudaStreamSynchronize(stream_));
    tr_.mark("tiled: scatter", stream_);
    part.alloc(static_cast<size_t>(sh.Q) * rows);
    smem = ring + static_cast<size_t>(sh.tc) * sizeof(T);
    const size_t fg = (rows + kThreads - 1) / kThreads;
    // (a grid of 32 instead of 8 CTAs per SM -- one row per thread on C5 -- was measured: no gain)
    fold_grid = static_cast<unsigned>(std::min<size_t>(fg > 0 ? fg : 1, static_cast<size_t>(ncta) * 8));
    ok = true;
    return true;
  }
};

}  // namespace pogs_b200
