// One-time Gram matrix G = A^T A of the (equilibrated) row-major fp32 operator on the
// 5th-generation tensor cores: tcgen05.mma kind::tf32 with 3xTF32 split precision.
//
// This is the only GEMM-shaped work of the path (reference: cblas_ssyrk in
// ProjectorDirect::Init, src/cpu/projector/projector_direct_dense.cpp:62-81; 2*m*n^2/2 =
// 1e13 flop for the BASELINE 100000 x 10000 problem, where the fp32 SIMT syrk of cuBLAS took
// ~200 ms -- a third of a one-shot PogsS call).
//
// Split precision.  Every fp32 entry is written as a = hi + lo with hi = a truncated to the
// 10-bit TF32 mantissa (exactly representable, so the tensor core's own conversion is exact
// whatever its rounding) and lo = a - hi (exact in fp32, <= 13 significant bits; the tensor
// core keeps 11 of them).  Then
//     a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b          (dropped: lo_a*lo_b <= 2^-22 |a*b|)
// i.e. three TF32 MMAs per tile step, products exact in the fp32 accumulator datapath, fp32
// accumulation in TMEM: fp32-GEMM accuracy at tensor-core speed.
//
// Shape.  G[i][j] = sum_r A[r][i] * A[r][j]: both operands are slices of the same row-major
// array with the reduction index r as the *slow* dimension, i.e. "MN-major" tiles in UMMA terms
// (the M / N index is contiguous in memory).  For 32-bit operands the tensor core transposes
// MN-major tiles in 32-byte units and accepts exactly one shared-memory layout for them, the
// "128-byte swizzle with 32-byte atoms" (UMMA layout type 1; any other type makes the MMA return
// zeros -- measured).  A 3-D TMA view (32 columns, m rows, n/32 column blocks) with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B drops a [32 x BK] box per column block into shared
// memory in exactly that canonical layout: 128-byte rows along the columns of A, swizzle period
// of 4 rows (stride byte offset 512 B), leading byte offset BK*128 B between column blocks.
//
// Tiles.  Only tiles that touch the lower triangle (i >= j) are computed; the epilogue stores
// each entry and its mirror image, so G comes out full and exactly symmetric.  One persistent
// CTA per SM walks a static list of 128 x 256 tiles; per tile the K loop streams all m rows.
//
// Accumulation.  The tensor core adds into its fp32 accumulator with truncation, a bias that
// grows linearly with the number of accumulation steps (measured: 2e-5 relative after 1000 rows,
// which would be 3e-4 after the 100000 rows of the BASELINE matrix -- no better than plain
// TF32; chunks of 256 rows still showed 5e-6).  So the K loop is cut into chunks of 64 rows: each chunk is accumulated on the tensor
// core from zero into one of two TMEM accumulators, and the chunk sums are added, rounded to
// nearest, in fp32 registers of the CUDA cores while the tensor core works on the next chunk.
//
// Roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one
// lane; also owns the TMEM allocation), warps 2-9 = accumulate + epilogue (tcgen05.ld of a
// chunk's accumulator; a TMEM lane quarter and half of the 256 columns each, 128 running sums in
// registers per thread).  Pipelines: smem full/empty (TMA <-> MMA, 4 stages of 48 KB) and TMEM
// full/empty (MMA <-> accumulate warps, 2 accumulators of 256 columns), all on mbarriers; every
// wait is bounded and raises an error flag instead of hanging the GPU.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pogs_b200 {

constexpr int kGramBM = 128;          // tile rows of G   (UMMA M)
constexpr int kGramBN = 256;          // tile columns of G (UMMA N)
constexpr int kGramBK = 16;           // rows of A per pipeline stage (2 UMMA K-steps of 8)
constexpr int kGramStages = 4;
constexpr int kGramThreads = 320;         // TMA warp, MMA warp, 8 accumulate/epilogue warps
constexpr int kGramChunkKB = 4;           // pipeline stages per accumulation chunk (4 x 16 = 64 rows of A)
constexpr int kGramColBlock = 32;     // fp32 columns per 128-byte swizzle row
constexpr uint32_t kGramBlockBytes = kGramBK * 128;                                   // one [32 x BK] box
constexpr uint32_t kGramStageBytes = 2 * (kGramBM + kGramBN) / kGramColBlock * kGramBlockBytes;   // hi+lo, I and J side
constexpr uint32_t kGramSmemBytes = kGramStages * kGramStageBytes + 1024;             // + alignment slack
constexpr uint32_t kGramTmemCols = 512;                                               // 2 accumulators x 256

// a = hi + lo, hi = a with the low 13 mantissa bits cleared.
static __global__ void __launch_bounds__(256)
k_split_tf32(const float4* __restrict__ a, float4* __restrict__ hi, float4* __restrict__ lo, size_t nvec) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = a[i];
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    hi[i] = h;
    lo[i] = l;
  }
}

namespace gram_detail {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: gives up after ~2 s (or at once when another role already gave up), raising
// *abort so that every later wait of the CTA falls through and the kernel ends.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*abort) return;
    if (clock64() - t0 > 4000000000LL) { *abort = 1; return; }
  }
}

// 3-D tiled TMA load (global -> shared), completion in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_addr(dst)), "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// UMMA shared-memory descriptor of an MN-major SWIZZLE_128B_BASE32B operand slice that starts at `addr`
// (one K-step: 8 rows x (blocks x 32 columns); see the file header for the layout).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t addr, uint64_t hi_bits) {
  return hi_bits | static_cast<uint64_t>((addr >> 4) & 0x3fffu);      // start address in the low 14 bits
}
// Everything of the descriptor but the start address.
__host__ __device__ constexpr uint64_t umma_desc_hi_bits(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16) |   // leading byte offset: next column block
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32) |   // stride byte offset: next group of 4 rows
         (1ull << 46) |                                                // descriptor version (sm_100)
         (static_cast<uint64_t>(layout) << 61);                        // 1 = SWIZZLE_128B_BASE32B
}

// Instruction descriptor: D fp32, A/B tf32, both MN-major, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// Arrive on an mbarrier once all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld_32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Static tile list: row blocks I (128 rows of G) x column blocks J (256 columns), only tiles
// with 256*J <= 128*I + 127, i.e. I >= 2J (they touch i >= j).  Order: super-rows of kSuper row
// blocks; inside a super-row column block by column block.  The 148 CTAs take consecutive tiles,
// so the tiles in flight form a compact ~12 x 12 patch of the tile grid and share their column
// slices of A in L2 (I-major order made every wave stream all n columns: 305 GB of DRAM traffic
// for the BASELINE matrix, the kernel was DRAM-bound).
struct TileWalk {
  static constexpr int kSuper = 12;
  int I = 0, J = 0, I0 = 0, Iend, NI, NJ;
  __device__ TileWalk(int ni, int nj) : Iend(ni < kSuper ? ni : kSuper), NI(ni), NJ(nj) {}
  __device__ bool valid() const { return I0 < NI; }
  __device__ void next() {
    if (++I < Iend) return;
    ++J;
    I = I0 > 2 * J ? I0 : 2 * J;
    if (J < NJ && I < Iend) return;
    I0 += kSuper;
    Iend = I0 + kSuper < NI ? I0 + kSuper : NI;
    I = I0; J = 0;
  }
  __device__ void advance(int steps) { for (int s = 0; s < steps && valid(); ++s) next(); }
};

}  // namespace gram_detail

struct GramArgs {
  float* G; size_t ldg;        // n x n output, row-major
  int n, m;
  int* error;                  // set to 1 when a wait timed out
  uint64_t desc_hi;            // shared-memory descriptor without the start address
  uint32_t idesc;              // instruction descriptor
  float* dbg_smem;             // debug: first pipeline stage as the MMA warp sees it (kGramStageBytes), or null
  float* dbg_acc;              // debug: raw accumulator of CTA 0's first tile (128 x 256), or null
};

static __global__ void __launch_bounds__(kGramThreads, 1)
k_gram_tf32x3(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, GramArgs a) {
  namespace gd = gram_detail;
  extern __shared__ unsigned char gram_smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kGramStages], s_empty[kGramStages], s_acc_full[2], s_acc_empty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_abort;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gram_smem_raw) + 1023) & ~uintptr_t(1023));

  if (threadIdx.x == 0) {
    s_abort = 0;
    for (int s = 0; s < kGramStages; ++s) { gd::mbar_init(&s_full[s], 1); gd::mbar_init(&s_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { gd::mbar_init(&s_acc_full[b], 1); gd::mbar_init(&s_acc_empty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    // TMEM allocation (whole warp), 512 columns = two 128 x 256 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gd::smem_addr(&s_tmem_base)), "r"(kGramTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  gd::tc_fence_before();
  __syncthreads();
  gd::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  volatile int* abort = &s_abort;

  const int NI = (a.n + kGramBM - 1) / kGramBM, NJ = (a.n + kGramBN - 1) / kGramBN;
  const int num_kb = (a.m + kGramBK - 1) / kGramBK;
  constexpr int kBlkI = kGramBM / kGramColBlock, kBlkJ = kGramBN / kGramColBlock;   // 4, 8
  // stage layout: [hi_I | lo_I | hi_J | lo_J]
  constexpr uint32_t kOffLoI = kBlkI * kGramBlockBytes, kOffHiJ = 2 * kBlkI * kGramBlockBytes,
                     kOffLoJ = kOffHiJ + kBlkJ * kGramBlockBytes;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      gd::TileWalk tw(NI, NJ);
      for (tw.advance(blockIdx.x); tw.valid(); tw.advance(gridDim.x)) {
        const int cbI = tw.I * kBlkI, cbJ = tw.J * kBlkJ;   // first 32-column block of each side
        for (int kb = 0; kb < num_kb; ++kb) {
          gd::mbar_wait(&s_empty[stage], phase ^ 1u, abort);
          unsigned char* st = smem + static_cast<size_t>(stage) * kGramStageBytes;
          gd::mbar_arrive_expect_tx(&s_full[stage], kGramStageBytes);
          const int r = kb * kGramBK;
          gd::tma_load_3d(st, &map_hi, &s_full[stage], 0, r, cbI);
          gd::tma_load_3d(st + kOffLoI, &map_lo, &s_full[stage], 0, r, cbI);
          gd::tma_load_3d(st + kOffHiJ, &map_hi, &s_full[stage], 0, r, cbJ);
          gd::tma_load_3d(st + kOffHiJ + kBlkI * kGramBlockBytes, &map_hi, &s_full[stage], 0, r, cbJ + kBlkI);
          gd::tma_load_3d(st + kOffLoJ, &map_lo, &s_full[stage], 0, r, cbJ);
          gd::tma_load_3d(st + kOffLoJ + kBlkI * kGramBlockBytes, &map_lo, &s_full[stage], 0, r, cbJ + kBlkI);
          if (++stage == kGramStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = a.idesc;
      const uint64_t dh = a.desc_hi;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool first = true;
      gd::TileWalk tw(NI, NJ);
      for (tw.advance(blockIdx.x); tw.valid(); tw.advance(gridDim.x)) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += kGramChunkKB) {
          const int kb1 = kb0 + kGramChunkKB < num_kb ? kb0 + kGramChunkKB : num_kb;
          gd::mbar_wait(&s_acc_empty[acc], acc_phase ^ 1u, abort);   // epilogue has drained this accumulator
          gd::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kGramBN;
          for (int kb = kb0; kb < kb1; ++kb) {
            gd::mbar_wait(&s_full[stage], phase, abort);
            gd::tc_fence_after();
            const uint32_t st = gd::smem_addr(smem + static_cast<size_t>(stage) * kGramStageBytes);
            if (first && a.dbg_smem != nullptr && blockIdx.x == 0) {
              const float* src = reinterpret_cast<const float*>(smem + static_cast<size_t>(stage) * kGramStageBytes);
              for (uint32_t w = 0; w < kGramStageBytes / 4; ++w) a.dbg_smem[w] = src[w];
            }
            first = false;
#pragma unroll
            for (int ks = 0; ks < kGramBK / 8; ++ks) {
              const uint32_t off = ks * 1024u;   // 8 rows x 128 B inside every column block
              const uint64_t hiI = gd::umma_desc_mn_sw128(st + off, dh), loI = gd::umma_desc_mn_sw128(st + kOffLoI + off, dh);
              const uint64_t hiJ = gd::umma_desc_mn_sw128(st + kOffHiJ + off, dh), loJ = gd::umma_desc_mn_sw128(st + kOffLoJ + off, dh);
              gd::umma_tf32(d_tmem, loI, hiJ, idesc, (kb != kb0) || (ks != 0));   // first MMA of a chunk overwrites
              gd::umma_tf32(d_tmem, hiI, loJ, idesc, true);
              gd::umma_tf32(d_tmem, hiI, hiJ, idesc, true);
            }
            gd::umma_commit(&s_empty[stage]);                    // smem slot free once these MMAs retire
            if (kb == kb1 - 1) gd::umma_commit(&s_acc_full[acc]);   // chunk complete
            if (++stage == kGramStages) { stage = 0; phase ^= 1u; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== accumulate + epilogue: warps 2..9.  TMEM lane quarter = warp % 4; the two warps of a
    //       quarter split the 256 columns.  Chunk sums are added in fp32 registers (round to
    //       nearest), so the truncating adder of the tensor core only ever sees 8 K-steps. =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int kCols = kGramBN / 2;                    // columns per thread
    uint32_t acc = 0, acc_phase = 0;
    bool first = true;
    gd::TileWalk tw(NI, NJ);
    for (tw.advance(blockIdx.x); tw.valid(); tw.advance(gridDim.x)) {
      float sum[kCols];
#pragma unroll
      for (int t = 0; t < kCols; ++t) sum[t] = 0.f;
      for (int kb0 = 0; kb0 < num_kb; kb0 += kGramChunkKB) {
        gd::mbar_wait(&s_acc_full[acc], acc_phase, abort);
        gd::tc_fence_after();
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kGramBN + half * kCols;
#pragma unroll
        for (int c = 0; c < kCols / 32; ++c) {
          float v[32];
          gd::tmem_ld_32(lane_addr + c * 32, v);
          if (first && a.dbg_acc != nullptr && blockIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < 32; ++t) a.dbg_acc[(q * 32 + lane) * kGramBN + half * kCols + c * 32 + t] = v[t];
          }
#pragma unroll
          for (int t = 0; t < 32; ++t) sum[c * 32 + t] += v[t];
        }
        first = false;
        gd::tc_fence_before();
        __syncwarp();
        if (lane == 0) gd::mbar_arrive(&s_acc_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int i = tw.I * kGramBM + q * 32 + lane;       // row of G held by this thread's TMEM lane
      const int j0 = tw.J * kGramBN + half * kCols;
      if (!*abort && i < a.n) {
#pragma unroll
        for (int t = 0; t < kCols; ++t) {
          const int j = j0 + t;
          if (j <= i) {   // lower triangle (j <= i < n) and its mirror image
            a.G[static_cast<size_t>(i) * a.ldg + j] = sum[t];
            a.G[static_cast<size_t>(j) * a.ldg + i] = sum[t];
          }
        }
      }
    }
  }

  gd::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    gd::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kGramTmemCols) : "memory");
  }
  if (threadIdx.x == 0 && s_abort) *a.error = 1;
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tensor_map_encoder() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    POGS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (p == nullptr || qres != cudaDriverEntryPointSuccess) throw Error("cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 3-D view of a row-major m x n fp32 array (leading dimension ld, ld % 4 == 0): (32 columns,
// m rows, ceil(n/32) column blocks), box = (32, BK, 4 blocks), 128-byte swizzle.  The last
// column block may reach past column ld of a row (it then reads the head of the next row, and
// up to 124 B past the end of the array for the last row: the caller allocates that slack);
// those columns only feed entries of G with an index >= n, which are never stored.
inline CUtensorMap gram_tensor_map(const float* base, size_t m, size_t n, size_t ld) {
  CUtensorMap map;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(kGramColBlock), static_cast<cuuint64_t>(m),
                              static_cast<cuuint64_t>((n + kGramColBlock - 1) / kGramColBlock)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld * sizeof(float)),
                                 static_cast<cuuint64_t>(kGramColBlock * sizeof(float))};
  const cuuint32_t box[3] = {kGramColBlock, kGramBK, kGramBM / kGramColBlock};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims,
                                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ")");
  return map;
}

// Slack (in floats) a row-major array needs behind its end to be read through gram_tensor_map.
constexpr size_t kGramSlackFloats = 64;

// G (n x n, row-major, leading dimension ldg, full and symmetric) = A^T A for the row-major
// m x n fp32 array A (leading dimension ld).  A must have kGramSlackFloats of slack.
inline void gram_tf32x3(cudaStream_t stream, const float* A, size_t m, size_t n, size_t ld, float* G, size_t ldg,
                        int sm_count, float* dbg_smem = nullptr, float* dbg_acc = nullptr) {
  if (ld % 4 != 0) throw Error("gram_tf32x3: leading dimension must be a multiple of 4");
  const size_t elems = m * ld;
  DevBuf<float> hi, lo;
  hi.alloc(elems, kGramSlackFloats);
  lo.alloc(elems, kGramSlackFloats);
  DevBuf<int> err(1);
  k_split_tf32<<<sm_count * 8, 256, 0, stream>>>(reinterpret_cast<const float4*>(A), reinterpret_cast<float4*>(hi.get()),
                                                 reinterpret_cast<float4*>(lo.get()), elems / 4);
  POGS_CUDA(cudaGetLastError());
  const CUtensorMap map_hi = gram_tensor_map(hi.get(), m, n, ld), map_lo = gram_tensor_map(lo.get(), m, n, ld);
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_index()];
  if (!attr_set) {
    POGS_CUDA(cudaFuncSetAttribute(k_gram_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGramSmemBytes)));
    attr_set = true;
  }
  GramArgs ga;
  ga.G = G; ga.ldg = ldg; ga.n = static_cast<int>(n); ga.m = static_cast<int>(m); ga.error = err.get();
  ga.desc_hi = gram_detail::umma_desc_hi_bits(kGramBlockBytes, 512u, 1u);
  ga.idesc = gram_detail::umma_idesc_tf32(kGramBM, kGramBN);
  ga.dbg_smem = dbg_smem; ga.dbg_acc = dbg_acc;
  // bring-up knobs: POGS_B200_GRAM_DESC="lbo_bytes,sbo_bytes,layout", POGS_B200_GRAM_IDESC=0x...
  if (const char* e = getenv("POGS_B200_GRAM_DESC")) {
    unsigned l = 0, sb = 0, ly = 0;
    if (sscanf(e, "%u,%u,%u", &l, &sb, &ly) == 3) ga.desc_hi = gram_detail::umma_desc_hi_bits(l, sb, ly);
  }
  if (const char* e = getenv("POGS_B200_GRAM_IDESC")) ga.idesc = static_cast<uint32_t>(strtoul(e, nullptr, 0));
  k_gram_tf32x3<<<sm_count, kGramThreads, kGramSmemBytes, stream>>>(map_hi, map_lo, ga);
  POGS_CUDA(cudaGetLastError());
  count_launch(2);
  int h_err = 0;
  POGS_CUDA(cudaMemcpyAsync(&h_err, err.get(), sizeof(int), cudaMemcpyDeviceToHost, stream));
  POGS_CUDA(cudaStreamSynchronize(stream));
  if (h_err != 0) throw Error("tensor-core Gram kernel: a pipeline wait timed out");
}

}  // namespace pogs_b200
