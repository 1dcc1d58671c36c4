// One-time factorisation of I + A^T A (or I + A A^T) on the device with the library's own kernels.
//
// The reference factors the Gram matrix with its blocked right-looking Cholesky
// (src/cpu/include/gsl/gsl_linalg.h:37-55: panel, trsm, syrk per block) and applies it with two
// triangular solves per iteration (gsl_linalg.h:57-61).  On the device the per-iteration apply is a
// streaming product with the explicit inverse (graph_solver.cuh), so the one-time work is
//     G + I = L L^T          blocked right-looking Cholesky, lower, in place          (chol_lower)
//     X = L^-1               blocked triangular inverse, block row by block row       (tri_inverse_lower)
//     (G + I)^-1 = X^T X     tensor-core Gram kernel for fp32 (gram_tc.cuh), k_gemm otherwise
// Round 1 used cuSOLVER potrf / potri and cuBLAS trsm / syrk here; their first use in a process cost
// up to 28 s of library page-in on a fresh machine.  Everything below is plain CUDA-core code: the
// one-time n^3/3 + n^3/6 multiply-adds of the factor and the inverse are ~5e11 for n = 10000, a few
// tens of milliseconds at a fraction of the fp32 FMA peak, and not part of the per-iteration roofline.
//
// All matrices are row-major.  k_gemm computes C = alpha * op(A) op(B) + beta * C on tiles held in shared
// memory (classic register-blocked SGEMM); the variants needed are
//     TA = false, TB = true     C_ij = sum_k A[i][k] B[j][k]    panel solve, trailing update, A A^T
//     TA = false, TB = false    C_ij = sum_k A[i][k] B[k][j]    the two products of the triangular inverse
//     TA = true,  TB = false    C_ij = sum_k A[k][i] B[k][j]    A^T A, X^T X
#pragma once

#include "common.cuh"

namespace pogs_b200 {

constexpr int kFacNb = 64;   // block size of the factorisation (diagonal blocks are factored by one CTA)
constexpr int kFacSuper = 256;   // super-panel width of the two-level Cholesky

enum GemmTri { kTriAll = 0, kTriLower = 2 };   // kTriLower: skip tiles that lie entirely above the diagonal
// kTrimBLower: B (K x N) is lower triangular: B[k][j] = 0 for k < j;  kTrimALower: A (M x K) is lower
// triangular: A[i][k] = 0 for k > i.  The zero blocks are skipped.
enum GemmTrim { kTrimNone = 0, kTrimBLower = 1, kTrimALower = 2 };

// Batch over blockIdx.z: operand z starts zsA / zsB / zsC elements further.  With zstep > 0 the problems shrink
// along z: M_z = min(M, ztotal - z * zstep) rows (none left: nothing to do), and K_z = M_z when zkm is set
// (square triangular left operand).
struct GemmBatch {
  size_t zsA = 0, zsB = 0, zsC = 0;
  int ztotal = 0, zstep = 0, zkm = 0;
  void* C2 = nullptr;   // optional second copy of the result (same element type), leading dimension ldc2;
  size_t ldc2 = 0;      // may alias A when the grid has one column of tiles and K <= BK * steps of one CTA's own rows
};

template <typename T, int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gemm(int M, int N, int K, T alpha, const T* __restrict__ A, size_t lda, const T* __restrict__ B, size_t ldb, T beta,
       T* __restrict__ C, size_t ldc, int tri, int trim, GemmBatch gb) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ T As[BK][BM + PAD];
  __shared__ T Bs[BK][BN + PAD];
  {
    const size_t z = blockIdx.z;
    A += z * gb.zsA; B += z * gb.zsB; C += z * gb.zsC;
    if (gb.zstep > 0) {
      const int avail = gb.ztotal - static_cast<int>(z) * gb.zstep;
      if (avail <= 0) return;
      if (avail < M) M = avail;
      if (gb.zkm) K = M;
    }
  }
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  if (i0 >= M) return;
  if (tri == kTriLower && j0 > i0 + BM - 1) return;   // tile entirely above the diagonal
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  T acc[TM][TN];
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = T(0);
  int kb = 0, ke = K;
  if (trim == kTrimBLower) kb = (j0 / BK) * BK;                  // rows k < j0 of this column block of B are zero
  if (trim == kTrimALower && i0 + BM < K) ke = i0 + BM;          // columns k > i of this row block of A are zero
  for (int k0 = kb; k0 < ke; k0 += BK) {
    // ---- tiles of op(A) (BK x BM, k-major) and op(B) (BK x BN) into shared memory, zero padded ----
    if (TA) {   // A stored K x M: contiguous in i
      for (int e = tid; e < BK * BM; e += NT) {
        const int k = e / BM, i = e % BM;
        As[k][i] = (k0 + k < K && i0 + i < M) ? A[static_cast<size_t>(k0 + k) * lda + i0 + i] : T(0);
      }
    } else {    // A stored M x K: contiguous in k
      for (int e = tid; e < BK * BM; e += NT) {
        const int i = e / BK, k = e % BK;
        As[k][i] = (k0 + k < K && i0 + i < M) ? A[static_cast<size_t>(i0 + i) * lda + k0 + k] : T(0);
      }
    }
    if (TB) {   // B stored N x K: contiguous in k
      for (int e = tid; e < BK * BN; e += NT) {
        const int j = e / BK, k = e % BK;
        Bs[k][j] = (k0 + k < K && j0 + j < N) ? B[static_cast<size_t>(j0 + j) * ldb + k0 + k] : T(0);
      }
    } else {    // B stored K x N: contiguous in j
      for (int e = tid; e < BK * BN; e += NT) {
        const int k = e / BN, j = e % BN;
        Bs[k][j] = (k0 + k < K && j0 + j < N) ? B[static_cast<size_t>(k0 + k) * ldb + j0 + j] : T(0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      T a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < TM; ++r) a[r] = As[k][ty * TM + r];
#pragma unroll
      for (int c = 0; c < TN; ++c) b[c] = Bs[k][tx * TN + c];
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] += a[r] * b[c];
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int i = i0 + ty * TM + r;
    if (i >= M) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int j = j0 + tx * TN + c;
      if (j >= N) continue;
      T* dst = C + static_cast<size_t>(i) * ldc + j;
      const T val = beta == T(0) ? alpha * acc[r][c] : alpha * acc[r][c] + beta * *dst;
      *dst = val;
      if (gb.C2 != nullptr) static_cast<T*>(gb.C2)[static_cast<size_t>(i) * gb.ldc2 + j] = val;
    }
  }
}

// fp32 fast path of k_gemm for row-major A (M x K, contiguous in k) on 128 x 128 x 16 tiles: 16 B global loads,
// the next tile's loads in flight (registers) while the current one is multiplied, two shared-memory buffers
// (one barrier per tile), and every thread's 8 x 8 block split into four 4 x 4 blocks 64 rows / columns apart
// so that its shared-memory reads are conflict-free 16 B vectors.  Same products, same order of the k sum as
// k_gemm (the generic kernel: one scalar load with a div / mod per element and two barriers per tile).
// Measured on C2 (n = 10000): triangular inverse 17 -> 11.8 ms (28 TFLOP/s incl. launches); the Cholesky phase
// did not move (47 ms): its 116 trailing / eager updates are 14 ms of it, the 157 single-CTA diagonal blocks
// 16 ms, the 157 skinny panel products 4 ms, the rest sits between the 630 dependent launches.
// Needs 16 B aligned operands: gemm() checks.
template <bool TB>
__global__ void __launch_bounds__(256)
k_gemm_f32(int M, int N, int K, float alpha, const float* __restrict__ A, size_t lda, const float* __restrict__ B, size_t ldb,
           float beta, float* __restrict__ C, size_t ldc, int tri, int trim, GemmBatch gb) {
  constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  {
    const size_t z = blockIdx.z;
    A += z * gb.zsA; B += z * gb.zsB; C += z * gb.zsC;
    if (gb.zstep > 0) {
      const int avail = gb.ztotal - static_cast<int>(z) * gb.zstep;
      if (avail <= 0) return;
      if (avail < M) M = avail;
      if (gb.zkm) K = M;
    }
  }
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  if (i0 >= M) return;
  if (tri == kTriLower && j0 > i0 + BM - 1) return;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  int kb = 0, ke = K;
  if (trim == kTrimBLower) kb = (j0 / BK) * BK;
  if (trim == kTrimALower && i0 + BM < K) ke = i0 + BM;
  const int nt = (ke - kb + BK - 1) / BK;
  float4 ra[2], rb[2];
  // k-contiguous operand: vector u of the thread = row (tid + 256 u) / 4, k quad (tid + 256 u) % 4
  auto load_kmajor = [&](const float* __restrict__ P, size_t ldp, int r0, int rmax, int k0, float4 (&reg)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + 256 * u, row = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + row < rmax) {
        const float* src = P + static_cast<size_t>(r0 + row) * ldp + k0 + kq;
        if (k0 + kq + 3 < K) v = *reinterpret_cast<const float4*>(src);
        else {
          if (k0 + kq + 0 < K) v.x = src[0];
          if (k0 + kq + 1 < K) v.y = src[1];
          if (k0 + kq + 2 < K) v.z = src[2];
        }
      }
      reg[u] = v;
    }
  };
  auto store_kmajor = [&](float (&S)[BK][BM + PAD], const float4 (&reg)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + 256 * u, row = idx >> 2, kq = (idx & 3) * 4;
      S[kq + 0][row] = reg[u].x; S[kq + 1][row] = reg[u].y; S[kq + 2][row] = reg[u].z; S[kq + 3][row] = reg[u].w;
    }
  };
  // B stored K x N (contiguous in j): vector u = k (tid + 256 u) / 32, column quad (tid + 256 u) % 32
  auto load_b_n = [&](int k0, float4 (&reg)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + 256 * u, k = idx >> 5, jq = (idx & 31) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < K) {
        const float* src = B + static_cast<size_t>(k0 + k) * ldb + j0 + jq;
        if (j0 + jq + 3 < N) v = *reinterpret_cast<const float4*>(src);
        else {
          if (j0 + jq + 0 < N) v.x = src[0];
          if (j0 + jq + 1 < N) v.y = src[1];
          if (j0 + jq + 2 < N) v.z = src[2];
        }
      }
      reg[u] = v;
    }
  };
  auto store_b_n = [&](float (&S)[BK][BN + PAD], const float4 (&reg)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + 256 * u, k = idx >> 5, jq = (idx & 31) * 4;
      *reinterpret_cast<float4*>(&S[k][jq]) = reg[u];
    }
  };
  auto load_tiles = [&](int k0) {
    load_kmajor(A, lda, i0, M, k0, ra);
    if (TB) load_kmajor(B, ldb, j0, N, k0, rb); else load_b_n(k0, rb);
  };
  auto store_tiles = [&](int buf) {
    store_kmajor(As[buf], ra);
    if (TB) store_kmajor(Bs[buf], rb); else store_b_n(Bs[buf], rb);
  };
  if (nt > 0) { load_tiles(kb); store_tiles(0); }
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    const int buf = t & 1;
    if (t + 1 < nt) load_tiles(kb + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] += a[r] * b[c];
    }
    if (t + 1 < nt) store_tiles(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = i0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + r - 4);
    if (i >= M) continue;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = j0 + (c < 4 ? tx * 4 + c : 64 + tx * 4 + c - 4);
      if (j >= N) continue;
      float* dst = C + static_cast<size_t>(i) * ldc + j;
      const float val = beta == 0.f ? alpha * acc[r][c] : alpha * acc[r][c] + beta * *dst;
      *dst = val;
      if (gb.C2 != nullptr) static_cast<float*>(gb.C2)[static_cast<size_t>(i) * gb.ldc2 + j] = val;
    }
  }
}

// Tile shapes: fp32 128 x 128 (8 x 8 per thread), skinny 64 x 128 when M <= 64; fp64 64 x 64 (4 x 4).
template <typename T, bool TA, bool TB>
inline void gemm(cudaStream_t st, int M, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C,
                 size_t ldc, int tri = kTriAll, int trim = kTrimNone, unsigned nbatch = 1, GemmBatch gb = GemmBatch()) {
  if (M <= 0 || N <= 0 || nbatch == 0) return;
  if constexpr (sizeof(T) == 4) {
    if (M <= 64) {
      dim3 grid((N + 127) / 128, (M + 63) / 64, nbatch);
      k_gemm<T, 64, 128, 16, 4, 8, TA, TB><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, trim, gb);
    } else if (N <= 64) {
      dim3 grid((N + 63) / 64, (M + 127) / 128, nbatch);
      k_gemm<T, 128, 64, 16, 8, 4, TA, TB><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, trim, gb);
    } else {
      dim3 grid((N + 127) / 128, (M + 127) / 128, nbatch);
      const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
      const bool fast = !TA && al16(A) && al16(B) && lda % 4 == 0 && ldb % 4 == 0 && gb.zsA % 4 == 0 && gb.zsB % 4 == 0 &&
                        getenv("POGS_B200_GEMM_PLAIN") == nullptr;
      if constexpr (!TA) {
        if (fast) {
          k_gemm_f32<TB><<<grid, 256, 0, st>>>(M, N, K, alpha, reinterpret_cast<const float*>(A), lda,
                                                reinterpret_cast<const float*>(B), ldb, beta, reinterpret_cast<float*>(C), ldc,
                                                tri, trim, gb);
          POGS_CUDA(cudaGetLastError());
          count_launch();
          return;
        }
      }
      k_gemm<T, 128, 128, 16, 8, 8, TA, TB><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, trim, gb);
    }
  } else {
    dim3 grid((N + 63) / 64, (M + 63) / 64, nbatch);
    k_gemm<T, 64, 64, 16, 4, 4, TA, TB><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, trim, gb);
  }
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

// Diagonal block (jb <= kFacNb): D = L L^T in place (lower triangle; the strict upper triangle of the block
// is left untouched) and W = L^-1 (lower, jb x jb, leading dimension kFacNb).  One CTA of 256 threads; the
// block and its inverse live in (dynamic) shared memory.  *info is set to j0 + column + 1 when a pivot is not
// positive (== LAPACK potrf).
// One barrier per column, factor and inverse in the same sweep: at step c every thread reads the pivot
// d = s[c][c] and the still unscaled column c, updates its fixed set of trailing entries with
// s_ij -= s_ic s_jc / d (no index arithmetic: thread t owns column t % 64 and rows t / 64 + 4 r) and the finished
// column L_ic = s_ic / sqrt(d) goes to a second array, so nothing that is read in a step is written in it.
// Row c of W = L^-1 is final at the same moment (forward substitution by rows, L W = I: W_c. = (e_c - sum_{k<c}
// L_ck W_k.) / L_cc, the sum kept up to date in a third array); behind the barrier all threads subtract
// L_ic W_c. from the rows below.  The block is latency-bound: 64 dependent steps of ~1.5 k cycles (pivot ->
// reciprocal / square root -> 16 shared-memory updates per thread -> barrier), 100 us per block in either form
// (the earlier form ran a forward substitution of 64 threads with a serial dot product each after the factor).
template <typename T>
__global__ void __launch_bounds__(256) k_potf2_inv(int jb, int j0, T* __restrict__ D, size_t ld, T* __restrict__ W, int* info) {
  extern __shared__ __align__(16) unsigned char potf2_smem[];
  T (*s)[kFacNb + 1] = reinterpret_cast<T (*)[kFacNb + 1]>(potf2_smem);
  T (*L)[kFacNb + 1] = s + kFacNb;
  T (*w)[kFacNb + 1] = L + kFacNb;                          // running right-hand sides of the inverse
  T (*wrow)[kFacNb] = reinterpret_cast<T (*)[kFacNb]>(w + kFacNb);   // [2][kFacNb]: finished row of W (double-buffered)
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int cj = tid & (kFacNb - 1), r0 = tid >> 6;   // own column, first own row (rows r0 + 4 r)
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < kFacNb * kFacNb; e += 256) {
    const int i = e >> 6, j = e & (kFacNb - 1);
    s[i][j] = (i < jb && j <= i) ? D[static_cast<size_t>(i) * ld + j] : (i == j ? T(1) : T(0));   // identity padding
    L[i][j] = T(0);
    w[i][j] = i == j ? T(1) : T(0);
  }
  __syncthreads();
  for (int c = 0; c < kFacNb; ++c) {
    T d = s[c][c];
    if (!(d > T(0))) { if (tid == 0 && c < jb) s_bad = c + 1; d = T(1); }
    const T inv_d = T(1) / d, inv_sq = T(1) / m_sqrt(d);
    const T lcc = d * inv_sq;
    const T sjc = s[cj][c];
    if (cj == c) {
      // finished column c of L (rows >= c), by the threads that own column c
#pragma unroll
      for (int r = 0; r < kFacNb / 4; ++r) {
        const int i = r0 + 4 * r;
        if (i >= c) L[i][c] = i == c ? lcc : s[i][c] * inv_sq;
      }
    } else if (cj > c) {
      const T f = sjc * inv_d;
#pragma unroll
      for (int r = 0; r < kFacNb / 4; ++r) {
        const int i = r0 + 4 * r;
        if (i >= cj) s[i][cj] -= s[i][c] * f;
      }
    }
    if (r0 == (c & 3)) {   // the threads that own row c: row c of the inverse is final
      const T wv = cj <= c ? w[c][cj] / lcc : T(0);
      wrow[c & 1][cj] = wv;
      if (c < jb && cj < jb) W[static_cast<size_t>(c) * kFacNb + cj] = wv;
    }
    __syncthreads();
    if (cj <= c) {
      const T wr = wrow[c & 1][cj];
#pragma unroll
      for (int r = 0; r < kFacNb / 4; ++r) {
        const int i = r0 + 4 * r;
        if (i > c) w[i][cj] -= L[i][c] * wr;
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < jb * kFacNb; e += 256) {
    const int i = e >> 6, j = e & (kFacNb - 1);
    if (j <= i && j < jb) D[static_cast<size_t>(i) * ld + j] = L[i][j];
  }
  if (tid == 0 && s_bad != 0 && *info == 0) *info = j0 + s_bad;
}
template <typename T>
inline void launch_potf2_inv(cudaStream_t st, int jb, int j0, T* D, size_t ld, T* W, int* info) {
  constexpr size_t smem = (3 * kFacNb * (kFacNb + 1) + 2 * kFacNb) * sizeof(T);
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_index()];
  if (!attr_set) {
    POGS_CUDA(cudaFuncSetAttribute(k_potf2_inv<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  k_potf2_inv<T><<<1, 256, smem, st>>>(jb, j0, D, ld, W, info);
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

// G (n x n, row-major, leading dimension ld; only the lower triangle is read and written) = L L^T in place.
// `work` must hold n * kFacNb + kFacNb * kFacNb elements: the inverses of the diagonal blocks (kept for
// tri_inverse_lower) and one panel copy.
// == gsl::linalg_cholesky_decomp (gsl_linalg.h:37-55): per block column the diagonal block, the panel below
// it (there trsm, here a product with the inverted diagonal block) and the trailing syrk update.
template <typename T>
inline void chol_lower(cudaStream_t st, int n, T* G, size_t ld, T* work, int* info_dev) {
  T* Wall = work;                                     // [nblocks][kFacNb][kFacNb]
  T* panel = work + static_cast<size_t>((n + kFacNb - 1) / kFacNb) * kFacNb * kFacNb;   // [n][kFacNb] scratch copy of the panel
  // Two-level blocking: inside a super-panel of kFacSuper columns the 64-column panels update only the
  // super-panel's own columns right away; everything to the right of it gets ONE update with K = kFacSuper
  // when the super-panel is done (a K = 64 update reads and writes the whole trailing matrix for 64 multiply-adds
  // per element: as much memory time as FMA time).
  for (int s0 = 0; s0 < n; s0 += kFacSuper) {
    const int sw = n - s0 < kFacSuper ? n - s0 : kFacSuper;
    for (int j0 = s0; j0 < s0 + sw; j0 += kFacNb) {
      const int blk = j0 / kFacNb;
      const int jb = n - j0 < kFacNb ? n - j0 : kFacNb;
      T* W = Wall + static_cast<size_t>(blk) * kFacNb * kFacNb;
      launch_potf2_inv<T>(st, jb, j0, G + static_cast<size_t>(j0) * ld + j0, ld, W, info_dev);
      const int rest = n - j0 - jb;
      if (rest <= 0) break;
      T* P = G + static_cast<size_t>(j0 + jb) * ld + j0;   // panel below the diagonal block: rest x jb
      // L_panel = G_panel * W^T  (C_ic = sum_k Gp[i][k] W[c][k]); written to the compact scratch copy (operand of
      // the updates) and, as second output, back in place: a CTA of this product owns whole rows of the panel
      // (jb <= 64 = one tile column) and writes them only after its last read
      GemmBatch two;
      two.C2 = P; two.ldc2 = ld;
      gemm<T, false, true>(st, rest, jb, jb, T(1), P, ld, W, kFacNb, T(0), panel, kFacNb, kTriAll, kTrimNone, 1, two);
      // eager update of the super-panel's remaining columns (lower triangle): G_ic -= sum_k L[i][k] L[c][k]
      const int cols_in = s0 + sw - (j0 + jb);
      if (cols_in > 0) {
        T* Gt = G + static_cast<size_t>(j0 + jb) * ld + j0 + jb;
        gemm<T, false, true>(st, rest, cols_in, jb, T(-1), panel, kFacNb, panel, kFacNb, T(1), Gt, ld, kTriLower);
      }
    }
    const int rest2 = n - s0 - sw;
    if (rest2 > 0) {
      // everything right of the super-panel: one update with K = sw from the finished columns of L
      const T* Ls = G + static_cast<size_t>(s0 + sw) * ld + s0;
      T* Gt = G + static_cast<size_t>(s0 + sw) * ld + s0 + sw;
      gemm<T, false, true>(st, rest2, rest2, sw, T(-1), Ls, ld, Ls, ld, T(1), Gt, ld, kTriLower);
    }
  }
}

// Places the inverted diagonal blocks (kept by chol_lower in `work`) on the diagonal of X.
template <typename T>
__global__ void __launch_bounds__(256) k_place_diag(int n, const T* __restrict__ Wall, T* __restrict__ X, size_t ldx) {
  const int blk = blockIdx.x, i0 = blk * kFacNb;
  const int jb = n - i0 < kFacNb ? n - i0 : kFacNb;
  const T* W = Wall + static_cast<size_t>(blk) * kFacNb * kFacNb;
  for (int e = threadIdx.x; e < jb * jb; e += 256) {
    const int i = e / jb, j = e % jb;
    X[static_cast<size_t>(i0 + i) * ldx + i0 + j] = W[static_cast<size_t>(i) * kFacNb + j];
  }
}

// X (n x n, row-major, leading dimension ldx, zero-initialised by the caller) = L^-1 for the lower triangular
// factor left in G by chol_lower (same `work`; `tmp` holds n * n / 2 + n * kFacNb elements).
// Recursive doubling: the inverses of the kFacNb diagonal blocks are known; at block size b every pair of
// adjacent diagonal blocks [X11 0; ? X22] of size b is completed with
//     X21 = -X22 * (L21 * X11)
// -- two products per level, all pairs of a level in ONE batched launch (blockIdx.z), so that even the last
// levels run as large products instead of the 64-row strips of a row-by-row inversion (which took 170 ms for
// n = 10000: 157 dependent steps with at most 79 CTAs each).
template <typename T>
inline void tri_inverse_lower(cudaStream_t st, int n, const T* G, size_t ld, T* X, size_t ldx, T* work, T* tmp) {
  const unsigned nblk = static_cast<unsigned>((n + kFacNb - 1) / kFacNb);
  k_place_diag<T><<<nblk, 256, 0, st>>>(n, work, X, ldx);
  POGS_CUDA(cudaGetLastError());
  count_launch();
  for (int b = kFacNb; b < n; b *= 2) {
    const unsigned npairs = static_cast<unsigned>((n - b + 2 * b - 1) / (2 * b));   // pairs whose second block is not empty
    GemmBatch g1;   // T_z (m_z x b) = L21_z * X11_z,  pair z covers rows/cols [2bz, 2bz + 2b)
    g1.zsA = static_cast<size_t>(2 * b) * (ld + 1); g1.zsB = static_cast<size_t>(2 * b) * (ldx + 1);
    g1.zsC = static_cast<size_t>(b) * b; g1.ztotal = n - b; g1.zstep = 2 * b;
    gemm<T, false, false>(st, b, b, b, T(1), G + static_cast<size_t>(b) * ld, ld, X, ldx, T(0), tmp, static_cast<size_t>(b),
                          kTriAll, kTrimBLower, npairs, g1);
    GemmBatch g2;   // X21_z = -X22_z * T_z  (X22 lower triangular, m_z x m_z)
    g2.zsA = static_cast<size_t>(2 * b) * (ldx + 1); g2.zsB = static_cast<size_t>(b) * b;
    g2.zsC = static_cast<size_t>(2 * b) * (ldx + 1); g2.ztotal = n - b; g2.zstep = 2 * b; g2.zkm = 1;
    gemm<T, false, false>(st, b, b, b, T(-1), X + static_cast<size_t>(b) * ldx + b, ldx, tmp, static_cast<size_t>(b), T(0),
                          X + static_cast<size_t>(b) * ldx, ldx, kTriAll, kTrimALower, npairs, g2);
  }
}

inline size_t factor_tmp_elems(size_t n) { return n * n / 2 + n * kFacNb + 4096; }
inline size_t factor_work_elems(size_t n) {
  return ((n + kFacNb - 1) / kFacNb) * kFacNb * kFacNb + n * kFacNb + 64;
}

// Mirror the lower triangle of a row-major square array into its upper triangle.
template <typename T>
__global__ void k_mirror_lower(size_t n, T* __restrict__ Mx, size_t ld) {
  const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t i = blockIdx.y;
  if (i < n && j < n && j > i) Mx[i * ld + j] = Mx[j * ld + i];
}

}  // namespace pogs_b200
