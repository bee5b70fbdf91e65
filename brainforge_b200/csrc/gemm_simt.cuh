// FP32 SIMT contraction core (the BF_PREC_FP32 path): C[M,N] = sum_k A(m,k) * B(k,n) with FP32 FMA
// accumulation, tiled BMxBNx16 through shared memory, 256 threads, register micro-tiles.
// Operands are described by small loader functors so the same mainloop serves dense forward /
// dW / dX, implicit-GEMM convolution (forward, full-mode dX, dF) and the LSTM / fitness GEMMs without
// ever materialising an im2col matrix or a transposed copy.
//
// Loader contract (A side shown; B is the same with n in place of m):
//   static constexpr bool kContigMN   -- true when consecutive m are contiguous in HBM for fixed k
//                                        (thread->element mapping is chosen for coalescing from it)
//   Row  row(int m) const             -- per-row state, hoisted out of the K loop
//   Tile tile(int k0) const           -- per-K-tile state (k0 = first k of the tile)
//   float load(Row, Tile, int kk, int k) const   -- element (m, k = k0+kk); only called in bounds
//   void prepare(int* smem_tables) const         -- optional cooperative table build (then __syncthreads)
// Epilogue contract:  void store(int m, int n, float acc, int split) const
#pragma once
#include "common.cuh"

namespace bf {

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_KALIGN = 32;   // K granularity shared by the SIMT (16) and tcgen05 (32) cores

template <int BM, int BN, class AL, class BL, class EP>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(AL al, BL bl, EP ep, int M, int N, int K, int k_per_split) {
  constexpr int BK = GEMM_BK;
  constexpr int TM = BM / 16, TN = BN / 16;  // 16x16 thread grid
  static_assert(BM % 16 == 0 && BN % 16 == 0, "tile");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  extern __shared__ int tables[];

  al.prepare(tables);
  bl.prepare(tables);
  if (AL::kHasTables || BL::kHasTables) __syncthreads();

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int split = blockIdx.z;
  const int k_begin = split * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);

  // ---- thread -> tile element mapping for the global->shared copies
  constexpr int A_PER = BM * BK / GEMM_THREADS;  // elements of the A tile per thread
  constexpr int B_PER = BN * BK / GEMM_THREADS;
  static_assert(A_PER >= 1 && B_PER >= 1, "tile too small for 256 threads");
  // contiguous-in-MN loaders: element e -> (mn = e % B?, kk = e / B?)  => fixed mn per thread
  // contiguous-in-K  loaders: element e -> (kk = e % BK, mn = e / BK) => fixed kk per thread
  constexpr int A_ROWS = AL::kContigMN ? 1 : A_PER;
  constexpr int B_ROWS = BL::kContigMN ? 1 : B_PER;
  typename AL::Row arow[A_ROWS];
  typename BL::Row brow[B_ROWS];
  bool arow_ok[A_ROWS], brow_ok[B_ROWS];
#pragma unroll
  for (int i = 0; i < A_ROWS; ++i) {
    int ml = AL::kContigMN ? (tid % BM) : (tid / BK + i * (GEMM_THREADS / BK));
    arow_ok[i] = (m0 + ml) < M;
    arow[i] = al.row(arow_ok[i] ? m0 + ml : 0);
  }
#pragma unroll
  for (int i = 0; i < B_ROWS; ++i) {
    int nl = BL::kContigMN ? (tid % BN) : (tid / BK + i * (GEMM_THREADS / BK));
    brow_ok[i] = (n0 + nl) < N;
    brow[i] = bl.row(brow_ok[i] ? n0 + nl : 0);
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    typename AL::Tile at = al.tile(k0);
    typename BL::Tile bt = bl.tile(k0);
    // stage the tiles through registers so that all global loads are in flight before any store
    float ra[A_PER], rb[B_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int kk, r;
      if (AL::kContigMN) { kk = tid / BM + i * (GEMM_THREADS / BM); r = 0; }
      else               { kk = tid % BK; r = i; }
      ra[i] = (arow_ok[r] && (k0 + kk) < k_end) ? al.load(arow[r], at, kk, k0 + kk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int kk, r;
      if (BL::kContigMN) { kk = tid / BN + i * (GEMM_THREADS / BN); r = 0; }
      else               { kk = tid % BK; r = i; }
      rb[i] = (brow_ok[r] && (k0 + kk) < k_end) ? bl.load(brow[r], bt, kk, k0 + kk) : 0.f;
    }
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      if (AL::kContigMN) As[tid / BM + i * (GEMM_THREADS / BM)][tid % BM] = ra[i];
      else               As[tid % BK][tid / BK + i * (GEMM_THREADS / BK)] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      if (BL::kContigMN) Bs[tid / BN + i * (GEMM_THREADS / BN)][tid % BN] = rb[i];
      else               Bs[tid % BK][tid / BK + i * (GEMM_THREADS / BK)] = rb[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      // micro-tile rows/cols are split in groups of 4 (128-bit shared loads, conflict-free):
      // group g of thread t covers [g*64 + t*4, +4) when the tile is 128 wide, else [t*T, +T)
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][TM == 8 ? ((i >> 2) * 64 + ty * 4 + (i & 3)) : (ty * TM + i)];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][TN == 8 ? ((j >> 2) * 64 + tx * 4 + (j & 3)) : (tx * TN + j)];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + (TM == 8 ? ((i >> 2) * 64 + ty * 4 + (i & 3)) : (ty * TM + i));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + (TN == 8 ? ((j >> 2) * 64 + tx * 4 + (j & 3)) : (tx * TN + j));
      if (n < N) ep.store(m, n, acc[i][j], split);
    }
  }
}

// ------------------------------------------------------------------ generic strided loaders
struct NoTile { __device__ NoTile() {} };

// element(mn, k) = p[mn*ld + k]  (contiguous along K)
struct LoadKContig {
  static constexpr bool kContigMN = false;
  static constexpr bool kHasTables = false;
  const float* p; long ld;
  typedef long Row; typedef NoTile Tile;
  __device__ void prepare(int*) const {}
  __device__ Row row(int mn) const { return (long)mn * ld; }
  __device__ Tile tile(int) const { return NoTile(); }
  __device__ float load(Row r, Tile, int, int k) const { return __ldg(p + r + k); }
};
// element(mn, k) = p[k*ld + mn]  (contiguous along M/N)
struct LoadMNContig {
  static constexpr bool kContigMN = true;
  static constexpr bool kHasTables = false;
  const float* p; long ld;
  typedef int Row; typedef NoTile Tile;
  __device__ void prepare(int*) const {}
  __device__ Row row(int mn) const { return mn; }
  __device__ Tile tile(int) const { return NoTile(); }
  __device__ float load(Row r, Tile, int, int k) const { return __ldg(p + (long)k * ld + r); }
};
// As LoadMNContig, plus one extra virtual row mn == ones_row whose elements are all 1.0: the
// contraction then yields the column sums of the other operand for free (db = E.sum(0)).
struct LoadMNContigOnes {
  static constexpr bool kContigMN = true;
  static constexpr bool kHasTables = false;
  const float* p; long ld; int ones_row;
  typedef int Row; typedef NoTile Tile;
  __device__ void prepare(int*) const {}
  __device__ Row row(int mn) const { return mn; }
  __device__ Tile tile(int) const { return NoTile(); }
  __device__ float load(Row r, Tile, int, int k) const { return r == ones_row ? 1.0f : __ldg(p + (long)k * ld + r); }
};

// ------------------------------------------------------------------ generic epilogues
// split-K partial: ws[split][m][n]
struct EpiPartial {
  float* ws; long MN; int N;
  __device__ void store(int m, int n, float v, int split) const { ws[(long)split * MN + (long)m * N + n] = v; }
};

// out[i] = (accumulate ? out[i] : 0) + sum_s ws[s][i], fixed order -> deterministic.
// Row `aux_row` (if >= 0) of the (rows x N) partial matrix is routed to `aux` instead (the ones-row trick).
// Element (r,c) of the partial matrix goes to out[r*out_rs + c*out_cs] (so a transposed result costs nothing).
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, float* __restrict__ aux,
                                     int rows, int N, int aux_row, int splits, int accumulate, long out_rs,
                                     long out_cs);

int launch_splitk_reduce(const float* ws, float* out, float* aux, int rows, int N, int aux_row, int splits,
                         int accumulate, long out_rs, long out_cs, cudaStream_t s);

// Choose a split count so that the grid has about two CTAs per SM, each split a multiple of BK.
inline int choose_splits(int tiles_mn, int K, int* k_per_split) {
  int target = 2 * num_sms();
  int splits = tiles_mn >= target ? 1 : (target + tiles_mn - 1) / tiles_mn;
  int max_splits = (K + 4 * GEMM_KALIGN - 1) / (4 * GEMM_KALIGN);  // at least 128 k per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int kps = (K + splits - 1) / splits;
  kps = (kps + GEMM_KALIGN - 1) / GEMM_KALIGN * GEMM_KALIGN;   // splits start on a K-chunk boundary of both cores
  splits = (K + kps - 1) / kps;
  *k_per_split = kps;
  return splits;
}

}  // namespace bf

namespace bf {

// Host-side tile selection + launch.  BM/BN are picked from the problem's M/N so that skinny
// problems (N = 10 classes, M = 32 filters) do not burn FMAs on padding.
template <class AL, class BL, class EP>
int launch_gemm_simt(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int splits, int k_per_split,
                     size_t table_bytes, cudaStream_t s) {
  if (M <= 0 || N <= 0) return BF_OK;
  const int bm = M <= 32 ? 32 : (M <= 64 ? 64 : 128);
  const int bn = N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 128));
  dim3 grid((M + bm - 1) / bm, (N + bn - 1) / bn, splits);
  if (grid.y > 65535u || grid.z > 65535u)
    return fail(BF_ERR_ARG, "gemm_simt: grid too large (%u,%u,%u)", grid.x, grid.y, grid.z);
  if (table_bytes > 180 * 1024) return fail(BF_ERR_ARG, "gemm_simt: index tables need %zu B of shared memory", table_bytes);
#define BF_GEMM_CASE(BM_, BN_)                                                                        \
  if (bm == BM_ && bn == BN_) {                                                                       \
    if (table_bytes > 24 * 1024)                                                                      \
      BF_CUDA(cudaFuncSetAttribute(gemm_simt_kernel<BM_, BN_, AL, BL, EP>,                            \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));   \
    gemm_simt_kernel<BM_, BN_, AL, BL, EP><<<grid, GEMM_THREADS, table_bytes, s>>>(al, bl, ep, M, N, K, k_per_split); \
  } else
  BF_GEMM_CASE(32, 16) BF_GEMM_CASE(32, 32) BF_GEMM_CASE(32, 64) BF_GEMM_CASE(32, 128)
  BF_GEMM_CASE(64, 16) BF_GEMM_CASE(64, 32) BF_GEMM_CASE(64, 64) BF_GEMM_CASE(64, 128)
  BF_GEMM_CASE(128, 16) BF_GEMM_CASE(128, 32) BF_GEMM_CASE(128, 64) BF_GEMM_CASE(128, 128)
  { return fail(BF_ERR_ARG, "gemm_simt: no tile for (%d,%d)", bm, bn); }
#undef BF_GEMM_CASE
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // namespace bf
