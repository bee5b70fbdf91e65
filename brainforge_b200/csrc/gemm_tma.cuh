// TMA-fed tcgen05 GEMM for operands that are both K-major in HBM:  C[m][n] = sum_k A[m][k] * B[n][k]
// (A: M x K row-major, B: N x K row-major, K % 4 == 0, 16-byte aligned).  Used where the gather-staged core
// (gemm_umma.cuh) was issue-bound on plain matrices: the LSTM gate GEMM per time step (A = Z[t], B = W^T) and the
// LSTM deltaZ GEMM (A = dgates[t], B = W as stored) -- atomic/recurrent_op.py:62,106.
//
// 128 x BN tiles, K blocks of 32 (one 128-byte swizzle row), 6-8-stage TMA ring, one MMA-issuing warp, two accumulator
// sets in TMEM so that the epilogue of a tile overlaps the MMAs of the next, persistent CTAs.  The epilogue functor is
// the same `store(m, n, value, 0)` the other contraction cores use.
#pragma once
#include "tma.cuh"

namespace bf {

constexpr int GT_THREADS = 256;     // warp 0 TMA, warp 1 MMA, warp 2 TMEM, warp 3 spare, warps 4-7 epilogue
constexpr int GT_SMEM_RING = 192 * 1024;   // ring depth = as many stages as fit: a CTA streaming its K range alone is latency-bound

struct GtParams {
  int M, N, K, nkb, m_tiles, n_tiles, ntiles, nstages;
};

template <int BN, class EP>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GtParams p, EP ep) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(2 * BN < 32 ? 32 : 2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    const bool leader = elect_one_sync();
    int s = 0, u = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
        if (leader) {
          const uint32_t st = base + (uint32_t)s * STAGE;
          mbar_expect_tx(&full[s], STAGE);
          tma_load_2d(st, &mapA, &full[s], kb * 32, mt * 128);
          tma_load_2d(st + A_BYTES, &mapB, &full[s], kb * 32, nt * BN);
        }
        if (++s == p.nstages) { s = 0; ++u; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, BN);
    int s = 0, u = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int set = it & 1, tu = it >> 1;
      if (tu > 0) mbar_wait(&t_empty[set], (uint32_t)((tu - 1) & 1));
      const uint32_t acc = tmem + (uint32_t)set * BN;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full[s], (uint32_t)(u & 1));
        tc_fence_after();
        if (leader) {
          const uint32_t st = base + (uint32_t)s * STAGE;
          const uint64_t ad = desc_k_sw128(st), bd = desc_k_sw128(st + A_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(acc, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        if (++s == p.nstages) { s = 0; ++u; }
      }
      if (leader) umma_commit(&t_full[set]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int set = it & 1, tu = it >> 1;
      const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
      const int m = mt * 128 + quad * 32 + lane;
      mbar_wait(&t_full[set], (uint32_t)(tu & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)set * BN;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = nt * BN + c0 + j;
            if (n < p.N) ep.store(m, n, __uint_as_float(v[j]), 0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN < 32 ? 32 : 2 * BN));
}

inline bool gemm_tma_ok(const float* A, const float* B, int M, int N, int K) {
  return tma_encode_fn() != nullptr && M >= 1 && N >= 1 && K >= 4 && (K & 3) == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0;
}

// C = A (M x K, lda = K) @ B^T (B: N x K, ldb = K), through `ep.store`
template <class EP>
int launch_gemm_tma(const float* A, const float* B, int M, int N, int K, EP ep, cudaStream_t s, long lda = 0, long ldb = 0) {
  if (lda == 0) lda = K;      // row strides in floats (multiples of 4: the tensor map wants 16-byte strides)
  if (ldb == 0) ldb = K;
  GtParams p;
  p.M = M; p.N = N; p.K = K; p.nkb = (K + 31) / 32; p.m_tiles = (M + 127) / 128;
  // narrower tiles when 128-wide ones would leave most SMs idle
  const int bn = (N > 64 && (long)p.m_tiles * ((N + 127) / 128) >= num_sms()) ? 128 : 64;
  p.n_tiles = (N + bn - 1) / bn;
  p.ntiles = p.m_tiles * p.n_tiles;
  CUtensorMap mA, mB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda};
    uint32_t box[2] = {32, 128};
    if (!tma_make_map(&mA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "gemm_tma: cuTensorMapEncodeTiled failed (A)");
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldb};
    uint32_t box[2] = {32, (uint32_t)bn};
    if (!tma_make_map(&mB, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "gemm_tma: cuTensorMapEncodeTiled failed (B)");
  }
  const int ctas = p.ntiles < num_sms() ? p.ntiles : num_sms();
  p.nstages = GT_SMEM_RING / (128 * 128 + bn * 128);
  if (p.nstages > 8) p.nstages = 8;
  if (bn == 128) {
    const size_t smem = 1024 + (size_t)p.nstages * (128 * 128 + 128 * 128);
    BF_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<128, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tma_kernel<128, EP><<<ctas, GT_THREADS, smem, s>>>(mA, mB, p, ep);
  } else {
    const size_t smem = 1024 + (size_t)p.nstages * (128 * 128 + 64 * 128);
    BF_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<64, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tma_kernel<64, EP><<<ctas, GT_THREADS, smem, s>>>(mA, mB, p, ep);
  }
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Contraction over the ROW index of two row-major matrices:  C[m][n] = sum_k A[k][m] * B[k][n]   (A: K x M, B: K x N)
// -- the weight gradients nabla_W = Z^T @ dgates of the LSTM (atomic/recurrent_op.py:109-110) and of a dense layer.
// Both operands are MN-major (the contraction index is the slow one): TMA writes them as SWIZZLE_128B/32-byte-atom
// boxes of 32 rows x 32 columns, kind::tf32 reads them as such (tma.cuh).  Split-K over the grid's second dimension,
// per-split partials [split][M + 1][N]; row M is the column sum of B (the bias gradient), produced for free by one
// more accumulator fed with an all-ones A operand in the CTAs of the first M tile.
constexpr int GM_THREADS = 256;
constexpr int GM_STAGES = 5;
constexpr int GM_STAGE = 2 * 4 * 4096;      // A: four 32-column groups x 32 rows x 128 B, B likewise

struct GmParams {
  int M, N, K, m_tiles, n_tiles, nkb_total, nkb_per_split;
  float* partial;
};

static __global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tma_mn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GmParams p) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[GM_STAGES], empty[GM_STAGES], done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t ones_base = base + GM_STAGES * GM_STAGE;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int nt = tile / p.m_tiles, mt = tile - nt * p.m_tiles;
  const int kb0 = split * p.nkb_per_split;
  int kb1 = kb0 + p.nkb_per_split;
  if (kb1 > p.nkb_total) kb1 = p.nkb_total;
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const bool with_ones = mt == 0;

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < GM_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t o = tid * 16u; o < 4096u; o += GM_THREADS * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ones_base + o), "r"(0x3f800000u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    const bool leader = elect_one_sync();
    int s = 0, u = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
      if (leader) {
        const uint32_t st = base + (uint32_t)s * GM_STAGE;
        mbar_expect_tx(&full[s], GM_STAGE);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          tma_load_2d(st + h * 4096, &mapA, &full[s], mt * 128 + h * 32, kb * 32);
          tma_load_2d(st + 16384 + h * 4096, &mapB, &full[s], nt * 128 + h * 32, kb * 32);
        }
      }
      if (++s == GM_STAGES) { s = 0; ++u; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, 128, true, true);
    const uint64_t od = desc_mn_sw128_32b(ones_base, 0u, 512u);          // every 32-row group of M reads the same ones
    int s = 0, u = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full[s], (uint32_t)(u & 1));
      tc_fence_after();
      if (leader) {
        const uint32_t st = base + (uint32_t)s * GM_STAGE;
        const uint64_t ad = desc_mn_sw128_32b(st, 4096u, 512u), bd = desc_mn_sw128_32b(st + 16384, 4096u, 512u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_tf32(tmem, ad + (uint64_t)(ks * 64), bd + (uint64_t)(ks * 64), idesc, (i > 0 || ks > 0) ? 1u : 0u);
          if (with_ones) umma_tf32(tmem + 128, od + (uint64_t)(ks * 64), bd + (uint64_t)(ks * 64), idesc, (i > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      if (++s == GM_STAGES) { s = 0; ++u; }
    }
    if (leader) umma_commit(&done);
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int m = mt * 128 + quad * 32 + lane;
    mbar_wait(&done, 0);
    tc_fence_after();
    float* prow = p.partial + ((size_t)split * (p.M + 1) + m) * p.N + nt * 128;
    float* brow = p.partial + ((size_t)split * (p.M + 1) + p.M) * p.N + nt * 128;
    const bool bias_writer = with_ones && quad == 0 && lane == 0;
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t v[16], w[16];
      if (nkb > 0) {
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + c0, v);
        if (with_ones) tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + 128 + c0, w);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] = 0u; w[j] = 0u; }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        if (nt * 128 + c0 + j < p.N) {       // N is a multiple of 4 (host-checked)
          if (m < p.M) *reinterpret_cast<uint4*>(prow + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (bias_writer) *reinterpret_cast<uint4*>(brow + c0 + j) = make_uint4(w[j], w[j + 1], w[j + 2], w[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

inline bool gemm_tma_mn_ok(const float* A, const float* B, int M, int N, long K) {
  return tma_encode_fn() != nullptr && M >= 1 && N >= 4 && K >= 1 && (M & 3) == 0 && (N & 3) == 0 && K < 2147483647L &&
         ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0;
}

// partial[split][M + 1][N] (row M = column sums of B); returns the number of splits through *splits_out
inline int launch_gemm_tma_mn(const float* A, const float* B, int M, int N, int K, float* partial, size_t partial_bytes,
                              int* splits_out, cudaStream_t s) {
  GmParams p;
  p.M = M; p.N = N; p.K = K; p.m_tiles = (M + 127) / 128; p.n_tiles = (N + 127) / 128;
  p.nkb_total = (K + 31) / 32;
  int splits = num_sms() / (p.m_tiles * p.n_tiles);
  if (splits < 1) splits = 1;
  if (splits > p.nkb_total) splits = p.nkb_total;
  while (splits > 1 && (size_t)splits * (M + 1) * N * 4 > partial_bytes) --splits;
  if ((size_t)splits * (M + 1) * N * 4 > partial_bytes) return fail(BF_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu (call bf_set_workspace)", (size_t)(M + 1) * N * 4, partial_bytes);
  p.nkb_per_split = (p.nkb_total + splits - 1) / splits;
  p.partial = partial;
  CUtensorMap mA, mB;
  {
    uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    uint64_t str[1] = {(uint64_t)M};
    uint32_t box[2] = {32, 32};
    if (!tma_make_map(&mA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return fail(BF_ERR_CUDA, "gemm_tma_mn: cuTensorMapEncodeTiled failed (A)");
  }
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    uint64_t str[1] = {(uint64_t)N};
    uint32_t box[2] = {32, 32};
    if (!tma_make_map(&mB, B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return fail(BF_ERR_CUDA, "gemm_tma_mn: cuTensorMapEncodeTiled failed (B)");
  }
  const size_t smem = 1024 + (size_t)GM_STAGES * GM_STAGE + 4096;
  BF_CUDA(cudaFuncSetAttribute(gemm_tma_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_tma_mn_kernel<<<dim3(p.m_tiles * p.n_tiles, splits), GM_THREADS, smem, s>>>(mA, mB, p);
  BF_LAUNCH_CHECK();
  *splits_out = splits;
  set_last_path(2);
  return BF_OK;
}

}  // namespace bf
