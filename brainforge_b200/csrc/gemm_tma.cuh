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
int launch_gemm_tma(const float* A, const float* B, int M, int N, int K, EP ep, cudaStream_t s) {
  GtParams p;
  p.M = M; p.N = N; p.K = K; p.nkb = (K + 31) / 32; p.m_tiles = (M + 127) / 128;
  // narrower tiles when 128-wide ones would leave most SMs idle
  const int bn = (N > 64 && (long)p.m_tiles * ((N + 127) / 128) >= num_sms()) ? 128 : 64;
  p.n_tiles = (N + bn - 1) / bn;
  p.ntiles = p.m_tiles * p.n_tiles;
  CUtensorMap mA, mB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)K};
    uint32_t box[2] = {32, 128};
    if (!tma_make_map(&mA, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "gemm_tma: cuTensorMapEncodeTiled failed (A)");
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K};
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

}  // namespace bf
