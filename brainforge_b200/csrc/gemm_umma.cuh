// Tensor-core contraction core (the BF_PREC_TF32 path): C[M,N] = sum_k A(m,k) * B(k,n) on the 5th-gen
// tensor cores -- tcgen05.mma.cta_group::1.kind::tf32, 128 x BN accumulator tile in TMEM, FP32 accumulate.
//
// It takes the SAME loader / epilogue functors as the SIMT core (gemm_simt.cuh), so every contraction of
// the hot path (dense fwd/dW/dX, implicit-GEMM conv fwd/dX/dF, LSTM gate GEMMs, the neuro-evolution layer)
// gets the tensor pipe without an im2col buffer or a transposed copy: 256 threads gather a 128x32 A tile and
// a BNx32 B tile straight from HBM/L2 (coalesced along whichever axis the loader says is contiguous), round
// to TF32 (cvt.rna, instead of the truncation the MMA would apply to raw FP32 bits) and store them into the
// canonical K-major SWIZZLE_128B shared-memory layout; one elected thread then issues four K=8 MMAs per
// 32-wide chunk.  A 3-stage ring of tiles, released by tcgen05.commit -> mbarrier, overlaps the gather of
// chunk i+1 with the MMAs of chunk i.  The epilogue reads the accumulator with tcgen05.ld (32x32b.x8),
// one TMEM lane = one output row per thread.
//
// Shared-memory operand layout (per stage), K-major, 128-byte swizzle (cute Layout_K_SW128_Atom<tf32>):
//   row r (an M or N index) occupies bytes [r*128, r*128+128); its 16-byte chunk j is stored at chunk
//   position j ^ (r & 7); 8-row groups are 1024 B apart (SBO = 1024), tiles are 1024-byte aligned.
#pragma once
#include "gemm_simt.cuh"

namespace bf {

constexpr int UMMA_BM = 128;
constexpr int UMMA_BK = 32;      // floats per K chunk = one 128-byte swizzle row
constexpr int UMMA_STAGES = 3;
constexpr int UMMA_THREADS = 256;
constexpr unsigned UMMA_SPIN_LIMIT = 200u * 1000u * 1000u;   // a wedged pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > UMMA_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
// K-major SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 | SBO(1024>>4)<<32
// | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

template <int BN>
struct UmmaCfg {
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;                       // power of two >= 32
  static constexpr int kStageBytes = UMMA_BM * 128 + BN * 128;
  static constexpr int kTileBytes = UMMA_STAGES * kStageBytes;
};

template <int BN, class AL, class BL, class EP>
__global__ void __launch_bounds__(UMMA_THREADS)
gemm_umma_kernel(AL al, BL bl, EP ep, int M, int N, int K, int k_per_split) {
  constexpr int BM = UMMA_BM, BK = UMMA_BK, S = UMMA_STAGES;
  using Cfg = UmmaCfg<BN>;
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t empty_bar[S];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // dynamic smem: [pad to 1024][S x (A tile | B tile)][index tables of the loaders]
  const uint32_t dyn_base = smem_u32(dyn_smem);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  int* tables = dyn_smem + ((tile_base - dyn_base) + Cfg::kTileBytes) / 4;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "n"(Cfg::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  al.prepare(tables);
  bl.prepare(tables);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;

  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int split = blockIdx.z;
  const int k_begin = split * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int nchunks = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

  // ---- thread -> element mapping of the gathers
  //  MN-contiguous loader: row = tid % ROWS (consecutive lanes -> consecutive rows, coalesced), the thread
  //    owns whole 16-byte K chunks (4 consecutive k) so that its shared stores are conflict-free 128-bit
  //  K-contiguous loader:  k = tid % 32 (a warp reads one 128-byte row segment), rows tid/32 + 8*i
  constexpr int A_PER = BM * BK / UMMA_THREADS;   // 16
  constexpr int B_PER = BN * BK / UMMA_THREADS;   // BN / 8
  constexpr int A_ROWS = AL::kContigMN ? 1 : A_PER;
  constexpr int B_ROWS = BL::kContigMN ? 1 : B_PER;
  typename AL::Row arow[A_ROWS];
  typename BL::Row brow[B_ROWS];
  bool arow_ok[A_ROWS], brow_ok[B_ROWS];
#pragma unroll
  for (int i = 0; i < A_ROWS; ++i) {
    int ml = AL::kContigMN ? (tid % BM) : (tid / BK + i * (UMMA_THREADS / BK));
    arow_ok[i] = (m0 + ml) < M;
    arow[i] = al.row(arow_ok[i] ? m0 + ml : 0);
  }
  constexpr int B_UNITS = (BN * 8 + UMMA_THREADS - 1) / UMMA_THREADS;   // 16-byte chunk units per thread (MN-contig B)
#pragma unroll
  for (int i = 0; i < B_ROWS; ++i) {
    int nl = BL::kContigMN ? (tid % BN) : (tid / BK + i * (UMMA_THREADS / BK));
    brow_ok[i] = (n0 + nl) < N;
    brow[i] = bl.row(brow_ok[i] ? n0 + nl : 0);
  }
  constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);

  for (int i = 0; i < nchunks; ++i) {
    const int s = i % S, u = i / S;
    const int k0 = k_begin + i * BK;
    typename AL::Tile at = al.tile(k0);
    typename BL::Tile bt = bl.tile(k0);
    uint32_t ra[A_PER], rb[BL::kContigMN ? B_UNITS * 4 : B_PER];
    // ---- gather (global -> registers); all loads are issued before anything waits
    if constexpr (AL::kContigMN) {
      // chunks c = tid/BM + 2*j (j = 0..3), 4 k each
#pragma unroll
      for (int j = 0; j < A_PER / 4; ++j) {
        int c = tid / BM + j * (UMMA_THREADS / BM);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int kk = c * 4 + e;
          float v = (arow_ok[0] && (k0 + kk) < k_end) ? al.load(arow[0], at, kk, k0 + kk) : 0.f;
          ra[j * 4 + e] = to_tf32(v);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < A_PER; ++r) {
        int kk = tid % BK;
        float v = (arow_ok[r] && (k0 + kk) < k_end) ? al.load(arow[r], at, kk, k0 + kk) : 0.f;
        ra[r] = to_tf32(v);
      }
    }
    if constexpr (BL::kContigMN) {
      // unit = tid + j*256 -> (row = unit % BN = tid % BN, chunk = unit / BN); BN*8 units in the tile
#pragma unroll
      for (int j = 0; j < B_UNITS; ++j) {
        int unit = tid + j * UMMA_THREADS;
        int c = unit / BN;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int kk = c * 4 + e;
          float v = (unit < BN * 8 && brow_ok[0] && (k0 + kk) < k_end) ? bl.load(brow[0], bt, kk, k0 + kk) : 0.f;
          rb[j * 4 + e] = to_tf32(v);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < B_PER; ++r) {
        int kk = tid % BK;
        float v = (brow_ok[r] && (k0 + kk) < k_end) ? bl.load(brow[r], bt, kk, k0 + kk) : 0.f;
        rb[r] = to_tf32(v);
      }
    }
    // ---- the MMAs that read this stage S chunks ago must have drained
    if (u > 0) mbar_wait(&empty_bar[s], (uint32_t)((u - 1) & 1));
    const uint32_t a_tile = tile_base + s * Cfg::kStageBytes;
    const uint32_t b_tile = a_tile + BM * 128;
    if constexpr (AL::kContigMN) {
      int r = tid % BM;
#pragma unroll
      for (int j = 0; j < A_PER / 4; ++j) {
        int c = tid / BM + j * (UMMA_THREADS / BM);
        uint32_t addr = a_tile + r * 128 + ((c ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ra[j * 4]), "r"(ra[j * 4 + 1]),
                     "r"(ra[j * 4 + 2]), "r"(ra[j * 4 + 3]) : "memory");
      }
    } else {
      int kk = tid % BK;
#pragma unroll
      for (int q = 0; q < A_PER; ++q) {
        int r = tid / BK + q * (UMMA_THREADS / BK);
        uint32_t addr = a_tile + r * 128 + (((kk >> 2) ^ (r & 7)) << 4) + ((kk & 3) << 2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(ra[q]) : "memory");
      }
    }
    if constexpr (BL::kContigMN) {
      int r = tid % BN;
#pragma unroll
      for (int j = 0; j < B_UNITS; ++j) {
        int unit = tid + j * UMMA_THREADS;
        int c = unit / BN;
        uint32_t addr = b_tile + r * 128 + ((c ^ (r & 7)) << 4);
        if (unit < BN * 8)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(rb[j * 4]), "r"(rb[j * 4 + 1]),
                       "r"(rb[j * 4 + 2]), "r"(rb[j * 4 + 3]) : "memory");
      }
    } else {
      int kk = tid % BK;
#pragma unroll
      for (int q = 0; q < B_PER; ++q) {
        int r = tid / BK + q * (UMMA_THREADS / BK);
        uint32_t addr = b_tile + r * 128 + (((kk >> 2) ^ (r & 7)) << 4) + ((kk & 3) << 2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(rb[q]) : "memory");
      }
    }
    // generic-proxy writes -> visible to the async proxy (the tensor core reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t adesc = umma_desc_k_sw128(a_tile), bdesc = umma_desc_k_sw128(b_tile);
#pragma unroll
      for (int k = 0; k < BK / 8; ++k)   // UMMA_K = 8 for tf32 (32 bytes): advance the start address inside the atom
        umma_tf32(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i > 0 || k > 0) ? 1u : 0u);
      umma_commit(&empty_bar[s]);                  // stage s reusable once these MMAs have read it
      if (i == nchunks - 1) umma_commit(&done_bar);  // accumulator complete
    }
  }

  // ---- epilogue: TMEM -> registers -> functor.  Warp w owns TMEM lanes 32*(w%4)..+31 (hardware rule);
  //      warps 0-3 take the low half of the columns, warps 4-7 the high half.
  if (nchunks > 0) {
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    constexpr int HALF = BN / 2;
#pragma unroll
    for (int c0 = 0; c0 < HALF; c0 += 8) {
      const int col = half * HALF + c0;
      uint32_t v[8];
      if (nchunks > 0) {
        uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0u;
      }
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int n = n0 + col + j;
          if (n < N) ep.store(m, n, __uint_as_float(v[j]), split);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(Cfg::kTmemCols));
  }
}

template <class AL, class BL, class EP>
int launch_gemm_umma(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int splits, int k_per_split,
                     size_t table_bytes, cudaStream_t s) {
  if (M <= 0 || N <= 0) return BF_OK;
  const int bn = N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 128));
  dim3 grid((M + UMMA_BM - 1) / UMMA_BM, (N + bn - 1) / bn, splits);
  if (grid.y > 65535u || grid.z > 65535u)
    return fail(BF_ERR_ARG, "gemm_umma: grid too large (%u,%u,%u)", grid.x, grid.y, grid.z);
#define BF_UMMA_CASE(BN_)                                                                                    \
  if (bn == BN_) {                                                                                           \
    size_t smem = 1024 + UmmaCfg<BN_>::kTileBytes + table_bytes;                                             \
    if (smem > 200 * 1024) return fail(BF_ERR_ARG, "gemm_umma: needs %zu B of shared memory", smem);         \
    BF_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<BN_, AL, BL, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem));                                                                \
    gemm_umma_kernel<BN_, AL, BL, EP><<<grid, UMMA_THREADS, smem, s>>>(al, bl, ep, M, N, K, k_per_split);    \
  } else
  BF_UMMA_CASE(16) BF_UMMA_CASE(32) BF_UMMA_CASE(64) BF_UMMA_CASE(128)
  { return fail(BF_ERR_ARG, "gemm_umma: no tile for N=%d", N); }
#undef BF_UMMA_CASE
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// One entry point for every contraction of the library: picks the arithmetic path.
template <class AL, class BL, class EP>
int launch_gemm(int precision, const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int splits,
                int k_per_split, size_t table_bytes, cudaStream_t s) {
  if (precision == BF_PREC_TF32) {
    set_last_path(1);
    return launch_gemm_umma(al, bl, ep, M, N, K, splits, k_per_split, table_bytes, s);
  }
  set_last_path(0);
  return launch_gemm_simt(al, bl, ep, M, N, K, splits, k_per_split, table_bytes, s);
}

}  // namespace bf
