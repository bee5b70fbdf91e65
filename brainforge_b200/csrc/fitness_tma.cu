// Neuro-evolution fitness fan-out on TMA + tcgen05 (evolution/_evolution.py:83-97, learner/neuroevolution.py:20-22,34-37):
// for P genomes the first layers are ONE contraction  H = X (N x nin) @ W1all (nin x P*nh)  on the shared input X.
//
//   1. relayout pre-pass: the genome matrix (P, loci) cannot be a TMA operand in place (loci*4 bytes is not a multiple
//      of 16), so W1 of every genome is written once, transposed, as Bt[p*64 + h][k] = (g - 0.5)*20 (as_weights,
//      neuroevolution.py:20-22), hidden units padded to 64, TF32-rounded: a (P*64 x nin) K-major B operand.
//   2. GEMM: CTA tile 256 samples x 128 columns (= 2 genomes), K blocks of 32; A = X through a K-major SWIZZLE_128B
//      map, B likewise (MN-major 32-byte-atom boxes measured 3x slower to load); 4-stage TMA ring; two MMA-issuing warps (one per 128-row half);
//      two accumulator sets in TMEM so that the epilogue of one tile overlaps the MMAs of the next.
//   3. epilogue = the rest of the network: +b1, sigmoid, the (nh x nout) second layer from shared memory, log-softmax,
//      -sum(Y * log p) per sample, summed over the tile's rows; per-tile partial sums go to a small buffer that a
//      last kernel adds in a fixed order (deterministic) and divides by N.  The hidden activations never reach HBM.
#include "tma.cuh"
#include <stdlib.h>

namespace bf {

constexpr int FT_THREADS = 640;      // warp 0 TMA, warps 1-2 MMA (row half 0 / 1), warp 3 TMEM, warps 4-19: four epilogue groups = (row half, genome of the tile)
constexpr int FT_STAGES = 4;
constexpr int FT_A_BYTES = 2 * 128 * 128;     // two 128-row halves x 32 k floats
constexpr int FT_B_BYTES = 128 * 128;         // 128 columns (2 genomes x 64 hidden) x 32 k floats, K-major
constexpr int FT_STAGE_BYTES = FT_A_BYTES + FT_B_BYTES;
constexpr int FT_HP = 64;                     // hidden units per genome, padded
constexpr int FT_MAX_OUT = 16;
constexpr int FT_PARAM_F = FT_HP + FT_HP * FT_MAX_OUT + FT_MAX_OUT;   // b1 | W2 [h][16] | b2 per genome

struct FtParams {
  int P, N, nin, nh, nout, nkb, m_passes, n_tiles, ntiles;
  long loci;            // floats between consecutive genome rows of G (>= the number of loci: padded stores allowed)
  const float* G;
  const float* Y;
  float* partial;       // [P][m_passes * 8]
  const int* inds;      // optional: slot j evaluates genome inds[j] (a subset / permutation of the store); null: j
  int bt_by_genome;     // the K-major operand is the RESIDENT one (64 rows per genome id) instead of this call's scratch (per slot)
};
__device__ __forceinline__ int ft_genome(const FtParams& p, int slot) {
  if (slot >= p.P) slot = p.P - 1;          // the padding slot of an odd count re-reads the last genome (result dropped)
  return p.inds ? __ldg(p.inds + slot) : slot;
}

// Bt[p*64 + h][k] = (G[p][k*nh + h] - 0.5) * 20, zero padding for h >= nh and p >= P: every genome's (nin x nh) first
// layer transposed through a padded shared-memory tile, so that the GEMM's B operand is K-major like A
// (a variant moving 112-row chunks per block with per-element index divisions measured 1.5x slower)
// `inds` (optional): block z handles genome inds[z]; `by_genome`: its rows go to the genome's own place in a resident
// operand (Bt[g*64 + h]) instead of slot z of a scratch one.
__global__ void fitness_relayout_kernel(const float* __restrict__ G, float* __restrict__ Bt, int P, int nin, int nh, long loci,
                                        const int* __restrict__ inds, int by_genome) {
  __shared__ float tile[32][33];
  const int slot = blockIdx.z, k0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  const int p = (slot < P && inds) ? __ldg(inds + slot) : slot;
  const int prow = by_genome ? p : slot;
  if (by_genome && slot >= P) return;
  const int tx = threadIdx.x, ty = threadIdx.y;     // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int k = k0 + ty + j, h = h0 + tx;
    float v = 0.f;
    if (slot < P && k < nin && h < nh) v = (__ldg(G + (size_t)p * loci + (size_t)k * nh + h) - 0.5f) * 20.f;
    tile[ty + j][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int h = h0 + ty + j, k = k0 + tx;
    if (k < nin) Bt[((size_t)prow * FT_HP + h) * nin + k] = __uint_as_float(to_tf32(tile[tx][ty + j]));
  }
}

template <int NQ>      // float4 groups of output classes actually computed: nout <= 4*NQ
__global__ void __launch_bounds__(FT_THREADS, 1)
fitness_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, FtParams p) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[FT_STAGES], empty[FT_STAGES], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sparam[2][2][FT_PARAM_F];      // [row half / epilogue group][genome of the tile]
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;

  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < FT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 2); mbar_init(&t_empty[i], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * 2 * FT_PARAM_F; i += FT_THREADS) (&sparam[0][0][0])[i] = 0.f;   // padding entries stay zero
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one_sync();
    int g = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int nt = tile / p.m_passes, mp = tile - nt * p.m_passes;       // consecutive tiles share the genome columns
      const int r0 = (p.bt_by_genome ? ft_genome(p, nt * 2) : nt * 2) * FT_HP;
      const int r1 = (p.bt_by_genome ? ft_genome(p, nt * 2 + 1) : nt * 2 + 1) * FT_HP;
      for (int kb = 0; kb < p.nkb; ++kb, ++g) {
        const int s = g % FT_STAGES, u = g / FT_STAGES;
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
        if (leader) {
          const uint32_t st = base + (uint32_t)s * FT_STAGE_BYTES;
          mbar_expect_tx(&full[s], FT_STAGE_BYTES);
          tma_load_2d(st, &mapA, &full[s], kb * 32, mp * 256);
          tma_load_2d(st + 128 * 128, &mapA, &full[s], kb * 32, mp * 256 + 128);
          tma_load_2d(st + FT_A_BYTES, &mapB, &full[s], kb * 32, r0);                     // 64 hidden rows of genome 0 of the tile
          tma_load_2d(st + FT_A_BYTES + FT_HP * 128, &mapB, &full[s], kb * 32, r1);       // ... of genome 1
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: warp w owns the rows 128*w.. of the tile
    const bool leader = elect_one_sync();
    const int half = warp - 1;
    constexpr uint32_t idesc = idesc_tf32(128, 128);
    int g = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int set = it & 1, tu = it >> 1;
      if (tu > 0) mbar_wait(&t_empty[set], (uint32_t)((tu - 1) & 1));
      const uint32_t acc = tmem + (uint32_t)(set * 2 + half) * 128u;
      for (int kb = 0; kb < p.nkb; ++kb, ++g) {
        const int s = g % FT_STAGES, u = g / FT_STAGES;
        mbar_wait(&full[s], (uint32_t)(u & 1));
        tc_fence_after();
        if (leader) {
          const uint32_t st = base + (uint32_t)s * FT_STAGE_BYTES;
          const uint64_t ad = desc_k_sw128(st + (uint32_t)half * 128u * 128u);
          const uint64_t bd = desc_k_sw128(st + FT_A_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(acc, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
      }
      if (leader) umma_commit(&t_full[set]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: the rest of the network, per sample row
    // group = (row half, genome gi of the tile): 128 threads, one TMEM lane quadrant per warp
    const int grp = (warp - 4) >> 2, half = grp & 1, gi = grp >> 1, quad = warp & 3, gt = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int set = it & 1, tu = it >> 1;
      const int nt = tile / p.m_passes, mp = tile - nt * p.m_passes;
      const int pg = nt * 2 + gi;                        // slot in this call's list; its genome row in the store:
      const size_t grow = (size_t)ft_genome(p, pg) * p.loci;
      // second-layer parameters of the group's genome -> shared memory (as_weights applied), while the MMAs run
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");       // previous tile's readers are done
      // (b1 | W2 | b2 are one contiguous run of nh + nh*nout + nout loci behind W1: coalesced reads, scattered into the
      // padded [b1 64 | W2 64x16 | b2 16] layout)
      const int ntail = p.nh + p.nh * p.nout + p.nout;
      constexpr int PF = (FT_PARAM_F + 127) / 128;       // all loads first (independent), then the scatter
      float pv[PF];
#pragma unroll
      for (int q = 0; q < PF; ++q) {
        const int j = gt + q * 128;
        pv[q] = (j < ntail && pg < p.P) ? __ldg(p.G + grow + (size_t)p.nin * p.nh + j) : 0.5f;
      }
#pragma unroll
      for (int q = 0; q < PF; ++q) {
        const int j = gt + q * 128;
        if (j < ntail) {
          int dst;
          if (j < p.nh) dst = j;
          else if (j < p.nh + p.nh * p.nout) { const int t = j - p.nh, h = t / p.nout; dst = FT_HP + h * FT_MAX_OUT + (t - h * p.nout); }
          else dst = FT_HP + FT_HP * FT_MAX_OUT + (j - p.nh - p.nh * p.nout);
          sparam[half][gi][dst] = (pv[q] - 0.5f) * 20.f;
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      const int row = mp * 256 + half * 128 + quad * 32 + lane;
      const bool rok = row < p.N;
      constexpr int NO = 4 * NQ;
      float yv[NO];
#pragma unroll
      for (int c = 0; c < NO; ++c) yv[c] = (rok && c < p.nout) ? __ldg(p.Y + (size_t)row * p.nout + c) : 0.f;
      mbar_wait(&t_full[set], (uint32_t)(tu & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * 2 + half) * 128u;
      {
        const float* pb1 = sparam[half][gi];
        const float* pw2 = pb1 + FT_HP;
        const float* pb2 = pw2 + FT_HP * FT_MAX_OUT;
        float logit[NO];
#pragma unroll
        for (int c = 0; c < NO; ++c) logit[c] = pb2[c];
#pragma unroll
        for (int c0 = 0; c0 < FT_HP; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + gi * FT_HP + c0, v);
          tmem_ld_wait();
          float act[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)     // sigmoid (atomic/activation_op.py:30); branch-free reciprocal so that the 16 chains interleave
            act[j] = __fdividef(1.0f, 1.0f + __expf(-(__uint_as_float(v[j]) + pb1[c0 + j])));
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int h = c0 + j;
            const float a = act[j];
            const float4* w4 = reinterpret_cast<const float4*>(pw2 + h * FT_MAX_OUT);        // zero rows for h >= nh
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const float4 w = w4[q];
              logit[4 * q + 0] = fmaf(a, w.x, logit[4 * q + 0]); logit[4 * q + 1] = fmaf(a, w.y, logit[4 * q + 1]);
              logit[4 * q + 2] = fmaf(a, w.z, logit[4 * q + 2]); logit[4 * q + 3] = fmaf(a, w.w, logit[4 * q + 3]);
            }
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < NO; ++c) if (c < p.nout) mx = fmaxf(mx, logit[c]);
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < NO; ++c) if (c < p.nout) se += expf(logit[c] - mx);
        const float lse = mx + logf(se);
        float loss = 0.f;
#pragma unroll
        for (int c = 0; c < NO; ++c) if (c < p.nout && yv[c] != 0.f) loss -= yv[c] * (logit[c] - lse);
        loss = warp_sum(rok ? loss : 0.f);
        if (lane == 0 && pg < p.P) p.partial[(size_t)pg * (p.m_passes * 8) + mp * 8 + half * 4 + quad] = loss;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// fitness[p] = (sum of the genome's partial sums, fixed order) / N
__global__ void fitness_finish_kernel(const float* __restrict__ partial, float* __restrict__ fitness, int P, int slots, int N) {
  const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= P) return;
  double s = 0.0;
  for (int i = 0; i < slots; ++i) s += (double)partial[(size_t)pidx * slots + i];
  fitness[pidx] = (float)(s / (double)N);
}

bool fitness_tma_supported(int P, int N, int nin, int nh, int nout) {
  return tma_encode_fn() != nullptr && nh <= FT_HP && nout <= FT_MAX_OUT && (nin % 4) == 0 && P >= 1 && N >= 1;
}

int fitness_tma(const float* genomes, long ldg, const float* bt_resident, const int* inds, const float* X, const float* Y,
                float* fitness, int P, int N, int nin, int nh, int nout, cudaStream_t s) {
  const int Ppad = (P + 1) & ~1;
  FtParams p;
  p.P = P; p.N = N; p.nin = nin; p.nh = nh; p.nout = nout; p.loci = ldg;
  p.nkb = (nin + 31) / 32;
  p.m_passes = (N + 255) / 256;
  p.n_tiles = Ppad / 2;
  p.ntiles = p.n_tiles * p.m_passes;
  p.G = genomes; p.Y = Y; p.inds = inds; p.bt_by_genome = bt_resident ? 1 : 0;
  const size_t bt_bytes = bt_resident ? 0 : (((size_t)nin * Ppad * FT_HP * 4 + 255) & ~(size_t)255);
  const size_t part_bytes = (size_t)P * p.m_passes * 8 * 4;
  void* ws = nullptr;
  int rc = workspace_require(bt_bytes + part_bytes, &ws);
  if (rc) return rc;
  const float* Bt = bt_resident ? bt_resident : (const float*)ws;
  p.partial = (float*)((char*)ws + bt_bytes);
  BF_REQUIRE(((uintptr_t)X & 15) == 0 && ((uintptr_t)Bt & 15) == 0, BF_ERR_ARG, "bf_fitness_mlp: operand alignment");
  BF_REQUIRE(Ppad <= 65535, BF_ERR_ARG, "bf_fitness_mlp: more than 65534 genomes per call");
  if (!bt_resident) {
    fitness_relayout_kernel<<<dim3((nin + 31) / 32, FT_HP / 32, Ppad), dim3(32, 8), 0, s>>>(genomes, (float*)ws, P, nin, nh, ldg, inds, 0);
    BF_LAUNCH_CHECK();
  }
  CUtensorMap mA, mB;
  {
    uint64_t dims[2] = {(uint64_t)nin, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)nin};
    uint32_t box[2] = {32, 128};
    if (!tma_make_map(&mA, X, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return fail(BF_ERR_CUDA, "bf_fitness_mlp: cuTensorMapEncodeTiled failed (X)");
  }
  {
    // rows beyond the tensor are never addressed (slots are clamped), so the extent only has to cover the operand
    uint64_t dims[2] = {(uint64_t)nin, (uint64_t)2147483647ull / (uint64_t)(nin > 0 ? nin : 1) / 8ull * 8ull};
    if (!bt_resident) dims[1] = (uint64_t)Ppad * FT_HP;
    uint64_t str[1] = {(uint64_t)nin};
    uint32_t box[2] = {32, FT_HP};
    if (!tma_make_map(&mB, Bt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, false))
      return fail(BF_ERR_CUDA, "bf_fitness_mlp: cuTensorMapEncodeTiled failed (genome operand)");
  }
  const size_t smem = 1024 + (size_t)FT_STAGES * FT_STAGE_BYTES;
  const int ctas = p.ntiles < num_sms() ? p.ntiles : num_sms();
  const int nq = (nout + 3) / 4;
#define BF_FT_CASE(NQ_)                                                                                              \
  if (nq == NQ_) {                                                                                                   \
    BF_CUDA(cudaFuncSetAttribute(fitness_tma_kernel<NQ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    fitness_tma_kernel<NQ_><<<ctas, FT_THREADS, smem, s>>>(mA, mB, p);                                               \
  }
  BF_FT_CASE(1) BF_FT_CASE(2) BF_FT_CASE(3) BF_FT_CASE(4)
#undef BF_FT_CASE
  BF_LAUNCH_CHECK();
  fitness_finish_kernel<<<(P + 127) / 128, 128, 0, s>>>(p.partial, fitness, P, p.m_passes * 8, N);
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}

// (re)build the resident K-major first-layer operand for the listed genomes (all P when inds is null)
int fitness_relayout_resident(const float* genomes, long ldg, float* bt, const int* inds, int n, int nin, int nh, cudaStream_t s) {
  if (n <= 0) return BF_OK;
  BF_REQUIRE(n <= 65535, BF_ERR_ARG, "bf_population_relayout: more than 65535 genomes per call");
  fitness_relayout_kernel<<<dim3((nin + 31) / 32, FT_HP / 32, n), dim3(32, 8), 0, s>>>(genomes, bt, n, nin, nh, ldg, inds, 1);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // namespace bf
