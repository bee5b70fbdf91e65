// Device-resident Population: the genome matrix of evolution/_evolution.py stays in HBM between
// generations, together with the K-major first-layer operand the fitness kernel contracts against, so that neither the
// 782 MB genome upload nor the 0.6 ms re-layout pre-pass of bf_fitness_mlp recurs per `Population.update`.
//
//   store:   G  (P, ldg) FP32, ldg = loci rounded up to 4 floats (16-byte rows), values in [0, 1)
//            Bt (P*64, nin) FP32: Bt[g*64 + h][k] = tf32((G[g][k*nh + h] - 0.5) * 20)  (as_weights, neuroevolution.py:20-22)
//   update:  bf_population_fitness   -- the fitness fan-out of _evolution.py:83-97 for a list of individuals, read in place
//   epoch:   bf_population_next_generation -- `np.where(survmask, individuals, candidates)` + mating + mutation
//            (_evolution.py:99-134,181-183,245-261) in ONE pass over the store: read parents / self, write the next
//            generation, flag the mutated rows; then bf_population_relayout refreshes Bt for the rows that changed.
// Random draws come from Philox4x32-10 keyed by (seed, generation) and counted by (individual, locus group), or -- for
// parity tests against the reference with its RNG draws recorded -- from caller-supplied arrays.
#include "common.cuh"

namespace bf {

int fitness_tma(const float* genomes, long ldg, const float* bt_resident, const int* inds, const float* X, const float* Y,
                float* fitness, int P, int N, int nin, int nh, int nout, cudaStream_t s);
int fitness_relayout_resident(const float* genomes, long ldg, float* bt, const int* inds, int n, int nin, int nh, cudaStream_t s);
bool fitness_tma_supported(int P, int N, int nin, int nh, int nout);

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: no state in memory, any (individual, locus) addressable
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

struct GaParams {
  const float* old;
  float* next;
  long ldg;
  int loci, P;
  const unsigned char* survive;      // [P] 1: the individual is kept (survmask, _evolution.py:136-141)
  const int* left;                   // [P] parents of candidate i (get_candidates, _evolution.py:99-134)
  const int* right;
  float rate;                        // per-locus mutation probability (already divided by loci, _evolution.py:176)
  uint32_t mut_threshold;            // rate * 2^32
  uint32_t seed, generation;
  const float* coin;                 // injected draws (tests): (P, loci) each, or null -> Philox
  const float* mut_u;
  const float* mut_val;
  unsigned char* mutant;             // [P] out: 1 where any locus of the row mutated
};

// one thread = 4 consecutive loci of one individual
__global__ void __launch_bounds__(256) ga_next_generation_kernel(GaParams p) {
  const int groups = (int)(p.ldg >> 2);
  const long total = (long)p.P * groups;
  const Philox rng{p.seed, p.generation};
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int i = (int)(t / groups), g = (int)(t - (long)i * groups), l0 = g * 4;
    float v[4];
    const bool keep = p.survive[i] != 0;
    if (keep) {
      const float4 s = *reinterpret_cast<const float4*>(p.old + (size_t)i * p.ldg + l0);
      v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w;
    } else {
      // _default_mate_function (_evolution.py:245-247): np.where(uniform(size) < 0.5, gen1, gen2)
      const float4 a = *reinterpret_cast<const float4*>(p.old + (size_t)p.left[i] * p.ldg + l0);
      const float4 b = *reinterpret_cast<const float4*>(p.old + (size_t)p.right[i] * p.ldg + l0);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
      if (p.coin) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (l0 + e < p.loci && p.coin[(size_t)i * p.loci + l0 + e] < 0.5f) ? av[e] : bv[e];
      } else {
        const uint32_t bits = rng((uint32_t)i, (uint32_t)g, 0u, 0u).x;
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = ((bits >> e) & 1u) ? av[e] : bv[e];
      }
    }
    // _default_mutate_function (_evolution.py:253-261): every locus of every row mutates with probability `rate`
    if (p.rate > 0.f) {
      bool any = false;
      if (p.mut_u) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (l0 + e < p.loci && p.mut_u[(size_t)i * p.loci + l0 + e] < p.rate) { v[e] = p.mut_val[(size_t)i * p.loci + l0 + e]; any = true; }
      } else {
        const uint4 u = rng((uint32_t)i, (uint32_t)g, 1u, 0u);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
        if (uu[0] < p.mut_threshold || uu[1] < p.mut_threshold || uu[2] < p.mut_threshold || uu[3] < p.mut_threshold) {
          const uint4 w = rng((uint32_t)i, (uint32_t)g, 2u, 0u);
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (l0 + e < p.loci && uu[e] < p.mut_threshold) { v[e] = u01(ww[e]); any = true; }
        }
      }
      if (any) p.mutant[i] = 1;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) if (l0 + e >= p.loci) v[e] = 0.f;          // row padding stays zero
    *reinterpret_cast<float4*>(p.next + (size_t)i * p.ldg + l0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// dst[r][:] = src[idx[r]][:]  (rows of `row_floats`; the shuffled mini-batches of util/_nnutil.py:4-15 taken from a
// data set that stays in HBM, and single genomes of the store)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                          float* __restrict__ dst, int m, size_t row_floats, int vec) {
  for (int r = blockIdx.x; r < m; r += gridDim.x) {
    const float* s = src + (size_t)__ldg(idx + r) * row_floats;
    float* d = dst + (size_t)r * row_floats;
    if (vec) {
      const size_t n4 = row_floats >> 2;
      for (size_t j = threadIdx.x; j < n4; j += blockDim.x)
        reinterpret_cast<float4*>(d)[j] = __ldg(reinterpret_cast<const float4*>(s) + j);
    } else {
      for (size_t j = threadIdx.x; j < row_floats; j += blockDim.x) d[j] = __ldg(s + j);
    }
  }
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_gather_rows(const float* src, const int* idx, float* dst, int m, size_t row_floats, void* stream) {
  BF_REQUIRE(m >= 0, BF_ERR_ARG, "bf_gather_rows: bad row count %d", m);
  if (m == 0 || row_floats == 0) return BF_OK;
  BF_REQUIRE(src && idx && dst, BF_ERR_ARG, "bf_gather_rows: null pointer");
  const int vec = ((row_floats & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) ? 1 : 0;
  // small rows: several rows per CTA would be better, but the callers' rows are >= 3 KB (images, sequences, genomes)
  int grid = m < num_sms() * 8 ? m : num_sms() * 8;
  gather_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, idx, dst, m, row_floats, vec);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_population_fitness(const float* G, long ldg, int Pstore, const float* Bt, const int* inds, int n, const float* X,
                          const float* Y, float* fitness, int N, int nin, int nh, int nout, int precision, void* stream) {
  BF_REQUIRE(G && Bt && X && Y && fitness, BF_ERR_ARG, "bf_population_fitness: null pointer");
  BF_REQUIRE(n >= 0 && n <= Pstore && N > 0 && nin > 0 && nh > 0 && nout > 0, BF_ERR_ARG, "bf_population_fitness: bad shape");
  BF_REQUIRE(ldg >= (long)nin * nh + nh + (long)nh * nout + nout && (ldg & 3) == 0, BF_ERR_ARG,
             "bf_population_fitness: genome row stride %ld too small or not a multiple of 4", ldg);
  BF_REQUIRE(precision == BF_PREC_TF32 && fitness_tma_supported(n > 0 ? n : 1, N, nin, nh, nout), BF_ERR_ARG,
             "bf_population_fitness: the resident store serves the TMA + tcgen05 fitness kernel only (BF_PREC_TF32, nh <= 64, nout <= 16, nin %% 4 == 0)");
  if (n == 0) return BF_OK;
  return fitness_tma(G, ldg, Bt, inds, X, Y, fitness, n, N, nin, nh, nout, as_stream(stream));
}

int bf_population_relayout(const float* G, long ldg, float* Bt, const int* inds, int n, int nin, int nh, void* stream) {
  BF_REQUIRE(G && Bt && n >= 0 && nin > 0 && nh > 0 && nh <= 64, BF_ERR_ARG, "bf_population_relayout: bad arguments");
  cudaStream_t s = as_stream(stream);
  for (int off = 0; off < n; off += 65535) {       // grid z limit
    const int c = n - off < 65535 ? n - off : 65535;
    if (inds) {
      int rc = fitness_relayout_resident(G, ldg, Bt, inds + off, c, nin, nh, s);
      if (rc) return rc;
    } else {
      int rc = fitness_relayout_resident(G + (size_t)off * ldg, ldg, Bt + (size_t)off * 64 * nin, nullptr, c, nin, nh, s);
      if (rc) return rc;
    }
  }
  return BF_OK;
}

int bf_population_next_generation(const float* G_old, float* G_new, long ldg, int loci, int P, const unsigned char* survive,
                                  const int* left, const int* right, float rate, unsigned seed, unsigned generation,
                                  const float* coin, const float* mut_u, const float* mut_val, unsigned char* mutant,
                                  void* stream) {
  BF_REQUIRE(G_old && G_new && G_old != G_new && survive && left && right && mutant, BF_ERR_ARG,
             "bf_population_next_generation: null pointer (or in-place call: the generations are double-buffered)");
  BF_REQUIRE(P >= 0 && loci > 0 && ldg >= loci && (ldg & 3) == 0, BF_ERR_ARG, "bf_population_next_generation: bad shape");
  BF_REQUIRE(rate >= 0.f && rate < 1.f, BF_ERR_ARG, "bf_population_next_generation: mutation rate %g out of range", (double)rate);
  BF_REQUIRE((mut_u == nullptr) == (mut_val == nullptr), BF_ERR_ARG, "bf_population_next_generation: mut_u and mut_val come together");
  BF_REQUIRE(((uintptr_t)G_old & 15) == 0 && ((uintptr_t)G_new & 15) == 0, BF_ERR_ARG, "bf_population_next_generation: alignment");
  if (P == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  BF_CUDA(cudaMemsetAsync(mutant, 0, (size_t)P, s));
  GaParams p;
  p.old = G_old; p.next = G_new; p.ldg = ldg; p.loci = loci; p.P = P; p.survive = survive; p.left = left; p.right = right;
  p.rate = rate;
  const double th = (double)rate * 4294967296.0;
  p.mut_threshold = th >= 4294967295.0 ? 4294967295u : (uint32_t)th;
  p.seed = seed; p.generation = generation; p.coin = coin; p.mut_u = mut_u; p.mut_val = mut_val; p.mutant = mutant;
  ga_next_generation_kernel<<<ew_grid((size_t)P * (size_t)(ldg >> 2), 256), 256, 0, s>>>(p);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // extern "C"
