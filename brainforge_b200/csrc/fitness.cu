// Neuro-evolution fitness fan-out: evolution/_evolution.py:83-97 (Population.update's serial loop over
// individuals) + learner/neuroevolution.py:20-22,34-37 (genome -> weights -> forward -> cost/m).
// All individuals share X, so the P first-layer products collapse into ONE wide GEMM
//   (N x nin) @ (nin x P*nh)      with B read straight out of the genome matrix, (g-0.5)*20 applied on load,
// followed by a per-individual fused second layer + log-softmax + cross-entropy reduction.
#include "gemm_umma.cuh"
#include <stdlib.h>

namespace bf {

// B(k, n=(p,j)) = (G[p*loci + k*nh + j] - 0.5) * 20
struct GenomeW1 {
  static constexpr bool kContigMN = true;
  static constexpr bool kHasTables = false;
  const float* G; long loci; int nh;
  typedef long Row; typedef NoTile Tile;
  __device__ void prepare(int*) const {}
  __device__ Row row(int n) const { int p = n / nh; return (long)p * loci + (n - p * nh); }
  __device__ Tile tile(int) const { return NoTile(); }
  __device__ float load(Row r, Tile, int, int k) const { return (__ldg(G + r + (long)k * nh) - 0.5f) * 20.0f; }
};

// Hbuf[p][j][m] = sigmoid(acc + b1_p[j])
struct EpiHidden {
  float* Hbuf; const float* G; long loci; int nin, nh, Ns;
  __device__ void store(int m, int n, float v, int) const {
    int p = n / nh, j = n - p * nh;
    float b1 = (__ldg(G + (long)p * loci + (long)nin * nh + j) - 0.5f) * 20.0f;
    Hbuf[((long)p * nh + j) * Ns + m] = sigmoidf_(v + b1);
  }
};

// One CTA per individual: second layer + log-softmax + -sum(Y*lsm)/Ns.
constexpr int FIT_MAX_OUT = 32;
__global__ void __launch_bounds__(256)
fitness_head_kernel(const float* __restrict__ Hbuf, const float* __restrict__ G, const float* __restrict__ Y,
                    float* __restrict__ fitness, long loci, int nin, int nh, int nout, int Ns) {
  extern __shared__ float sm[];          // W2 (nh*nout) | b2 (nout) | warp partials (8 doubles)
  float* W2 = sm;
  float* b2 = sm + nh * nout;
  const int p = blockIdx.x;
  const float* g = G + (long)p * loci + (long)nin * nh + nh;
  for (int i = threadIdx.x; i < nh * nout + nout; i += blockDim.x) sm[i] = (g[i] - 0.5f) * 20.0f;
  __syncthreads();
  const float* Hp = Hbuf + (long)p * nh * Ns;
  double acc = 0.0;
  for (int m = threadIdx.x; m < Ns; m += blockDim.x) {
    float logit[FIT_MAX_OUT];
#pragma unroll
    for (int o = 0; o < FIT_MAX_OUT; ++o) logit[o] = o < nout ? b2[o] : 0.f;
    for (int j = 0; j < nh; ++j) {
      float h = Hp[(long)j * Ns + m];
      const float* w = W2 + j * nout;
#pragma unroll
      for (int o = 0; o < FIT_MAX_OUT; ++o)
        if (o < nout) logit[o] = fmaf(h, w[o], logit[o]);
    }
    float mx = logit[0];
#pragma unroll
    for (int o = 1; o < FIT_MAX_OUT; ++o)
      if (o < nout) mx = fmaxf(mx, logit[o]);
    float se = 0.f;
#pragma unroll
    for (int o = 0; o < FIT_MAX_OUT; ++o)
      if (o < nout) se += expf(logit[o] - mx);
    float lse = mx + logf(se);
    const float* y = Y + (long)m * nout;
    float loss = 0.f;
#pragma unroll
    for (int o = 0; o < FIT_MAX_OUT; ++o)
      if (o < nout) {
        float t = y[o];
        if (t != 0.f) loss -= t * (logit[o] - lse);
      }
    acc += (double)loss;
  }
  acc = warp_sum(acc);
  __shared__ double part[8];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    fitness[p] = (float)(s / (double)Ns);
  }
}

// fitness_tma.cu: the same computation on TMA-fed tcgen05 with the whole network in the epilogue
bool fitness_tma_supported(int P, int N, int nin, int nh, int nout);
int fitness_tma(const float* genomes, long ldg, const float* bt_resident, const int* inds, const float* X, const float* Y,
                float* fitness, int P, int N, int nin, int nh, int nout,
                cudaStream_t s);

}  // namespace bf

using namespace bf;

extern "C" {

int bf_fitness_mlp(const float* genomes, const float* X, const float* Y, float* fitness, int P, int N, int nin, int nh,
                   int nout, int precision, void* stream) {
  BF_REQUIRE(genomes && X && Y && fitness, BF_ERR_ARG, "bf_fitness_mlp: null pointer");
  BF_REQUIRE(P >= 0 && N > 0 && nin > 0 && nh > 0 && nout > 0, BF_ERR_ARG, "bf_fitness_mlp: bad shape");
  BF_REQUIRE(nout <= FIT_MAX_OUT, BF_ERR_ARG, "bf_fitness_mlp: nout %d > %d", nout, FIT_MAX_OUT);
  if (P == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  if (precision == BF_PREC_TF32 && fitness_tma_supported(P, N, nin, nh, nout) && getenv("BF_FITNESS_GATHER") == nullptr)
    return fitness_tma(genomes, (long)nin * nh + nh + (long)nh * nout + nout, nullptr, nullptr, X, Y, fitness, P, N, nin, nh, nout, s);
  const long loci = (long)nin * nh + nh + (long)nh * nout + nout;
  size_t per_ind = (size_t)nh * N * sizeof(float);
  size_t have = workspace_bytes();
  if (workspace_ptr() == nullptr || have < per_ind)
    return fail(BF_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu (call bf_set_workspace)", per_ind, have);
  long chunk = (long)(have / per_ind);
  if (chunk > P) chunk = P;
  if (chunk * nh > 2000000000L / 1) chunk = 2000000000L / nh;
  float* Hbuf = (float*)workspace_ptr();
  size_t head_smem = (size_t)(nh * nout + nout) * sizeof(float);
  BF_REQUIRE(head_smem <= 40 * 1024, BF_ERR_ARG, "bf_fitness_mlp: second layer too large for shared memory");
  for (long p0 = 0; p0 < P; p0 += chunk) {
    int pc = (int)((P - p0) < chunk ? (P - p0) : chunk);
    const float* G = genomes + p0 * loci;
    LoadMNContig al_unused{nullptr, 0};
    (void)al_unused;
    LoadKContig al{X, (long)nin};
    GenomeW1 bl{G, loci, nh};
    EpiHidden ep{Hbuf, G, loci, nin, nh, N};
    int rc = launch_gemm(precision, al, bl, ep, N, pc * nh, nin, 1, (nin + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    if (rc) return rc;
    fitness_head_kernel<<<pc, 256, head_smem, s>>>(Hbuf, G, Y, fitness + p0, loci, nin, nh, nout, N);
    BF_LAUNCH_CHECK();
  }
  return BF_OK;
}

}  // extern "C"
