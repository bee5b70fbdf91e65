// Whole training step of a two-layer perceptron at small batch in ONE kernel (BASELINE configs[0]: 784-60-10, batch 20):
//   Dense(h, act1) -> Dense(n, act2) -> cost value, accuracy, cost', both backward passes        (layers/core.py:27-39,
//   atomic/core_op.py:12-20, metrics/costs.py:19-37, metrics/metrics.py:15-16, learner/backpropagation.py:16-30)
// At batch 20 the six launches this replaces were pure latency (79 us per step for 2.8 MFLOP; one of them ran the whole
// 784-deep contraction on ONE SM).  Here a thread-block CLUSTER of 8 CTAs splits the contraction index of the first
// layer: CTA c keeps rows [c*ks, (c+1)*ks) of W1 and the matching columns of X in shared memory for the whole step,
//   1. partial  Hp_c = X[:, slice] W1[slice, :]
//   2. every CTA adds the 8 partials in rank order through distributed shared memory (identical bits in all CTAs), + b1, act1
//   3. (replicated, tiny) out = act2(H W2 + b2), cost, hits, delta2, dW2, db2, dH = (delta2 W2^T) * act1'(H), db1
//   4. dW1[slice, :] = X[:, slice]^T dH  -- each CTA its own rows
// Gradients are ADDED to nabla (Dense accumulates, core.py:37-38); the fused optimizer kernel follows as usual.
// FP32 FMA throughout (both precision settings: the arithmetic is exact-order SIMT, parity <= 1e-5).
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace bf {

constexpr int M2_CLUSTER = 8;
constexpr int M2_THREADS = 512;          // 16 warps: every phase is a chain of shared-memory latencies
constexpr int M2_GROUPS = M2_THREADS / 64;   // row / K groups next to the 64 columns of the hidden layer
constexpr int M2_MAX_M = 32, M2_MAX_H = 64, M2_MAX_N = 16, M2_MAX_KS = 128;

struct Mlp2Params {
  const float *X, *W1, *b1, *W2, *b2, *Y, *sw;
  float *H, *out, *gW1, *gb1, *gW2, *gb2, *cost, *hits;
  int m, k, h, n, ks, act1, act2, cost_kind;
};

__global__ void __launch_bounds__(M2_THREADS, 1) mlp2_step_kernel(Mlp2Params p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  extern __shared__ __align__(16) float sm[];
  const int m = p.m, h = p.h, n = p.n, ks = p.ks;
  const int xp = ks + 1;                          // row pitch of the X slice (odd: conflict-free column walks)
  float* Xs = sm;                                 // [m][xp]
  float* W1s = Xs + m * xp;                       // [ks][h]
  float* Hp = W1s + ks * h;                       // [m][h]  this CTA's partial (read by the peers)
  float* Hs = Hp + m * h;                         // [m][h]  act1(sum + b1)
  float* dH = Hs + m * h;                         // [m][h]
  float* W2s = dH + m * h;                        // [h][n]
  float* Zs = W2s + h * n;                        // [m][n]  logits -> out
  float* D2 = Zs + m * n;                         // [m][n]  delta of the output layer
  __shared__ double s_cost[M2_MAX_M];
  __shared__ unsigned s_hit[M2_MAX_M];
  const int tid = threadIdx.x;
  const int k0 = rank * ks;
  int kn = p.k - k0;                              // rows of W1 this CTA really owns (the last slice may be short)
  if (kn > ks) kn = ks;
  if (kn < 0) kn = 0;

  // ---- stage the slices (rows beyond the matrix are zero)
  for (int i = tid; i < m * ks; i += M2_THREADS) {
    const int r = i / ks, c = i - r * ks;
    Xs[r * xp + c] = c < kn ? __ldg(p.X + (size_t)r * p.k + k0 + c) : 0.f;
  }
#pragma unroll 8
  for (int i = tid; i < ks * h; i += M2_THREADS) W1s[i] = i < kn * h ? __ldg(p.W1 + (size_t)k0 * h + i) : 0.f;
  for (int i = tid; i < h * n; i += M2_THREADS) W2s[i] = __ldg(p.W2 + i);
  __syncthreads();

  // ---- 1. partial first-layer pre-activations: thread = (column j, row group), 64 columns x M2_GROUPS row groups
  {
    constexpr int NR = M2_MAX_M / M2_GROUPS;
    const int j = tid & 63, rg = tid >> 6;
    float acc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] = 0.f;
    if (j < h) {
      for (int kk = 0; kk < kn; ++kk) {
        const float w = W1s[kk * h + j];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const int r = rg + M2_GROUPS * i;
          if (r < m) acc[i] = fmaf(Xs[r * xp + kk], w, acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = rg + M2_GROUPS * i;
        if (r < m) Hp[r * h + j] = acc[i];
      }
    }
  }
  cluster.sync();

  // ---- 2. sum of the partials in rank order (the same order in every CTA), bias, activation
  for (int i = tid; i < m * h; i += M2_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < M2_CLUSTER; ++c) s += cluster.map_shared_rank(Hp, c)[i];
    const int j = i % h;
    const float a = act_fwd_rt(p.act1, s + (p.b1 ? __ldg(p.b1 + j) : 0.f));
    Hs[i] = a;
    if (rank == 0) p.H[i] = a;
  }
  __syncthreads();

  // ---- 3. output layer (replicated in every CTA: 2 m h n FLOP)
  for (int i = tid; i < m * n; i += M2_THREADS) {
    const int r = i / n, j = i - r * n;
    float z = 0.f;
    for (int q = 0; q < h; ++q) z = fmaf(Hs[r * h + q], W2s[q * n + j], z);
    Zs[i] = z + (p.b2 ? __ldg(p.b2 + j) : 0.f);
  }
  __syncthreads();
  if (tid < m) {                                  // one thread per row: softmax / activation, cost, arg-max, delta
    const int r = tid;
    float* z = Zs + r * n;
    if (p.act2 == BF_ACT_SOFTMAX) {
      float mx = -INFINITY;
      for (int j = 0; j < n; ++j) mx = fmaxf(mx, z[j]);
      float se = 0.f;
      for (int j = 0; j < n; ++j) { z[j] = expf(z[j] - mx); se += z[j]; }
      for (int j = 0; j < n; ++j) z[j] = z[j] / se;
    } else {
      for (int j = 0; j < n; ++j) z[j] = act_fwd_rt(p.act2, z[j]);
    }
    double c = 0.0;
    int ao = 0, at = 0;
    float bo = -INFINITY, bt = -INFINITY;
    const float wrow = p.sw ? __ldg(p.sw + r) : 1.f;
    for (int j = 0; j < n; ++j) {
      const float o = z[j], t = __ldg(p.Y + (size_t)r * n + j);
      if (p.cost_kind == BF_COST_MSE) { const float d = o - t; c += 0.5 * (double)d * (double)d; }
      else if (p.cost_kind == BF_COST_CXENT) { if (t != 0.0f) c -= (double)t * (double)logf(o); }
      else {
        if (t != 0.0f) c -= (double)t * (double)logf(o);
        if (t != 1.0f) c -= (double)(1.0f - t) * (double)logf(1.0f - o);
      }
      if (o > bo) { bo = o; ao = j; }             // first index of the maximum (np.argmax)
      if (t > bt) { bt = t; at = j; }
      float d = o - t;                            // costs.py:19-21
      if (p.sw) d *= wrow;
      D2[r * n + j] = d * act_bwd_rt(p.act2, o);  // layers/core.py:35 (1 for softmax / linear)
      if (rank == 0) p.out[(size_t)r * n + j] = o;
    }
    s_cost[r] = c;
    s_hit[r] = ao == at ? 1u : 0u;
  }
  __syncthreads();
  if (rank == 0 && tid == 0) {
    double c = 0.0;
    unsigned hh = 0;
    for (int r = 0; r < m; ++r) { c += s_cost[r]; hh += s_hit[r]; }
    p.cost[0] = (float)c;
    if (p.hits) p.hits[0] = (float)hh;
  }
  // dH = (D2 W2^T) * act1'(H)
  for (int i = tid; i < m * h; i += M2_THREADS) {
    const int r = i / h, q = i - r * h;
    float s = 0.f;
    for (int j = 0; j < n; ++j) s = fmaf(D2[r * n + j], W2s[q * n + j], s);
    dH[i] = s * act_bwd_rt(p.act1, Hs[i]);
  }
  if (rank == 0) {
    // dW2 = H^T D2, db2 = column sums of D2   (added: Dense accumulates)
    for (int i = tid; i < h * n; i += M2_THREADS) {
      const int q = i / n, j = i - q * n;
      float s = 0.f;
      for (int r = 0; r < m; ++r) s = fmaf(Hs[r * h + q], D2[r * n + j], s);
      p.gW2[i] += s;
    }
    if (tid < n) {
      float s = 0.f;
      for (int r = 0; r < m; ++r) s += D2[r * n + tid];
      p.gb2[tid] += s;
    }
  }
  __syncthreads();
  if (rank == 0 && tid < h) {
    float s = 0.f;
    for (int r = 0; r < m; ++r) s += dH[r * h + tid];
    p.gb1[tid] += s;
  }
  // ---- 4. this CTA's rows of dW1 = X[:, slice]^T dH: thread = (column j, K group), eight rows of W1 at a time so that one
  // read of dH serves eight independent sums; the accumulating read-modify-write asks for its eight old values together
  {
    const int j = tid & 63, kg = tid >> 6;
    if (j < h) {
      for (int base = kg; base < kn; base += M2_GROUPS * 8) {
        float old[8], sum[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = base + M2_GROUPS * i;
          old[i] = kk < kn ? p.gW1[(size_t)(k0 + kk) * h + j] : 0.f;
          sum[i] = 0.f;
        }
        for (int r = 0; r < m; ++r) {
          const float d = dH[r * h + j];
          const float* xr = Xs + r * xp + base;          // (columns >= kn read zeros or the next array: never stored)
#pragma unroll
          for (int i = 0; i < 8; ++i) sum[i] = fmaf(xr[M2_GROUPS * i], d, sum[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = base + M2_GROUPS * i;
          if (kk < kn) p.gW1[(size_t)(k0 + kk) * h + j] = old[i] + sum[i];
        }
      }
    }
  }
  cluster.sync();                                 // nobody leaves while a peer may still read its partial
}

static inline size_t mlp2_smem(int m, int ks, int h, int n) {
  return ((size_t)m * (ks + 1) + (size_t)ks * h + 3 * (size_t)m * h + (size_t)h * n + 2 * (size_t)m * n) * sizeof(float);
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_mlp2_step_supported(int m, int k, int h, int n, int act1, int act2, int cost) {
  if (m < 1 || m > M2_MAX_M || h < 1 || h > M2_MAX_H || n < 1 || n > M2_MAX_N || k < 1) return 0;
  const int ks = (k + M2_CLUSTER - 1) / M2_CLUSTER;
  if (ks > M2_MAX_KS) return 0;
  if (act1 < BF_ACT_LINEAR || act1 > BF_ACT_RELU || act2 < BF_ACT_LINEAR || act2 > BF_ACT_SOFTMAX) return 0;
  if (cost < BF_COST_MSE || cost > BF_COST_BXENT) return 0;
  return 1;
}

int bf_mlp2_step(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* Y,
                 const float* sample_w, float* H, float* out, float* gW1, float* gb1, float* gW2, float* gb2, float* cost_result,
                 float* hits_result, int m, int k, int h, int n, int act1, int act2, int cost, void* stream) {
  BF_REQUIRE(X && W1 && W2 && Y && H && out && gW1 && gb1 && gW2 && gb2 && cost_result, BF_ERR_ARG, "bf_mlp2_step: null pointer");
  BF_REQUIRE(bf_mlp2_step_supported(m, k, h, n, act1, act2, cost), BF_ERR_ARG,
             "bf_mlp2_step: unsupported network (m=%d k=%d h=%d n=%d act %d/%d cost %d)", m, k, h, n, act1, act2, cost);
  Mlp2Params p;
  p.X = X; p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2; p.Y = Y; p.sw = sample_w;
  p.H = H; p.out = out; p.gW1 = gW1; p.gb1 = gb1; p.gW2 = gW2; p.gb2 = gb2; p.cost = cost_result; p.hits = hits_result;
  p.m = m; p.k = k; p.h = h; p.n = n; p.ks = (k + M2_CLUSTER - 1) / M2_CLUSTER; p.act1 = act1; p.act2 = act2; p.cost_kind = cost;
  const size_t smem = mlp2_smem(m, p.ks, h, n);
  BF_CUDA(cudaFuncSetAttribute(mlp2_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(M2_CLUSTER);
  cfg.blockDim = dim3(M2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = M2_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BF_CUDA(cudaLaunchKernelEx(&cfg, mlp2_step_kernel, p));
  BF_LAUNCH_CHECK();
  set_last_path(0);
  return BF_OK;
}

}  // extern "C"
