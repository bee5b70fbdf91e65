// Dense layer contractions: atomic/core_op.py:9-20 (DenseOp.forward / backward).
#include <stdlib.h>
#include <stdint.h>
#include "gemm_umma.cuh"
#include "gemm_tma.cuh"

namespace bf {

constexpr int SK_RED_SLICES = 16;
// block = 32 consecutive outputs x 16 slices of the split range, fixed-order tree (deterministic): the dense head's
// 128-chunk partials used to be added by one thread per output, a 128-long chain of L2 latencies
__global__ void __launch_bounds__(32 * SK_RED_SLICES)
splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, float* __restrict__ aux,
                     int rows, int N, int aux_row, int splits, int accumulate, long out_rs, long out_cs) {
  __shared__ float red[SK_RED_SLICES][33];
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)rows * N;
  const long i = (long)blockIdx.x * 32 + threadIdx.x;
  const int zs = threadIdx.y;
  float s = 0.f;
  if (i < total)
    for (int z = zs; z < splits; z += SK_RED_SLICES) s += ws[(long)z * total + i];
  red[zs][threadIdx.x] = s;
  __syncthreads();
#pragma unroll
  for (int w = SK_RED_SLICES / 2; w >= 1; w >>= 1) {
    if (zs < w) red[zs][threadIdx.x] += red[zs + w][threadIdx.x];
    __syncthreads();
  }
  if (zs == 0 && i < total) {
    s = red[0][threadIdx.x];
    int r = (int)(i / N), c = (int)(i % N);
    float* dst;
    if (r == aux_row) dst = aux ? aux + c : nullptr;
    else dst = out + (long)(r > aux_row && aux_row >= 0 ? r - 1 : r) * out_rs + (long)c * out_cs;
    if (dst) *dst = accumulate ? *dst + s : s;
  }
}

// Few splits over a LARGE output (the LSTM weight gradient: 385 x 1024 outputs, 6 splits): one thread per four consecutive
// outputs of a row, the splits added in order (deterministic) from 16-byte loads.  The sliced kernel above launches one
// 512-thread block per 32 outputs -- 12 k blocks for that matrix, most of their threads without work (58 us).
__global__ void __launch_bounds__(256)
splitk_reduce_wide_kernel(const float* __restrict__ ws, float* __restrict__ out, float* __restrict__ aux, int rows, int N,
                          int aux_row, int splits, int accumulate, long out_rs) {
  pdl_launch_dependents();
  pdl_wait();
  const int n4 = N >> 2;
  const long total4 = (long)rows * n4, stride4 = (long)rows * N / 4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c4 = (int)(i - (long)r * n4);
    const float4* src = reinterpret_cast<const float4*>(ws) + i;
    float4 a = src[0];
    for (int z = 1; z < splits; ++z) {
      const float4 b = src[(long)z * stride4];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float* dst;
    if (r == aux_row) dst = aux ? aux + c4 * 4 : nullptr;
    else dst = out + (long)(r > aux_row && aux_row >= 0 ? r - 1 : r) * out_rs + c4 * 4;
    if (dst) {
      if (accumulate) { dst[0] += a.x; dst[1] += a.y; dst[2] += a.z; dst[3] += a.w; }
      else { dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; }
    }
  }
}

int launch_splitk_reduce(const float* ws, float* out, float* aux, int rows, int N, int aux_row, int splits,
                         int accumulate, long out_rs, long out_cs, cudaStream_t s) {
  long total = (long)rows * N;
  if (total == 0) return BF_OK;
  if (splits <= 16 && out_cs == 1 && (N & 3) == 0 && total >= (1L << 16) && (((uintptr_t)ws) & 15) == 0) {
    launch_pdl(splitk_reduce_wide_kernel, dim3((unsigned)ew_grid((size_t)(total / 4), 256)), dim3(256), 0, s, ws, out, aux, rows, N,
               aux_row, splits, accumulate, out_rs);
    BF_LAUNCH_CHECK();
    return BF_OK;
  }
  launch_pdl(splitk_reduce_kernel, dim3((unsigned)((total + 31) / 32)), dim3(32, SK_RED_SLICES), 0, s, ws, out, aux, rows, N, aux_row,
             splits, accumulate, out_rs, out_cs);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// out[m*N+n] = act(acc + b[n])   (softmax rows are finished by a second, in-place kernel)
struct EpiDenseFwd {
  float* out; const float* b; int N; int act;
  __device__ void store(int m, int n, float v, int) const {
    if (b) v += __ldg(b + n);
    out[(long)m * N + n] = act_fwd_rt(act, v);
  }
};
struct EpiPlain {
  float* out; int N;
  __device__ void store(int m, int n, float v, int) const { out[(long)m * N + n] = v; }
};


// ---------------------------------------------------------------------------------------------------------------
// Skinny dense layers (n <= 32 output columns, e.g. the 10-class heads of every BASELINE config).  At n = 10 the
// contraction has 5 FLOP per byte of X: it is HBM-bound, a 128 x 16 tensor-core tile would idle 3/4 of its columns
// and (measured, B200) costs ~40 us of fixed pipeline latency per call, so these run as FP32 FMA kernels that
// stream X once at full width, with W resident in shared memory / registers.  (Exact FP32 on both precision paths.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int SKINNY_ROWS = 2;          // rows per warp iteration (forward): more warps in flight beats W reuse here

// forward: one warp per SKINNY_ROWS rows, lanes split K; W staged once per CTA as [k][NT + pad] (conflict-free LDS.128)
template <int NT>
__global__ void __launch_bounds__(256)
dense_skinny_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                        float* __restrict__ out, int m, int k, int n, int act, int pc, int phw) {
  extern __shared__ __align__(16) float sw[];
  constexpr int PITCH = NT == 4 ? 4 : NT + 4;
  pdl_launch_dependents();
  pdl_wait();
  // W is (k x n) contiguous: read it flat (coalesced, 8 independent loads in flight per thread -- a one-load-at-a-time
  // loop made this staging the whole cost of the kernel at small batch), scatter into the padded [k][PITCH] tile
  for (int i = threadIdx.x; i < k * PITCH; i += blockDim.x) sw[i] = 0.f;
  __syncthreads();
  const int total = k * n;
  for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < total ? __ldg(W + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total) {
        const int kl = i / n;                                         // logical row of W: feature (c, yx) of an NCHW flatten
        const int kk = pc ? (kl % phw) * pc + kl / phw : kl;          // row of X's PHYSICAL (channels-last) feature order
        sw[kk * PITCH + (i - kl * n)] = v[u];
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int ngroups = (m + SKINNY_ROWS - 1) / SKINNY_ROWS;
  for (int grp = blockIdx.x * wpb + warp; grp < ngroups; grp += gridDim.x * wpb) {
    const int m0 = grp * SKINNY_ROWS;
    float acc[SKINNY_ROWS][NT];
#pragma unroll
    for (int r = 0; r < SKINNY_ROWS; ++r)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[r][j] = 0.f;
#pragma unroll 4
    for (int kk = lane; kk < k; kk += 32) {
      float x[SKINNY_ROWS];
#pragma unroll
      for (int r = 0; r < SKINNY_ROWS; ++r) x[r] = (m0 + r < m) ? __ldg(X + (size_t)(m0 + r) * k + kk) : 0.f;
      const float4* wr = reinterpret_cast<const float4*>(sw + kk * PITCH);
#pragma unroll
      for (int q = 0; q < NT / 4; ++q) {
        const float4 w4 = wr[q];
#pragma unroll
        for (int r = 0; r < SKINNY_ROWS; ++r) {
          acc[r][4 * q + 0] = fmaf(x[r], w4.x, acc[r][4 * q + 0]);
          acc[r][4 * q + 1] = fmaf(x[r], w4.y, acc[r][4 * q + 1]);
          acc[r][4 * q + 2] = fmaf(x[r], w4.z, acc[r][4 * q + 2]);
          acc[r][4 * q + 3] = fmaf(x[r], w4.w, acc[r][4 * q + 3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SKINNY_ROWS; ++r) {
      float mine = 0.f;                      // lane j ends up with column j of row m0 + r
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float t = warp_sum(acc[r][j]);
        if (lane == j) mine = t;
      }
      const bool col = lane < n;
      if (b && col) mine += __ldg(b + lane);
      if (act == BF_ACT_SOFTMAX) {           // SoftMax.t1, atomic/activation_op.py:127-130
        const float mx = warp_max(col ? mine : -INFINITY);
        const float e = col ? expf(mine - mx) : 0.f;
        mine = e / warp_sum(e);
      } else {
        mine = act_fwd_rt(act, mine);
      }
      if (col && m0 + r < m) out[(size_t)(m0 + r) * n + lane] = mine;
    }
  }
}

// forward without staging: a warp owns DIRECT_ROWS rows, a lane four consecutive features of the PHYSICAL order (one
// 128-bit load of X per row), W rows come straight from L1/L2 (n floats each, the whole matrix is a few tens of KB).
// No per-CTA preamble: at small batch the staged kernel above spends its time copying W into shared memory once per CTA.
constexpr int DIRECT_ROWS = 4;
template <int NT>
__global__ void __launch_bounds__(128)
dense_skinny_fwd_direct_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                               float* __restrict__ out, int m, int k, int n, int act, int pc, int phw) {
  // the four warps of a CTA share DIRECT_ROWS rows and split K four ways (a single warp per row group left the
  // L1/L2 latency of the W rows exposed); their partial sums meet in shared memory
  __shared__ float red[4][DIRECT_ROWS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * DIRECT_ROWS;
  float acc[DIRECT_ROWS][NT];
#pragma unroll
  for (int r = 0; r < DIRECT_ROWS; ++r)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[r][j] = 0.f;
  for (int kq = threadIdx.x * 4; kq < k; kq += 512) {
    float4 x[DIRECT_ROWS];
#pragma unroll
    for (int r = 0; r < DIRECT_ROWS; ++r)
      x[r] = (m0 + r < m) ? __ldcs(reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * k + kq)) : make_float4(0.f, 0.f, 0.f, 0.f);
    // logical W row of physical feature kq: (c, yx) -> c * phw + yx; the four features of a quad share yx (pc % 4 == 0)
    const int yx = pc ? kq / pc : 0;
    const int kl0 = pc ? (kq - yx * pc) * phw + yx : kq;
    const int kstep = pc ? phw : 1;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2* wr = reinterpret_cast<const float2*>(W + (size_t)(kl0 + u * kstep) * n);   // n even: 8-byte aligned rows
      float w[NT];
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        const float2 t = (j < n) ? __ldg(wr + (j >> 1)) : make_float2(0.f, 0.f);
        w[j] = t.x; w[j + 1] = t.y;
      }
#pragma unroll
      for (int r = 0; r < DIRECT_ROWS; ++r) {
        const float xv = u == 0 ? x[r].x : u == 1 ? x[r].y : u == 2 ? x[r].z : x[r].w;
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[r][j] = fmaf(xv, w[j], acc[r][j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < DIRECT_ROWS; ++r) {
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float t = warp_sum(acc[r][j]);
      if (lane == j) mine = t;
    }
    red[warp][r][lane] = mine;
  }
  __syncthreads();
  if (warp < DIRECT_ROWS && m0 + warp < m) {      // warp r finishes row m0 + r
    const int r = warp;
    float mine = (red[0][r][lane] + red[1][r][lane]) + (red[2][r][lane] + red[3][r][lane]);
    const bool col = lane < n;
    if (b && col) mine += __ldg(b + lane);
    if (act == BF_ACT_SOFTMAX) {
      const float mx = warp_max(col ? mine : -INFINITY);
      const float e = col ? expf(mine - mx) : 0.f;
      mine = e / warp_sum(e);
    } else {
      mine = act_fwd_rt(act, mine);
    }
    if (col) out[(size_t)(m0 + r) * n + lane] = mine;
  }
}

// backward: CTA = (row chunk, 256-column slab of K), thread = one column kk.  Per row: dX[m][kk] = E[m,:] . W[kk,:]
// and the running dW[kk,:] += X[m][kk] * E[m,:]; per-chunk partials (row k of a partial = db) go to the workspace and
// are summed in chunk order by splitk_reduce_kernel (deterministic).
template <int NT>
__global__ void __launch_bounds__(256)
dense_skinny_bwd_kernel(const float* __restrict__ X, const float* __restrict__ E, const float* __restrict__ W,
                        float* __restrict__ partial, float* __restrict__ dX, int m, int k, int n, int rows_per_chunk,
                        int pc, int phw) {
  extern __shared__ __align__(16) float se[];     // [rows_per_chunk][NT]
  pdl_launch_dependents();
  pdl_wait();
  const int m0 = blockIdx.x * rows_per_chunk;
  int rows = m - m0;
  if (rows > rows_per_chunk) rows = rows_per_chunk;
  for (int i = threadIdx.x; i < rows * NT; i += blockDim.x) {
    const int r = i / NT, j = i - r * NT;
    se[i] = j < n ? E[(size_t)(m0 + r) * n + j] : 0.f;
  }
  __syncthreads();
  const int kk = blockIdx.y * blockDim.x + threadIdx.x;
  float* part = partial + (size_t)blockIdx.x * (k + 1) * n;
  if (kk < k) {
    // kk = PHYSICAL column of X / dX; kl = the row of W / dW it belongs to (differs for a channels-last flatten)
    const int kl = pc ? (kk % pc) * phw + kk / pc : kk;
    float w[NT], acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) { w[j] = j < n ? __ldg(W + (size_t)kl * n + j) : 0.f; acc[j] = 0.f; }
    for (int r0 = 0; r0 < rows; r0 += 8) {
      float xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xv[u] = r0 + u < rows ? __ldg(X + (size_t)(m0 + r0 + u) * k + kk) : 0.f;   // 8 loads in flight
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u;
        if (r < rows) {
          const float x = xv[u];
          const float4* e4 = reinterpret_cast<const float4*>(se + r * NT);
          float dx = 0.f;
#pragma unroll
          for (int q = 0; q < NT / 4; ++q) {
            const float4 e = e4[q];
            acc[4 * q + 0] = fmaf(x, e.x, acc[4 * q + 0]); dx = fmaf(e.x, w[4 * q + 0], dx);
            acc[4 * q + 1] = fmaf(x, e.y, acc[4 * q + 1]); dx = fmaf(e.y, w[4 * q + 1], dx);
            acc[4 * q + 2] = fmaf(x, e.z, acc[4 * q + 2]); dx = fmaf(e.z, w[4 * q + 2], dx);
            acc[4 * q + 3] = fmaf(x, e.w, acc[4 * q + 3]); dx = fmaf(e.w, w[4 * q + 3], dx);
          }
          if (dX) dX[(size_t)(m0 + r) * k + kk] = dx;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
      if (j < n) part[(size_t)kl * n + j] = acc[j];
  }
  if (blockIdx.y == 0 && threadIdx.x < n) {       // db partial = column sums of the chunk's E rows
    float sacc = 0.f;
    for (int r = 0; r < rows; ++r) sacc += se[r * NT + threadIdx.x];
    part[(size_t)k * n + threadIdx.x] = sacc;
  }
}

static inline int skinny_nt(int n) { return n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }
static inline bool skinny_ok(int m, int k, int n) {
  if (n > 32 || m < 1) return false;
  const int nt = skinny_nt(n);
  return (size_t)k * (nt == 4 ? 4 : nt + 4) * 4 <= 200 * 1024;
}

static int launch_skinny_fwd(const float* X, const float* W, const float* b, float* out, int m, int k, int n, int act,
                             int pc, int phw, cudaStream_t s) {
  const int nt = skinny_nt(n);
  static const int direct_max = getenv("BF_DENSE_DIRECT_MAX") ? atoi(getenv("BF_DENSE_DIRECT_MAX")) : 2048;
  if (m <= direct_max && (k & 3) == 0 && (n & 1) == 0 && n <= 16 && (pc == 0 || (pc & 3) == 0) && ((uintptr_t)X & 15) == 0 &&
      ((uintptr_t)W & 7) == 0) {
    const int dgrid = (m + DIRECT_ROWS - 1) / DIRECT_ROWS;
    const int dnt = n <= 4 ? 4 : n <= 8 ? 8 : n <= 12 ? 12 : 16;
#define BF_SKD(NT_) if (dnt == NT_) dense_skinny_fwd_direct_kernel<NT_><<<dgrid, 128, 0, s>>>(X, W, b, out, m, k, n, act, pc, phw);
    BF_SKD(4) BF_SKD(8) BF_SKD(12) BF_SKD(16)
#undef BF_SKD
    BF_LAUNCH_CHECK();
    set_last_path(0);
    return BF_OK;
  }
  const size_t smem = (size_t)k * (nt == 4 ? 4 : nt + 4) * 4;
  int grid = (m + 8 * SKINNY_ROWS - 1) / (8 * SKINNY_ROWS);
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
#define BF_SK(NT_)                                                                                                  \
  if (nt == NT_) {                                                                                                  \
    BF_CUDA(cudaFuncSetAttribute(dense_skinny_fwd_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_pdl(dense_skinny_fwd_kernel<NT_>, dim3(grid), dim3(256), smem, s, X, W, b, out, m, k, n, act, pc, phw);                               \
  }
  BF_SK(4) BF_SK(8) BF_SK(16) BF_SK(32)
#undef BF_SK
  BF_LAUNCH_CHECK();
  set_last_path(0);
  return BF_OK;
}

static int launch_skinny_bwd(const float* X, const float* E, const float* W, float* dW, float* db, float* dX, int m,
                             int k, int n, int accumulate, int pc, int phw, cudaStream_t s) {
  const int nt = skinny_nt(n);
  int rows_per_chunk = (m + 127) / 128;
  if (rows_per_chunk < 16) rows_per_chunk = 16;
  if (rows_per_chunk > 256) rows_per_chunk = 256;
  const int chunks = (m + rows_per_chunk - 1) / rows_per_chunk;
  void* ws = nullptr;
  int rc = workspace_require((size_t)chunks * (k + 1) * n * sizeof(float), &ws);
  if (rc) return rc;
  dim3 grid(chunks, (k + 255) / 256);
  const size_t smem = (size_t)rows_per_chunk * nt * 4;
#define BF_SK(NT_) if (nt == NT_) launch_pdl(dense_skinny_bwd_kernel<NT_>, dim3(grid), dim3(256), smem, s, X, E, W, (float*)ws, dX, m, k, n, rows_per_chunk, pc, phw);
  BF_SK(4) BF_SK(8) BF_SK(16) BF_SK(32)
#undef BF_SK
  BF_LAUNCH_CHECK();
  set_last_path(0);
  return launch_splitk_reduce((const float*)ws, dW, db, k + 1, n, k, chunks, accumulate, n, 1, s);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused output head: Dense(n <= 32, act) forward + cost value + cost derivative + Dense backward in ONE kernel, for the
// last layer of a network inside learn_batch (layers/core.py:27-39 + metrics/costs.py:19-37 + metrics/metrics.py:15-16 +
// learner/backpropagation.py:19-22).  A CTA owns a chunk of R rows: their X rows are staged in shared memory ONCE and
// serve both directions (forward dot products, then dW = X^T delta and dX = delta W^T); W (k x n) sits next to them.
// The unfused sequence -- skinny forward, cost/delta kernel, accuracy, skinny backward, split reduction -- read X
// twice and cost five launches of ~4-20 us each for ~8 MB of data (cfg3 head: 51 us; HBM floor ~3 us).
//   out[m][n]  = act(X W + b)                                   (stored: layer.output)
//   cost      += cost(out, Y) partials (double, per CTA; the last CTA adds them in CTA order: deterministic)
//   hits      += argmax(out) == argmax(Y)
//   delta      = (out - Y) * w_row * act'(out)                 (softmax / linear: act' = 1, activation_op.py:79,134)
//   partial[cta] = [X^T delta ; column sums of delta]          (added by splitk_reduce_kernel in CTA order)
//   dX[m][k]   = delta W^T                                     (physical column order of X)
__device__ unsigned int g_head_ticket = 0;
__device__ unsigned int g_head_hits = 0;

constexpr int HEAD_THREADS = 512;     // 16 warps: the kernel is a chain of latencies (staging, warp reductions), more warps hide them
template <int NT>
__global__ void __launch_bounds__(HEAD_THREADS)
dense_head_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                  const float* __restrict__ Y, const float* __restrict__ sw_rows, float* __restrict__ out,
                  float* __restrict__ partial, float* __restrict__ dX, double* __restrict__ cost_partial,
                  float* __restrict__ cost_out, float* __restrict__ hits_out, int m, int k, int n, int R, int act,
                  int cost_kind, int pc, int phw) {
  extern __shared__ __align__(16) float hs[];
  constexpr int PITCH = NT == 4 ? 4 : NT + 4;
  float* sW = hs;                                   // [k][PITCH], rows in X's PHYSICAL column order
  float* sX = sW + (size_t)k * PITCH;               // [R][k]
  float* sD = sX + (size_t)R * k;                   // [R][NT]  delta
  __shared__ double s_cost[HEAD_THREADS / 32];
  __shared__ unsigned s_hits[HEAD_THREADS / 32];
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int m0 = blockIdx.x * R;
  int rows = m - m0;
  if (rows > R) rows = R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- stage W (permuted rows) and the chunk's X rows.  The rows of a chunk are one contiguous run of X: ONE bulk copy
  // (cp.async.bulk, completes on an mbarrier) fetches them while the threads re-lay W; a per-thread load loop here was
  // a chain of ~30 dependent HBM round trips (a third of the kernel).
  __shared__ __align__(8) uint64_t xbar;
  const float* xsrc = X + (size_t)m0 * k;
  const uint32_t xbytes = (uint32_t)rows * (uint32_t)k * 4u;
  const bool bulk = (xbytes & 15u) == 0 && (((uintptr_t)xsrc) & 15) == 0 && ((k * PITCH) & 3) == 0 && xbytes < (1u << 20);
  if (threadIdx.x == 0) {
    mbar_init(&xbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < k * PITCH; i += blockDim.x) sW[i] = 0.f;
  __syncthreads();
  if (bulk && threadIdx.x == 0) {
    mbar_expect_tx(&xbar, xbytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sX)), "l"(reinterpret_cast<uint64_t>(xsrc)), "r"(xbytes), "r"(smem_u32(&xbar)) : "memory");
  }
  // W by rows: thread = row kl of W (its n values are contiguous), ONE index permutation per row instead of two integer
  // divisions per element (that loop was a fifth of the kernel's instructions)
  for (int kl = threadIdx.x; kl < k; kl += blockDim.x) {
    const int kk = pc ? (kl % phw) * pc + kl / phw : kl;
    const float* wr = W + (size_t)kl * n;
    float* dst = sW + kk * PITCH;
    for (int j = 0; j < n; ++j) dst[j] = __ldg(wr + j);
  }
  if (bulk) {
    mbar_wait(&xbar, 0u);
  } else {
    for (int i = threadIdx.x; i < rows * k; i += blockDim.x) sX[i] = __ldg(xsrc + i);
  }
  __syncthreads();
  // ---- forward + cost + delta: one warp per row, lanes split K
  double cost_acc = 0.0;
  unsigned hits = 0;
  for (int r = warp; r < rows; r += HEAD_THREADS / 32) {
    float acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] = 0.f;
    const float* xr = sX + (size_t)r * k;
#pragma unroll 4
    for (int kk = lane; kk < k; kk += 32) {
      const float x = xr[kk];
      const float4* wr = reinterpret_cast<const float4*>(sW + kk * PITCH);
#pragma unroll
      for (int q = 0; q < NT / 4; ++q) {
        const float4 w4 = wr[q];
        acc[4 * q + 0] = fmaf(x, w4.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x, w4.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x, w4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, w4.w, acc[4 * q + 3]);
      }
    }
    float mine = 0.f;                        // lane j ends up with column j
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float t = warp_sum(acc[j]);
      if (lane == j) mine = t;
    }
    const bool col = lane < n;
    if (b && col) mine += __ldg(b + lane);
    if (act == BF_ACT_SOFTMAX) {
      const float mx = warp_max(col ? mine : -INFINITY);
      const float e = col ? expf(mine - mx) : 0.f;
      mine = e / warp_sum(e);
    } else {
      mine = act_fwd_rt(act, mine);
    }
    const size_t row = (size_t)(m0 + r);
    const float tv = col ? __ldg(Y + row * n + lane) : 0.f;
    if (col) out[row * n + lane] = mine;
    // cost terms (metrics/costs.py:26-27,36-37,48-49), summed in double like bf_cost_value
    double c = 0.0;
    if (col) {
      if (cost_kind == BF_COST_MSE) { const float d = mine - tv; c = 0.5 * (double)d * (double)d; }
      else if (cost_kind == BF_COST_CXENT) { if (tv != 0.0f) c = -(double)tv * (double)logf(mine); }
      else {
        if (tv != 0.0f) c -= (double)tv * (double)logf(mine);
        if (tv != 1.0f) c -= (double)(1.0f - tv) * (double)logf(1.0f - mine);
      }
    }
    cost_acc += warp_sum(c);
    if (hits_out) {                          // first-index argmax of both rows (np.argmax)
      float bo = col ? mine : -INFINITY, bt = col ? tv : -INFINITY;
      int ao = lane, at = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float vo = __shfl_xor_sync(0xffffffffu, bo, o), vt = __shfl_xor_sync(0xffffffffu, bt, o);
        const int io = __shfl_xor_sync(0xffffffffu, ao, o), it = __shfl_xor_sync(0xffffffffu, at, o);
        if (vo > bo || (vo == bo && io < ao)) { bo = vo; ao = io; }
        if (vt > bt || (vt == bt && it < at)) { bt = vt; at = it; }
      }
      hits += (ao == at) ? 1u : 0u;
    }
    float d = mine - tv;                     // costs.py:19-21 (hinge is not routed here)
    if (sw_rows) d *= __ldg(sw_rows + row);
    d *= act_bwd_rt(act, mine);              // layers/core.py:35 (1 for softmax / linear)
    if (lane < NT) sD[r * NT + lane] = col ? d : 0.f;
  }
  if (lane == 0) { s_cost[warp] = cost_acc; s_hits[warp] = hits; }
  __syncthreads();
  // ---- backward: thread = physical column kk (strided), rows of the chunk from shared memory
  float* part = partial + (size_t)blockIdx.x * (k + 1) * n;
  for (int kk = threadIdx.x; kk < k; kk += blockDim.x) {
    const int kl = pc ? (kk % pc) * phw + kk / pc : kk;          // row of W / dW this physical column belongs to
    float w[NT], acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) { w[j] = sW[kk * PITCH + j]; acc[j] = 0.f; }
    for (int r = 0; r < rows; ++r) {
      const float x = sX[(size_t)r * k + kk];
      const float4* e4 = reinterpret_cast<const float4*>(sD + r * NT);
      float dx = 0.f;
#pragma unroll
      for (int q = 0; q < NT / 4; ++q) {
        const float4 e = e4[q];
        dx = fmaf(e.x, w[4 * q + 0], dx); dx = fmaf(e.y, w[4 * q + 1], dx);
        dx = fmaf(e.z, w[4 * q + 2], dx); dx = fmaf(e.w, w[4 * q + 3], dx);
        acc[4 * q + 0] = fmaf(x, e.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x, e.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x, e.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x, e.w, acc[4 * q + 3]);
      }
      if (dX) dX[(size_t)(m0 + r) * k + kk] = dx;
    }
    // the weight-gradient partial goes out through shared memory (W's row is dead for this column now): written
    // directly, every thread's n values were a strided 40-byte store -- a quarter of the kernel's stall samples
#pragma unroll
    for (int j = 0; j < NT; ++j) sW[kk * PITCH + j] = acc[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k * n; i += blockDim.x) {
    const int kl = i / n, j = i - kl * n;
    const int kk = pc ? (kl % phw) * pc + kl / phw : kl;
    part[i] = sW[kk * PITCH + j];
  }
  if (threadIdx.x < n) {                     // db partial = column sums of the chunk's delta
    float sacc = 0.f;
    for (int r = 0; r < rows; ++r) sacc += sD[r * NT + threadIdx.x];
    part[(size_t)k * n + threadIdx.x] = sacc;
  }
  // ---- cost / hits: per-CTA partial, the last CTA to arrive adds them in CTA order
  if (threadIdx.x == 0) {
    double c = 0.0;
    unsigned h = 0;
    for (int w8 = 0; w8 < HEAD_THREADS / 32; ++w8) { c += s_cost[w8]; h += s_hits[w8]; }
    cost_partial[blockIdx.x] = c;
    if (hits_out && h) atomicAdd(&g_head_hits, h);
    __threadfence();
    s_last = atomicAdd(&g_head_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double c = 0.0;
    for (unsigned i = 0; i < gridDim.x; ++i) c += ((volatile double*)cost_partial)[i];
    cost_out[0] = (float)c;
    if (hits_out) hits_out[0] = (float)atomicExch(&g_head_hits, 0u);
    g_head_ticket = 0;
  }
}

// rows per CTA and shared-memory bytes of the fused head for a (k, n) layer; R = 0: does not fit
static inline int head_rows(int m, int k, int n, size_t* smem_out) {
  if (n > 32 || k < 1) return 0;
  const int nt = skinny_nt(n), pitch = nt == 4 ? 4 : nt + 4;
  for (int R = 128; R >= 4; R >>= 1) {
    const size_t smem = ((size_t)k * pitch + (size_t)R * k + (size_t)R * nt) * sizeof(float);
    if (smem <= 200 * 1024) {
      // about two CTAs per SM: every CTA pays for staging W once, so more, smaller chunks only help until the machine is full
      int r = 4;
      static const int max_ctas = getenv("BF_HEAD_MAX_CTAS") ? atoi(getenv("BF_HEAD_MAX_CTAS")) : num_sms();
      while (r < R && (m + r - 1) / r > max_ctas) r <<= 1;
      *smem_out = ((size_t)k * pitch + (size_t)r * k + (size_t)r * nt) * sizeof(float);
      return r;
    }
  }
  return 0;
}

// (rows x cols) -> (cols x rows): W^T as the K-major B operand of the TMA-fed forward GEMM
__global__ void dense_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? __ldg(in + (size_t)r * cols + c) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

// Which of the three contractions of a dense layer with n > 32 columns run on the TMA-fed tcgen05 GEMMs (operands
// staged by TMA exactly as stored, north_star) rather than on the gather-staged core: 0 = forward (K = k), 1 = weight
// gradient (both operands MN-major, K = m), 2 = data gradient (K = n).
static inline bool dense_tma_site(int m, int k, int n, int site) {
  if (tma_encode_fn() == nullptr || getenv("BF_DENSE_GATHER") != nullptr || m < 1) return false;
  if (site == 0) return (k & 3) == 0 && k >= 4;
  if (site == 1) return (k & 3) == 0 && (n & 3) == 0;
  return (n & 3) == 0 && n >= 4;
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_dense_cl_supported(int k, int n) { return skinny_ok(1, k, n) ? 1 : 0; }

int bf_dense_head_supported(int k, int n, int act, int cost) {
  size_t smem;
  return (head_rows(1 << 20, k, n, &smem) > 0 && act >= BF_ACT_LINEAR && act <= BF_ACT_SOFTMAX && cost >= BF_COST_MSE &&
          cost <= BF_COST_BXENT) ? 1 : 0;
}

int bf_dense_head(const float* X, const float* W, const float* b, const float* Y, const float* sample_w, float* out,
                  float* dW, float* db, float* dX, float* cost_result, float* hits_result, int m, int chan, int hw, int n,
                  int act, int cost, int channels_last, int accumulate, void* stream) {
  const int k = chan * hw;
  BF_REQUIRE(m > 0 && k > 0 && n > 0 && X && W && Y && out && dW && db && cost_result, BF_ERR_ARG, "bf_dense_head: bad arguments");
  BF_REQUIRE(bf_dense_head_supported(k, n, act, cost), BF_ERR_ARG, "bf_dense_head: unsupported layer (k=%d n=%d act=%d cost=%d)", k, n, act, cost);
  cudaStream_t s = as_stream(stream);
  size_t smem = 0;
  const int R = head_rows(m, k, n, &smem);
  const int chunks = (m + R - 1) / R;
  const size_t part_bytes = ((size_t)chunks * (k + 1) * n * sizeof(float) + 255) & ~(size_t)255;
  void* ws = nullptr;
  int rc = workspace_require(part_bytes + (size_t)chunks * sizeof(double), &ws);
  if (rc) return rc;
  float* partial = (float*)ws;
  double* cpart = (double*)((char*)ws + part_bytes);
  const int nt = skinny_nt(n);
  const int pc = channels_last ? chan : 0, phw = channels_last ? hw : 0;
#define BF_HEAD(NT_)                                                                                                       \
  if (nt == NT_) {                                                                                                         \
    BF_CUDA(cudaFuncSetAttribute(dense_head_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    /* several CTAs per SM: without the carve-out hint the driver sizes shared memory for ONE 75 KB CTA (measured: two waves) */ \
    BF_CUDA(cudaFuncSetAttribute(dense_head_kernel<NT_>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)); \
    launch_pdl(dense_head_kernel<NT_>, dim3(chunks), dim3(HEAD_THREADS), smem, s, X, W, b, Y, sample_w, out, partial, dX, cpart,    \
               cost_result, hits_result, m, k, n, R, act, cost, pc, phw);                                                  \
  }
  BF_HEAD(4) BF_HEAD(8) BF_HEAD(16) BF_HEAD(32)
#undef BF_HEAD
  BF_LAUNCH_CHECK();
  set_last_path(0);
  return launch_splitk_reduce(partial, dW, db, k + 1, n, k, chunks, accumulate, n, 1, s);
}

int bf_dense_tma_supported(int m, int k, int n, int site) {
  return (n > 32 && site >= 0 && site <= 2 && dense_tma_site(m, k, n, site)) ? 1 : 0;
}

int bf_dense_forward(const float* X, const float* W, const float* b, float* out, int m, int k, int n, int act,
                     int precision, void* stream) {
  BF_REQUIRE(m >= 0 && k > 0 && n > 0, BF_ERR_ARG, "bf_dense_forward: bad shape m=%d k=%d n=%d", m, k, n);
  BF_REQUIRE(act >= BF_ACT_LINEAR && act <= BF_ACT_SOFTMAX, BF_ERR_ARG, "bf_dense_forward: bad activation %d", act);
  if (m == 0) return BF_OK;
  BF_REQUIRE(X && W && out, BF_ERR_ARG, "bf_dense_forward: null pointer");
  cudaStream_t s = as_stream(stream);
  if (skinny_ok(m, k, n)) return launch_skinny_fwd(X, W, b, out, m, k, n, act, 0, 0, s);
  int ep_act = act == BF_ACT_SOFTMAX ? BF_ACT_LINEAR : act;
  int rc = BF_OK;
  if (precision == BF_PREC_TF32 && dense_tma_site(m, k, n, 0) && ((uintptr_t)X & 15) == 0) {
    // X is K-major as stored; W (k x n) becomes the K-major B operand by one small transpose into the workspace
    void* ws = nullptr;
    rc = workspace_require(((size_t)k * n * sizeof(float) + 255) & ~(size_t)255, &ws);
    if (rc) return rc;
    dense_transpose_kernel<<<dim3((n + 31) / 32, (k + 31) / 32), dim3(32, 8), 0, s>>>(W, (float*)ws, k, n);
    BF_LAUNCH_CHECK();
    EpiDenseFwd ep{out, b, n, ep_act};
    rc = launch_gemm_tma(X, (const float*)ws, m, n, k, ep, s);
  } else {
    LoadKContig al{X, (long)k};
    LoadMNContig bl{W, (long)n};
    EpiDenseFwd ep{out, b, n, ep_act};
    rc = launch_gemm(precision, al, bl, ep, m, n, k, 1, (k + GEMM_KALIGN - 1) / GEMM_KALIGN * GEMM_KALIGN, 0, s);
  }
  if (rc) return rc;
  if (act == BF_ACT_SOFTMAX) return bf_softmax_forward(out, out, m, n, stream);
  return BF_OK;
}

int bf_dense_backward(const float* X, const float* E, const float* W, float* dW, float* db, float* dX, int m, int k,
                      int n, int accumulate, int precision, void* stream) {
  BF_REQUIRE(m >= 0 && k > 0 && n > 0, BF_ERR_ARG, "bf_dense_backward: bad shape m=%d k=%d n=%d", m, k, n);
  BF_REQUIRE(W && dW && (m == 0 || (X && E)), BF_ERR_ARG, "bf_dense_backward: null pointer");
  cudaStream_t s = as_stream(stream);
  if (skinny_ok(m, k, n)) return launch_skinny_bwd(X, E, W, dW, db, dX, m, k, n, accumulate, 0, 0, s);
  int rc;
  // dW = X^T @ E with a virtual all-ones row appended to X^T: row k of the product is db = E.sum(0).
  {
    int rows = k + 1;
    int bm = rows <= 32 ? 32 : (rows <= 64 ? 64 : 128), bn = n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128));
    int tiles = ((rows + bm - 1) / bm) * ((n + bn - 1) / bn);
    int kps = 0;
    int splits = m > 0 ? choose_splits(tiles, m, &kps) : 1;
    if (m == 0) kps = GEMM_BK;
    void* ws = nullptr;
    rc = workspace_require((size_t)splits * rows * n * sizeof(float), &ws);
    if (rc) return rc;
    const bool tma_dw = precision == BF_PREC_TF32 && m > 0 && dense_tma_site(m, k, n, 1) && ((uintptr_t)X & 15) == 0 &&
                        ((uintptr_t)E & 15) == 0;
    if (tma_dw) {
      // X (m x k) and E (m x n) are both MN-major for this contraction (K = m): read in place; db = column sums of E
      // from the all-ones operand of the first M tile
      const int tiles_mn = ((k + 127) / 128) * ((n + 127) / 128);
      int want = num_sms() / tiles_mn;
      if (want < 1) want = 1;
      if (want > (m + 31) / 32) want = (m + 31) / 32;
      const size_t pbytes = (size_t)want * rows * n * sizeof(float);
      rc = workspace_require(pbytes, &ws);
      if (rc) return rc;
      rc = launch_gemm_tma_mn(X, E, k, n, m, (float*)ws, pbytes, &splits, s);
      if (rc) return rc;
    } else if (m == 0) {
      BF_CUDA(cudaMemsetAsync(ws, 0, (size_t)rows * n * sizeof(float), s));
    } else {
      LoadMNContigOnes al{X, (long)k, k};
      LoadMNContig bl{E, (long)n};
      EpiPartial ep{(float*)ws, (long)rows * n, n};
      rc = launch_gemm(precision, al, bl, ep, rows, n, m, splits, kps, 0, s);
      if (rc) return rc;
    }
    rc = launch_splitk_reduce((const float*)ws, dW, db, rows, n, k, splits, accumulate, n, 1, s);
    if (rc) return rc;
  }
  // dX = E @ W^T
  if (dX && m > 0) {
    EpiPlain ep{dX, k};
    if (precision == BF_PREC_TF32 && dense_tma_site(m, k, n, 2) && ((uintptr_t)E & 15) == 0 && ((uintptr_t)W & 15) == 0) {
      rc = launch_gemm_tma(E, W, m, k, n, ep, s);       // E (m x n) and W (k x n) are both K-major as stored
    } else {
      LoadKContig al{E, (long)n};
      LoadKContig bl{W, (long)n};
      rc = launch_gemm(precision, al, bl, ep, m, k, n, 1, (n + GEMM_KALIGN - 1) / GEMM_KALIGN * GEMM_KALIGN, 0, s);
    }
    if (rc) return rc;
  }
  return BF_OK;
}

// Dense layer directly behind a channels-last convolutional block: X (m, hw*chan) holds the features of an
// (chan, hw) activation in PHYSICAL channels-last order (yx, c) while W (k, n) is indexed by the reference's NCHW flatten
// (c, yx) -- layers/core.py:82-86 + layers/core.py:27-39.  The permutation is applied while W is staged / addressed, so
// the two layout-change passes around Flatten disappear.  n <= 32 only (the skinny kernels); dX comes back in X's order.
int bf_dense_forward_cl(const float* X, const float* W, const float* b, float* out, int m, int chan, int hw, int n, int act,
                        void* stream) {
  const int k = chan * hw;
  BF_REQUIRE(m >= 0 && chan > 0 && hw > 0 && n > 0 && skinny_ok(m > 0 ? m : 1, k, n), BF_ERR_ARG, "bf_dense_forward_cl: unsupported shape");
  BF_REQUIRE(act >= BF_ACT_LINEAR && act <= BF_ACT_SOFTMAX, BF_ERR_ARG, "bf_dense_forward_cl: bad activation %d", act);
  if (m == 0) return BF_OK;
  return launch_skinny_fwd(X, W, b, out, m, k, n, act, chan, hw, as_stream(stream));
}

int bf_dense_backward_cl(const float* X, const float* E, const float* W, float* dW, float* db, float* dX, int m, int chan,
                         int hw, int n, int accumulate, void* stream) {
  const int k = chan * hw;
  BF_REQUIRE(m > 0 && chan > 0 && hw > 0 && n > 0 && skinny_ok(m, k, n), BF_ERR_ARG, "bf_dense_backward_cl: unsupported shape");
  return launch_skinny_bwd(X, E, W, dW, db, dX, m, k, n, accumulate, chan, hw, as_stream(stream));
}

}  // extern "C"
