// Memory-bound kernels: activations and their derivatives, softmax, cost derivative / value,
// accuracy and the fused optimizer step.  All are 128-bit vectorised grid-stride kernels sized to a
// multiple of the SM count; reductions use warp shuffles and a deterministic last-block finish.
#include "common.cuh"

namespace bf {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------ activations
template <int KIND>
__global__ void act_fwd_kernel(const float* __restrict__ Z, float* __restrict__ A, size_t n, int vec) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    size_t n4 = n >> 2;
    const float4* z4 = reinterpret_cast<const float4*>(Z);
    float4* a4 = reinterpret_cast<float4*>(A);
    for (size_t i = tid; i < n4; i += stride) {
      float4 v = z4[i];
      v.x = act_fwd<KIND>(v.x); v.y = act_fwd<KIND>(v.y); v.z = act_fwd<KIND>(v.z); v.w = act_fwd<KIND>(v.w);
      a4[i] = v;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) A[i] = act_fwd<KIND>(Z[i]);
  } else {
    for (size_t i = tid; i < n; i += stride) A[i] = act_fwd<KIND>(Z[i]);
  }
}

template <int KIND>
__global__ void act_bwd_kernel(const float* __restrict__ A, const float* __restrict__ din, float* __restrict__ dout,
                               size_t n, int vec) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if (vec) {
    size_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(A);
    const float4* d4 = reinterpret_cast<const float4*>(din);
    float4* o4 = reinterpret_cast<float4*>(dout);
    for (size_t i = tid; i < n4; i += stride) {
      float4 a = a4[i], d = d4[i];
      d.x *= act_bwd<KIND>(a.x); d.y *= act_bwd<KIND>(a.y); d.z *= act_bwd<KIND>(a.z); d.w *= act_bwd<KIND>(a.w);
      o4[i] = d;
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += stride) dout[i] = din[i] * act_bwd<KIND>(A[i]);
  } else {
    for (size_t i = tid; i < n; i += stride) dout[i] = din[i] * act_bwd<KIND>(A[i]);
  }
}

// Row softmax, n <= 32: one thread per row, three passes over a row that sits in L1 (rows are
// n*4 <= 128 B; a warp covers 32 consecutive rows = one contiguous span).
__global__ void softmax_small_kernel(const float* __restrict__ Z, float* __restrict__ A, int m, int n) {
  int row = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (; row < m; row += stride) {
    const float* z = Z + (size_t)row * n;
    float* a = A + (size_t)row * n;
    float mx = z[0];
    for (int j = 1; j < n; ++j) mx = fmaxf(mx, z[j]);
    float s = 0.f;
    for (int j = 0; j < n; ++j) {
      float e = expf(z[j] - mx);
      a[j] = e;
      s += e;
    }
    float inv = 1.0f / s;
    for (int j = 0; j < n; ++j) a[j] *= inv;
  }
}
// n > 32: one warp per row, shuffle reductions.
__global__ void softmax_warp_kernel(const float* __restrict__ Z, float* __restrict__ A, int m, int n) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < m; row += nwarps) {
    const float* z = Z + (size_t)row * n;
    float* a = A + (size_t)row * n;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, z[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
      float e = expf(z[j] - mx);
      a[j] = e;
      s += e;
    }
    s = warp_sum(s);
    float inv = 1.0f / s;
    for (int j = lane; j < n; j += 32) a[j] *= inv;
  }
}

// ------------------------------------------------------------------ cost / metrics
__global__ void cost_delta_kernel(const float* __restrict__ o, const float* __restrict__ t,
                                  const float* __restrict__ w, float* __restrict__ d, size_t total, int n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    float v = o[i] - t[i];
    if (w) v *= w[i / n];
    d[i] = v;
  }
}

// Deterministic two-level sum: each CTA reduces its grid-stride slice in double, writes one partial,
// and the last CTA to arrive (atomic ticket) adds the partials in index order.
__device__ unsigned int g_reduce_ticket = 0;

// delta != null: also writes delta = (o - t) * (w ? w[row] : 1) -- CostFunction.derivative (metrics/costs.py:19-21)
// and the cost value of the same forward pass in ONE pass over the predictions (learner/backpropagation.py:19-27)
__global__ void mul_rowmask_kernel(const float* __restrict__ x, const float* __restrict__ m, float* __restrict__ out, size_t total,
                                   size_t cols, float scale) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) out[i] = x[i] * (m ? m[i % cols] : 1.0f) * scale;
}

template <int KIND>
__global__ void cost_value_kernel(const float* __restrict__ o, const float* __restrict__ t, size_t total,
                                  double* __restrict__ partial, float* __restrict__ result, float* __restrict__ delta,
                                  const float* __restrict__ w, int n, double* __restrict__ record = nullptr,
                                  const unsigned* __restrict__ record_ctr = nullptr, int record_slot = 0) {
  __shared__ double sm[32];
  __shared__ bool last;
  pdl_launch_dependents();
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (; i < total; i += stride) {
    float ov = o[i], tv = t[i];
    if (delta) {   // every cost uses outputs - targets (costs.py:19-21) except hinge: -y where output <= 1 (costs.py:58-66)
      const float d = KIND == BF_COST_HINGE ? (ov > 1.0f ? 0.0f : -tv) : ov - tv;
      delta[i] = d * (w ? w[i / n] : 1.0f);
    }
    if (KIND == BF_COST_MSE) {
      float d = ov - tv;
      acc += 0.5 * (double)d * (double)d;
    } else if (KIND == BF_COST_CXENT) {
      if (tv != 0.0f) acc -= (double)tv * (double)logf(ov);
    } else if (KIND == BF_COST_BXENT) {        // -(t log o + (1-t) log(1-o)), costs.py:48-49; 0*log 0 terms contribute 0
      if (tv != 0.0f) acc -= (double)tv * (double)logf(ov);
      if (tv != 1.0f) acc -= (double)(1.0f - tv) * (double)logf(1.0f - ov);
    } else {                                   // hinge: max(0, 1 - t*o), costs.py:54-55
      acc += (double)fmaxf(0.0f, 1.0f - tv * ov);
    }
  }
  acc = warp_sum(acc);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
    acc = warp_sum(acc);
    if (lane == 0) {
      partial[blockIdx.x] = acc;
      __threadfence();
      unsigned int ticket = atomicAdd(&g_reduce_ticket, 1u);
      last = (ticket == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += ((volatile double*)partial)[b];
    if (result) result[0] = (float)s;
    if (record) record[3 * (size_t)record_ctr[0] + record_slot] = s;       // gradient check: the double-precision sum, by parameter
    g_reduce_ticket = 0;
  }
}

__global__ void accuracy_kernel(const float* __restrict__ o, const float* __restrict__ t, int m, int n,
                                unsigned int* __restrict__ count) {
  int row = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  unsigned int hits = 0;
  for (; row < m; row += stride) {
    const float* orow = o + (size_t)row * n;
    const float* trow = t + (size_t)row * n;
    int ao = 0, at = 0;
    float bo = orow[0], bt = trow[0];
    for (int j = 1; j < n; ++j) {
      float ov = orow[j], tv = trow[j];
      if (ov > bo) { bo = ov; ao = j; }   // strict '>' keeps the FIRST maximum, like numpy.argmax
      if (tv > bt) { bt = tv; at = j; }
    }
    hits += (ao == at);
  }
  hits = __reduce_add_sync(0xffffffffu, hits);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd(count, hits);
}
__global__ void count_to_float_kernel(const unsigned int* c, float* r) { r[0] = (float)c[0]; }

// ------------------------------------------------------------------ optimizer
struct OptHP {
  float eta_m;  // eta / m
  float inv_m;  // 1 / m
  float eta;
  float hp0, hp1, hp2;
};

template <int KIND>
__device__ __forceinline__ void opt_update(float& w, float g, float& s1, float& s2, const OptHP& h) {
  if (KIND == BF_OPT_SGD) {                 // gradient_descent.py:8-9
    w = w - g * h.eta_m;
  } else if (KIND == BF_OPT_MOMENTUM) {     // gradient_descent.py:25-28
    s1 = s1 * h.hp0 + g * h.eta_m;
    w = w - s1;
  } else if (KIND == BF_OPT_NESTEROV) {     // gradient_descent.py:44-50 (literal)
    float nabla = g * h.eta_m;
    float wn = s2 - s1 + nabla;
    s2 = w;
    s1 = s1 * h.hp0 + nabla;
    w = wn;
  } else if (KIND == BF_OPT_ADAGRAD) {      // adaptive_gd.py:16-20
    float nabla = g * h.inv_m;
    s2 = s2 + nabla * nabla;
    w = w - (h.eta / sqrtf(s2 + h.hp0)) * nabla;
  } else if (KIND == BF_OPT_RMSPROP) {      // adaptive_gd.py:71-76
    float nabla = g * h.inv_m;
    s2 = s2 * h.hp0 + (1.0f - h.hp0) * nabla * nabla;
    w = w - (h.eta * nabla) / sqrtf(s2 + h.hp1);
  } else {                                  // ADAM, adaptive_gd.py:94-99 (non-standard form)
    s1 = h.hp0 * s1 + (1.0f - h.hp0) * g;
    s2 = h.hp1 * s2 + (1.0f - h.hp1) * g * g;
    w = w - (h.eta_m * s1) / sqrtf(s2 + h.hp2);
  }
}

template <int KIND, bool HAS1, bool HAS2>
__global__ void opt_kernel(float* __restrict__ W, float* __restrict__ G, float* __restrict__ S1,
                           float* __restrict__ S2, size_t n, OptHP h, int zero_grad, int vec) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  size_t n4 = vec ? (n >> 2) : 0;
  float4* W4 = reinterpret_cast<float4*>(W);
  float4* G4 = reinterpret_cast<float4*>(G);
  float4* S14 = reinterpret_cast<float4*>(S1);
  float4* S24 = reinterpret_cast<float4*>(S2);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = tid; i < n4; i += stride) {
    float4 w = W4[i], g = G4[i];
    float4 a = HAS1 ? S14[i] : zero4, b = HAS2 ? S24[i] : zero4;
    opt_update<KIND>(w.x, g.x, a.x, b.x, h);
    opt_update<KIND>(w.y, g.y, a.y, b.y, h);
    opt_update<KIND>(w.z, g.z, a.z, b.z, h);
    opt_update<KIND>(w.w, g.w, a.w, b.w, h);
    W4[i] = w;
    if (HAS1) S14[i] = a;
    if (HAS2) S24[i] = b;
    if (zero_grad) G4[i] = zero4;
  }
  for (size_t i = (n4 << 2) + tid; i < n; i += stride) {
    float w = W[i], g = G[i], a = HAS1 ? S1[i] : 0.f, b = HAS2 ? S2[i] : 0.f;
    opt_update<KIND>(w, g, a, b, h);
    W[i] = w;
    if (HAS1) S1[i] = a;
    if (HAS2) S2[i] = b;
    if (zero_grad) G[i] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void fill_kernel(float* __restrict__ x, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = v;
}

// Data-parallel step tail in ONE kernel: one-shot all-reduce of the flat gradient vector over peer memory (NVLink /
// NVSwitch) fused with the optimizer update.  PUSH form, one synchronisation per step:
//   every rank r, CTA b (owner of slice b of the vector on every rank):
//     1. stores its slice of G into slot r of every peer's receive buffer R[parity][r][.] (remote stores are fire and
//        forget), clears (or keeps) its own G slice, and -- tail CTA -- posts its local sample count in the tail;
//     2. release-stores its epoch into flag (b, r) of every peer's signal pad;
//     3. waits until its own flags (b, 0..world-1) have reached the epoch: all world copies of slice b have landed;
//     4. sums them in rank order (every rank adds the same numbers in the same order: identical bits everywhere),
//        applies the optimizer to W[slice], writes the summed step scalars back to the tail of G.
// The receive buffer is double-buffered by epoch parity: a peer can run at most one step ahead (it needs this rank's
// next push to finish its own next step), so its next push lands in the other half while this rank is still summing.
// The previous pull form needed a second barrier ("peers are done reading my G") before G could be reused.
// A peer that never arrives (crashed rank, mismatched call sequence) must not wedge the GPU or corrupt the replica:
// after ~10 s the CTA gives up, leaves W and the optimizer state of its slice untouched and raises *timeout (a word
// in mapped host memory that the host checks when it reads the step's scalars).
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_recv4(const float* p) {      // written by a peer: never served from this SM's L1
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

constexpr int AR_MAX_WORLD = 16;
struct ArParams {
  float* const* recv;         // device array [world]: every rank's receive buffer R[2][world][n_total] (mine included)
  unsigned* const* pads;      // device array [world] of signal pads, [blocks][world] words each
  unsigned* epoch;            // [gridDim.x] local per-CTA epochs (device memory: the kernel replays unchanged in a CUDA graph)
  unsigned* timeout;          // one word, mapped host memory: set when a peer failed to arrive
  int rank, world;
  size_t n_params, n_total;   // multiples of 4
  float m_local;              // this rank's sample count, posted in tail slot 2 (the sum lets the host detect uneven shards)
};

template <int KIND, bool HAS1, bool HAS2>
__global__ void __launch_bounds__(256)
opt_allreduce_kernel(float* __restrict__ W, float* __restrict__ G, float* __restrict__ S1, float* __restrict__ S2,
                     ArParams p, OptHP h, int zero_grad) {
  __shared__ unsigned s_epoch;
  __shared__ int s_ok;
  if (threadIdx.x == 0) { s_epoch = ++p.epoch[blockIdx.x]; s_ok = 1; }
  __syncthreads();
  const unsigned e = s_epoch;
  const size_t n4 = p.n_total >> 2, np4 = p.n_params >> 2;
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = per * blockIdx.x, hi = lo + per < n4 ? lo + per : n4;
  const size_t half = (size_t)p.world * p.n_total;                  // floats per parity half of a receive buffer
  const size_t my_slot = (size_t)(e & 1u) * half + (size_t)p.rank * p.n_total;
  float4* G4 = reinterpret_cast<float4*>(G);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // 1. push
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float4 g = G4[i];
    if (i == np4) g.z = p.m_local;                                  // tail: [cost sum, hits, local m, -]
    for (int d = 0; d < p.world; ++d) {
      const int dst = (p.rank + d) % p.world;                       // start with myself, spread the peers
      *reinterpret_cast<float4*>(p.recv[dst] + my_slot + 4 * i) = g;
    }
    if (zero_grad && i < np4) G4[i] = zero4;
  }
  // 2. signal, 3. wait
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    __threadfence_system();
    const int t = threadIdx.x;
    st_release_sys(p.pads[t] + (size_t)blockIdx.x * p.world + p.rank, e);
    const unsigned* mine = p.pads[p.rank] + (size_t)blockIdx.x * p.world + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 20000000000LL) {
        s_ok = 0;
        *reinterpret_cast<volatile unsigned*>(p.timeout) = 1u;
        break;
      }
    }
  }
  __syncthreads();
  if (!s_ok) return;                                                // replica left untouched; the host raises
  // 4. sum in rank order + update
  float4* W4 = reinterpret_cast<float4*>(W);
  float4* S14 = reinterpret_cast<float4*>(S1);
  float4* S24 = reinterpret_cast<float4*>(S2);
  const float* R = p.recv[p.rank] + (size_t)(e & 1u) * half;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float4 g = zero4;
    for (int r = 0; r < p.world; ++r) {
      const float4 v = ld_recv4(R + (size_t)r * p.n_total + 4 * i);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
    if (i < np4) {
      float4 w = W4[i];
      float4 a = HAS1 ? S14[i] : zero4, b = HAS2 ? S24[i] : zero4;
      opt_update<KIND>(w.x, g.x, a.x, b.x, h);
      opt_update<KIND>(w.y, g.y, a.y, b.y, h);
      opt_update<KIND>(w.z, g.z, a.z, b.z, h);
      opt_update<KIND>(w.w, g.w, a.w, b.w, h);
      W4[i] = w;
      if (HAS1) S14[i] = a;
      if (HAS2) S24[i] = b;
      if (!zero_grad) G4[i] = g;
    } else {
      G4[i] = g;                                                    // summed step scalars
    }
  }
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_act_forward(int kind, const float* Z, float* A, size_t n, void* stream) {
  if (n == 0) return BF_OK;
  BF_REQUIRE(Z && A, BF_ERR_ARG, "bf_act_forward: null pointer");
  cudaStream_t s = as_stream(stream);
  int vec = aligned16(Z) && aligned16(A);
  int grid = ew_grid((n + 3) / 4, 256);
  switch (kind) {
    case BF_ACT_LINEAR:
      if (Z != A) BF_CUDA(cudaMemcpyAsync(A, Z, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
      return BF_OK;
    case BF_ACT_SIGMOID: act_fwd_kernel<BF_ACT_SIGMOID><<<grid, 256, 0, s>>>(Z, A, n, vec); break;
    case BF_ACT_TANH: act_fwd_kernel<BF_ACT_TANH><<<grid, 256, 0, s>>>(Z, A, n, vec); break;
    case BF_ACT_RELU: act_fwd_kernel<BF_ACT_RELU><<<grid, 256, 0, s>>>(Z, A, n, vec); break;
    default: return fail(BF_ERR_ARG, "bf_act_forward: unsupported activation kind %d (softmax is bf_softmax_forward)", kind);
  }
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_softmax_forward(const float* Z, float* A, int m, int n, void* stream) {
  BF_REQUIRE(m >= 0 && n > 0, BF_ERR_ARG, "bf_softmax_forward: bad shape (%d,%d)", m, n);
  if (m == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  if (n <= 32)
    softmax_small_kernel<<<ew_grid(m, 128), 128, 0, s>>>(Z, A, m, n);
  else
    softmax_warp_kernel<<<ew_grid((size_t)m * 32, 256), 256, 0, s>>>(Z, A, m, n);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_act_backward(int kind, const float* A, const float* din, float* dout, size_t n, void* stream) {
  if (n == 0) return BF_OK;
  BF_REQUIRE(din && dout, BF_ERR_ARG, "bf_act_backward: null pointer");
  cudaStream_t s = as_stream(stream);
  int vec = aligned16(A) && aligned16(din) && aligned16(dout);
  int grid = ew_grid((n + 3) / 4, 256);
  switch (kind) {
    case BF_ACT_LINEAR:
    case BF_ACT_SOFTMAX:  // activation_op.py:79-80,134-135: derivative is the scalar 1
      if (din != dout) BF_CUDA(cudaMemcpyAsync(dout, din, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
      return BF_OK;
    case BF_ACT_SIGMOID: act_bwd_kernel<BF_ACT_SIGMOID><<<grid, 256, 0, s>>>(A, din, dout, n, vec); break;
    case BF_ACT_TANH: act_bwd_kernel<BF_ACT_TANH><<<grid, 256, 0, s>>>(A, din, dout, n, vec); break;
    case BF_ACT_RELU: act_bwd_kernel<BF_ACT_RELU><<<grid, 256, 0, s>>>(A, din, dout, n, vec); break;
    default: return fail(BF_ERR_ARG, "bf_act_backward: unknown activation kind %d", kind);
  }
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_cost_delta(const float* out, const float* tgt, const float* w, float* delta, int m, int n, void* stream) {
  BF_REQUIRE(m >= 0 && n > 0, BF_ERR_ARG, "bf_cost_delta: bad shape (%d,%d)", m, n);
  size_t total = (size_t)m * n;
  if (total == 0) return BF_OK;
  cost_delta_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(out, tgt, w, delta, total, n);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_cost_value(int kind, const float* out, const float* tgt, int m, int n, float* result, void* stream) {
  BF_REQUIRE(m >= 0 && n > 0 && result, BF_ERR_ARG, "bf_cost_value: bad arguments");
  size_t total = (size_t)m * n;
  cudaStream_t s = as_stream(stream);
  if (total == 0) {
    BF_CUDA(cudaMemsetAsync(result, 0, sizeof(float), s));
    return BF_OK;
  }
  int grid = ew_grid(total, 256, 1);
  void* ws = nullptr;
  int rc = workspace_require((size_t)grid * sizeof(double), &ws);
  if (rc) return rc;
#define BF_COST_CASE(K_) case K_: launch_pdl(cost_value_kernel<K_>, dim3(grid), dim3(256), 0, s, out, tgt, total, (double*)ws, result, (float*)nullptr, (const float*)nullptr, n, (double*)nullptr, (const unsigned*)nullptr, 0); break;
  switch (kind) {
    BF_COST_CASE(BF_COST_MSE) BF_COST_CASE(BF_COST_CXENT) BF_COST_CASE(BF_COST_BXENT) BF_COST_CASE(BF_COST_HINGE)
    default: return fail(BF_ERR_ARG, "bf_cost_value: unknown cost kind %d", kind);
  }
#undef BF_COST_CASE
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// delta = (out - tgt) * w[row]  AND  result[0] = cost(out, tgt)  in one kernel
int bf_cost_delta_value(int kind, const float* out, const float* tgt, const float* w, float* delta, float* result, int m, int n,
                        void* stream) {
  BF_REQUIRE(m >= 0 && n > 0 && result && (m == 0 || (out && tgt && delta)), BF_ERR_ARG, "bf_cost_delta_value: bad arguments");
  BF_REQUIRE(kind >= BF_COST_MSE && kind <= BF_COST_HINGE, BF_ERR_ARG, "bf_cost_delta_value: unknown cost kind %d", kind);
  cudaStream_t s = as_stream(stream);
  size_t total = (size_t)m * n;
  if (total == 0) {
    BF_CUDA(cudaMemsetAsync(result, 0, sizeof(float), s));
    return BF_OK;
  }
  int grid = ew_grid(total, 256, 1);
  void* ws = nullptr;
  int rc = workspace_require((size_t)grid * sizeof(double), &ws);
  if (rc) return rc;
#define BF_COST_CASE(K_) case K_: launch_pdl(cost_value_kernel<K_>, dim3(grid), dim3(256), 0, s, out, tgt, total, (double*)ws, result, delta, w, n, (double*)nullptr, (const unsigned*)nullptr, 0); break;
  switch (kind) {
    BF_COST_CASE(BF_COST_MSE) BF_COST_CASE(BF_COST_CXENT) BF_COST_CASE(BF_COST_BXENT) BF_COST_CASE(BF_COST_HINGE)
    default: break;
  }
#undef BF_COST_CASE
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_accuracy(const float* out, const float* tgt, int m, int n, float* result, void* stream) {
  BF_REQUIRE(m >= 0 && n > 0 && result, BF_ERR_ARG, "bf_accuracy: bad arguments");
  cudaStream_t s = as_stream(stream);
  void* ws = nullptr;
  int rc = workspace_require(sizeof(unsigned int), &ws);
  if (rc) return rc;
  BF_CUDA(cudaMemsetAsync(ws, 0, sizeof(unsigned int), s));
  if (m > 0) accuracy_kernel<<<ew_grid(m, 128), 128, 0, s>>>(out, tgt, m, n, (unsigned int*)ws);
  count_to_float_kernel<<<1, 1, 0, s>>>((const unsigned int*)ws, result);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_opt_step(int kind, float* W, float* g, float* s1, float* s2, size_t n, float eta, float m, float hp0, float hp1,
                float hp2, int zero_grad, void* stream) {
  if (n == 0) return BF_OK;
  BF_REQUIRE(W && g, BF_ERR_ARG, "bf_opt_step: null W or g");
  BF_REQUIRE(m > 0.f, BF_ERR_ARG, "bf_opt_step: m must be positive");
  cudaStream_t s = as_stream(stream);
  OptHP h;
  h.eta = eta; h.eta_m = eta / m; h.inv_m = 1.0f / m; h.hp0 = hp0; h.hp1 = hp1; h.hp2 = hp2;
  int vec = aligned16(W) && aligned16(g) && (!s1 || aligned16(s1)) && (!s2 || aligned16(s2));
  int grid = ew_grid((n + 3) / 4, 256);
  switch (kind) {
    case BF_OPT_SGD:
      opt_kernel<BF_OPT_SGD, false, false><<<grid, 256, 0, s>>>(W, g, nullptr, nullptr, n, h, zero_grad, vec);
      break;
    case BF_OPT_MOMENTUM:
      BF_REQUIRE(s1, BF_ERR_ARG, "bf_opt_step: momentum needs velocity (s1)");
      opt_kernel<BF_OPT_MOMENTUM, true, false><<<grid, 256, 0, s>>>(W, g, s1, nullptr, n, h, zero_grad, vec);
      break;
    case BF_OPT_NESTEROV:
      BF_REQUIRE(s1 && s2, BF_ERR_ARG, "bf_opt_step: nesterov needs velocity (s1) and memory (s2)");
      opt_kernel<BF_OPT_NESTEROV, true, true><<<grid, 256, 0, s>>>(W, g, s1, s2, n, h, zero_grad, vec);
      break;
    case BF_OPT_ADAGRAD:
      BF_REQUIRE(s2, BF_ERR_ARG, "bf_opt_step: adagrad needs memory (s2)");
      opt_kernel<BF_OPT_ADAGRAD, false, true><<<grid, 256, 0, s>>>(W, g, nullptr, s2, n, h, zero_grad, vec);
      break;
    case BF_OPT_RMSPROP:
      BF_REQUIRE(s2, BF_ERR_ARG, "bf_opt_step: rmsprop needs memory (s2)");
      opt_kernel<BF_OPT_RMSPROP, false, true><<<grid, 256, 0, s>>>(W, g, nullptr, s2, n, h, zero_grad, vec);
      break;
    case BF_OPT_ADAM:
      BF_REQUIRE(s1 && s2, BF_ERR_ARG, "bf_opt_step: adam needs velocity (s1) and memory (s2)");
      opt_kernel<BF_OPT_ADAM, true, true><<<grid, 256, 0, s>>>(W, g, s1, s2, n, h, zero_grad, vec);
      break;
    default: return fail(BF_ERR_ARG, "bf_opt_step: unknown optimizer kind %d", kind);
  }
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// out[r, i] = x[r, i] * (m ? m[i] : 1) * scale -- DropOut's mask has the shape of ONE sample and is shared by the batch
// (layers/fancy.py:73-76,80)
int bf_mul_rowmask(const float* x, const float* m, float* out, size_t rows, size_t cols, float scale, void* stream) {
  if (rows * cols == 0) return BF_OK;
  BF_REQUIRE(x && out, BF_ERR_ARG, "bf_mul_rowmask: null pointer");
  mul_rowmask_kernel<<<ew_grid(rows * cols, 256), 256, 0, as_stream(stream)>>>(x, m, out, rows * cols, cols, scale);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// Data-parallel optimizer step: all-reduce(sum) of the flat gradient vector over peer memory fused with bf_opt_step.
// G is this rank's buffer inside a symmetric allocation; peers / pads are DEVICE arrays of `world` pointers (every
// rank's buffer and signal pad, as torch's symmetric memory hands them out); epoch is a zero-initialised local array
// of `blocks` + 1 words (per-CTA epochs; the last one is set to 1 if a peer failed to arrive within ~10 s -- the
// results of that step are then undefined), gsum a local scratch of n_total floats.  Elements [n_params, n_total) (the step scalars) are
// summed but not optimised.  Every rank must make the same sequence of calls with the same `blocks`.
int bf_opt_step_allreduce(int kind, float* W, float* G, float* s1, float* s2, size_t n_params, size_t n_total, float eta,
                          float m, float hp0, float hp1, float hp2, int zero_grad, void* const* recv, void* const* pads,
                          unsigned* epoch, unsigned* timeout, float m_local, int rank, int world, int blocks, void* stream) {
  BF_REQUIRE(W && G && recv && pads && epoch && timeout, BF_ERR_ARG, "bf_opt_step_allreduce: null pointer");
  BF_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world && blocks >= 1, BF_ERR_ARG,
             "bf_opt_step_allreduce: bad rank/world/blocks %d/%d/%d", rank, world, blocks);
  BF_REQUIRE((n_params & 3) == 0 && (n_total & 3) == 0 && n_params + 4 <= n_total, BF_ERR_ARG,
             "bf_opt_step_allreduce: sizes must be multiples of 4 with a scalar tail behind the parameters");
  BF_REQUIRE(aligned16(W) && aligned16(G) && (!s1 || aligned16(s1)) && (!s2 || aligned16(s2)), BF_ERR_ARG,
             "bf_opt_step_allreduce: buffers must be 16-byte aligned");
  BF_REQUIRE(m > 0.f, BF_ERR_ARG, "bf_opt_step_allreduce: m must be positive");
  cudaStream_t s = as_stream(stream);
  OptHP h;
  h.eta = eta; h.eta_m = eta / m; h.inv_m = 1.0f / m; h.hp0 = hp0; h.hp1 = hp1; h.hp2 = hp2;
  ArParams p;
  p.recv = reinterpret_cast<float* const*>(recv); p.pads = reinterpret_cast<unsigned* const*>(pads); p.epoch = epoch;
  p.timeout = timeout; p.rank = rank; p.world = world; p.n_params = n_params; p.n_total = n_total; p.m_local = m_local;
#define BF_AR(K_, H1_, H2_) opt_allreduce_kernel<K_, H1_, H2_><<<blocks, 256, 0, s>>>(W, G, s1, s2, p, h, zero_grad)
  switch (kind) {
    case BF_OPT_SGD: BF_AR(BF_OPT_SGD, false, false); break;
    case BF_OPT_MOMENTUM:
      BF_REQUIRE(s1, BF_ERR_ARG, "bf_opt_step_allreduce: momentum needs velocity (s1)");
      BF_AR(BF_OPT_MOMENTUM, true, false); break;
    case BF_OPT_NESTEROV:
      BF_REQUIRE(s1 && s2, BF_ERR_ARG, "bf_opt_step_allreduce: nesterov needs velocity (s1) and memory (s2)");
      BF_AR(BF_OPT_NESTEROV, true, true); break;
    case BF_OPT_ADAGRAD:
      BF_REQUIRE(s2, BF_ERR_ARG, "bf_opt_step_allreduce: adagrad needs memory (s2)");
      BF_AR(BF_OPT_ADAGRAD, false, true); break;
    case BF_OPT_RMSPROP:
      BF_REQUIRE(s2, BF_ERR_ARG, "bf_opt_step_allreduce: rmsprop needs memory (s2)");
      BF_AR(BF_OPT_RMSPROP, false, true); break;
    case BF_OPT_ADAM:
      BF_REQUIRE(s1 && s2, BF_ERR_ARG, "bf_opt_step_allreduce: adam needs velocity (s1) and memory (s2)");
      BF_AR(BF_OPT_ADAM, true, true); break;
    default: return fail(BF_ERR_ARG, "bf_opt_step_allreduce: unknown optimizer kind %d", kind);
  }
#undef BF_AR
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// ---- batched on-device gradient check (gradientcheck/raw_gradients.py:15-42): building blocks, graph-capturable --------
__global__ void gradcheck_perturb_kernel(float* __restrict__ W, const int* __restrict__ offsets, unsigned* __restrict__ counter,
                                         float* __restrict__ saved, double* __restrict__ record, float eps, int phase) {
  const unsigned c = counter[0];
  float* w = W + offsets[c];
  if (phase == 0) {
    saved[0] = *w;
    *w = saved[0] + eps;
    record[3 * (size_t)c + 2] = (double)*w;
  } else if (phase == 1) {
    *w = saved[0] - eps;
    record[3 * (size_t)c + 2] -= (double)*w;              // the step actually taken between the two FP32 weights
  } else {
    *w = saved[0];
    counter[0] = c + 1;
  }
}

int bf_gradcheck_perturb(float* W, const int* offsets, unsigned* counter, float* saved, double* record, float epsilon, int phase,
                         void* stream) {
  BF_REQUIRE(W && offsets && counter && saved && record && phase >= 0 && phase <= 2, BF_ERR_ARG, "bf_gradcheck_perturb: bad arguments");
  gradcheck_perturb_kernel<<<1, 1, 0, as_stream(stream)>>>(W, offsets, counter, saved, record, epsilon, phase);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_gradcheck_cost(int kind, const float* out, const float* tgt, int m, int n, double* record, const unsigned* counter,
                      int slot, void* stream) {
  BF_REQUIRE(m > 0 && n > 0 && out && tgt && record && counter && (slot == 0 || slot == 1), BF_ERR_ARG, "bf_gradcheck_cost: bad arguments");
  cudaStream_t s = as_stream(stream);
  const size_t total = (size_t)m * n;
  int grid = ew_grid(total, 256, 1);
  void* ws = nullptr;
  int rc = workspace_require((size_t)grid * sizeof(double), &ws);
  if (rc) return rc;
#define BF_COST_CASE(K_) case K_: cost_value_kernel<K_><<<grid, 256, 0, s>>>(out, tgt, total, (double*)ws, (float*)nullptr, (float*)nullptr, (const float*)nullptr, n, record, counter, slot); break;
  switch (kind) {
    BF_COST_CASE(BF_COST_MSE) BF_COST_CASE(BF_COST_CXENT) BF_COST_CASE(BF_COST_BXENT) BF_COST_CASE(BF_COST_HINGE)
    default: return fail(BF_ERR_ARG, "bf_gradcheck_cost: unknown cost kind %d", kind);
  }
#undef BF_COST_CASE
  BF_LAUNCH_CHECK();
  return BF_OK;
}

/* memory fills that do not read the destination (a NaN in it must not survive): zero through the copy engine, any other
 * constant through a store-only kernel */
int bf_fill(float* x, size_t n, float value, void* stream) {
  if (n == 0) return BF_OK;
  BF_REQUIRE(x, BF_ERR_ARG, "bf_fill: null pointer");
  cudaStream_t s = as_stream(stream);
  if (value == 0.0f) {
    BF_CUDA(cudaMemsetAsync(x, 0, n * sizeof(float), s));
    return BF_OK;
  }
  fill_kernel<<<ew_grid(n, 256), 256, 0, s>>>(x, n, value);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // extern "C"
