// LSTM forward / BPTT: atomic/recurrent_op.py:46-112 (LSTMOp), llatomic/_lllstm.py:6-67.
// Time is a true recurrence and stays a host loop over T; each step is one gate GEMM
// (B x (D+H)) @ ((D+H) x 4H) plus one fused elementwise gate kernel.  The weight gradient is ONE
// split-K contraction over all T*B rows after the loop (recurrent_op.py:109), with the bias gradient
// riding along as an all-ones row.
#include "gemm_tma.cuh"
#include <stdlib.h>

namespace bf {

// Z[t,b,0:D] = X[t,b,:] for all t;  Z[0,b,D:] = 0  (O[-1] is the still-zero last row, recurrent_op.py:60).
// With `bm` the source is batch-major, X (B,T,D): the layer's `X.transpose(1, 0, 2)` (layers/recurrent.py:26) is folded in.
__global__ void lstm_fill_z_kernel(const float* __restrict__ X, float* __restrict__ Z, long TB, int B, int D, int H, int T = 0,
                                   int bm = 0) {
  long total = TB * D + (long)B * H;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    if (i < TB * D) {
      long r = i / D;
      int c = (int)(i - r * D);
      long src = i;
      if (bm) {
        const long t = r / B, b = r - t * B;
        src = (b * T + t) * D + c;
      }
      Z[r * (D + H) + c] = X[src];
    } else {
      long j = i - TB * D;
      long r = j / H;
      int c = (int)(j - r * H);
      Z[r * (D + H) + D + c] = 0.f;
    }
  }
}

// act'(z) computed from the pre-activation z (a = act(z) given), free of cancellation
__device__ __forceinline__ float act_deriv_from_z(int act, float z, float a) {
  switch (act) {
    case BF_ACT_SIGMOID: return a * sigmoidf_(-z);
    case BF_ACT_TANH: {   // sech^2(z) = 4 e^{-2|z|} / (1 + e^{-2|z|})^2
      float e = expf(-2.0f * fabsf(z));
      float d = 1.0f + e;
      return 4.0f * e / (d * d);
    }
    case BF_ACT_RELU: return a > 0.0f ? 1.0f : 0.0f;
    default: return 1.0f;
  }
}

// gates (B,4H) pre-activations [cand|f|i|o] -> cache rows at time t, O[t], and the recurrent half of Z[t+1]
__global__ void lstm_gates_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Cprev,
                                      float* __restrict__ C, float* __restrict__ Ca, float* __restrict__ cand,
                                      float* __restrict__ f, float* __restrict__ ig, float* __restrict__ o,
                                      float* __restrict__ O, float* __restrict__ Znext, float* __restrict__ bw,
                                      long bw_stride, int B, int D, int H, int act) {
  long total = (long)B * H;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    long b = idx / H;
    int j = (int)(idx - b * H);
    const float* p = P + b * 4 * H;
    float cv = act_fwd_rt(act, p[j]);
    float fv = sigmoidf_(p[H + j]);
    float iv = sigmoidf_(p[2 * H + j]);
    float ov = sigmoidf_(p[3 * H + j]);
    float cp = Cprev ? Cprev[idx] : 0.f;
    float c = cp * fv + cv * iv;       // recurrent_op.py:72
    float ca = act_fwd_rt(act, c);
    float out = ca * ov;               // recurrent_op.py:74-75
    C[idx] = c; Ca[idx] = ca; cand[idx] = cv; f[idx] = fv; ig[idx] = iv; o[idx] = ov;
    O[idx] = out;
    if (bw) {
      // act'(.) of the five cached activations, evaluated from the PRE-activations: forming A(1-A) or
      // 1-A^2 from FP32-rounded saturated outputs (all biases start at +7) would cancel catastrophically
      // stored as the five per-element COEFFICIENTS of the backward recurrence (recurrent_op.py:94-104):
      // o*act'(C), i*act'(cand), C[t-1]*f', cand*i', Ca*o'
      bw[idx] = ov * act_deriv_from_z(act, c, ca);
      bw[bw_stride + idx] = iv * act_deriv_from_z(act, p[j], cv);
      bw[2 * bw_stride + idx] = cp * (fv * sigmoidf_(-p[H + j]));
      bw[3 * bw_stride + idx] = cv * (iv * sigmoidf_(-p[2 * H + j]));
      bw[4 * bw_stride + idx] = ca * (ov * sigmoidf_(-p[3 * H + j]));
    }
    if (Znext) Znext[b * (D + H) + D + j] = out;
  }
}

// One BPTT step of the elementwise part (recurrent_op.py:94-104); dH is deltaZ[t+1][:, D:] (or null at t=T-1).
__global__ void lstm_gates_bwd_kernel(const float* __restrict__ E, const float* __restrict__ dH,
                                      const float* __restrict__ Cprev, const float* __restrict__ Ca,
                                      const float* __restrict__ cand, const float* __restrict__ f,
                                      const float* __restrict__ ig, const float* __restrict__ o,
                                      const float* __restrict__ bw, long bw_stride, float* __restrict__ deltaC,
                                      float* __restrict__ dgates, int B, int H, int act, int first) {
  long total = (long)B * H;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    long b = idx / H;
    int j = (int)(idx - b * H);
    float e = E[idx] + (dH ? dH[idx] : 0.f);
    float fv = f[idx];
    float k0, k1, k2, k3, k4;      // o*act'(C), i*act'(cand), C[t-1]*f', cand*i', Ca*o'
    if (bw) {
      k0 = bw[idx]; k1 = bw[bw_stride + idx]; k2 = bw[2 * bw_stride + idx];
      k3 = bw[3 * bw_stride + idx]; k4 = bw[4 * bw_stride + idx];
    } else {  // derivative-from-output exactly as recurrent_op.py:85-88 (FP32 cancellation when saturated)
      float cav = Ca[idx], cv = cand[idx], iv = ig[idx], ov = o[idx];
      float cp = Cprev ? Cprev[idx] : 0.f;
      k0 = ov * act_bwd_rt(act, cav); k1 = iv * act_bwd_rt(act, cv);
      k2 = cp * (fv * (1.f - fv)); k3 = cv * (iv * (1.f - iv)); k4 = cav * (ov * (1.f - ov));
    }
    float dC = (first ? 0.f : deltaC[idx]) + e * k0;
    float* dg = dgates + b * 4 * H;
    dg[j] = dC * k1;
    dg[H + j] = dC * k2;
    dg[2 * H + j] = dC * k3;
    dg[3 * H + j] = e * k4;
    deltaC[idx] = dC * fv;
  }
}

// deltaZ[t] = dgates[t] @ W^T, split on the fly: columns < D are dX[t], the rest feed E[t-1]
struct EpiLstmDz {
  float* dX; float* dH; int D, H;
  __device__ void store(int m, int n, float v, int) const {
    if (n < D) { if (dX) dX[(long)m * D + n] = v; }
    else dH[(long)m * H + (n - D)] = v;
  }
};
struct EpiBias {
  float* out; const float* b; int N;
  __device__ void store(int m, int n, float v, int) const { out[(long)m * N + n] = v + __ldg(b + n); }
};

// ---------------------------------------------------------------------------------------------------------------
// One forward time step as ONE kernel (TF32 path): the gate GEMM Z[t] @ W with the whole elementwise gate update in
// its epilogue.  A tile is 128 batch rows x 16 hidden units x 4 gates: the B operand gathers the four gates' rows of
// W^T with four TMA boxes, so one thread owns all four pre-activations of a (row, unit) pair and the (B, 4H) gate
// matrix never goes to HBM.  Writes the six cache rows, O[t], the recurrent half of Z[t+1] and the derivative rows.
constexpr int LF_UNITS = 16, LF_N = 4 * LF_UNITS, LF_THREADS = 384, LF_STAGES = 8;
constexpr uint32_t LF_A_BYTES = 128 * 128, LF_B_BYTES = LF_N * 128, LF_STAGE = LF_A_BYTES + LF_B_BYTES;

struct LstmFwdParams {
  int t, T, B, D, H, nkb, m_tiles, act;
  long TBH;
  const float* bias;
  float *cache, *O, *Z, *bw;
};

__device__ __forceinline__ void store8(float* p, const float (&v)[8], int nvalid, bool vec) {
  if (vec && nvalid == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nvalid) p[i] = v[i];
  }
}

__global__ void __launch_bounds__(LF_THREADS, 1)
lstm_step_fwd_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapW, LstmFwdParams p) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[LF_STAGES], empty[LF_STAGES], done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const int nt = blockIdx.x / p.m_tiles, mt = blockIdx.x - nt * p.m_tiles;
  const int j0 = nt * LF_UNITS;

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(LF_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < LF_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    if (elect_one_sync()) {
      int sl = 0, u = 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (u > 0) mbar_wait(&empty[sl], (uint32_t)((u - 1) & 1));
        const uint32_t st = base + (uint32_t)sl * LF_STAGE;
        mbar_expect_tx(&full[sl], LF_STAGE);
        tma_load_3d(st, &mapZ, &full[sl], kb * 32, mt * 128, p.t);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tma_load_2d(st + LF_A_BYTES + g * (LF_UNITS * 128), &mapW, &full[sl], kb * 32, g * p.H + j0);
        if (++sl == LF_STAGES) { sl = 0; ++u; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, LF_N);
    int sl = 0, u = 0;
    for (int kb = 0; kb < p.nkb; ++kb) {
      mbar_wait(&full[sl], (uint32_t)(u & 1));
      tc_fence_after();
      if (leader) {
        const uint32_t st = base + (uint32_t)sl * LF_STAGE;
        const uint64_t ad = desc_k_sw128(st), bd = desc_k_sw128(st + LF_A_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_tf32(tmem, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&empty[sl]);
      }
      __syncwarp();
      if (++sl == LF_STAGES) { sl = 0; ++u; }
    }
    if (leader) umma_commit(&done);
  } else if (warp >= 4) {
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m = mt * 128 + quad * 32 + lane;
    const int jb = j0 + half * 8;                 // this thread's eight hidden units
    int nvalid = p.H - jb;
    nvalid = nvalid > 8 ? 8 : nvalid;
    const bool vec = (p.H & 3) == 0 && (p.D & 3) == 0;
    // operands that do not depend on the accumulator are fetched while the MMAs run
    float cp[8], bc[8], bf_[8], bi[8], bo[8];
    const long row = (long)m * p.H + jb;
    const long off = (long)p.t * p.B * p.H;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = m < p.B && i < nvalid;
      cp[i] = (ok && p.t > 0) ? __ldg(p.cache + off - (long)p.B * p.H + row + i) : 0.f;
      bc[i] = ok ? __ldg(p.bias + jb + i) : 0.f;
      bf_[i] = ok ? __ldg(p.bias + p.H + jb + i) : 0.f;
      bi[i] = ok ? __ldg(p.bias + 2 * p.H + jb + i) : 0.f;
      bo[i] = ok ? __ldg(p.bias + 3 * p.H + jb + i) : 0.f;
    }
    mbar_wait(&done, 0);
    tc_fence_after();
    uint32_t pc[8], pf[8], pi[8], po[8];
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 8);
    tmem_ld8(taddr, pc);
    tmem_ld8(taddr + LF_UNITS, pf);
    tmem_ld8(taddr + 2 * LF_UNITS, pi);
    tmem_ld8(taddr + 3 * LF_UNITS, po);
    tmem_ld_wait();
    if (m < p.B && nvalid > 0) {
      float c[8], ca[8], cv[8], fv[8], iv[8], ov[8], out[8], d0[8], d1[8], d2[8], d3[8], d4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float zc = __uint_as_float(pc[i]) + bc[i], zf = __uint_as_float(pf[i]) + bf_[i];
        const float zi = __uint_as_float(pi[i]) + bi[i], zo = __uint_as_float(po[i]) + bo[i];
        cv[i] = act_fwd_rt(p.act, zc);
        fv[i] = sigmoidf_(zf);
        iv[i] = sigmoidf_(zi);
        ov[i] = sigmoidf_(zo);
        c[i] = cp[i] * fv[i] + cv[i] * iv[i];       // recurrent_op.py:72
        ca[i] = act_fwd_rt(p.act, c[i]);
        out[i] = ca[i] * ov[i];                     // recurrent_op.py:74-75
        d0[i] = ov[i] * act_deriv_from_z(p.act, c[i], ca[i]);
        d1[i] = iv[i] * act_deriv_from_z(p.act, zc, cv[i]);
        d2[i] = cp[i] * (fv[i] * sigmoidf_(-zf));
        d3[i] = cv[i] * (iv[i] * sigmoidf_(-zi));
        d4[i] = ca[i] * (ov[i] * sigmoidf_(-zo));
      }
      float* q = p.cache + off + row;
      store8(q, c, nvalid, vec);
      store8(q + p.TBH, ca, nvalid, vec);
      store8(q + 2 * p.TBH, cv, nvalid, vec);
      store8(q + 3 * p.TBH, fv, nvalid, vec);
      store8(q + 4 * p.TBH, iv, nvalid, vec);
      store8(q + 5 * p.TBH, ov, nvalid, vec);
      store8(p.O + off + row, out, nvalid, vec);
      if (p.bw) {
        float* w = p.bw + off + row;
        store8(w, d0, nvalid, vec);
        store8(w + p.TBH, d1, nvalid, vec);
        store8(w + 2 * p.TBH, d2, nvalid, vec);
        store8(w + 3 * p.TBH, d3, nvalid, vec);
        store8(w + 4 * p.TBH, d4, nvalid, vec);
      }
      if (p.t + 1 < p.T)
        store8(p.Z + ((long)(p.t + 1) * p.B + m) * (p.D + p.H) + p.D + jb, out, nvalid, vec);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(LF_N));
}

// ---------------------------------------------------------------------------------------------------------------
// The whole forward recurrence as ONE persistent cooperative kernel.  Same tiling as the per-step kernel, one CTA per
// (128 batch rows, 16 hidden units) for all T steps:
//   * the CTA's slice of W^T (64 rows x (D+H)) is loaded once and stays in shared memory;
//   * the cell state of a (row, unit) pair lives in the registers of the thread that owns it;
//   * Z[t]'s input columns do not depend on the recurrence: their K blocks are prefetched and multiplied while step
//     t-1 is still finishing; only the K blocks holding O[t-1] wait;
//   * the wait is a release/acquire counter per (batch tile, step) shared by the CTAs of one batch tile -- a tile of
//     batch rows never waits for the rest of the grid.  The producer orders the acquire before its TMA reads with a
//     generic->async proxy fence;
//   * of the thirteen output rows only O[t] -> Z[t+1] is on the critical path; it is stored and published first.
constexpr int LS_RING = 6;

// The recurrence's critical path evaluates five transcendentals per (row, unit) and step; this kernel only runs on the
// TF32 path (5e-3 tolerance), so it uses the SFU forms: absolute error ~1e-7 for sigmoid, ~2e-7 for tanh.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_act(int act, float x) {
  switch (act) {
    case BF_ACT_SIGMOID: return fast_sigmoid(x);
    case BF_ACT_TANH: return 2.f * fast_sigmoid(2.f * x) - 1.f;
    case BF_ACT_RELU: return x > 0.f ? x : 0.f;
    default: return x;
  }
}

struct LstmSeqParams {
  int T, B, D, H, nkb, kb_dep, m_tiles, n_tiles, act;
  long TBH;
  const float* bias;
  float *cache, *O, *Z, *bw;
  unsigned* flags;      // [m_tiles][T], zeroed before the launch
  int tma_rows;         // the twelve bulk rows of a step (cache, O, bw) leave through TMA stores from a staged tile
};
constexpr uint32_t LS_ROW_TILE = 128 * LF_UNITS * 4;       // one staged [128 rows][16 units] tile of a bulk row

__global__ void __launch_bounds__(LF_THREADS, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapW,
                    const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapO,
                    const __grid_constant__ CUtensorMap mapBW, LstmSeqParams p) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t w_full, full[LS_RING], empty[LS_RING], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t wbase = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t ring = wbase + (uint32_t)p.nkb * LF_B_BYTES;
  const uint32_t rowst = ring + (uint32_t)LS_RING * LF_A_BYTES;          // two staged row tiles (tma_rows)
  const int nt = blockIdx.x / p.m_tiles, mt = blockIdx.x - nt * p.m_tiles;
  const int j0 = nt * LF_UNITS;

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(2 * LF_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < LS_RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(&w_full, (uint32_t)p.nkb * LF_B_BYTES);
      for (int kb = 0; kb < p.nkb; ++kb)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tma_load_2d(wbase + (uint32_t)kb * LF_B_BYTES + g * (LF_UNITS * 128), &mapW, &w_full, kb * 32, g * p.H + j0);
      int sl = 0, u = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (kb == p.kb_dep && t > 0) {
            const unsigned* flag = p.flags + (size_t)mt * p.T + (t - 1);
            unsigned v;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            } while (v < (unsigned)p.n_tiles);
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (u > 0) mbar_wait(&empty[sl], (uint32_t)((u - 1) & 1));
          mbar_expect_tx(&full[sl], LF_A_BYTES);
          tma_load_3d(ring + (uint32_t)sl * LF_A_BYTES, &mapZ, &full[sl], kb * 32, mt * 128, t);
          if (++sl == LS_RING) { sl = 0; ++u; }
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, LF_N);
    mbar_wait(&w_full, 0);
    int sl = 0, u = 0;
    for (int t = 0; t < p.T; ++t) {
      const int set = t & 1, tu = t >> 1;
      if (tu > 0) mbar_wait(&t_empty[set], (uint32_t)((tu - 1) & 1));
      tc_fence_after();
      const uint32_t acc = tmem + (uint32_t)set * LF_N;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full[sl], (uint32_t)(u & 1));
        tc_fence_after();
        if (leader) {
          const uint64_t ad = desc_k_sw128(ring + (uint32_t)sl * LF_A_BYTES), bd = desc_k_sw128(wbase + (uint32_t)kb * LF_B_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(acc, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty[sl]);
        }
        __syncwarp();
        if (++sl == LS_RING) { sl = 0; ++u; }
      }
      if (leader) umma_commit(&t_full[set]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m = mt * 128 + quad * 32 + lane;
    const int jb = j0 + half * 8;
    int nvalid = p.H - jb;
    nvalid = nvalid > 8 ? 8 : nvalid;
    const bool vec = (p.H & 3) == 0 && (p.D & 3) == 0;
    const bool live = m < p.B && nvalid > 0;
    const long row = (long)m * p.H + jb;
    float c[8], bc[8], bf_[8], bi[8], bo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = live && i < nvalid;
      c[i] = 0.f;
      bc[i] = ok ? __ldg(p.bias + jb + i) : 0.f;
      bf_[i] = ok ? __ldg(p.bias + p.H + jb + i) : 0.f;
      bi[i] = ok ? __ldg(p.bias + 2 * p.H + jb + i) : 0.f;
      bo[i] = ok ? __ldg(p.bias + 3 * p.H + jb + i) : 0.f;
    }
    uint32_t rowseq = 0;             // staged row tiles written so far (tma_rows)
    for (int t = 0; t < p.T; ++t) {
      const int set = t & 1, tu = t >> 1;
      mbar_wait(&t_full[set], (uint32_t)(tu & 1));
      tc_fence_after();
      uint32_t pc[8], pf[8], pi[8], po[8];
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * LF_N + half * 8);
      tmem_ld8(taddr, pc);
      tmem_ld8(taddr + LF_UNITS, pf);
      tmem_ld8(taddr + 2 * LF_UNITS, pi);
      tmem_ld8(taddr + 3 * LF_UNITS, po);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[set]);      // the accumulator is in registers: the MMA warp may reuse the set
      float ca[8], cv[8], fv[8], iv[8], ov[8], out[8], zc[8], zf[8], zi[8], zo[8], cprev[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        zc[i] = __uint_as_float(pc[i]) + bc[i]; zf[i] = __uint_as_float(pf[i]) + bf_[i];
        zi[i] = __uint_as_float(pi[i]) + bi[i]; zo[i] = __uint_as_float(po[i]) + bo[i];
        cv[i] = fast_act(p.act, zc[i]);
        fv[i] = fast_sigmoid(zf[i]);
        iv[i] = fast_sigmoid(zi[i]);
        ov[i] = fast_sigmoid(zo[i]);
        cprev[i] = c[i];
        c[i] = c[i] * fv[i] + cv[i] * iv[i];        // recurrent_op.py:72
        ca[i] = fast_act(p.act, c[i]);
        out[i] = ca[i] * ov[i];                     // recurrent_op.py:74-75
      }
      const long off = (long)t * p.B * p.H;
      if (t + 1 < p.T) {
        if (live) store8(p.Z + ((long)(t + 1) * p.B + m) * (p.D + p.H) + p.D + jb, out, nvalid, vec);
        // the barrier orders every epilogue thread's stores before thread 128's gpu-scope release (cumulativity):
        // no per-thread membar needed
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 128) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + (size_t)mt * p.T + t) : "memory");
      }
      if (p.tma_rows) {
        // EXPERIMENT (BF_LSTM_TMA_ROWS=1, parity-green, off by default): the twelve rows nobody waits for leave through the
        // async proxy -- staged 128 x 16 tile -> one TMA store per row -- so that they do not sit in front of the next step's
        // gpu-scope release.  Measured SLOWER on a B200 (cfg4 forward 706 vs 635 us): the 24 group barriers and staging
        // stores per step cost more than the release saves.
        float dv[8];
        const uint32_t mine = (uint32_t)((quad * 32 + lane) * LF_UNITS + half * 8) * 4u;
#define LS_PUT_ROW(VALS, MAP, ROWIDX, SEQ)                                                                              \
        {                                                                                                               \
          const uint32_t tile = rowst + (rowseq++ & 1u) * LS_ROW_TILE;     /* alternates across rows AND steps */        \
          if (tid == 128) tma_store_wait_read<1>();          /* the store that read this tile two rows ago is done */   \
          asm volatile("bar.sync 1, 256;" ::: "memory");                                                                \
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + mine), "f"(VALS[0]), "f"(VALS[1]), "f"(VALS[2]), "f"(VALS[3]) : "memory"); \
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + mine + 16u), "f"(VALS[4]), "f"(VALS[5]), "f"(VALS[6]), "f"(VALS[7]) : "memory"); \
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                                  \
          asm volatile("bar.sync 1, 256;" ::: "memory");                                                                \
          if (tid == 128) { tma_store_4d(&MAP, tile, j0, mt * 128, t, ROWIDX); tma_store_commit(); }                    \
        }
        LS_PUT_ROW(c, mapC, 0, 0) LS_PUT_ROW(ca, mapC, 1, 1) LS_PUT_ROW(cv, mapC, 2, 2) LS_PUT_ROW(fv, mapC, 3, 3)
        LS_PUT_ROW(iv, mapC, 4, 4) LS_PUT_ROW(ov, mapC, 5, 5) LS_PUT_ROW(out, mapO, 0, 6)
        if (p.bw) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = ov[i] * act_deriv_from_z(p.act, c[i], ca[i]);
          LS_PUT_ROW(dv, mapBW, 0, 7)
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = iv[i] * act_deriv_from_z(p.act, zc[i], cv[i]);
          LS_PUT_ROW(dv, mapBW, 1, 8)
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = cprev[i] * (fv[i] * sigmoidf_(-zf[i]));
          LS_PUT_ROW(dv, mapBW, 2, 9)
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = cv[i] * (iv[i] * sigmoidf_(-zi[i]));
          LS_PUT_ROW(dv, mapBW, 3, 10)
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = ca[i] * (ov[i] * sigmoidf_(-zo[i]));
          LS_PUT_ROW(dv, mapBW, 4, 11)
        }
#undef LS_PUT_ROW
      } else if (live) {
        float* q = p.cache + off + row;
        store8(q, c, nvalid, vec);
        store8(q + p.TBH, ca, nvalid, vec);
        store8(q + 2 * p.TBH, cv, nvalid, vec);
        store8(q + 3 * p.TBH, fv, nvalid, vec);
        store8(q + 4 * p.TBH, iv, nvalid, vec);
        store8(q + 5 * p.TBH, ov, nvalid, vec);
        store8(p.O + off + row, out, nvalid, vec);
        if (p.bw) {
          float d[8];
          float* w = p.bw + off + row;
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = ov[i] * act_deriv_from_z(p.act, c[i], ca[i]);
          store8(w, d, nvalid, vec);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = iv[i] * act_deriv_from_z(p.act, zc[i], cv[i]);
          store8(w + p.TBH, d, nvalid, vec);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = cprev[i] * (fv[i] * sigmoidf_(-zf[i]));
          store8(w + 2 * p.TBH, d, nvalid, vec);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = cv[i] * (iv[i] * sigmoidf_(-zi[i]));
          store8(w + 3 * p.TBH, d, nvalid, vec);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = ca[i] * (ov[i] * sigmoidf_(-zo[i]));
          store8(w + 4 * p.TBH, d, nvalid, vec);
        }
      }
    }
  }
  if (p.tma_rows && tid == 128) tma_store_wait<0>();       // every row store has left the staged tiles and landed
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * LF_N));
}

// ---------------------------------------------------------------------------------------------------------------
// The backward recurrence (BPTT) as ONE persistent cooperative kernel, the mirror image of lstm_seq_fwd_kernel.
// A CTA owns 64 batch rows x 32 hidden units for all T steps:
//   * its slice of W_h (rows D + j of W, 32 rows x 4H) is loaded once and stays in shared memory;
//   * deltaC of a (row, unit) pair lives in the registers of the thread that owns it;
//   * step t: the epilogue threads turn dH (their own accumulator columns of the previous GEMM) into the four gate
//     deltas of their units (recurrent_op.py:94-104, coefficients from `bw`), store them into dgates[t] and publish a
//     release/acquire counter per (batch tile, step); when all unit tiles of the batch tile have published, the
//     producer streams dgates[t] (64 rows x 4H) through a TMA ring and the MMA warp forms dH for step t-1;
//   * the input half of deltaZ (dX) is not on the recurrence: it is one large GEMM over all T*B rows afterwards.
constexpr int LB_ROWS = 64, LB_UNITS = 32, LB_RING = 8, LB_THREADS = 384;
constexpr uint32_t LB_A_BYTES = LB_ROWS * 128, LB_W_BYTES = LB_UNITS * 128;

struct LstmBwdParams {
  int T, B, D, H, nkb, m_tiles, n_tiles;
  long TBH;
  const float *E, *cache, *bw;
  float* dgates;
  unsigned* flags;      // [m_tiles][T], zeroed before the launch
};

__device__ __forceinline__ void load8(float (&v)[8], const float* p, int nvalid, bool vec, bool live) {
  if (live && vec && nvalid == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (live && i < nvalid) ? __ldg(p + i) : 0.f;
  }
}

__global__ void __launch_bounds__(LB_THREADS, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapW, LstmBwdParams p) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t w_full, full[LB_RING], empty[LB_RING], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t wbase = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t ring = wbase + (uint32_t)p.nkb * LB_W_BYTES;
  const int nt = blockIdx.x / p.m_tiles, mt = blockIdx.x - nt * p.m_tiles;
  const int j0 = nt * LB_UNITS;

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(4 * LB_UNITS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < LB_RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 2); mbar_init(&t_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(&w_full, (uint32_t)p.nkb * LB_W_BYTES);
      for (int kb = 0; kb < p.nkb; ++kb)
        tma_load_2d(wbase + (uint32_t)kb * LB_W_BYTES, &mapW, &w_full, kb * 32, p.D + j0);
      int sl = 0, u = 0;
      for (int t = p.T - 1; t >= 1; --t) {         // the GEMM of step t produces dH for step t-1
        const unsigned* flag = p.flags + (size_t)mt * p.T + t;
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        } while (v < (unsigned)p.n_tiles);
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (u > 0) mbar_wait(&empty[sl], (uint32_t)((u - 1) & 1));
          mbar_expect_tx(&full[sl], LB_A_BYTES);
          tma_load_3d(ring + (uint32_t)sl * LB_A_BYTES, &mapG, &full[sl], kb * 32, mt * LB_ROWS, t);
          if (++sl == LB_RING) { sl = 0; ++u; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // two issuers: a single thread issues one of these small MMAs every 50-100 cycles, 4 * nkb of them per step would be
    // the critical path.  Even K blocks accumulate into one 32-column accumulator, odd ones into a second; the epilogue
    // adds the two.
    const int which = warp == 1 ? 0 : 1;
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(LB_ROWS, LB_UNITS);
    mbar_wait(&w_full, 0);
    int sl = 0, u = 0, it = 0;
    for (int t = p.T - 1; t >= 1; --t, ++it) {
      const int set = it & 1, tu = it >> 1;
      if (tu > 0) mbar_wait(&t_empty[set], (uint32_t)((tu - 1) & 1));
      tc_fence_after();
      const uint32_t acc = tmem + (uint32_t)(set * 2 + which) * LB_UNITS;
      bool first = true;
      for (int kb = 0; kb < p.nkb; ++kb) {
        if ((kb & 1) == which) {
          mbar_wait(&full[sl], (uint32_t)(u & 1));
          tc_fence_after();
          if (leader) {
            const uint64_t ad = desc_k_sw128(ring + (uint32_t)sl * LB_A_BYTES), bd = desc_k_sw128(wbase + (uint32_t)kb * LB_W_BYTES);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32(acc, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (!first || ks > 0) ? 1u : 0u);
            umma_commit(&empty[sl]);
          }
          first = false;
          __syncwarp();
        }
        if (++sl == LB_RING) { sl = 0; ++u; }
      }
      if (leader) umma_commit(&t_full[set]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // M = 64 accumulators: row r sits in TMEM lane (r / 16) * 32 + r % 16 -> lanes 0..15 of each warp carry rows
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m = mt * LB_ROWS + quad * 16 + (lane & 15);
    const bool row_ok = lane < 16 && m < p.B;
    const bool vec = (p.H & 3) == 0;
    const int jb = j0 + half * 16;                 // this thread's 16 hidden units, as two chunks of 8
    int nv[2];
    nv[0] = p.H - jb; nv[0] = nv[0] > 8 ? 8 : (nv[0] < 0 ? 0 : nv[0]);
    nv[1] = p.H - jb - 8; nv[1] = nv[1] > 8 ? 8 : (nv[1] < 0 ? 0 : nv[1]);
    float dC[2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) dC[c][i] = 0.f;
    int it = 0;
    for (int t = p.T - 1; t >= 0; --t) {
      const long base = ((long)t * p.B + m) * p.H + jb;
      // operands of chunk 0 do not depend on the recurrence: fetched while the GEMM of step t+1 is still running
      float e[8], fv[8], k0[8], k1[8], k2[8], k3[8], k4[8];
      {
        const bool live = row_ok && nv[0] > 0;
        load8(e, p.E + base, nv[0], vec, live);
        load8(fv, p.cache + 3 * p.TBH + base, nv[0], vec, live);
        load8(k0, p.bw + base, nv[0], vec, live);
        load8(k1, p.bw + p.TBH + base, nv[0], vec, live);
        load8(k2, p.bw + 2 * p.TBH + base, nv[0], vec, live);
        load8(k3, p.bw + 3 * p.TBH + base, nv[0], vec, live);
        load8(k4, p.bw + 4 * p.TBH + base, nv[0], vec, live);
      }
      uint32_t dh[16];
      if (t < p.T - 1) {
        const int set = it & 1, tu = it >> 1;
        mbar_wait(&t_full[set], (uint32_t)(tu & 1));
        tc_fence_after();
        uint32_t dh1[16];
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * 2 * LB_UNITS + half * 16), dh);
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)((set * 2 + 1) * LB_UNITS + half * 16), dh1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) dh[i] = __float_as_uint(__uint_as_float(dh[i]) + (p.nkb > 1 ? __uint_as_float(dh1[i]) : 0.f));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[set]);
        ++it;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) dh[i] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float e2[8], fv2[8], q0[8], q1[8], q2[8], q3[8], q4[8];
        if (c == 0) {                                   // issue chunk 1's loads before chunk 0's arithmetic and stores
          const bool live = row_ok && nv[1] > 0;
          load8(e2, p.E + base + 8, nv[1], vec, live);
          load8(fv2, p.cache + 3 * p.TBH + base + 8, nv[1], vec, live);
          load8(q0, p.bw + base + 8, nv[1], vec, live);
          load8(q1, p.bw + p.TBH + base + 8, nv[1], vec, live);
          load8(q2, p.bw + 2 * p.TBH + base + 8, nv[1], vec, live);
          load8(q3, p.bw + 3 * p.TBH + base + 8, nv[1], vec, live);
          load8(q4, p.bw + 4 * p.TBH + base + 8, nv[1], vec, live);
        }
        float g0[8], g1[8], g2[8], g3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ev = e[i] + __uint_as_float(dh[c * 8 + i]);
          const float d = dC[c][i] + ev * k0[i];
          g0[i] = d * k1[i];
          g1[i] = d * k2[i];
          g2[i] = d * k3[i];
          g3[i] = ev * k4[i];
          dC[c][i] = d * fv[i];
        }
        if (row_ok && nv[c] > 0) {
          float* dg = p.dgates + ((long)t * p.B + m) * 4 * p.H + jb + c * 8;
          store8(dg, g0, nv[c], vec);
          store8(dg + p.H, g1, nv[c], vec);
          store8(dg + 2 * p.H, g2, nv[c], vec);
          store8(dg + 3 * p.H, g3, nv[c], vec);
        }
        if (c == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { e[i] = e2[i]; fv[i] = fv2[i]; k0[i] = q0[i]; k1[i] = q1[i]; k2[i] = q2[i]; k3[i] = q3[i]; k4[i] = q4[i]; }
        }
      }
      if (t >= 1) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 128) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + (size_t)mt * p.T + t) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(4 * LB_UNITS));
}

// BF_LSTM_PERSIST_BWD=0 keeps the per-step backward (gate kernel + GEMM per time step)
static inline bool lstm_persistent_bwd_enabled() {
  const char* e = getenv("BF_LSTM_PERSIST_BWD");
  return e == nullptr || e[0] != '0';
}

struct EpiStoreLd {
  float* out; long ld;
  __device__ void store(int m, int n, float v, int) const { out[(long)m * ld + n] = v; }
};

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// dst (C x R) = src (R x C)^T through a padded shared-memory tile
__global__ void lstm_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ty + j, c = c0 + tx;
    if (r < R && c < C) tile[ty + j][tx] = src[(size_t)r * C + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, r = r0 + tx;
    if (r < R && c < C) dst[(size_t)c * R + r] = tile[tx][ty + j];
  }
}

// ---- plain recurrent layer (atomic/recurrent_op.py:9-43): O[t] = act([X[t], O[t-1]] @ W + b)
struct EpiRnnFwd {
  float* O; float* Znext; const float* b; int D, H, act;
  __device__ void store(int m, int n, float v, int) const {
    const float a = act_fwd_rt(act, v + __ldg(b + n));
    O[(long)m * H + n] = a;
    if (Znext) Znext[(long)m * (D + H) + D + n] = a;
  }
};
// E'[t-1] = (E[t-1] + deltaZ[t][:, D:]) * act'(O[t-1])  -- recurrent_op.py:35-38 with the next step's product folded in
struct EpiRnnBwd {
  const float* Eprev; const float* Oprev; float* out; int H, act;
  __device__ void store(int m, int n, float v, int) const {
    const long i = (long)m * H + n;
    out[i] = (__ldg(Eprev + i) + v) * act_bwd_rt(act, __ldg(Oprev + i));
  }
};
__global__ void rnn_scale_kernel(const float* __restrict__ E, const float* __restrict__ O, float* __restrict__ out, long n, int act) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = E[i] * act_bwd_rt(act, O[i]);
}

// ---- GRU (layers/recurrent.py:116-190; the reference keeps it at layer level, no atomic op).  W (D+H, 3H), columns [U|R|O].
// gate GEMM epilogue: U, R = sigmoid(Z @ Wur + bur);  Zk[:, D:] = R * h_prev   (recurrent.py:137-140)
struct EpiGruUR {
  float *U, *R, *Zk; const float* Z; const float* b; int D, H;
  __device__ void store(int m, int n, float v, int) const {
    const float a = sigmoidf_(v + __ldg(b + n));
    if (n < H) {
      U[(long)m * H + n] = a;
    } else {
      const int j = n - H;
      R[(long)m * H + j] = a;
      Zk[(long)m * (D + H) + D + j] = a * __ldg(Z + (long)m * (D + H) + D + j);
    }
  }
};
// candidate GEMM epilogue: O = act(Zk @ Wo + bo);  h = U h_prev + (1 - U) O   (recurrent.py:141-142)
struct EpiGruO {
  float *O, *hs, *Znext; const float *U, *Z, *b; int D, H, act;
  __device__ void store(int m, int n, float v, int) const {
    const long i = (long)m * H + n;
    const float o = act_fwd_rt(act, v + __ldg(b + n));
    const float u = __ldg(U + i), hp = __ldg(Z + (long)m * (D + H) + D + n);
    const float h = u * hp + (1.f - u) * o;
    O[i] = o; hs[i] = h;
    if (Znext) Znext[(long)m * (D + H) + D + n] = h;
  }
};
// dh += delta[t];  dU = (h_prev - O) U(1-U) dh;  dO = (1-U) act'(O) dh   (recurrent.py:170-172)
__global__ void gru_bwd_pre_kernel(const float* __restrict__ delta, float* __restrict__ dh, const float* __restrict__ Z,
                                   const float* __restrict__ U, const float* __restrict__ O, float* __restrict__ dUR,
                                   float* __restrict__ dO, int B, int D, int H, int act) {
  long total = (long)B * H;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const long m = i / H;
    const int j = (int)(i - m * H);
    const float d = dh[i] + delta[i], u = U[i], o = O[i], hp = Z[m * (D + H) + D + j];
    dUR[m * 2 * H + j] = (hp - o) * (u * (1.f - u)) * d;
    dO[i] = (1.f - u) * act_bwd_rt(act, o) * d;
    dh[i] = d;
  }
}
// dZk = dO @ Wo^T:  columns < D start dX[t];  dK = columns >= D:  dR = h_prev R(1-R) dK,  keep R dK   (recurrent.py:173-175,181)
struct EpiGruDzk {
  float *dX, *dUR, *RdK; const float *Z, *R; int D, H;
  __device__ void store(int m, int n, float v, int) const {
    if (n < D) {
      dX[(long)m * D + n] = v;
    } else {
      const int j = n - D;
      const long i = (long)m * H + j;
      const float r = __ldg(R + i), hp = __ldg(Z + (long)m * (D + H) + n);
      dUR[(long)m * 2 * H + H + j] = hp * (r * (1.f - r)) * v;
      RdK[i] = r * v;
    }
  }
};
// dZ = [dU, dR] @ Wur^T:  dX[t] += columns < D;  dh = dZ[:, D:] + R dK + U dh   (recurrent.py:176,181-182)
struct EpiGruDz {
  float *dX, *dh; const float *RdK, *U; int D, H;
  __device__ void store(int m, int n, float v, int) const {
    if (n < D) {
      dX[(long)m * D + n] += v;
    } else {
      const long i = (long)m * H + (n - D);
      dh[i] = v + __ldg(RdK + i) + __ldg(U + i) * dh[i];
    }
  }
};

// ---- Clockwork layer (layers/recurrent.py:193-283): a plain recurrence whose output columns only update on their ticks.
// gate(t, n) = (t % tick[n] == 0) for t = 1..T  (:242; a zero tick never fires, like NumPy's nan remainder)
__device__ __forceinline__ bool cw_gate(const float* tick, int n, int t) { return fmodf((float)t, __ldg(tick + n)) == 0.f; }
// output = act(Z @ (W * gate) + b * gate)   (:243-246): gated-off columns see a zero pre-activation
struct EpiCwFwd {
  float* O; float* Znext; const float *b, *tick; int D, H, act, t;
  __device__ void store(int m, int n, float v, int) const {
    const float a = act_fwd_rt(act, cw_gate(tick, n, t) ? v + __ldg(b + n) : 0.f);
    O[(long)m * H + n] = a;
    if (Znext) Znext[(long)m * (D + H) + D + n] = a;
  }
};
// G[t] = (dh + delta[t]) * act'(O[t]) * gate_t  -- the quantity every use of dh in :266-273 reduces to
__global__ void cw_first_kernel(const float* __restrict__ delta, const float* __restrict__ O, const float* __restrict__ tick,
                                float* __restrict__ G, long n, int H, int act, int t) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) G[i] = cw_gate(tick, (int)(i % H), t) ? delta[i] * act_bwd_rt(act, O[i]) : 0.f;
}
// deltaZ = G[t] @ W^T:  columns < D are dX[t];  the rest is dh for step t-1, turned into G[t-1] on the spot   (:273-275)
struct EpiCwBwd {
  float* dX; const float *delta_prev, *O_prev, *tick; float* G_prev; int D, H, act, t_prev;
  __device__ void store(int m, int n, float v, int) const {
    if (n < D) {
      dX[(long)m * D + n] = v;
    } else if (G_prev) {
      const int j = n - D;
      const long i = (long)m * H + j;
      G_prev[i] = cw_gate(tick, j, t_prev) ? (v + __ldg(delta_prev + i)) * act_bwd_rt(act, __ldg(O_prev + i)) : 0.f;
    }
  }
};

}  // namespace bf

using namespace bf;

extern "C" {

static int lstm_forward_impl(const float* X, const float* W, const float* b, float* O, float* Z, float* cache, float* bw,
                             int T, int B, int D, int H, int act, int precision, void* stream, int x_batch_major) {
  BF_REQUIRE(X && W && b && O && Z && cache, BF_ERR_ARG, "bf_lstm_forward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_lstm_forward: bad shape");
  BF_REQUIRE(act == BF_ACT_TANH || act == BF_ACT_RELU || act == BF_ACT_SIGMOID || act == BF_ACT_LINEAR, BF_ERR_ARG,
             "bf_lstm_forward: bad activation %d", act);
  if (T == 0 || B == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H, ZD = D + H;
  void* ws = nullptr;
  const size_t gates_bytes = align_up((size_t)B * 4 * H * sizeof(float));
  const size_t wt_bytes = align_up((size_t)ZD * 4 * H * sizeof(float));
  const size_t flag_bytes = (size_t)((B + 127) / 128) * T * sizeof(unsigned);
  int rc = workspace_require(gates_bytes + wt_bytes + flag_bytes, &ws);
  if (rc) return rc;
  float* gates = (float*)ws;
  float* Wt = (float*)((char*)ws + gates_bytes);
  unsigned* flags = (unsigned*)((char*)ws + gates_bytes + wt_bytes);     // W^T (4H x (D+H)): the K-major B operand of the TMA-fed gate GEMM
  const bool tma = precision == BF_PREC_TF32 && gemm_tma_ok(Z, Wt, B, 4 * H, (int)ZD) && ((ZD * 4) % 16) == 0 &&
                   getenv("BF_LSTM_GATHER") == nullptr;
  if (tma) {
    lstm_transpose_kernel<<<dim3((4 * H + 31) / 32, ((int)ZD + 31) / 32), dim3(32, 8), 0, s>>>(W, Wt, (int)ZD, 4 * H);
    BF_LAUNCH_CHECK();
  }
  lstm_fill_z_kernel<<<ew_grid(TB * D + BH, 256), 256, 0, s>>>(X, Z, TB, B, D, H, T, x_batch_major);
  BF_LAUNCH_CHECK();
  float* C = cache;               // cache (6,T,B,H): C, Ca, cand, f, i, o  (recurrent_op.py:77)
  float* Ca = cache + 1 * TB * H;
  float* cand = cache + 2 * TB * H;
  float* f = cache + 3 * TB * H;
  float* ig = cache + 4 * TB * H;
  float* o = cache + 5 * TB * H;
  const int K = D + H, N = 4 * H;
  // whole step in one kernel
  if (tma && getenv("BF_LSTM_UNFUSED") == nullptr) {
    CUtensorMap mZ, mW;
    {
      uint64_t dims[3] = {(uint64_t)ZD, (uint64_t)B, (uint64_t)T};
      uint64_t str[2] = {(uint64_t)ZD, (uint64_t)B * ZD};
      uint32_t box[3] = {32, 128, 1};
      if (!tma_make_map(&mZ, Z, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "bf_lstm_forward: cuTensorMapEncodeTiled failed (Z)");
    }
    {
      uint64_t dims[2] = {(uint64_t)ZD, (uint64_t)N};
      uint64_t str[1] = {(uint64_t)ZD};
      uint32_t box[2] = {32, LF_UNITS};
      if (!tma_make_map(&mW, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "bf_lstm_forward: cuTensorMapEncodeTiled failed (W)");
    }
    {
      // persistent form: every tile resident at once (one CTA per SM), W slice + A ring within shared memory
      const int nkb = (K + 31) / 32, m_tiles = (B + 127) / 128, n_tiles = (H + LF_UNITS - 1) / LF_UNITS;
      // bulk rows through TMA stores: needs 16-byte row strides and room for two staged row tiles
      CUtensorMap mC, mO, mBW;
      memset(&mC, 0, sizeof(mC)); memset(&mO, 0, sizeof(mO)); memset(&mBW, 0, sizeof(mBW));
      bool tma_rows = (H % 4) == 0 && getenv("BF_LSTM_TMA_ROWS") != nullptr &&   // opt-in: measured slower (see the kernel)
                      1024 + (size_t)nkb * LF_B_BYTES + (size_t)LS_RING * LF_A_BYTES + 2 * LS_ROW_TILE <= 224 * 1024;
      if (tma_rows) {
        uint64_t str[3] = {(uint64_t)H, (uint64_t)B * H, (uint64_t)TB * H};
        uint32_t box[4] = {LF_UNITS, 128, 1, 1};
        uint64_t dc[4] = {(uint64_t)H, (uint64_t)B, (uint64_t)T, 6}, dO[4] = {(uint64_t)H, (uint64_t)B, (uint64_t)T, 1},
                 db[4] = {(uint64_t)H, (uint64_t)B, (uint64_t)T, 5};
        tma_rows = tma_make_map(&mC, cache, 4, dc, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, false) &&
                   tma_make_map(&mO, O, 4, dO, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, false) &&
                   (bw == nullptr || tma_make_map(&mBW, bw, 4, db, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, false));
      }
      const size_t smem = 1024 + (size_t)nkb * LF_B_BYTES + (size_t)LS_RING * LF_A_BYTES + (tma_rows ? 2 * LS_ROW_TILE : 0);
      bool fits = m_tiles * n_tiles <= num_sms() && smem <= 224 * 1024 && getenv("BF_LSTM_STEPWISE") == nullptr;
      if (fits) {
        // the flag protocol needs every CTA resident: ask the runtime instead of assuming one CTA per SM is available
        BF_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_fwd_kernel, LF_THREADS, smem));
        fits = (long)per_sm * num_sms() >= (long)m_tiles * n_tiles;
      }
      if (fits) {
        LstmSeqParams q;
        q.T = T; q.B = B; q.D = D; q.H = H; q.nkb = nkb; q.kb_dep = D / 32; q.m_tiles = m_tiles; q.n_tiles = n_tiles;
        q.act = act; q.TBH = TB * H; q.bias = b; q.cache = cache; q.O = O; q.Z = Z; q.bw = bw; q.flags = flags;
        q.tma_rows = tma_rows ? 1 : 0;
        BF_CUDA(cudaMemsetAsync(flags, 0, flag_bytes, s));
        void* args[] = {(void*)&mZ, (void*)&mW, (void*)&mC, (void*)&mO, (void*)&mBW, (void*)&q};
        BF_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_seq_fwd_kernel, dim3(m_tiles * n_tiles), dim3(LF_THREADS), args, smem, s));
        count_launch();
        set_last_path(2);
        return BF_OK;
      }
    }
    LstmFwdParams p;
    p.T = T; p.B = B; p.D = D; p.H = H; p.nkb = (K + 31) / 32; p.m_tiles = (B + 127) / 128; p.act = act;
    p.TBH = TB * H; p.bias = b; p.cache = cache; p.O = O; p.Z = Z; p.bw = bw;
    const int n_tiles = (H + LF_UNITS - 1) / LF_UNITS;
    const size_t smem = 1024 + (size_t)LF_STAGES * LF_STAGE;
    BF_CUDA(cudaFuncSetAttribute(lstm_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int t = 0; t < T; ++t) {
      p.t = t;
      lstm_step_fwd_kernel<<<p.m_tiles * n_tiles, LF_THREADS, smem, s>>>(mZ, mW, p);
      BF_LAUNCH_CHECK();
    }
    set_last_path(2);
    return BF_OK;
  }
  for (int t = 0; t < T; ++t) {
    EpiBias ep{gates, b, N};
    if (tma) {
      rc = launch_gemm_tma(Z + (long)t * B * ZD, Wt, B, N, K, ep, s);
    } else {
      LoadKContig al{Z + (long)t * B * ZD, (long)ZD};
      LoadMNContig bl{W, (long)N};
      rc = launch_gemm(precision, al, bl, ep, B, N, K, 1, (K + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
    long off = (long)t * BH;
    lstm_gates_fwd_kernel<<<ew_grid(BH, 256), 256, 0, s>>>(
        gates, t ? C + off - BH : nullptr, C + off, Ca + off, cand + off, f + off, ig + off, o + off, O + off,
        t + 1 < T ? Z + (long)(t + 1) * B * ZD : nullptr, bw ? bw + off : nullptr, TB * H, B, D, H, act);
    BF_LAUNCH_CHECK();
  }
  return BF_OK;
}

int bf_lstm_forward(const float* X, const float* W, const float* b, float* O, float* Z, float* cache, float* bw,
                    int T, int B, int D, int H, int act, int precision, void* stream) {
  return lstm_forward_impl(X, W, b, O, Z, cache, bw, T, B, D, H, act, precision, stream, 0);
}

int bf_lstm_forward_bm(const float* X, const float* W, const float* b, float* O, float* Z, float* cache, float* bw,
                       int T, int B, int D, int H, int act, int precision, void* stream) {
  return lstm_forward_impl(X, W, b, O, Z, cache, bw, T, B, D, H, act, precision, stream, 1);
}

int bf_lstm_backward(const float* Z, const float* O, const float* E, const float* W, const float* cache,
                     const float* bw, float* dX, float* nablaW, float* nablab, int T, int B, int D, int H, int act,
                     int precision, void* stream) {
  (void)O;
  BF_REQUIRE(Z && E && W && cache && nablaW && nablab, BF_ERR_ARG, "bf_lstm_backward: null pointer");   // dX may be NULL
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_lstm_backward: bad shape");
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H, N4 = 4 * H;
  if (T == 0 || B == 0) {
    BF_CUDA(cudaMemsetAsync(nablaW, 0, (size_t)ZD * N4 * sizeof(float), s));
    BF_CUDA(cudaMemsetAsync(nablab, 0, (size_t)N4 * sizeof(float), s));
    return BF_OK;
  }
  // workspace: dgates (T,B,4H) | dH (B,H) | deltaC (B,H) | split-K partials for nablaW
  const int rows = ZD + 1;
  int bm = rows <= 32 ? 32 : (rows <= 64 ? 64 : 128), bn = N4 <= 16 ? 16 : (N4 <= 32 ? 32 : (N4 <= 64 ? 64 : 128));
  int tiles = ((rows + bm - 1) / bm) * ((N4 + bn - 1) / bn);
  BF_REQUIRE(TB < 2147483647L, BF_ERR_ARG, "bf_lstm_backward: T*B too large");
  int kps = 0;
  int splits = choose_splits(tiles, (int)TB, &kps);
  size_t off_dg = 0;
  size_t off_dh = align_up(off_dg + (size_t)TB * N4 * sizeof(float));
  size_t off_dc = align_up(off_dh + (size_t)BH * sizeof(float));
  size_t off_pw = align_up(off_dc + (size_t)BH * sizeof(float));
  // Z and dgates are both stored with the contraction index (t, b) slowest: the MN-major TMA kernel reads them in place
  const bool use_mn = precision == BF_PREC_TF32 && getenv("BF_LSTM_GATHER") == nullptr && gemm_tma_mn_ok(Z, nullptr, ZD, N4, TB);
  int mn_splits = num_sms() / (((ZD + 127) / 128) * ((N4 + 127) / 128));
  if (mn_splits < 1) mn_splits = 1;
  const size_t partial_bytes = (size_t)(use_mn ? mn_splits : splits) * rows * N4 * sizeof(float);
  size_t need = off_pw + partial_bytes;
  void* ws = nullptr;
  int rc = workspace_require(need, &ws);
  if (rc) return rc;
  char* base = (char*)ws;
  float* dgates = (float*)(base + off_dg);
  float* dH = (float*)(base + off_dh);
  float* deltaC = (float*)(base + off_dc);
  float* partial = (float*)(base + off_pw);
  const float* C = cache;
  const float* Ca = cache + 1 * TB * H;
  const float* cand = cache + 2 * TB * H;
  const float* f = cache + 3 * TB * H;
  const float* ig = cache + 4 * TB * H;
  const float* o = cache + 5 * TB * H;
  // persistent form: needs the coefficient rows `bw`, every tile resident at once, W_h slice + ring within shared memory
  bool persistent_done = false;
  if (precision == BF_PREC_TF32 && bw && T >= 1 && (N4 & 3) == 0 && (H & 3) == 0 && ((uintptr_t)W & 15) == 0 && tma_encode_fn() != nullptr &&
      getenv("BF_LSTM_GATHER") == nullptr && getenv("BF_LSTM_STEPWISE") == nullptr && lstm_persistent_bwd_enabled()) {
    const int nkb = (N4 + 31) / 32, m_tiles = (B + LB_ROWS - 1) / LB_ROWS, n_tiles = (H + LB_UNITS - 1) / LB_UNITS;
    const size_t smem = 1024 + (size_t)nkb * LB_W_BYTES + (size_t)LB_RING * LB_A_BYTES;
    const size_t flag_bytes = (size_t)m_tiles * T * sizeof(unsigned);
    bool fits = m_tiles * n_tiles <= num_sms() && smem <= 220 * 1024 && flag_bytes <= (size_t)BH * sizeof(float);
    if (fits) {
      BF_CUDA(cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      BF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_bwd_kernel, LB_THREADS, smem));
      fits = (long)per_sm * num_sms() >= (long)m_tiles * n_tiles;
    }
    if (fits) {
      CUtensorMap mG, mW;
      {
        uint64_t dims[3] = {(uint64_t)N4, (uint64_t)B, (uint64_t)T};
        uint64_t str[2] = {(uint64_t)N4, (uint64_t)B * N4};
        uint32_t box[3] = {32, LB_ROWS, 1};
        if (!tma_make_map(&mG, dgates, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "bf_lstm_backward: cuTensorMapEncodeTiled failed (dgates)");
      }
      {
        uint64_t dims[2] = {(uint64_t)N4, (uint64_t)ZD};
        uint64_t str[1] = {(uint64_t)N4};
        uint32_t box[2] = {32, LB_UNITS};
        if (!tma_make_map(&mW, W, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return fail(BF_ERR_CUDA, "bf_lstm_backward: cuTensorMapEncodeTiled failed (W)");
      }
      unsigned* flags = (unsigned*)dH;          // the dH scratch is unused on this path
      BF_CUDA(cudaMemsetAsync(flags, 0, flag_bytes, s));
      LstmBwdParams q;
      q.T = T; q.B = B; q.D = D; q.H = H; q.nkb = nkb; q.m_tiles = m_tiles; q.n_tiles = n_tiles; q.TBH = TB * H;
      q.E = E; q.cache = cache; q.bw = bw; q.dgates = dgates; q.flags = flags;
      void* args[] = {(void*)&mG, (void*)&mW, (void*)&q};
      BF_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_seq_bwd_kernel, dim3(m_tiles * n_tiles), dim3(LB_THREADS), args, smem, s));
      count_launch();
      // dX = dgates @ W[:D]^T over all T*B rows at once: B(n, k) = W[n * 4H + k]
      EpiStoreLd ep{dX, (long)D};
      if (dX == nullptr) {
        rc = BF_OK;            // the caller discards the input delta (first layer inside learn_batch)
      } else if (gemm_tma_ok(dgates, W, (int)TB, D, N4)) {
        rc = launch_gemm_tma(dgates, W, (int)TB, D, N4, ep, s);
      } else {
        LoadKContig al{dgates, (long)N4};
        LoadKContig bl{W, (long)N4};
        rc = launch_gemm(precision, al, bl, ep, (int)TB, D, N4, 1, (N4 + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
      }
      if (rc) return rc;
      persistent_done = true;
    }
  }
  for (int t = T - 1; t >= 0 && !persistent_done; --t) {
    long off = (long)t * BH;
    float* dg_t = dgates + (long)t * B * N4;
    lstm_gates_bwd_kernel<<<ew_grid(BH, 256), 256, 0, s>>>(E + off, t == T - 1 ? nullptr : dH,
                                                          t ? C + off - BH : nullptr, Ca + off, cand + off, f + off,
                                                          ig + off, o + off, bw ? bw + off : nullptr, TB * H, deltaC, dg_t, B, H,
                                                          act, t == T - 1);
    BF_LAUNCH_CHECK();
    // deltaZ[t] = dgates[t] @ W^T : M=B, N=D+H, K=4H ; B(k, n) = W[n*4H + k]
    EpiLstmDz ep{dX ? dX + (long)t * B * D : nullptr, dH, D, H};
    if (precision == BF_PREC_TF32 && gemm_tma_ok(dg_t, W, B, ZD, N4) && getenv("BF_LSTM_GATHER") == nullptr) {
      rc = launch_gemm_tma(dg_t, W, B, ZD, N4, ep, s);     // both operands are K-major as stored: no copy, no transpose
    } else {
      LoadKContig al{dg_t, (long)N4};
      LoadKContig bl{W, (long)N4};
      rc = launch_gemm(precision, al, bl, ep, B, ZD, N4, 1, (N4 + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
  }
  // nablaW = sum_t Z[t]^T @ dgates[t]  (+ ones row -> nablab)
  if (use_mn) {
    int used = 1;
    rc = launch_gemm_tma_mn(Z, dgates, ZD, N4, (int)TB, partial, partial_bytes, &used, s);
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, nablaW, nablab, rows, N4, ZD, used, 0, N4, 1, s);
    if (rc) return rc;
  } else {
    LoadMNContigOnes al{Z, (long)ZD, ZD};
    LoadMNContig bl{dgates, (long)N4};
    EpiPartial ep{partial, (long)rows * N4, N4};
    rc = launch_gemm(precision, al, bl, ep, rows, N4, (int)TB, splits, kps, 0, s);
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, nablaW, nablab, rows, N4, ZD, splits, 0, N4, 1, s);
    if (rc) return rc;
  }
  return BF_OK;
}

// Plain recurrent layer, atomic/recurrent_op.py:14-27.  X (T,B,D), W (D+H,H), b (H) -> O (T,B,H), Z (T,B,D+H).
int bf_rnn_forward(const float* X, const float* W, const float* b, float* O, float* Z, int T, int B, int D, int H,
                   int act, int precision, void* stream) {
  BF_REQUIRE(X && W && b && O && Z, BF_ERR_ARG, "bf_rnn_forward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_rnn_forward: bad shape");
  BF_REQUIRE(act == BF_ACT_TANH || act == BF_ACT_RELU || act == BF_ACT_SIGMOID || act == BF_ACT_LINEAR, BF_ERR_ARG,
             "bf_rnn_forward: bad activation %d", act);
  if (T == 0 || B == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H;
  void* ws = nullptr;
  int rc = workspace_require((size_t)ZD * H * sizeof(float), &ws);
  if (rc) return rc;
  float* Wt = (float*)ws;
  const bool tma = precision == BF_PREC_TF32 && gemm_tma_ok(Z, Wt, B, H, ZD) && getenv("BF_LSTM_GATHER") == nullptr;
  if (tma) {
    lstm_transpose_kernel<<<dim3((H + 31) / 32, (ZD + 31) / 32), dim3(32, 8), 0, s>>>(W, Wt, ZD, H);
    BF_LAUNCH_CHECK();
  }
  lstm_fill_z_kernel<<<ew_grid(TB * D + BH, 256), 256, 0, s>>>(X, Z, TB, B, D, H);
  BF_LAUNCH_CHECK();
  for (int t = 0; t < T; ++t) {
    EpiRnnFwd ep{O + (long)t * BH, t + 1 < T ? Z + (long)(t + 1) * B * ZD : nullptr, b, D, H, act};
    if (tma) {
      rc = launch_gemm_tma(Z + (long)t * B * ZD, Wt, B, H, ZD, ep, s);
    } else {
      LoadKContig al{Z + (long)t * B * ZD, (long)ZD};
      LoadMNContig bl{W, (long)H};
      rc = launch_gemm(precision, al, bl, ep, B, H, ZD, 1, (ZD + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
  }
  return BF_OK;
}

// atomic/recurrent_op.py:29-43.  Eout (T,B,H) receives the reference's in-place modified E (its `dX` is the slice
// Eout[:, :, :D], taken by the caller); nablaW (D+H,H), nablab (H).
int bf_rnn_backward(const float* Z, const float* O, const float* E, const float* W, float* Eout, float* nablaW,
                    float* nablab, int T, int B, int D, int H, int act, int precision, void* stream) {
  BF_REQUIRE(Z && O && E && W && Eout && nablaW && nablab, BF_ERR_ARG, "bf_rnn_backward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_rnn_backward: bad shape");
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H;
  if (T == 0 || B == 0) {
    BF_CUDA(cudaMemsetAsync(nablaW, 0, (size_t)ZD * H * sizeof(float), s));
    BF_CUDA(cudaMemsetAsync(nablab, 0, (size_t)H * sizeof(float), s));
    return BF_OK;
  }
  BF_REQUIRE(TB < 2147483647L, BF_ERR_ARG, "bf_rnn_backward: T*B too large");
  const int rows = ZD + 1;
  int bm = rows <= 32 ? 32 : (rows <= 64 ? 64 : 128), bn = H <= 16 ? 16 : (H <= 32 ? 32 : (H <= 64 ? 64 : 128));
  int tiles = ((rows + bm - 1) / bm) * ((H + bn - 1) / bn);
  int kps = 0;
  int splits = choose_splits(tiles, (int)TB, &kps);
  const bool use_mn = precision == BF_PREC_TF32 && getenv("BF_LSTM_GATHER") == nullptr && gemm_tma_mn_ok(Z, Eout, ZD, H, TB);
  int mn_splits = num_sms() / (((ZD + 127) / 128) * ((H + 127) / 128));
  if (mn_splits < 1) mn_splits = 1;
  const size_t partial_bytes = (size_t)(use_mn ? mn_splits : splits) * rows * H * sizeof(float);
  void* ws = nullptr;
  int rc = workspace_require(partial_bytes, &ws);
  if (rc) return rc;
  float* partial = (float*)ws;
  const float* Wh = W + (long)D * H;      // rows D.. of W: B(n, k) = W[(D + n) * H + k], K-major as stored
  rnn_scale_kernel<<<ew_grid(BH, 256), 256, 0, s>>>(E + (T - 1) * BH, O + (T - 1) * BH, Eout + (T - 1) * BH, BH, act);
  BF_LAUNCH_CHECK();
  for (int t = T - 1; t >= 1; --t) {
    const float* Et = Eout + (long)t * BH;
    EpiRnnBwd ep{E + (long)(t - 1) * BH, O + (long)(t - 1) * BH, Eout + (long)(t - 1) * BH, H, act};
    if (precision == BF_PREC_TF32 && gemm_tma_ok(Et, Wh, B, H, H) && getenv("BF_LSTM_GATHER") == nullptr) {
      rc = launch_gemm_tma(Et, Wh, B, H, H, ep, s);
    } else {
      LoadKContig al{Et, (long)H};
      LoadKContig bl{Wh, (long)H};
      rc = launch_gemm(precision, al, bl, ep, B, H, H, 1, (H + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
  }
  // nablaW = sum_t Z[t]^T @ E'[t], nablab = column sums of E'  (recurrent_op.py:41-42)
  if (use_mn) {
    int used = 1;
    rc = launch_gemm_tma_mn(Z, Eout, ZD, H, (int)TB, partial, partial_bytes, &used, s);
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, nablaW, nablab, rows, H, ZD, used, 0, H, 1, s);
  } else {
    LoadMNContigOnes al{Z, (long)ZD, ZD};
    LoadMNContig bl{Eout, (long)H};
    EpiPartial ep{partial, (long)rows * H, H};
    rc = launch_gemm(precision, al, bl, ep, rows, H, (int)TB, splits, kps, 0, s);
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, nablaW, nablab, rows, H, ZD, splits, 0, H, 1, s);
  }
  return rc;
}

// GRU layer forward, layers/recurrent.py:123-153.  X (T,B,D) time-major, W (D+H,3H) columns [U|R|O], b (3H).
// Outputs: hs (T,B,H) hidden states, Z and Zk (T,B,D+H) the two concatenated operands, gates (3,T,B,H) = U, R, O.
int bf_gru_forward(const float* X, const float* W, const float* b, float* hs, float* Z, float* Zk, float* gates, int T,
                   int B, int D, int H, int act, int precision, void* stream) {
  BF_REQUIRE(X && W && b && hs && Z && Zk && gates, BF_ERR_ARG, "bf_gru_forward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_gru_forward: bad shape");
  BF_REQUIRE(act == BF_ACT_TANH || act == BF_ACT_RELU || act == BF_ACT_SIGMOID || act == BF_ACT_LINEAR, BF_ERR_ARG,
             "bf_gru_forward: bad activation %d", act);
  if (T == 0 || B == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H;
  void* ws = nullptr;
  int rc = workspace_require((size_t)ZD * 3 * H * sizeof(float), &ws);
  if (rc) return rc;
  float* Wt = (float*)ws;      // W^T (3H x ZD): rows [0,2H) = Wur^T, rows [2H,3H) = Wo^T
  const bool tma = precision == BF_PREC_TF32 && gemm_tma_ok(Z, Wt, B, H, ZD) && gemm_tma_ok(Zk, Wt + (long)2 * H * ZD, B, H, ZD) &&
                   getenv("BF_LSTM_GATHER") == nullptr;
  if (tma) {
    lstm_transpose_kernel<<<dim3((3 * H + 31) / 32, (ZD + 31) / 32), dim3(32, 8), 0, s>>>(W, Wt, ZD, 3 * H);
    BF_LAUNCH_CHECK();
  }
  lstm_fill_z_kernel<<<ew_grid(TB * D + BH, 256), 256, 0, s>>>(X, Z, TB, B, D, H);
  BF_LAUNCH_CHECK();
  lstm_fill_z_kernel<<<ew_grid(TB * D + BH, 256), 256, 0, s>>>(X, Zk, TB, B, D, H);
  BF_LAUNCH_CHECK();
  float *U = gates, *R = gates + TB * H, *O = gates + 2 * TB * H;
  for (int t = 0; t < T; ++t) {
    const long zo = (long)t * B * ZD, ho = (long)t * BH;
    EpiGruUR e1{U + ho, R + ho, Zk + zo, Z + zo, b, D, H};
    EpiGruO e2{O + ho, hs + ho, t + 1 < T ? Z + zo + (long)B * ZD : nullptr, U + ho, Z + zo, b + 2 * H, D, H, act};
    if (tma) {
      rc = launch_gemm_tma(Z + zo, Wt, B, 2 * H, ZD, e1, s);
      if (rc) return rc;
      rc = launch_gemm_tma(Zk + zo, Wt + (long)2 * H * ZD, B, H, ZD, e2, s);
    } else {
      const int kp = (ZD + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
      LoadKContig a1{Z + zo, (long)ZD};
      LoadMNContig b1{W, (long)3 * H};
      rc = launch_gemm(precision, a1, b1, e1, B, 2 * H, ZD, 1, kp, 0, s);
      if (rc) return rc;
      LoadKContig a2{Zk + zo, (long)ZD};
      LoadMNContig b2{W + 2 * H, (long)3 * H};
      rc = launch_gemm(precision, a2, b2, e2, B, H, ZD, 1, kp, 0, s);
    }
    if (rc) return rc;
  }
  return BF_OK;
}

// GRU layer backward, layers/recurrent.py:155-190.  delta (T,B,H) -> dX (T,B,D), nablaW (D+H,3H), nablab (3H) (the
// reference zeroes and then accumulates them over time, i.e. assigns).
int bf_gru_backward(const float* Z, const float* Zk, const float* gates, const float* delta, const float* W, float* dX,
                    float* nablaW, float* nablab, int T, int B, int D, int H, int act, int precision, void* stream) {
  BF_REQUIRE(Z && Zk && gates && delta && W && dX && nablaW && nablab, BF_ERR_ARG, "bf_gru_backward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_gru_backward: bad shape");
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H, H2 = 2 * H, H3 = 3 * H;
  if (T == 0 || B == 0) {
    BF_CUDA(cudaMemsetAsync(nablaW, 0, (size_t)ZD * H3 * sizeof(float), s));
    BF_CUDA(cudaMemsetAsync(nablab, 0, (size_t)H3 * sizeof(float), s));
    return BF_OK;
  }
  BF_REQUIRE(TB < 2147483647L, BF_ERR_ARG, "bf_gru_backward: T*B too large");
  const int rows = ZD + 1;
  int bm = rows <= 32 ? 32 : (rows <= 64 ? 64 : 128), bn = H2 <= 16 ? 16 : (H2 <= 32 ? 32 : (H2 <= 64 ? 64 : 128));
  int tiles = ((rows + bm - 1) / bm) * ((H2 + bn - 1) / bn);
  int kps = 0;
  int splits = choose_splits(tiles, (int)TB, &kps);      // sized for the wider (2H) product, reused for the H-wide one
  const bool tf = precision == BF_PREC_TF32 && getenv("BF_LSTM_GATHER") == nullptr;
  const bool use_mn = tf && gemm_tma_mn_ok(Z, nullptr, ZD, H, TB) && (H & 3) == 0;
  int mn_splits = num_sms() / ((ZD + 127) / 128);
  if (mn_splits < 1) mn_splits = 1;
  size_t off_dh = 0;
  size_t off_rdk = align_up(off_dh + (size_t)BH * sizeof(float));
  size_t off_dur = align_up(off_rdk + (size_t)BH * sizeof(float));
  size_t off_do = align_up(off_dur + (size_t)TB * H2 * sizeof(float));
  size_t off_pw = align_up(off_do + (size_t)TB * H * sizeof(float));
  const size_t partial_bytes = (size_t)(use_mn ? mn_splits : splits) * rows * H2 * sizeof(float);
  void* ws = nullptr;
  int rc = workspace_require(off_pw + partial_bytes, &ws);
  if (rc) return rc;
  char* base = (char*)ws;
  float* dh = (float*)(base + off_dh);
  float* RdK = (float*)(base + off_rdk);
  float* dUR = (float*)(base + off_dur);
  float* dO = (float*)(base + off_do);
  float* partial = (float*)(base + off_pw);
  const float *U = gates, *R = gates + TB * H, *O = gates + 2 * TB * H;
  BF_CUDA(cudaMemsetAsync(dh, 0, (size_t)BH * sizeof(float), s));
  const bool tma3 = tf && (H & 3) == 0 && gemm_tma_ok(dO, W + H2, B, ZD, H);        // B(n, k) = W[n * 3H + 2H + k]
  const bool tma4 = tf && (H & 3) == 0 && gemm_tma_ok(dUR, W, B, ZD, H2);           // B(n, k) = W[n * 3H + k]
  for (int t = T - 1; t >= 0; --t) {
    const long zo = (long)t * B * ZD, ho = (long)t * BH;
    float* dUR_t = dUR + (long)t * B * H2;
    float* dO_t = dO + ho;
    gru_bwd_pre_kernel<<<ew_grid(BH, 256), 256, 0, s>>>(delta + ho, dh, Z + zo, U + ho, O + ho, dUR_t, dO_t, B, D, H, act);
    BF_LAUNCH_CHECK();
    EpiGruDzk e3{dX + (long)t * B * D, dUR_t, RdK, Z + zo, R + ho, D, H};
    if (tma3) {
      rc = launch_gemm_tma(dO_t, W + H2, B, ZD, H, e3, s, H, H3);
    } else {
      LoadKContig al{dO_t, (long)H};
      LoadKContig bl{W + H2, (long)H3};
      rc = launch_gemm(precision, al, bl, e3, B, ZD, H, 1, (H + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
    EpiGruDz e4{dX + (long)t * B * D, dh, RdK, U + ho, D, H};
    if (tma4) {
      rc = launch_gemm_tma(dUR_t, W, B, ZD, H2, e4, s, H2, H3);
    } else {
      LoadKContig al{dUR_t, (long)H2};
      LoadKContig bl{W, (long)H3};
      rc = launch_gemm(precision, al, bl, e4, B, ZD, H2, 1, (H2 + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
  }
  // nabla_W[:, :2H] = sum_t Z[t]^T [dU, dR],  nabla_W[:, 2H:] = sum_t Zk[t]^T dO;  biases = column sums (recurrent.py:178-186)
  for (int part = 0; part < 2; ++part) {
    const float* A = part ? Zk : Z;
    const float* Bm = part ? dO : dUR;
    const int N = part ? H : H2;
    float* outW = nablaW + (part ? H2 : 0);
    float* outb = nablab + (part ? H2 : 0);
    int used = splits;
    if (use_mn) {
      rc = launch_gemm_tma_mn(A, Bm, ZD, N, (int)TB, partial, partial_bytes, &used, s);
    } else {
      LoadMNContigOnes al{A, (long)ZD, ZD};
      LoadMNContig bl{Bm, (long)N};
      EpiPartial ep{partial, (long)rows * N, N};
      rc = launch_gemm(precision, al, bl, ep, rows, N, (int)TB, splits, kps, 0, s);
    }
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, outW, outb, rows, N, ZD, used, 0, H3, 1, s);
    if (rc) return rc;
  }
  return BF_OK;
}

// ClockworkLayer.feedforward, layers/recurrent.py:239-257.  X (T,B,D) time-major, W (D+H,H), b (H), tick (H) the layer's
// tick_array -> O (T,B,H) (self.cache), Z (T,B,D+H) (self.Zs).
int bf_clockwork_forward(const float* X, const float* W, const float* b, const float* tick, float* O, float* Z, int T,
                         int B, int D, int H, int act, int precision, void* stream) {
  BF_REQUIRE(X && W && b && tick && O && Z, BF_ERR_ARG, "bf_clockwork_forward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_clockwork_forward: bad shape");
  BF_REQUIRE(act == BF_ACT_TANH || act == BF_ACT_RELU || act == BF_ACT_SIGMOID || act == BF_ACT_LINEAR, BF_ERR_ARG,
             "bf_clockwork_forward: bad activation %d", act);
  if (T == 0 || B == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H;
  void* ws = nullptr;
  int rc = workspace_require((size_t)ZD * H * sizeof(float), &ws);
  if (rc) return rc;
  float* Wt = (float*)ws;
  const bool tma = precision == BF_PREC_TF32 && gemm_tma_ok(Z, Wt, B, H, ZD) && getenv("BF_LSTM_GATHER") == nullptr;
  if (tma) {
    lstm_transpose_kernel<<<dim3((H + 31) / 32, (ZD + 31) / 32), dim3(32, 8), 0, s>>>(W, Wt, ZD, H);
    BF_LAUNCH_CHECK();
  }
  lstm_fill_z_kernel<<<ew_grid(TB * D + BH, 256), 256, 0, s>>>(X, Z, TB, B, D, H);
  BF_LAUNCH_CHECK();
  for (int t = 0; t < T; ++t) {
    EpiCwFwd ep{O + (long)t * BH, t + 1 < T ? Z + (long)(t + 1) * B * ZD : nullptr, b, tick, D, H, act, t + 1};
    if (tma) {
      rc = launch_gemm_tma(Z + (long)t * B * ZD, Wt, B, H, ZD, ep, s);
    } else {
      LoadKContig al{Z + (long)t * B * ZD, (long)ZD};
      LoadMNContig bl{W, (long)H};
      rc = launch_gemm(precision, al, bl, ep, B, H, ZD, 1, (ZD + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
  }
  return BF_OK;
}

// ClockworkLayer.backpropagate, layers/recurrent.py:259-277.  delta (T,B,H) -> dX (T,B,D), nablaW (D+H,H), nablab (H).
int bf_clockwork_backward(const float* Z, const float* O, const float* delta, const float* W, const float* tick, float* dX,
                          float* nablaW, float* nablab, int T, int B, int D, int H, int act, int precision, void* stream) {
  BF_REQUIRE(Z && O && delta && W && tick && dX && nablaW && nablab, BF_ERR_ARG, "bf_clockwork_backward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_clockwork_backward: bad shape");
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H;
  const int ZD = D + H;
  if (T == 0 || B == 0) {
    BF_CUDA(cudaMemsetAsync(nablaW, 0, (size_t)ZD * H * sizeof(float), s));
    BF_CUDA(cudaMemsetAsync(nablab, 0, (size_t)H * sizeof(float), s));
    return BF_OK;
  }
  BF_REQUIRE(TB < 2147483647L, BF_ERR_ARG, "bf_clockwork_backward: T*B too large");
  const int rows = ZD + 1;
  int bm = rows <= 32 ? 32 : (rows <= 64 ? 64 : 128), bn = H <= 16 ? 16 : (H <= 32 ? 32 : (H <= 64 ? 64 : 128));
  int tiles = ((rows + bm - 1) / bm) * ((H + bn - 1) / bn);
  int kps = 0;
  int splits = choose_splits(tiles, (int)TB, &kps);
  const bool tf = precision == BF_PREC_TF32 && getenv("BF_LSTM_GATHER") == nullptr;
  const bool use_mn = tf && gemm_tma_mn_ok(Z, nullptr, ZD, H, TB);
  int mn_splits = num_sms() / (((ZD + 127) / 128) * ((H + 127) / 128));
  if (mn_splits < 1) mn_splits = 1;
  const size_t g_bytes = align_up((size_t)TB * H * sizeof(float));
  const size_t partial_bytes = (size_t)(use_mn ? mn_splits : splits) * rows * H * sizeof(float);
  void* ws = nullptr;
  int rc = workspace_require(g_bytes + partial_bytes, &ws);
  if (rc) return rc;
  float* G = (float*)ws;
  float* partial = (float*)((char*)ws + g_bytes);
  cw_first_kernel<<<ew_grid(BH, 256), 256, 0, s>>>(delta + (T - 1) * BH, O + (T - 1) * BH, tick, G + (T - 1) * BH, BH, H, act, T);
  BF_LAUNCH_CHECK();
  for (int t = T - 1; t >= 0; --t) {
    const float* Gt = G + (long)t * BH;
    EpiCwBwd ep{dX + (long)t * B * D, t ? delta + (long)(t - 1) * BH : nullptr, t ? O + (long)(t - 1) * BH : nullptr, tick,
                t ? G + (long)(t - 1) * BH : nullptr, D, H, act, t};
    if (tf && gemm_tma_ok(Gt, W, B, ZD, H)) {
      rc = launch_gemm_tma(Gt, W, B, ZD, H, ep, s);          // B(n, k) = W[n * H + k]: K-major as stored
    } else {
      LoadKContig al{Gt, (long)H};
      LoadKContig bl{W, (long)H};
      rc = launch_gemm(precision, al, bl, ep, B, ZD, H, 1, (H + GEMM_BK - 1) / GEMM_BK * GEMM_BK, 0, s);
    }
    if (rc) return rc;
  }
  if (use_mn) {
    int used = 1;
    rc = launch_gemm_tma_mn(Z, G, ZD, H, (int)TB, partial, partial_bytes, &used, s);
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, nablaW, nablab, rows, H, ZD, used, 0, H, 1, s);
  } else {
    LoadMNContigOnes al{Z, (long)ZD, ZD};
    LoadMNContig bl{G, (long)H};
    EpiPartial ep{partial, (long)rows * H, H};
    rc = launch_gemm(precision, al, bl, ep, rows, H, (int)TB, splits, kps, 0, s);
    if (rc) return rc;
    rc = launch_splitk_reduce(partial, nablaW, nablab, rows, H, ZD, splits, 0, H, 1, s);
  }
  return rc;
}

}  // extern "C"
