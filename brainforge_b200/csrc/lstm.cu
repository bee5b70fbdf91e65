// LSTM forward / BPTT: atomic/recurrent_op.py:46-112 (LSTMOp), llatomic/_lllstm.py:6-67.
// Time is a true recurrence and stays a host loop over T; each step is one gate GEMM
// (B x (D+H)) @ ((D+H) x 4H) plus one fused elementwise gate kernel.  The weight gradient is ONE
// split-K contraction over all T*B rows after the loop (recurrent_op.py:109), with the bias gradient
// riding along as an all-ones row.
#include "gemm_tma.cuh"
#include <stdlib.h>

namespace bf {

// Z[t,b,0:D] = X[t,b,:] for all t;  Z[0,b,D:] = 0  (O[-1] is the still-zero last row, recurrent_op.py:60)
__global__ void lstm_fill_z_kernel(const float* __restrict__ X, float* __restrict__ Z, long TB, int B, int D, int H) {
  long total = TB * D + (long)B * H;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    if (i < TB * D) {
      long r = i / D;
      int c = (int)(i - r * D);
      Z[r * (D + H) + c] = X[i];
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
      bw[idx] = act_deriv_from_z(act, c, ca);
      bw[bw_stride + idx] = act_deriv_from_z(act, p[j], cv);
      bw[2 * bw_stride + idx] = fv * sigmoidf_(-p[H + j]);
      bw[3 * bw_stride + idx] = iv * sigmoidf_(-p[2 * H + j]);
      bw[4 * bw_stride + idx] = ov * sigmoidf_(-p[3 * H + j]);
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
    float cav = Ca[idx], cv = cand[idx], fv = f[idx], iv = ig[idx], ov = o[idx];
    float bCa, bcand, bf_, bi, bo;
    if (bw) {
      bCa = bw[idx]; bcand = bw[bw_stride + idx]; bf_ = bw[2 * bw_stride + idx];
      bi = bw[3 * bw_stride + idx]; bo = bw[4 * bw_stride + idx];
    } else {  // derivative-from-output exactly as recurrent_op.py:85-88 (FP32 cancellation when saturated)
      bCa = act_bwd_rt(act, cav); bcand = act_bwd_rt(act, cv);
      bf_ = fv * (1.f - fv); bi = iv * (1.f - iv); bo = ov * (1.f - ov);
    }
    float dC = (first ? 0.f : deltaC[idx]) + e * ov * bCa;
    float cp = Cprev ? Cprev[idx] : 0.f;
    float* dg = dgates + b * 4 * H;
    dg[j] = dC * iv * bcand;
    dg[H + j] = dC * cp * bf_;
    dg[2 * H + j] = dC * cv * bi;
    dg[3 * H + j] = cav * e * bo;
    deltaC[idx] = dC * fv;
  }
}

// deltaZ[t] = dgates[t] @ W^T, split on the fly: columns < D are dX[t], the rest feed E[t-1]
struct EpiLstmDz {
  float* dX; float* dH; int D, H;
  __device__ void store(int m, int n, float v, int) const {
    if (n < D) dX[(long)m * D + n] = v;
    else dH[(long)m * H + (n - D)] = v;
  }
};
struct EpiBias {
  float* out; const float* b; int N;
  __device__ void store(int m, int n, float v, int) const { out[(long)m * N + n] = v + __ldg(b + n); }
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

}  // namespace bf

using namespace bf;

extern "C" {

int bf_lstm_forward(const float* X, const float* W, const float* b, float* O, float* Z, float* cache, float* bw,
                    int T, int B, int D, int H, int act, int precision, void* stream) {
  BF_REQUIRE(X && W && b && O && Z && cache, BF_ERR_ARG, "bf_lstm_forward: null pointer");
  BF_REQUIRE(T >= 0 && B >= 0 && D > 0 && H > 0, BF_ERR_ARG, "bf_lstm_forward: bad shape");
  BF_REQUIRE(act == BF_ACT_TANH || act == BF_ACT_RELU || act == BF_ACT_SIGMOID || act == BF_ACT_LINEAR, BF_ERR_ARG,
             "bf_lstm_forward: bad activation %d", act);
  if (T == 0 || B == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const long TB = (long)T * B, BH = (long)B * H, ZD = D + H;
  void* ws = nullptr;
  const size_t gates_bytes = align_up((size_t)B * 4 * H * sizeof(float));
  int rc = workspace_require(gates_bytes + (size_t)ZD * 4 * H * sizeof(float), &ws);
  if (rc) return rc;
  float* gates = (float*)ws;
  float* Wt = (float*)((char*)ws + gates_bytes);     // W^T (4H x (D+H)): the K-major B operand of the TMA-fed gate GEMM
  const bool tma = precision == BF_PREC_TF32 && gemm_tma_ok(Z, Wt, B, 4 * H, (int)ZD) && ((ZD * 4) % 16) == 0 &&
                   getenv("BF_LSTM_GATHER") == nullptr;
  if (tma) {
    lstm_transpose_kernel<<<dim3((4 * H + 31) / 32, ((int)ZD + 31) / 32), dim3(32, 8), 0, s>>>(W, Wt, (int)ZD, 4 * H);
    BF_LAUNCH_CHECK();
  }
  lstm_fill_z_kernel<<<ew_grid(TB * D + BH, 256), 256, 0, s>>>(X, Z, TB, B, D, H);
  BF_LAUNCH_CHECK();
  float* C = cache;               // cache (6,T,B,H): C, Ca, cand, f, i, o  (recurrent_op.py:77)
  float* Ca = cache + 1 * TB * H;
  float* cand = cache + 2 * TB * H;
  float* f = cache + 3 * TB * H;
  float* ig = cache + 4 * TB * H;
  float* o = cache + 5 * TB * H;
  const int K = D + H, N = 4 * H;
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

int bf_lstm_backward(const float* Z, const float* O, const float* E, const float* W, const float* cache,
                     const float* bw, float* dX, float* nablaW, float* nablab, int T, int B, int D, int H, int act,
                     int precision, void* stream) {
  (void)O;
  BF_REQUIRE(Z && E && W && cache && dX && nablaW && nablab, BF_ERR_ARG, "bf_lstm_backward: null pointer");
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
  for (int t = T - 1; t >= 0; --t) {
    long off = (long)t * BH;
    float* dg_t = dgates + (long)t * B * N4;
    lstm_gates_bwd_kernel<<<ew_grid(BH, 256), 256, 0, s>>>(E + off, t == T - 1 ? nullptr : dH,
                                                          t ? C + off - BH : nullptr, Ca + off, cand + off, f + off,
                                                          ig + off, o + off, bw ? bw + off : nullptr, TB * H, deltaC, dg_t, B, H,
                                                          act, t == T - 1);
    BF_LAUNCH_CHECK();
    // deltaZ[t] = dgates[t] @ W^T : M=B, N=D+H, K=4H ; B(k, n) = W[n*4H + k]
    EpiLstmDz ep{dX + (long)t * B * D, dH, D, H};
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

}  // extern "C"
