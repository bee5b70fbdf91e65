// Dense layer contractions: atomic/core_op.py:9-20 (DenseOp.forward / backward).
#include "gemm_umma.cuh"

namespace bf {

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, float* __restrict__ aux,
                                     int rows, int N, int aux_row, int splits, int accumulate, long out_rs,
                                     long out_cs) {
  long total = (long)rows * N;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(long)z * total + i];
    int r = (int)(i / N), c = (int)(i % N);
    float* dst;
    if (r == aux_row) dst = aux ? aux + c : nullptr;
    else dst = out + (long)(r > aux_row && aux_row >= 0 ? r - 1 : r) * out_rs + (long)c * out_cs;
    if (dst) *dst = accumulate ? *dst + s : s;
  }
}

int launch_splitk_reduce(const float* ws, float* out, float* aux, int rows, int N, int aux_row, int splits,
                         int accumulate, long out_rs, long out_cs, cudaStream_t s) {
  long total = (long)rows * N;
  if (total == 0) return BF_OK;
  splitk_reduce_kernel<<<ew_grid(total, 256), 256, 0, s>>>(ws, out, aux, rows, N, aux_row, splits, accumulate,
                                                             out_rs, out_cs);
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

}  // namespace bf

using namespace bf;

extern "C" {

int bf_dense_forward(const float* X, const float* W, const float* b, float* out, int m, int k, int n, int act,
                     int precision, void* stream) {
  BF_REQUIRE(m >= 0 && k > 0 && n > 0, BF_ERR_ARG, "bf_dense_forward: bad shape m=%d k=%d n=%d", m, k, n);
  BF_REQUIRE(act >= BF_ACT_LINEAR && act <= BF_ACT_SOFTMAX, BF_ERR_ARG, "bf_dense_forward: bad activation %d", act);
  if (m == 0) return BF_OK;
  BF_REQUIRE(X && W && out, BF_ERR_ARG, "bf_dense_forward: null pointer");
  cudaStream_t s = as_stream(stream);
  int ep_act = act == BF_ACT_SOFTMAX ? BF_ACT_LINEAR : act;
  int rc = BF_OK;
  {
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
    if (m == 0) {
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
    LoadKContig al{E, (long)n};
    LoadKContig bl{W, (long)n};
    EpiPlain ep{dX, k};
    rc = launch_gemm(precision, al, bl, ep, m, k, n, 1, (n + GEMM_KALIGN - 1) / GEMM_KALIGN * GEMM_KALIGN, 0, s);
    if (rc) return rc;
  }
  return BF_OK;
}

}  // extern "C"
