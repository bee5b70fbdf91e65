// Convolution as implicit GEMM on NCHW FP32 tensors: atomic/tensor_op.py:6-75 (ConvolutionOp).
// The im2col matrix of the reference (tensor_op.py:15-24, llatomic/lltensor_op.py:9-22) is never
// materialised: the GEMM loaders below compute patch addresses on the fly from small index tables
// held in shared memory.
//   forward  : Y[(img,y,x), f]   = sum_{c,ky,kx} X[img,c,y+ky-py,x+kx-px] * F[f,c,ky,kx]
//   dX       : dX[(img,y,x), c]  = sum_{f,ky,kx} E[img,f,y+ky-(fy-1),x+kx-(fx-1)] * F[f,c,fy-1-ky,fx-1-kx]
//   dF (+db) : dF[(c,ky,kx), f]  = sum_{img,y,x} X[img,c,y+ky,x+kx] * E[img,f,y,x]   (+ ones row -> db[f])
#include "gemm_umma.cuh"

namespace bf {

// A(m=(img,y,x), k=(c,ky,kx)) = X[img, c, y+ky-py, x+kx-px]  (zero outside the image when PAD)
template <bool PAD>
struct ConvPatchA {
  static constexpr bool kContigMN = true;
  static constexpr bool kHasTables = true;
  const float* X;
  int ic, iy, ix;      // input tensor dims
  int oy, ox;          // output spatial dims
  int fy, fx, py, px;  // filter dims and zero padding
  int K;               // ic*fy*fx
  int tab;             // offset (ints) of this loader's tables in dynamic smem
  mutable const int* koff;
  mutable const int* kyx;
  struct Row { long base; int y0, x0; };
  typedef NoTile Tile;
  __device__ void prepare(int* tables) const {
    int* t0 = tables + tab;
    int* t1 = t0 + K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      int c = k / (fy * fx), r = k % (fy * fx), ky = r / fx, kx = r % fx;
      t0[k] = c * iy * ix + ky * ix + kx;
      t1[k] = (ky << 16) | kx;
    }
    koff = t0;
    kyx = t1;
  }
  __device__ Row row(int m) const {
    int P = oy * ox;
    int img = m / P, p = m % P, y = p / ox, x = p % ox;
    Row r;
    r.y0 = y - py;
    r.x0 = x - px;
    r.base = (long)img * ic * iy * ix + (long)r.y0 * ix + r.x0;
    return r;
  }
  __device__ Tile tile(int) const { return NoTile(); }
  __device__ float load(const Row& r, Tile, int, int k) const {
    if (PAD) {
      int v = kyx[k];
      int yy = r.y0 + (v >> 16), xx = r.x0 + (v & 0xffff);
      if ((unsigned)yy >= (unsigned)iy || (unsigned)xx >= (unsigned)ix) return 0.f;
    }
    return __ldg(X + r.base + koff[k]);
  }
};

// B(k=(f,ky,kx), n=c) = F[f, c, fy-1-ky, fx-1-kx]   (flipped + channel-transposed filter, never built)
struct ConvFlipB {
  static constexpr bool kContigMN = true;
  static constexpr bool kHasTables = true;
  const float* F;
  int nf, ic, fy, fx, K;  // K = nf*fy*fx
  int tab;
  mutable const int* foff;
  typedef int Row; typedef NoTile Tile;
  __device__ void prepare(int* tables) const {
    int* t0 = tables + tab;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      int f = k / (fy * fx), r = k % (fy * fx), ky = r / fx, kx = r % fx;
      t0[k] = f * ic * fy * fx + (fy - 1 - ky) * fx + (fx - 1 - kx);
    }
    foff = t0;
  }
  __device__ Row row(int n) const { return n * fy * fx; }
  __device__ Tile tile(int) const { return NoTile(); }
  __device__ float load(Row r, Tile, int, int k) const { return __ldg(F + foff[k] + r); }
};

// Reduction index for dF: r = img*Ppad + p, Ppad = roundup(oy*ox, 16) so that K tiles never straddle images.
struct ImgTile { long base; int p0; };

// A'(m=(c,ky,kx) | ones row, r=(img,p)) = X[img, c, y+ky, x+kx]
struct ConvPatchT {
  static constexpr bool kContigMN = false;
  static constexpr bool kHasTables = true;
  const float* X;
  int ic, iy, ix, oy, ox, fy, fx;
  int K, P, Ppad;
  int tab;
  mutable const int* koff;
  mutable const int* poff;
  typedef int Row; typedef ImgTile Tile;
  __device__ void prepare(int* tables) const {
    int* t0 = tables + tab;
    int* t1 = t0 + K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      int c = k / (fy * fx), r = k % (fy * fx), ky = r / fx, kx = r % fx;
      t0[k] = c * iy * ix + ky * ix + kx;
    }
    for (int p = threadIdx.x; p < Ppad; p += blockDim.x) t1[p] = p < P ? (p / ox) * ix + (p % ox) : -1;
    koff = t0;
    poff = t1;
  }
  __device__ Row row(int m) const { return m < K ? koff[m] : -1; }  // -1: the virtual all-ones row
  __device__ Tile tile(int k0) const {
    Tile t;
    int img = k0 / Ppad;
    t.p0 = k0 - img * Ppad;
    t.base = (long)img * ic * iy * ix;
    return t;
  }
  __device__ float load(Row r, const Tile& t, int kk, int) const {
    int po = poff[t.p0 + kk];
    if (po < 0) return 0.f;
    if (r < 0) return 1.0f;
    return __ldg(X + t.base + r + po);
  }
};

// B'(r=(img,p), n=f) = E[img, f, p]
struct ConvErrB {
  static constexpr bool kContigMN = false;
  static constexpr bool kHasTables = false;
  const float* E;
  int nf, P, Ppad;
  typedef long Row; typedef ImgTile Tile;
  __device__ void prepare(int*) const {}
  __device__ Row row(int n) const { return (long)n * P; }
  __device__ Tile tile(int k0) const {
    Tile t;
    int img = k0 / Ppad;
    t.p0 = k0 - img * Ppad;
    t.base = (long)img * nf * P;
    return t;
  }
  __device__ float load(Row r, const Tile& t, int kk, int) const {
    int p = t.p0 + kk;
    return p < P ? __ldg(E + t.base + r + p) : 0.f;
  }
};

// Y[img, n, p] = act(acc) + bias[n]     (bias AFTER the activation: layers/tensor.py:77-78)
struct EpiConv {
  float* Y; const float* bias; int N, P, act;
  __device__ void store(int m, int n, float v, int) const {
    int img = m / P, p = m - img * P;
    v = act_fwd_rt(act, v);
    if (bias) v += __ldg(bias + n);
    Y[((long)img * N + n) * P + p] = v;
  }
};

}  // namespace bf

using namespace bf;

extern "C" {

int bf_conv_forward(const float* A, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix,
                    int nf, int fc, int fy, int fx, int mode, int act, int precision, void* stream) {
  BF_REQUIRE(F && (im == 0 || (A && Y)), BF_ERR_ARG, "bf_conv_forward: null pointer");
  BF_REQUIRE(im >= 0 && ic > 0 && iy > 0 && ix > 0 && nf > 0 && fy > 0 && fx > 0, BF_ERR_ARG,
             "bf_conv_forward: bad shape");
  if (fc != ic)  // atomic/tensor_op.py:18-21
    return fail(BF_ERR_VALUE,
                "Supplied filter (F) is incompatible with supplied input! (X)\ninput depth: %d != %d :filter depth",
                ic, fc);
  BF_REQUIRE(mode == BF_CONV_VALID || mode == BF_CONV_FULL, BF_ERR_ARG, "bf_conv_forward: bad mode %d", mode);
  BF_REQUIRE(act >= BF_ACT_LINEAR && act <= BF_ACT_RELU, BF_ERR_ARG, "bf_conv_forward: bad activation %d", act);
  BF_REQUIRE(fy < 65536 && fx < 65536, BF_ERR_ARG, "bf_conv_forward: filter too large");
  int py = mode == BF_CONV_FULL ? fy - 1 : 0, px = mode == BF_CONV_FULL ? fx - 1 : 0;
  int oy = iy + 2 * py - fy + 1, ox = ix + 2 * px - fx + 1;
  if (mode == BF_CONV_VALID)
    BF_REQUIRE(oy > 0 && ox > 0, BF_ERR_VALUE, "bf_conv_forward: filter (%d,%d) larger than input (%d,%d)", fy, fx, iy, ix);
  if (im == 0) return BF_OK;
  long M = (long)im * oy * ox;
  BF_REQUIRE(M < 2147483647L, BF_ERR_ARG, "bf_conv_forward: too many output pixels");
  int K = ic * fy * fx;
  cudaStream_t s = as_stream(stream);
  LoadKContig bl{F, (long)K};
  EpiConv ep{Y, bias, nf, oy * ox, act};
  int kps = (K + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
  if (mode == BF_CONV_VALID) {
    ConvPatchA<false> al{A, ic, iy, ix, oy, ox, fy, fx, 0, 0, K, 0, nullptr, nullptr};
    return launch_gemm(precision, al, bl, ep, (int)M, nf, K, 1, kps, (size_t)2 * K * sizeof(int), s);
  }
  ConvPatchA<true> al{A, ic, iy, ix, oy, ox, fy, fx, py, px, K, 0, nullptr, nullptr};
  return launch_gemm(precision, al, bl, ep, (int)M, nf, K, 1, kps, (size_t)2 * K * sizeof(int), s);
}

int bf_conv_backward(const float* X, const float* E, const float* F, float* dF, float* db, float* dX, int im, int ic,
                     int iy, int ix, int nf, int fy, int fx, int precision, void* stream) {
  BF_REQUIRE(F && (im == 0 || (X && E)), BF_ERR_ARG, "bf_conv_backward: null pointer");
  BF_REQUIRE(im >= 0 && ic > 0 && iy > 0 && ix > 0 && nf > 0 && fy > 0 && fx > 0, BF_ERR_ARG,
             "bf_conv_backward: bad shape");
  int oy = iy - fy + 1, ox = ix - fx + 1;
  BF_REQUIRE(oy > 0 && ox > 0, BF_ERR_VALUE, "bf_conv_backward: filter larger than input");
  cudaStream_t s = as_stream(stream);
  int K = ic * fy * fx, P = oy * ox, Ppad = (P + GEMM_KALIGN - 1) / GEMM_KALIGN * GEMM_KALIGN;
  int rc;
  // ---- dF and db in one split-K contraction over (img, pixel)
  if (dF) {
    int rows = K + 1;
    long Kred = (long)im * Ppad;
    BF_REQUIRE(Kred < 2147483647L, BF_ERR_ARG, "bf_conv_backward: reduction too long");
    int bm = rows <= 32 ? 32 : (rows <= 64 ? 64 : 128), bn = nf <= 16 ? 16 : (nf <= 32 ? 32 : (nf <= 64 ? 64 : 128));
    int tiles = ((rows + bm - 1) / bm) * ((nf + bn - 1) / bn);
    int kps = GEMM_BK, splits = 1;
    if (im > 0) splits = choose_splits(tiles, (int)Kred, &kps);
    void* ws = nullptr;
    rc = workspace_require((size_t)splits * rows * nf * sizeof(float), &ws);
    if (rc) return rc;
    if (im == 0) {
      BF_CUDA(cudaMemsetAsync(ws, 0, (size_t)rows * nf * sizeof(float), s));
    } else {
      ConvPatchT al{X, ic, iy, ix, oy, ox, fy, fx, K, P, Ppad, 0, nullptr, nullptr};
      ConvErrB bl{E, nf, P, Ppad};
      EpiPartial ep{(float*)ws, (long)rows * nf, nf};
      rc = launch_gemm(precision, al, bl, ep, rows, nf, (int)Kred, splits, kps, (size_t)(K + Ppad) * sizeof(int), s);
      if (rc) return rc;
    }
    // partial is (kidx | ones, f); dF is (f, kidx): write transposed, route the ones row to db
    rc = launch_splitk_reduce((const float*)ws, dF, db, rows, nf, K, splits, /*accumulate=*/0, 1, K, s);
    if (rc) return rc;
  }
  // ---- dX = full(E, flip(F)^T)
  if (dX && im > 0) {
    long M = (long)im * iy * ix;
    BF_REQUIRE(M < 2147483647L, BF_ERR_ARG, "bf_conv_backward: too many input pixels");
    int K2 = nf * fy * fx;
    ConvPatchA<true> al{E, nf, oy, ox, iy, ix, fy, fx, fy - 1, fx - 1, K2, 0, nullptr, nullptr};
    ConvFlipB bl{F, nf, ic, fy, fx, K2, 2 * K2, nullptr};
    EpiConv ep{dX, nullptr, ic, iy * ix, BF_ACT_LINEAR};
    int kps = (K2 + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
    rc = launch_gemm(precision, al, bl, ep, (int)M, ic, K2, 1, kps, (size_t)3 * K2 * sizeof(int), s);
    if (rc) return rc;
  }
  return BF_OK;
}

}  // extern "C"
