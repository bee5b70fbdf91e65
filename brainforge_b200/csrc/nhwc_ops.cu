// Memory-bound companions of the channels-last convolution path (conv_nhwc.cu, conv_c3.cu):
//   * layout changes NCHW <-> NHWC (a batched 2-D transpose through a padded shared-memory tile),
//   * max pooling on channels-last activations with the reference's multi-hot equality mask
//     (atomic/tensor_op.py:93-119), optionally fused with the ReLU that precedes it in the stack
//     (layers/core.py:45-59): forward reads the raw convolution output once and applies max(0,.) on the fly,
//     backward multiplies by (pooled output > 0) while scattering, so the ReLU layer costs no HBM pass.
// One thread owns four consecutive channels of one pooling window: 128-bit loads and stores, a warp covers
// 512 contiguous bytes of every input pixel row it touches.
#include "common.cuh"

namespace bf {

// src (B, R, C) -> dst (B, C, R)
__global__ void batched_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int R, int C) {
  __shared__ float tile[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const float* s = src + (size_t)b * R * C;
    float* d = dst + (size_t)b * R * C;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int r = r0 + ty + j, c = c0 + tx;
      if (r < R && c < C) tile[ty + j][tx] = s[(size_t)r * C + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int c = c0 + ty + j, r = r0 + tx;
      if (r < R && c < C) d[(size_t)c * R + r] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

// Small planes (R or C below a tile): one thread per destination element, reads strided but L1/L2 resident.
__global__ void batched_transpose_small_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t total, int R, int C) {
  pdl_launch_dependents();
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const size_t plane = (size_t)R * C;
  for (; i < total; i += stride) {
    const size_t b = i / plane;
    const int rem = (int)(i - b * plane);
    const int c = rem / R, r = rem - c * R;      // destination (b, c, r)
    dst[i] = src[b * plane + (size_t)r * C + c];
  }
}

static int batched_transpose(const float* src, float* dst, int B, int R, int C, cudaStream_t s) {
  if (B <= 0 || R <= 0 || C <= 0) return BF_OK;
  if (R == 1 || C == 1) {
    BF_CUDA(cudaMemcpyAsync(dst, src, (size_t)B * R * C * 4, cudaMemcpyDeviceToDevice, s));
    return BF_OK;
  }
  if (R < 16 || C < 16) {
    const size_t total = (size_t)B * R * C;
    launch_pdl(batched_transpose_small_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, s, src, dst, total, R, C);
  } else {
    dim3 grid((C + 31) / 32, (R + 31) / 32, B < 16384 ? B : 16384), block(32, 8);
    launch_pdl(batched_transpose_kernel, dim3(grid), dim3(block), 0, s, src, dst, B, R, C);
  }
  BF_LAUNCH_CHECK();
  return BF_OK;
}

__device__ __forceinline__ float4 relu4(float4 v) { return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)); }

// byte-mask windows (fdim*fdim <= 8, i.e. fdim 1 or 2), FD = compile-time window edge
template <int FD, bool RELU>
__global__ void pool_nhwc_fwd_kernel(const float4* __restrict__ A, float4* __restrict__ out, uchar4* __restrict__ mask,
                                     long total, int oy, int ox, int c4, int ix) {
  pdl_launch_dependents();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int q = (int)(i % c4);
    long w = i / c4;
    const int x = (int)(w % ox);
    w /= ox;
    const int y = (int)(w % oy);
    const long b = w / oy;
    const float4* src = A + ((b * oy * FD + (long)y * FD) * ix + (long)x * FD) * c4 + q;
    float4 v[FD * FD];
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < FD; ++dy)
#pragma unroll
      for (int dx = 0; dx < FD; ++dx) {
        float4 t = src[((long)dy * ix + dx) * c4];
        if (RELU) t = relu4(t);
        v[dy * FD + dx] = t;
        mx.x = fmaxf(mx.x, t.x); mx.y = fmaxf(mx.y, t.y); mx.z = fmaxf(mx.z, t.z); mx.w = fmaxf(mx.w, t.w);
      }
    uchar4 mk = make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < FD * FD; ++k) {
      const float4 t = v[k];
      mk.x |= (unsigned char)((t.x == mx.x) << k);
      mk.y |= (unsigned char)((t.y == mx.y) << k);
      mk.z |= (unsigned char)((t.z == mx.z) << k);
      mk.w |= (unsigned char)((t.w == mx.w) << k);
    }
    out[i] = mx;
    mask[i] = mk;
  }
}

// DBP: also emit, per CTA, the per-channel sums of everything written to dX (row blockIdx.x of `dbp`, c floats): the
// bias gradient db = E.sum((0,2,3)) of the convolution below (atomic/tensor_op.py:55) as a by-product of the pass that
// produces its E, instead of an extra all-ones MMA per K step in that layer's filter-gradient kernel.  Needs 256 % c4 == 0
// (a thread then keeps its channel quad over the whole grid-stride loop); fixed-order block tree -> deterministic.
template <int FD, bool GATE, bool DBP>
__global__ void __launch_bounds__(256)
pool_nhwc_bwd_kernel(const float4* __restrict__ E, const uchar4* __restrict__ mask, const float4* __restrict__ gate,
                                     float4* __restrict__ dX, long total, int oy, int ox, int c4, int ix, float4* __restrict__ dbp) {
  __shared__ float4 sred[DBP ? 256 : 1];
  pdl_launch_dependents();
  pdl_wait();
  constexpr int fd = FD;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int q = (int)(i % c4);
    long w = i / c4;
    const int x = (int)(w % ox);
    w /= ox;
    const int y = (int)(w % oy);
    const long b = w / oy;
    float4 e = E[i];
    const uchar4 mk = mask[i];
    if (GATE) {   // delta * relu'(pooled output): every masked element of the window equals the pooled output
      const float4 g = gate[i];
      e.x = g.x > 0.f ? e.x : 0.f; e.y = g.y > 0.f ? e.y : 0.f; e.z = g.z > 0.f ? e.z : 0.f; e.w = g.w > 0.f ? e.w : 0.f;
    }
    if (DBP) {     // a tied window writes its value to every tied element
      acc.x += e.x * (float)__popc(mk.x); acc.y += e.y * (float)__popc(mk.y);
      acc.z += e.z * (float)__popc(mk.z); acc.w += e.w * (float)__popc(mk.w);
    }
    float4* dst = dX + ((b * oy * fd + (long)y * fd) * ix + (long)x * fd) * c4 + q;
#pragma unroll
    for (int dy = 0; dy < fd; ++dy)
#pragma unroll
      for (int dx = 0; dx < fd; ++dx) {
        const int k = dy * fd + dx;
        dst[((long)dy * ix + dx) * c4] = make_float4((mk.x >> k) & 1 ? e.x : 0.f, (mk.y >> k) & 1 ? e.y : 0.f,
                                                     (mk.z >> k) & 1 ? e.z : 0.f, (mk.w >> k) & 1 ? e.w : 0.f);
      }
  }
  if (DBP) {
    sred[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w >= c4; w >>= 1) {          // threads t and t + w hold the same channel quad (c4 divides w)
      if ((int)threadIdx.x < w) {
        const float4 o = sred[threadIdx.x + w];
        float4 m = sred[threadIdx.x];
        m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
        sred[threadIdx.x] = m;
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < c4) dbp[(size_t)blockIdx.x * c4 + threadIdx.x] = sred[threadIdx.x];
  }
}

// 64-bit mask words (3 <= fdim <= 8), one channel per thread
template <bool RELU>
__global__ void pool_nhwc_fwd_wide_kernel(const float* __restrict__ A, float* __restrict__ out, unsigned long long* __restrict__ mask,
                                          long total, int oy, int ox, int c, int ix, int fd) {
  pdl_launch_dependents();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int q = (int)(i % c);
    long w = i / c;
    const int x = (int)(w % ox);
    w /= ox;
    const int y = (int)(w % oy);
    const long b = w / oy;
    const float* src = A + ((b * oy * fd + (long)y * fd) * ix + (long)x * fd) * c + q;
    float mx = -INFINITY;
    for (int dy = 0; dy < fd; ++dy)
      for (int dx = 0; dx < fd; ++dx) {
        float t = src[((long)dy * ix + dx) * c];
        if (RELU) t = fmaxf(t, 0.f);
        mx = fmaxf(mx, t);
      }
    unsigned long long mk = 0ull;
    for (int dy = 0; dy < fd; ++dy)
      for (int dx = 0; dx < fd; ++dx) {
        float t = src[((long)dy * ix + dx) * c];
        if (RELU) t = fmaxf(t, 0.f);
        if (t == mx) mk |= 1ull << (dy * fd + dx);
      }
    out[i] = mx;
    mask[i] = mk;
  }
}

__global__ void pool_nhwc_bwd_wide_kernel(const float* __restrict__ E, const unsigned long long* __restrict__ mask,
                                          const float* __restrict__ gate, float* __restrict__ dX, long total, int oy, int ox,
                                          int c, int ix, int fd) {
  pdl_launch_dependents();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int q = (int)(i % c);
    long w = i / c;
    const int x = (int)(w % ox);
    w /= ox;
    const int y = (int)(w % oy);
    const long b = w / oy;
    float e = E[i];
    if (gate && !(gate[i] > 0.f)) e = 0.f;
    const unsigned long long mk = mask[i];
    float* dst = dX + ((b * oy * fd + (long)y * fd) * ix + (long)x * fd) * c + q;
    for (int dy = 0; dy < fd; ++dy)
      for (int dx = 0; dx < fd; ++dx) dst[((long)dy * ix + dx) * c] = ((mk >> (dy * fd + dx)) & 1ull) ? e : 0.f;
  }
}

// ---- global average pooling: layers/tensor.py:95-114 (X.mean(axis=(2,3)); backward repeats delta / (H*W))
// NCHW: one warp per (image, channel) plane, coalesced along the plane; channels-last: one thread per (image, channel),
// a warp reads 32 consecutive channels of each pixel.
__global__ void gap_fwd_nchw_kernel(const float* __restrict__ A, float* __restrict__ out, long planes, int hw) {
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long pl = warp; pl < planes; pl += nwarps) {
    const float* src = A + pl * hw;
    float s = 0.f;
    for (int i = lane; i < hw; i += 32) s += src[i];
    s = warp_sum(s);
    if (lane == 0) out[pl] = s / (float)hw;
  }
}
__global__ void gap_fwd_nhwc_kernel(const float* __restrict__ A, float* __restrict__ out, long total, int c, int hw) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const long b = i / c;
    const int ch = (int)(i - b * c);
    const float* src = A + b * hw * c + ch;
    float s = 0.f;
    for (int p = 0; p < hw; ++p) s += src[(long)p * c];
    out[i] = s / (float)hw;
  }
}
// dX[b, c, p] = delta[b, c] / hw   (either layout: idx -> (b, c))
__global__ void gap_bwd_kernel(const float* __restrict__ delta, float* __restrict__ dX, long total, int c, int hw, int nhwc) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  const float inv = 1.0f / (float)hw;
  for (; i < total; i += stride) {
    long bc;
    if (nhwc) { const long b = i / ((long)hw * c); bc = b * c + (i % c); }
    else bc = i / hw;
    dX[i] = delta[bc] * inv;
  }
}

static int pool_nhwc_check(const char* who, int im, int c, int iy, int ix, int fdim) {
  if (im < 0 || c <= 0 || iy <= 0 || ix <= 0) return fail(BF_ERR_ARG, "%s: bad shape", who);
  if (fdim < 1 || fdim > 8) return fail(BF_ERR_ARG, "%s: fdim %d not in [1,8]", who, fdim);
  if (iy % fdim || ix % fdim) return fail(BF_ERR_VALUE, "Incompatible shapes: (%d, %d) %% %d", ix, iy, fdim);
  if (c % 4) return fail(BF_ERR_ARG, "%s: channels-last pooling needs a multiple of 4 channels (got %d)", who, c);
  return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" {

int bf_layout_nchw_to_nhwc(const float* src, float* dst, int im, int c, int hw, void* stream) {
  BF_REQUIRE(im >= 0 && c > 0 && hw > 0, BF_ERR_ARG, "bf_layout_nchw_to_nhwc: bad shape");
  return batched_transpose(src, dst, im, c, hw, as_stream(stream));     // (im, c, hw) -> (im, hw, c)
}

int bf_layout_nhwc_to_nchw(const float* src, float* dst, int im, int c, int hw, void* stream) {
  BF_REQUIRE(im >= 0 && c > 0 && hw > 0, BF_ERR_ARG, "bf_layout_nhwc_to_nchw: bad shape");
  return batched_transpose(src, dst, im, hw, c, as_stream(stream));     // (im, hw, c) -> (im, c, hw)
}

int bf_pool_nhwc_forward(const float* A, float* out, uint8_t* mask, int im, int c, int iy, int ix, int fdim, int relu_in,
                         void* stream) {
  int rc = pool_nhwc_check("bf_pool_nhwc_forward", im, c, iy, ix, fdim);
  if (rc) return rc;
  if (im == 0) return BF_OK;
  BF_REQUIRE(A && out && mask, BF_ERR_ARG, "bf_pool_nhwc_forward: null pointer");
  const int oy = iy / fdim, ox = ix / fdim;
  cudaStream_t s = as_stream(stream);
  if (fdim * fdim <= 8) {
    const int c4 = c / 4;
    const long total = (long)im * oy * ox * c4;
    if (total == 0) return BF_OK;
    const int grid = ew_grid(total, 256);
#define BF_PF(FD_, R_) launch_pdl(pool_nhwc_fwd_kernel<FD_, R_>, dim3(grid), dim3(256), 0, s, (const float4*)A, (float4*)out, (uchar4*)mask, total, oy, ox, c4, ix)
    if (fdim == 2) { if (relu_in) BF_PF(2, true); else BF_PF(2, false); }
    else { if (relu_in) BF_PF(1, true); else BF_PF(1, false); }
#undef BF_PF
  } else {
    const long total = (long)im * oy * ox * c;
    if (total == 0) return BF_OK;
    const int grid = ew_grid(total, 256);
    if (relu_in) launch_pdl(pool_nhwc_fwd_wide_kernel<true>, dim3(grid), dim3(256), 0, s, A, out, (unsigned long long*)mask, total, oy, ox, c, ix, fdim);
    else launch_pdl(pool_nhwc_fwd_wide_kernel<false>, dim3(grid), dim3(256), 0, s, A, out, (unsigned long long*)mask, total, oy, ox, c, ix, fdim);
  }
  BF_LAUNCH_CHECK();
  return BF_OK;
}

// rows of the per-CTA bias-gradient partials bf_pool_nhwc_backward_db writes (0: that shape cannot produce them)
int bf_pool_nhwc_db_rows(int im, int c, int iy, int ix, int fdim) {
  if (fdim != 2 || im <= 0 || c <= 0 || (c & 3) || 256 % (c / 4) != 0 || (iy % 2) || (ix % 2)) return 0;
  return ew_grid((size_t)im * (iy / 2) * (ix / 2) * (c / 4), 256);
}

static int pool_nhwc_backward_impl(const float* E, const uint8_t* mask, const float* gate, float* dX, float* db_partial, int im,
                                   int c, int iy, int ix, int fdim, void* stream) {
  int rc = pool_nhwc_check("bf_pool_nhwc_backward", im, c, iy, ix, fdim);
  if (rc) return rc;
  if (im == 0) return BF_OK;
  BF_REQUIRE(E && dX && mask, BF_ERR_ARG, "bf_pool_nhwc_backward: null pointer");
  const int oy = iy / fdim, ox = ix / fdim;
  cudaStream_t s = as_stream(stream);
  if (fdim * fdim <= 8) {
    const int c4 = c / 4;
    const long total = (long)im * oy * ox * c4;
    if (total == 0) return BF_OK;
    const int grid = ew_grid(total, 256);
#define BF_PB(FD_, G_, D_) launch_pdl(pool_nhwc_bwd_kernel<FD_, G_, D_>, dim3(grid), dim3(256), 0, s, (const float4*)E, (const uchar4*)mask, (const float4*)gate, (float4*)dX, total, oy, ox, c4, ix, (float4*)db_partial)
    if (db_partial) {
      BF_REQUIRE(fdim == 2 && 256 % c4 == 0 && grid == bf_pool_nhwc_db_rows(im, c, iy, ix, fdim), BF_ERR_ARG,
                 "bf_pool_nhwc_backward: per-CTA bias-gradient partials need fdim 2 and a channel count dividing 1024");
      if (gate) BF_PB(2, true, true); else BF_PB(2, false, true);
    } else if (fdim == 2) { if (gate) BF_PB(2, true, false); else BF_PB(2, false, false); }
    else { if (gate) BF_PB(1, true, false); else BF_PB(1, false, false); }
#undef BF_PB
  } else {
    const long total = (long)im * oy * ox * c;
    if (total == 0) return BF_OK;
    launch_pdl(pool_nhwc_bwd_wide_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, s, E, (const unsigned long long*)mask, gate, dX, total, oy, ox, c, ix, fdim);
  }
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_pool_nhwc_backward(const float* E, const uint8_t* mask, const float* gate, float* dX, int im, int c, int iy, int ix,
                          int fdim, void* stream) {
  return pool_nhwc_backward_impl(E, mask, gate, dX, nullptr, im, c, iy, ix, fdim, stream);
}

int bf_pool_nhwc_backward_db(const float* E, const uint8_t* mask, const float* gate, float* dX, float* db_partial, int im, int c,
                             int iy, int ix, int fdim, void* stream) {
  BF_REQUIRE(db_partial && bf_pool_nhwc_db_rows(im, c, iy, ix, fdim) > 0, BF_ERR_ARG, "bf_pool_nhwc_backward_db: unsupported shape");
  return pool_nhwc_backward_impl(E, mask, gate, dX, db_partial, im, c, iy, ix, fdim, stream);
}

int bf_gap_forward(const float* A, float* out, int im, int c, int hw, int nhwc, void* stream) {
  BF_REQUIRE(im >= 0 && c > 0 && hw > 0, BF_ERR_ARG, "bf_gap_forward: bad shape");
  if (im == 0) return BF_OK;
  BF_REQUIRE(A && out, BF_ERR_ARG, "bf_gap_forward: null pointer");
  cudaStream_t s = as_stream(stream);
  const long planes = (long)im * c;
  if (nhwc) gap_fwd_nhwc_kernel<<<ew_grid(planes, 256), 256, 0, s>>>(A, out, planes, c, hw);
  else gap_fwd_nchw_kernel<<<ew_grid(planes * 32, 256), 256, 0, s>>>(A, out, planes, hw);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_gap_backward(const float* delta, float* dX, int im, int c, int hw, int nhwc, void* stream) {
  BF_REQUIRE(im >= 0 && c > 0 && hw > 0, BF_ERR_ARG, "bf_gap_backward: bad shape");
  if (im == 0) return BF_OK;
  BF_REQUIRE(delta && dX, BF_ERR_ARG, "bf_gap_backward: null pointer");
  const long total = (long)im * c * hw;
  gap_bwd_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(delta, dX, total, c, hw, nhwc);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // extern "C"
