// Max pooling with the reference's multi-hot equality mask: atomic/tensor_op.py:78-128,
// llatomic/lltensor_op.py:43-72.  Memory-bound: one pass over A, one word of mask per window.
//   mask word: bit (dy*fdim+dx) is set for EVERY element of the window equal to the window max
//   (ties are multi-hot, tensor_op.py:105-106).  1 byte/window when fdim*fdim <= 8, else 8 bytes.
#include "common.cuh"

namespace bf {

__host__ __device__ inline int pool_word_bytes(int fdim) { return fdim * fdim <= 8 ? 1 : 8; }

// fdim == 2 fast path: each thread owns one window, reads two float2 (coalesced: a warp covers
// 64 consecutive floats of each of the two input rows), writes one float and one mask byte.
__global__ void pool2_fwd_kernel(const float* __restrict__ A, float* __restrict__ out, uint8_t* __restrict__ mask,
                                 long windows, int oy, int ox, int ix) {
  long w = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; w < windows; w += stride) {
    int x = (int)(w % ox);
    long t = w / ox;
    int y = (int)(t % oy);
    long plane = t / oy;
    const float* src = A + (plane * (2 * oy) + 2 * y) * ix + 2 * x;
    float a, b, c, d;
    if ((ix & 1) == 0) {  // rows are 8-byte aligned
      float2 r0 = *reinterpret_cast<const float2*>(src);
      float2 r1 = *reinterpret_cast<const float2*>(src + ix);
      a = r0.x; b = r0.y; c = r1.x; d = r1.y;
    } else {
      a = src[0]; b = src[1]; c = src[ix]; d = src[ix + 1];
    }
    float mx = fmaxf(fmaxf(a, b), fmaxf(c, d));
    out[w] = mx;
    mask[w] = (uint8_t)((a == mx) | ((b == mx) << 1) | ((c == mx) << 2) | ((d == mx) << 3));
  }
}

__global__ void pool2_bwd_kernel(const float* __restrict__ E, const uint8_t* __restrict__ mask,
                                 float* __restrict__ dX, long windows, int oy, int ox, int ix) {
  long w = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; w < windows; w += stride) {
    int x = (int)(w % ox);
    long t = w / ox;
    int y = (int)(t % oy);
    long plane = t / oy;
    float e = E[w];
    unsigned mk = mask[w];
    float* dst = dX + (plane * (2 * oy) + 2 * y) * ix + 2 * x;
    float2 r0 = make_float2((mk & 1) ? e : 0.f, (mk & 2) ? e : 0.f);
    float2 r1 = make_float2((mk & 4) ? e : 0.f, (mk & 8) ? e : 0.f);
    if ((ix & 1) == 0) {
      *reinterpret_cast<float2*>(dst) = r0;
      *reinterpret_cast<float2*>(dst + ix) = r1;
    } else {
      dst[0] = r0.x; dst[1] = r0.y; dst[ix] = r1.x; dst[ix + 1] = r1.y;
    }
  }
}

// generic fdim <= 8
template <typename WORD>
__global__ void pool_fwd_kernel(const float* __restrict__ A, float* __restrict__ out, WORD* __restrict__ mask,
                                long windows, int oy, int ox, int ix, int fdim) {
  long w = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; w < windows; w += stride) {
    int x = (int)(w % ox);
    long t = w / ox;
    int y = (int)(t % oy);
    long plane = t / oy;
    const float* src = A + (plane * ((long)fdim * oy) + (long)fdim * y) * ix + (long)fdim * x;
    float mx = src[0];
    for (int dy = 0; dy < fdim; ++dy)
      for (int dx = 0; dx < fdim; ++dx) mx = fmaxf(mx, src[(long)dy * ix + dx]);
    unsigned long long mk = 0ull;
    for (int dy = 0; dy < fdim; ++dy)
      for (int dx = 0; dx < fdim; ++dx)
        if (src[(long)dy * ix + dx] == mx) mk |= 1ull << (dy * fdim + dx);
    out[w] = mx;
    mask[w] = (WORD)mk;
  }
}

template <typename WORD>
__global__ void pool_bwd_kernel(const float* __restrict__ E, const WORD* __restrict__ mask, float* __restrict__ dX,
                                long windows, int oy, int ox, int ix, int fdim) {
  long w = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; w < windows; w += stride) {
    int x = (int)(w % ox);
    long t = w / ox;
    int y = (int)(t % oy);
    long plane = t / oy;
    float e = E[w];
    unsigned long long mk = mask[w];
    float* dst = dX + (plane * ((long)fdim * oy) + (long)fdim * y) * ix + (long)fdim * x;
    for (int dy = 0; dy < fdim; ++dy)
      for (int dx = 0; dx < fdim; ++dx) dst[(long)dy * ix + dx] = ((mk >> (dy * fdim + dx)) & 1ull) ? e : 0.f;
  }
}

template <typename WORD>
__global__ void pool_expand_kernel(const WORD* __restrict__ mask, float* __restrict__ filt, long windows, int oy,
                                   int ox, int ix, int fdim) {
  long w = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; w < windows; w += stride) {
    int x = (int)(w % ox);
    long t = w / ox;
    int y = (int)(t % oy);
    long plane = t / oy;
    unsigned long long mk = mask[w];
    float* dst = filt + (plane * ((long)fdim * oy) + (long)fdim * y) * ix + (long)fdim * x;
    for (int dy = 0; dy < fdim; ++dy)
      for (int dx = 0; dx < fdim; ++dx) dst[(long)dy * ix + dx] = ((mk >> (dy * fdim + dx)) & 1ull) ? 1.f : 0.f;
  }
}

static int pool_check(const char* who, int im, int ic, int iy, int ix, int fdim) {
  if (im < 0 || ic <= 0 || iy <= 0 || ix <= 0) return fail(BF_ERR_ARG, "%s: bad shape", who);
  if (fdim < 1 || fdim > 8) return fail(BF_ERR_ARG, "%s: fdim %d not in [1,8]", who, fdim);
  if (iy % fdim || ix % fdim)  // layers/tensor.py:21-24
    return fail(BF_ERR_VALUE, "Incompatible shapes: (%d, %d) %% %d", ix, iy, fdim);
  return BF_OK;
}

}  // namespace bf

using namespace bf;

extern "C" {

size_t bf_pool_mask_bytes(int im, int ic, int iy, int ix, int fdim) {
  if (fdim < 1) return 0;
  return (size_t)im * ic * (iy / fdim) * (ix / fdim) * pool_word_bytes(fdim);
}

int bf_pool_forward(const float* A, float* out, uint8_t* mask, int im, int ic, int iy, int ix, int fdim, void* stream) {
  int rc = pool_check("bf_pool_forward", im, ic, iy, ix, fdim);
  if (rc) return rc;
  BF_REQUIRE(A && out && mask, BF_ERR_ARG, "bf_pool_forward: null pointer");
  int oy = iy / fdim, ox = ix / fdim;
  long windows = (long)im * ic * oy * ox;
  if (windows == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  int grid = ew_grid(windows, 256);
  if (fdim == 2)
    pool2_fwd_kernel<<<grid, 256, 0, s>>>(A, out, mask, windows, oy, ox, ix);
  else if (pool_word_bytes(fdim) == 1)
    pool_fwd_kernel<uint8_t><<<grid, 256, 0, s>>>(A, out, mask, windows, oy, ox, ix, fdim);
  else
    pool_fwd_kernel<unsigned long long><<<grid, 256, 0, s>>>(A, out, (unsigned long long*)mask, windows, oy, ox, ix, fdim);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_pool_backward(const float* E, const uint8_t* mask, float* dX, int im, int ic, int iy, int ix, int fdim,
                     void* stream) {
  int rc = pool_check("bf_pool_backward", im, ic, iy, ix, fdim);
  if (rc) return rc;
  BF_REQUIRE(E && dX && mask, BF_ERR_ARG, "bf_pool_backward: null pointer");
  int oy = iy / fdim, ox = ix / fdim;
  long windows = (long)im * ic * oy * ox;
  if (windows == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  int grid = ew_grid(windows, 256);
  if (fdim == 2)
    pool2_bwd_kernel<<<grid, 256, 0, s>>>(E, mask, dX, windows, oy, ox, ix);
  else if (pool_word_bytes(fdim) == 1)
    pool_bwd_kernel<uint8_t><<<grid, 256, 0, s>>>(E, mask, dX, windows, oy, ox, ix, fdim);
  else
    pool_bwd_kernel<unsigned long long><<<grid, 256, 0, s>>>(E, (const unsigned long long*)mask, dX, windows, oy, ox, ix, fdim);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_pool_mask_expand(const uint8_t* mask, float* filt, int im, int ic, int iy, int ix, int fdim, void* stream) {
  int rc = pool_check("bf_pool_mask_expand", im, ic, iy, ix, fdim);
  if (rc) return rc;
  int oy = iy / fdim, ox = ix / fdim;
  long windows = (long)im * ic * oy * ox;
  if (windows == 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  int grid = ew_grid(windows, 256);
  if (pool_word_bytes(fdim) == 1)
    pool_expand_kernel<uint8_t><<<grid, 256, 0, s>>>(mask, filt, windows, oy, ox, ix, fdim);
  else
    pool_expand_kernel<unsigned long long><<<grid, 256, 0, s>>>((const unsigned long long*)mask, filt, windows, oy, ox, ix, fdim);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

}  // extern "C"
