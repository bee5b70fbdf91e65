// First-layer convolution (few input channels, ic <= 4) on the tcgen05 tensor cores: NCHW input as the reference
// holds it, channels-last output.  K = ic*fy*fx is tiny (27 for the 3x3 RGB layer of cfg3), so both kernels are
// HBM-bound on the 32-channel side (the forward writes it, the filter gradient reads it) and are organised
// around moving those bytes at full rate.
//
//   forward:  the input is re-laid once as pixels x 4 channels ("NHWC4", 16 B per pixel).  With K-major
//             SWIZZLE_NONE core matrices whose rows OVERLAP (row pitch 16 B = one pixel, LBO = 16 B) the flat pixel
//             stream IS the patch matrix of one filter row: A[m][dx*4 + c] = X4[m + ky*ix + dx][c].  TMA drops
//             (128 + halo) pixels into shared memory, 2*fy MMAs (K = 8: two pixels x 4 channels) produce 128
//             output positions x nf filters; no im2col is ever formed (reference: tensor_op.py:9-32).
//   dF + db:  contraction over pixels.  E arrives by TMA as [pixel][32 filters] = MN-major operand (128B swizzle,
//             32B atoms), the patch rows (c,ky,kx) x 32 pixels are gathered by four SIMT warps straight from the
//             NCHW planes (27 coalesced 128-byte row reads per 32 pixels) into a K-major SWIZZLE_128B tile; an
//             all-ones row yields db.  One accumulator (rows = (c,ky,kx), columns = filters) lives in TMEM for
//             the whole kernel; CTAs split the pixels, a small kernel adds the partials (tensor_op.py:49-55).
#include "tma.cuh"
#include "conv_nhwc.cuh"
#include <stdlib.h>

namespace bf {

constexpr int C3_THREADS = 224;       // warp 0 TMA, warp 1 MMA, warp 2 spare, warps 3-6 epilogue
constexpr int C3_STAGE_LD = 36;
constexpr int C3_STAGING_BYTES = 4 * 32 * C3_STAGE_LD * 4;

struct C3FwdParams {
  int nimg, IR, ix, oy, ox, RT, tiles_per_img, ntiles, fy, fx, ic, N, box_rows, na;
};

// NCHW (im, ic, iy*ix) -> (im*iy*ix, 4): channel-padded channels-last copy of the input
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ X, float4* __restrict__ X4, long npix_total, int plane, int ic) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < npix_total; i += stride) {
    const long b = i / plane;
    const int q = (int)(i - b * plane);
    const float* src = X + b * ic * plane + q;
    float4 v = make_float4(src[0], ic > 1 ? src[plane] : 0.f, ic > 2 ? src[2 * (long)plane] : 0.f, ic > 3 ? src[3 * (long)plane] : 0.f);
    X4[i] = v;
  }
}

template <int NB, bool RELU>
__global__ void __launch_bounds__(C3_THREADS, 1)
conv_c3_fwd_kernel(const __grid_constant__ CUtensorMap mapX, C3FwdParams p, const float* __restrict__ F,
                   float* __restrict__ out, const float* __restrict__ bias) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t a_full[8], a_empty[8], t_full[2], t_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[NB];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t a_slot = (uint32_t)p.box_rows * 16u;
  const int Kt = p.fy * 16;                                   // K per output position: fy rows x (4 pixels x 4 channels)
  const uint32_t b_sbo = (uint32_t)(Kt / 4) * 128u;           // bytes per 8-filter group of the weight operand
  const uint32_t b_base = base + (uint32_t)p.na * a_slot;
  const uint32_t b_bytes = (uint32_t)(NB / 8) * b_sbo;
  const uint32_t stage_base = (b_base + b_bytes + 127u) & ~127u;
  const int n0 = blockIdx.y * NB;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    for (int i = 0; i < 8; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < NB) sbias[tid] = bias ? bias[n0 + tid] : 0.f;
  // weight operand W[n][ky*16 + dx*4 + c] = F[n][c][ky][dx], K-major SWIZZLE_NONE core matrices (8 filters x 16 B)
  for (int idx = tid; idx < NB * Kt; idx += C3_THREADS) {
    const int n = idx / Kt, k = idx - n * Kt;
    const int ky = k >> 4, dx = (k >> 2) & 3, c = k & 3;
    float v = 0.f;
    if (c < p.ic && dx < p.fx) v = F[(((long)(n0 + n) * p.ic + c) * p.fy + ky) * p.fx + dx];
    const uint32_t addr = b_base + (uint32_t)(n >> 3) * b_sbo + (uint32_t)(n & 7) * 16u + (uint32_t)(k >> 2) * 128u + (uint32_t)(k & 3) * 4u;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(to_tf32(v)) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    const bool leader = elect_one_sync();
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int as = it % p.na, au = it / p.na;
      const int b = tile / p.tiles_per_img, t = tile - b * p.tiles_per_img;
      if (au > 0) mbar_wait(&a_empty[as], (uint32_t)((au - 1) & 1));
      if (leader) {
        mbar_expect_tx(&a_full[as], a_slot);
        tma_load_2d(base + as * a_slot, &mapX, &a_full[as], 0, b * p.IR + t * p.RT * p.ix);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, NB);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int as = it % p.na, au = it / p.na, acc_set = it & 1, tu = it >> 1;
      if (tu > 0) mbar_wait(&t_empty[acc_set], (uint32_t)((tu - 1) & 1));
      mbar_wait(&a_full[as], (uint32_t)(au & 1));
      tc_fence_after();
      if (leader) {
        const uint32_t a0 = base + as * a_slot;
        for (int ky = 0; ky < p.fy; ++ky) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t ad = desc_k_none(a0 + (uint32_t)(ky * p.ix) * 16u + j * 32u, 16u, 128u);
            const uint64_t bd = desc_k_none(b_base + (uint32_t)(ky * 4 + j * 2) * 128u, 128u, b_sbo);
            umma_tf32(tmem + acc_set * NB, ad, bd, idesc, (ky > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(&a_empty[as]);
        umma_commit(&t_full[acc_set]);
      }
    }
  } else if (warp >= 3) {
    const int quad = warp & 3;
    float* stg = reinterpret_cast<float*>(dyn_smem) + (stage_base - smem_u32(dyn_smem)) / 4 + quad * 32 * C3_STAGE_LD;
    int it = 0;
    const int m = quad * 32 + lane;
    const int r = m / p.ix, x = m - r * p.ix;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc_set = it & 1, tu = it >> 1;
      const int b = tile / p.tiles_per_img, t = tile - b * p.tiles_per_img;
      const int y = t * p.RT + r;
      const bool ok = r < p.RT && x < p.ox && y < p.oy;
      const int orow = (b * p.oy + y) * p.ox + x;
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      mbar_wait(&t_full[acc_set], (uint32_t)(tu & 1));
      tc_fence_after();
      if (okmask != 0u) {
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + acc_set * NB;
#pragma unroll
        for (int cb = 0; cb < NB; cb += 32) {
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + cb + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(&sbias[cb + c0 + j]);
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
              *reinterpret_cast<float4*>(stg + lane * C3_STAGE_LD + c0 + j) = o;
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            const int dst_row = __shfl_sync(0xffffffffu, orow, rr);
            const float4 val = *reinterpret_cast<const float4*>(stg + rr * C3_STAGE_LD + (lane & 7) * 4);
            if ((okmask >> rr) & 1u)
              *reinterpret_cast<float4*>(out + (size_t)dst_row * p.N + n0 + cb + (lane & 7) * 4) = val;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc_set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

// --------------------------------------------------------------------------------------------- dF + db
constexpr int C3DF_THREADS = 224;     // warp 0 TMA, warps 1-2 MMA (even / odd chunks, one accumulator each), warps 3-6 epilogue
struct C3DfParams {
  int nimg, ic, iy, ix, oy, ox, fy, fx, nseg, nchunks, Mrows, ns, nf, KB;
  long long* prof;
};
#define C3_T(var) long long var = p.prof ? clock64() : 0
#define C3_ACC(slot, t0) do { if (p.prof && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.prof[slot] += clock64() - (t0); } while (0)

// TMA can only start a box on a 16-byte boundary of the innermost dimension (an unaligned coordinate traps), but a
// horizontal filter tap kx shifts the pixels by 4*kx bytes.  So the (small) input is re-laid once per call as fx
// copies, copy kx shifted left by kx pixels, rows padded to a multiple of 4 floats:
//   Xs[kx][plane row r][x] = X[r][x + kx]   (0 past the end of the row)
__global__ void shift_rows_kernel(const float* __restrict__ X, float* __restrict__ Xs, long rows, int ix, int pitch, int fx) {
  const long per = rows * pitch, total = per * fx;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int kx = (int)(i / per);
    const long j = i - (long)kx * per;
    const long r = j / pitch;
    const int x = (int)(j - r * pitch) + kx;
    Xs[i] = x < ix ? X[r * ix + x] : 0.f;
  }
}

// One chunk = 32 output pixels of one output row (b, y).  Patch tile (K-major SWIZZLE_128B, 128 B per row):
// row kx*KB + c*fy + ky = X[b][c][y+ky][x0+kx .. +31] (KB = ic*fy) -- all fx*ic*fy rows are ONE 5-D TMA box
// (32 x, fy rows, ic planes, 1 image, fx shifted copies); row Mrows (= fx*KB) is constant 1 (-> db).
template <int MH>
__global__ void __launch_bounds__(C3DF_THREADS, 1)
conv_c3_df_kernel(const __grid_constant__ CUtensorMap mapE, const __grid_constant__ CUtensorMap mapX, C3DfParams p,
                  float* __restrict__ partial) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  constexpr uint32_t A_SLOT = 128 * 128, B_SLOT = MH * 4096;
  const uint32_t b_base = base + (uint32_t)p.ns * A_SLOT;
  constexpr int NF = MH * 32;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // patch tiles: rows >= Mrows stay 0 except row Mrows = all ones (-> db); TMA only ever writes rows < Mrows
  for (uint32_t o = tid * 16u; o < (uint32_t)p.ns * A_SLOT; o += C3DF_THREADS * 16u) {
    const uint32_t row = (o % A_SLOT) / 128u;
    const uint32_t v = (row == (uint32_t)p.Mrows) ? 0x3f800000u : 0u;
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(v) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  int my_chunks = 0;
  for (int c = blockIdx.x; c < p.nchunks; c += gridDim.x) ++my_chunks;

  if (warp == 0) {
    const bool leader = elect_one_sync();
    const uint32_t kx_bytes = (uint32_t)p.KB * 128u, box_bytes = (uint32_t)(p.ic * p.fy) * 128u;
    int it = 0;
    for (int c = blockIdx.x; c < p.nchunks; c += gridDim.x, ++it) {
      const int s = it % p.ns, u = it / p.ns;
      const int seg = c % p.nseg, row = c / p.nseg;
      const int y = row % p.oy, b = row / p.oy;
      { C3_T(t0); if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1)); C3_ACC(0, t0); }
      if (leader) {
        mbar_expect_tx(&full[s], B_SLOT + (uint32_t)p.fx * box_bytes);
#pragma unroll
        for (int h = 0; h < MH; ++h) tma_load_4d(b_base + s * B_SLOT + h * 4096, &mapE, &full[s], h * 32, seg * 32, y, b);
        tma_load_5d(base + s * A_SLOT, &mapX, &full[s], seg * 32, y, 0, b, 0);
      }
    }
  } else if (warp == 1 || warp == 2) {
    const bool leader = elect_one_sync();
    const int wsel = warp - 1;
    constexpr uint32_t idesc = idesc_tf32(128, NF, false, true);
    int mine = 0;
    C3_T(tm);
    for (int it = wsel; it < my_chunks; it += 2, ++mine) {
      const int s = it % p.ns, u = it / p.ns;
      { C3_T(t0); mbar_wait(&full[s], (uint32_t)(u & 1)); if (wsel == 0) C3_ACC(1, t0); }
      tc_fence_after();
      if (leader) {
        const uint64_t ad = desc_k_sw128(base + s * A_SLOT);
        const uint64_t bd = desc_mn_sw128_32b(b_base + s * B_SLOT, 4096u, 512u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_tf32(tmem + wsel * NF, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 64), idesc, (mine > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
    }
    if (leader) umma_commit(&done);
    if (wsel == 0) C3_ACC(2, tm);
  } else if (warp >= 3) {
    // ---- final epilogue: partial[cta][m][NF] = accumulator of the even chunks + accumulator of the odd chunks
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    mbar_wait(&done, 0);
    tc_fence_after();
    float* dst = partial + ((size_t)blockIdx.x * 128 + m) * NF;
#pragma unroll
    for (int c0 = 0; c0 < NF; c0 += 16) {
      uint32_t v0[16], v1[16];
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = 0.f;
      if (my_chunks > 0) {
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + c0, v0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v0[j]);
      }
      if (my_chunks > 1) {
        tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + NF + c0, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] += __uint_as_float(v1[j]);
      }
      if (m <= p.Mrows) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

// dF[f][c][ky][kx] = sum_cta partial[cta][kx*KB + c*fy + ky][f];  db[f] = sum_cta partial[cta][fx*KB][f]
__global__ void conv_c3_df_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dF, float* __restrict__ db,
                                         int nctas, int ic, int fy, int fx, int nf, int KB) {
  const int Mrows = ic * fy * fx, total = nf * (Mrows + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int f = i / (Mrows + 1), r = i - f * (Mrows + 1);   // r = (c, ky, kx) in the reference's order, or Mrows
    int m = fx * KB;
    if (r < Mrows) {
      const int kx = r % fx, t = r / fx, ky = t % fy, c = t / fy;
      m = kx * KB + c * fy + ky;
    }
    float s = 0.f;
    for (int z = 0; z < nctas; ++z) s += partial[((size_t)z * 128 + m) * nf + f];
    if (r == Mrows) { if (db) db[f] = s; }
    else dF[(size_t)f * Mrows + r] = s;
  }
}

static inline size_t c3_al256(size_t x) { return (x + 255) & ~(size_t)255; }

bool conv_c3_supported(int ic, int iy, int ix, int nf, int fy, int fx) {
  if (!tma_encode_fn()) return false;
  if (ic < 1 || ic > 4 || fx < 1 || fx > 4 || fy < 1 || fy > 8) return false;
  if (!(nf == 32 || nf == 64 || nf == 128)) return false;
  if (ix < fx || iy < fy || ix > 120) return false;
  if (128 + (fy - 1) * ix + 4 > 256) return false;          // TMA box rows
  if (fx * ic * fy + 1 > 128) return false;  // patch rows of the dF accumulator
  return true;
}

}  // namespace bf

using namespace bf;

int bf_conv_c3_supported(int ic, int iy, int ix, int nf, int fy, int fx) { return conv_c3_supported(ic, iy, ix, nf, fy, fx) ? 1 : 0; }

// Y (im,oy,ox,nf) NHWC = act(valid_corr(X (im,ic,iy,ix) NCHW, F)) + bias
int bf_conv_c3_forward(const float* X, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix, int nf,
                       int fy, int fx, int act, void* stream) {
  BF_REQUIRE(conv_c3_supported(ic, iy, ix, nf, fy, fx), BF_ERR_ARG, "bf_conv_c3_forward: unsupported shape");
  BF_REQUIRE(act == BF_ACT_LINEAR || act == BF_ACT_RELU, BF_ERR_ARG, "bf_conv_c3_forward: only linear / relu are fused");
  if (im <= 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const int oy = iy - fy + 1, ox = ix - fx + 1, IR = iy * ix;
  BF_REQUIRE((long)im * IR < 2147483647L, BF_ERR_ARG, "bf_conv_c3_forward: too many pixels");
  void* ws = nullptr;
  int rc = workspace_require(c3_al256((size_t)im * IR * 16), &ws);
  if (rc) return rc;
  nchw_to_nhwc4_kernel<<<ew_grid((size_t)im * IR, 256), 256, 0, s>>>(X, (float4*)ws, (long)im * IR, IR, ic);
  BF_LAUNCH_CHECK();
  C3FwdParams p;
  p.nimg = im; p.IR = IR; p.ix = ix; p.oy = oy; p.ox = ox; p.RT = 128 / ix; if (p.RT > oy) p.RT = oy;
  p.tiles_per_img = (oy + p.RT - 1) / p.RT; p.ntiles = im * p.tiles_per_img; p.fy = fy; p.fx = fx; p.ic = ic; p.N = nf;
  p.box_rows = (128 + (fy - 1) * ix + 4 + 7) / 8 * 8; p.na = 8;
  CUtensorMap mX;
  {
    uint64_t dims[2] = {4, (uint64_t)im * IR};
    uint64_t str[1] = {4};
    uint32_t box[2] = {4, (uint32_t)p.box_rows};
    if (!tma_make_map(&mX, (const float*)ws, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
      return fail(BF_ERR_CUDA, "bf_conv_c3_forward: cuTensorMapEncodeTiled failed");
  }
  const size_t smem = 1024 + (size_t)p.na * p.box_rows * 16 + (size_t)nf * fy * 16 * 4 + 128 + C3_STAGING_BYTES;
  const int ctas = p.ntiles < num_sms() ? p.ntiles : num_sms();
  const bool relu = act == BF_ACT_RELU;
#define BF_C3_CASE(NB_)                                                                                                  \
  if (nf == NB_) {                                                                                                       \
    if (relu) {                                                                                                          \
      BF_CUDA(cudaFuncSetAttribute(conv_c3_fwd_kernel<NB_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      conv_c3_fwd_kernel<NB_, true><<<ctas, C3_THREADS, smem, s>>>(mX, p, F, Y, bias);                                   \
    } else {                                                                                                             \
      BF_CUDA(cudaFuncSetAttribute(conv_c3_fwd_kernel<NB_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      conv_c3_fwd_kernel<NB_, false><<<ctas, C3_THREADS, smem, s>>>(mX, p, F, Y, bias);                                  \
    }                                                                                                                    \
  }
  BF_C3_CASE(32) BF_C3_CASE(64) BF_C3_CASE(128)
#undef BF_C3_CASE
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}

// dF (nf,ic,fy,fx), db (nf) from X (im,ic,iy,ix) NCHW and E (im,oy,ox,nf) NHWC
int bf_conv_c3_backward_filter(const float* X, const float* E, float* dF, float* db, int im, int ic, int iy, int ix, int nf,
                               int fy, int fx, void* stream) {
  BF_REQUIRE(conv_c3_supported(ic, iy, ix, nf, fy, fx), BF_ERR_ARG, "bf_conv_c3_backward_filter: unsupported shape");
  cudaStream_t s = as_stream(stream);
  const int oy = iy - fy + 1, ox = ix - fx + 1, KB = ic * fy, Mrows = fx * KB;   // Mrows: index of the ones row
  if (im <= 0) {
    BF_CUDA(cudaMemsetAsync(dF, 0, (size_t)nf * ic * fy * fx * 4, s));
    if (db) BF_CUDA(cudaMemsetAsync(db, 0, (size_t)nf * 4, s));
    return BF_OK;
  }
  C3DfParams p;
  p.nimg = im; p.ic = ic; p.iy = iy; p.ix = ix; p.oy = oy; p.ox = ox; p.fy = fy; p.fx = fx; p.nseg = (ox + 31) / 32;
  BF_REQUIRE((long)im * oy * p.nseg < 2147483647L, BF_ERR_ARG, "bf_conv_c3_backward_filter: too many rows");
  p.nchunks = im * oy * p.nseg; p.Mrows = Mrows; p.nf = nf; p.KB = KB; p.prof = nullptr;
  static long long* d_prof = nullptr;
  const bool prof = getenv("BF_PROF_IGEMM") != nullptr;
  if (prof) {
    if (!d_prof) cudaMalloc(&d_prof, 16 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), s);
    p.prof = d_prof;
  }
  const int MH = nf / 32;
  p.ns = 8;
  while (p.ns > 2 && 1024 + (size_t)p.ns * (128 * 128 + MH * 4096) > 200 * 1024) --p.ns;
  const size_t smem = 1024 + (size_t)p.ns * (128 * 128 + MH * 4096);
  const int ctas = p.nchunks < num_sms() ? p.nchunks : num_sms();
  // workspace: [split-K partials | fx shifted, row-padded copies of X]
  const int pitch = (ix + 3) / 4 * 4;
  const long xrows = (long)im * ic * iy;
  const size_t part_bytes = c3_al256((size_t)ctas * 128 * nf * 4);
  const size_t xs_bytes = c3_al256((size_t)fx * xrows * pitch * 4);
  void* ws = nullptr;
  int rc = workspace_require(part_bytes + xs_bytes, &ws);
  if (rc) return rc;
  float* Xs = (float*)((char*)ws + part_bytes);
  shift_rows_kernel<<<ew_grid((size_t)fx * xrows * pitch, 256), 256, 0, s>>>(X, Xs, xrows, ix, pitch, fx);
  BF_LAUNCH_CHECK();
  CUtensorMap mE, mX;
  {
    uint64_t dims[4] = {(uint64_t)nf, (uint64_t)ox, (uint64_t)oy, (uint64_t)im};
    uint64_t str[3] = {(uint64_t)nf, (uint64_t)nf * ox, (uint64_t)nf * ox * oy};
    uint32_t box[4] = {32, 32, 1, 1};
    if (!tma_make_map(&mE, E, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return fail(BF_ERR_CUDA, "bf_conv_c3_backward_filter: cuTensorMapEncodeTiled failed (E)");
  }
  {
    uint64_t dims[5] = {(uint64_t)pitch, (uint64_t)iy, (uint64_t)ic, (uint64_t)im, (uint64_t)fx};
    uint64_t str[4] = {(uint64_t)pitch, (uint64_t)pitch * iy, (uint64_t)pitch * iy * ic, (uint64_t)pitch * xrows};
    uint32_t box[5] = {32, (uint32_t)fy, (uint32_t)ic, 1, (uint32_t)fx};
    if (!tma_make_map(&mX, Xs, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return fail(BF_ERR_CUDA, "bf_conv_c3_backward_filter: cuTensorMapEncodeTiled failed (X)");
  }
#define BF_C3DF_CASE(MH_)                                                                                            \
  if (MH == MH_) {                                                                                                   \
    BF_CUDA(cudaFuncSetAttribute(conv_c3_df_kernel<MH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    conv_c3_df_kernel<MH_><<<ctas, C3DF_THREADS, smem, s>>>(mE, mX, p, (float*)ws);                                  \
  }
  BF_C3DF_CASE(1) BF_C3DF_CASE(2) BF_C3DF_CASE(4)
#undef BF_C3DF_CASE
  BF_LAUNCH_CHECK();
  if (prof) {
    long long h[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[c3 dF prof, CTA 0, %d chunks] producer wait empty %lld | mma warp 1: wait full %lld of total %lld\n",
            (p.nchunks + ctas - 1) / ctas, h[0], h[1], h[2]);
  }
  conv_c3_df_reduce_kernel<<<ew_grid((size_t)nf * (ic * fy * fx + 1), 128), 128, 0, s>>>((const float*)ws, dF, db, ctas, ic, fy, fx, nf, KB);
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}
