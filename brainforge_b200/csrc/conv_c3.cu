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
//   dF + db:  contraction over pixels, in bands of RB output rows.  E is the MN-major operand (128B swizzle, 32B
//             atoms): either one tensor-TMA box per band, used in place, or -- the un-pooling variant -- rebuilt in
//             shared memory from the pooling layer's delta, tie mask and output, so that the full-resolution delta
//             never exists in HBM.  The raw NCHW rows of the band arrive by 1-D bulk copies; eight builder warps turn
//             them shared->shared into K-major SWIZZLE_128B patch tiles (row (c,ky,kx) x 32 pixels, plus an all-ones
//             row that yields db).  Two MMA-issuing warps (M = 64) accumulate in TMEM for the whole kernel; CTAs split
//             the bands, a small kernel adds the per-CTA partials (tensor_op.py:49-55).
#include "tma.cuh"
#include "conv_nhwc.cuh"
#include <stdlib.h>
#include <string.h>

namespace bf {

constexpr int C3_GROUPS = 4;          // epilogue groups of four warps; group g drains accumulator set g (tiles it = g mod 4)
constexpr int C3_THREADS = 96 + C3_GROUPS * 128;   // warp 0 loads, warps 1-2 MMA (even / odd tiles), then the epilogue groups
constexpr int C3_STAGE_LD = 36;
constexpr int C3_STAGE_TILE = 4 * 32 * C3_STAGE_LD;      // floats of one staged tile [128 positions][32 channels + pad]
constexpr int C3_STAGING_BYTES = C3_GROUPS * 2 * C3_STAGE_TILE * 4;   // two tiles per group: one barrier per drain (see the epilogue)

// 1-D bulk copy global -> shared (16-byte aligned on both sides, size a multiple of 16), completes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct C3FwdParams {
  int nimg, IR, ix, oy, ox, RT, tiles_per_img, ntiles, fy, fx, ic, N, box_rows, na;
  int pool_relu;            // POOL variant: apply max(0,.) before pooling (the Activation("relu") layer in between)
  uint8_t* mask;            // POOL variant: 2x2 tie mask, one byte per pooled element, (im, oy/2, ox/2, N)
  const float4* X4;         // the pixel stream (channel-padded channels-last copy of the input)
  long long npix;           // pixels in X4
};

// NCHW (im, ic, iy*ix) -> (im*iy*ix, 4): channel-padded channels-last copy of the input
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ X, float4* __restrict__ X4, long npix_total, int plane, int ic) {
  pdl_launch_dependents();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < npix_total; i += stride) {
    const long b = i / plane;
    const int q = (int)(i - b * plane);
    const float* src = X + b * ic * plane + q;
    // rounded to TF32 here (round-to-nearest; the MMA alone would truncate): the stream is only ever an MMA operand
    float4 v = make_float4(__uint_as_float(to_tf32(src[0])), ic > 1 ? __uint_as_float(to_tf32(src[plane])) : 0.f,
                           ic > 2 ? __uint_as_float(to_tf32(src[2 * (long)plane])) : 0.f,
                           ic > 3 ? __uint_as_float(to_tf32(src[3 * (long)plane])) : 0.f);
    X4[i] = v;
  }
}

// POOL: the epilogue max-pools 2x2 windows of the tile (staged in shared memory) and writes the pooled tensor + tie
// mask instead of the convolution output -- ConvLayer -> [ReLU] -> PoolLayer(2) in one pass (layers/tensor.py:28-30,
// 75-79; atomic/tensor_op.py:93-107): the full-resolution activation never goes to HBM.
template <int NB, bool RELU, bool POOL>
__global__ void __launch_bounds__(C3_THREADS, 1)
conv_c3_fwd_kernel(const __grid_constant__ CUtensorMap mapX, C3FwdParams p, const float* __restrict__ F,
                   float* __restrict__ out, const float* __restrict__ bias) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t a_full[32], a_empty[32], t_full[2 * C3_GROUPS], t_empty[2 * C3_GROUPS];
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
  constexpr int NSETS = 2 * C3_GROUPS * NB <= 512 ? 2 * C3_GROUPS : C3_GROUPS;     // accumulator sets in the 512 TMEM columns
  pdl_launch_dependents();

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * C3_GROUPS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    for (int i = 0; i < 32; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();          // everything above touched only this CTA's shared memory / TMEM
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
    int as = 0, au = 0;                // ring slot and its use count (running counters: no divisions by p.na in the loop)
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_img, t = tile - b * p.tiles_per_img;
      if (au > 0) mbar_wait(&a_empty[as], (uint32_t)((au - 1) & 1));
      if (leader) {
        // the tile's pixels + halo are ONE contiguous run of the stream: a 1-D bulk copy (a tensor box with a
        // 16-byte inner extent is serviced row by row and was the bottleneck of this kernel)
        const long long px0 = (long long)b * p.IR + (long long)t * p.RT * p.ix;
        long long npx = p.box_rows;
        if (px0 + npx > p.npix) npx = p.npix - px0;
        mbar_expect_tx(&a_full[as], (uint32_t)npx * 16u);
        bulk_load_1d(base + as * a_slot, p.X4 + px0, (uint32_t)npx * 16u, &a_full[as]);
      }
      if (++as == p.na) { as = 0; ++au; }
    }
  } else if (warp == 1 || warp == 2) {
    // two issuing warps, one per accumulator set (even / odd tiles): a single thread issues a K=8 MMA only every
    // ~50-100 cycles, which at 2*fy MMAs per 128-position tile was the limit of this kernel
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, NB);
    int it = 0, as = 0, au = 0;
    // descriptors: everything but the ring slot's offset is tile-invariant (the address field counts 16-byte units, so a
    // slot or filter-row offset is a plain add on the encoded descriptor)
    const uint64_t ad0 = desc_k_none(base, 16u, 128u), bd0 = desc_k_none(b_base, 128u, b_sbo);
    const uint32_t a_slot_u = a_slot >> 4, row_u = (uint32_t)p.ix;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it, as = (as + 1 == p.na ? 0 : as + 1), au += (as == 0)) {
      if ((it & 1) != warp - 1) continue;
      // NSETS accumulator sets: tile it -> set it % NSETS, drained by epilogue group it % C3_GROUPS.  With two sets per group
      // (NB <= 64: 8 x NB <= 512 TMEM columns) the MMAs of a group's NEXT tile are issued while it drains the current one
      // (ncu: with one set per group the epilogue warps spent a third of their time waiting for the accumulator)
      const int acc_set = it % NSETS, tu = it / NSETS;
      if (tu > 0) mbar_wait(&t_empty[acc_set], (uint32_t)((tu - 1) & 1));
      mbar_wait(&a_full[as], (uint32_t)(au & 1));
      tc_fence_after();
      if (leader) {
        const uint64_t a0 = ad0 + (uint64_t)((uint32_t)as * a_slot_u);
        const uint32_t d = tmem + acc_set * NB;
        if (p.fy == 3) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            umma_tf32(d, a0 + (uint64_t)(ky * row_u), bd0 + (uint64_t)(ky * 32), idesc, ky > 0 ? 1u : 0u);
            umma_tf32(d, a0 + (uint64_t)(ky * row_u + 2u), bd0 + (uint64_t)(ky * 32 + 16), idesc, 1u);
          }
        } else {
          for (int ky = 0; ky < p.fy; ++ky) {
            umma_tf32(d, a0 + (uint64_t)(ky * row_u), bd0 + (uint64_t)(ky * 32), idesc, ky > 0 ? 1u : 0u);
            umma_tf32(d, a0 + (uint64_t)(ky * row_u + 2u), bd0 + (uint64_t)(ky * 32 + 16), idesc, 1u);
          }
        }
        umma_commit(&a_empty[as]);
        umma_commit(&t_full[acc_set]);
      }
    }
  } else if (warp >= 3) {
    // C3_GROUPS epilogue groups of four warps (one warp per TMEM lane quadrant): group g drains accumulator set g,
    // so that the drains of consecutive tiles overlap (one drain is a serial TMEM -> shared -> global chain)
    // The staged tile is double-buffered: drain k writes buffer k & 1, ONE group barrier separates its writes from its
    // pooled reads, and a thread that has passed the barrier of drain k + 1 knows every thread finished reading drain k.
    const int quad = warp & 3, grp = (warp - 3) >> 2;
    float* stg0 = reinterpret_cast<float*>(dyn_smem) + (stage_base - smem_u32(dyn_smem)) / 4 + grp * 2 * C3_STAGE_TILE + quad * 32 * C3_STAGE_LD;
    int it = 0, drain = 0;
    // bias of a 32-channel block in registers (loads from shared memory between the staging stores would be serialised
    // behind them: the compiler must assume they alias)
    float4 bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = *reinterpret_cast<const float4*>(&sbias[4 * j]);
    // (image, tile of the image) of this group's next tile, advanced without divisions
    const int step_b = (4 * (int)gridDim.x) / p.tiles_per_img, step_t = (4 * (int)gridDim.x) - step_b * p.tiles_per_img;
    int b = 0, t = 0;
    {
      const int first = (int)blockIdx.x + grp * (int)gridDim.x;
      b = first / p.tiles_per_img; t = first - b * p.tiles_per_img;
    }
    const int m = quad * 32 + lane;
    const int r = m / p.ix, x = m - r * p.ix;
    // POOL: tile-invariant geometry of the (at most two) pooled items this thread produces per tile and channel block:
    // item = (pooled position of the tile, channel quad), quads fastest
    bool pi_ok[2];
    int pi_y2[2], pi_soff[2], pi_ooff[2];
    {
      const int pw = p.ox >> 1, npool = (p.RT >> 1) * pw;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int item = m + i * 128;
        const int q = item & 7, pp = item >> 3;
        const int py = pp / (pw > 0 ? pw : 1), pxx = pp - py * pw;
        pi_ok[i] = POOL && item < npool * 8;
        pi_y2[i] = 2 * py;
        pi_soff[i] = (2 * py * p.ix + 2 * pxx) * C3_STAGE_LD + q * 4;
        pi_ooff[i] = (py * pw + pxx) * p.N + q * 4;
      }
    }
    for (int tile = (int)blockIdx.x + grp * (int)gridDim.x; tile < p.ntiles;
         tile += 4 * (int)gridDim.x, ++it, b += step_b, t += step_t, b += (t >= p.tiles_per_img), t -= (t >= p.tiles_per_img) ? p.tiles_per_img : 0) {
      const int acc_set = NSETS == C3_GROUPS ? grp : grp + C3_GROUPS * (it & 1), tu = NSETS == C3_GROUPS ? it : it >> 1;
      const int y = t * p.RT + r;
      const bool ok = r < p.RT && x < p.ox && y < p.oy;
      const int orow = (b * p.oy + y) * p.ox + x;
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      mbar_wait(&t_full[acc_set], (uint32_t)(tu & 1));
      tc_fence_after();
      if (POOL || okmask != 0u) {
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + acc_set * NB;
#pragma unroll
        for (int cb = 0; cb < NB; cb += 32, ++drain) {
          float* stg = stg0 + (drain & 1) * C3_STAGE_TILE;
          if (NB > 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bb[j] = *reinterpret_cast<const float4*>(&sbias[cb + 4 * j]);
          }
          {
            uint32_t v[32];
            tmem_ld32(taddr + cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              o.x += bb[j >> 2].x; o.y += bb[j >> 2].y; o.z += bb[j >> 2].z; o.w += bb[j >> 2].w;
              if (POOL && p.pool_relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              *reinterpret_cast<float4*>(stg + lane * C3_STAGE_LD + j) = o;
            }
          }
          if (POOL) {
            // the four warps of the group have staged the whole tile [128 positions][32 channels]
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            const float* tile_stg = stg - quad * 32 * C3_STAGE_LD;
            const size_t obase = ((size_t)(b * (p.oy >> 1) + ((t * p.RT) >> 1)) * (p.ox >> 1)) * p.N + n0 + cb;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (pi_ok[i] && t * p.RT + pi_y2[i] + 1 < p.oy) {
                const float* s00 = tile_stg + pi_soff[i];
                const float4 a = *reinterpret_cast<const float4*>(s00);
                const float4 bq = *reinterpret_cast<const float4*>(s00 + C3_STAGE_LD);
                const float4 c = *reinterpret_cast<const float4*>(s00 + p.ix * C3_STAGE_LD);
                const float4 d = *reinterpret_cast<const float4*>(s00 + (p.ix + 1) * C3_STAGE_LD);
                float4 mx;
                mx.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(c.x, d.x)); mx.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(c.y, d.y));
                mx.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(c.z, d.z)); mx.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(c.w, d.w));
                uchar4 mk;
                mk.x = (unsigned char)((a.x == mx.x) | ((bq.x == mx.x) << 1) | ((c.x == mx.x) << 2) | ((d.x == mx.x) << 3));
                mk.y = (unsigned char)((a.y == mx.y) | ((bq.y == mx.y) << 1) | ((c.y == mx.y) << 2) | ((d.y == mx.y) << 3));
                mk.z = (unsigned char)((a.z == mx.z) | ((bq.z == mx.z) << 1) | ((c.z == mx.z) << 2) | ((d.z == mx.z) << 3));
                mk.w = (unsigned char)((a.w == mx.w) | ((bq.w == mx.w) << 1) | ((c.w == mx.w) << 2) | ((d.w == mx.w) << 3));
                const size_t o = obase + pi_ooff[i];
                *reinterpret_cast<float4*>(out + o) = mx;
                *reinterpret_cast<uchar4*>(p.mask + o) = mk;
              }
            }
          } else {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc_set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// --------------------------------------------------------------------------------------------- dF + db
// Work unit = one BAND of RB output rows of one image (px = RB*ox pixels, cut into k-blocks of 32 pixels).
//   warp 0      producer: per band one tensor-TMA box of E ([px][32 filters] per 32-filter group, MN-major
//               SWIZZLE_128B/32B-atom operand exactly as it lies in HBM) + one 1-D bulk copy per input plane of the
//               (RB+fy-1) raw NCHW rows the band's patches touch (16-byte aligned start, `lead` floats of slack);
//   warps 4-7   builders: turn the raw rows into the K-major SWIZZLE_128B patch tile of one k-block
//               A[r = (c,ky,kx)][k = pixel] = X[c][y+ky][x+kx]  (row Mrows = 1 -> db; padding pixels = 0),
//               shared -> shared, TF32-rounded, conflict-free; a ring of `na` tiles decouples them from the MMAs;
//   warps 1-2   MMA issuers: even / odd k-blocks, one TMEM accumulator each (two issuing threads reach twice the
//               issue rate of one for these small N), four K=8 tcgen05.mma per k-block;
//   warps 4-7   final epilogue: sum of the two accumulators -> per-CTA partial; a small kernel adds the partials.
constexpr int C3DF_BUILDERS = 8;        // warps per builder group
constexpr int C3DF_ISSUERS = 3;         // MMA-issuing warps (1-3), one accumulator each: k-block j of a band goes to issuer j % 3
// warp 0 loads, 1-3 MMA (3 also owns the TMEM allocation), 4-11 patch-tile builders (4-7 also drain the accumulators); the un-pooling variant
// adds warps 12-19, which build the E tile concurrently (one group doing both, in sequence, was the kernel's limit)
constexpr int c3df_threads(bool unpool) { return 128 + (unpool ? 2 : 1) * C3DF_BUILDERS * 32; }
constexpr int C3DF_A_SLOT = 32 * 128;     // 32 patch rows per tile; the M=128 descriptor runs on into the next tiles (rows >= 32 of D are never read)
// after the ring: zeros, >= 4 KB (the M=64 descriptor of the last tile reads 8 KB) and >= one raw stage (source of padding lanes)
struct C3DfParams {
  int nimg, ic, iy, ix, oy, ox, fy, fx, nf;
  int RB, nbands, nunits, px_band, nkb, Mrows, ns, na;
  int e_stage_bytes, x_plane_f, x_stage_bytes;
  long long x_total_f;
  int a_tail_bytes;
  // UNPOOL variant: E is not read from HBM but rebuilt per band from the pooling layer's state
  const float* dP;          // delta w.r.t. the pooled output   (im, oy/2, ox/2, nf)
  const float* gate;        // pooled forward output (ReLU gate: delta passes where gate > 0), or null
  const uint8_t* mask;      // 2x2 tie mask, one byte per pooled element
  int pool_bytes;           // bytes of dP (= of gate) of one full band; a load stage = [raw rows | dP | gate | mask]
  int stage_bytes;          // stride of the load ring (x_stage_bytes, + the pooled-resolution state when un-pooling)
};

// UNPOOL: the E operand (delta w.r.t. the convolution output) is never materialised in HBM: the builder warps
// scatter `dP * mask * (gate > 0)` -- MaxPoolOp.backward followed by the ReLU layer's derivative
// (atomic/tensor_op.py:110-119, layers/core.py:51-52) -- straight into the MN-major SWIZZLE_128B/32B-atom tile the
// MMA reads.  The band's dP / gate / mask slices are contiguous in HBM and arrive by 1-D bulk copies in the load ring.
// Per band that is 16 KB of reads instead of a 25 KB E box, and the 411 MB write + read of the full-resolution
// delta disappears.
template <int MH, int NR, bool UNPOOL>
__global__ void __launch_bounds__(c3df_threads(UNPOOL), 1)
conv_c3_df_kernel(const __grid_constant__ CUtensorMap mapE, const float* __restrict__ X, C3DfParams p, float* __restrict__ partial) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[6], empty[6], built[4], afree[4], done;
  __shared__ uint32_t tmem_slot;
  __shared__ int roff[128];
  __shared__ int acc_used[C3DF_ISSUERS];
  constexpr int NF = MH * 32;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  // shared memory: [E tiles: ns load stages (TMA) or na band tiles (UNPOOL)][na band tiles of nkb patch tiles][zeros][raw-row ring]
  const uint32_t e_base = base;
  const uint32_t a_base = e_base + (uint32_t)(UNPOOL ? p.na : p.ns) * p.e_stage_bytes;
  const uint32_t zero_base = a_base + (uint32_t)(p.na * p.nkb) * C3DF_A_SLOT;
  const uint32_t x_base = zero_base + (uint32_t)p.a_tail_bytes;

  pdl_launch_dependents();
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(4 * NF));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    constexpr int NBW = (UNPOOL ? 2 : 1) * C3DF_BUILDERS;      // builder warps that arrive per band
    for (int i = 0; i < 6; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], (UNPOOL ? 0 : C3DF_ISSUERS) + NBW); }
    for (int i = 0; i < 4; ++i) { mbar_init(&built[i], NBW); mbar_init(&afree[i], C3DF_ISSUERS); }
    mbar_init(&done, 2 * C3DF_ISSUERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // E stages and patch tiles start as zeros: rows the loads / builders never write must read as finite zeros
  for (uint32_t o = tid * 16u; o < x_base - base; o += blockDim.x * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
  // row r = (c*fy + ky)*fx + kx (the reference's filter order) -> offset of its first element inside the raw stage
  for (int r = tid; r < 128; r += blockDim.x) {
    int v = 0;
    if (r < p.Mrows) {
      const int kx = r % p.fx, t = r / p.fx, ky = t % p.fy, c = t / p.fy;
      v = (c << 24) | (c * p.x_plane_f + ky * p.ix + kx);
    }
    roff[r] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  pdl_wait();          // the prologue above touched only this CTA's shared memory / TMEM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  int my_units = 0;
  for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) ++my_units;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    const bool leader = elect_one_sync();
    int s = 0, u = 0;                  // load-ring slot and its use count (running counters: no divisions in the loop)
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const int b = unit / p.nbands, band = unit - b * p.nbands, y0 = band * p.RB;
      if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
      if (leader) {
        int rows_in = p.RB + p.fy - 1;
        if (rows_in > p.iy - y0) rows_in = p.iy - y0;
        uint32_t bytes = UNPOOL ? 0u : (uint32_t)MH * (uint32_t)p.px_band * 128u;
        uint32_t xb[4];
        long long xo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          xb[c] = 0; xo[c] = 0;
          if (c < p.ic) {
            const long long off = ((long long)(b * p.ic + c) * p.iy + y0) * p.ix;
            const long long st = off & ~3ll;
            long long nfl = ((off - st) + (long long)rows_in * p.ix + 3) & ~3ll;
            if (st + nfl > p.x_total_f) nfl = p.x_total_f - st;
            xo[c] = st; xb[c] = (uint32_t)nfl * 4u;
            bytes += xb[c];
          }
        }
        uint32_t pb = 0;
        size_t pe = 0;
        if (UNPOOL) {
          int rbp = p.RB;
          if (rbp > p.oy - y0) rbp = p.oy - y0;
          pb = (uint32_t)((rbp >> 1) * (p.ox >> 1) * p.nf) * 4u;                                  // bytes of dP of this band
          pe = ((size_t)b * (p.oy >> 1) + (y0 >> 1)) * (size_t)(p.ox >> 1) * p.nf;                // first pooled element
          bytes += pb + (p.gate ? pb : 0u) + pb / 4u;
        }
        mbar_expect_tx(&full[s], bytes);
        const uint32_t st0 = x_base + (uint32_t)s * p.stage_bytes;
        if (!UNPOOL) {
          const int px0 = (b * p.oy + y0) * p.ox;
#pragma unroll
          for (int h = 0; h < MH; ++h)
            tma_load_2d(e_base + s * p.e_stage_bytes + h * p.nkb * 4096, &mapE, &full[s], h * 32, px0);
        }
        if (UNPOOL) {
          bulk_load_1d(st0 + p.x_stage_bytes, p.dP + pe, pb, &full[s]);
          if (p.gate) bulk_load_1d(st0 + p.x_stage_bytes + p.pool_bytes, p.gate + pe, pb, &full[s]);
          bulk_load_1d(st0 + p.x_stage_bytes + 2 * p.pool_bytes, p.mask + pe, pb / 4u, &full[s]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < p.ic) bulk_load_1d(st0 + c * p.x_plane_f * 4, X + xo[c], xb[c], &full[s]);
      }
      if (++s == p.ns) { s = 0; ++u; }
    }
  } else if (warp >= 1 && warp <= C3DF_ISSUERS) {
    // ------------------------------------------------------------------ MMA issuers: k-block j of a band -> issuer j % 3
    // (one thread issues a K=8 MMA only every 50-70 cycles; the tensor pipe takes an M=64 x N=32 one every ~24)
    const bool leader = elect_one_sync();
    const int wsel = warp - 1;
    constexpr uint32_t idesc = idesc_tf32(64, NF, false, true);   // M=64: half the shared-memory reads of M=128 per MMA
    int mine = 0, s = 0, u = 0, t = 0, ut = 0;
    for (int it = 0; it < my_units; ++it) {
      if (!UNPOOL) mbar_wait(&full[s], (uint32_t)(u & 1));          // E landed
      mbar_wait(&built[t], (uint32_t)(ut & 1));        // patch tiles (and, UNPOOL, the E tile) of the band written
      tc_fence_after();
      if (leader) {
        const int unit = blockIdx.x + it * gridDim.x, band = unit % p.nbands;
        int rb = p.RB;
        if (rb > p.oy - band * p.RB) rb = p.oy - band * p.RB;
        const int nkb_eff = (rb * p.ox + 31) >> 5;         // k-blocks that hold pixels of this band
        const uint32_t e_tile = e_base + (uint32_t)(UNPOOL ? t : s) * p.e_stage_bytes;
        bool any = false;
        for (int j = wsel; j < nkb_eff; j += C3DF_ISSUERS) {
          const uint64_t ad = desc_k_sw128(a_base + (uint32_t)(t * p.nkb + j) * C3DF_A_SLOT);
          const uint64_t bd = desc_mn_sw128_32b(e_tile + j * 4096, (uint32_t)p.nkb * 4096u, 512u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(tmem + wsel * NF, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 64), idesc, (mine > 0 || ks > 0) ? 1u : 0u);
          ++mine;
          any = true;
        }
        if (any) { umma_commit(&afree[t]); if (!UNPOOL) umma_commit(&empty[s]); }
        else { mbar_arrive(&afree[t]); if (!UNPOOL) mbar_arrive(&empty[s]); }
      }
      __syncwarp();
      if (++s == p.ns) { s = 0; ++u; }
      if (++t == p.na) { t = 0; ++ut; }
    }
    if (leader) {
      acc_used[wsel] = mine > 0 ? 1 : 0;     // an accumulator nobody wrote holds garbage: the epilogue must skip it
      umma_commit(&done);                    // arrives when this warp's MMAs have completed
      mbar_arrive(&done);                    // releases the flag write
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ patch-tile builders
    // builder warp bw writes the k-block tiles j = bw, bw+8, .. of every band (hand-off to the MMA warps once per
    // band).  Lane = pixel of the k-block.  Per patch row: add, shared load, round (add half an ulp of TF32; the
    // MMA drops the low 13 bits), conflict-free swizzled store.  Every per-row offset is band-invariant and pinned
    // in a register (planes are multiples of 4 floats, host-checked, so the alignment slack `lead` of the raw rows
    // is the same for every plane).  Lanes beyond the band's pixels read zeros from the never-written tail of the
    // tile ring.  Rows Mrows+1 .. NR-1 of a tile are don't-care rows (their accumulator rows are never read).
    const int bw = warp - 4;
    uint32_t rowoff[NR];                 // byte offset of row r's first element inside a raw stage
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      rowoff[r] = r < p.Mrows ? (uint32_t)(roff[r] & 0xffffff) * 4u : 0u;
      asm volatile("" : "+r"(rowoff[r]));          // keep it in a register (no rematerialisation from the table)
    }
    const uint32_t sw = lane >> 2, l3 = (lane & 3) * 4;
    uint32_t swz[8];                     // byte offset of this lane's 16-byte chunk in a tile row with (r & 7) = i
#pragma unroll
    for (int i = 0; i < 8; ++i) { swz[i] = ((sw ^ i) << 4) + l3; asm volatile("" : "+r"(swz[i])); }
    const uint32_t ones_off = (uint32_t)p.Mrows * 128u + (((sw ^ (uint32_t)(p.Mrows & 7)) << 4) + l3);
    const uint32_t zero_src = zero_base;     // a_tail_bytes of zeros
    // band-invariant geometry, computed once: this lane's pixel of the warp's own k-block (j = bw), and for the
    // un-pooling variant the four E-tile addresses of each pooled item this thread expands
    const bool do_a = bw < C3DF_BUILDERS, do_e = UNPOOL && bw >= C3DF_BUILDERS;
    const int bi = do_e ? bw - C3DF_BUILDERS : bw;           // index inside the group
    const int pxA = bi * 32 + lane, yyA = pxA / p.ox;
    const uint32_t offA = (uint32_t)(yyA * p.ix + (pxA - yyA * p.ox)) * 4u;
    constexpr int UI = 2;                                     // items per builder thread (host checks the band fits)
    const int pw = p.ox >> 1, nq = p.nf >> 2;
    uint32_t eoff[UI][4];
    if (UNPOOL) {
#pragma unroll
      for (int i = 0; i < UI; ++i) {
        const int item = (bi * 32 + lane) + i * (C3DF_BUILDERS * 32);
        const int q = item % nq, pp = item / nq;
        const int py = pp / pw, pxx = pp - py * pw;
        const int f0 = q * 4;
        const uint32_t sub = (uint32_t)(f0 >> 5) * (uint32_t)p.nkb * 4096u;
        const uint32_t chunk = (uint32_t)((f0 & 31) >> 3), inner = (uint32_t)(f0 & 7) * 4u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {           // window element k = dy*2 + dx -> conv-output pixel of the band
          const int pl = (2 * py + (k >> 1)) * p.ox + 2 * pxx + (k & 1);
          eoff[i][k] = sub + (uint32_t)pl * 128u + ((chunk ^ (uint32_t)(pl & 3)) << 5) + inner;
        }
      }
    }
    int s = 0, u = 0, t = 0, ut = 0;
    for (int unit = blockIdx.x; unit < p.nunits; unit += gridDim.x) {
      const int band = unit % p.nbands, y0 = band * p.RB;
      int rb = p.RB;
      if (rb > p.oy - y0) rb = p.oy - y0;
      const int npx = rb * p.ox;
      const int lead = (y0 * p.ix) & 3;
      const uint32_t xs = x_base + (uint32_t)s * p.stage_bytes;
      const int nitems = UNPOOL ? (rb >> 1) * pw * nq : 0;      // (pooled pixel of the band, channel quad), quads fastest
      mbar_wait(&full[s], (uint32_t)(u & 1));                       // raw rows (and the pooled-resolution state) landed
      if (ut > 0) mbar_wait(&afree[t], (uint32_t)((ut - 1) & 1));   // MMAs of the band tile's previous use are done
      const int nkb_eff = (npx + 31) >> 5;
      for (int j = do_a ? bi : nkb_eff; j < nkb_eff; j += C3DF_BUILDERS) {
        const int px = j * 32 + lane;
        const bool valid = px < npx;
        uint32_t off = offA;
        if (j != bi) { const int yy = px / p.ox; off = (uint32_t)(yy * p.ix + (px - yy * p.ox)) * 4u; }
        const uint32_t src = valid ? xs + (uint32_t)lead * 4u + off : zero_src;
        const uint32_t tile = a_base + (uint32_t)(t * p.nkb + j) * C3DF_A_SLOT;
        uint32_t v[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v[r]) : "r"(src + rowoff[r]));
#pragma unroll
        for (int r = 0; r < NR; ++r)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + swz[r & 7] + (uint32_t)(r * 128)), "r"(v[r] + 0x1000u));
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + ones_off), "r"(valid ? 0x3f800000u : 0u));
      }
      if (do_e) {
        const uint32_t e_tile = e_base + (uint32_t)t * p.e_stage_bytes;
        const uint32_t dp_s = xs + (uint32_t)p.x_stage_bytes, g_s = dp_s + (uint32_t)p.pool_bytes, m_s = g_s + (uint32_t)p.pool_bytes;
#pragma unroll
        for (int i = 0; i < UI; ++i) {
          const int item = (bi * 32 + lane) + i * (C3DF_BUILDERS * 32);
          if (item < nitems) {
            uint32_t d0, d1, d2, d3, g0 = 0x3f800000u, g1 = 0x3f800000u, g2 = 0x3f800000u, g3 = 0x3f800000u, mw;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(dp_s + (uint32_t)item * 16u));
            if (p.gate)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(g0), "=r"(g1), "=r"(g2), "=r"(g3) : "r"(g_s + (uint32_t)item * 16u));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(mw) : "r"(m_s + (uint32_t)item * 4u));
            // relu'(gate) folded in, then TF32 rounding (+half ulp; the MMA drops the low 13 bits)
            d0 = __uint_as_float(g0) > 0.f ? d0 + 0x1000u : 0u;
            d1 = __uint_as_float(g1) > 0.f ? d1 + 0x1000u : 0u;
            d2 = __uint_as_float(g2) > 0.f ? d2 + 0x1000u : 0u;
            d3 = __uint_as_float(g3) > 0.f ? d3 + 0x1000u : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t addr = e_tile + eoff[i][k];
              const uint32_t e0 = (mw >> k) & 1u ? d0 : 0u, e1 = (mw >> (8 + k)) & 1u ? d1 : 0u;
              const uint32_t e2 = (mw >> (16 + k)) & 1u ? d2 : 0u, e3 = (mw >> (24 + k)) & 1u ? d3 : 0u;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(e0), "r"(e1), "r"(e2), "r"(e3));
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&built[t]); mbar_arrive(&empty[s]); }   // tiles written; raw rows no longer read
      if (++s == p.ns) { s = 0; ++u; }
      if (++t == p.na) { t = 0; ++ut; }
    }
    // ---- final epilogue: partial[cta][r][NF] = accumulator of the even k-blocks + accumulator of the odd ones
    // M=64 accumulator (cta_group::1): row r lives in TMEM lane (r/16)*32 + r%16, 16 rows per 32-lane quadrant
    // (pinned on a B200: the other candidate, lane = r, fails parity)
    const int m = lane < 16 ? bw * 16 + lane : 1 << 20;
    if (bw < 4 && bw * 16 <= p.Mrows) {
      mbar_wait(&done, 0);
      tc_fence_after();
      float* dst = partial + ((size_t)blockIdx.x * (p.Mrows + 1) + m) * NF;
#pragma unroll
      for (int c0 = 0; c0 < NF; c0 += 16) {
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = 0.f;
#pragma unroll
        for (int a = 0; a < C3DF_ISSUERS; ++a) {
          if (acc_used[a]) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(bw * 32) << 16) + a * NF + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] += __uint_as_float(v[j]);
          }
        }
        if (m <= p.Mrows) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(4 * NF));
}

// dF[f][r] = sum_cta partial[cta][r][f] (r = (c,ky,kx) in the reference's order);  db[f] = sum_cta partial[cta][Mrows][f]
// block = 32 consecutive outputs x 32 slices of the CTA range (a thread adds ~5 partials: the kernel is a chain of
// dependent L2 latencies, so the chain is kept short); a fixed-order shared-memory tree adds the slices.
constexpr int C3_RED_SLICES = 32;
__global__ void __launch_bounds__(32 * C3_RED_SLICES)
conv_c3_df_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dF, float* __restrict__ db,
                         int nctas, int Mrows, int nf) {
  __shared__ float red[C3_RED_SLICES][33];
  pdl_launch_dependents();
  pdl_wait();
  const int total = nf * (Mrows + 1);
  const int o = blockIdx.x * 32 + threadIdx.x, zs = threadIdx.y;
  float s = 0.f;
  if (o < total)
    for (int z = zs; z < nctas; z += C3_RED_SLICES) s += partial[(size_t)z * total + o];
  red[zs][threadIdx.x] = s;
  __syncthreads();
#pragma unroll
  for (int w = C3_RED_SLICES / 2; w >= 1; w >>= 1) {
    if (zs < w) red[zs][threadIdx.x] += red[zs + w][threadIdx.x];
    __syncthreads();
  }
  if (zs == 0 && o < total) {
    const float sum = red[0][threadIdx.x];
    const int r = o / nf, f = o - r * nf;        // partial rows are [r][f]: consecutive threads -> consecutive filters
    if (r == Mrows) { if (db) db[f] = sum; }
    else dF[(size_t)f * Mrows + r] = sum;
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

static int c3_forward_impl(const float* X, const float* F, const float* bias, float* Y, uint8_t* mask, int pool, int pool_relu,
                           int im, int ic, int iy, int ix, int nf, int fy, int fx, int act, cudaStream_t s) {
  const int oy = iy - fy + 1, ox = ix - fx + 1, IR = iy * ix;
  BF_REQUIRE((long)im * IR < 2147483647L, BF_ERR_ARG, "bf_conv_c3_forward: too many pixels");
  void* ws = nullptr;
  int rc = workspace_require(c3_al256((size_t)im * IR * 16), &ws);
  if (rc) return rc;
  launch_pdl(nchw_to_nhwc4_kernel, dim3(ew_grid((size_t)im * IR, 256)), dim3(256), 0, s, X, (float4*)ws, (long)im * IR, IR, ic);
  BF_LAUNCH_CHECK();
  C3FwdParams p;
  p.nimg = im; p.IR = IR; p.ix = ix; p.oy = oy; p.ox = ox; p.RT = 128 / ix; if (p.RT > oy) p.RT = oy;
  if (pool) {
    p.RT &= ~1;              // tiles start on even rows and hold whole 2x2 windows
    BF_REQUIRE(p.RT >= 2 && (oy & 1) == 0 && (ox & 1) == 0, BF_ERR_ARG, "bf_conv_c3_forward_pool: plane %dx%d cannot be pooled in-tile", oy, ox);
  }
  p.tiles_per_img = (oy + p.RT - 1) / p.RT; p.ntiles = im * p.tiles_per_img; p.fy = fy; p.fx = fx; p.ic = ic; p.N = nf;
  p.box_rows = (128 + (fy - 1) * ix + 4 + 7) / 8 * 8;
  p.na = 16;   // tiles in flight (8..32 measured equal: the kernel is not load-bound)
  while (p.na > 4 && (size_t)p.na * p.box_rows * 16 > 96 * 1024) --p.na;
  p.pool_relu = pool_relu; p.mask = mask;
  p.X4 = (const float4*)ws; p.npix = (long long)im * IR;
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
#define BF_C3_LAUNCH(NB_, R_, P_)                                                                                           \
  {                                                                                                                         \
    BF_CUDA(cudaFuncSetAttribute(conv_c3_fwd_kernel<NB_, R_, P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    launch_pdl(conv_c3_fwd_kernel<NB_, R_, P_>, dim3(ctas), dim3(C3_THREADS), smem, s, mX, p, F, Y, bias);                                      \
  }
#define BF_C3_CASE(NB_)                                                                              \
  if (nf == NB_) {                                                                                   \
    if (pool) { if (relu) BF_C3_LAUNCH(NB_, true, true) else BF_C3_LAUNCH(NB_, false, true) }        \
    else { if (relu) BF_C3_LAUNCH(NB_, true, false) else BF_C3_LAUNCH(NB_, false, false) }           \
  }
  BF_C3_CASE(32) BF_C3_CASE(64) BF_C3_CASE(128)
#undef BF_C3_CASE
#undef BF_C3_LAUNCH
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}

// Y (im,oy,ox,nf) NHWC = act(valid_corr(X (im,ic,iy,ix) NCHW, F)) + bias
int bf_conv_c3_forward(const float* X, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix, int nf,
                       int fy, int fx, int act, void* stream) {
  BF_REQUIRE(conv_c3_supported(ic, iy, ix, nf, fy, fx), BF_ERR_ARG, "bf_conv_c3_forward: unsupported shape");
  BF_REQUIRE(act == BF_ACT_LINEAR || act == BF_ACT_RELU, BF_ERR_ARG, "bf_conv_c3_forward: only linear / relu are fused");
  if (im <= 0) return BF_OK;
  return c3_forward_impl(X, F, bias, Y, nullptr, 0, 0, im, ic, iy, ix, nf, fy, fx, act, as_stream(stream));
}

// P (im,oy/2,ox/2,nf), mask = maxpool2x2( [relu]( act(valid_corr(X, F)) + bias ) ): the convolution output itself is never
// written.  relu_in != 0: an Activation("relu") layer sits between the convolution and the pooling layer.
int bf_conv_c3_forward_pool(const float* X, const float* F, const float* bias, float* P, uint8_t* mask, int im, int ic, int iy,
                            int ix, int nf, int fy, int fx, int act, int relu_in, void* stream) {
  BF_REQUIRE(conv_c3_supported(ic, iy, ix, nf, fy, fx), BF_ERR_ARG, "bf_conv_c3_forward_pool: unsupported shape");
  BF_REQUIRE(act == BF_ACT_LINEAR || act == BF_ACT_RELU, BF_ERR_ARG, "bf_conv_c3_forward_pool: only linear / relu are fused");
  BF_REQUIRE(P && mask, BF_ERR_ARG, "bf_conv_c3_forward_pool: null pointer");
  if (im <= 0) return BF_OK;
  return c3_forward_impl(X, F, bias, P, mask, 1, relu_in ? 1 : 0, im, ic, iy, ix, nf, fy, fx, act, as_stream(stream));
}

static int c3_backward_filter_impl(const float* X, const float* E, const float* dP, const uint8_t* mask, const float* gate,
                                   int unpool, float* dF, float* db, int im, int ic, int iy, int ix, int nf, int fy, int fx,
                                   cudaStream_t s) {
  const int oy = iy - fy + 1, ox = ix - fx + 1, Mrows = ic * fy * fx;
  if (im <= 0) {
    BF_CUDA(cudaMemsetAsync(dF, 0, (size_t)nf * ic * fy * fx * 4, s));
    if (db) BF_CUDA(cudaMemsetAsync(db, 0, (size_t)nf * 4, s));
    return BF_OK;
  }
  const int MH = nf / 32;
  const int cap_px = nf == 32 ? 224 : nf == 64 ? 128 : 64;     // E tile <= 32 KB
  const long long x_total = (long long)im * ic * iy * ix;
  BF_REQUIRE(Mrows + 1 <= 32, BF_ERR_ARG, "bf_conv_c3_backward_filter: more than 31 patch rows (ic*fy*fx = %d)", Mrows);
  BF_REQUIRE(((iy * ix) & 3) == 0, BF_ERR_ARG, "bf_conv_c3_backward_filter: plane size %dx%d is not a multiple of 4", iy, ix);
  BF_REQUIRE(ox <= cap_px && (x_total & 3) == 0 && ((uintptr_t)X & 15) == 0, BF_ERR_ARG,
             "bf_conv_c3_backward_filter: unsupported plane / alignment");
  BF_REQUIRE((long long)im * oy * ox < 2147483647LL, BF_ERR_ARG, "bf_conv_c3_backward_filter: too many pixels");
  C3DfParams p;
  p.nimg = im; p.ic = ic; p.iy = iy; p.ix = ix; p.oy = oy; p.ox = ox; p.fy = fy; p.fx = fx; p.nf = nf;
  p.RB = cap_px / ox; if (p.RB > oy) p.RB = oy;
  if (unpool) {
    p.RB &= ~1;              // bands hold whole 2x2 windows
    BF_REQUIRE(p.RB >= 2 && (oy & 1) == 0 && (ox & 1) == 0 && ((uintptr_t)dP & 15) == 0 && ((uintptr_t)mask & 3) == 0 &&
                   ((uintptr_t)gate & 15) == 0 && (p.RB / 2) * (ox / 2) * (nf / 4) <= 2 * C3DF_BUILDERS * 32,
               BF_ERR_ARG, "bf_conv_c3_backward_filter_unpool: plane %dx%d x %d filters cannot be un-pooled in-band", oy, ox, nf);
  } else {
    BF_REQUIRE(((uintptr_t)E & 15) == 0, BF_ERR_ARG, "bf_conv_c3_backward_filter: E is not 16-byte aligned");
  }
  p.nbands = (oy + p.RB - 1) / p.RB;
  p.nunits = im * p.nbands;
  p.px_band = p.RB * ox;
  p.nkb = (p.px_band + 31) / 32;
  p.Mrows = Mrows;
  p.e_stage_bytes = MH * p.nkb * 4096;
  p.x_plane_f = ((p.RB + fy - 1) * ix + 3 + 3) & ~3;
  p.x_stage_bytes = ic * p.x_plane_f * 4;
  p.x_total_f = x_total;
  p.a_tail_bytes = p.x_stage_bytes > 4096 ? (p.x_stage_bytes + 1023) / 1024 * 1024 : 4096;
  p.dP = dP; p.gate = gate; p.mask = mask;
  p.pool_bytes = unpool ? (p.RB / 2) * (ox / 2) * nf * 4 : 0;
  p.stage_bytes = p.x_stage_bytes + (unpool ? 2 * p.pool_bytes + p.pool_bytes / 4 : 0);
  // ring depths: na band tiles of patch tiles (+ E tiles when un-pooling), ns load stages (E box + raw rows, or raw rows only)
  auto smem_of = [&](int ns, int na) {
    return (size_t)1024 + (size_t)(unpool ? na : ns) * p.e_stage_bytes + (size_t)na * p.nkb * C3DF_A_SLOT + p.a_tail_bytes +
           (size_t)ns * p.stage_bytes;
  };
  p.na = 2;
  p.ns = 6;
  while (p.ns > 2 && smem_of(p.ns, p.na) > 224 * 1024) --p.ns;
  if (smem_of(p.ns, p.na) > 224 * 1024 && p.na > 2) p.na = 2;
  BF_REQUIRE(smem_of(p.ns, p.na) <= 224 * 1024, BF_ERR_ARG, "bf_conv_c3_backward_filter: stage does not fit shared memory");
  const size_t smem = smem_of(p.ns, p.na);
  const int ctas = p.nunits < num_sms() ? p.nunits : num_sms();
  void* ws = nullptr;
  int rc = workspace_require(c3_al256((size_t)ctas * (Mrows + 1) * nf * 4), &ws);
  if (rc) return rc;
  CUtensorMap mE;
  memset(&mE, 0, sizeof(mE));
  if (!unpool) {
    uint64_t dims[2] = {(uint64_t)nf, (uint64_t)im * oy * ox};
    uint64_t str[1] = {(uint64_t)nf};
    uint32_t box[2] = {32, (uint32_t)p.px_band};
    if (!tma_make_map(&mE, E, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return fail(BF_ERR_CUDA, "bf_conv_c3_backward_filter: cuTensorMapEncodeTiled failed (E)");
  }
  const int NR = Mrows <= 12 ? 12 : Mrows <= 20 ? 20 : Mrows <= 27 ? 27 : 31;   // patch rows a builder writes per tile
#define BF_C3DF_LAUNCH(MH_, NR_, U_)                                                                                          \
  {                                                                                                                           \
    BF_CUDA(cudaFuncSetAttribute(conv_c3_df_kernel<MH_, NR_, U_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    launch_pdl(conv_c3_df_kernel<MH_, NR_, U_>, dim3(ctas), dim3(c3df_threads(U_)), smem, s, mE, X, p, (float*)ws);                                   \
  }
#define BF_C3DF_CASE(MH_, NR_) \
  if (MH == MH_ && NR == NR_) { if (unpool) BF_C3DF_LAUNCH(MH_, NR_, true) else BF_C3DF_LAUNCH(MH_, NR_, false) }
  BF_C3DF_CASE(1, 12) BF_C3DF_CASE(1, 20) BF_C3DF_CASE(1, 27) BF_C3DF_CASE(1, 31)
  BF_C3DF_CASE(2, 12) BF_C3DF_CASE(2, 20) BF_C3DF_CASE(2, 27) BF_C3DF_CASE(2, 31)
  BF_C3DF_CASE(4, 12) BF_C3DF_CASE(4, 20) BF_C3DF_CASE(4, 27) BF_C3DF_CASE(4, 31)
#undef BF_C3DF_CASE
#undef BF_C3DF_LAUNCH
  BF_LAUNCH_CHECK();
  launch_pdl(conv_c3_df_reduce_kernel, dim3((nf * (Mrows + 1) + 31) / 32), dim3(32, C3_RED_SLICES), 0, s, (const float*)ws, dF, db, ctas, Mrows, nf);
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}

// dF (nf,ic,fy,fx), db (nf) from X (im,ic,iy,ix) NCHW and E (im,oy,ox,nf) NHWC
int bf_conv_c3_backward_filter(const float* X, const float* E, float* dF, float* db, int im, int ic, int iy, int ix, int nf,
                               int fy, int fx, void* stream) {
  BF_REQUIRE(conv_c3_supported(ic, iy, ix, nf, fy, fx), BF_ERR_ARG, "bf_conv_c3_backward_filter: unsupported shape");
  return c3_backward_filter_impl(X, E, nullptr, nullptr, nullptr, 0, dF, db, im, ic, iy, ix, nf, fy, fx, as_stream(stream));
}

// Same result with E = MaxPoolOp.backward(dP, mask) [* (gate > 0)] formed on the fly (2x2 pooling, channels-last pooled
// tensors): the filter gradient of ConvLayer -> [Activation("relu")] -> PoolLayer(2) straight from the pooling layer's delta.
int bf_conv_c3_backward_filter_unpool(const float* X, const float* dP, const uint8_t* mask, const float* gate, float* dF,
                                      float* db, int im, int ic, int iy, int ix, int nf, int fy, int fx, void* stream) {
  BF_REQUIRE(conv_c3_supported(ic, iy, ix, nf, fy, fx), BF_ERR_ARG, "bf_conv_c3_backward_filter_unpool: unsupported shape");
  BF_REQUIRE(im <= 0 || (dP && mask), BF_ERR_ARG, "bf_conv_c3_backward_filter_unpool: null pointer");
  return c3_backward_filter_impl(X, nullptr, dP, mask, gate, 1, dF, db, im, ic, iy, ix, nf, fy, fx, as_stream(stream));
}
