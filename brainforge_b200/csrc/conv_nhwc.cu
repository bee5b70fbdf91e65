// Channels-last (NHWC) implicit-GEMM convolution on the tcgen05 tensor cores, operands staged by TMA only.
//
// Why channels-last: with the activations stored as [pixel][channel] a 32-channel slice of a pixel is one
// 128-byte row of the canonical SWIZZLE_128B shared-memory layout, so TMA drops a (rows x width x 32 channel)
// box of the tensor into shared memory in exactly the form tcgen05.mma reads, and a filter tap (ky,kx) is a
// ROW OFFSET of ky*pitch + kx into that staged box: the nine taps of a 3x3 filter are nine matrix descriptors
// over the same bytes.  No im2col, no SIMT instruction touches an operand (reference: atomic/tensor_op.py:9-61
// builds the patch matrix with a Python loop, llatomic/lltensor_op.py:9-40 materialises it).
//
//   forward / dX ("igemm"):  out[q, n] = sum_{tap, c} in[q + shift(tap), c] * Wt[tap][n][c]
//        q runs over the staged box in row-major (pitch PW) order; positions whose x >= out_x (or y >= out rows)
//        are computed and dropped.  The "full" correlation of the data gradient (tensor_op.py:34-40,56-60) is the
//        same kernel: TMA's out-of-bounds zero fill supplies the padding (box origin at -(f-1)), the flipped and
//        transposed filter is a different repack of Wt.
//   dF (+db):  dF[tap][f, c] = sum_{b, q} E[b, q, f] * X[b, q + shift(tap), c]
//        both operands are MN-major here (the contraction index is the pixel = the row); kind::tf32 takes
//        MN-major operands in the 128B-swizzle/32B-atom layout that TMA's SWIZZLE_128B_ATOM_32B writes.
//        E is loaded with the pitch of X (TMA zero-fills x >= ox, y >= oy), one accumulator per tap lives in
//        TMEM for the whole kernel, an all-ones operand adds db as a tenth accumulator.
//
// Pipeline (both kernels): warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-5 = epilogue
// (TMEM -> registers -> global); mbarrier rings between them; persistent CTAs, one per SM.
#include "tma.cuh"
#include "conv_nhwc.cuh"
#include <stdlib.h>

namespace bf {

constexpr int IG_THREADS = 352;   // igemm: warp 0 TMA, warps 1-2 MMA issuers (tiles of even / odd index), warps 3-6 / 7-10 two epilogue groups (tiles of even / odd index)
constexpr int DF_THREADS = 192;
constexpr int IG_ACCSET_COLS = 256;   // two accumulator sets in the 512 TMEM columns

struct IgemmParams {
  int nstages, nbands, G, IR, PW, T, slot_rows, ntaps, fx;
  int x0, y0, RO, out_y, out_x, nimg, N, relu, nbs;
  int resident;   // the whole repacked filter stays in shared memory: no per-tap traffic or waits
  int nkp;        // K passes per stage: the contraction channels are staged KH*32 at a time (C = nkp * KH * 32)
  int na;         // depth of the A (activation tile) ring: HBM latency is hidden by na-1 tiles in flight
  int fy;         // filter rows
  int TS;         // positions between consecutive tiles of a stage: 128, or 128 - (fx - 1) when the filter columns are folded
  int pool_relu;  // POOL: max(0, .) between the convolution (+ bias) and the pooling (an Activation("relu") layer in between)
  uint8_t* pool_mask;   // POOL: 2x2 tie mask, one byte per pooled element, (im, out_y/2, out_x/2, N)
  long long* prof;   // debug (BF_PROF_IGEMM=1): per-role cycle counters of CTA 0, else null
};

#define IG_T(var) long long var = p.prof ? clock64() : 0
#define IG_ACC(slot, t0) do { if (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0) p.prof[slot] += clock64() - (t0); } while (0)
constexpr int IG_STAGE_LD = 36;                                  // floats per staged epilogue row (32 + 4 pad)
constexpr int IG_STAGING_BYTES = 8 * 32 * IG_STAGE_LD * 4;       // eight epilogue warps

// FOLD: the fx taps of a filter row are folded into the N dimension of ONE MMA -- D'[q][tx*NB + n] = sum_c in[q + ty*PW][c] *
// Wt[(ty,tx)][n][c] -- and the epilogue adds the fx column blocks with a shift of tx positions: out[q][n] = sum_tx D'[q + tx]
// [tx*NB + n].  With NB = 32 an unfolded MMA (128 x 32 x 8) is bound by the 4 KB shared-memory read of its A operand
// (40 clk for 16 clk of tensor work); folding amortises that read over 3x the columns and issues a third of the MMAs.
// Tiles of a stage then overlap by fx - 1 positions (stride TS = 128 - (fx - 1)) so that every sum stays inside one tile.
constexpr int IG_FOLD_ROWS = 132, IG_FOLD_LD = 20;      // 16 columns at a time (+4 pad: conflict-free float4 rows)
// POOL: ConvLayer -> [ReLU] -> PoolLayer(2) in one pass (layers/tensor.py:28-30,75-79; atomic/tensor_op.py:93-107): the
// epilogue stages the stage's positions (whole images) 16 channels at a time in shared memory, max-pools the 2x2 windows
// and writes the pooled tensor + tie mask; the full-resolution activation never goes to HBM.  `out` is the pooled tensor.
constexpr int IG_POOL_LD = 20;
template <int NB, int KH, bool RELU, bool FOLD = false, bool POOL = false>
__global__ void __launch_bounds__(IG_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, IgemmParams p,
                  float* __restrict__ out, const float* __restrict__ bias) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t a_full[8], a_empty[8], t_full[2], t_empty[2], b_full[16], b_empty[16];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[NB];
  // warp-uniform role index (the shuffle tells the compiler it is uniform: the TMA / MMA operands then live in
  // uniform registers instead of being broadcast lane by lane before every UTMALDG / UTCHMMA)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t half_bytes = (uint32_t)p.slot_rows * 128u;
  const uint32_t a_slot_bytes = KH * half_bytes;
  const uint32_t b_base0 = base + (uint32_t)p.na * a_slot_bytes;
  constexpr uint32_t B_SLOT = KH * NB * 128;
  const uint32_t stage_base = b_base0 + (uint32_t)p.nbs * B_SLOT;
  const int n0 = blockIdx.y * NB;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 2); mbar_init(&t_empty[i], 8); }
    for (int i = 0; i < 8; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); }
    for (int i = 0; i < 16; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_launch_dependents();
  // rows of the A slots that TMA never writes are read by the taps of dropped positions only; give them a
  // defined value once so that no NaN pattern lingers there
  for (uint32_t o = tid * 16u; o < (uint32_t)p.na * a_slot_bytes; o += IG_THREADS * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
  pdl_wait();          // the prologue above touched only this CTA's shared memory / TMEM
  if (tid < NB) sbias[tid] = bias ? bias[n0 + tid] : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp loops, one
    // elected lane issues)
    const bool leader = elect_one_sync();
    int as = 0, au = 0;            // A-ring slot and its use count; B-ring likewise (running counters, no divisions)
    uint32_t bs = 0, bu = 0;
    IG_T(tp);
    if (p.resident && leader) {
      mbar_expect_tx(&b_full[0], (uint32_t)(p.ntaps * p.nkp) * B_SLOT);
      if (FOLD) {     // [kp][ty][h][tx][n]: the fx taps of a row are one box (fx * NB consecutive operand rows)
        for (int kp = 0; kp < p.nkp; ++kp)
          for (int ty = 0; ty < p.fy; ++ty)
#pragma unroll
            for (int h = 0; h < KH; ++h)
              tma_load_3d(b_base0 + (uint32_t)(((kp * p.fy + ty) * KH + h) * p.fx) * (NB * 128), &mapB, &b_full[0], (kp * KH + h) * 32,
                          n0, ty * p.fx);
      } else {
        for (int kp = 0; kp < p.nkp; ++kp)
          for (int tap = 0; tap < p.ntaps; ++tap)
#pragma unroll
            for (int h = 0; h < KH; ++h)
              tma_load_3d(b_base0 + (kp * p.ntaps + tap) * B_SLOT + h * NB * 128, &mapB, &b_full[0], (kp * KH + h) * 32, n0, tap);
      }
    }
    for (int st = blockIdx.x; st < p.nstages; st += gridDim.x) {
      const int grp = st / p.nbands, band = st - grp * p.nbands;
      for (int kp = 0; kp < p.nkp; ++kp) {
        { IG_T(t0); if (au > 0) mbar_wait(&a_empty[as], (uint32_t)((au - 1) & 1)); IG_ACC(0, t0); }
        if (leader) {
          mbar_expect_tx(&a_full[as], (uint32_t)(KH * p.G * p.IR * 128));
#pragma unroll
          for (int h = 0; h < KH; ++h)
            tma_load_4d(base + as * a_slot_bytes + h * half_bytes, &mapA, &a_full[as], (kp * KH + h) * 32, p.x0,
                        p.y0 + band * p.RO, grp * p.G);
        }
        if (!p.resident) {
          for (int tap = 0; tap < p.ntaps; ++tap) {
            if (bu > 0) mbar_wait(&b_empty[bs], (bu - 1) & 1);
            if (leader) {
              mbar_expect_tx(&b_full[bs], B_SLOT);
#pragma unroll
              for (int h = 0; h < KH; ++h)
                tma_load_3d(b_base0 + bs * B_SLOT + h * NB * 128, &mapB, &b_full[bs], (kp * KH + h) * 32, n0, tap);
            }
            if (++bs == (uint32_t)p.nbs) { bs = 0; ++bu; }
          }
        }
        if (++as == p.na) { as = 0; ++au; }
      }
    }
    IG_ACC(1, tp);
  } else if (warp == 1 || warp == 2) {
    // ===================================================================== MMA issuers: warp 1 takes the even tiles of
    // every stage, warp 2 the odd ones (one thread cannot issue tcgen05.mma faster than ~1 per 50-100 cycles,
    // scripts/umma_rate2.cu); both wait on the same full barriers and both arrive on the empty ones
    const int wsel = warp - 1;
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = idesc_tf32(128, NB);
    const uint32_t half_units = half_bytes >> 4;
    uint32_t bsm = 0, bum = 0;       // B-ring slot / use count of the next tap (running counters)
    int it = 0, as = 0, au = 0;
    IG_T(tm);
    for (int st = blockIdx.x; st < p.nstages; st += gridDim.x, ++it) {
      const int acc_set = it & 1, tu = it >> 1;
      { IG_T(t0); if (tu > 0) mbar_wait(&t_empty[acc_set], (uint32_t)((tu - 1) & 1)); if (wsel == 0) IG_ACC(2, t0); }
      const uint32_t acc = tmem + acc_set * IG_ACCSET_COLS;
      for (int kp = 0; kp < p.nkp; ++kp) {
        { IG_T(t0); mbar_wait(&a_full[as], (uint32_t)(au & 1)); if (wsel == 0) IG_ACC(3, t0); }
        tc_fence_after();
        const uint64_t a_desc0 = desc_k_sw128(base + as * a_slot_bytes);
        if (FOLD) {
          if (as == 0 && au == 0) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
          const uint32_t idesc_f = idesc_tf32(128, NB * p.fx);
          const uint32_t nbf = (uint32_t)(NB * p.fx);
          for (int ty = 0; ty < p.fy; ++ty) {
            const uint64_t a_row = a_desc0 + (uint64_t)((uint32_t)(ty * p.PW) * 8u);
            const uint32_t b_row = b_base0 + (uint32_t)((kp * p.fy + ty) * KH * p.fx) * (NB * 128);
            for (int t = wsel; t < p.T; t += 2) {
#pragma unroll
              for (int h = 0; h < KH; ++h) {
                const uint64_t ad = a_row + (uint64_t)(h * half_units + (uint32_t)(t * p.TS) * 8u);
                const uint64_t bd = desc_k_sw128(b_row + (uint32_t)(h * p.fx) * (NB * 128));
                if (leader) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_tf32(acc + t * nbf, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc_f,
                              (kp > 0 || ty > 0 || h > 0 || ks > 0) ? 1u : 0u);
                }
              }
            }
          }
          if (leader) umma_commit(&a_empty[as]);
          if (++as == p.na) { as = 0; ++au; }
          continue;
        }
        int ty = 0, tx = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          uint32_t bs;
          if (p.resident) {
            bs = (uint32_t)(kp * p.ntaps + tap);
            if (as == 0 && au == 0 && tap == 0) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
          } else {
            bs = bsm;
            mbar_wait(&b_full[bs], bum & 1);
            tc_fence_after();
            if (++bsm == (uint32_t)p.nbs) { bsm = 0; ++bum; }
          }
          const uint64_t a_tap = a_desc0 + (uint64_t)((uint32_t)(ty * p.PW + tx) * 8u);   // 128 B per pixel row = 8 units
          const uint64_t b_tap = desc_k_sw128(b_base0 + bs * B_SLOT);
          if (++tx == p.fx) { tx = 0; ++ty; }
          for (int t = wsel; t < p.T; t += 2) {
#pragma unroll
            for (int h = 0; h < KH; ++h) {
              const uint64_t ad = a_tap + (uint64_t)(h * half_units + (uint32_t)t * 1024u);
              const uint64_t bd = b_tap + (uint64_t)(h * NB * 8);
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_tf32(acc + t * NB, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc,
                            (kp > 0 || tap > 0 || h > 0 || ks > 0) ? 1u : 0u);
              }
            }
          }
          if (!p.resident && leader) umma_commit(&b_empty[bs]);
        }
        if (leader) umma_commit(&a_empty[as]);
        if (++as == p.na) { as = 0; ++au; }
      }
      if (leader) umma_commit(&t_full[acc_set]);
    }
    if (wsel == 0) IG_ACC(4, tm);
  } else if (warp >= 3) {
    // ===================================================================== epilogue (TMEM lanes 32*(warp%4)..+31)
    // TMEM -> registers (+ReLU, +bias) -> a padded per-warp staging tile -> global, so that every store
    // instruction writes whole 128-byte lines (4 rows x 32 channels) instead of 32 scattered 16-byte pieces.
    // two groups of four warps (one TMEM lane quadrant each); group eg drains the tiles of index eg, eg + 2, ...
    const int quad = warp & 3, eg = (warp - 3) >> 2;
    float* stg = reinterpret_cast<float*>(dyn_smem) + (stage_base - smem_u32(dyn_smem)) / 4 + (eg * 4 + quad) * 32 * IG_STAGE_LD;
    int it = 0;
    int pool_g[2] = {0, 0}, pool_py[2] = {0, 0}, pool_px[2] = {0, 0};
    if (POOL) {
      const int PX = p.out_x >> 1, wins = (p.out_y >> 1) * PX;
#pragma unroll
      for (int wi = 0; wi < 2; ++wi) {
        const int win = ((warp - 3) * 32 + lane + wi * 256) >> 2;
        pool_g[wi] = win / wins;
        const int rem = win - pool_g[wi] * wins;
        pool_py[wi] = rem / PX; pool_px[wi] = rem - pool_py[wi] * PX;
      }
    }
    IG_T(te);
    for (int st = blockIdx.x; st < p.nstages; st += gridDim.x, ++it) {
      const int as = it & 1, au = it >> 1;
      const int grp = st / p.nbands, band = st - grp * p.nbands;
      { IG_T(t0); mbar_wait(&t_full[as], (uint32_t)(au & 1)); if (warp == 3 && lane == 0) IG_ACC(5, t0); }
      tc_fence_after();
      if (POOL) {
        float* pt = reinterpret_cast<float*>(dyn_smem) + (stage_base - smem_u32(dyn_smem)) / 4;
        const int et = (warp - 3) * 32 + lane;                       // 0..255 among the epilogue threads
        const int PY = p.out_y >> 1, PX = p.out_x >> 1, wins = PY * PX, nitems = p.G * wins * 4;
#pragma unroll 1
        for (int cc = 0; cc < NB; cc += 16) {
          // bias of the pass in registers BEFORE the staging stores (a load between them would have to wait for them: the
          // compiler must assume the two shared arrays alias)
          float4 bb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) bb[j] = *reinterpret_cast<const float4*>(&sbias[cc + 4 * j]);
          for (int t = eg; t < p.T; t += 2) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + as * IG_ACCSET_COLS + t * NB + cc, v);
            tmem_ld_wait();
            float* d = pt + (size_t)(t * 128 + quad * 32 + lane) * IG_POOL_LD;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 r = make_float4(__uint_as_float(v[j]) + bb[j >> 2].x, __uint_as_float(v[j + 1]) + bb[j >> 2].y,
                                     __uint_as_float(v[j + 2]) + bb[j >> 2].z, __uint_as_float(v[j + 3]) + bb[j >> 2].w);
              if (p.pool_relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
              *reinterpret_cast<float4*>(d + j) = r;
            }
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          for (int w = et, wi = 0; w < nitems; w += 256, ++wi) {
            // the geometry of a thread's first two items does not change from stage to stage (pool_geo, set up once)
            int g, py, px;
            const int c4 = w & 3;
            if (wi < 2) { g = pool_g[wi]; py = pool_py[wi]; px = pool_px[wi]; }
            else {
              const int win = w >> 2;
              g = win / wins;
              const int rem = win - g * wins;
              py = rem / PX; px = rem - py * PX;
            }
            const int img = grp * p.G + g;
            if (img < p.nimg) {
              const float* s00 = pt + (size_t)(g * p.IR + (2 * py) * p.PW + 2 * px) * IG_POOL_LD + c4 * 4;
              const float4 a = *reinterpret_cast<const float4*>(s00), b = *reinterpret_cast<const float4*>(s00 + IG_POOL_LD);
              const float4 c = *reinterpret_cast<const float4*>(s00 + (size_t)p.PW * IG_POOL_LD);
              const float4 d = *reinterpret_cast<const float4*>(s00 + (size_t)(p.PW + 1) * IG_POOL_LD);
              float4 mx;
              mx.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)); mx.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
              mx.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)); mx.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
              uchar4 mk;       // bit k = dy * 2 + dx set where the element equals the window's max (multi-hot on ties)
              mk.x = (unsigned char)((a.x == mx.x) | ((b.x == mx.x) << 1) | ((c.x == mx.x) << 2) | ((d.x == mx.x) << 3));
              mk.y = (unsigned char)((a.y == mx.y) | ((b.y == mx.y) << 1) | ((c.y == mx.y) << 2) | ((d.y == mx.y) << 3));
              mk.z = (unsigned char)((a.z == mx.z) | ((b.z == mx.z) << 1) | ((c.z == mx.z) << 2) | ((d.z == mx.z) << 3));
              mk.w = (unsigned char)((a.w == mx.w) | ((b.w == mx.w) << 1) | ((c.w == mx.w) << 2) | ((d.w == mx.w) << 3));
              const size_t o = ((size_t)(img * PY + py) * PX + px) * p.N + n0 + cc + c4 * 4;
              *reinterpret_cast<float4*>(out + o) = mx;
              *reinterpret_cast<uchar4*>(p.pool_mask + o) = mk;
            }
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[as]);
        continue;
      }
      for (int t = eg; t < p.T; t += 2) {
        const int rloc = quad * 32 + lane;                              // row of the tile
        const int q = t * p.TS + rloc;
        const int g = q / p.IR, ql = q - g * p.IR;
        const int y = ql / p.PW, x = ql - y * p.PW;
        const int img = grp * p.G + g, yo = band * p.RO + y;
        const bool ok = rloc < p.TS && g < p.G && img < p.nimg && y < p.RO && yo < p.out_y && x < p.out_x;
        const int orow = (img * p.out_y + yo) * p.out_x + x;          // output pixel index (rows of N floats)
        const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
        if (FOLD) {
          // column blocks tx >= 1 go through the group's shared tile, shifted back by tx rows when they are read; the sums
          // return to the tile (slot 0, own row) for the coalesced copy-out.  All four warps of the group pass the same
          // named barriers (no early exit for a warp without valid rows).  16 columns per pass.
          float* fsh = reinterpret_cast<float*>(dyn_smem) + (stage_base - smem_u32(dyn_smem)) / 4 +
                       (size_t)eg * (p.fx - 1) * IG_FOLD_ROWS * IG_FOLD_LD;
          const uint32_t taddr_f = tmem + ((uint32_t)(quad * 32) << 16) + as * IG_ACCSET_COLS + (uint32_t)(t * NB * p.fx);
          const int bar_id = 1 + eg;
#pragma unroll 1
          for (int cc = 0; cc < NB; cc += 16) {
            uint32_t v0[16], v1[16];
            tmem_ld16(taddr_f + NB + cc, v0);
            if (p.fx > 2) tmem_ld16(taddr_f + 2 * NB + cc, v1);
            tmem_ld_wait();
            {
              float* d0 = fsh + (size_t)rloc * IG_FOLD_LD;
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(d0 + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
              if (p.fx > 2) {
                float* d1 = fsh + (size_t)(IG_FOLD_ROWS + rloc) * IG_FOLD_LD;
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(d1 + j) = make_uint4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
              }
            }
            for (int tx = 3; tx < p.fx; ++tx) {
              tmem_ld16(taddr_f + tx * NB + cc, v0);
              tmem_ld_wait();
              float* d = fsh + ((size_t)(tx - 1) * IG_FOLD_ROWS + rloc) * IG_FOLD_LD;
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(d + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
            }
            tmem_ld16(taddr_f + cc, v0);                      // the unshifted block meanwhile
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            tmem_ld_wait();
            float4 r4[4];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 r = make_float4(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1]), __uint_as_float(v0[j + 2]),
                                     __uint_as_float(v0[j + 3]));
              for (int tx = 1; tx < p.fx; ++tx) {
                const float4 a = *reinterpret_cast<const float4*>(fsh + ((size_t)(tx - 1) * IG_FOLD_ROWS + rloc + tx) * IG_FOLD_LD + j);
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
              }
              const float4 bb = *reinterpret_cast<const float4*>(&sbias[cc + j]);
              if (RELU) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
              r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
              r4[j >> 2] = r;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");       // every shifted read of this pass is done
            {
              float* d0 = fsh + (size_t)rloc * IG_FOLD_LD;
#pragma unroll
              for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(d0 + 4 * j) = r4[j];
            }
            __syncwarp();
            // copy-out: the warp's 32 rows x 16 columns, four lanes per row (64 contiguous bytes)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + (lane >> 2);
              const int dst_row = __shfl_sync(0xffffffffu, orow, r);
              const float4 val = *reinterpret_cast<const float4*>(fsh + (size_t)(quad * 32 + r) * IG_FOLD_LD + (lane & 3) * 4);
              if ((okmask >> r) & 1u)
                *reinterpret_cast<float4*>(out + (size_t)dst_row * p.N + n0 + cc + (lane & 3) * 4) = val;
            }
            __syncwarp();         // the warp's rows of slot 0 are rewritten by the next pass
          }
          continue;
        }
        if (okmask == 0u) continue;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + as * IG_ACCSET_COLS + t * NB;
#pragma unroll
        for (int cb = 0; cb < NB; cb += 32) {
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + cb + c0, v);
            float4 bb4[4];            // before the staging stores: a shared-memory load between them would wait for them
#pragma unroll
            for (int j = 0; j < 4; ++j) bb4[j] = *reinterpret_cast<const float4*>(&sbias[cb + c0 + 4 * j]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bb = bb4[j >> 2];
              float4 r = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (RELU) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
              r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
              *reinterpret_cast<float4*>(stg + lane * IG_STAGE_LD + c0 + j) = r;
            }
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const int dst_row = __shfl_sync(0xffffffffu, orow, r);
            const float4 val = *reinterpret_cast<const float4*>(stg + r * IG_STAGE_LD + (lane & 7) * 4);
            if ((okmask >> r) & 1u)
              *reinterpret_cast<float4*>(out + (size_t)dst_row * p.N + n0 + cb + (lane & 7) * 4) = val;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[as]);
    }
    if (warp == 3 && lane == 0) IG_ACC(6, te);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// --------------------------------------------------------------------------------------------- dF (+ db)
struct DfParams {
  int nstages, nbands, G, IR, PW, RO, Krows, x_rows, ntaps, fx, nimg, nsplit, e_rows_box;
  int ns;   // ring depth (stages in flight)
  int with_db;   // accumulate db through an all-ones operand (0: the caller has it from the kernel that produced E)
  int kimg;      // > 0: K steps per image, E holds only the oy output rows of every image (no zero rows between images)
                 // and the K loop restarts at every image of the stage (X rows of image g begin at g * IR); 0: one run of Krows
};

template <int MH>
__global__ void __launch_bounds__(DF_THREADS, 1)
conv_df_kernel(const __grid_constant__ CUtensorMap mapE, const __grid_constant__ CUtensorMap mapX, DfParams p,
               float* __restrict__ partial) {
  extern __shared__ int dyn_smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t e_half = (uint32_t)p.Krows * 128u, e_slot = MH * e_half, x_slot = (uint32_t)p.x_rows * 128u;
  const uint32_t x_base = base + (uint32_t)p.ns * e_slot, ones_base = x_base + (uint32_t)p.ns * x_slot;
  const uint32_t total_bytes = (uint32_t)p.ns * (e_slot + x_slot) + 1024u;
  const int split = blockIdx.x, ch = blockIdx.y;
  constexpr int DF_M = MH <= 2 ? 64 : 128;         // rows of the accumulator (filters)
  constexpr int DF_ROWS = MH * 32;                 // rows of a partial block
  pdl_launch_dependents();
  const int ntap_acc = p.ntaps + ((ch == 0 && p.with_db) ? 1 : 0);   // channel half 0 also accumulates db (all-ones operand)

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero everything once: the K-padding rows of E and the tail rows of X are never written by TMA and must stay 0
  for (uint32_t o = tid * 16u; o < total_bytes; o += DF_THREADS * 16u) {
    const uint32_t v = (base + o >= ones_base) ? 0x3f800000u : 0u;
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(v) : "memory");
  }
  pdl_wait();          // the prologue above touched only this CTA's shared memory / TMEM
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  int my_stages = 0;
  for (int st = split; st < p.nstages; st += p.nsplit) ++my_stages;

  if (warp == 0) {
    const bool leader = elect_one_sync();
    const uint32_t e_box_bytes = (uint32_t)(p.G * p.e_rows_box * p.PW) * 128u;
    const uint32_t x_box_bytes = (uint32_t)(p.G * p.IR) * 128u;
    // ring slot / use count and (image group, band) of the stage as running counters: no divisions in the loop
    int s = 0, u = 0;
    int grp = split / p.nbands, band = split - grp * p.nbands;
    const int dgrp = p.nsplit / p.nbands, dband = p.nsplit - dgrp * p.nbands;
    for (int st = split; st < p.nstages; st += p.nsplit) {
      if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
      if (leader) {
        mbar_expect_tx(&full[s], MH * e_box_bytes + x_box_bytes);
#pragma unroll
        for (int h = 0; h < MH; ++h)
          tma_load_4d(base + s * e_slot + h * e_half, &mapE, &full[s], h * 32, 0, band * p.RO, grp * p.G);
        tma_load_4d(x_base + s * x_slot, &mapX, &full[s], ch * 32, 0, band * p.RO, grp * p.G);
      }
      if (++s == p.ns) { s = 0; ++u; }
      grp += dgrp; band += dband;
      if (band >= p.nbands) { band -= p.nbands; ++grp; }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    // One MMA per filter ROW: the fx taps of a row are fx 32-channel N groups of the SAME operand, one pixel
    // (128 B = LBO) apart, landing in fx consecutive 32-column accumulators.
    // nf <= 64: M = 64 MMAs (half the shared-memory read of the E operand per MMA, no junk filter groups)
    const int fy = p.ntaps / p.fx;
    const uint32_t idesc_row = idesc_tf32(DF_M, 32 * p.fx, true, true);
    constexpr uint32_t idesc_one = idesc_tf32(DF_M, 32, true, true);
    const uint64_t ones_desc = desc_mn_sw128_32b(ones_base, 4096u, 512u);
    const uint32_t row_step = (uint32_t)(p.PW * 128) >> 4;   // descriptor units (16 B) per filter row
    // K runs: one of Krows / 8 steps, or (kimg > 0) one of kimg steps per image with the X operand restarting at the image
    const int nrun = p.kimg > 0 ? p.G : 1, ksteps = p.kimg > 0 ? p.kimg : p.Krows / 8;
    const uint32_t x_run = (uint32_t)(p.IR * 128) >> 4;      // descriptor units between the X rows of consecutive images
    const bool with_db = ch == 0 && p.with_db;
    const uint64_t ad0 = desc_mn_sw128_32b(base, e_half, 512u), bd0 = desc_mn_sw128_32b(x_base, 128u, 512u);
    const uint32_t e_slot_u = e_slot >> 4, x_slot_u = x_slot >> 4;
    const uint32_t acc_step = (uint32_t)(p.fx * 32);
    int s = 0, u = 0;
    uint32_t acc_flag = 0u;
    for (int st = split; st < p.nstages; st += p.nsplit) {
      mbar_wait(&full[s], (uint32_t)(u & 1));
      tc_fence_after();
      if (leader) {
        uint64_t ad = ad0 + (uint64_t)((uint32_t)s * e_slot_u);
        const uint64_t bds = bd0 + (uint64_t)((uint32_t)s * x_slot_u);
        for (int g = 0; g < nrun; ++g) {
          uint64_t bd = bds + (uint64_t)((uint32_t)g * x_run);
          if (fy == 3 && !with_db) {
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ++ks, ad += 64, bd += 64) {   // 8 pixel rows = 1024 B per K step
              umma_tf32(tmem, ad, bd, idesc_row, acc_flag);
              umma_tf32(tmem + acc_step, ad, bd + (uint64_t)row_step, idesc_row, acc_flag);
              umma_tf32(tmem + 2 * acc_step, ad, bd + (uint64_t)(2 * row_step), idesc_row, acc_flag);
              acc_flag = 1u;
            }
          } else {
            for (int ks = 0; ks < ksteps; ++ks, ad += 64, bd += 64) {
              uint64_t bdr = bd;
              for (int ty = 0; ty < fy; ++ty, bdr += row_step)
                umma_tf32(tmem + ty * acc_step, ad, bdr, idesc_row, acc_flag);
              if (with_db) umma_tf32(tmem + p.ntaps * 32, ad, ones_desc, idesc_one, acc_flag);
              acc_flag = 1u;
            }
          }
        }
        umma_commit(&empty[s]);
      }
      if (++s == p.ns) { s = 0; ++u; }
    }
    if (leader) umma_commit(&done);
  } else if (warp >= 2) {
    const int quad = warp & 3;
    // M = 128: row m of the accumulator is TMEM lane m; M = 64: row r sits in lane (r / 16) * 32 + r % 16
    const int m = DF_M == 128 ? quad * 32 + lane : (lane < 16 ? quad * 16 + lane : DF_ROWS);
    if (my_stages > 0) {
      mbar_wait(&done, 0);
      tc_fence_after();
    }
    // partial[split][ch][tap][m][32], m < DF_ROWS
    float* dst = partial + (((size_t)split * gridDim.y + ch) * (p.ntaps + 1)) * DF_ROWS * 32 + (size_t)m * 32;
    for (int tap = 0; tap < p.ntaps + 1; ++tap) {
      uint32_t v[32];
      if (my_stages > 0 && tap < ntap_acc) {
        tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + tap * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < DF_ROWS) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + (size_t)tap * DF_ROWS * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// dF[f][c][tap] = sum_z partial[z][c/32][tap][f][c%32];  db[f] = sum_z partial[z][0][ntaps][f][0]
// Threads walk the PARTIAL's own order (32 channels innermost), so every split is read with full 128-byte lines; the
// (small) result is scattered into the reference's (nf, ic, fy, fx) order.  block = 32 outputs x 16 slices of the split
// range: a thread adds ~9 partials instead of ~148 (the kernel is latency-, not bandwidth-bound), then a fixed-order
// shared-memory tree adds the slices (deterministic).
constexpr int DF_RED_SLICES = 16;
__global__ void __launch_bounds__(32 * DF_RED_SLICES)
conv_df_reduce_nhwc_kernel(const float* __restrict__ partial, float* __restrict__ dF, float* __restrict__ db,
                           int nsplit, int nch, int ntaps, int nf, int ic, const float* __restrict__ dbp, int dbrows) {
  // thread = FOUR consecutive channels (one 16-byte load per split) of one (channel group, tap, filter) row x one slice of
  // the split range; the last blocks take the nf bias sums (scalar)
  __shared__ float4 red[DF_RED_SLICES][32];
  pdl_launch_dependents();
  pdl_wait();
  const long nmain4 = (long)nch * ntaps * nf * 8;               // float4 items of the filter gradient
  const long main_blocks = (nmain4 + 31) / 32;
  const size_t zstride = (size_t)nch * (ntaps + 1) * nf * 32;
  const int zs = threadIdx.y;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  long dst = -1;
  int dbf = -1;
  if ((long)blockIdx.x < main_blocks) {
    const long i = (long)blockIdx.x * 32 + threadIdx.x;
    if (i < nmain4) {
      const int c4 = (int)(i & 7);
      long r = i >> 3;
      const int f = (int)(r % nf);
      r /= nf;
      const int tap = (int)(r % ntaps), ch = (int)(r / ntaps);
      const float4* src = reinterpret_cast<const float4*>(partial + ((size_t)(ch * (ntaps + 1) + tap) * nf + f) * 32) + c4;
      const size_t z4 = zstride / 4;
      float4 s1 = s, s2 = s, s3 = s;            // independent chains: four splits' loads in flight together
      int z = zs;
      for (; z + 3 * DF_RED_SLICES < nsplit; z += 4 * DF_RED_SLICES) {
        const float4 a0 = src[(size_t)z * z4], a1 = src[(size_t)(z + DF_RED_SLICES) * z4];
        const float4 a2 = src[(size_t)(z + 2 * DF_RED_SLICES) * z4], a3 = src[(size_t)(z + 3 * DF_RED_SLICES) * z4];
        s.x += a0.x; s.y += a0.y; s.z += a0.z; s.w += a0.w;
        s1.x += a1.x; s1.y += a1.y; s1.z += a1.z; s1.w += a1.w;
        s2.x += a2.x; s2.y += a2.y; s2.z += a2.z; s2.w += a2.w;
        s3.x += a3.x; s3.y += a3.y; s3.z += a3.z; s3.w += a3.w;
      }
      for (; z < nsplit; z += DF_RED_SLICES) {
        const float4 a0 = src[(size_t)z * z4];
        s.x += a0.x; s.y += a0.y; s.z += a0.z; s.w += a0.w;
      }
      s.x = (s.x + s1.x) + (s2.x + s3.x); s.y = (s.y + s1.y) + (s2.y + s3.y);
      s.z = (s.z + s1.z) + (s2.z + s3.z); s.w = (s.w + s1.w) + (s2.w + s3.w);
      dst = ((long)f * ic + ch * 32 + c4 * 4) * ntaps + tap;
    }
  } else {
    const int f = (int)(((long)blockIdx.x - main_blocks) * 32 + threadIdx.x);
    if (f < nf) {
      dbf = f;
      if (dbp) {                                  // db from the per-CTA channel sums of the kernel that wrote E (bf_pool_nhwc_backward_db)
        for (int z = zs; z < dbrows; z += DF_RED_SLICES) s.x += dbp[(size_t)z * nf + f];
      } else {
        const size_t off = ((size_t)ntaps * nf + f) * 32;
        for (int z = zs; z < nsplit; z += DF_RED_SLICES) s.x += partial[(size_t)z * zstride + off];
      }
    }
  }
  red[zs][threadIdx.x] = s;
  __syncthreads();
#pragma unroll
  for (int w = DF_RED_SLICES / 2; w >= 1; w >>= 1) {
    if (zs < w) {
      float4 a = red[zs][threadIdx.x];
      const float4 c = red[zs + w][threadIdx.x];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      red[zs][threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (zs == 0) {
    const float4 sum = red[0][threadIdx.x];
    if (dst >= 0) { dF[dst] = sum.x; dF[dst + ntaps] = sum.y; dF[dst + 2 * ntaps] = sum.z; dF[dst + 3 * ntaps] = sum.w; }
    else if (dbf >= 0 && db) db[dbf] = sum.x;
  }
}

// Wt[tap][n][k]:  forward  n = f, k = c, tap = (ky,kx)        -> F[f][c][ky][kx]
//                 data grad n = c, k = f, tap = (ty,tx)        -> F[f][c][fy-1-ty][fx-1-tx]   (flip + transpose)
__global__ void conv_repack_nhwc_kernel(const float* __restrict__ F, float* __restrict__ Wt, int nf, int ic, int fy, int fx,
                                        int transpose) {
  pdl_launch_dependents();
  pdl_wait();
  const int ntaps = fy * fx, N = transpose ? ic : nf, K = transpose ? nf : ic;
  const long total = (long)ntaps * N * K;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int k = (int)(i % K);
    long r = i / K;
    const int n = (int)(r % N), tap = (int)(r / N);
    const int ty = tap / fx, tx = tap - ty * fx;
    float v;
    if (transpose) v = F[(((long)k * ic + n) * fy + (fy - 1 - ty)) * fx + (fx - 1 - tx)];
    else v = F[(((long)n * ic + k) * fy + ty) * fx + tx];
    Wt[i] = v;
  }
}

// both forms in one pass over F: the forward operand and the flipped / transposed data-gradient operand
__global__ void conv_repack_both_kernel(const float* __restrict__ F, float* __restrict__ Wf, float* __restrict__ Wd, int nf, int ic,
                                        int fy, int fx) {
  pdl_launch_dependents();
  pdl_wait();
  const int ntaps = fy * fx;
  const long total = (long)ntaps * nf * ic;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    {   // forward: Wf[tap][n = f][k = c]
      const int k = (int)(i % ic);
      long r = i / ic;
      const int n = (int)(r % nf), tap = (int)(r / nf);
      const int ty = tap / fx, tx = tap - ty * fx;
      Wf[i] = F[(((long)n * ic + k) * fy + ty) * fx + tx];
    }
    {   // data gradient: Wd[tap][n = c][k = f] = F[f][c][fy-1-ty][fx-1-tx]
      const int k = (int)(i % nf);
      long r = i / nf;
      const int n = (int)(r % ic), tap = (int)(r / ic);
      const int ty = tap / fx, tx = tap - ty * fx;
      Wd[i] = F[(((long)k * ic + n) * fy + (fy - 1 - ty)) * fx + (fx - 1 - tx)];
    }
  }
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
// dF kernel: [E slots | X slots | ones]; the A operand always spans four 32-filter groups (M = 128), so for
// nf < 128 the junk groups read behind the last E slot must still fall inside the allocation
static inline size_t df_smem_bytes(int MH, int Krows, int x_rows, int ns) {
  size_t rows = (size_t)ns * ((size_t)MH * Krows + x_rows) + 8;
  const size_t reach = (size_t)(ns - 1) * MH * Krows + 4 * (size_t)Krows;
  if (rows < reach) rows = reach;
  return rows * 128 + 1024;
}
constexpr size_t IG_SMEM_BUDGET = 224 * 1024;

struct IgemmPlan {
  IgemmParams p;
  int PH_box;
  size_t smem;
  bool ok;
  double score;
};

// Picks images-per-stage or output-row bands so that the accumulators fit one TMEM set and two A slots + the
// filter (resident if it fits, else a ring of taps) + the epilogue staging fit shared memory.  With `allow_fold` the
// column-folded form (see the kernel) competes as well, where fx * NB accumulator columns and a resident filter fit.
static IgemmPlan plan_igemm(int B, int H, int W, int KH, int nkp, int NB, int N, int fy, int fx, int pad_y, int pad_x,
                            int relu, bool allow_fold, bool* folded, bool pool = false) {
  IgemmPlan pl;
  pl.ok = false;
  pl.score = -1;
  *folded = false;
  const int PW = W + 2 * pad_x, PH = H + 2 * pad_y;
  const int out_y = PH - fy + 1, out_x = PW - fx + 1;
  if (out_y <= 0 || out_x <= 0 || PW > 256) return pl;
  const int ntaps = fy * fx;
  const size_t b_slot = (size_t)KH * NB * 128;
  double best = -1;
  // cost model (cycles per stage on one SM): the MMAs are bound by shared-memory operand reads for small N
  // ((128 + N) rows x 32 B per K=8 step at 128 B/clk) or by the tensor pipe (N/2 clk per step); the loads by the
  // SM's share of HBM/L2 bandwidth (~24 B/clk); a shallow ring leaves the load latency exposed.
  auto consider = [&](int G, int nbands, int RO, int PHb, bool fold) {
    const int IR = PHb * PW;
    const int NBF = fold ? NB * fx : NB, TS = fold ? 128 - (fx - 1) : 128;
    const int Tmax = IG_ACCSET_COLS / NBF;
    const int max_shift = fold ? (fy - 1) * PW : (fy - 1) * PW + fx - 1;
    const long last = (long)(G - 1) * IR + (long)(RO - 1) * PW + out_x;   // positions to cover
    const int T = (int)((last + TS - 1) / TS);
    if (T > Tmax || T < 1) return false;
    if (pool && nbands != 1) return true;            // pooling windows must not straddle stages: whole images only
    // epilogue scratch: per-warp staging tiles, (folded) one shifted-read tile per epilogue group, or (pooling) the stage's positions
    const size_t fixed = 2048 + (pool ? (size_t)T * 128 * IG_POOL_LD * 4
                                      : fold ? (size_t)2 * (fx - 1) * IG_FOLD_ROWS * IG_FOLD_LD * 4 : (size_t)IG_STAGING_BYTES);
    long rows = (long)(T - 1) * TS + 128 + max_shift;
    if (rows < (long)G * IR) rows = (long)G * IR;
    rows = (rows + 7) / 8 * 8;
    const size_t a_slot = (size_t)KH * rows * 128;
    if (2 * a_slot + 2 * b_slot + fixed > IG_SMEM_BUDGET) return false;
    const size_t b_all = (size_t)ntaps * nkp * b_slot;
    const bool resident = 2 * a_slot + b_all + fixed <= IG_SMEM_BUDGET;
    if (fold && !resident) return false;
    int na, nbs;
    if (resident) {
      nbs = ntaps * nkp;
      na = (int)((IG_SMEM_BUDGET - fixed - b_all) / a_slot);
    } else {
      na = 4;
      while (na > 2 && (IG_SMEM_BUDGET - fixed - na * a_slot) / b_slot < (size_t)(ntaps < 4 ? ntaps : 4)) --na;
      if (na * a_slot + 2 * b_slot + fixed > IG_SMEM_BUDGET) na = 2;
      nbs = (int)((IG_SMEM_BUDGET - fixed - na * a_slot) / b_slot);
      if (nbs > 16) nbs = 16;
    }
    if (na > 8) na = 8;
    // per K=8 step a 128-row tile costs max(tensor N/2, shared-memory operand reads (128+N)/4) cycles, but one
    // issuing thread sustains only ~1 MMA per 100 cycles (two issuers when T >= 2)
    const double mma_step = NBF / 2.0 > (128 + NBF) / 4.0 ? NBF / 2.0 : (128 + NBF) / 4.0;
    const double issue = T >= 2 ? 60.0 : 120.0;
    const double mma_clk = (double)(fold ? fy : ntaps) * nkp * T * KH * 4 * (mma_step > issue ? mma_step : issue);
    const double bytes = (double)KH * nkp * G * IR * 128 + (resident ? 0.0 : (double)b_all);
    double stage_clk = mma_clk > bytes / 24.0 ? mma_clk : bytes / 24.0;
    stage_clk += 600.0 + (na == 2 ? 1500.0 : 0.0) + (!resident && nbs < 4 ? 1500.0 : 0.0);
    const double score = (double)G * RO * out_x / stage_clk;
    if (score > best) {
      best = score;
      pl.p = IgemmParams{nbands == 1 ? (B + G - 1) / G : B * nbands, nbands, G, IR, PW, T, (int)rows, ntaps, fx, -pad_x, -pad_y,
                         RO, out_y, out_x, B, N, relu, nbs, resident ? 1 : 0, nkp, na, fy, TS, 0, nullptr, nullptr};
      pl.PH_box = PHb;
      pl.smem = 1024 + (size_t)na * a_slot + (size_t)nbs * b_slot + fixed - 2048;
      pl.ok = true;
      pl.score = score;
      *folded = fold;
    }
    return true;
  };
  static const bool fold_off = getenv("BF_IGEMM_NOFOLD") != nullptr;
  for (int fold = 0; fold <= ((allow_fold && !fold_off && fx >= 2 && fx <= 4 && NB * fx <= IG_ACCSET_COLS) ? 1 : 0); ++fold) {
    if (PH <= 256)
      for (int G = 1; G <= 256 && G <= B + 8; ++G)
        if (!consider(G, 1, out_y, PH, fold != 0)) break;
    // "full" correlation (data gradient, pad = f - 1): the bottom padding rows of an image need not be staged -- the reads
    // that would touch them land in the TOP padding rows of the next image of the stage (zero-filled by TMA) or, behind the
    // last image, in the never-written tail of the slot (zeros).  Every staged row is then an output row: on a 4x4 plane
    // 48 instead of 64 positions per image.
    if (pad_y > 0 && pad_y == fy - 1 && PH - pad_y <= 256)
      for (int G = 1; G <= 256 && G <= B + 8; ++G)
        if (!consider(G, 1, out_y, PH - pad_y, fold != 0)) break;
    for (int RO = (out_y < 254 ? out_y - 1 : 254); RO >= 1; --RO) {
      if (RO + fy - 1 > 256) continue;
      consider(1, (out_y + RO - 1) / RO, RO, RO + fy - 1, fold != 0);
    }
  }
  return pl;
}

template <int NB, int KH>
static int launch_igemm(const CUtensorMap& mA, const CUtensorMap& mB, IgemmPlan& pl, float* out, const float* bias, bool fold,
                        bool pool, cudaStream_t s) {
  static long long* d_prof = nullptr;
  const bool prof = getenv("BF_PROF_IGEMM") != nullptr;
  if (prof) {
    if (!d_prof) cudaMalloc(&d_prof, 16 * sizeof(long long));
    cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), s);
    pl.p.prof = d_prof;
  }
  const int ctas = pl.p.nstages < num_sms() ? pl.p.nstages : num_sms();
  dim3 grid(ctas, pl.p.N / NB);
#define BF_IG_LAUNCH(RELU_, FOLD_)                                                                                                          \
  {                                                                                                                                        \
    BF_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<NB, KH, RELU_, FOLD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));      \
    launch_pdl(conv_igemm_kernel<NB, KH, RELU_, FOLD_>, dim3(grid), dim3(IG_THREADS), pl.smem, s, mA, mB, pl.p, out, bias);                 \
  }
  if (pool) {
    BF_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<NB, KH, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    launch_pdl(conv_igemm_kernel<NB, KH, false, false, true>, dim3(grid), dim3(IG_THREADS), pl.smem, s, mA, mB, pl.p, out, bias);
  } else if (fold) {
    if constexpr (NB <= 64) {
      if (pl.p.relu) BF_IG_LAUNCH(true, true) else BF_IG_LAUNCH(false, true)
    }
  } else {
    if (pl.p.relu) BF_IG_LAUNCH(true, false) else BF_IG_LAUNCH(false, false)
  }
#undef BF_IG_LAUNCH
  BF_LAUNCH_CHECK();
  if (prof) {
    long long h[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
    const int mine = (pl.p.nstages + ctas - 1) / ctas;
    fprintf(stderr, "[igemm prof, CTA 0, %d stages] producer: wait a_empty %lld, total %lld | mma: wait t_empty %lld, wait a_full %lld, "
            "total %lld | epilogue: wait t_full %lld, total %lld  (cycles)\n", mine, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  }
  return BF_OK;
}

static inline int pick_nb(int N) {
  if (N == 32 || N == 64 || N == 128) return N;
  if (N > 128 && N % 128 == 0) return 128;
  return 0;
}

bool conv_nhwc_supported(int C, int N, int fy, int fx) {
  if (!tma_encode_fn()) return false;
  if (C % 32 != 0 || C <= 0 || C > 1024) return false;
  if (pick_nb(N) == 0) return false;
  if (fy * fx > 15 || fy < 1 || fx < 1) return false;
  return true;
}

// out[(b,y,x), n] = act( sum_{ty,tx,c} in[b, y - pad_y + ty, x - pad_x + tx, c] * Wt[(ty,tx)][n][c] ) + bias[n]
static int conv_nhwc_igemm(const float* in, const float* Wt, const float* bias, float* out, int B, int H, int W, int C,
                           int N, int fy, int fx, int pad_y, int pad_x, int relu, bool allow_fold, cudaStream_t s,
                           uint8_t* pool_mask = nullptr, int pool_relu = 0) {
  int KH = (C % 64 == 0) ? 2 : 1, nkp = C / (32 * KH);
  const int NB = pick_nb(N);
  bool fold = false;
  const bool pool = pool_mask != nullptr;
  IgemmPlan pl = plan_igemm(B, H, W, KH, nkp, NB, N, fy, fx, pad_y, pad_x, relu, allow_fold && !pool, &fold, pool);
  if (pool && pl.ok) { pl.p.pool_mask = pool_mask; pl.p.pool_relu = pool_relu; }
  if (allow_fold && !pool && KH == 2) {
    // 64 channels per stage may leave no room for the resident filter + the folded epilogue's tile next to whole images:
    // 32 at a time (two K passes) competes
    bool fold1 = false;
    IgemmPlan pl1 = plan_igemm(B, H, W, 1, C / 32, NB, N, fy, fx, pad_y, pad_x, relu, true, &fold1);
    if (pl1.ok && fold1 && (!pl.ok || pl1.score > pl.score)) { pl = pl1; fold = true; KH = 1; nkp = C / 32; }
  }
  if (!pl.ok) return fail(BF_ERR_ARG, "conv_nhwc: no tiling for plane %dx%d, %d->%d channels", H, W, C, N);
  if ((long)B * pl.p.out_y * pl.p.out_x >= 2147483647L) return fail(BF_ERR_ARG, "conv_nhwc: too many output pixels");
  if (getenv("BF_DEBUG_PLAN"))
    fprintf(stderr, "[igemm %dx%dx%d->%d pad %d] G=%d bands=%d RO=%d T=%d rows=%d KH=%d nkp=%d NB=%d na=%d nbs=%d resident=%d fold=%d smem=%zu stages=%d\n",
            H, W, C, N, pad_y, pl.p.G, pl.p.nbands, pl.p.RO, pl.p.T, pl.p.slot_rows, KH, nkp, NB, pl.p.na, pl.p.nbs, pl.p.resident, fold ? 1 : 0,
            pl.smem, pl.p.nstages);
  CUtensorMap mA, mB;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C, (uint64_t)C * W, (uint64_t)C * W * H};
    uint32_t box[4] = {32, (uint32_t)pl.p.PW, (uint32_t)pl.PH_box, (uint32_t)pl.p.G};
    if (!tma_make_map(&mA, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return fail(BF_ERR_CUDA, "conv_nhwc: cuTensorMapEncodeTiled failed (activations %dx%dx%dx%d)", B, H, W, C);
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)(fy * fx)};
    uint64_t str[2] = {(uint64_t)C, (uint64_t)C * N};
    uint32_t box[3] = {32, (uint32_t)NB, (uint32_t)(fold ? fx : 1)};
    if (!tma_make_map(&mB, Wt, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return fail(BF_ERR_CUDA, "conv_nhwc: cuTensorMapEncodeTiled failed (filter)");
  }
#define BF_IG_CASE(NB_, KH_) if (NB == NB_ && KH == KH_) return launch_igemm<NB_, KH_>(mA, mB, pl, out, bias, fold, pool, s);
  BF_IG_CASE(32, 1) BF_IG_CASE(32, 2) BF_IG_CASE(64, 1) BF_IG_CASE(64, 2) BF_IG_CASE(128, 1) BF_IG_CASE(128, 2)
#undef BF_IG_CASE
  return fail(BF_ERR_ARG, "conv_nhwc: unsupported channel counts %d -> %d", C, N);
}

}  // namespace bf

using namespace bf;

// Y (im,oy,ox,nf) = act(valid_corr(X (im,iy,ix,ic), F (nf,ic,fy,fx))) + bias      [NHWC activations]
int bf_conv_nhwc_forward(const float* X, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix, int nf,
                         int fy, int fx, int act, void* stream) {
  BF_REQUIRE(conv_nhwc_supported(ic, nf, fy, fx), BF_ERR_ARG, "bf_conv_nhwc_forward: unsupported shape (ic=%d nf=%d)", ic, nf);
  BF_REQUIRE(act == BF_ACT_LINEAR || act == BF_ACT_RELU, BF_ERR_ARG,
             "bf_conv_nhwc_forward: only linear / relu are fused into the epilogue (got act %d)", act);
  if (im <= 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  void* ws = nullptr;
  int rc = workspace_require(al256((size_t)nf * ic * fy * fx * 4), &ws);
  if (rc) return rc;
  launch_pdl(conv_repack_nhwc_kernel, dim3(ew_grid((size_t)nf * ic * fy * fx, 256)), dim3(256), 0, s, F, (float*)ws, nf, ic, fy, fx, 0);
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return conv_nhwc_igemm(X, (const float*)ws, bias, Y, im, iy, ix, ic, nf, fy, fx, 0, 0, act == BF_ACT_RELU ? 1 : 0,
                         getenv("BF_IGEMM_FOLD_FWD") != nullptr, s);
}

// dX (im,iy,ix,ic) = full_corr(E (im,oy,ox,nf), flip(F)^T)                           [NHWC]
int bf_conv_nhwc_backward_data(const float* E, const float* F, float* dX, int im, int ic, int iy, int ix, int nf, int fy,
                               int fx, void* stream) {
  BF_REQUIRE(conv_nhwc_supported(nf, ic, fy, fx), BF_ERR_ARG, "bf_conv_nhwc_backward_data: unsupported shape (ic=%d nf=%d)", ic, nf);
  if (im <= 0) return BF_OK;
  cudaStream_t s = as_stream(stream);
  const int oy = iy - fy + 1, ox = ix - fx + 1;
  void* ws = nullptr;
  int rc = workspace_require(al256((size_t)nf * ic * fy * fx * 4), &ws);
  if (rc) return rc;
  launch_pdl(conv_repack_nhwc_kernel, dim3(ew_grid((size_t)nf * ic * fy * fx, 256)), dim3(256), 0, s, F, (float*)ws, nf, ic, fy, fx, 1);
  BF_LAUNCH_CHECK();
  set_last_path(2);
  // few output channels (N = ic): the column-folded form, when its accumulators and the resident filter fit
  return conv_nhwc_igemm(E, (const float*)ws, nullptr, dX, im, oy, ox, nf, ic, fy, fx, fy - 1, fx - 1, 0, true, s);
}

// dF (nf,ic,fy,fx), db (nf) from X (im,iy,ix,ic), E (im,oy,ox,nf)                      [NHWC activations]
static int conv_nhwc_backward_filter_impl(const float* X, const float* E, float* dF, float* db, const float* db_partial,
                                          int db_rows, int im, int ic, int iy, int ix, int nf, int fy, int fx, void* stream) {
  BF_REQUIRE(ic % 32 == 0 && (nf == 32 || nf == 64 || nf == 128) && fy * fx <= 15 && fx <= 8 && tma_encode_fn() != nullptr, BF_ERR_ARG,
             "bf_conv_nhwc_backward_filter: unsupported shape (ic=%d nf=%d)", ic, nf);
  cudaStream_t s = as_stream(stream);
  const int oy = iy - fy + 1, ox = ix - fx + 1, ntaps = fy * fx, MH = nf / 32, nch = ic / 32;
  BF_REQUIRE(oy > 0 && ox > 0 && ix <= 256, BF_ERR_ARG, "bf_conv_nhwc_backward_filter: bad plane %dx%d", iy, ix);
  if (im <= 0) {
    BF_CUDA(cudaMemsetAsync(dF, 0, (size_t)nf * ic * ntaps * 4, s));
    if (db) BF_CUDA(cudaMemsetAsync(db, 0, (size_t)nf * 4, s));
    return BF_OK;
  }
  const int max_shift = (fy - 1) * ix + fx - 1;
  DfParams p;
  bool ok = false;
  size_t smem = 0;
  double best = -1;
  // candidates: G whole images per stage (E loaded with iy rows so that both operands share the row index) or
  // bands of RO filter-gradient rows of one image; score = useful pixels per estimated stage time, where a
  // ring of fewer than 4 stages leaves the load latency exposed
  auto consider = [&](int G, int nbands, int RO, int e_rows, int xr, bool per_image = false) {
    const int IRx = xr * ix;
    // per_image: E carries only the oy * ix rows of every image (a multiple of 8), the K loop restarts per image
    const int Krows = per_image ? G * oy * ix : (G * (nbands == 1 ? IRx : RO * ix) + 7) / 8 * 8;
    int x_rows = per_image ? (G - 1) * IRx + oy * ix + max_shift : Krows + max_shift;
    if (x_rows < G * IRx) x_rows = G * IRx;
    x_rows = (x_rows + 7) / 8 * 8;
    int ns = 8;
    while (ns >= 2 && df_smem_bytes(MH, Krows, x_rows, ns) > IG_SMEM_BUDGET) --ns;
    if (ns < 2) return false;
    // one MMA per filter row and K step: max(shared-memory read of both operands at 128 B/clk, tensor floor 128 x N / 256)
    const double m_rows = MH <= 2 ? 64.0 : 128.0, n_cols = 32.0 * fx;
    const double one_mma = (m_rows + n_cols) / 4.0 > n_cols / 2.0 ? (m_rows + n_cols) / 4.0 : n_cols / 2.0;
    const double mma_clk = (Krows / 8) * (fy * one_mma + 40.0);
    const double bytes = (double)G * (MH * e_rows + xr) * ix * 128.0;
    double stage_clk = (mma_clk > bytes / 24.0 ? mma_clk : bytes / 24.0) + 400.0 + (ns == 2 ? 3000.0 : ns == 3 ? 1000.0 : 0.0);
    // useful filter-gradient pixels per stage: the last band of an image may be partly empty
    const double score = (double)G * ((double)oy / nbands) * ox / stage_clk;
    if (score > best) {
      best = score;
      p = DfParams{nbands == 1 ? (im + G - 1) / G : im * nbands, nbands, G, IRx, ix, RO, Krows, x_rows, ntaps, fx, im, 1, e_rows, ns,
                   db_partial ? 0 : 1, per_image ? oy * ix / 8 : 0};
      smem = df_smem_bytes(MH, Krows, x_rows, ns);
      ok = true;
    }
    return true;
  };
  if (iy <= 256)
    for (int G = 1; G <= 256 && G <= im + 8; ++G)
      if (!consider(G, 1, oy, iy, iy)) break;
  if (iy <= 256 && (oy * ix) % 8 == 0 && oy < iy)
    for (int G = 1; G <= 256 && G <= im + 8; ++G)
      if (!consider(G, 1, oy, oy, iy, true)) break;
  for (int RO = (oy < 254 ? oy - 1 : 254); RO >= 1; --RO) {
    if (RO + fy - 1 > 256) continue;
    consider(1, (oy + RO - 1) / RO, RO, RO, RO + fy - 1);
  }
  BF_REQUIRE(ok, BF_ERR_ARG, "bf_conv_nhwc_backward_filter: no tiling for plane %dx%d", iy, ix);
  if (getenv("BF_DEBUG_PLAN"))
    fprintf(stderr, "[df %dx%dx%d->%d] G=%d bands=%d RO=%d Krows=%d x_rows=%d ns=%d kimg=%d smem=%zu stages=%d\n", iy, ix, ic, nf, p.G,
            p.nbands, p.RO, p.Krows, p.x_rows, p.ns, p.kimg, smem, p.nstages);
  int nsplit = num_sms() / nch;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > p.nstages) nsplit = p.nstages;
  // small batches: every split costs a 160 KB partial that the reduction has to read back -- give a CTA at least four
  // stages of work before adding another split
  {
    static const int min_stages = getenv("BF_DF_MIN_STAGES") ? atoi(getenv("BF_DF_MIN_STAGES")) : 4;
    const int cap = (p.nstages + min_stages - 1) / min_stages;
    if (nsplit > cap) nsplit = cap < 1 ? 1 : cap;
  }
  p.nsplit = nsplit;
  const size_t part_bytes = al256((size_t)nsplit * nch * (ntaps + 1) * nf * 32 * 4);
  void* ws = nullptr;
  int rc = workspace_require(part_bytes, &ws);
  if (rc) return rc;
  CUtensorMap mE, mX;
  {
    uint64_t dims[4] = {(uint64_t)nf, (uint64_t)ox, (uint64_t)oy, (uint64_t)im};
    uint64_t str[3] = {(uint64_t)nf, (uint64_t)nf * ox, (uint64_t)nf * ox * oy};
    uint32_t box[4] = {32, (uint32_t)ix, (uint32_t)p.e_rows_box, (uint32_t)p.G};
    if (!tma_make_map(&mE, E, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return fail(BF_ERR_CUDA, "bf_conv_nhwc_backward_filter: cuTensorMapEncodeTiled failed (E)");
  }
  {
    uint64_t dims[4] = {(uint64_t)ic, (uint64_t)ix, (uint64_t)iy, (uint64_t)im};
    uint64_t str[3] = {(uint64_t)ic, (uint64_t)ic * ix, (uint64_t)ic * ix * iy};
    uint32_t box[4] = {32, (uint32_t)ix, (uint32_t)(p.IR / ix), (uint32_t)p.G};
    if (!tma_make_map(&mX, X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return fail(BF_ERR_CUDA, "bf_conv_nhwc_backward_filter: cuTensorMapEncodeTiled failed (X)");
  }
  dim3 grid(nsplit, nch);
#define BF_DF_CASE(MH_)                                                                                              \
  if (MH == MH_) {                                                                                                   \
    BF_CUDA(cudaFuncSetAttribute(conv_df_kernel<MH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    launch_pdl(conv_df_kernel<MH_>, dim3(grid), dim3(DF_THREADS), smem, s, mE, mX, p, (float*)ws);                                       \
  }
  BF_DF_CASE(1) BF_DF_CASE(2) BF_DF_CASE(4)
#undef BF_DF_CASE
  BF_LAUNCH_CHECK();
  launch_pdl(conv_df_reduce_nhwc_kernel, dim3((unsigned)(((size_t)nf * ic * ntaps / 4 + 31) / 32 + (nf + 31) / 32)), dim3(32, DF_RED_SLICES), 0, s,
             (const float*)ws, dF, db, nsplit, nch, ntaps, nf, ic, db_partial, db_rows);
  BF_LAUNCH_CHECK();
  set_last_path(2);
  return BF_OK;
}

int bf_conv_nhwc_backward_filter(const float* X, const float* E, float* dF, float* db, int im, int ic, int iy, int ix,
                                 int nf, int fy, int fx, void* stream) {
  return conv_nhwc_backward_filter_impl(X, E, dF, db, nullptr, 0, im, ic, iy, ix, nf, fy, fx, stream);
}

// As above with db taken from `db_rows` rows of per-CTA channel sums of E (written by bf_pool_nhwc_backward_db, the kernel
// that produced E): no all-ones MMA per K step.
int bf_conv_nhwc_backward_filter_dbp(const float* X, const float* E, float* dF, float* db, const float* db_partial, int db_rows,
                                     int im, int ic, int iy, int ix, int nf, int fy, int fx, void* stream) {
  BF_REQUIRE(db_partial && db_rows > 0, BF_ERR_ARG, "bf_conv_nhwc_backward_filter_dbp: no partials");
  return conv_nhwc_backward_filter_impl(X, E, dF, db, db_partial, db_rows, im, ic, iy, ix, nf, fy, fx, stream);
}

// ---- pre-packed filters: a layer re-lays its filter ONCE per weight update (both forms in one launch) and hands the
// packed operands to the forward / data-gradient kernels, instead of one re-layout launch in front of each of them
int bf_conv_nhwc_repack(const float* F, float* Wt_fwd, float* Wt_dx, int nf, int ic, int fy, int fx, void* stream) {
  BF_REQUIRE(F && Wt_fwd && Wt_dx && nf > 0 && ic > 0 && fy > 0 && fx > 0, BF_ERR_ARG, "bf_conv_nhwc_repack: bad arguments");
  launch_pdl(conv_repack_both_kernel, dim3(ew_grid((size_t)nf * ic * fy * fx, 256)), dim3(256), 0, as_stream(stream), F, Wt_fwd, Wt_dx,
             nf, ic, fy, fx);
  BF_LAUNCH_CHECK();
  return BF_OK;
}

int bf_conv_nhwc_forward_packed(const float* X, const float* Wt_fwd, const float* bias, float* Y, int im, int ic, int iy, int ix,
                                int nf, int fy, int fx, int act, void* stream) {
  BF_REQUIRE(conv_nhwc_supported(ic, nf, fy, fx), BF_ERR_ARG, "bf_conv_nhwc_forward_packed: unsupported shape (ic=%d nf=%d)", ic, nf);
  BF_REQUIRE(act == BF_ACT_LINEAR || act == BF_ACT_RELU, BF_ERR_ARG, "bf_conv_nhwc_forward_packed: only linear / relu are fused");
  if (im <= 0) return BF_OK;
  set_last_path(2);
  return conv_nhwc_igemm(X, Wt_fwd, bias, Y, im, iy, ix, ic, nf, fy, fx, 0, 0, act == BF_ACT_RELU ? 1 : 0,
                         getenv("BF_IGEMM_FOLD_FWD") != nullptr, as_stream(stream));
}

// ConvLayer -> [Activation("relu")] -> PoolLayer(2) in one kernel: P (im, oy/2, ox/2, nf) = maxpool2([relu](corr(X, F) + bias)),
// mask = the pooling layer's tie mask (one byte per pooled element).  Returns BF_ERR_ARG when no whole-image tiling exists.
int bf_conv_nhwc_forward_pool_packed(const float* X, const float* Wt_fwd, const float* bias, float* P, uint8_t* mask, int im, int ic,
                                     int iy, int ix, int nf, int fy, int fx, int relu_in, void* stream) {
  BF_REQUIRE(conv_nhwc_supported(ic, nf, fy, fx), BF_ERR_ARG, "bf_conv_nhwc_forward_pool_packed: unsupported shape (ic=%d nf=%d)", ic, nf);
  const int oy = iy - fy + 1, ox = ix - fx + 1;
  BF_REQUIRE(oy > 0 && ox > 0 && (oy & 1) == 0 && (ox & 1) == 0 && P && mask, BF_ERR_ARG,
             "bf_conv_nhwc_forward_pool_packed: the convolution output %dx%d does not pool 2x2", oy, ox);
  if (im <= 0) return BF_OK;
  set_last_path(2);
  return conv_nhwc_igemm(X, Wt_fwd, bias, P, im, iy, ix, ic, nf, fy, fx, 0, 0, 0, false, as_stream(stream), mask, relu_in ? 1 : 0);
}

int bf_conv_nhwc_backward_data_packed(const float* E, const float* Wt_dx, float* dX, int im, int ic, int iy, int ix, int nf, int fy,
                                      int fx, void* stream) {
  BF_REQUIRE(conv_nhwc_supported(nf, ic, fy, fx), BF_ERR_ARG, "bf_conv_nhwc_backward_data_packed: unsupported shape (ic=%d nf=%d)", ic, nf);
  if (im <= 0) return BF_OK;
  set_last_path(2);
  return conv_nhwc_igemm(E, Wt_dx, nullptr, dX, im, iy - fy + 1, ix - fx + 1, nf, ic, fy, fx, fy - 1, fx - 1, 0, true, as_stream(stream));
}

int bf_conv_nhwc_supported(int ic, int nf, int fy, int fx) {
  // forward needs (ic -> nf), the data gradient (nf -> ic), the filter gradient nf in {32,64,128}
  return (conv_nhwc_supported(ic, nf, fy, fx) && conv_nhwc_supported(nf, ic, fy, fx) && (nf == 32 || nf == 64 || nf == 128))
             ? 1 : 0;
}
