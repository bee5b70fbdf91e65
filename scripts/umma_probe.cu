// umma_probe: discovers how tcgen05.mma maps a shared-memory matrix descriptor to shared-memory words.
//
// Shared memory is filled with word ids; the OTHER operand is an 8x8 identity in the known-good K-major
// SWIZZLE_128B layout, so one K=8 MMA returns, for every (row, k<8), the id of the word the tensor core
// read for that element.  Two passes (low / high 11 bits of the id: TF32 holds integers < 2048 exactly)
// reconstruct the byte address.  Used once to pin the descriptor semantics the conv kernels rely on:
// row-shifted K-major views, MN-major SWIZZLE_128B, and overlapping ("Toeplitz") no-swizzle views.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
//   ./umma_probe <side:A|B> <major:K|MN> <layout:0|2|4|6> <start_off_bytes> <lbo_bytes> <sbo_bytes> <base_off> <MN>
// prints, for every row r < MN and k < 8, the byte offset (relative to the probe region) that was read.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  int side;        // 0: probe A (B = identity), 1: probe B (A = identity)
  int mn_major;    // probed operand is MN-major
  uint32_t layout, start_off, lbo, sbo, base_off;
  int MN;          // rows of the probed operand (A: M = 64/128, B: N multiple of 16)
  int shift;       // id bits to expose: (word >> shift) & 2047
  int f16;         // 1: kind::f16 with half elements (K = 16), 0: kind::tf32 (K = 8)
};

constexpr int PROBE_BYTES = 96 * 1024;
constexpr int IDENT_BYTES = 16 * 1024;

__global__ void __launch_bounds__(128) probe_kernel(Params p, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* probe = reinterpret_cast<float*>(gen);
  float* ident = reinterpret_cast<float*>(gen + PROBE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.f16) {
    __half* ph = reinterpret_cast<__half*>(gen);
    for (int w = tid; w < PROBE_BYTES / 2; w += 128) ph[w] = __float2half((float)((w >> p.shift) & 2047));
  } else {
    for (int w = tid; w < PROBE_BYTES / 4; w += 128) probe[w] = (float)((w >> p.shift) & 2047);
  }
  for (int w = tid; w < IDENT_BYTES / 4; w += 128) ident[w] = 0.f;
  __syncthreads();
  if (!p.f16 && tid < 8) {   // identity rows n = k = tid, K-major SW128: row r at r*128, chunk c at (c ^ (r&7))*16
    int r = tid, k = tid;
    ident[(r * 128 + (((k >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2)) / 4] = 1.0f;
  }
  if (p.f16 && tid < 16) {
    int r = tid, k = tid;
    reinterpret_cast<__half*>(ident)[(r * 128 + (((k >> 3) ^ (r & 7)) << 4) + ((k & 7) << 1)) / 2] = __float2half(1.0f);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int M = p.side == 0 ? p.MN : 128;
  const int N = p.side == 0 ? 16 : p.MN;
  if (tid == 0) {
    uint64_t pd = (uint64_t)(((base + p.start_off) >> 4) & 0x3FFFu) | ((uint64_t)((p.lbo >> 4) & 0x3FFFu) << 16) |
                  ((uint64_t)((p.sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)(p.base_off & 7u) << 49) |
                  ((uint64_t)p.layout << 61);
    uint64_t id = (uint64_t)(((base + PROBE_BYTES) >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (!p.f16) idesc |= (2u << 7) | (2u << 10);   // f16 kind: A = B = F16 (format 0)
    if (p.mn_major) idesc |= (p.side == 0 ? (1u << 15) : (1u << 16));
    uint64_t ad = p.side == 0 ? pd : id, bd = p.side == 0 ? id : pd;
    if (p.f16)
      asm volatile(
          "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
          : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (spin > 100000000u) __trap();
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D is M x N: lane = row.  out[row][col], 128 x 128 floats
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

int main(int argc, char** argv) {
  if (argc < 9) { printf("usage\n"); return 1; }
  Params p;
  p.side = argv[1][0] == 'B';
  p.mn_major = strcmp(argv[2], "MN") == 0;
  p.layout = atoi(argv[3]); p.start_off = atoi(argv[4]); p.lbo = atoi(argv[5]); p.sbo = atoi(argv[6]);
  p.base_off = atoi(argv[7]); p.MN = atoi(argv[8]);
  p.f16 = argc > 9 && atoi(argv[9]) == 1;
  const int KK = p.f16 ? 16 : 8, ES = p.f16 ? 2 : 4;
  float* d;
  CK(cudaMalloc(&d, 128 * 128 * 4));
  static float h[2][128 * 128];
  size_t smem = 1024 + PROBE_BYTES + IDENT_BYTES;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int pass = 0; pass < 2; ++pass) {
    p.shift = pass * 11;
    CK(cudaMemset(d, 0, 128 * 128 * 4));
    probe_kernel<<<1, 128, smem>>>(p, d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h[pass], d, sizeof(h[0]), cudaMemcpyDeviceToHost));
  }
  printf("EXP side=%s major=%s layout=%u start=%u lbo=%u sbo=%u boff=%u MN=%d f16=%d\n", argv[1], argv[2], p.layout, p.start_off,
         p.lbo, p.sbo, p.base_off, p.MN, p.f16);
  // element (r, k): probing A -> D[r][k];  probing B -> D[k][r]
  int rows = p.MN;
  for (int r = 0; r < rows; ++r) {
    if (!(r < 10 || (r % 8) < 2 || r == rows - 1 || (r >= 30 && r < 36))) continue;
    printf("r=%3d:", r);
    for (int k = 0; k < KK; ++k) {
      float lo = p.side == 0 ? h[0][r * 128 + k] : h[0][k * 128 + r];
      float hi = p.side == 0 ? h[1][r * 128 + k] : h[1][k * 128 + r];
      long word = (long)lo + ((long)hi << 11);
      printf(" %6ld", word * ES);
    }
    printf("\n");
  }
  return 0;
}
