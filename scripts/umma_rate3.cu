// umma_rate3: issue rate of the filter-gradient MMA (kind::tf32, BOTH operands MN-major SWIZZLE_128B / 32-byte atoms, K = 8)
// for M in {64, 128}, N in {32, 64, 96, 128, 192}, with 1..3 issuing warps (own accumulator each), and the K-major form next
// to it.  Operand addresses advance by one K step per MMA inside a 16-step window, like the kernels do.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) k(int W, int N, int iters, int M, int mn, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t o = tid * 4; o < 200 * 1024; o += 1024) *(float*)(smem_raw + (base - smem_u32(smem_raw)) + o) = 1.0f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (warp < W && lane == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (mn ? (3u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    // MN-major: 16 K steps (128 pixel rows of 128 B) per operand window; A groups of 32 filters 16 KB apart, B groups one row apart
    auto mk_mn = [&](uint32_t addr, uint32_t lbo) { return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (1ull << 61); };
    auto mk_k = [&](uint32_t addr) { return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); };
    const uint32_t a_addr = base + warp * 16384u * 0u, b_addr = base + 128 * 1024;    // all warps read the same operands
    uint64_t ad0 = mn ? mk_mn(a_addr, 16384u) : mk_k(a_addr), bd0 = mn ? mk_mn(b_addr, 128u) : mk_k(b_addr);
    const uint32_t d = tmem + (uint32_t)(warp * (N <= 128 ? 128 : 256));
    for (int i = 0; i < iters; i += 4) {
      const uint64_t step = mn ? (uint64_t)(((i >> 2) & 3) * 256) : 0ull;     // four K steps of 1024 B per unrolled group
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = ad0 + step + (mn ? (uint64_t)(j * 64) : (uint64_t)(j * 2));
        const uint64_t bd = bd0 + step + (mn ? (uint64_t)(j * 64) : (uint64_t)(j * 2));
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[warp])), "r"(0u) : "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
int main() {
  long long* d; CK(cudaMalloc(&d, 148 * 8));
  size_t smem = 1024 + 200 * 1024;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int IT = 4096;
  for (int mn : {1, 0})
    for (int M : {128, 64})
      for (int N : {32, 64, 96, 128, 192}) {
        for (int W : {1, 2, 3}) {
          if (W * (N <= 128 ? 128 : 256) > 512) continue;
          for (int r = 0; r < 2; ++r) { k<<<148, 256, smem>>>(W, N, IT, M, mn, d); CK(cudaDeviceSynchronize()); }
          long long h[148]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
          double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
          printf("%s M=%3d N=%3d issuing warps=%d : %7.1f cycles per MMA (all warps together), floor smem %5.1f tensor %5.1f\n",
                 mn ? "MN-major" : "K-major ", M, N, W, s / 148 / (IT * W), (M + N) / 4.0, (M > 128 ? M : 128) * N / 256.0);
        }
      }
  return 0;
}
