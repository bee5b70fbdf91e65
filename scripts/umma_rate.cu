// umma_rate: cycles per tcgen05.mma (kind::tf32, M=128) for different N and operand addressing modes.
//   ./umma_rate  -> table.  One CTA per SM (148), each issues ITERS MMAs back to back, then one commit.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct P { int M, f16, unroll; int N, iters, a_mn, b_mn; uint32_t a_off, a_step, b_off, b_step, a_lbo, a_sbo, b_lbo, b_sbo, a_layout, b_layout; int nacc; };
__global__ void __launch_bounds__(128) k(P p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t o = tid * 4; o < 160 * 1024; o += 512) *(float*)(smem_raw + (base - smem_u32(smem_raw)) + o) = 1.0f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
    if (!p.f16) idesc |= (2u << 7) | (2u << 10); else idesc |= (1u << 7) | (1u << 10);
    if (p.a_mn) idesc |= 1u << 15;
    if (p.b_mn) idesc |= 1u << 16;
    auto mk = [&](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
    };
    long long t0 = clock64();
    if (p.unroll) {
      uint64_t ad[8], bd[8];
      for (int j = 0; j < 8; ++j) {
        ad[j] = mk(base + p.a_off + (j & 3) * p.a_step, p.a_lbo, p.a_sbo, p.a_layout);
        bd[j] = mk(base + 96 * 1024 + p.b_off + (j & 3) * p.b_step, p.b_lbo, p.b_sbo, p.b_layout);
      }
      for (int i = 0; i < p.iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (p.f16)
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem + (uint32_t)((j % p.nacc) * p.N)), "l"(ad[j]), "l"(bd[j]), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem + (uint32_t)((j % p.nacc) * p.N)), "l"(ad[j]), "l"(bd[j]), "r"(idesc), "r"(1u) : "memory");
        }
      }
    } else
    for (int i = 0; i < p.iters; ++i) {
      const uint32_t j = i & 15;
      uint64_t ad = mk(base + p.a_off + j * p.a_step, p.a_lbo, p.a_sbo, p.a_layout);
      uint64_t bd = mk(base + 96 * 1024 + p.b_off + j * p.b_step, p.b_lbo, p.b_sbo, p.b_layout);
      uint32_t d = tmem + (uint32_t)((i % p.nacc) * p.N);
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i >= p.nacc ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
static void run(const char* name, P p) {
  long long* d; CK(cudaMalloc(&d, 148 * 8));
  size_t smem = 1024 + 200 * 1024;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<148, 128, smem>>>(p, d); CK(cudaDeviceSynchronize());
  k<<<148, 128, smem>>>(p, d); CK(cudaDeviceSynchronize());
  long long h[148]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  printf("%-58s N=%3d : %7.1f cycles / MMA\n", name, p.N, s / 148 / p.iters);
  cudaFree(d);
}
int main() {
  const int IT = 4096;
  for (int M : {128, 64})
    for (int N : {32, 64, 128, 256}) {
      int nacc = 512 / N; if (nacc > 4) nacc = 4;
      char nm[96];
      snprintf(nm, sizeof nm, "tf32 M=%d K-major sw128 unrolled", M);
      run(nm, P{M, 0, 1, N, IT, 0, 0, 0, 32, 0, 32, 16, 1024, 16, 1024, 2, 2, nacc});
      snprintf(nm, sizeof nm, "tf32 M=%d K-major sw128 unrolled, 1 accumulator", M);
      run(nm, P{M, 0, 1, N, IT, 0, 0, 0, 32, 0, 32, 16, 1024, 16, 1024, 2, 2, 1});
      snprintf(nm, sizeof nm, "f16  M=%d K-major sw128 unrolled", M);
      run(nm, P{M, 1, 1, N, IT, 0, 0, 0, 32, 0, 32, 16, 1024, 16, 1024, 2, 2, nacc});
      snprintf(nm, sizeof nm, "tf32 M=%d K-major sw128 rolled loop", M);
      run(nm, P{M, 0, 0, N, IT, 0, 0, 0, 32, 0, 32, 16, 1024, 16, 1024, 2, 2, nacc});
    }
  for (int N : {32, 96}) {
    run("tf32 M=128 MN-major both unrolled", P{128, 0, 1, N, IT, 1, 1, 0, 1024, 0, 1024, 25600, 512, 128, 512, 1, 1, 512 / N > 4 ? 4 : 512 / N});
  }
  return 0;
}
