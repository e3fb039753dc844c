// Measures the issue-to-completion rate of back-to-back tcgen05.mma kind::i8 (SS, K-major no-swizzle)
// for N = 128 and N = 256 on one SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int N, int PC = 0>
__global__ void __launch_bounds__(128) rate(int reps, long long *out, int per_commit) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < (128 + N) * 32 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01000100u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(smem), 16 * 128, 128);
    const uint64_t db = make_desc(smem_u32(smem) + 128 * 32, (N / 8) * 128, 128);
    const long long t0 = clock64();
    if (PC > 0) {
      // compile-time commit spacing: PC MMAs (fully unrolled), then one commit -- no runtime division in the issue loop
      for (int r = 0; r < reps; r += PC) {
#pragma unroll
        for (int q = 0; q < PC; ++q)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db),
                       "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bars[(r / PC) & 1])) : "memory");
      }
    } else
    for (int r = 0; r < reps; ++r) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db),
                   "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
      if (per_commit && (r % per_commit) == per_commit - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bars[(r / per_commit) & 1])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}
int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  const int reps = 2000;
  for (int rep = 0; rep < 1; ++rep) {
    for (int pc : {0, 5, 10, 20}) {
      rate<128><<<1, 128, (128 + 128) * 32 + 1024>>>(reps, d, pc);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("N=128 commit every %d MMAs: %.1f clk/MMA\n", pc, (double)h[1] / reps);
    }
#define RUNPC(PCV)                                                                                         \
    rate<128, PCV><<<1, 128, (128 + 128) * 32 + 1024>>>(1920, d, 0);                                       \
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);                                                          \
    printf("N=128 commit every %d MMAs (compile-time): %.1f clk/MMA\n", PCV, (double)h[1] / 1920);
    RUNPC(1) RUNPC(2) RUNPC(3) RUNPC(4) RUNPC(5) RUNPC(6) RUNPC(8) RUNPC(10) RUNPC(12) RUNPC(16) RUNPC(20) RUNPC(32)
    rate<128><<<1, 128, (128 + 128) * 32 + 1024>>>(reps, d, 0);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=128: issue %.1f clk/MMA, complete %.1f clk/MMA  (%s)\n", (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(cudaGetLastError()));
    rate<256><<<1, 128, (128 + 256) * 32 + 1024>>>(reps, d, 0);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=256: issue %.1f clk/MMA, complete %.1f clk/MMA  (%s)\n", (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
