// Probe: one CTA, tcgen05.mma kind::i8, M=128 N=128, operands written to shared memory by the threads in
// the canonical K-major no-swizzle layout, accumulator read back from TMEM. Checks against the CPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes (128 B contiguous)
// element (row r, byte k): off = (k/16)*LBO + (r/8)*SBO + (r%8)*16 + k%16
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;               // layout_type = 0 (no swizzle), base_offset = 0
}

constexpr int M = 128, N = 128;

__global__ void __launch_bounds__(128) probe(const int8_t *A, const int8_t *B, int K, int32_t *C, int lbo_is_k) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t *sA = smem, *sB = smem + (size_t)M * K;
  const int kch = K / 16;
  // layout [kchunk][rowgroup][8][16]
  for (int idx = tid; idx < M * kch; idx += 128) {
    const int r = idx % M, kc = idx / M;
    const uint4 va = *reinterpret_cast<const uint4 *>(A + (size_t)r * K + kc * 16);
    const uint4 vb = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + kc * 16);
    const size_t off = ((size_t)kc * (M / 8) + r / 8) * 128 + (r % 8) * 16;
    *reinterpret_cast<uint4 *>(sA + off) = va;
    *reinterpret_cast<uint4 *>(sB + off) = vb;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // generic-proxy writes -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    // idesc: c_format S32 (2) @4, a_format INT8 (1) @7, b_format INT8 (1) @10, K-major both, n>>3 @17, m>>4 @24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t k_stride = (M / 8) * 128;  // bytes between K-adjacent core matrices
    const uint32_t m_stride = 128;            // bytes between M-adjacent 8-row groups
    for (int k = 0; k < K / 32; ++k) {
      const uint32_t a0 = smem_u32(sA) + k * 2 * k_stride, b0 = smem_u32(sB) + k * 2 * k_stride;
      const uint64_t da = lbo_is_k ? make_desc(a0, k_stride, m_stride) : make_desc(a0, m_stride, k_stride);
      const uint64_t db = lbo_is_k ? make_desc(b0, k_stride, m_stride) : make_desc(b0, m_stride, k_stride);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the MMAs
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32w..32w+31 (row = lane), 128 columns in 4 chunks of 32
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) C[(size_t)tid * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  for (int lbo_is_k = 1; lbo_is_k >= 0; --lbo_is_k) {
    for (int K : {32, 64, 160, 320}) {
      std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
      srand(1 + K);
      for (auto &x : hA) x = (int8_t)((rand() % 5 == 0) ? -3 : (rand() & 1));
      for (auto &x : hB) x = (int8_t)(rand() & 1);
      int8_t *dA, *dB; int32_t *dC;
      cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dC, (size_t)M * N * 4);
      cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
      cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
      cudaMemset(dC, 0xFF, (size_t)M * N * 4);
      const size_t smem = (size_t)(M + N) * K;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      probe<<<1, 128, smem>>>(dA, dB, K, dC, lbo_is_k);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<int32_t> hC((size_t)M * N);
      cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
      long bad = 0;
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          int32_t ref = 0;
          for (int k = 0; k < K; ++k) ref += (int32_t)hA[(size_t)i * K + k] * (int32_t)hB[(size_t)j * K + k];
          if (ref != hC[(size_t)i * N + j]) {
            if (bad < 3) printf("  mismatch (%d,%d): got %d want %d\n", i, j, hC[(size_t)i * N + j], ref);
            ++bad;
          }
        }
      printf("lbo_is_k=%d K=%d: %s (%ld bad) cuda=%s\n", lbo_is_k, K, bad ? "FAIL" : "PASS", bad, cudaGetErrorString(e));
      cudaFree(dA); cudaFree(dB); cudaFree(dC);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
