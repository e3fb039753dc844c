// Measured int8 tensor-pipe peak (the denominator of k_sweep_tc's roofline): every SM issues back-to-back
// tcgen05.mma.cta_group::1.kind::i8 (M = 128, K = 32, N = 128 or 256; operands in shared memory, K-major,
// no swizzle; accumulators in TMEM) with no commits in between, one commit at the end. Timed with CUDA events
// over the whole grid, so clocks under a chip-wide tensor load are part of the number; SM 0 also reports
// clock64 cycles per MMA. Included at the end of sweep.cu (uses umma_desc / smem_u32 of sweep_tc.inl).

namespace tracs {

template <int N>
__global__ void __launch_bounds__(128) k_tc_rate(int reps, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t tcr_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < (128 + N) * 32 / 4; i += 128) reinterpret_cast<uint32_t *>(tcr_smem)[i] = 0x01000100u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
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
    const uint64_t da = umma_desc(smem_u32(tcr_smem), 16 * 128, 128);
    const uint64_t db = umma_desc(smem_u32(tcr_smem) + 128 * 32, (N / 8) * 128, 128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    if (blockIdx.x == 0) *cycles = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

}  // namespace tracs

extern "C" int tracs_tc_peak(double out[4]) {
  using namespace tracs;
  return guarded([&] {
    require_device();
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    DevBuf<long long> cyc(1);
    Timer T(0);
    const int reps = 100000;  // ~3-7 ms per launch
    auto run = [&](int N, double &tops, double &clk) {
      const size_t smem = (size_t)(128 + N) * 32 + 1024;
      float best = 1e30f;
      long long c = 0;
      for (int rep = 0; rep < 3; ++rep) {
        T.start();
        if (N == 128) k_tc_rate<128><<<n_sm, 128, smem>>>(reps, cyc.p);
        else k_tc_rate<256><<<n_sm, 128, smem>>>(reps, cyc.p);
        const float ms = T.stop();
        TRACS_CK(cudaGetLastError());
        if (rep > 0 && ms < best) {
          best = ms;
          TRACS_CK(cudaMemcpy(&c, cyc.p, sizeof c, cudaMemcpyDeviceToHost));
        }
      }
      tops = (double)n_sm * reps * 2.0 * 128.0 * N * 32.0 / ((double)best * 1e-3) / 1e12;
      clk = (double)c / reps;
    };
    run(128, out[0], out[2]);
    run(256, out[1], out[3]);
  });
}

// The same probe held for `seconds` (back-to-back launches of the N = 256 shape): out[0] = TOP/s over the second half of
// the run, out[1] = TOP/s over the whole run. A multi-second tensor-core kernel runs under the board's power cap
// (sw_power_cap at ~1 kW on B200); the burst figure of tracs_tc_peak is not what the pipe can sustain then.
extern "C" int tracs_tc_peak_sustained(double seconds, double out[2]) {
  using namespace tracs;
  return guarded([&] {
    require_device();
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    DevBuf<long long> cyc(1);
    const int reps = 100000;  // ~6.5 ms per launch
    const size_t smem = (size_t)(128 + 256) * 32 + 1024;
    const int launches = std::max(4, (int)(std::min(std::max(seconds, 0.05), 10.0) / 6.5e-3)) & ~1;
    cudaEvent_t e0, e1, e2;
    TRACS_CK(cudaEventCreate(&e0));
    TRACS_CK(cudaEventCreate(&e1));
    TRACS_CK(cudaEventCreate(&e2));
    TRACS_CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < launches; ++i) {
      if (i == launches / 2) TRACS_CK(cudaEventRecord(e1, 0));
      k_tc_rate<256><<<n_sm, 128, smem>>>(reps, cyc.p);
    }
    TRACS_CK(cudaEventRecord(e2, 0));
    TRACS_CK(cudaEventSynchronize(e2));
    TRACS_CK(cudaGetLastError());
    float ms_all = 0, ms_half = 0;
    cudaEventElapsedTime(&ms_all, e0, e2);
    cudaEventElapsedTime(&ms_half, e1, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    const double ops = (double)n_sm * reps * 2.0 * 128.0 * 256.0 * 32.0;
    out[0] = ops * (launches / 2) / ((double)ms_half * 1e-3) / 1e12;
    out[1] = ops * launches / ((double)ms_all * 1e-3) / 1e12;
  });
}

