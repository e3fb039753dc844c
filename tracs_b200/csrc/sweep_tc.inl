// K1': the full-length pair sweep on the 5th-generation tensor cores (tcgen05, kind::i8, accumulators
// in TMEM). Included by sweep.cu (same translation unit: shares SweepArgs and the tile list).
//
// Identity. Expand every sample to one int8 per (site, plane): x[s][p] = bit p of the base
// mask, plus a fifth "N" column n[s] = 1 iff the mask is 1111. For masks that are single bases or N,
//     |{s : m_i[s] & m_j[s] != 0}|  =  sum_s ( sum_p x_i[s][p] x_j[s][p]  -  3 n_i[s] n_j[s] )
// (base/base 1 or 0, base/N 1, N/N 4 - 3 = 1), i.e. ONE int8 GEMM whose row operand carries -3 in the
// N column and whose column operand carries +1. int32 accumulation is exact. Partial ambiguity codes
// (two or three bases) break the identity, so ingest records whether any occurs at a variable site and
// such alignments stay on the LOP3/POPC kernel.
//
// The int8 operands never exist in HBM (at config-5 size they would take 250 GB): 8 producer warps
// expand the bit-planes of one 32-site word (5 planes x 32 bytes per sample) straight into shared memory
// in the canonical K-major no-swizzle UMMA layout, one elected thread issues 5 tcgen05.mma
// (M=128, N=128, K=32) per word, tcgen05.commit recycles the stage, and 4 warps read the 128x128 int32
// tile back with tcgen05.ld for the fused threshold + append epilogue.

namespace tracs {

constexpr int TC_STAGES = 4;
constexpr int TC_PLANES = 5;                               // A, C, G, T, N
constexpr int TC_KCH = TC_PLANES * 2;                      // 16-byte K chunks per word (32 bytes per plane)
constexpr uint32_t TC_SIDE_BYTES = TC_KCH * (TILE / 8) * 128;   // 20480: one operand, one word
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_SIDE_BYTES;
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024;
constexpr int TC_PRODUCERS = 256;   // 8 warps: thread = (operand side, row)
constexpr int TC_THREADS = TC_PRODUCERS + 32;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrix = 8 rows x 16 bytes;
// LBO = bytes between K-adjacent core matrices, SBO = bytes between 8-row groups (tools/tc_probe.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_sweep_tc(const SweepArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *stage_base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[TC_STAGES], empty[TC_STAGES], tmem_full, tmem_empty;
  __shared__ uint32_t tmem_slot;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], TC_PRODUCERS / 32);  // one arrival per producer warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full, 1);
    mbar_init(&tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_PRODUCERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  const uint32_t nw = a.Wp;  // one pipeline step per 32-site word
  uint32_t it = 0;           // running word counter (stage = it % TC_STAGES)
  uint32_t tile_iter = 0;

  uint2 rc_next = blockIdx.x < a.n_tiles ? __ldg(a.tile_table + blockIdx.x) : make_uint2(0, 0);
  for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tile_iter) {
    const uint32_t rb = rc_next.x, cb = rc_next.y;
    if (tile + gridDim.x < a.n_tiles) rc_next = __ldg(a.tile_table + tile + gridDim.x);  // in flight during this tile

    if (warp < TC_PRODUCERS / 32) {
      // ===== producers: bit-planes -> int8 operands in the UMMA layout =====
      const uint32_t side = tid >> 7, r = tid & 127;
      const uint4 *src = a.planes + (size_t)(side ? cb : rb) * TILE + r;
      const uint32_t row_off = (r >> 3) * 128 + (r & 7) * 16;
      const uint32_t nmul = side ? 1u : 0xFDu;  // N column: -3 on the row operand, +1 on the column operand
      // Two words per trip: both stages are written, then ONE proxy fence and two arrivals, so the
      // store -> fence -> arrive latency chain is paid once per pair of words (the producers were
      // latency-bound on it, not ALU- or store-bound: profiles/r1_tc_ncu.md). Wp is a multiple of 8.
      auto expand = [&](const uint4 &x, uint32_t slot) {
        uint8_t *dst = stage_base + (size_t)slot * TC_STAGE_BYTES + side * TC_SIDE_BYTES + row_off;
        const uint32_t pl[TC_PLANES] = {x.x, x.y, x.z, x.w, x.x & x.y & x.z & x.w};
#pragma unroll
        for (int p = 0; p < TC_PLANES; ++p) {
          const uint32_t v = pl[p];
          const uint32_t mul = (p == 4) ? nmul : 1u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 o;
            o.x = ((v >> (4 * h + 0)) & 0x01010101u) * mul;
            o.y = ((v >> (4 * h + 1)) & 0x01010101u) * mul;
            o.z = ((v >> (4 * h + 2)) & 0x01010101u) * mul;
            o.w = ((v >> (4 * h + 3)) & 0x01010101u) * mul;
            *reinterpret_cast<uint4 *>(dst + (size_t)(2 * p + h) * ((TILE / 8) * 128)) = o;
          }
        }
      };
      // A pipeline step is a PAIR of words (two 40 KB stages): one "empty" wait, one proxy fence and one
      // "full" arrival per pair, and the MMA side commits once per pair. Probes (tools/tc_rate.cu and
      // handshake-only runs of this kernel) show that MMAs queued behind a tcgen05.commit start only
      // after it retires: ~370 clk of drained pipe per commit, against 64 clk per MMA. Per-word commits
      // cost 690 clk/word, per-pair commits 505 (the MMAs alone: 320); shared memory holds 4 words.
      uint4 c0 = __ldg(src), c1 = __ldg(src + (size_t)a.Npad);
      for (uint32_t w = 0; w < nw; w += 2, it += 2) {
        const uint4 x0 = c0, x1 = c1;
        if (w + 2 < nw) {
          c0 = __ldg(src + (size_t)(w + 2) * a.Npad);
          c1 = __ldg(src + (size_t)(w + 3) * a.Npad);
        }
        const uint32_t pair = (it >> 1) % (TC_STAGES / 2);
        mbar_wait(&empty[pair], (((it >> 1) / (TC_STAGES / 2)) & 1u) ^ 1u);
        expand(x0, 2 * pair);
        expand(x1, 2 * pair + 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[pair]);  // 8 arrivals per stage instead of 256 on one shared-memory word
      }
    } else {
      // ===== MMA issuer: one elected thread =====
      if (lane == 0) {
        // instruction descriptor: D = S32, A = B = signed int8, both K-major, N >> 3, M >> 4
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
        const uint32_t k_stride = (TILE / 8) * 128, m_stride = 128;
        mbar_wait(&tmem_empty, (tile_iter & 1u) ^ 1u);  // epilogue of the previous tile has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // The issuing thread is ONE thread running dependent scalar code between MMAs (each MMA occupies the
        // tensor pipe for only 64 clk), so everything that can be folded at compile time is: the loop is
        // unrolled over the two pair-stages, descriptors are a base plus immediate offsets.
        const uint64_t desc0 = umma_desc(smem_u32(stage_base), k_stride, m_stride);
        for (uint32_t w = 0; w < nw; w += 4, it += 4) {
          const uint32_t phase = ((it >> 1) / (TC_STAGES / 2)) & 1u;
#pragma unroll
          for (int pair = 0; pair < TC_STAGES / 2; ++pair) {
            mbar_wait(&full[pair], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
              for (int p = 0; p < TC_PLANES; ++p) {
                const uint64_t da = desc0 + (uint64_t)(((2 * pair + q) * TC_STAGE_BYTES + p * 2 * k_stride) >> 4);
                const uint64_t db = da + (uint64_t)(TC_SIDE_BYTES >> 4);
                const uint32_t acc = (pair | q | p) != 0 ? 1u : (uint32_t)(w != 0u);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                    "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
                    : "memory");
              }
            }
            // commit: both stages of the pair may be overwritten once these MMAs have read them
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&empty[pair])) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&tmem_full)) : "memory");
      } else {
        it += nw;
      }
      __syncwarp();
    }

    if (warp < 4) {
      // ===== epilogue: TMEM -> registers -> threshold -> append =====
      mbar_wait(&tmem_full, tile_iter & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t total_bits = a.Wp * 32u;
      const uint32_t gi = rb * TILE + warp * 32 + lane;  // TMEM lane = tile row
      for (uint32_t c0 = 0; c0 < (uint32_t)TILE; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((warp * 32u) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
            "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t keep = 0, cnt = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const uint32_t gj = cb * TILE + c0 + j;
          const int32_t d = (int32_t)(total_bits - v[j]);
          if (gi < a.i_end && gj < a.n && gj > gi && gj >= a.j_start && d <= a.dist) {
            keep |= 1u << j;
            cnt++;
          }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (wtot) {
          unsigned long long base = 0;
          if (lane == 31) base = atomicAdd(a.counter, (unsigned long long)wtot);
          base = __shfl_sync(0xFFFFFFFFu, base, 31);
          unsigned long long pos = base + (incl - cnt);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if ((keep >> j) & 1u) {
              if (pos < a.cap) {
                a.keys[pos] = ((uint64_t)gi << 32) | (cb * TILE + c0 + j);
                a.dvals[pos] = total_bits - v[j];
              }
              pos++;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_PRODUCERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

}  // namespace tracs
