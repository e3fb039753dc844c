// K1''' : the tensor-core sweep on 128 x 512 super-tiles (tcgen05, kind::i8, N = 256 MMAs, the whole TMEM as four
// 128 x 128 int32 accumulators). Included by sweep.cu after sweep_tc2.inl (same identity, same operand encoding).
//
// ncu on k_sweep_tc2 (profiles/r2_tc.md): tensor pipe 30 % active, issue slots 68 % active with 17 warps -- the kernel is
// bound by the INSTRUCTIONS that expand bit-planes into int8 operands (~1 instruction per operand byte, 24 KB per word
// and 128 x 128 tile), not by MMAs, commits, shared-memory bandwidth or load latency. The expansion of a row panel is
// the same for every column block it meets, so this kernel lets one row panel meet FOUR column blocks per word: per
// word it expands 128 + 512 rows (60 KB at three planes) for four tiles instead of 4 x 256 rows, i.e. 15 KB instead of
// 24 KB per tile-word, and the six N = 256 MMAs of a word keep the tensor pipe busy for 768 clk -- longer than the
// expansion takes. 20 producer warps (two 16-row x 2-chunk items each per word), one MMA-issuing warp, the first
// four warps also drain TMEM in the epilogue.

namespace tracs {

constexpr int TC3_PRODUCER_WARPS = 20;
constexpr int TC3_THREADS = TC3_PRODUCER_WARPS * 32 + 32;
template <int NP> struct Tc3Geom {
  static constexpr uint32_t A_BYTES = NP * 2 * 2048;    // 128 rows: NP planes x 2 K-chunks x (16 row groups x 128 B)
  static constexpr uint32_t BH_BYTES = NP * 2 * 4096;   // 256 rows (one N = 256 operand)
  static constexpr uint32_t STAGE_BYTES = A_BYTES + 2 * BH_BYTES;
  static constexpr int STAGES = NP == 3 ? 3 : 2;        // 180 KB / 160 KB
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
};

// super-tile s -> (row-block, first col-block | number of col-blocks << 28): up to four consecutive tiles of a row-block
__global__ void k_stile_table(const uint32_t *__restrict__ rb_list, const uint32_t *__restrict__ sprefix, uint32_t n_rb, uint32_t cb_min,
                              uint32_t n_cb, uint32_t n_stiles, uint2 *__restrict__ table) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_stiles) return;
  uint32_t lo = 0, hi = n_rb;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sprefix[mid] <= s) lo = mid; else hi = mid;
  }
  const uint32_t rb = rb_list[lo];
  const uint32_t cb0 = max(rb, cb_min) + 4u * (s - sprefix[lo]);
  table[s] = make_uint2(rb, cb0 | (min(4u, n_cb - cb0) << 28));
}

template <int NP>
__global__ void __launch_bounds__(TC3_THREADS, 1) k_sweep_tc3(const SweepArgs a, const uint2 *__restrict__ stiles, uint32_t n_stiles,
                                                              uint32_t dbg /* timing experiments only (TRACS_TC3_DBG, profiles/r2_tc.md): 1 no proxy fence, 2 no operand
                                                                              loads, 4 no expansion arithmetic, 8 no MMAs -- results are wrong when set */) {
  using G = Tc3Geom<NP>;
  constexpr int S = G::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *stage_base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[S], empty[S], tmem_full, tmem_empty;
  __shared__ uint32_t tmem_slot;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t MMA_WARP = TC3_PRODUCER_WARPS;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], TC3_PRODUCER_WARPS);  // one arrival per producer warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full, 1);
    mbar_init(&tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {  // the whole tensor memory: one CTA per SM (launch bounds + shared memory guarantee it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  const uint32_t nw = a.Wp;
  uint32_t it = 0;  // running word counter (stage = it % S)
  uint32_t tile_iter = 0;

  for (uint32_t st = blockIdx.x; st < n_stiles; st += gridDim.x, ++tile_iter) {
    const uint2 rc = __ldg(stiles + st);
    const uint32_t rb = rc.x, cb0 = rc.y & 0x0FFFFFFFu, nblk = rc.y >> 28, nhalf = (nblk + 1) >> 1;

    if (warp < MMA_WARP) {
      // ===== producers: two items per warp and word; item j = 16 consecutive rows of the 640-row stack x 2 K-chunks =====
      // rows 0..127 = row operand (row-block rb), rows 128..639 = the 512 consecutive samples of the column blocks
      const uint32_t h = lane >> 4;
      const uint4 *src[2];
      uint32_t off[2];
      bool is_row[2], live[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t r = (warp + TC3_PRODUCER_WARPS * t) * 16u + (lane & 15u);  // 0..639
        is_row[t] = r < 128u;
        const uint32_t rB = r - 128u;
        const uint32_t sample = is_row[t] ? rb * TILE + r : cb0 * TILE + rB;
        live[t] = is_row[t] || ((rB >> 8) < nhalf && sample < a.Npad);
        src[t] = a.planes + (live[t] ? sample : 0u);
        off[t] = is_row[t] ? (r >> 3) * 128u + (r & 7u) * 16u + h * 2048u
                           : G::A_BYTES + (rB >> 8) * G::BH_BYTES + ((rB & 255u) >> 3) * 128u + (rB & 7u) * 16u + h * 4096u;
      }
      // One plane word half (16 sites) -> 16 operand bytes: the bits are masked in place into PRMT selector nibbles
      // (bit 4k + q of the half -> byte k of register q; any fixed site permutation is fine as long as both operands
      // use it) and PRMT picks +1 / -1 out of a two-register byte pool: 9 instructions per 16 bytes instead of 16.
      // Stores go through explicit shared-space addresses (STS.128; the generic form cost 64-bit address arithmetic).
      auto sts128 = [](uint32_t saddr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
      };
      auto pm = [](uint32_t a, uint32_t b, uint32_t sel) {
        uint32_t r;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
        return r;
      };
      auto expand = [&](const uint4 &x, uint32_t dst, uint32_t kstride, uint32_t nmul) {
        const uint32_t hi = x.z | x.w, lo = x.y | x.w;   // A = 00, C = 01, G = 10, T = 11
        const uint32_t pl[3] = {hi, lo, ~(hi ^ lo)};
        if (NP == 3) {
          // selector 0 -> -1; 1, 2, 4 -> +1. The pool is derived from a kernel argument so that it lives in ONE register
          // (as a literal the compiler re-materialised it in front of every PRMT: 16 extra instructions per word)
          const uint32_t PA = 0x000101FEu + a.one, PB = a.one;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const uint32_t v = pl[p] >> (16u * h);
            sts128(dst + 2 * p * kstride, pm(PA, PB, v & 0x1111u), pm(PA, PB, v & 0x2222u), pm(PA, PB, v & 0x4444u),
                   pm(PA, PB, (v >> 3) & 0x1111u));
          }
        } else {
          const uint32_t nb = x.x & x.y & x.z & x.w;
          uint32_t nm[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) nm[q] = ((nb >> (4 * h + q)) & 0x01010101u) * 0xFFu;  // 0xFF in the bytes of N sites
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = ((((pl[p] >> (4 * h + q)) & 0x01010101u) * 0xFEu) ^ 0xFFFFFFFFu) & ~nm[q];
            sts128(dst + 2 * p * kstride, o[0], o[1], o[2], o[3]);
          }
          sts128(dst + 2 * 3 * kstride, (nm[0] & 0x01010101u) * nmul, (nm[1] & 0x01010101u) * nmul, (nm[2] & 0x01010101u) * nmul,
                 (nm[3] & 0x01010101u) * nmul);
        }
      };
      // Operand words: the load of word w + 1 is issued before word w is expanded and must have RETURNED by the proxy
      // fence that closes the expansion (the fence is a CTA-wide memory barrier and waits for this thread's outstanding
      // loads: with two words in flight it exposed a DRAM round trip per word, profiles/r2_tc.md). So the loads are made
      // L2 hits: the lines of word w + PFD are requested into L2 well ahead, one request per 128-byte line.
      constexpr uint32_t PFD = 24;
      const bool pf_lane = (lane & 0x17u) == 0u;  // lanes 0 and 8: the two lines of this warp's 16 rows
      const uint32_t sbase_u32 = smem_u32(stage_base);
      const uint4 zero = make_uint4(0, 0, 0, 0);
      uint4 c[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        c[t] = live[t] ? __ldg(src[t]) : zero;
        if (live[t] && pf_lane)
          for (uint32_t w = 1; w < PFD && w < nw; ++w) asm volatile("prefetch.global.L2 [%0];" ::"l"(src[t] + (size_t)w * a.Npad));
      }
#pragma unroll 2
      for (uint32_t w = 0; w < nw; ++w, ++it) {
        uint4 x[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          x[t] = c[t];
          if (live[t] && w + 1 < nw && !(dbg & 2u)) c[t] = __ldg(src[t] + (size_t)(w + 1) * a.Npad);
          if (live[t] && pf_lane && w + PFD < nw) asm volatile("prefetch.global.L2 [%0];" ::"l"(src[t] + (size_t)(w + PFD) * a.Npad));
        }
        const uint32_t s = it % S;
        mbar_wait(&empty[s], ((it / S) & 1u) ^ 1u);
        const uint32_t sb = sbase_u32 + s * G::STAGE_BYTES;
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (live[t]) {
            if (dbg & 4u) {
              sts128(sb + off[t], x[t].x, x[t].y, x[t].z, x[t].w);
              sts128(sb + off[t] + 2 * (is_row[t] ? 2048u : 4096u), x[t].x, x[t].y, x[t].z, x[t].w);
              sts128(sb + off[t] + 4 * (is_row[t] ? 2048u : 4096u), x[t].x, x[t].y, x[t].z, x[t].w);
            } else {
              expand(x[t], sb + off[t], is_row[t] ? 2048u : 4096u, is_row[t] ? 0xFDu : 1u);
            }
          }
        if (!(dbg & 1u)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    } else {
      // ===== MMA issuer: one elected thread; per word NP MMAs (M = 128, N = 256, K = 32) per 256-column half =====
      if (lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
        mbar_wait(&tmem_empty, (tile_iter & 1u) ^ 1u);  // epilogue of the previous super-tile has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sbase = smem_u32(stage_base);
        for (uint32_t w = 0; w < nw; ++w, ++it) {
          const uint32_t s = it % S;
          mbar_wait(&full[s], (it / S) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st0 = sbase + s * G::STAGE_BYTES;
          for (uint32_t b = 0; b < ((dbg & 8u) ? 0u : nhalf); ++b) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const uint64_t da = umma_desc(st0 + p * 2 * 2048, 2048, 128);
              const uint64_t db = umma_desc(st0 + G::A_BYTES + b * G::BH_BYTES + p * 2 * 4096, 4096, 128);
              const uint32_t acc = p != 0 ? 1u : (uint32_t)(w != 0u);
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem + b * 256u),
                  "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
                  : "memory");
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&tmem_full)) : "memory");
      } else {
        it += nw;
      }
      __syncwarp();
    }

    if (warp < 4) {
      // ===== epilogue: TMEM -> registers -> threshold -> append, one 128 x 128 block after the other =====
      mbar_wait(&tmem_full, tile_iter & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int32_t W3 = 3 * (int32_t)(a.Wp * 32u);
      const uint32_t gi = rb * TILE + warp * 32 + lane;  // TMEM lane = tile row
      const int32_t ci = NP == 4 ? 3 * (int32_t)__ldg(a.tc_ncnt + gi) : 0;
      for (uint32_t blk = 0; blk < nblk; ++blk) {
        const uint32_t cb = cb0 + blk;
        for (uint32_t c0 = 0; c0 < (uint32_t)TILE; c0 += 32) {
          uint32_t v[32];
          const uint32_t taddr = tmem + ((warp * 32u) << 16) + blk * TILE + c0;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
              "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // d = W - matches = (3 W - T - 3 (cnt_i + cnt_j)) / 4 (exact)
          const int32_t cj_mine = NP == 4 ? 3 * (int32_t)__ldg(a.tc_ncnt + cb * TILE + c0 + lane) : 0;
          uint32_t keep = 0, cnt = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t gj = cb * TILE + c0 + j;
            const int32_t cj = NP == 4 ? __shfl_sync(0xFFFFFFFFu, cj_mine, j) : 0;
            const int32_t d = (W3 - (int32_t)v[j] - ci - cj) >> 2;
            v[j] = (uint32_t)d;
            if (gi < a.i_end && gj < a.n && gj > gi && gj >= a.j_start && d <= a.dist) {
              keep |= 1u << j;
              cnt++;
            }
          }
          if (!__any_sync(0xFFFFFFFFu, cnt != 0u)) continue;
          uint32_t incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
          }
          const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
          unsigned long long base = 0;
          if (lane == 31) base = atomicAdd(a.counter, (unsigned long long)wtot);
          base = __shfl_sync(0xFFFFFFFFu, base, 31);
          unsigned long long pos = base + (incl - cnt);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if ((keep >> j) & 1u) {
              if (pos < a.cap) {
                a.keys[pos] = ((uint64_t)gi << 32) | (cb * TILE + c0 + j);
                a.dvals[pos] = v[j];
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
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace tracs
