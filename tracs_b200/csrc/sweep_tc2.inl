// K1'' : second generation of the tensor-core sweep (tcgen05, kind::i8, int32 accumulators in TMEM). Included by
// sweep.cu after sweep_tc.inl (shares umma_desc, the barriers helpers, SweepArgs and the tile table).
//
// What changed against k_sweep_tc, and why (profiles/r2_tc.md):
//   * A clean probe (tools/tc_rate.cu, commit spacing as a compile-time constant) shows tcgen05.commit is FREE from
//     3 MMAs per commit on (64.2 clk per M128 N128 K32 MMA with or without commits); the "370 clk drain per commit" of
//     round 1 was the probe's own runtime division in the issue loop. The kernel was PRODUCER-bound: expanding two
//     words took ~1000 clk against 640 clk of MMAs.
//   * Three int8 planes per site instead of five. With the 2-bit code (hi, lo) of an unambiguous base and
//     u = 2 hi - 1, v = 2 lo - 1, w = u v (all +-1):   [base_i == base_j] = (1 + u_i u_j + v_i v_j + w_i w_j) / 4,
//     so  4 * matches = W + sum_s (u u' + v v' + w w')  -- three K = 32 MMAs per 32-site word, no N column when the
//     alignment has no N at a variable site (NP = 3). With N (mask 1111 -> zero vector in the three planes) a fourth
//     plane carries the N indicator with weights -3 (row operand) / +1 (column operand), and per-sample N counts over
//     the swept words complete the identity (NP = 4):
//         4 * matches = W + T + 3 (cnt_i + cnt_j),   T = sum_s (u u' + v v' + w w' - 3 n n')
//     (a site where either sample is N contributes 1 to W, 0 to the three planes, and 3 via the counts, minus 3 if both
//     are N: always 4 = one match, as the reference's "N matches everything", src/pairsnp.hpp:107-199, 398-403).
//     Exact in int32; ambiguity codes with two or three bases still rule the kernel out (ingest flag).
//   * Sixteen producer warps (two threads per operand row, one per 16-byte K chunk) instead of eight, six word stages
//     in three groups of two: expansion of a group now takes less time than its MMAs, so the tensor pipe is the limit.

namespace tracs {

constexpr int TC2_GROUPS = 3;                      // groups of two word-stages in flight
constexpr int TC2_PRODUCERS = 512;                 // 16 warps: (operand side, row, K chunk)
constexpr int TC2_THREADS = TC2_PRODUCERS + 32;
template <int NP> struct Tc2Geom {
  static constexpr uint32_t SIDE_BYTES = NP * 2 * (TILE / 8) * 128;   // one operand, one word: NP planes x 2 chunks x 2 KB
  static constexpr uint32_t STAGE_BYTES = 2 * SIDE_BYTES;
  static constexpr size_t SMEM = (size_t)TC2_GROUPS * 2 * STAGE_BYTES + 1024;
};

// per-sample count of N sites (mask 1111) over the first `words` words of the word-major planes: the cnt_i of the identity
__global__ void k_tc_ncount(const uint4 *__restrict__ planes, uint32_t Npad, uint32_t words, uint32_t *__restrict__ ncnt) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= Npad) return;
  uint32_t c = 0;
  for (uint32_t w = 0; w < words; ++w) {
    const uint4 x = __ldg(planes + (size_t)w * Npad + s);
    c += __popc(x.x & x.y & x.z & x.w);
  }
  ncnt[s] = c;
}

template <int NP>
__global__ void __launch_bounds__(TC2_THREADS, 1) k_sweep_tc2(const SweepArgs a) {
  using G = Tc2Geom<NP>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *stage_base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[TC2_GROUPS], empty[TC2_GROUPS], tmem_full, tmem_empty;
  __shared__ uint32_t tmem_slot;

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t MMA_WARP = TC2_PRODUCERS / 32;
  if (tid == 0) {
    for (int s = 0; s < TC2_GROUPS; ++s) {
      mbar_init(&full[s], TC2_PRODUCERS / 32);  // one arrival per producer warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full, 1);
    mbar_init(&tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  const uint32_t nw = a.Wp;  // words per tile (a multiple of 8)
  uint32_t gi_run = 0;       // running group counter (group slot = gi_run % TC2_GROUPS)
  uint32_t tile_iter = 0;

  uint2 rc_next = blockIdx.x < a.n_tiles ? __ldg(a.tile_table + blockIdx.x) : make_uint2(0, 0);
  for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tile_iter) {
    const uint32_t rb = rc_next.x, cb = rc_next.y;
    if (tile + gridDim.x < a.n_tiles) rc_next = __ldg(a.tile_table + tile + gridDim.x);

    if (warp < MMA_WARP) {
      // ===== producers: bit-planes -> +-1 int8 operands in the canonical K-major no-swizzle UMMA layout =====
      // lanes 0-15 write K chunk 0 of rows r .. r+15, lanes 16-31 chunk 1: every quarter-warp stores 128 contiguous bytes
      const uint32_t side = warp >> 3, r = (warp & 7u) * 16u + (lane & 15u), h = lane >> 4;
      const uint4 *src = a.planes + (size_t)(side ? cb : rb) * TILE + r;
      const uint32_t row_off = (r >> 3) * 128 + (r & 7) * 16;
      const uint32_t nmul = side ? 1u : 0xFDu;  // N plane: -3 on the row operand, +1 on the column operand
      auto expand = [&](const uint4 &x, uint32_t slot) {
        uint8_t *dst = stage_base + (size_t)slot * G::STAGE_BYTES + side * G::SIDE_BYTES + row_off;
        const uint32_t hi = x.z | x.w, lo = x.y | x.w;   // A = 00, C = 01, G = 10, T = 11
        const uint32_t pl[3] = {hi, lo, ~(hi ^ lo)};
        const uint32_t nb = NP == 4 ? (x.x & x.y & x.z & x.w) : 0u;
        uint32_t nm[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) nm[q] = ((nb >> (4 * h + q)) & 0x01010101u) * 0xFFu;  // 0xFF in the bytes of N sites
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          uint4 o;
          uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t b01 = (pl[p] >> (4 * h + q)) & 0x01010101u;
            uint32_t v = (b01 * 0xFEu) ^ 0xFFFFFFFFu;     // bit 1 -> 0x01 (+1), bit 0 -> 0xFF (-1)
            if (NP == 4) v &= ~nm[q];                     // N -> 0
            ow[q] = v;
          }
          *reinterpret_cast<uint4 *>(dst + (size_t)(2 * p + h) * ((TILE / 8) * 128)) = o;
        }
        if (NP == 4) {
          uint4 o;
          o.x = (nm[0] & 0x01010101u) * nmul; o.y = (nm[1] & 0x01010101u) * nmul;
          o.z = (nm[2] & 0x01010101u) * nmul; o.w = (nm[3] & 0x01010101u) * nmul;
          *reinterpret_cast<uint4 *>(dst + (size_t)(2 * 3 + h) * ((TILE / 8) * 128)) = o;
        }
      };
      // The panels of a full-length tile are tens of MB each and stream from HBM: two groups are held in registers
      // ahead of the one being expanded and the lines of a group 16 further on are requested into L2 (the first
      // version kept one group in flight and ran at DRAM latency per group, not at the MMA rate).
      constexpr uint32_t PFD = 32;  // words between the L2 prefetch and the use
      uint4 c0 = __ldg(src), c1 = __ldg(src + (size_t)a.Npad);
      uint4 e0 = c0, e1 = c1;
      if (nw > 2) {
        e0 = __ldg(src + (size_t)2 * a.Npad);
        e1 = __ldg(src + (size_t)3 * a.Npad);
      }
      for (uint32_t w = 0; w < nw; w += 2, ++gi_run) {
        const uint4 x0 = c0, x1 = c1;
        c0 = e0;
        c1 = e1;
        if (w + 4 < nw) {
          e0 = __ldg(src + (size_t)(w + 4) * a.Npad);
          e1 = __ldg(src + (size_t)(w + 5) * a.Npad);
        }
        if (w + PFD < nw && (lane & 7u) == 0) {  // one request per 128-byte line (8 rows x 16 B)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(w + PFD) * a.Npad));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(w + PFD + 1) * a.Npad));
        }
        const uint32_t g = gi_run % TC2_GROUPS;
        mbar_wait(&empty[g], ((gi_run / TC2_GROUPS) & 1u) ^ 1u);
        expand(x0, 2 * g);
        expand(x1, 2 * g + 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[g]);
      }
    } else {
      // ===== MMA issuer: one elected thread =====
      if (lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
        const uint32_t k_stride = (TILE / 8) * 128, m_stride = 128;
        mbar_wait(&tmem_empty, (tile_iter & 1u) ^ 1u);  // epilogue of the previous tile has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t desc0 = umma_desc(smem_u32(stage_base), k_stride, m_stride);
        for (uint32_t w = 0; w < nw; w += 2, ++gi_run) {
          const uint32_t g = gi_run % TC2_GROUPS;
          mbar_wait(&full[g], (gi_run / TC2_GROUPS) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dg = desc0 + (uint64_t)((2 * g * G::STAGE_BYTES) >> 4);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const uint64_t da = dg + (uint64_t)((q * G::STAGE_BYTES + p * 2 * k_stride) >> 4);
              const uint64_t db = da + (uint64_t)(G::SIDE_BYTES >> 4);
              const uint32_t acc = (q | p) != 0 ? 1u : (uint32_t)(w != 0u);
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                  "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
                  : "memory");
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&empty[g])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&tmem_full)) : "memory");
      } else {
        gi_run += nw / 2;
      }
      __syncwarp();
    }

    if (warp < 4) {
      // ===== epilogue: TMEM -> registers -> threshold -> append =====
      mbar_wait(&tmem_full, tile_iter & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int32_t W3 = 3 * (int32_t)(a.Wp * 32u);
      const uint32_t gi = rb * TILE + warp * 32 + lane;  // TMEM lane = tile row
      const int32_t ci = NP == 4 ? 3 * (int32_t)__ldg(a.tc_ncnt + gi) : 0;
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
        // d = W - matches = (3 W - T - 3 (cnt_i + cnt_j)) / 4 (exact)
        const int32_t cj_mine = NP == 4 ? 3 * (int32_t)__ldg(a.tc_ncnt + cb * TILE + c0 + lane) : 0;  // lane j holds column j's count
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
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

}  // namespace tracs
