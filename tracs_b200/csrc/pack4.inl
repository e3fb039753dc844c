// Ingest of 4-bit packed alignments (included by sweep.cu, same translation unit as k_pack / k_pack_x).
//
// Packed format: nib[n][pitch4] bytes, site s of a sample in byte s >> 1, bits [4 * (s & 1), +4) -- i.e. site
// s of a 32-bit word sits in nibble s % 8 -- and the nibble IS the base mask of src/pairsnp.hpp:107-199
// (bit0 = A, bit1 = C, bit2 = G, bit3 = T; N / gap / anything else = 1111). It is half the bytes of the ASCII
// matrix, needs no table lookup, and is the layout of the column mask itself, so the pack pass is pure
// streaming: 16 bytes per thread and sample, 4 ANDs for the column mask, 4 x 4 logic ops for the is-N word.
// The 100 000 x 2 Mb alignment of BASELINE.json configs[2] is 100 GB in this format and fits one B200.
//
//   k_encode        ASCII chunk -> packed rows (used by the host-streaming entry point)
//   k_pack4<false>  column AND + N-plane + N counts + block summaries                     [HBM bound]
//   k_pack4<true>   the same + early extraction of the listed sites into X (see k_pack_x)
//
// N-plane bit order of a word in this family: bit 4k + q  <->  site 8q + k (q = register of the uint4,
// k = nibble). Like the ASCII family's order it is a fixed permutation and only ever consumed by
// population counts of ANDs between rows produced by the same family.

namespace tracs {

constexpr int P4_BATCH = 8;   // samples per trip: 8 x 16 B in flight per thread
constexpr int P4_ROUNDS = 3;  // extraction items per lane held in registers (a warp rarely lists more than 12 sites)

// bit 4k + 3 of the result = nibble k of v is 1111: adding 1 to the low three bits carries into bit 3 iff they are all
// set (three operations per register; the shift-and-AND form needs four, and this kernel's ALU pipe is two thirds busy)
__device__ __forceinline__ uint32_t nib_is_n_hi(uint32_t v) { return ((v & 0x77777777u) + 0x11111111u) & v & 0x88888888u; }
__device__ __forceinline__ uint32_t p4_isn(const uint4 &v) {
  return (nib_is_n_hi(v.x) >> 3) | (nib_is_n_hi(v.y) >> 2) | (nib_is_n_hi(v.z) >> 1) | nib_is_n_hi(v.w);
}
// valid sites of a word, in this family's N-plane bit order
__device__ __forceinline__ uint32_t p4_validp(uint32_t nvalid) {
  if (nvalid >= 32u) return 0xFFFFFFFFu;
  uint32_t m = 0;
  for (uint32_t s = 0; s < nvalid; ++s) m |= 1u << (4u * (s & 7u) + (s >> 3));
  return m;
}
// column-AND word q of a 32-site word (sites 8q .. 8q+7): sites >= L must not look variable
__device__ __forceinline__ uint32_t p4_colword(uint32_t a, int q, uint32_t nvalid) {
  const uint32_t v = nvalid > (uint32_t)q * 8u ? min(8u, nvalid - (uint32_t)q * 8u) : 0u;
  if (v < 8u) a |= (v == 0u ? 0xFFFFFFFFu : (0xFFFFFFFFu << (4u * v)));
  return a;
}

// eight ASCII bytes (two registers) -> eight mask nibbles in site order
__device__ __forceinline__ uint32_t encode8(uint32_t w1, uint32_t w2, const uint8_t *slut) {
  uint32_t A, B;  // masks of sites (0, 4, 1, 5) and (2, 6, 3, 7), one per byte (+ flag bits above bit 3)
  pack_unit<0, 0>(w1, w2, slut, A, B);
  return (A & 0xFu) | (((A >> 16) & 0xFu) << 4) | ((B & 0xFu) << 8) | (((B >> 16) & 0xFu) << 12) | (((A >> 8) & 0xFu) << 16) |
         (((A >> 24) & 0xFu) << 20) | (((B >> 8) & 0xFu) << 24) | (((B >> 24) & 0xFu) << 28);
}

// ASCII rows [0, rows) of a staging buffer (pitch % 32 == 0, >= L rounded up to 32) -> packed rows. One thread per
// 32-site word; sites >= L become 1111.
__global__ void __launch_bounds__(256)
k_encode(const uint8_t *__restrict__ ascii, uint64_t rows, uint64_t L, uint64_t pitch, uint8_t *__restrict__ nib, uint64_t pitch4) {
  __shared__ uint8_t slut[256];
  slut[threadIdx.x] = (uint8_t)base_mask(threadIdx.x);
  __syncthreads();
  const uint64_t wpr = pitch4 / 16;  // words per packed row
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * wpr) return;
  const uint64_t s = gid / wpr, w = gid - s * wpr, site0 = w * 32;
  uint4 o = make_uint4(~0u, ~0u, ~0u, ~0u);
  if (site0 < L) {
    const uint4 *src = reinterpret_cast<const uint4 *>(ascii + s * pitch + site0);
    const uint4 a = __ldcs(src), b = __ldcs(src + 1);
    o = make_uint4(encode8(a.x, a.y, slut), encode8(a.z, a.w, slut), encode8(b.x, b.y, slut), encode8(b.z, b.w, slut));
    const uint32_t nvalid = (uint32_t)min((uint64_t)32, L - site0);
    o.x = p4_colword(o.x, 0, nvalid); o.y = p4_colword(o.y, 1, nvalid);
    o.z = p4_colword(o.z, 2, nvalid); o.w = p4_colword(o.w, 3, nvalid);
  }
  __stcs(reinterpret_cast<uint4 *>(nib + s * pitch4 + w * 16), o);
}

template <bool EXTRACT, bool SPARSE_N = false>
__global__ void __launch_bounds__(PACK_THREADS, 3)
k_pack4(const uint8_t *__restrict__ nib, uint64_t s_begin, uint64_t s_end, uint64_t L, uint64_t pitch4, uint32_t *__restrict__ colmask,
        uint32_t *__restrict__ nplane, uint64_t npitch /*words*/, uint8_t *__restrict__ nsum, uint64_t spitch /*bytes*/,
        uint32_t *__restrict__ ncount, const uint32_t *__restrict__ elist, uint32_t VE, uint8_t *__restrict__ X, uint64_t XP) {
  // per warp and sample of the batch: the 512 bytes (1024 sites) the warp has just loaded
  __shared__ __align__(16) uint8_t wbuf[EXTRACT ? PACK_THREADS / 32 : 1][EXTRACT ? P4_BATCH : 1][EXTRACT ? 512 : 16];
  __shared__ uint32_t s_ncnt[PACK_SCHUNK];
  for (int i = threadIdx.x; i < PACK_SCHUNK; i += PACK_THREADS) s_ncnt[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t w = (uint64_t)blockIdx.x * PACK_THREADS + threadIdx.x;  // word index
  const uint64_t s0 = s_begin + (uint64_t)blockIdx.y * PACK_SCHUNK;
  const uint64_t s1 = min(s_end, s0 + PACK_SCHUNK);
  const uint64_t site0 = w * 32;
  const bool in_row = w < npitch;  // npitch is a multiple of 32 words: warp-uniform
  const bool has_sites = site0 < L;
  uint4 acc = make_uint4(~0u, ~0u, ~0u, ~0u);
  const uint32_t nvalid = has_sites ? (uint32_t)min((uint64_t)32, L - site0) : 0u;
  const uint32_t validp = p4_validp(nvalid);
  // listed sites inside this warp's 1024 sites: elist[e_lo .. e_lo + nE)
  const uint64_t wsite0 = (w - lane) * 32;
  uint32_t e_lo = 0, nE = 0;
  if (EXTRACT) {
    auto lower = [&](uint64_t key) {
      uint32_t a = 0, b = VE;
      while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if (__ldg(elist + mid) < key) a = mid + 1; else b = mid;
      }
      return a;
    };
    e_lo = lower(wsite0);
    nE = lower(wsite0 + 1024) - e_lo;
  }
  // work item i of a batch = (sample t = i / nE, listed site e = i % nE). Its nibble sits in the warp's slot at byte
  // t * 512 + (o >> 5) * 16 + ((o & 31) >> 1), high nibble if o is odd (bit 12 of the item word); o = site - wsite0
  auto slot = [&](uint32_t t, uint32_t e) {
    const uint32_t o = __ldg(elist + e_lo + e) - (uint32_t)wsite0;
    return (t * 512u + (o >> 5) * 16u + ((o & 31u) >> 1)) | ((o & 1u) << 12);
  };
  const uint32_t items = nE * P4_BATCH;
  uint32_t it_a[P4_ROUNDS], it_x[P4_ROUNDS];
#pragma unroll
  for (int r = 0; r < P4_ROUNDS; ++r) {
    const uint32_t i = lane + 32u * r;
    it_a[r] = 0xFFFFFFFFu;
    it_x[r] = 0;
    if (EXTRACT && i < items) {
      const uint32_t t = i / nE, e = i - t * nE;
      it_a[r] = slot(t, e);
      it_x[r] = t * (uint32_t)XP + e;  // XP <= L / 16 < 2^27, t < 8
    }
  }
  uint8_t *mine = &wbuf[EXTRACT ? warp : 0][0][EXTRACT ? lane * 16 : 0];
  const uint8_t *wb = &wbuf[EXTRACT ? warp : 0][0][0];
  if (in_row) {  // warp-uniform
    const uint4 kN = make_uint4(~0u, ~0u, ~0u, ~0u);
    // lanes past the end of the alignment read the start of the row instead: whatever they see is masked out
    // (validp == 0, no column word, never listed), so the loads need no per-lane predicate
    const uint8_t *src = nib + s0 * pitch4 + (has_sites ? w * 16 : 0);
    uint32_t *np = nplane + s0 * npitch + w;
    uint8_t *sp = nsum + s0 * spitch + (w >> 5);
    uint32_t *cnt = s_ncnt;
    uint8_t *xrow = EXTRACT ? X + s0 * XP + e_lo : nullptr;
    const bool more_items = items > 32u * P4_ROUNDS;  // warp-uniform
    auto batch = [&](auto full_tag, uint32_t rows, bool prefetch_next) {
      constexpr bool FULL = decltype(full_tag)::value;
      uint4 va[P4_BATCH];
#pragma unroll
      for (int t = 0; t < P4_BATCH; ++t)
        va[t] = (FULL || (uint32_t)t < rows) ? __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)t * pitch4)) : kN;
      if (prefetch_next) {  // the next batch on its way into L2 while this one is handled
#pragma unroll
        for (int t = 0; t < P4_BATCH; ++t) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(P4_BATCH + t) * pitch4));
      }
#pragma unroll
      for (int t = 0; t < P4_BATCH; ++t) {
        if (EXTRACT && nE) *reinterpret_cast<uint4 *>(mine + t * 512) = va[t];  // warp-uniform condition
        if (FULL || (uint32_t)t < rows) {
          acc.x &= va[t].x; acc.y &= va[t].y; acc.z &= va[t].z; acc.w &= va[t].w;
          pack_emit_n<SPARSE_N>(p4_isn(va[t]) & validp, np + (size_t)t * npitch, sp + (size_t)t * spitch, cnt + t, lane);
        }
      }
      if (EXTRACT && nE) {  // warp-uniform
        __syncwarp();
#pragma unroll
        for (int r = 0; r < P4_ROUNDS; ++r) {  // the lookups run unconditionally (slot 0xFFF for idle lanes), only the store is predicated
          const uint8_t m = (uint8_t)((wb[it_a[r] & 0xFFFu] >> ((it_a[r] >> 10) & 4u)) & 15u);
          if (FULL ? it_a[r] != 0xFFFFFFFFu : (it_a[r] != 0xFFFFFFFFu && ((it_a[r] & 0xFFFu) >> 9) < rows)) xrow[it_x[r]] = m;
        }
        if (more_items)
          for (uint32_t i = lane + 32u * P4_ROUNDS; i < items; i += 32) {
            const uint32_t t = i / nE, e = i - t * nE;
            if (t < rows) {
              const uint32_t a = slot(t, e);
              xrow[(size_t)t * XP + e] = (uint8_t)((wb[a & 0xFFFu] >> ((a >> 10) & 4u)) & 15u);
            }
          }
        __syncwarp();
      }
      src += (size_t)P4_BATCH * pitch4;
      np += (size_t)P4_BATCH * npitch;
      sp += (size_t)P4_BATCH * spitch;
      cnt += P4_BATCH;
      if (EXTRACT) xrow += (size_t)P4_BATCH * XP;
    };
    uint64_t b0 = s0;
    for (; b0 + 2 * P4_BATCH <= s1; b0 += P4_BATCH) batch(std::true_type{}, P4_BATCH, true);
    for (; b0 + P4_BATCH <= s1; b0 += P4_BATCH) batch(std::true_type{}, P4_BATCH, false);
    if (b0 < s1) batch(std::false_type{}, (uint32_t)(s1 - b0), false);
  }
  __syncthreads();
  for (uint64_t i = threadIdx.x; i < s1 - s0; i += PACK_THREADS)
    if (s_ncnt[i]) atomicAdd(ncount + s0 + i, s_ncnt[i]);
  if (has_sites) {
    uint32_t *cm = colmask + w * 4;
    const uint32_t cw[4] = {p4_colword(acc.x, 0, nvalid), p4_colword(acc.y, 1, nvalid), p4_colword(acc.z, 2, nvalid),
                            p4_colword(acc.w, 3, nvalid)};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cw[j] != ~0u) atomicAnd(cm + j, cw[j]);
  }
}

}  // namespace tracs
