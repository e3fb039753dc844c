// Single-pass ingest (included in sweep.cu after k_pack / k_gather: same translation unit).
//
// The two-kernel ingest reads the ASCII matrix twice: k_pack streams it once (column AND, N-plane), and
// k_gather then fetches one 64-byte DRAM atom per (sample, variable site) -- 32-37 GB for 0.5 GB of useful
// bytes at C2. k_ingest removes the second DRAM pass: the whole grid walks the alignment in STRIPS of sw
// 32-site words x all samples, narrow enough that three strips fit in L2, and gathers a strip's variable
// sites right after the strip's column AND is complete, while its bytes are still L2-resident.
//
//   item (k, c) = strip k, sample chunk c; items are handed out in order by one atomic counter.
//     1. pack   : the k_pack loop on [chunk samples] x [sw words] -> N-plane, summaries, N counts, column AND
//     2. last item of the strip to finish ("last block" pattern on done[k]) turns the strip's column AND into
//        variable-site indices; base[k] chains the running count from strip k-1 so sites keep their order
//     3. every item of the strip waits for ready[k], then bit-slices ITS OWN samples at the plane words the
//        strip completed (the bytes it read a few microseconds ago)
//   No item blocks before its pack phase and items are dequeued in order, so with at least n_items
//   resident CTAs every wait is on CTAs that are running: no deadlock.
// Results are identical to k_pack + k_siteflags + select + k_gather (same site order, same planes).

namespace tracs {

constexpr int ING_THREADS = 256;
constexpr int ING_CHUNK_MAX = 256;   // samples per item (s_ncnt)
constexpr int ING_SW_MAX = 256;      // words per strip

struct IngestArgs {
  const uint8_t *seqs;
  uint64_t n, L, pitch;
  uint32_t *colmask, *nplane;
  uint64_t npitch;
  uint8_t *nsum;
  uint64_t spitch;
  uint32_t *ncount;
  uint32_t sw, n_strips, n_items, chunk;  // words per strip (32..256, power of two), strips, items per strip, samples per item
  uint32_t *done, *ready, *base;          // [n_strips], [n_strips], [n_strips + 1] (all zero at launch)
  uint32_t *site_idx;
  uint32_t cap_sites;                     // capacity of site_idx / planes in sites (multiple of 32)
  uint4 *planes;
  uint64_t Npad;
  uint32_t *flags;                        // [0] partial ambiguity seen, [1] capacity overflow
  unsigned int *next;
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(ING_THREADS) k_ingest(const IngestArgs a) {
  __shared__ uint8_t slut[256];
  __shared__ uint32_t s_ncnt[ING_CHUNK_MAX];
  __shared__ uint32_t s_cm[ING_SW_MAX * 4];
  __shared__ uint32_t s_scan[ING_THREADS / 32];
  __shared__ uint32_t s_item, s_last, s_b0, s_b1;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  slut[tid] = (uint8_t)base_mask(tid);
  const uint32_t subs = ING_THREADS / a.sw, per = a.chunk / subs;
  const uint32_t wi = tid & (a.sw - 1), sub = tid / a.sw;
  const uint64_t Lw4 = ((a.L + 31) / 32) * 4;  // column-AND words with sites
  const uint32_t total_items = a.n_strips * a.n_items;
  for (;;) {
    __syncthreads();  // shared arrays of the previous item are free
    if (tid == 0) s_item = atomicAdd(a.next, 1u);
    for (uint32_t i = tid; i < a.chunk; i += ING_THREADS) s_ncnt[i] = 0;
    for (uint32_t i = tid; i < a.sw * 4; i += ING_THREADS) s_cm[i] = ~0u;
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= total_items) break;
    const uint32_t k = item / a.n_items, c = item - k * a.n_items;
    const uint64_t cs0 = (uint64_t)c * a.chunk, cs1 = min(a.n, cs0 + a.chunk);  // samples of the item

    // ---- 1. pack -----------------------------------------------------------------------------
    {
      const uint64_t w = (uint64_t)k * a.sw + wi;
      const uint64_t site0 = w * 32;
      const bool in_row = w < a.npitch, has_sites = site0 < a.L;
      const uint64_t sA = min(cs1, cs0 + (uint64_t)sub * per), sB = min(cs1, sA + per);
      uint32_t acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = ~0u;
      const uint32_t nvalid = has_sites ? (uint32_t)min((uint64_t)32, a.L - site0) : 0u;
      const uint32_t validp = pack_validp(nvalid);
      if (in_row && sB > sA)  // warp-uniform: 32 consecutive words of one sample range
        pack_rows<false>(a.seqs + sA * a.pitch + (has_sites ? site0 : 0), a.pitch, (uint32_t)(sB - sA), has_sites,
                         a.nplane + sA * a.npitch + w, a.npitch, a.nsum + sA * a.spitch + (w >> 5), a.spitch,
                         s_ncnt + (sA - cs0), slut, acc, validp, lane);
      if (has_sites) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t cw = pack_colword(acc, j, nvalid);
          if (cw != ~0u) atomicAnd(&s_cm[wi * 4 + j], cw);
        }
      }
      __syncthreads();
      for (uint32_t i = tid; cs0 + i < cs1; i += ING_THREADS)
        if (s_ncnt[i]) atomicAdd(a.ncount + cs0 + i, s_ncnt[i]);
      for (uint32_t i = tid; i < a.sw * 4; i += ING_THREADS) {
        const uint64_t gw = (uint64_t)k * a.sw * 4 + i;
        if (gw < Lw4 && s_cm[i] != ~0u) atomicAnd(a.colmask + gw, s_cm[i]);
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last = (atomicAdd(a.done + k, 1u) == a.n_items - 1) ? 1u : 0u;
      __syncthreads();
    }

    // ---- 2. the last item of the strip turns its column AND into site indices ---------------------
    if (s_last) {
      __threadfence();
      // thread t owns `own` consecutive column words (8 sites each), so that the scan keeps site order
      const uint32_t nw = a.sw * 4, own = (nw + ING_THREADS - 1) / ING_THREADS;
      uint32_t fl[ING_SW_MAX * 4 / ING_THREADS];
      uint32_t cnt = 0;
#pragma unroll
      for (uint32_t r = 0; r < ING_SW_MAX * 4 / ING_THREADS; ++r) {
        fl[r] = 0;
        const uint32_t i = tid * own + r;
        const uint64_t gw = (uint64_t)k * nw + i;
        if (r < own && i < nw && gw < Lw4) {
          const uint32_t m = __ldcg(a.colmask + gw);
          // bit s of fl = nibble s of m is zero (variable site); sites >= L carry non-zero nibbles
          uint32_t z = m | (m >> 1);
          z |= z >> 2;
          z = ~z & 0x11111111u;
          z = (z | (z >> 3)) & 0x03030303u;
          z = (z | (z >> 6)) & 0x000F000Fu;
          fl[r] = (z | (z >> 12)) & 0xFFu;
          cnt += __popc(fl[r]);
        }
      }
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
      }
      if (lane == 31) s_scan[warp] = incl;
      __syncthreads();
      uint32_t woff = 0, tot = 0;
#pragma unroll
      for (int q = 0; q < ING_THREADS / 32; ++q) {
        if ((uint32_t)q < warp) woff += s_scan[q];
        tot += s_scan[q];
      }
      if (tid == 0) {
        if (k > 0)
          while (ld_acquire_u32(a.ready + k - 1) == 0u) __nanosleep(64);
        s_b0 = __ldcg(a.base + k);
      }
      __syncthreads();
      const uint32_t b0 = s_b0;
      uint32_t pos = b0 + woff + incl - cnt;
#pragma unroll
      for (uint32_t r = 0; r < ING_SW_MAX * 4 / ING_THREADS; ++r) {
        uint32_t f = fl[r];
        const uint32_t i = tid * own + r;
        while (f) {
          const uint32_t b = __ffs(f) - 1;
          f &= f - 1;
          if (pos < a.cap_sites) a.site_idx[pos] = (uint32_t)(((uint64_t)k * nw + i) * 8 + b);
          ++pos;
        }
      }
      if (tid == 0) {
        if (b0 + tot > a.cap_sites) a.flags[1] = 1u;
        a.base[k + 1] = b0 + tot;
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_u32(a.ready + k, 1u);
    }

    // ---- 3. gather this item's samples at the plane words the strip completed -----------------------
    if (tid == 0) {
      while (ld_acquire_u32(a.ready + k) == 0u) __nanosleep(64);
      s_b0 = __ldcg(a.base + k);
      s_b1 = __ldcg(a.base + k + 1);
    }
    __syncthreads();
    {
      const uint32_t b0 = s_b0, b1 = min(s_b1, a.cap_sites);
      const uint32_t w_lo = b0 >> 5, w_hi = (k == a.n_strips - 1) ? ((b1 + 31) >> 5) : (b1 >> 5);
      constexpr int GB = 8;  // byte loads in flight per lane
      for (uint32_t w = w_lo; w < w_hi; ++w) {
        const uint32_t v = w * 32 + lane;
        const bool live = v < b1;
        const uint64_t site = live ? __ldcg(a.site_idx + v) : 0;
        // warp q takes samples cs0 + q, cs0 + q + 8, ... of the item
        for (uint64_t sb = cs0 + warp; sb < cs1; sb += (uint64_t)GB * (ING_THREADS / 32)) {
          uint8_t ch[GB];
#pragma unroll
          for (int t = 0; t < GB; ++t) {
            const uint64_t s = sb + (uint64_t)t * (ING_THREADS / 32);
            ch[t] = (live && s < cs1) ? __ldcs(a.seqs + s * a.pitch + site) : (uint8_t)'N';
          }
#pragma unroll
          for (int t = 0; t < GB; ++t) {
            const uint64_t s = sb + (uint64_t)t * (ING_THREADS / 32);
            if (s >= cs1) break;
            const uint32_t m = live ? slut[ch[t]] : 15u;
            if (__any_sync(0xFFFFFFFFu, m != 15u && (m & (m - 1)) != 0u) && lane == 0) a.flags[0] = 1u;
            const uint32_t A = __ballot_sync(0xFFFFFFFFu, m & 1);
            const uint32_t C = __ballot_sync(0xFFFFFFFFu, m & 2);
            const uint32_t G = __ballot_sync(0xFFFFFFFFu, m & 4);
            const uint32_t T = __ballot_sync(0xFFFFFFFFu, m & 8);
            if (lane == 0) a.planes[(uint64_t)w * a.Npad + s] = make_uint4(A, C, G, T);
          }
        }
      }
    }
  }
}

// planesT[s][w] = planes[w][s]  (sample-major copy for the per-pair kernels)
__global__ void __launch_bounds__(256) k_planes_transpose(const uint4 *__restrict__ planes, uint64_t Npad, uint64_t n, uint32_t Wp,
                                                          uint4 *__restrict__ planesT) {
  __shared__ uint4 tile[32][33];
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const uint64_t s0 = (uint64_t)blockIdx.x * 32;
  const uint32_t w0 = blockIdx.y * 32;
  for (uint32_t r = ty; r < 32; r += 8) {
    const uint32_t w = w0 + r;
    const uint64_t s = s0 + tx;
    if (w < Wp && s < n) tile[r][tx] = planes[(uint64_t)w * Npad + s];
  }
  __syncthreads();
  for (uint32_t r = ty; r < 32; r += 8) {
    const uint64_t s = s0 + r;
    const uint32_t w = w0 + tx;
    if (w < Wp && s < n) planesT[s * Wp + w] = tile[tx][r];
  }
}

}  // namespace tracs
