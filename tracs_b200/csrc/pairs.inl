// Evaluation of the prefilter's candidate pairs: full-length mismatch count d(i,j) over the ingested words and
// |N_i u N_j| (for the compared-site count), included by sweep.cu after `Ingested`.
//
// Candidates are not random pairs: they are the edges of a graph whose connected components are (nearly) the
// transmission clusters, i.e. small and dense. The per-candidate kernels (one warp per pair: k_pairs_sparse)
// stream two full rows per pair and re-read every row once per partner; on the benchmark shapes they were
// latency-bound and 25 % of the step. Here the candidate graph is cut into its connected components first;
// a component that is dense (m(m-1)/2 <= 4 x its candidates) is evaluated as a BLOCK:
//   k_block_d   64 x 64 member blocks, 4 x 4 register micro-tiles over the sample-major planes staged with
//               cp.async: every row of a component is read once per block row, not once per partner; diagonal
//               blocks only run the micro-tiles of the upper triangle                        [INT-pipe work]
//   k_block_n   lane <-> 128-site block position; the members' block summaries are transposed across the warp so
//               that a lane visits only the members with an N in ITS block, four loads in flight; pairs are
//               formed only where two members really share a site                            [L2 latency]
// and the candidates read their values out of the per-component matrices (k_pairs_gather). Components that
// are large and sparse (chains, hubs) keep the per-candidate kernel. Same quantities as src/pairsnp.hpp:398-403
// (d) and :417-419 (N union); every path is compared with the oracle in tests/.

namespace tracs {

constexpr uint32_t PP_MAXM = 8192;      // largest component evaluated as blocks
constexpr int PP_KC = 8;                // words per staged chunk (Wp is a multiple of KC == 8)

__global__ void k_pp_init(uint32_t *parent, uint32_t *iota, uint32_t n) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) {
    parent[v] = v;
    iota[v] = v;
  }
}
__device__ __forceinline__ uint32_t pp_find(uint32_t *parent, uint32_t v) {
  uint32_t p = parent[v];
  while (p != v) {  // path halving
    const uint32_t gp = parent[p];
    if (gp != p) parent[v] = gp;
    v = p;
    p = gp;
  }
  return v;
}
// union-find over the candidate edges: the larger root is hooked under the smaller one, so the root of a finished
// component is its smallest sample (deterministic whatever the interleaving)
__global__ void k_pp_hook(const uint64_t *__restrict__ keys, uint64_t E, uint32_t *parent) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint64_t k = keys[e];
  uint32_t ru = pp_find(parent, (uint32_t)(k >> 32)), rv = pp_find(parent, (uint32_t)k);
  while (ru != rv) {
    const uint32_t hi = max(ru, rv), lo = min(ru, rv);
    const uint32_t old = atomicCAS(parent + hi, hi, lo);
    if (old == hi) break;
    ru = pp_find(parent, old);
    rv = pp_find(parent, lo);
  }
}
__global__ void k_pp_flatten(uint32_t *parent, uint32_t n, uint32_t *size) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  uint32_t r = v;
  while (parent[r] != r) r = parent[r];
  parent[v] = r;
  atomicAdd(size + r, 1u);
}
__global__ void k_pp_count(const uint64_t *__restrict__ keys, uint64_t E, const uint32_t *__restrict__ root, uint32_t *ecnt) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) atomicAdd(ecnt + root[keys[e] >> 32], 1u);
}
// per root: is the component evaluated as blocks; how many 64 x 64 block tasks and matrix entries it needs
__global__ void k_pp_decide(const uint32_t *__restrict__ root, const uint32_t *__restrict__ size, const uint32_t *__restrict__ ecnt, uint32_t n,
                            uint8_t *__restrict__ dense, uint32_t *__restrict__ ntasks, uint64_t *__restrict__ msq,
                            uint32_t *__restrict__ dense_list, uint32_t *n_dense) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n) return;
  if (s == n) {  // sentinel entries of the exclusive scans
    ntasks[n] = 0;
    msq[n] = 0;
    return;
  }
  const uint32_t m = size[s];
  const bool is_dense = root[s] == s && m >= 2 && m <= PP_MAXM && (uint64_t)m * (m - 1) / 2 <= 4ull * ecnt[s];
  dense[s] = is_dense ? 1 : 0;
  const uint32_t nb = (m + 63) / 64;
  ntasks[s] = is_dense ? nb * (nb + 1) / 2 : 0u;
  msq[s] = is_dense ? (uint64_t)m * m : 0ull;
  if (is_dense) dense_list[atomicAdd(n_dense, 1u)] = s;
}
__global__ void k_pp_heads(const uint32_t *__restrict__ sroot, uint32_t n, uint32_t *__restrict__ comp_start) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n && (p == 0 || sroot[p] != sroot[p - 1])) comp_start[sroot[p]] = p;
}
__global__ void k_pp_rank(const uint32_t *__restrict__ sroot, const uint32_t *__restrict__ members, uint32_t n,
                          const uint32_t *__restrict__ comp_start, uint32_t *__restrict__ rank) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) rank[members[p]] = p - comp_start[sroot[p]];
}
__global__ void k_pp_tasks(const uint32_t *__restrict__ ntasks, const uint32_t *__restrict__ task_off, const uint32_t *__restrict__ size,
                           uint32_t n, uint2 *__restrict__ tasks) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || ntasks[s] == 0) return;
  const uint32_t nb = (size[s] + 63) / 64;
  uint32_t o = task_off[s];
  for (uint32_t br = 0; br < nb; ++br)
    for (uint32_t bc = br; bc < nb; ++bc) tasks[o++] = make_uint2(s, (br << 16) | bc);
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// shared-memory slot of block row r: rows 4t .. 4t+3 of one micro-tile land 16 slots apart, so that the lanes of a
// warp (consecutive micro-tiles) read consecutive 16-byte slots: conflict-free LDS.128
__device__ __forceinline__ uint32_t pp_slot(uint32_t r) { return (r & 3u) * 16u + (r >> 2); }

constexpr int BD_THREADS = 256;
// MODE 0: rows = sample-major bit-planes (uint4 = A, C, G, T words of 32 sites); result = mismatches = 32 * Wp - matches.
// MODE 1: rows = N-plane rows (uint4 = 128 sites); result = |N_x n N_y| (the dense-N variant of k_block_n: when most
//         128-site blocks of most samples hold an N -- low-coverage metagenomic alignments, BASELINE configs[3] -- the
//         summaries skip nothing and the intersection is a plain AND + POPC contraction over the whole row).
template <int MODE>
__global__ void __launch_bounds__(BD_THREADS, 2)
k_block_d(const uint2 *__restrict__ tasks, const uint32_t *__restrict__ task_off, uint32_t n, const uint32_t *__restrict__ members,
          const uint32_t *__restrict__ comp_start, const uint32_t *__restrict__ size, const uint64_t *__restrict__ sq_off,
          const uint4 *__restrict__ planesT, uint64_t Wp /* uint4 per row */, uint32_t one, uint32_t ksplit,
          uint32_t *__restrict__ scratch_d) {
  __shared__ __align__(16) uint4 sm[2][2][PP_KC * 64];  // [stage][side][kk * 64 + slot]: 32 KB
  const uint32_t tid = threadIdx.x;
  const uint32_t n_tasks = task_off[n];
  const uint32_t nchunks_all = (uint32_t)(Wp / PP_KC);
  // ksplit > 1 (MODE 1 only): the row is cut into `ksplit` slices handled by different CTAs that add their counts up
  // with atomics -- a handful of components over very long rows would otherwise keep only a handful of SMs busy
  for (uint32_t vt = blockIdx.x; vt < n_tasks * ksplit; vt += gridDim.x) {
    const uint32_t task = vt / ksplit, slice = vt - task * ksplit;
    const uint32_t ch_lo = (uint32_t)((uint64_t)nchunks_all * slice / ksplit), ch_hi = (uint32_t)((uint64_t)nchunks_all * (slice + 1) / ksplit);
    const uint32_t nchunks = ch_hi - ch_lo;
    if (nchunks == 0) continue;
    const uint2 t = tasks[task];
    const uint32_t c = t.x, br = t.y >> 16, bc = t.y & 0xFFFFu;
    const uint32_t m = size[c], base = comp_start[c];
    const uint64_t sq = sq_off[c];
    const uint32_t r0 = br * 64, c0 = bc * 64;
    const uint32_t nr = min(64u, m - r0), nc = min(64u, m - c0);
    const bool diag = br == bc;
    // the two (row, word) pairs this thread stages per side and chunk: idx = tid + 256 k -> row = idx / 8, word = idx % 8
    const uint4 *srcA[2], *srcB[2];
    uint32_t dst[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t idx = tid + 256u * k, row = idx >> 3, kk = idx & 7u;
      dst[k] = kk * 64 + pp_slot(row);
      srcA[k] = row < nr ? planesT + (size_t)members[base + r0 + row] * Wp + kk + (size_t)ch_lo * PP_KC : nullptr;
      srcB[k] = (!diag && row < nc) ? planesT + (size_t)members[base + c0 + row] * Wp + kk + (size_t)ch_lo * PP_KC : nullptr;
    }
    auto stage = [&](uint32_t ch, uint32_t s) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (srcA[k]) cp_async16(&sm[s][0][dst[k]], srcA[k] + (size_t)ch * PP_KC);
        if (srcB[k]) cp_async16(&sm[s][1][dst[k]], srcB[k] + (size_t)ch * PP_KC);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // micro-tile of this thread: 4 rows x 4 columns; diagonal blocks only enumerate the upper triangle
    const uint32_t nt_r = (nr + 3) / 4, nt_c = (nc + 3) / 4;
    uint32_t ty = 0, tx = 0;
    bool active;
    if (diag) {
      uint32_t rem = tid;
      while (ty < nt_r && rem >= nt_r - ty) {
        rem -= nt_r - ty;
        ++ty;
      }
      active = ty < nt_r;
      tx = ty + rem;
    } else {
      active = tid < nt_r * nt_c;
      ty = tid / nt_c;
      tx = tid - ty * nt_c;
    }
    uint32_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    stage(0, 0);
    for (uint32_t ch = 0; ch < nchunks; ++ch) {
      if (ch + 1 < nchunks) {
        stage(ch + 1, (ch + 1) & 1u);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      if (active) {
        const uint4 *pr = &sm[ch & 1u][0][ty];
        const uint4 *pc = &sm[ch & 1u][diag ? 0 : 1][tx];
#pragma unroll
        for (int kk = 0; kk < PP_KC; ++kk) {
          uint4 r[4], cv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) r[i] = pr[kk * 64 + i * 16];
#pragma unroll
          for (int j = 0; j < 4; ++j) cv[j] = pc[kk * 64 + j * 16];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (MODE == 0) {
                const uint32_t mt = (r[i].x & cv[j].x) | (r[i].y & cv[j].y) | (r[i].z & cv[j].z) | (r[i].w & cv[j].w);
                acc[i][j] = __popc(mt) * one + acc[i][j];
              } else {
                acc[i][j] = (__popc(r[i].x & cv[j].x) + __popc(r[i].y & cv[j].y)) * one + acc[i][j];
                acc[i][j] = (__popc(r[i].z & cv[j].z) + __popc(r[i].w & cv[j].w)) * one + acc[i][j];
              }
            }
        }
      }
      __syncthreads();
    }
    if (active) {
      const uint32_t total_bits = (uint32_t)Wp * 32u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a = r0 + 4 * ty + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t b = c0 + 4 * tx + j;
          if (a < b && b < m) {
            if (MODE == 1 && ksplit > 1) {
              if (acc[i][j]) atomicAdd(scratch_d + sq + (uint64_t)a * m + b, acc[i][j]);
            } else {
              scratch_d[sq + (uint64_t)a * m + b] = MODE == 0 ? total_bits - acc[i][j] : acc[i][j];
            }
          }
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t popc4(const uint4 &v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
__device__ __forceinline__ bool any4(const uint4 &v) { return (v.x | v.y | v.z | v.w) != 0u; }
__device__ __forceinline__ uint4 and4(const uint4 &a, const uint4 &b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }

// 32 x 32 bit-matrix transpose across a warp: afterwards bit k of lane b = bit b of lane k's input
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t mlo = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
    x = (lane & j) ? ((x & ~mlo) | ((y & ~mlo) >> j)) : ((x & mlo) | ((y & mlo) << j));
  }
  return x;
}

// |N_x n N_y| for all member pairs of the dense components. One warp per (component, group q of 32 block positions);
// lane <-> one 128-site block. The block summaries of 32 members at a time are transposed across the warp, so each
// lane knows WHICH members have an N in its block (a handful) and visits only those, four loads in flight at a time.
// Pairs inside a group of four are intersected directly; a member that shares a site with an earlier group
// (`once` accumulator) walks back over the earlier members of its block (rare).
__global__ void __launch_bounds__(256, 4)
k_block_n(const uint32_t *__restrict__ dense_list, const uint32_t *__restrict__ n_dense_p, const uint32_t *__restrict__ members,
          const uint32_t *__restrict__ comp_start, const uint32_t *__restrict__ size, const uint64_t *__restrict__ sq_off,
          const uint32_t *__restrict__ nplane, uint64_t npitch, const uint8_t *__restrict__ nsum, uint64_t spitch,
          uint32_t *__restrict__ scratch_i) {
  // per thread: the members (index in the component) met so far with an N in this thread's block; a member that
  // shares a site with an earlier one only has to look at those instead of walking over all earlier members
  constexpr int HIST = 28;
  __shared__ uint16_t s_hist[HIST][256];
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t nq = spitch / 4;
  const uint64_t total = (uint64_t)(*n_dense_p) * nq;
  for (uint64_t item = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += n_warps) {
    uint32_t n_hist = 0;  // entries of s_hist[.][threadIdx.x] in use (members of earlier groups); HIST + 1 = overflowed
    const uint32_t ci = (uint32_t)(item / nq);
    const uint64_t q = item - (uint64_t)ci * nq;
    const uint32_t c = dense_list[ci];
    const uint32_t m = size[c], base = comp_start[c];
    const uint64_t sq = sq_off[c];
    const uint64_t blk = q * 32 + lane;  // this lane's 128-site block
    auto block_of = [&](uint32_t s) { return __ldg(reinterpret_cast<const uint4 *>(nplane + (size_t)s * npitch) + blk); };
    auto word_of = [&](uint32_t s) { return __ldg(reinterpret_cast<const uint32_t *>(nsum + (size_t)s * spitch) + q); };
    uint4 once = make_uint4(0, 0, 0, 0);  // N sites of the members of earlier groups
    for (uint32_t k0 = 0; k0 < m; k0 += 32) {
      const uint32_t kn = min(32u, m - k0);
      const uint32_t my_s = lane < kn ? __ldg(members + base + k0 + lane) : 0u;
      const uint32_t my_w = lane < kn ? word_of(my_s) : 0u;
      if (!__any_sync(0xFFFFFFFFu, my_w != 0u)) continue;
      uint32_t mask = warp_transpose32(my_w, lane);  // bit k: member k0 + k has an N in this lane's block
      while (__any_sync(0xFFFFFFFFu, mask != 0u)) {
        uint4 v[4];
        uint32_t idx[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const bool have = mask != 0u;
          const uint32_t k = have ? (uint32_t)__ffs(mask) - 1u : 0u;
          mask &= mask - 1u;  // (0 stays 0)
          const uint32_t s = __shfl_sync(0xFFFFFFFFu, my_s, k);
          idx[t] = have ? k0 + k : 0xFFFFFFFFu;
          v[t] = have ? block_of(s) : make_uint4(0, 0, 0, 0);
        }
        const uint32_t group_first = idx[0];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = a + 1; b < 4; ++b) {
            const uint32_t cnt = popc4(and4(v[a], v[b]));
            if (cnt) atomicAdd(scratch_i + sq + (uint64_t)idx[a] * m + idx[b], cnt);  // idx[b] valid whenever cnt != 0
          }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (any4(and4(v[t], once))) {  // shares a site with a member of an earlier group: find out which
            if (n_hist <= (uint32_t)HIST) {
              for (uint32_t k = 0; k < n_hist; ++k) {
                const uint32_t x = s_hist[k][threadIdx.x];
                const uint32_t cnt = popc4(and4(block_of(__ldg(members + base + x)), v[t]));
                if (cnt) atomicAdd(scratch_i + sq + (uint64_t)x * m + idx[t], cnt);
              }
            } else {  // history overflowed (a block where very many members have an N): walk over all earlier members
              for (uint32_t x = 0; x < group_first; ++x) {
                const uint32_t sx = __ldg(members + base + x);
                if (!((word_of(sx) >> lane) & 1u)) continue;
                const uint32_t cnt = popc4(and4(block_of(sx), v[t]));
                if (cnt) atomicAdd(scratch_i + sq + (uint64_t)x * m + idx[t], cnt);
              }
            }
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          once.x |= v[t].x; once.y |= v[t].y; once.z |= v[t].z; once.w |= v[t].w;
          if (idx[t] != 0xFFFFFFFFu) {
            if (n_hist < (uint32_t)HIST) s_hist[n_hist][threadIdx.x] = (uint16_t)idx[t];
            n_hist = min(n_hist + 1u, (uint32_t)HIST + 1u);
          }
        }
      }
    }
  }
}

// candidates of dense components read their values out of the component matrices
__global__ void k_pairs_gather(const uint64_t *__restrict__ keys, uint64_t E, const uint32_t *__restrict__ root, const uint8_t *__restrict__ dense,
                               const uint32_t *__restrict__ rank, const uint32_t *__restrict__ size, const uint64_t *__restrict__ sq_off,
                               const uint32_t *__restrict__ scratch_d, const uint32_t *__restrict__ scratch_i,
                               const uint32_t *__restrict__ ncount, uint32_t *__restrict__ d_out, uint32_t *__restrict__ u_out) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint64_t k = keys[e];
  const uint32_t i = (uint32_t)(k >> 32), j = (uint32_t)k;
  const uint32_t c = root[i];
  if (!dense[c]) return;
  const uint64_t idx = sq_off[c] + (uint64_t)rank[i] * size[c] + rank[j];  // i < j and members are in sample order: rank[i] < rank[j]
  if (d_out) d_out[e] = scratch_d[idx];
  if (u_out) u_out[e] = ncount[i] + ncount[j] - scratch_i[idx];
}

// the per-candidate kernel (one warp per pair) for candidates of components that are not evaluated as blocks;
// `dense` == nullptr: every candidate
__global__ void __launch_bounds__(256)
k_pairs_sparse(const uint64_t *__restrict__ keys, uint64_t n_keys, const uint32_t *__restrict__ root, const uint8_t *__restrict__ dense,
               const uint4 *__restrict__ planesT, uint32_t Wp, const uint32_t *__restrict__ nplane, uint64_t npitch,
               const uint8_t *__restrict__ nsum, uint64_t spitch, const uint32_t *__restrict__ ncount, uint32_t *__restrict__ d_out,
               uint32_t *__restrict__ u_out) {
  const uint32_t lane = threadIdx.x & 31;
  // every lane pre-screens one candidate (most belong to components evaluated as blocks), then the warp works through
  // the remaining ones together
  const uint64_t e0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  if (e0 >= n_keys) return;
  const uint64_t mine = e0 + lane;
  const bool todo = mine < n_keys && !(dense && dense[root[keys[mine] >> 32]]);
  uint32_t pending = __ballot_sync(0xFFFFFFFFu, todo);
  while (pending) {
  const uint64_t e = e0 + (uint32_t)(__ffs(pending) - 1);
  pending &= pending - 1;
  const uint64_t k = keys[e];
  const uint64_t i = k >> 32, j = k & 0xFFFFFFFFull;
  const uint4 *ri = planesT + i * Wp, *rj = planesT + j * Wp;
  uint32_t mism = 0;
  if (d_out) {
#pragma unroll 4
    for (uint32_t w = lane; w < Wp; w += 32) {
      const uint4 x = __ldg(ri + w), y = __ldg(rj + w);
      mism += __popc(~((x.x & y.x) | (x.y & y.y) | (x.z & y.z) | (x.w & y.w)));
    }
  }
  uint32_t inter = 0;
  if (u_out) {
    const uint32_t *si = reinterpret_cast<const uint32_t *>(nsum + i * spitch);
    const uint32_t *sj = reinterpret_cast<const uint32_t *>(nsum + j * spitch);
    const uint4 *ni = reinterpret_cast<const uint4 *>(nplane + i * npitch);
    const uint4 *nj = reinterpret_cast<const uint4 *>(nplane + j * npitch);
    for (uint64_t q = lane; q < spitch / 4; q += 32) {
      uint32_t m = __ldg(si + q) & __ldg(sj + q);
      while (m) {
        const uint32_t b = __ffs(m) - 1;
        m &= m - 1;
        const uint4 x = __ldg(ni + q * 32 + b), y = __ldg(nj + q * 32 + b);
        inter += __popc(x.x & y.x) + __popc(x.y & y.y) + __popc(x.z & y.z) + __popc(x.w & y.w);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mism += __shfl_xor_sync(0xFFFFFFFFu, mism, o);
    inter += __shfl_xor_sync(0xFFFFFFFFu, inter, o);
  }
  if (lane == 0) {
    if (d_out) d_out[e] = mism;
    if (u_out) u_out[e] = ncount[i] + ncount[j] - inter;
  }
  }
}

// keys: E candidate keys (i << 32 | j, i < j), any order, no duplicates. d_out[e] = mismatches of the pair over all
// ingested words; u_out[e] = |N_i u N_j| over the ingested sites. Two phases, so that a caller can hand the distances
// on (threshold, likelihood, copies to the host) while the compared-sites part is still running: distances() first,
// then unions(). TRACS_PAIRS=sparse forces the per-candidate kernel for everything (tests compare both).
struct PairEval {
  bool blocks = false;
  uint32_t n = 0;
  uint64_t sq_cap = 0, task_cap = 0;
  int n_sm = 148;
  DevBuf<uint32_t> parent, iota, size, ecnt, ntasks, task_off, dense_list, n_dense, sroot, members, comp_start, rank;
  DevBuf<uint64_t> msq, sq_off;
  DevBuf<uint8_t> dense, tmp;
  DevBuf<uint2> tasks;
  DevBuf<uint32_t> scratch_d, scratch_i;
  static unsigned grid1(uint64_t items) { return (unsigned)((items + 255) / 256); }

  void distances(const Ingested &g, const uint64_t *keys, uint64_t E, uint32_t *d_out, cudaStream_t st) {
    tracs_stats_t &S = g_stats;
    if (E == 0) return;
    n = (uint32_t)g.n;
    const char *mode = getenv("TRACS_PAIRS");
    blocks = !(mode && !strcmp(mode, "sparse")) && E < (1ull << 31);
    if (!blocks) {
      k_pairs_sparse<<<grid1(E), 256, 0, st>>>(keys, E, nullptr, nullptr, g.planesT.p, g.Wp, g.nplane.p, g.npitch, g.nsum.p, g.spitch,
                                                  g.ncount.p, d_out, nullptr);
      S.kernel_launches++;
      TRACS_CK(cudaGetLastError());
      return;
    }
    // ---- components of the candidate graph --------------------------------------------------------------------
    parent.alloc(n); iota.alloc(n); size.alloc(n); ecnt.alloc(n); ntasks.alloc(n + 1); task_off.alloc(n + 1); dense_list.alloc(n);
    n_dense.alloc(1); sroot.alloc(n); members.alloc(n); comp_start.alloc(n); rank.alloc(n);
    msq.alloc(n + 1); sq_off.alloc(n + 1);
    dense.alloc(n);
    TRACS_CK(cudaMemsetAsync(size.p, 0, n * sizeof(uint32_t), st));
    TRACS_CK(cudaMemsetAsync(ecnt.p, 0, n * sizeof(uint32_t), st));
    TRACS_CK(cudaMemsetAsync(n_dense.p, 0, sizeof(uint32_t), st));
    k_pp_init<<<grid1(n), 256, 0, st>>>(parent.p, iota.p, n);
    k_pp_hook<<<grid1(E), 256, 0, st>>>(keys, E, parent.p);
    k_pp_flatten<<<grid1(n), 256, 0, st>>>(parent.p, n, size.p);
    k_pp_count<<<grid1(E), 256, 0, st>>>(keys, E, parent.p, ecnt.p);
    k_pp_decide<<<grid1((uint64_t)n + 1), 256, 0, st>>>(parent.p, size.p, ecnt.p, n, dense.p, ntasks.p, msq.p, dense_list.p, n_dense.p);
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    int end_bit = 1;
    while ((1ull << end_bit) < n) end_bit++;
    cub::DeviceScan::ExclusiveSum(nullptr, tb1, ntasks.p, task_off.p, (int64_t)n + 1, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, msq.p, sq_off.p, (int64_t)n + 1, st);
    cub::DeviceRadixSort::SortPairs(nullptr, tb3, parent.p, sroot.p, iota.p, members.p, (int64_t)n, 0, end_bit, st);
    tmp.alloc(std::max(tb1, std::max(tb2, tb3)));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb1, ntasks.p, task_off.p, (int64_t)n + 1, st);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb2, msq.p, sq_off.p, (int64_t)n + 1, st);
    cub::DeviceRadixSort::SortPairs(tmp.p, tb3, parent.p, sroot.p, iota.p, members.p, (int64_t)n, 0, end_bit, st);  // stable: members stay in sample order
    k_pp_heads<<<grid1(n), 256, 0, st>>>(sroot.p, n, comp_start.p);
    k_pp_rank<<<grid1(n), 256, 0, st>>>(sroot.p, members.p, n, comp_start.p, rank.p);
    // capacity bounds that need no round trip to the host: a dense component has m(m-1)/2 <= 4 x its candidates, so
    // sum m^2 <= 8 E + n matrix entries, and (m/64 + 1)^2 block tasks
    sq_cap = 8 * E + n + 16;
    task_cap = E / 256 + 2ull * n + 16;
    tasks.alloc(task_cap);
    scratch_d.alloc(sq_cap);
    k_pp_tasks<<<grid1(n), 256, 0, st>>>(ntasks.p, task_off.p, size.p, n, tasks.p);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned bgrid = (unsigned)std::min<uint64_t>(task_cap, (uint64_t)n_sm * 16);
    k_block_d<0><<<bgrid, BD_THREADS, 0, st>>>(tasks.p, task_off.p, n, members.p, comp_start.p, size.p, sq_off.p, g.planesT.p, g.Wp, 1u, 1u,
                                             scratch_d.p);
    k_pairs_gather<<<grid1(E), 256, 0, st>>>(keys, E, parent.p, dense.p, rank.p, size.p, sq_off.p, scratch_d.p, nullptr, g.ncount.p, d_out,
                                           nullptr);
    k_pairs_sparse<<<grid1(E), 256, 0, st>>>(keys, E, parent.p, dense.p, g.planesT.p, g.Wp, g.nplane.p, g.npitch, g.nsum.p, g.spitch,
                                                g.ncount.p, d_out, nullptr);
    S.kernel_launches += 16 + (end_bit + 7) / 8;
    TRACS_CK(cudaGetLastError());
  }

  void unions(const Ingested &g, const uint64_t *keys, uint64_t E, uint32_t *u_out, cudaStream_t st) {
    tracs_stats_t &S = g_stats;
    if (E == 0 || !u_out) return;
    if (!blocks) {
      k_pairs_sparse<<<grid1(E), 256, 0, st>>>(keys, E, nullptr, nullptr, g.planesT.p, g.Wp, g.nplane.p, g.npitch, g.nsum.p, g.spitch,
                                                  g.ncount.p, nullptr, u_out);
      S.kernel_launches++;
      TRACS_CK(cudaGetLastError());
      return;
    }
    // N intersections: summary-guided (k_block_n) while most 128-site blocks are free of N; a dense AND + POPC
    // contraction over whole N-plane rows once they are not (estimated block occupancy 1 - (1 - p_N)^128 > 35 %)
    const char *nmode = getenv("TRACS_NBLOCKS");  // "dense" / "sparse": tests force both
    const double p_n = g.n && g.L ? (double)g.n_total / ((double)g.n * (double)g.L) : 0.0;
    bool dense_n = 1.0 - pow(1.0 - std::min(1.0, p_n), 128.0) > 0.35;
    if (nmode && !strcmp(nmode, "dense")) dense_n = true;
    if (nmode && !strcmp(nmode, "sparse")) dense_n = false;
    if (g.nplane_sparse) dense_n = false;  // only the 256-site groups that hold an N were stored: whole rows cannot be contracted
    scratch_i.alloc(sq_cap);
    TRACS_CK(cudaMemsetAsync(scratch_i.p, 0, sq_cap * sizeof(uint32_t), st));
    if (dense_n) {
      const uint32_t ksplit = (uint32_t)std::min<uint64_t>(32, std::max<uint64_t>(1, g.npitch / 4 / PP_KC / 64));  // >= 64 chunks per slice
      k_block_d<1><<<(unsigned)std::min<uint64_t>(task_cap * ksplit, (uint64_t)n_sm * 16), BD_THREADS, 0, st>>>(
          tasks.p, task_off.p, n, members.p, comp_start.p, size.p, sq_off.p, reinterpret_cast<const uint4 *>(g.nplane.p), g.npitch / 4, 1u,
          ksplit, scratch_i.p);
    } else {
      k_block_n<<<n_sm * 8, 256, 0, st>>>(dense_list.p, n_dense.p, members.p, comp_start.p, size.p, sq_off.p, g.nplane.p, g.npitch, g.nsum.p,
                                        g.spitch, scratch_i.p);
    }
    k_pairs_gather<<<grid1(E), 256, 0, st>>>(keys, E, parent.p, dense.p, rank.p, size.p, sq_off.p, scratch_d.p, scratch_i.p, g.ncount.p, nullptr,
                                           u_out);
    k_pairs_sparse<<<grid1(E), 256, 0, st>>>(keys, E, parent.p, dense.p, g.planesT.p, g.Wp, g.nplane.p, g.npitch, g.nsum.p, g.spitch,
                                                g.ncount.p, nullptr, u_out);
    S.kernel_launches += 4;
    TRACS_CK(cudaGetLastError());
  }
};

static void eval_pairs(const Ingested &g, const uint64_t *keys, uint64_t E, uint32_t *d_out, uint32_t *u_out, cudaStream_t st) {
  PairEval pe;
  pe.distances(g, keys, E, d_out, st);
  pe.unions(g, keys, E, u_out, st);
}

}  // namespace tracs
