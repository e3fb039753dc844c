// Site-sharded multi-GPU sweep (included at the end of sweep.cu: same translation unit as k_sweep).
//
// Every quantity of the path is ADDITIVE over disjoint site ranges: d(i,j) = sum_r d_r(i,j) and
// |N_i u N_j| = sum_r |N_i u N_j|_r. So for alignments too large for one GPU (C3: 100 000 x 2 Mb =
// 200 GB of ASCII) rank r ingests only the column slab [L*r/R, L*(r+1)/R) of every sequence and
//   1. prefilters ITS share of the triangle row-blocks with the tile kernel on its own first words
//      (a partial distance over any subset of sites is a lower bound of d, so pairs it rejects are
//      decided for good);                                              -> candidate pairs (few)
//   2. the candidate lists are all-gathered (the caller does this with NCCL);
//   3. every rank evaluates its slab's partial d and |N_i u N_j| for ALL candidates;  [pairs.inl]
//   4. the two integer vectors are all-reduced (sum) and thresholded by the caller.
// No bit-plane or N-plane ever crosses NVLink; traffic is O(candidates).
// Reference semantics unchanged: src/pairsnp.hpp:398-403 (d), :417-419 (compared sites).

namespace tracs {

struct SiteShard {
  Ingested ing;
  DevBuf<uint64_t> cand;   // candidate keys (i << 32 | j) found by this rank, sorted
  uint64_t n_cand = 0;
  // likelihood table prepared at open() when the options already carry the sampling days: it is computed on the
  // auxiliary stream under the ingest, and the emit step of the same sweep picks it up (same days / parameters)
  std::unique_ptr<TransLut> lut;
  std::vector<int32_t> lut_days;
  double lut_lamb = 0, lut_beta = 0, lut_thr = 0;
  int32_t lut_dist = -1;
};
static thread_local SiteShard *g_open_shard = nullptr;  // the handle between open() and close() on this thread

// finish: candidates with summed d <= dist, in key order
__global__ void k_finish_flags(const uint32_t *__restrict__ d, uint64_t n, int32_t dist, uint8_t *__restrict__ flags) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) flags[e] = (d[e] <= (uint32_t)dist) ? 1 : 0;
}
__global__ void k_finish_gather(const uint32_t *__restrict__ idx, const uint64_t *__restrict__ n_sel, const uint64_t *__restrict__ keys,
                                const uint32_t *__restrict__ d, const uint32_t *__restrict__ u, uint64_t L_total,
                                uint64_t *__restrict__ keys_o, uint32_t *__restrict__ d_o, uint64_t *__restrict__ rows,
                                uint64_t *__restrict__ cols, uint64_t *__restrict__ dist, uint64_t *__restrict__ ncomp) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= *n_sel) return;
  const uint32_t c = idx[e];
  const uint64_t k = keys[c];
  keys_o[e] = k;
  d_o[e] = d[c];
  rows[e] = k >> 32;
  cols[e] = k & 0xFFFFFFFFull;
  dist[e] = d[c];
  if (u) ncomp[e] = L_total - u[c];
}

// Candidate vectors (device; keys ascending, d = full-length distance, union = |N_i u N_j| or null) -> edge columns on
// the host, in two phases: SELECT (threshold, compared sites; leaves the selected columns on the device and returns the
// count) and EMIT (fused likelihood + copies into host columns given by the caller). finish_candidates() runs both into
// library-owned page-locked columns; the site-sharded sweep at N > 1 runs them apart, with an all-gather of the counts
// in between, so that every rank writes its share of the edge table over its own PCIe link (tracs_site_shard_select /
// _emit).
struct Selection {
  uint64_t E = 0, n = 0;
  bool want_n = false;
  tracs_opts_t o;
  std::vector<int32_t> days;  // copy: the caller's array need not outlive select()
  DevBuf<uint64_t> keys_o, rows, cols, dist, nc;
  DevBuf<uint32_t> d_o;
};
struct HostColumns {  // destination of emit(): caller-owned host arrays (page-locked / registered), element offset applied
  uint64_t *rows, *cols, *dist, *ncomp;
  double *p0_log, *eK, *datediff;
};

static void select_candidates(const uint64_t *dev_keys, const uint32_t *dev_d, const uint32_t *dev_union, uint64_t n_keys, uint64_t n,
                              uint64_t L_total, const tracs_opts_t &o, Selection &sel, cudaStream_t st) {
  tracs_stats_t &S = g_stats;
  sel.E = 0;
  sel.n = n;
  sel.o = o;
  sel.want_n = dev_union != nullptr;
  if (o.days) {
    sel.days.assign(o.days, o.days + n);
    sel.o.days = sel.days.data();
  }
  if (n_keys == 0) return;
  if (n_keys >= (1ull << 32)) throw std::runtime_error("too many candidates");
  Timer T(st);
  T.start();
  DevBuf<uint8_t> flags(n_keys);
  DevBuf<uint32_t> idx(n_keys);
  DevBuf<uint64_t> n_sel(1);
  k_finish_flags<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(dev_d, n_keys, o.dist, flags.p);
  size_t tb = 0;
  cub::CountingInputIterator<uint32_t> cnt_it(0);
  cub::DeviceSelect::Flagged(nullptr, tb, cnt_it, flags.p, idx.p, n_sel.p, (int64_t)n_keys, st);
  DevBuf<uint8_t> tmp(tb);
  cub::DeviceSelect::Flagged(tmp.p, tb, cnt_it, flags.p, idx.p, n_sel.p, (int64_t)n_keys, st);
  // columns are sized by the candidate count (an upper bound): no host round trip before the gather
  sel.keys_o.alloc(n_keys); sel.rows.alloc(n_keys); sel.cols.alloc(n_keys); sel.dist.alloc(n_keys); sel.nc.alloc(n_keys);
  sel.d_o.alloc(n_keys);
  k_finish_gather<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(idx.p, n_sel.p, dev_keys, dev_d, dev_union, L_total, sel.keys_o.p,
                                                                   sel.d_o.p, sel.rows.p, sel.cols.p, sel.dist.p, sel.nc.p);
  S.kernel_launches += 4;
  TRACS_CK(cudaGetLastError());
  uint64_t E = 0;
  TRACS_CK(cudaMemcpyAsync(&E, n_sel.p, 8, cudaMemcpyDeviceToHost, st));
  TRACS_CK(cudaStreamSynchronize(st));
  S.ms_sort += T.stop();
  S.n_edges += E;
  sel.E = E;
}

// returns whether the likelihood columns were written
static bool emit_selection(Selection &sel, const HostColumns &dst, TransLut *lut_in, cudaStream_t st) {
  tracs_stats_t &S = g_stats;
  const uint64_t E = sel.E;
  if (E == 0) return false;
  Timer T(st);
  TRACS_CK(cudaMemcpyAsync(dst.rows, sel.rows.p, E * 8, cudaMemcpyDeviceToHost, st));
  TRACS_CK(cudaMemcpyAsync(dst.cols, sel.cols.p, E * 8, cudaMemcpyDeviceToHost, st));
  TRACS_CK(cudaMemcpyAsync(dst.dist, sel.dist.p, E * 8, cudaMemcpyDeviceToHost, st));
  if (sel.want_n) TRACS_CK(cudaMemcpyAsync(dst.ncomp, sel.nc.p, E * 8, cudaMemcpyDeviceToHost, st));
  S.d2h_bytes += E * (sel.want_n ? 32 : 24);
  TransLut lut_own;
  bool fuse = lut_in != nullptr;
  if (!fuse && g_open_shard && g_open_shard->lut && sel.o.want_trans && sel.o.days && g_open_shard->lut_dist == sel.o.dist &&
      g_open_shard->lut_lamb == sel.o.lamb && g_open_shard->lut_beta == sel.o.beta && g_open_shard->lut_thr == sel.o.threshold_Ek &&
      g_open_shard->lut_days.size() == sel.n && !memcmp(g_open_shard->lut_days.data(), sel.o.days, sel.n * sizeof(int32_t))) {
    lut_in = g_open_shard->lut.get();  // prepared at open()
    fuse = true;
  }
  if (!fuse) fuse = lut_own.setup(sel.o, sel.n, (uint64_t)std::max<int64_t>(sel.o.dist, 0) + 1, st);
  TransLut &lut = lut_in ? *lut_in : lut_own;
  DevBuf<double> d_p0, d_eK, d_dt;
  if (fuse) {
    T.start();
    d_p0.alloc(E); d_eK.alloc(E); d_dt.alloc(E);
    lut.apply(sel.keys_o.p, sel.d_o.p, E, d_p0.p, d_eK.p, d_dt.p, st);
    S.ms_trans += T.stop();
    TRACS_CK(cudaMemcpyAsync(dst.p0_log, d_p0.p, E * 8, cudaMemcpyDeviceToHost, st));
    TRACS_CK(cudaMemcpyAsync(dst.eK, d_eK.p, E * 8, cudaMemcpyDeviceToHost, st));
    TRACS_CK(cudaMemcpyAsync(dst.datediff, d_dt.p, E * 8, cudaMemcpyDeviceToHost, st));
    S.d2h_bytes += E * 24;
  }
  T.start();
  TRACS_CK(cudaStreamSynchronize(st));
  S.ms_d2h += T.stop();
  return fuse;
}

static void finish_candidates(const uint64_t *dev_keys, const uint32_t *dev_d, const uint32_t *dev_union, uint64_t n_keys, uint64_t n,
                              uint64_t L_total, const tracs_opts_t &o, TransLut *lut_in, HostEdges &out, cudaStream_t st) {
  Selection sel;
  select_candidates(dev_keys, dev_d, dev_union, n_keys, n, L_total, o, sel, st);
  const uint64_t E = sel.E;
  if (E == 0) return;
  const bool fuse = lut_in != nullptr || (o.want_trans && o.days && o.dist >= 0);
  const size_t old = out.rows.size();
  out.rows.resize(old + E); out.cols.resize(old + E); out.dist.resize(old + E);
  if (sel.want_n) out.ncomp.resize(old + E);
  if (fuse) { out.p0_log.resize(old + E); out.eK.resize(old + E); out.datediff.resize(old + E); }
  HostColumns dst{out.rows.data() + old, out.cols.data() + old, out.dist.data() + old, sel.want_n ? out.ncomp.data() + old : nullptr,
                  fuse ? out.p0_log.data() + old : nullptr, fuse ? out.eK.data() + old : nullptr, fuse ? out.datediff.data() + old : nullptr};
  if (emit_selection(sel, dst, lut_in, st)) {
    out.has_trans = true;
  } else if (fuse) {  // table too large for the device-side path: finish_edges() takes the unique-key route
    out.p0_log.resize(old); out.eK.resize(old); out.datediff.resize(old);
  }
}

__global__ void k_finish_ncomp(const uint32_t *__restrict__ idx, const uint64_t *__restrict__ n_sel, const uint32_t *__restrict__ u,
                               uint64_t L_total, uint64_t *__restrict__ ncomp) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < *n_sel) ncomp[e] = L_total - u[idx[e]];
}

// Single-GPU tail of a thresholded call: candidates (device, keys ascending) -> edge columns on the host, with the
// copies of the finished columns UNDER the compared-sites kernels. The distances are complete after the first phase of
// PairEval, so threshold, selection and the likelihood columns are enqueued right behind it, followed by the
// compared-sites phase; while that runs (5 of the 7 ms of the refinement at C3), the host learns the edge count and
// the auxiliary stream copies rows / cols / dist / likelihood columns (100 of the 140 MB). Only the compared-sites
// column is copied after its kernel. No host synchronisation sits between two kernels of the main stream.
static void eval_finish_overlapped(const Ingested &g, const uint64_t *keys, uint64_t n_keys, uint32_t *d, uint32_t *u, uint64_t n,
                                   uint64_t L_total, const tracs_opts_t &o, TransLut *lut, HostEdges &out, cudaStream_t st) {
  tracs_stats_t &S = g_stats;
  if (n_keys == 0) return;
  if (n_keys >= (1ull << 32)) throw std::runtime_error("too many candidates");
  const bool want_n = u != nullptr, fuse = lut != nullptr;
  cudaStream_t aux = aux_stream();
  // events belong to the device they were created on: one pair per device (a thread may drive several)
  static thread_local cudaEvent_t ev_sel_dev[64] = {}, ev_aux_dev[64] = {};
  static thread_local uint64_t *h_E = nullptr;  // page-locked landing place of the edge count
  int cur_dev = 0;
  TRACS_CK(cudaGetDevice(&cur_dev));
  cur_dev &= 63;
  if (!ev_sel_dev[cur_dev]) {
    TRACS_CK(cudaEventCreateWithFlags(&ev_sel_dev[cur_dev], cudaEventDisableTiming));
    TRACS_CK(cudaEventCreateWithFlags(&ev_aux_dev[cur_dev], cudaEventDisableTiming));
  }
  if (!h_E) TRACS_CK(cudaHostAlloc((void **)&h_E, sizeof(uint64_t), cudaHostAllocPortable));
  const cudaEvent_t ev_sel = ev_sel_dev[cur_dev], ev_aux = ev_aux_dev[cur_dev];
  DeferredTimers DT(st);
  PairEval pe;
  DT.start(&S.ms_refine);
  pe.distances(g, keys, n_keys, d, st);
  DT.stop();
  // ---- threshold + selection + likelihood columns, all sized by the candidate count (an upper bound) -------------
  DT.start(&S.ms_sort);
  DevBuf<uint8_t> flags(n_keys);
  DevBuf<uint32_t> idx(n_keys), d_o(n_keys);
  DevBuf<uint64_t> n_sel(1), keys_o(n_keys), rows(n_keys), cols(n_keys), dist(n_keys), nc(want_n ? n_keys : 1);
  k_finish_flags<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(d, n_keys, o.dist, flags.p);
  size_t tb = 0;
  cub::CountingInputIterator<uint32_t> cnt_it(0);
  cub::DeviceSelect::Flagged(nullptr, tb, cnt_it, flags.p, idx.p, n_sel.p, (int64_t)n_keys, st);
  DevBuf<uint8_t> tmp(tb);
  cub::DeviceSelect::Flagged(tmp.p, tb, cnt_it, flags.p, idx.p, n_sel.p, (int64_t)n_keys, st);
  k_finish_gather<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(idx.p, n_sel.p, keys, d, nullptr, L_total, keys_o.p, d_o.p, rows.p, cols.p,
                                                                   dist.p, nullptr);
  S.kernel_launches += 4;
  TRACS_CK(cudaGetLastError());
  TRACS_CK(cudaMemcpyAsync(h_E, n_sel.p, 8, cudaMemcpyDeviceToHost, st));
  DT.stop();
  DevBuf<double> d_p0, d_eK, d_dt;
  if (fuse) {
    DT.start(&S.ms_trans);
    d_p0.alloc(n_keys); d_eK.alloc(n_keys); d_dt.alloc(n_keys);
    lut->apply(keys_o.p, d_o.p, n_keys, d_p0.p, d_eK.p, d_dt.p, st, n_sel.p);
    DT.stop();
  }
  TRACS_CK(cudaEventRecord(ev_sel, st));
  // ---- compared sites: enqueued before the host looks at the count ---------------------------------------------------
  if (want_n) {
    DT.start(&S.ms_refine);
    pe.unions(g, keys, n_keys, u, st);
    k_finish_ncomp<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>(idx.p, n_sel.p, u, L_total, nc.p);
    S.kernel_launches++;
    TRACS_CK(cudaGetLastError());
    DT.stop();
  }
  // ---- the finished columns leave while that runs ---------------------------------------------------------------------
  TRACS_CK(cudaEventSynchronize(ev_sel));
  const uint64_t E = *h_E;
  S.n_edges += E;
  if (E == 0) {
    TRACS_CK(cudaStreamSynchronize(st));
    return;
  }
  const size_t old = out.rows.size();
  out.rows.resize(old + E); out.cols.resize(old + E); out.dist.resize(old + E);
  if (want_n) out.ncomp.resize(old + E);
  if (fuse) { out.p0_log.resize(old + E); out.eK.resize(old + E); out.datediff.resize(old + E); }
  TRACS_CK(cudaStreamWaitEvent(aux, ev_sel, 0));
  TRACS_CK(cudaMemcpyAsync(out.rows.data() + old, rows.p, E * 8, cudaMemcpyDeviceToHost, aux));
  TRACS_CK(cudaMemcpyAsync(out.cols.data() + old, cols.p, E * 8, cudaMemcpyDeviceToHost, aux));
  TRACS_CK(cudaMemcpyAsync(out.dist.data() + old, dist.p, E * 8, cudaMemcpyDeviceToHost, aux));
  S.d2h_bytes += E * 24;
  if (fuse) {
    TRACS_CK(cudaMemcpyAsync(out.p0_log.data() + old, d_p0.p, E * 8, cudaMemcpyDeviceToHost, aux));
    TRACS_CK(cudaMemcpyAsync(out.eK.data() + old, d_eK.p, E * 8, cudaMemcpyDeviceToHost, aux));
    TRACS_CK(cudaMemcpyAsync(out.datediff.data() + old, d_dt.p, E * 8, cudaMemcpyDeviceToHost, aux));
    S.d2h_bytes += E * 24;
    out.has_trans = true;
  }
  TRACS_CK(cudaEventRecord(ev_aux, aux));
  if (want_n) {
    DT.start(&S.ms_d2h);
    TRACS_CK(cudaMemcpyAsync(out.ncomp.data() + old, nc.p, E * 8, cudaMemcpyDeviceToHost, st));
    S.d2h_bytes += E * 8;
    DT.stop();
  }
  TRACS_CK(cudaStreamSynchronize(st));
  TRACS_CK(cudaEventSynchronize(ev_aux));  // the device columns go back to the cache below
  DT.resolve();
}

void site_shard_finish_device(const uint64_t *dev_keys, const uint32_t *dev_d, const uint32_t *dev_union, uint64_t n_keys, uint64_t n,
                              uint64_t L_total, const tracs_opts_t &o, HostEdges &out, cudaStream_t st) {
  finish_candidates(dev_keys, dev_d, dev_union, n_keys, n, L_total, o, nullptr, out, st);
}

}  // namespace tracs

using namespace tracs;

extern "C" {

int tracs_site_shard_open(const uint8_t *dev_slab, size_t n, size_t L_slab, size_t pitch, const tracs_opts_t *opts,
                          void **handle, const uint64_t **dev_cand_keys, size_t *n_cand) {
  *handle = nullptr;
  *dev_cand_keys = nullptr;
  *n_cand = 0;
  memset(&g_stats, 0, sizeof g_stats);
  return guarded([&] {
    require_device();
    if (!opts) throw std::runtime_error("site shard: options required");
    const tracs_opts_t &o = *opts;
    if (o.dist < 0 || (uint64_t)o.dist >= (uint64_t)PREFILTER_WORDS * 32)
      throw std::runtime_error("site-sharded sweep needs a SNP threshold in [0, 2048): it relies on the prefilter");
    if (o.filter) throw std::runtime_error("site-sharded sweep does not run the recombination filter");
    if (n == 0 || n >= (1ull << 31)) throw std::runtime_error("site shard: bad sample count");
    cudaStream_t st = 0;
    Timer T(st), Ttot(st);
    Ttot.start();
    std::unique_ptr<SiteShard> sh(new SiteShard());
    g_stats.n_samples = n;
    g_stats.seq_length = L_slab;
    if (o.want_trans && o.days) {  // optional: the table kernel then runs under the ingest
      sh->lut.reset(new TransLut());
      if (sh->lut->setup(o, n, (uint64_t)o.dist + 1, st)) {
        sh->lut_days.assign(o.days, o.days + n);
        sh->lut_lamb = o.lamb; sh->lut_beta = o.beta; sh->lut_thr = o.threshold_Ek; sh->lut_dist = o.dist;
      } else {
        sh->lut.reset();
      }
    }
    ingest_device(dev_slab, n, L_slab, pitch, o.packed_input != 0, false, sh->ing, st);
    const Ingested &g = sh->ing;
    const uint64_t i_end = (o.i_end == 0 || o.i_end > n) ? n : o.i_end;
    const int world = std::max(1, (int)o.shard_world), rank = std::max(0, (int)o.shard_rank);
    const TilePlan plan = plan_tiles(n, i_end, o.j_start, g.Npad, rank, world);
    // all of this rank's row-blocks append to ONE candidate list; launches are cut so that the tile
    // count stays in 32 bits, the list capacity bounds the (few) survivors
    const uint64_t CAND_CAP = 1ull << 26;
    DevBuf<uint64_t> keys(CAND_CAP), keys2(CAND_CAP);
    DevBuf<uint32_t> dv(CAND_CAP), dv2(CAND_CAP);
    DevBuf<unsigned long long> counter(1);
    TRACS_CK(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), st));
    const bool tc_ok = o.sweep_variant != 1 && !g.partial_ambiguity;
    DevBuf<uint32_t> d_rb(std::max<size_t>(1, plan.my_rb.size())), d_prefix(plan.my_rb.size() + 1);
    uint64_t my_pairs = 0;
    for (uint64_t p : plan.rb_pairs) my_pairs += p;
    g_stats.n_pairs = my_pairs;
    unsigned long long nc = 0;
    uint32_t words_used = 0;
    // windows of 4, 8, 16, 64 local words (any prefix of the slab's words gives a lower bound of d), widened while more
    // than 4 % of this rank's pairs survive; Wp is a multiple of KC. Kernel choice as in sweep_device.
    const uint32_t first = first_window(o.dist);
    for (uint32_t pw = first;; pw = pw < 8 ? 8 : pw < 16 ? 16 : PREFILTER_WORDS) {
      const uint32_t words = std::min<uint32_t>(pw, g.Wp);
      TRACS_CK(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), st));
      size_t k0 = 0;
      T.start();
      while (k0 < plan.my_rb.size()) {
        std::vector<uint32_t> rbs, prefix{0};
        uint64_t tiles = 0;
        size_t k1 = k0;
        while (k1 < plan.my_rb.size() && tiles < (1ull << 30)) {
          tiles += plan.n_cb - std::max(plan.my_rb[k1], plan.cb_min);
          rbs.push_back(plan.my_rb[k1]);
          prefix.push_back((uint32_t)tiles);
          ++k1;
        }
        TRACS_CK(cudaMemcpyAsync(d_rb.p, rbs.data(), rbs.size() * 4, cudaMemcpyHostToDevice, st));
        TRACS_CK(cudaMemcpyAsync(d_prefix.p, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, st));
        SweepArgs a;
        a.planes = g.planes.p; a.Wp = words; a.Npad = g.Npad; a.n = (uint32_t)n; a.i_end = (uint32_t)i_end;
        a.j_start = (uint32_t)o.j_start; a.dist = o.dist; a.rb_list = d_rb.p; a.tile_prefix = d_prefix.p;
        a.n_rb = (uint32_t)rbs.size(); a.n_tiles = (uint32_t)tiles; a.cb_min = plan.cb_min; a.counter = counter.p;
        a.keys = keys.p; a.dvals = dv.p; a.cap = CAND_CAP; a.one = 1;
        DevBuf<uint2> d_table(std::max<uint64_t>(1, tiles));
        if (tiles) {
          k_tile_table<<<(unsigned)((tiles + 255) / 256), 256, 0, st>>>(d_rb.p, d_prefix.p, a.n_rb, plan.cb_min, (uint32_t)tiles, d_table.p);
          g_stats.kernel_launches++;
        }
        a.tile_table = d_table.p;
        launch_tile_sweep(a, tc_ok && words >= tc_min_words(), st, g.has_n_var);
        TRACS_CK(cudaStreamSynchronize(st));  // rbs/prefix are reused by the next cut
        g_stats.n_tiles += tiles;
        k0 = k1;
      }
      g_stats.ms_sweep += T.stop();
      g_stats.swept_wordpairs += my_pairs * words;
      TRACS_CK(cudaMemcpyAsync(&nc, counter.p, sizeof nc, cudaMemcpyDeviceToHost, st));
      TRACS_CK(cudaStreamSynchronize(st));
      words_used = words;
      if ((nc <= CAND_CAP && nc * 25 <= my_pairs) || words >= std::min<uint32_t>(PREFILTER_WORDS, g.Wp)) break;
    }
    if (nc > CAND_CAP)
      throw std::runtime_error("site-sharded sweep: the prefilter left too many candidate pairs; use the single-GPU / tile-sharded path");
    if (nc && words_used < 16 && g.Wp > words_used) {
      // narrow window: trim this rank's candidates over the next local words (k_cand_trim), so that the few unrelated
      // pairs it let through do not bridge clusters in the candidate graph
      T.start();
      const uint32_t extra = std::min<uint32_t>(g.Wp - words_used, 32 - words_used);
      DevBuf<uint8_t> flags(nc);
      DevBuf<uint64_t> n_sel(1);
      k_cand_trim<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(keys.p, dv.p, nc, g.planesT.p, g.Wp, words_used, extra, o.dist, flags.p);
      size_t tb = 0;
      cub::DeviceSelect::Flagged(nullptr, tb, keys.p, flags.p, keys2.p, n_sel.p, (int64_t)nc, st);
      DevBuf<uint8_t> tmp(tb);
      cub::DeviceSelect::Flagged(tmp.p, tb, keys.p, flags.p, keys2.p, n_sel.p, (int64_t)nc, st);
      uint64_t kept = 0;
      TRACS_CK(cudaMemcpyAsync(&kept, n_sel.p, 8, cudaMemcpyDeviceToHost, st));
      TRACS_CK(cudaStreamSynchronize(st));
      std::swap(keys.p, keys2.p);
      g_stats.kernel_launches += 3;
      g_stats.ms_refine += T.stop();
      nc = kept;
    }
    g_stats.n_candidates = nc;
    if (nc) {
      T.start();
      int end_bit = 32;
      while ((1ull << (end_bit - 32)) < n) end_bit++;
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys2.p, dv.p, dv2.p, (int64_t)nc, 0, end_bit, st);
      DevBuf<uint8_t> tmp(tb);
      cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, dv.p, dv2.p, (int64_t)nc, 0, end_bit, st);
      g_stats.kernel_launches += 2 + (end_bit + 7) / 8;
      sh->cand.alloc(nc);
      TRACS_CK(cudaMemcpyAsync(sh->cand.p, keys2.p, nc * 8, cudaMemcpyDeviceToDevice, st));
      TRACS_CK(cudaStreamSynchronize(st));
      g_stats.ms_sort += T.stop();
    }
    sh->n_cand = nc;
    g_stats.ms_total = Ttot.stop();
    *dev_cand_keys = sh->cand.p;
    *n_cand = nc;
    g_open_shard = sh.get();
    *handle = sh.release();
  });
}

int tracs_site_shard_partials(void *handle, const uint64_t *dev_keys, size_t n_keys, uint32_t *dev_d, uint32_t *dev_union) {
  return guarded([&] {
    if (!handle) throw std::runtime_error("site shard: null handle");
    SiteShard *sh = (SiteShard *)handle;
    const Ingested &g = sh->ing;
    if (!n_keys) return;
    Timer T(0);
    T.start();
    eval_pairs(g, dev_keys, n_keys, dev_d, dev_union, 0);
    g_stats.ms_refine += T.stop();
  });
}

int tracs_site_shard_select(const uint64_t *dev_keys, const uint32_t *dev_d, const uint32_t *dev_union, size_t n_keys, size_t n_samples,
                            size_t L_total, const tracs_opts_t *opts, void **selection, size_t *n_selected) {
  *selection = nullptr;
  *n_selected = 0;
  memset(&g_stats, 0, sizeof g_stats);
  return guarded([&] {
    require_device();
    if (!opts) throw std::runtime_error("site shard: options required");
    std::unique_ptr<Selection> sel(new Selection());
    tracs_opts_t o = *opts;
    o.filter = 0;
    select_candidates(dev_keys, dev_d, dev_union, n_keys, n_samples, L_total, o, *sel, 0);
    *n_selected = sel->E;
    *selection = sel.release();
  });
}

int tracs_site_shard_emit(void *selection, const tracs_edges_t *dst, size_t at, int *has_trans) {
  std::unique_ptr<Selection> sel((Selection *)selection);
  if (has_trans) *has_trans = 0;
  return guarded([&] {
    if (!sel) throw std::runtime_error("site shard: null selection");
    if (!dst || !dst->rows || !dst->cols || !dst->dist) throw std::runtime_error("site shard: destination columns required");
    const bool fuse = sel->o.want_trans && sel->o.days && dst->p0_log && dst->eK && dst->datediff;
    if (!fuse) sel->o.want_trans = 0;
    if (sel->want_n && !dst->ncomp) throw std::runtime_error("site shard: destination for the compared-sites column required");
    HostColumns hc{dst->rows + at, dst->cols + at, dst->dist + at, dst->ncomp ? dst->ncomp + at : nullptr,
                   fuse ? dst->p0_log + at : nullptr, fuse ? dst->eK + at : nullptr, fuse ? dst->datediff + at : nullptr};
    const bool wrote = emit_selection(*sel, hc, nullptr, 0);
    if (fuse && sel->E && !wrote) throw std::runtime_error("site shard: the (distance, day) table is too large for the fused likelihood");
    if (has_trans) *has_trans = wrote ? 1 : 0;
  });
}

int tracs_site_shard_close(void *handle) {
  return guarded([&] {
    if (g_open_shard == (SiteShard *)handle) g_open_shard = nullptr;
    delete (SiteShard *)handle;
  });
}

}  // extern "C"
