/*
 * tracs_oracle.c -- CPU restatement of the TRACS pairwise-distance hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity checker for the CUDA library in
 * tracs_b200/csrc. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it; the product never does (and has no CPU fallback).
 *
 * Parity status: PINNED. tests/test_oracle.py checks this file against
 *   - the reference's own known-answer tests (tests/test_llk.py:21-29,
 *     tests/test_trans_distance.py:29-42, tests/test_pairsnp.py:7-9 ordering pin), and
 *   - outputs of the UNMODIFIED reference compiled here (oracle/_ref, see build_ref.sh), both
 *     live (when oracle/_ref is present) and through fixtures committed in tests/golden/.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * Plain C99 + zlib + optional OpenMP; scalar and deliberately simple.
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
  uint64_t n_edges;
  uint64_t *rows, *cols, *dist, *filt, *ncomp;
  uint64_t n_names;
  char **names;
  uint64_t seq_length;
} orc_edges;

static void set_err(char *err, size_t n, const char *msg) {
  if (err && n) { strncpy(err, msg, n - 1); err[n - 1] = 0; }
}

/* ---- character -> 4-bit base mask (bit0=A bit1=C bit2=G bit3=T) -------------------------
 * src/pairsnp.hpp:107-199: toupper, then A C G T, the ten 2-/3-base IUPAC codes, and
 * EVERYTHING else (N, '-', X, '?', digits ...) = all four bases. */
uint8_t orc_base_mask(int ch) {
  switch (toupper(ch)) {
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;  case 'T': return 8;
    case 'M': return 1 | 2;  case 'R': return 1 | 4;  case 'W': return 1 | 8;
    case 'S': return 2 | 4;  case 'Y': return 2 | 8;  case 'K': return 4 | 8;
    case 'V': return 1 | 2 | 4;  case 'H': return 1 | 2 | 8;
    case 'D': return 1 | 4 | 8;  case 'B': return 2 | 4 | 8;
    default:  return 15;
  }
}

/* ---- FASTA/FASTQ record reader with klib-kseq semantics ---------------------------------
 * src/kseq.h:170-208 (behaviour, restated): skip to the first '>' or '@'; name = header up
 * to the first whitespace; rest of the header line is a comment; sequence = every isgraph()
 * byte until the next '>', '@' or '+' ANYWHERE; after '+' a FASTQ quality block of the same
 * length follows (-2 if it is short). gz or plain input through zlib (src/pairsnp.hpp:75). */
typedef struct { gzFile f; unsigned char buf[1 << 16]; int beg, end, eof, last; } rd_t;
static int rd_getc(rd_t *r) {
  if (r->beg >= r->end) {
    if (r->eof) return -1;
    r->beg = 0;
    r->end = gzread(r->f, r->buf, sizeof r->buf);
    if (r->end < (int)sizeof r->buf) r->eof = 1;
    if (r->end <= 0) return -1;
  }
  return r->buf[r->beg++];
}
typedef struct { char *s; size_t l, m; } str_t;
static void str_push(str_t *s, int c) {
  if (s->l + 2 > s->m) { s->m = s->m ? s->m * 2 : 256; s->s = (char *)realloc(s->s, s->m); }
  s->s[s->l++] = (char)c; s->s[s->l] = 0;
}
/* returns seq length >=0, -1 EOF, -2 truncated quality */
static long rd_record(rd_t *r, str_t *name, str_t *seq) {
  int c;
  if (r->last == 0) {
    while ((c = rd_getc(r)) != -1 && c != '>' && c != '@') {}
    if (c == -1) return -1;
    r->last = c;
  }
  name->l = 0; seq->l = 0;
  if (!name->s) str_push(name, 0), name->l = 0, name->s[0] = 0;
  if (!seq->s) str_push(seq, 0), seq->l = 0, seq->s[0] = 0;
  /* name: up to first whitespace; if the stream is exhausted before any byte -> EOF */
  int got = 0;
  while ((c = rd_getc(r)) != -1) { got = 1; if (isspace(c)) break; str_push(name, c); }
  if (!got) return -1;
  if (c != -1 && c != '\n') while ((c = rd_getc(r)) != -1 && c != '\n') {}
  while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@')
    if (isgraph(c)) str_push(seq, c);
  if (c == '>' || c == '@') r->last = c;
  if (c != '+') return (long)seq->l;
  while ((c = rd_getc(r)) != -1 && c != '\n') {}
  if (c == -1) return -2;
  size_t q = 0;
  while (q < seq->l && (c = rd_getc(r)) != -1) if (c >= 33 && c <= 127) q++;
  r->last = 0;
  if (q != seq->l) return -2;
  return (long)seq->l;
}

typedef struct { uint8_t *ascii; uint64_t n, L; char **names; } aln_t;

/* src/pairsnp.hpp:62-220: read every record, demand equal lengths WITHIN a file. */
static int load_file(const char *path, uint8_t **ascii, uint64_t *n_io, uint64_t *L_out, char ***names,
                     char *err, size_t errlen) {
  rd_t *r = (rd_t *)calloc(1, sizeof(rd_t));
  r->f = gzopen(path, "r");
  if (!r->f) { set_err(err, errlen, "Error reading FASTA!"); free(r); return 1; }
  str_t name = {0, 0, 0}, seq = {0, 0, 0};
  uint64_t cnt = 0, L = 0, n0 = *n_io, cap = 0;
  uint8_t *rows = NULL;
  long l;
  int rc = 0;
  while ((l = rd_record(r, &name, &seq)) != -1) {
    if (l < 0) { set_err(err, errlen, "Error reading FASTA!"); rc = 1; break; }
    if (cnt > 0 && (uint64_t)l != L) { set_err(err, errlen, "Error reading FASTA, variable sequence lengths!"); rc = 1; break; }
    L = (uint64_t)l;
    if ((cnt + 1) * (L ? L : 1) > cap) { cap = (cnt + 1) * (L ? L : 1) * 2; rows = (uint8_t *)realloc(rows, cap); }
    memcpy(rows + cnt * L, seq.s, L);
    *names = (char **)realloc(*names, (n0 + cnt + 1) * sizeof(char *));
    (*names)[n0 + cnt] = strdup(name.s);
    cnt++;
  }
  gzclose(r->f); free(r); free(name.s); free(seq.s);
  if (rc) { free(rows); return rc; }
  *ascii = rows; *n_io = n0 + cnt; *L_out = L;
  return 0;
}

/* ---- bit-plane helpers -------------------------------------------------------------------*/
static inline uint64_t popc64(uint64_t x) { return (uint64_t)__builtin_popcountll(x); }

/* Recombination filter, restating src/pairsnp.hpp:223-318 on the SNP-position list of one pair.
 * snp[] = ascending positions where the pair mismatches (the flipped `res`), d of them.
 * Window = [max(0,i-h), min(L,i+h+1)); count SNPs in it and span = last-first+1 (range_count
 * :223-248); keep the SNP if the window holds one SNP or 1-BinomCDF(count; span, p) >= 0.05/d.
 * The binomial CDF is the same direct-summation stand-in oracle/_ref uses (Boost absent): the
 * filter sub-path is therefore "parity unpinned" against real Boost. */
static double binom_cdf(double n, double p, double k) {
  if (k >= n) return 1.0;
  if (k < 0) return 0.0;
  if (p <= 0) return 1.0;
  if (p >= 1) return 0.0;
  double lp = log(p), lq = log1p(-p), s = 0.0;
  for (double i = 0; i <= floor(k); i += 1.0)
    s += exp(lgamma(n + 1) - lgamma(i + 1) - lgamma(n - i + 1) + i * lp + (n - i) * lq);
  return s > 1.0 ? 1.0 : s;
}
uint64_t orc_filter_recomb(const uint64_t *snp, uint64_t d, uint64_t L) {
  if (d <= 1) return d;
  double dd = (double)d;
  int aln = (int)L;
  double p = dd / aln, thr = 0.05 / dd;
  int h = (int)(1.0 / p / 2.0 + 1);
  if (h > 5000) h = 5000;
  if (h < 50) h = 50;
  uint64_t kept = 0;
  for (uint64_t s = 0; s < d; s++) {
    int i = (int)snp[s];
    int64_t left = i - h > 0 ? i - h : 0;
    int64_t right = i + h + 1 < aln ? i + h + 1 : aln;
    uint64_t cnt = 0, first = 0, last = 0;
    for (uint64_t t = 0; t < d; t++) {
      if ((int64_t)snp[t] >= right) break;
      if ((int64_t)snp[t] >= left) { if (!cnt) first = snp[t]; last = snp[t]; cnt++; }
    }
    uint64_t span = last - first + 1;
    if (cnt > 1) {
      double pv = 1.0 - binom_cdf((double)(int)span, p, (double)(int)cnt);
      if (pv >= thr) kept++;
    } else kept++;
  }
  return kept;
}

/* ---- the pair sweep ------------------------------------------------------------------------
 * src/pairsnp.hpp:380-432 on an ASCII matrix seqs[n][L]:
 *   d(i,j)  = L - popcount((Ai&Aj)|(Ci&Cj)|(Gi&Gj)|(Ti&Tj))             (:398-403)
 *   emit iff d <= dist (int compare)                                     (:405-410)
 *   nn(i,j) = L - popcount((Ai&Ci&Gi&Ti)|(Aj&Cj&Gj&Tj))                  (:417-419)
 *   rows i in [0,i_end), cols j in [max(j_start,i+1), n); output in (i,j) order (:395,:450-457)
 *   filt = zeros unless filter (:411-414, :452). */
int orc_pairsnp_ascii(const uint8_t *seqs, uint64_t n, uint64_t L, uint64_t i_end, uint64_t j_start,
                      int n_threads, int dist, int filter, orc_edges *out) {
  uint64_t W = (L + 63) / 64;
  uint64_t *pl = (uint64_t *)calloc((size_t)(n * 4 * W + 1), sizeof(uint64_t)); /* [n][4][W] */
  for (uint64_t s = 0; s < n; s++)
    for (uint64_t k = 0; k < L; k++) {
      uint8_t m = orc_base_mask(seqs[s * L + k]);
      for (int b = 0; b < 4; b++)
        if (m & (1 << b)) pl[(s * 4 + b) * W + (k >> 6)] |= (uint64_t)1 << (k & 63);
    }
  uint64_t **er = (uint64_t **)calloc(n + 1, sizeof(uint64_t *)); /* per-row edge triples+ */
  uint64_t *ec = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (uint64_t i = 0; i < i_end; i++) {
    uint64_t cap = 0, cnt = 0, *buf = NULL;
    uint64_t *snp = filter ? (uint64_t *)malloc((L + 1) * sizeof(uint64_t)) : NULL;
    const uint64_t *Ai = pl + (i * 4 + 0) * W, *Ci = pl + (i * 4 + 1) * W, *Gi = pl + (i * 4 + 2) * W, *Ti = pl + (i * 4 + 3) * W;
    uint64_t j0 = j_start > i + 1 ? j_start : i + 1;
    for (uint64_t j = j0; j < n; j++) {
      const uint64_t *Aj = pl + (j * 4 + 0) * W, *Cj = pl + (j * 4 + 1) * W, *Gj = pl + (j * 4 + 2) * W, *Tj = pl + (j * 4 + 3) * W;
      uint64_t match = 0;
      for (uint64_t w = 0; w < W; w++) match += popc64((Ai[w] & Aj[w]) | (Ci[w] & Cj[w]) | (Gi[w] & Gj[w]) | (Ti[w] & Tj[w]));
      int d = (int)(L - match);
      if (d > dist) continue;
      uint64_t both = 0;
      for (uint64_t w = 0; w < W; w++) both += popc64((Ai[w] & Ci[w] & Gi[w] & Ti[w]) | (Aj[w] & Cj[w] & Gj[w] & Tj[w]));
      uint64_t fd = 0;
      if (filter) {
        uint64_t ns = 0;
        for (uint64_t k = 0; k < L; k++) {
          uint64_t w = k >> 6, b = k & 63;
          uint64_t mm = (Ai[w] & Aj[w]) | (Ci[w] & Cj[w]) | (Gi[w] & Gj[w]) | (Ti[w] & Tj[w]);
          if (!((mm >> b) & 1)) snp[ns++] = k;
        }
        fd = orc_filter_recomb(snp, ns, L);
      }
      if (cnt == cap) { cap = cap ? cap * 2 : 16; buf = (uint64_t *)realloc(buf, cap * 4 * sizeof(uint64_t)); }
      buf[cnt * 4 + 0] = j; buf[cnt * 4 + 1] = (uint64_t)d; buf[cnt * 4 + 2] = (uint64_t)(int)(L - both); buf[cnt * 4 + 3] = fd;
      cnt++;
    }
    free(snp);
    er[i] = buf; ec[i] = cnt;
  }
  uint64_t E = 0;
  for (uint64_t i = 0; i < i_end; i++) E += ec[i];
  out->n_edges = E;
  out->rows = (uint64_t *)malloc((E + 1) * 8); out->cols = (uint64_t *)malloc((E + 1) * 8);
  out->dist = (uint64_t *)malloc((E + 1) * 8); out->filt = (uint64_t *)malloc((E + 1) * 8);
  out->ncomp = (uint64_t *)malloc((E + 1) * 8);
  uint64_t e = 0;
  for (uint64_t i = 0; i < i_end; i++) {
    for (uint64_t t = 0; t < ec[i]; t++, e++) {
      out->rows[e] = i; out->cols[e] = er[i][t * 4]; out->dist[e] = er[i][t * 4 + 1];
      out->ncomp[e] = er[i][t * 4 + 2]; out->filt[e] = filter ? er[i][t * 4 + 3] : 0;
    }
    free(er[i]);
  }
  free(er); free(ec); free(pl);
  out->seq_length = L;
  return 0;
}

/* src/pairsnp.hpp:320-458: one or two FASTA files; two files = query x db with global indices. */
int orc_pairsnp(const char *const *paths, int n_paths, int n_threads, int dist, int filter, orc_edges *out,
                char *err, size_t errlen) {
  memset(out, 0, sizeof *out);
  if (n_paths < 1 || n_paths > 2) { set_err(err, errlen, "Invalid number of fasta files!"); return 1; }
  uint8_t *a0 = NULL, *a1 = NULL;
  uint64_t n = 0, L0 = 0, L1 = 0;
  char **names = NULL;
  if (load_file(paths[0], &a0, &n, &L0, &names, err, errlen)) return 1;
  uint64_t n1 = n, i_end = n, j_start = 0;
  if (n_paths == 2) {
    if (load_file(paths[1], &a1, &n, &L1, &names, err, errlen)) { free(a0); return 1; }
    j_start = n1;
    /* the reference does not cross-check the two files' lengths (bitset sizes would differ: UB).
       The restatement demands equality when the second file is non-empty. */
    if (n > n1 && L1 != L0) { set_err(err, errlen, "Error reading FASTA, variable sequence lengths!"); free(a0); free(a1); return 1; }
  }
  uint8_t *all = (uint8_t *)malloc((size_t)(n * L0 + 1));
  if (n1) memcpy(all, a0, (size_t)(n1 * L0));
  if (n > n1) memcpy(all + n1 * L0, a1, (size_t)((n - n1) * L0));
  free(a0); free(a1);
  int rc = orc_pairsnp_ascii(all, n, L0, i_end, j_start, n_threads, dist, filter, out);
  free(all);
  out->n_names = n; out->names = names;
  return rc;
}

void orc_edges_free(orc_edges *e) {
  free(e->rows); free(e->cols); free(e->dist); free(e->filt); free(e->ncomp);
  for (uint64_t i = 0; i < e->n_names; i++) free(e->names[i]);
  free(e->names);
  memset(e, 0, sizeof *e);
}

/* ================= transcluster ============================================================ */

/* src/transcluster.hpp:62-75 */
static double lae(double x, double y) {
  double t = x - y;
  if (x == y) return x + M_LN2;
  if (t > 0) return x + log1p(exp(-t));
  else if (t <= 0) return y + log1p(exp(t));
  return t;
}

/* src/transcluster.hpp:90-129 (first formulation; test-only entry point). lg[x] = lgamma(x). */
void orc_lprob_k_given_N(uint64_t N, uint64_t k, double delta, double lamb, double beta, const double *lg, double out[2]) {
  double lprob, lhs;
  if (delta > 0) {
    lprob = (N + 1) * log(lamb) - delta * (lamb + beta) + k * log(beta) - lg[k + 1];
    double pois = -INFINITY;
    for (uint64_t i = 0; i <= N; i++) pois = lae(i * log(lamb * delta) - lg[i + 1], pois);
    pois -= lamb * delta;
    lprob -= pois;
    double integ = -INFINITY;
    for (uint64_t i = 0; i <= N + k; i++)
      integ = lae(lg[N + k + 1] - lg[i + 1] - lg[N + k - i + 1] + (N + k - i) * log(delta) + lg[i + 1] - (i + 1) * log(lamb + beta), integ);
    integ -= lg[N + 1];
    lhs = lprob;
    lprob += integ;
  } else {
    lprob = (N + 1) * log(lamb) + k * log(beta) + lg[N + k + 1] - lg[N + 1] - lg[k + 1] - (N + k + 1) * log(lamb + beta);
    lhs = lprob;
  }
  out[0] = lprob; out[1] = lhs;
}

/* src/transcluster.hpp:131-170 */
static void lprob2(uint64_t N, uint64_t k, double delta, double lamb, double beta, const double *lg, double *lprob_o, double *lhs_o) {
  double lprob, lhs;
  if (delta > 0) {
    lprob = (N + 1) * log(lamb) + k * log(beta) + lg[N + k + 1];
    lprob = lprob - lg[N + 1] - lg[k + 1] - delta * beta;
    double pois = -INFINITY;
    for (uint64_t i = 0; i <= N; i++) pois = lae(i * log(lamb * delta) - lg[i + 1], pois);
    lprob -= pois;
    double integ = -INFINITY;
    for (uint64_t i = 0; i <= N + k; i++)
      integ = lae((N + k - i) * log(delta) - lg[N + k - i + 1] - (i + 1) * log(lamb + beta), integ);
    lhs = lprob;
    lprob += integ;
  } else {
    lprob = (N + 1) * log(lamb) + k * log(beta) + lg[N + k + 1] - lg[N + 1] - lg[k + 1] - (N + k + 1) * log(lamb + beta);
    lhs = lprob;
  }
  *lprob_o = lprob; *lhs_o = lhs;
}
void orc_lprob_k_given_N_2(uint64_t N, uint64_t k, double delta, double lamb, double beta, double out[2]) {
  uint64_t n = N + k + 3;
  double *lg = (double *)malloc(n * sizeof(double));
  for (uint64_t i = 0; i < n; i++) lg[i] = lgamma((double)i);
  lprob2(N, k, delta, lamb, beta, lg, &out[0], &out[1]);
  free(lg);
}

/* src/transcluster.hpp:173-188 */
static double upper_bound_E(const double *lg, double delta, double lamb, double beta, uint64_t N) {
  double pois = -INFINITY;
  for (uint64_t i = 0; i <= N; i++) pois = lae(i * log(lamb * delta) - lg[i + 1], pois);
  return exp(log(beta) + delta * lamb + log((double)(N + 1)) - (log(lamb) + pois));
}

/* src/transcluster.hpp:191-238. Stopping rule restated as the SHIPPED (-ffast-math, setup.py:46)
 * build behaves: a NaN bound (delta == 0: 0*log(0)) never satisfies the exit test, so the loop
 * runs to k = 9999 and the series converges to (N+1)*beta/lamb [SURVEY F6, probed]. The lgamma
 * table here is long enough that no access is out of bounds (the reference's is 10000 entries
 * and is over-read in that case). *k_exit reports the k at which the loop stopped. */
static double expected_k(int N, double delta, double lamb, double beta, double thr, const double *lg, int *k_exit) {
  double lprob = -INFINITY, elprob = -INFINITY;
  double ub = upper_bound_E(lg, delta, lamb, beta, (uint64_t)N);
  double diff = thr + 1;
  int k = 1;
  while ((diff > thr || isnan(diff)) && k < 10000) {
    double lp, lhs;
    lprob2((uint64_t)N, (uint64_t)k, delta, lamb, beta, lg, &lp, &lhs);
    lprob = lae(lprob, lp + log((double)k));
    elprob = lae(elprob, lhs + log((double)k) + delta * (lamb + beta) - (N + k + 1) * log(lamb + beta));
    diff = ub - exp(elprob);
    k++;
  }
  if (k_exit) *k_exit = k;
  return exp(lprob);
}

/* src/transcluster.hpp:240-287. Memoised on (N, delta) exactly like the reference (here by a
 * sort-free linear probe over previously seen keys hashed on the bit patterns).
 * p0 = log P(k=0 | N, delta); eK = E[K]. k_exit (optional, may be NULL) = loop exit k per edge,
 * used by the tests to delimit the reference's valid domain (N + k_exit + 1 < 10000). */
int orc_trans_dist(const int32_t *snp, const double *dt, uint64_t n, double lamb, double beta, double thr,
                   double *p0, double *eK, int32_t *k_exit) {
  int32_t maxN = 0;
  for (uint64_t i = 0; i < n; i++) { if (snp[i] < 0) return 1; if (snp[i] > maxN) maxN = snp[i]; }
  uint64_t nlg = (uint64_t)maxN + 10000 + 8;
  double *lg = (double *)malloc(nlg * sizeof(double));
  for (uint64_t i = 0; i < nlg; i++) lg[i] = lgamma((double)i);
  uint64_t cap = 1; while (cap < 2 * n + 16) cap <<= 1;
  int64_t *slot = (int64_t *)malloc(cap * sizeof(int64_t));
  for (uint64_t i = 0; i < cap; i++) slot[i] = -1;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t bits; memcpy(&bits, &dt[i], 8);
    uint64_t h = (bits * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)snp[i] * 0xC2B2AE3D27D4EB4Full);
    h ^= h >> 29;
    uint64_t s = h & (cap - 1);
    int64_t hit = -1;
    while (slot[s] >= 0) {
      int64_t o = slot[s];
      if (snp[o] == snp[i] && memcmp(&dt[o], &dt[i], 8) == 0) { hit = o; break; }
      s = (s + 1) & (cap - 1);
    }
    if (hit >= 0) { p0[i] = p0[hit]; eK[i] = eK[hit]; if (k_exit) k_exit[i] = k_exit[hit]; continue; }
    slot[s] = (int64_t)i;
    int ke = 0;
    eK[i] = expected_k(snp[i], dt[i], lamb, beta, thr, lg, &ke);
    if (k_exit) k_exit[i] = ke;
    double lp, lhs;
    lprob2((uint64_t)snp[i], 0, dt[i], lamb, beta, lg, &lp, &lhs);
    p0[i] = lp;
  }
  free(slot); free(lg);
  return 0;
}
