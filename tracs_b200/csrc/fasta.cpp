// FASTA / FASTQ (.gz or plain) record reader for the pair sweep.
//
// Behavioural contract = the klib kseq reader the reference uses (reference src/kseq.h:170-208,
// driven from src/pairsnp.hpp:75-99), re-implemented as a chunked state machine:
//   * a record starts at the first '>' or '@' seen while looking for a header;
//   * name   = header bytes up to the first whitespace; the rest of the line is ignored;
//   * seq    = every printable non-space byte (33..126) up to the next '>', '@' or '+' ANYWHERE;
//   * '+'    = FASTQ: skip that line, then consume as many quality bytes as sequence bytes;
//   * all records of one file must have equal length (src/pairsnp.hpp:94-98).
// Plain files are parsed by n_threads workers (read_fasta_parallel below) when that is provably the same.
#include <emmintrin.h>
#include <immintrin.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace tracs {

namespace {
enum State { SEEK, NAME, REST_OF_HEADER, SEQ, PLUS_LINE, QUAL, QUAL_TRAIL };

inline bool is_space(unsigned c) { return c == ' ' || (c >= 9 && c <= 13); }

// 32 bytes at a time where the CPU has AVX2 (run-time check; the SSE2 / scalar code below it handles whatever is
// left). Copies sequence bytes from p to o, dropping non-printable ones; returns at a record delimiter ('>', '@',
// '+'), which is left for the caller, or when fewer than 32 input bytes / 32 bytes of room (o_end, may be null =
// unlimited) remain. Every store writes 32 bytes at o, of which only the clean prefix is kept.
__attribute__((target("avx2"))) void bulk_copy_avx2(const unsigned char *&p, const unsigned char *end, uint8_t *&o, const uint8_t *o_end) {
  const __m256i c_gt = _mm256_set1_epi8('>'), c_at = _mm256_set1_epi8('@'), c_pl = _mm256_set1_epi8('+');
  const __m256i c33 = _mm256_set1_epi8(33), c126 = _mm256_set1_epi8(126);
  while (end - p >= 32 && (!o_end || o_end - o >= 32)) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
    const __m256i ok_lo = _mm256_cmpeq_epi8(_mm256_max_epu8(v, c33), v);   // v >= 33
    const __m256i ok_hi = _mm256_cmpeq_epi8(_mm256_min_epu8(v, c126), v);  // v <= 126
    const __m256i delim = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, c_gt), _mm256_cmpeq_epi8(v, c_at)), _mm256_cmpeq_epi8(v, c_pl));
    const __m256i good = _mm256_andnot_si256(delim, _mm256_and_si256(ok_lo, ok_hi));
    const unsigned bad = ~(unsigned)_mm256_movemask_epi8(good);
    _mm256_storeu_si256(reinterpret_cast<__m256i *>(o), v);
    if (!bad) {
      o += 32;
      p += 32;
      continue;
    }
    const unsigned k = (unsigned)__builtin_ctz(bad);
    o += k;
    p += k;
    const unsigned c = *p;
    if (c == '>' || c == '@' || c == '+') return;
    ++p;  // a dropped byte (newline, CR, blank, ...)
  }
}
const bool g_have_avx2 = __builtin_cpu_supports("avx2") != 0;
}  // namespace

// The sequential reader: a state machine fed with consecutive byte ranges of the (decompressed) file.
struct FastaMachine {
  ByteBuf &ascii;
  std::vector<std::string> &names;
  RowSink *sink = nullptr;      // completed records are handed over in batches while parsing goes on
  uint64_t sink_row0 = 0;       // global number of this file's first record
  uint64_t sunk = 0;            // records of this file already handed over
  size_t base_len = 0;          // bytes of `ascii` in use before this file
  State st = SEEK;
  std::string name;
  uint64_t count = 0, L = 0;
  size_t len;        // bytes of `ascii` in use
  size_t rec_start;  // where the current record's bases begin in `ascii`
  uint64_t qual_seen = 0;
  bool name_started = false, failed_len = false, truncated = false;

  FastaMachine(ByteBuf &a, std::vector<std::string> &n) : ascii(a), names(n), base_len(a.size()), len(a.size()), rec_start(a.size()) {}

  // records [sunk, count) are complete and equally long: pass them on (at least `min_rows` at a time)
  void flush_sink(uint64_t min_rows) {
    if (!sink || failed_len || L == 0 || count - sunk < min_rows || count == sunk) return;
    sink->rows(ascii.data() + base_len + sunk * L, sink_row0 + sunk, count - sunk, L);
    sunk = count;
  }

  void finish_record() {
    const uint64_t rec_len = len - rec_start;
    if (count > 0 && rec_len != L) failed_len = true;
    L = rec_len;
    names.push_back(name);
    count++;
    rec_start = len;
    flush_sink(std::max<uint64_t>(1, ((uint64_t)32 << 20) / std::max<uint64_t>(1, L)));  // ~32 MB batches
  }

  void feed(const unsigned char *p, const unsigned char *end) {
    while (p < end && !failed_len) {
      switch (st) {
        case SEEK:
          while (p < end && *p != '>' && *p != '@') ++p;
          if (p < end) {
            ++p;
            st = NAME;
            name.clear();
            name_started = false;
          }
          break;
        case NAME:
          while (p < end && !is_space(*p)) {
            name.push_back((char)*p++);
            name_started = true;
          }
          if (p < end) {
            name_started = true;
            st = (*p == '\n') ? SEQ : REST_OF_HEADER;
            ++p;
          }
          break;
        case REST_OF_HEADER:
          while (p < end && *p != '\n') ++p;
          if (p < end) {
            ++p;
            st = SEQ;
          }
          break;
        case SEQ: {
          // Bulk copy of sequence bytes, 16 at a time (SSE2). A byte is "special" if it ends the
          // sequence ('>', '@', '+') or is not printable (newline, CR, space, ...): those are dropped.
          ascii.reserve(len + (size_t)(end - p) + 32);
          uint8_t *o = ascii.data() + len;
          const __m128i c_gt = _mm_set1_epi8('>'), c_at = _mm_set1_epi8('@'), c_pl = _mm_set1_epi8('+');
          const __m128i c33 = _mm_set1_epi8(33), c126 = _mm_set1_epi8(126);
          bool ended = false;
          unsigned endc = 0;
          while (p < end) {
            if (g_have_avx2) {
              bulk_copy_avx2(p, end, o, nullptr);
              if (p >= end) break;
            }
            if (end - p >= 16) {
              const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p));
              const __m128i ok_lo = _mm_cmpeq_epi8(_mm_max_epu8(v, c33), v);    // v >= 33
              const __m128i ok_hi = _mm_cmpeq_epi8(_mm_min_epu8(v, c126), v);   // v <= 126
              const __m128i delim = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, c_gt), _mm_cmpeq_epi8(v, c_at)), _mm_cmpeq_epi8(v, c_pl));
              const __m128i good = _mm_andnot_si128(delim, _mm_and_si128(ok_lo, ok_hi));
              const unsigned bad = (~(unsigned)_mm_movemask_epi8(good)) & 0xFFFFu;
              if (!bad) {
                _mm_storeu_si128(reinterpret_cast<__m128i *>(o), v);
                o += 16;
                p += 16;
                continue;
              }
              const unsigned k = (unsigned)__builtin_ctz(bad);  // bytes before the first special one
              _mm_storeu_si128(reinterpret_cast<__m128i *>(o), v);  // over-copy, only k bytes are kept
              o += k;
              p += k;
            }
            const unsigned c = *p;
            if (c == '>' || c == '@' || c == '+') {
              ended = true;
              endc = c;
              ++p;
              break;
            }
            *o = (uint8_t)c;
            o += (c >= 33 && c <= 126);
            ++p;
          }
          len = (size_t)(o - ascii.data());
          if (ended) {
            if (endc == '+') {
              st = PLUS_LINE;
            } else {
              finish_record();
              st = NAME;
              name.clear();
              name_started = false;
            }
          }
          break;
        }
        case PLUS_LINE:
          while (p < end && *p != '\n') ++p;
          if (p < end) {
            ++p;
            st = QUAL;
            qual_seen = 0;
          }
          break;
        case QUAL: {
          uint64_t need = len - rec_start;
          while (p < end && qual_seen < need) {
            unsigned c = *p++;
            if (c >= 33 && c <= 127) qual_seen++;
          }
          if (qual_seen >= need) st = QUAL_TRAIL;
          break;
        }
        case QUAL_TRAIL:
          // kseq consumes one more byte before it notices the quality string is complete
          ++p;
          finish_record();
          st = SEEK;
          break;
      }
    }
  }

  // end of stream
  void finish() {
    if (!failed_len) {
      switch (st) {
        case SEEK: break;
        case NAME:
          if (name_started) finish_record();  // header only, stream ended: empty sequence
          break;
        case REST_OF_HEADER:
        case SEQ: finish_record(); break;
        case PLUS_LINE: truncated = true; break;
        case QUAL: truncated = (qual_seen != len - rec_start); if (!truncated) finish_record(); break;
        case QUAL_TRAIL: finish_record(); break;
      }
    }
    ascii.len = len;
    if (failed_len) throw std::runtime_error("Error reading FASTA, variable sequence lengths!");
    if (truncated) throw std::runtime_error("Error reading FASTA!");
    flush_sink(1);
  }
};

// ---- parallel reader for plain (uncompressed) multi-FASTA --------------------------------------------
// Records are located first (every '>' at the start of a line), then parsed by a pool of threads straight
// into their final place. This equals the sequential machine exactly when (a) the file starts with '>',
// and (b) no '>', '@' or '+' occurs inside a sequence (kseq would start a record / a quality block there):
// both are checked while parsing, and the reader falls back to the sequential machine otherwise. The last
// record goes through the sequential machine itself (end-of-stream rules). Names, order and bases are
// identical to the sequential result (tests/test_fasta_property.py runs both).
namespace {

// one record [p, end) that starts with '>' and is followed by another record: header line, then bases.
// false = not a simple record. `nb` counts the bases; only the first `cap` are stored.
bool parse_simple_record(const unsigned char *p, const unsigned char *end, std::string &name, uint8_t *out, uint64_t cap,
                         uint64_t &nb) {
  ++p;
  const unsigned char *q = p;
  while (q < end && !is_space(*q)) ++q;
  if (q == end) return false;
  name.assign(reinterpret_cast<const char *>(p), (size_t)(q - p));
  if (*q != '\n') {
    q = static_cast<const unsigned char *>(memchr(q, '\n', (size_t)(end - q)));
    if (!q) return false;
  }
  p = q + 1;
  uint8_t *o = out, *const o_end = out + cap;
  uint64_t extra = 0;  // bases beyond cap
  const __m128i c_gt = _mm_set1_epi8('>'), c_at = _mm_set1_epi8('@'), c_pl = _mm_set1_epi8('+');
  const __m128i c33 = _mm_set1_epi8(33), c126 = _mm_set1_epi8(126);
  while (p < end) {
    if (g_have_avx2 && out) {
      bulk_copy_avx2(p, end, o, o_end);
      if (p >= end) break;
    }
    if (end - p >= 16 && o_end - o >= 16) {  // the 16-byte store must stay inside this record's slot
      const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p));
      const __m128i ok_lo = _mm_cmpeq_epi8(_mm_max_epu8(v, c33), v);
      const __m128i ok_hi = _mm_cmpeq_epi8(_mm_min_epu8(v, c126), v);
      const __m128i delim = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, c_gt), _mm_cmpeq_epi8(v, c_at)), _mm_cmpeq_epi8(v, c_pl));
      const __m128i good = _mm_andnot_si128(delim, _mm_and_si128(ok_lo, ok_hi));
      const unsigned bad = (~(unsigned)_mm_movemask_epi8(good)) & 0xFFFFu;
      _mm_storeu_si128(reinterpret_cast<__m128i *>(o), v);
      if (!bad) {
        o += 16;
        p += 16;
        continue;
      }
      const unsigned k = (unsigned)__builtin_ctz(bad);
      o += k;
      p += k;
    }
    const unsigned c = *p++;
    if (c == '>' || c == '@' || c == '+') return false;
    if (c >= 33 && c <= 126) {
      if (o < o_end) *o++ = (uint8_t)c; else ++extra;
    }
  }
  nb = (uint64_t)(o - out) + extra;
  return true;
}

struct Mapped {
  const unsigned char *p = nullptr;
  size_t size = 0;
  int fd = -1;
  ~Mapped() {
    if (p) munmap(const_cast<unsigned char *>(p), size);
    if (fd >= 0) close(fd);
  }
};

size_t parallel_min_bytes() {
  if (const char *e = getenv("TRACS_FASTA_PAR_MIN")) return (size_t)strtoull(e, nullptr, 10);
  return (size_t)8 << 20;
}

// returns false when the file is not handled here (nothing appended); throws like the sequential reader
bool read_fasta_parallel(const char *path, int n_threads, ByteBuf &ascii, std::vector<std::string> &names, uint64_t &L_io,
                         uint64_t &count_out, RowSink *sink, uint64_t row0) {
  Mapped m;
  m.fd = open(path, O_RDONLY);
  if (m.fd < 0) return false;
  struct stat sb;
  if (fstat(m.fd, &sb) != 0 || !S_ISREG(sb.st_mode) || (size_t)sb.st_size < std::max<size_t>(2, parallel_min_bytes())) return false;
  m.size = (size_t)sb.st_size;
  void *mp = mmap(nullptr, m.size, PROT_READ, MAP_PRIVATE, m.fd, 0);
  if (mp == MAP_FAILED) return false;
  m.p = static_cast<const unsigned char *>(mp);
  madvise(mp, m.size, MADV_SEQUENTIAL);
  const unsigned char *d = m.p;
  if (d[0] != '>') return false;
  const int T = std::max(1, std::min(n_threads, (int)std::max(1u, std::thread::hardware_concurrency())));

  // 1. record starts: '>' at the start of a line
  std::vector<std::vector<size_t>> found((size_t)T);
  std::atomic<int> scan_failed(0);
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
      pool.emplace_back([&, t] {
        try {
          const size_t lo = m.size / T * t, hi = (t == T - 1) ? m.size : m.size / T * (t + 1);
          const unsigned char *q = d + lo;
          while (q < d + hi) {
            q = static_cast<const unsigned char *>(memchr(q, '>', (size_t)(d + hi - q)));
            if (!q) break;
            if (q == d || q[-1] == '\n') found[t].push_back((size_t)(q - d));
            ++q;
          }
        } catch (...) {
          scan_failed.store(1);
        }
      });
    for (auto &th : pool) th.join();
  }
  if (scan_failed.load()) return false;
  std::vector<size_t> starts;
  for (auto &v : found) starts.insert(starts.end(), v.begin(), v.end());
  const size_t R = starts.size();  // >= 1

  // 2. length from record 0
  uint64_t L = 0;
  if (R >= 2) {
    std::string nm;
    if (!parse_simple_record(d + starts[0], d + starts[1], nm, nullptr, 0, L)) return false;
  }

  // 3. the last record through the sequential machine (end-of-stream rules; it may hold FASTQ or several
  //    records). With records in front of it, the machine starts as if it had just finished one of length L,
  //    so that it raises the errors the sequential pass would, in the same order.
  ByteBuf tail;
  std::vector<std::string> tail_names;
  FastaMachine tm(tail, tail_names);
  if (R >= 2) {
    tm.count = 1;
    tm.L = L;
  }
  tm.feed(d + starts[R - 1], d + m.size);
  std::exception_ptr tail_error;  // raised only after the records in front of it turned out fine
  try {
    tm.finish();
  } catch (...) {
    tail_error = std::current_exception();
  }
  if (R == 1) {
    if (tail_error) std::rethrow_exception(tail_error);
    ascii.reserve(ascii.size() + tail.size() + 16);
    memcpy(ascii.data() + ascii.size(), tail.data(), tail.size());
    if (sink && tm.count > 0 && tm.L > 0) sink->rows(ascii.data() + ascii.size(), row0, tm.count, tm.L);
    ascii.len += tail.size();
    names.insert(names.end(), tail_names.begin(), tail_names.end());
    if (tm.count > 0) L_io = tm.L;
    count_out = tm.count;
    return true;
  }
  const uint64_t tail_count = tm.count - 1;

  // 4. everything before the last record in parallel, in place
  const size_t base = ascii.size();
  const size_t n_par = R - 1;
  ascii.reserve(base + (n_par + tail_count) * L + 16);
  std::vector<std::string> par_names(n_par);
  std::atomic<size_t> next(0);
  std::atomic<int> not_simple(0), bad_len(0), workers_left(T);
  // rows are handed to the sink in chunks (~32 MB) as soon as every record of a chunk is parsed; the workers take
  // the records in file order, so chunks complete (nearly) in order while later ones are still being parsed
  const size_t chunk_rows = std::max<size_t>(1, ((size_t)32 << 20) / std::max<uint64_t>(1, L));
  const size_t n_chunks = (n_par + chunk_rows - 1) / chunk_rows;
  std::vector<std::atomic<uint32_t>> chunk_done(sink ? n_chunks : 0);
  for (auto &c : chunk_done) c.store(0);
  if (sink && L > 0) sink->expect(row0 + n_par + tail_count, L);
  {
    std::vector<std::thread> pool;
    uint8_t *out0 = ascii.data() + base;
    for (int t = 0; t < T; ++t)
      pool.emplace_back([&] {
        try {
          for (;;) {
            const size_t r = next.fetch_add(1);
            if (r >= n_par || not_simple.load(std::memory_order_relaxed)) break;
            uint64_t nb = 0;
            if (!parse_simple_record(d + starts[r], d + starts[r + 1], par_names[r], out0 + r * L, L, nb)) {
              not_simple.store(1);
              break;
            }
            if (nb != L) bad_len.store(1);
            if (sink) chunk_done[r / chunk_rows].fetch_add(1, std::memory_order_release);
          }
        } catch (...) {  // out of memory while storing a name: let the sequential reader report it
          not_simple.store(1);
        }
        workers_left.fetch_sub(1);
      });
    bool sunk_any = false;
    std::exception_ptr sink_error;
    if (sink && L > 0) {
      for (size_t c = 0; c < n_chunks; ++c) {
        const size_t r0 = c * chunk_rows, nr = std::min(chunk_rows, n_par - r0);
        while (chunk_done[c].load(std::memory_order_acquire) < nr && workers_left.load() > 0 && !not_simple.load() && !bad_len.load())
          std::this_thread::sleep_for(std::chrono::microseconds(50));
        if (chunk_done[c].load(std::memory_order_acquire) < nr || not_simple.load() || bad_len.load()) break;
        try {
          sink->rows(out0 + r0 * L, row0 + r0, nr, L);
          sunk_any = true;
        } catch (...) {
          sink_error = std::current_exception();
          not_simple.store(1);  // stops the workers
          break;
        }
      }
    }
    for (auto &th : pool) th.join();
    if (sink_error) std::rethrow_exception(sink_error);
    if (sink && (not_simple.load() || bad_len.load() || tail_error) && sunk_any) sink->reset();
    else if (sink && L > 0 && !not_simple.load() && !bad_len.load() && !tail_error) {
      // chunks the loop above did not reach (it only breaks on failure), then the records of the tail
      if (tail_count) sink->rows(tail.data(), row0 + n_par, tail_count, L);
    }
  }
  if (not_simple.load()) return false;  // `ascii.len` was never advanced: nothing appended
  if (bad_len.load()) throw std::runtime_error("Error reading FASTA, variable sequence lengths!");
  if (tail_error) std::rethrow_exception(tail_error);
  memcpy(ascii.data() + base + n_par * L, tail.data(), tail.size());
  ascii.len = base + n_par * L + tail.size();
  names.insert(names.end(), par_names.begin(), par_names.end());
  names.insert(names.end(), tail_names.begin(), tail_names.end());
  L_io = L;
  count_out = n_par + tail_count;
  return true;
}

}  // namespace

uint64_t read_fasta(const char *path, int n_threads, ByteBuf &ascii, std::vector<std::string> &names, uint64_t &L_io, RowSink *sink,
                    uint64_t row0) {
  gzFile f = gzopen(path, "r");
  if (!f) throw std::runtime_error("Error reading FASTA!");
  const bool plain = gzdirect(f) != 0;
  if (plain && n_threads > 1) {
    uint64_t cnt = 0;
    bool done = false;
    try {
      done = read_fasta_parallel(path, n_threads, ascii, names, L_io, cnt, sink, row0);
    } catch (...) {
      gzclose(f);
      throw;
    }
    if (done) {
      gzclose(f);
      return cnt;
    }
  }
  {
    // size the output once: a plain file cannot hold more bases than bytes; a gzip stream of
    // nucleotides rarely inflates more than ~4.5x
    struct stat sb;
    if (stat(path, &sb) == 0 && sb.st_size > 0) ascii.reserve(ascii.size() + (plain ? (size_t)sb.st_size : (size_t)sb.st_size * 9 / 2) + 64);
  }
  gzbuffer(f, 1 << 20);
  std::vector<unsigned char> buf((size_t(1) << 22) + 16);
  FastaMachine fm(ascii, names);
  fm.sink = sink;
  fm.sink_row0 = row0;
  for (;;) {
    int got = gzread(f, buf.data(), (unsigned)(buf.size() - 16));
    if (got < 0) {
      gzclose(f);
      throw std::runtime_error("Error reading FASTA!");
    }
    if (got == 0) break;
    fm.feed(buf.data(), buf.data() + got);
    if (fm.failed_len) break;
  }
  gzclose(f);
  fm.finish();
  if (fm.count > 0) L_io = fm.L;
  return fm.count;
}

}  // namespace tracs
