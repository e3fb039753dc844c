// FASTA / FASTQ (.gz or plain) record reader for the pair sweep.
//
// Behavioural contract = the klib kseq reader the reference uses (reference src/kseq.h:170-208,
// driven from src/pairsnp.hpp:75-99), re-implemented as a chunked state machine:
//   * a record starts at the first '>' or '@' seen while looking for a header;
//   * name   = header bytes up to the first whitespace; the rest of the line is ignored;
//   * seq    = every printable non-space byte (33..126) up to the next '>', '@' or '+' ANYWHERE;
//   * '+'    = FASTQ: skip that line, then consume as many quality bytes as sequence bytes;
//   * all records of one file must have equal length (src/pairsnp.hpp:94-98).
#include <emmintrin.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace tracs {

namespace {
enum State { SEEK, NAME, REST_OF_HEADER, SEQ, PLUS_LINE, QUAL, QUAL_TRAIL };

inline bool is_space(unsigned c) { return c == ' ' || (c >= 9 && c <= 13); }
}  // namespace

uint64_t read_fasta(const char *path, int /*n_threads*/, ByteBuf &ascii, std::vector<std::string> &names, uint64_t &L_io) {
  gzFile f = gzopen(path, "r");
  if (!f) throw std::runtime_error("Error reading FASTA!");
  {
    // size the output once: a plain file cannot hold more bases than bytes; a gzip stream of
    // nucleotides rarely inflates more than ~4.5x
    struct stat sb;
    if (stat(path, &sb) == 0 && sb.st_size > 0) ascii.reserve(ascii.size() + (gzdirect(f) ? (size_t)sb.st_size : (size_t)sb.st_size * 9 / 2) + 64);
  }
  gzbuffer(f, 1 << 20);
  std::vector<unsigned char> buf((size_t(1) << 22) + 16);
  State st = SEEK;
  std::string name;
  uint64_t count = 0, L = 0;
  size_t len = ascii.size();        // bytes of `ascii` in use
  size_t rec_start = len;           // where the current record's bases begin in `ascii`
  uint64_t qual_seen = 0;
  bool name_started = false;
  bool failed_len = false, truncated = false;

  auto finish_record = [&]() {
    const uint64_t rec_len = len - rec_start;
    if (count > 0 && rec_len != L) failed_len = true;
    L = rec_len;
    names.push_back(name);
    count++;
    rec_start = len;
  };

  for (;;) {
    int got = gzread(f, buf.data(), (unsigned)(buf.size() - 16));
    if (got < 0) {
      gzclose(f);
      throw std::runtime_error("Error reading FASTA!");
    }
    if (got == 0) break;
    const unsigned char *p = buf.data(), *end = p + got;
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
          ascii.reserve(len + (size_t)(end - p) + 16);
          uint8_t *o = ascii.data() + len;
          const __m128i c_gt = _mm_set1_epi8('>'), c_at = _mm_set1_epi8('@'), c_pl = _mm_set1_epi8('+');
          const __m128i c33 = _mm_set1_epi8(33), c126 = _mm_set1_epi8(126);
          bool ended = false;
          unsigned endc = 0;
          while (p < end) {
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
    if (failed_len) break;
  }
  gzclose(f);
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
  if (count > 0) L_io = L;
  return count;
}

}  // namespace tracs
