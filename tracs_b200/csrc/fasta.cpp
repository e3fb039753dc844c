// FASTA / FASTQ (.gz or plain) record reader for the pair sweep.
//
// Behavioural contract = the klib kseq reader the reference uses (reference src/kseq.h:170-208,
// driven from src/pairsnp.hpp:75-99), re-implemented as a chunked state machine:
//   * a record starts at the first '>' or '@' seen while looking for a header;
//   * name   = header bytes up to the first whitespace; the rest of the line is ignored;
//   * seq    = every printable non-space byte (33..126) up to the next '>', '@' or '+' ANYWHERE;
//   * '+'    = FASTQ: skip that line, then consume as many quality bytes as sequence bytes;
//   * all records of one file must have equal length (src/pairsnp.hpp:94-98).
#include <zlib.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace tracs {

namespace {
enum State { SEEK, NAME, REST_OF_HEADER, SEQ, PLUS_LINE, QUAL, QUAL_TRAIL };

inline bool is_space(unsigned c) { return c == ' ' || (c >= 9 && c <= 13); }
}  // namespace

uint64_t read_fasta(const char *path, int /*n_threads*/, std::vector<uint8_t> &ascii, std::vector<std::string> &names,
                    uint64_t &L_io) {
  gzFile f = gzopen(path, "r");
  if (!f) throw std::runtime_error("Error reading FASTA!");
  gzbuffer(f, 1 << 20);
  std::vector<unsigned char> buf(size_t(1) << 22);
  State st = SEEK;
  std::string name;
  uint64_t count = 0, L = 0;
  size_t rec_start = ascii.size();  // where the current record's bases begin in `ascii`
  uint64_t qual_seen = 0;
  bool name_started = false;
  bool failed_len = false, truncated = false;

  auto finish_record = [&]() {
    uint64_t len = ascii.size() - rec_start;
    if (count > 0 && len != L) failed_len = true;
    L = len;
    names.push_back(name);
    count++;
    rec_start = ascii.size();
  };

  for (;;) {
    int got = gzread(f, buf.data(), (unsigned)buf.size());
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
          const unsigned char *q = p;
          while (q < end) {
            unsigned c = *q;
            if (c == '>' || c == '@' || c == '+') break;
            ++q;
          }
          // append printable bytes of [p, q)
          size_t old = ascii.size();
          ascii.resize(old + (q - p));
          uint8_t *o = ascii.data() + old;
          for (const unsigned char *r = p; r < q; ++r) {
            unsigned c = *r;
            *o = (uint8_t)c;
            o += (c >= 33 && c <= 126);
          }
          ascii.resize(o - ascii.data());
          p = q;
          if (p < end) {
            unsigned c = *p++;
            if (c == '+') {
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
          uint64_t need = ascii.size() - rec_start;
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
      case QUAL: truncated = (qual_seen != ascii.size() - rec_start); if (!truncated) finish_record(); break;
      case QUAL_TRAIL: finish_record(); break;
    }
  }
  if (failed_len) throw std::runtime_error("Error reading FASTA, variable sequence lengths!");
  if (truncated) throw std::runtime_error("Error reading FASTA!");
  if (count > 0) L_io = L;
  return count;
}

}  // namespace tracs
