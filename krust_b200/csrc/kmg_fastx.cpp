// Host utility: FASTA / FASTQ record splitter for hosts that do not bring the Rust reader layer
// (the C++ CLI and the Python harness).  It stands in for rust-bio 3.0.0's
// bio::io::fasta::Reader / bio::io::fastq::Reader as used at src/reader.rs:91,96,176,181 and
// follows their observable behaviour:
//   * a record starts at a line beginning with '>' (FASTA) or '@' (FASTQ), anything else there is an error;
//   * every sequence / quality line is trim_end()-ed and the pieces are concatenated;
//   * a FASTA record runs to the next '>' line or EOF;  a FASTQ record has sequence lines up to the
//     '+' line followed by the same number of quality lines;
//   * reading stops at the first record that is completely empty (no id, no sequence).
// Pure host code, no CUDA; output is the back-to-back record layout kmg_count_ascii() consumes.
#include <cstdio>
#include <cstring>

#include "../../include/kmerust_gpu.h"

namespace {

struct Line {
  const uint8_t *p;
  uint64_t raw;      // length including the terminating '\n' (0 at EOF)
  uint64_t trimmed;  // length after trim_end()
};

struct Cursor {
  const uint8_t *cur, *end;
  Line next() {
    Line l{cur, 0, 0};
    if (cur == end) return l;
    const void *nl = memchr(cur, '\n', (size_t)(end - cur));
    const uint8_t *stop = nl ? (const uint8_t *)nl + 1 : end;
    l.raw = (uint64_t)(stop - cur);
    uint64_t t = l.raw;
    while (t > 0) {  // ASCII subset of Rust's White_Space: ' ' and \t \n \v \f \r
      const uint8_t b = cur[t - 1];
      if (b == ' ' || (b >= 9 && b <= 13)) --t; else break;
    }
    l.trimmed = t;
    cur = stop;
    return l;
  }
};

void set_err(char *buf, size_t n, const char *msg, uint64_t rec) {
  if (buf && n) snprintf(buf, n, "%s (record %llu)", msg, (unsigned long long)rec);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) kmg_status kmg_parse_fastx(
    const uint8_t *buf, uint64_t len, int is_fastq, uint8_t *seq_out, uint8_t *qual_out, uint64_t *offsets_out,
    uint64_t max_records, uint64_t *n_records_out, char *errbuf, size_t errbuf_len) {
  if (!offsets_out || !n_records_out || (len && (!buf || !seq_out))) return KMG_ERR_INVALID_ARG;
  Cursor c{buf, buf + len};
  uint64_t n = 0, w = 0;
  offsets_out[0] = 0;
  *n_records_out = 0;
  const uint8_t marker = is_fastq ? '@' : '>';
  Line l = c.next();
  while (l.raw) {
    if (l.p[0] != marker) {
      set_err(errbuf, errbuf_len, is_fastq ? "expected '@' at record start" : "expected '>' at record start", n);
      return KMG_ERR_PARSE;
    }
    const bool has_id = l.trimmed > 1;
    const uint64_t w0 = w;
    if (!is_fastq) {
      for (l = c.next(); l.raw && l.p[0] != '>'; l = c.next()) {
        memcpy(seq_out + w, l.p, l.trimmed);
        w += l.trimmed;
      }
    } else {
      uint64_t n_lines = 0;
      for (l = c.next();; l = c.next()) {
        if (!l.raw) { set_err(errbuf, errbuf_len, "incomplete FASTQ record", n); return KMG_ERR_PARSE; }
        if (l.p[0] == '+') break;
        memcpy(seq_out + w, l.p, l.trimmed);
        w += l.trimmed; ++n_lines;
      }
      uint64_t q = w0;
      for (uint64_t i = 0; i < n_lines; ++i) {
        l = c.next();
        if (q + l.trimmed > w) { set_err(errbuf, errbuf_len, "quality longer than sequence", n); return KMG_ERR_PARSE; }
        if (qual_out) memcpy(qual_out + q, l.p, l.trimmed);
        q += l.trimmed;
      }
      if (q != w) { set_err(errbuf, errbuf_len, "sequence and quality lengths differ", n); return KMG_ERR_PARSE; }
      l = c.next();
    }
    if (!has_id && w == w0) break;  // empty record ends the iteration
    if (n == max_records) { set_err(errbuf, errbuf_len, "more records than offsets_out can hold", n); return KMG_ERR_CAPACITY; }
    offsets_out[++n] = w;
  }
  *n_records_out = n;
  return KMG_OK;
}
