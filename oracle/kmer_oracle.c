/*
 * kmer_oracle.c -- CPU ORACLE for the canonical k-mer counting path of kmerust v0.3.1.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (krust_b200/csrc, libkmerust_gpu.so) never links, loads or calls it.
 *
 * It is a plain-C restatement of the reference algorithm; every function cites the
 * reference file:line (paths relative to /root/reference) whose results it reproduces.
 * The Rust reference cannot be compiled in this image (no cargo/rustc), so the oracle
 * is pinned against the reference's own known-answer tests instead (tests/test_oracle_golden.py,
 * list in SURVEY.md section 8c).
 *
 * Two independent counters are provided and cross-checked against each other:
 *   oracle #1 "literal"  : per-window validate -> pack -> bytewise canonical compare, with the
 *                          reference's skip-ahead loop (run.rs:526-571, kmer.rs:266-390);
 *   oracle #2 "rolling"  : rolling forward/reverse-complement words + min(), sort + run-length.
 * plus the multi-threaded "restated reference CPU path" used as the timed CPU baseline.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * k-mer primitive (src/kmer.rs)
 * ---------------------------------------------------------------------------------------- */

/* kmer.rs:78-81, :100-111  KmerLength::new -- valid k is 1..=32. */
ORC_API int orc_kmer_length_ok(uint64_t k) { return k >= 1 && k <= 32; }

/* kmer.rs:21-32 PACK_TABLE: A/a=0 C/c=1 G/g=2 T/t=3. Returns -1 for any other byte. */
static inline int orc_code(uint8_t b) {
  switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

/* kmer.rs:266-286 Kmer::from_sub: accept ACGTacgt, upper-case them into `out`; the first other
 * byte gives InvalidBaseError{base, position}.  Returns -1 when valid, else the position. */
ORC_API int64_t orc_from_sub(const uint8_t *sub, uint64_t k, uint8_t *out, uint8_t *bad_base) {
  for (uint64_t i = 0; i < k; i++) {
    uint8_t b = sub[i];
    switch (b) {
      case 'A': case 'C': case 'G': case 'T': out[i] = b; break;
      case 'a': case 'c': case 'g': case 't': out[i] = (uint8_t)(b - 32); break;
      default:
        if (bad_base) *bad_base = b;
        return (int64_t)i;
    }
  }
  return -1;
}

/* kmer.rs:467-471 pack_bytes: MSB-first fold acc = (acc << 2) | code. */
ORC_API uint64_t orc_pack_bytes(const uint8_t *bytes, uint64_t k) {
  uint64_t acc = 0;
  for (uint64_t i = 0; i < k; i++) acc = (acc << 2) | (uint64_t)orc_code(bytes[i]);
  return acc;
}

/* kmer.rs:36-47 COMPLEMENT_TABLE (upper-case output). */
static inline uint8_t orc_complement(uint8_t b) {
  switch (b) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    case 'T': case 't': return 'A';
    default: return 0;
  }
}

/* kmer.rs:348-390 Kmer<Packed>::canonical: compare forward with reverse complement byte by byte
 * from the left; Less -> forward, Greater -> reverse complement, all equal -> forward.
 * `norm` is the upper-cased k-mer.  Returns the packed bits of the winner; *is_rc says which. */
ORC_API uint64_t orc_canonical(const uint8_t *norm, uint64_t k, int *is_rc) {
  int use_rc = 0;
  for (uint64_t i = 0; i < k; i++) {
    uint8_t fwd = norm[i];
    uint8_t rc = orc_complement(norm[k - 1 - i]);
    if (fwd < rc) { use_rc = 0; break; }
    if (fwd > rc) { use_rc = 1; break; }
  }
  if (is_rc) *is_rc = use_rc;
  if (!use_rc) return orc_pack_bytes(norm, k);
  uint8_t tmp[32];
  for (uint64_t i = 0; i < k; i++) tmp[i] = orc_complement(norm[k - 1 - i]);
  return orc_pack_bytes(tmp, k);
}

/* kmer.rs:431-440 unpack_to_bytes: base i = (bits >> 2(k-1-i)) & 3 -> "ACGT". */
ORC_API void orc_unpack(uint64_t bits, uint64_t k, uint8_t *out) {
  static const uint8_t T[4] = {'A', 'C', 'G', 'T'};
  for (uint64_t i = 0; i < k; i++) out[i] = T[(bits >> ((k - 1 - i) * 2)) & 3];
}

/* ------------------------------------------------------------------------------------------
 * u64 -> u64 count map (stands in for DashMap<u64,u64,Fx>, run.rs:489; only its *contents*
 * matter for parity -- dashmap/rustc-hash affect speed and iteration order, never results).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t *keys;
  uint64_t *vals;
  uint8_t *used;
  uint64_t cap;  /* power of two */
  uint64_t n;
} orc_map;

static inline uint64_t orc_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

static void orc_map_init(orc_map *m, uint64_t cap) {
  uint64_t c = 16;
  while (c < cap) c <<= 1;
  m->cap = c; m->n = 0;
  m->keys = (uint64_t *)malloc(c * 8);
  m->vals = (uint64_t *)malloc(c * 8);
  m->used = (uint8_t *)calloc(c, 1);
}
static void orc_map_free(orc_map *m) { free(m->keys); free(m->vals); free(m->used); memset(m, 0, sizeof *m); }

static void orc_map_add(orc_map *m, uint64_t key, uint64_t add);
static void orc_map_grow(orc_map *m) {
  orc_map old = *m;
  orc_map_init(m, old.cap * 2);
  for (uint64_t i = 0; i < old.cap; i++)
    if (old.used[i]) orc_map_add(m, old.keys[i], old.vals[i]);
  free(old.keys); free(old.vals); free(old.used);
}
/* run.rs:565-571: entry(key).and_modify(|c| *c = c.saturating_add(1)).or_insert(1). */
static void orc_map_add(orc_map *m, uint64_t key, uint64_t add) {
  if ((m->n + 1) * 10 > m->cap * 7) orc_map_grow(m);
  uint64_t mask = m->cap - 1, i = orc_mix(key) & mask;
  while (m->used[i]) {
    if (m->keys[i] == key) {
      uint64_t v = m->vals[i] + add;
      m->vals[i] = v < add ? UINT64_MAX : v; /* saturating_add */
      return;
    }
    i = (i + 1) & mask;
  }
  m->used[i] = 1; m->keys[i] = key; m->vals[i] = add; m->n++;
}

/* ------------------------------------------------------------------------------------------
 * counter object
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t k;
  int has_q;
  uint8_t min_quality;
  int mode;            /* 0 literal, 1 rolling */
  orc_map map;         /* literal mode */
  uint64_t *keys;      /* rolling mode: raw canonical keys */
  uint64_t nkeys, capkeys;
  uint64_t windows;    /* counted windows */
  uint64_t records, bases;
} orc_counter;

ORC_API orc_counter *orc_counter_new(uint32_t k, int has_min_quality, uint8_t min_quality, int mode) {
  if (!orc_kmer_length_ok(k)) return NULL;
  orc_counter *c = (orc_counter *)calloc(1, sizeof *c);
  c->k = k; c->has_q = has_min_quality; c->min_quality = min_quality; c->mode = mode;
  if (mode == 0) orc_map_init(&c->map, 1024);
  return c;
}
ORC_API void orc_counter_free(orc_counter *c) {
  if (!c) return;
  if (c->mode == 0) orc_map_free(&c->map);
  free(c->keys);
  free(c);
}

static inline uint8_t sat_add_u8(uint8_t a, uint8_t b) { unsigned s = (unsigned)a + b; return s > 255 ? 255 : (uint8_t)s; }

/* oracle #1 -- literal restatement of run.rs:526-563 (process_sequence_with_quality), identical
 * logic duplicated at streaming.rs:622-660 / :791-829 / :1068-1105. */
static void orc_add_literal(orc_counter *c, const uint8_t *seq, uint64_t len, const uint8_t *qual) {
  uint64_t k = c->k;
  if (len < k) return;                                             /* run.rs:533 */
  int have_thr = c->has_q && qual != NULL;
  uint8_t thr = sat_add_u8(c->min_quality, 33);                    /* run.rs:538 */
  uint8_t norm[32];
  uint64_t i = 0;
  while (i <= len - k) {                                           /* run.rs:541 */
    if (have_thr) {                                                /* run.rs:543-548 */
      int64_t bad = -1;
      for (uint64_t j = 0; j < k; j++) if (qual[i + j] < thr) { bad = (int64_t)j; break; }
      if (bad >= 0) { i += (uint64_t)bad + 1; continue; }
    }
    int64_t pos = orc_from_sub(seq + i, k, norm, NULL);            /* run.rs:550-552 */
    if (pos < 0) {
      uint64_t key = orc_canonical(norm, k, NULL);                 /* run.rs:565-566 */
      orc_map_add(&c->map, key, 1);                                /* run.rs:567-570 */
      c->windows++;
      i += 1;
    } else {
      i += (uint64_t)pos + 1;                                      /* run.rs:557-560 */
    }
  }
}

static void orc_push_key(orc_counter *c, uint64_t key) {
  if (c->nkeys == c->capkeys) {
    c->capkeys = c->capkeys ? c->capkeys * 2 : 4096;
    c->keys = (uint64_t *)realloc(c->keys, c->capkeys * 8);
  }
  c->keys[c->nkeys++] = key;
}

/* oracle #2 -- rolling form: window ending at e is counted iff the last k positions are all
 * ACGTacgt and (no filter or all k quals >= thr); key = min(fwd, rc) as unsigned integers
 * (equivalent to the bytewise compare because A<C<G<T holds in ASCII and in the 2-bit code). */
static void orc_add_rolling(orc_counter *c, const uint8_t *seq, uint64_t len, const uint8_t *qual) {
  uint64_t k = c->k;
  if (len < k) return;
  int have_thr = c->has_q && qual != NULL;
  uint8_t thr = sat_add_u8(c->min_quality, 33);
  uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
  uint64_t fwd = 0, rc = 0, run = 0;
  for (uint64_t e = 0; e < len; e++) {
    int code = orc_code(seq[e]);
    int ok = code >= 0 && (!have_thr || qual[e] >= thr);
    if (!ok) { run = 0; fwd = 0; rc = 0; continue; }
    fwd = ((fwd << 2) | (uint64_t)code) & mask;
    rc = (rc >> 2) | ((uint64_t)(3 - code) << (2 * (k - 1)));
    if (++run >= k) { orc_push_key(c, fwd < rc ? fwd : rc); c->windows++; }
  }
}

ORC_API void orc_counter_add(orc_counter *c, const uint8_t *seq, uint64_t len, const uint8_t *qual) {
  c->records++; c->bases += len;
  if (c->mode == 0) orc_add_literal(c, seq, len, qual);
  else orc_add_rolling(c, seq, len, qual);
}

/* records laid back to back; offsets has n_records+1 entries. */
ORC_API void orc_counter_add_batch(orc_counter *c, const uint8_t *seq, const uint8_t *qual,
                                   const uint64_t *offsets, uint64_t n_records) {
  for (uint64_t r = 0; r < n_records; r++)
    orc_counter_add(c, seq + offsets[r], offsets[r + 1] - offsets[r], qual ? qual + offsets[r] : NULL);
}

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

/* LSD radix sort of u64 keys, 8 passes x 8 bit (qsort is too slow for 1e8 keys). */
static void radix_sort_u64(uint64_t *a, uint64_t n) {
  if (n < 4096) { qsort(a, n, 8, cmp_u64); return; }
  uint64_t *b = (uint64_t *)malloc(n * 8);
  for (int pass = 0; pass < 8; pass++) {
    uint64_t cnt[257] = {0};
    int sh = pass * 8;
    for (uint64_t i = 0; i < n; i++) cnt[((a[i] >> sh) & 255) + 1]++;
    if (cnt[1] == n) continue; /* all zero digit: nothing to do */
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (uint64_t i = 0; i < n; i++) b[cnt[(a[i] >> sh) & 255]++] = a[i];
    memcpy(a, b, n * 8);
  }
  free(b);
}

typedef struct { uint64_t key, val; } orc_pair;
static int cmp_pair(const void *a, const void *b) {
  uint64_t x = ((const orc_pair *)a)->key, y = ((const orc_pair *)b)->key;
  return x < y ? -1 : x > y;
}

ORC_API uint64_t orc_counter_windows(const orc_counter *c) { return c->windows; }
ORC_API uint64_t orc_counter_records(const orc_counter *c) { return c->records; }
ORC_API uint64_t orc_counter_bases(const orc_counter *c) { return c->bases; }

/* Number of distinct keys (after this call rolling-mode keys are sorted). */
ORC_API uint64_t orc_counter_distinct(orc_counter *c) {
  if (c->mode == 0) return c->map.n;
  radix_sort_u64(c->keys, c->nkeys);
  uint64_t d = 0;
  for (uint64_t i = 0; i < c->nkeys; i++) if (i == 0 || c->keys[i] != c->keys[i - 1]) d++;
  return d;
}

/* Sorted (key,count) dump -- the parity object (SURVEY.md 8c "parity definition" (i)).
 * Caller provides arrays of orc_counter_distinct() entries.  Returns entries written. */
ORC_API uint64_t orc_counter_export_sorted(orc_counter *c, uint64_t *keys, uint64_t *counts) {
  uint64_t n = 0;
  if (c->mode == 0) {
    orc_pair *p = (orc_pair *)malloc((c->map.n + 1) * sizeof *p);
    for (uint64_t i = 0; i < c->map.cap; i++)
      if (c->map.used[i]) { p[n].key = c->map.keys[i]; p[n].val = c->map.vals[i]; n++; }
    qsort(p, n, sizeof *p, cmp_pair);
    for (uint64_t i = 0; i < n; i++) { keys[i] = p[i].key; counts[i] = p[i].val; }
    free(p);
    return n;
  }
  radix_sort_u64(c->keys, c->nkeys);
  for (uint64_t i = 0; i < c->nkeys;) {
    uint64_t j = i;
    while (j < c->nkeys && c->keys[j] == c->keys[i]) j++;
    keys[n] = c->keys[i]; counts[n] = j - i; n++;
    i = j;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * post-processing (L4)
 * ---------------------------------------------------------------------------------------- */

/* run.rs:447-450 / builder.rs:251-258: retain count >= min_count.  In place; returns new n. */
ORC_API uint64_t orc_filter_min_count(uint64_t *keys, uint64_t *counts, uint64_t n, uint64_t min_count) {
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; i++)
    if (counts[i] >= min_count) { keys[m] = keys[i]; counts[m] = counts[i]; m++; }
  return m;
}

/* histogram.rs:88-94 / :110-116 compute_histogram(_packed): BTreeMap<count, #distinct k-mers
 * with that count>, i.e. ascending by count.  Applied after the min-count filter
 * (run.rs:447-476).  vals/freqs need room for n entries.  Returns number of bins. */
ORC_API uint64_t orc_histogram(const uint64_t *counts, uint64_t n, uint64_t min_count,
                               uint64_t *vals, uint64_t *freqs) {
  uint64_t *tmp = (uint64_t *)malloc((n + 1) * 8), m = 0;
  for (uint64_t i = 0; i < n; i++) if (counts[i] >= min_count) tmp[m++] = counts[i];
  radix_sort_u64(tmp, m);
  uint64_t bins = 0;
  for (uint64_t i = 0; i < m;) {
    uint64_t j = i;
    while (j < m && tmp[j] == tmp[i]) j++;
    vals[bins] = tmp[i]; freqs[bins] = j - i; bins++;
    i = j;
  }
  free(tmp);
  return bins;
}

/* histogram.rs:148-169 histogram_stats: distinct = sum f, total = sum c*f, mode = max_by_key(f)
 * (Rust's max_by_key returns the LAST maximum on ties), mean = total/distinct as f64. */
ORC_API void orc_histogram_stats(const uint64_t *vals, const uint64_t *freqs, uint64_t bins,
                                 uint64_t *total, uint64_t *distinct, uint64_t *mode_count,
                                 uint64_t *mode_freq, double *mean) {
  uint64_t t = 0, d = 0, mc = 0, mf = 0;
  for (uint64_t i = 0; i < bins; i++) {
    d += freqs[i]; t += vals[i] * freqs[i];
    if (i == 0 || freqs[i] >= mf) { mf = freqs[i]; mc = vals[i]; }
  }
  *total = t; *distinct = d; *mode_count = mc; *mode_freq = mf;
  *mean = d ? (double)t / (double)d : 0.0;
}

/* index.rs:404-431 crc32: reflected IEEE polynomial 0xEDB88320, init and xor-out 0xFFFFFFFF. */
ORC_API uint32_t orc_crc32(const uint8_t *data, uint64_t len) {
  static uint32_t table[256];
  static int init = 0;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t crc = i;
      for (int j = 0; j < 8; j++) crc = (crc & 1) ? (crc >> 1) ^ 0xEDB88320u : crc >> 1;
      table[i] = crc;
    }
    init = 1;
  }
  uint32_t crc = ~0u;
  for (uint64_t i = 0; i < len; i++) crc = table[(crc ^ data[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

/* index.rs:222-279 write_index: "KMIX" 0x01 k n:u64le n*(key:u64le,count:u64le) crc32:u32le.
 * buf needs 18 + 16 n bytes.  Returns bytes written. */
ORC_API uint64_t orc_kmix_encode(uint32_t k, const uint64_t *keys, const uint64_t *counts, uint64_t n,
                                 uint8_t *buf) {
  uint64_t o = 0;
  memcpy(buf, "KMIX", 4); o = 4;
  buf[o++] = 1;
  buf[o++] = (uint8_t)k;
  for (int b = 0; b < 8; b++) buf[o++] = (uint8_t)(n >> (8 * b));
  for (uint64_t i = 0; i < n; i++) {
    for (int b = 0; b < 8; b++) buf[o++] = (uint8_t)(keys[i] >> (8 * b));
    for (int b = 0; b < 8; b++) buf[o++] = (uint8_t)(counts[i] >> (8 * b));
  }
  uint32_t crc = orc_crc32(buf, o);
  for (int b = 0; b < 4; b++) buf[o++] = (uint8_t)(crc >> (8 * b));
  return o;
}

/* index.rs:282-401 read_index checks, in the reference's order: size >= 18 (-1), magic (-2),
 * CRC (-3), version (-4), k in 1..=32 (-5), data size == n*16 (-6).  On success returns n and
 * fills keys/counts when non-NULL (cap entries). */
ORC_API int64_t orc_kmix_decode(const uint8_t *buf, uint64_t len, uint32_t *k_out, uint64_t *keys,
                                uint64_t *counts, uint64_t cap) {
  if (len < 18) return -1;
  if (memcmp(buf, "KMIX", 4) != 0) return -2;
  uint32_t stored = 0;
  for (int b = 0; b < 4; b++) stored |= (uint32_t)buf[len - 4 + b] << (8 * b);
  if (orc_crc32(buf, len - 4) != stored) return -3;
  if (buf[4] != 1) return -4;
  uint32_t k = buf[5];
  if (!orc_kmer_length_ok(k)) return -5;
  uint64_t n = 0;
  for (int b = 0; b < 8; b++) n |= (uint64_t)buf[6 + b] << (8 * b);
  if (len - 18 != n * 16) return -6;
  if (k_out) *k_out = k;
  if (keys && counts) {
    const uint8_t *p = buf + 14;
    for (uint64_t i = 0; i < n && i < cap; i++, p += 16) {
      uint64_t a = 0, c = 0;
      for (int b = 0; b < 8; b++) { a |= (uint64_t)p[b] << (8 * b); c |= (uint64_t)p[8 + b] << (8 * b); }
      keys[i] = a; counts[i] = c;
    }
  }
  return (int64_t)n;
}

/* ------------------------------------------------------------------------------------------
 * FASTA / FASTQ record parser following rust-bio 3.0.0 (Cargo.toml:17; call sites
 * reader.rs:91,96,176,181).  bio is NOT vendored under /root/reference; this restates its
 * published behaviour:
 *   FASTA: header line must start with '>'; sequence = concatenation of the following lines,
 *          each trim_end()-ed, until the next '>' line or EOF; iteration stops at the first
 *          record whose id, desc and seq are all empty.
 *   FASTQ: header '@'; sequence lines until a line starting with '+'; then the same number of
 *          quality lines; each trim_end()-ed.
 * Parser details beyond what the reference's fixtures exercise are PARITY UNPINNED.
 * Output: records appended back to back into seq_out/qual_out with offsets (n+1 entries).
 * Returns number of records, or -1 on a format error.
 * ---------------------------------------------------------------------------------------- */
static uint64_t trim_end_len(const uint8_t *p, uint64_t n) {
  /* Rust str::trim_end strips Unicode White_Space; for ASCII: space, \t, \n, \v, \f, \r. */
  while (n > 0) {
    uint8_t b = p[n - 1];
    if (b == ' ' || (b >= 9 && b <= 13)) n--; else break;
  }
  return n;
}
/* returns pointer past the line (incl. '\n'); *line_len excludes nothing (read_line keeps '\n') */
static const uint8_t *next_line(const uint8_t *p, const uint8_t *end, uint64_t *line_len) {
  const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
  const uint8_t *q = nl ? nl + 1 : end;
  *line_len = (uint64_t)(q - p);
  return q;
}

ORC_API int64_t orc_parse_fastx(const uint8_t *buf, uint64_t len, int is_fastq, uint8_t *seq_out,
                                uint8_t *qual_out, uint64_t *offsets, uint64_t max_records) {
  const uint8_t *p = buf, *end = buf + len;
  uint64_t n = 0, o = 0, ll;
  offsets[0] = 0;
  if (!is_fastq) {
    const uint8_t *line = p; p = next_line(p, end, &ll);
    while (ll > 0) {
      if (line[0] != '>') return -1;
      uint64_t hl = trim_end_len(line + 1, ll - 1);
      uint64_t start = o;
      for (;;) {
        line = p; p = next_line(p, end, &ll);
        if (ll == 0 || line[0] == '>') break;
        uint64_t t = trim_end_len(line, ll);
        memcpy(seq_out + o, line, t); o += t;
      }
      if (hl == 0 && o == start) break; /* Record::is_empty -> iterator ends */
      if (n >= max_records) return -1;
      offsets[++n] = o;
    }
    return (int64_t)n;
  }
  for (;;) {
    const uint8_t *line = p; p = next_line(p, end, &ll);
    if (ll == 0) break;
    if (line[0] != '@') return -1;
    uint64_t hl = trim_end_len(line + 1, ll - 1);
    uint64_t start = o, lines = 0, qo = o;
    for (;;) {
      line = p; p = next_line(p, end, &ll);
      if (ll == 0) return -1; /* incomplete record */
      if (line[0] == '+') break;
      uint64_t t = trim_end_len(line, ll);
      memcpy(seq_out + o, line, t); o += t; lines++;
    }
    for (uint64_t i = 0; i < lines; i++) {
      line = p; p = next_line(p, end, &ll);
      uint64_t t = trim_end_len(line, ll);
      if (qual_out) memcpy(qual_out + qo, line, t);
      qo += t;
    }
    if (qo != o) return -1; /* seq/qual length mismatch: reference would panic on slice OOB or miscount */
    if (hl == 0 && o == start) break;
    if (n >= max_records) return -1;
    offsets[++n] = o;
  }
  return (int64_t)n;
}

/* ------------------------------------------------------------------------------------------
 * "Restated reference CPU path" -- the timed CPU baseline.
 * Follows the reference's cost structure, not just its results:
 *   - record-parallel worker threads (rayon for_each over records, run.rs:500-520);
 *   - per window: heap allocation + validate/upper-case (from_sub, kmer.rs:266-286), O(k) pack
 *     (kmer.rs:304-312), bytewise canonical compare with a second allocation + repack when the
 *     reverse complement wins (kmer.rs:348-390);
 *   - upsert into a sharded, lock-protected hash map with an Fx-style multiplicative hash
 *     (DashMap<u64,u64,FxHasher>, run.rs:489, :565-571; dashmap 5.5.3 default shard count =
 *     4 x cores rounded up to a power of two).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  pthread_rwlock_t lock;
  orc_map map;
  char pad[64];
} orc_shard;

typedef struct {
  uint32_t k; int has_q; uint8_t min_quality;
  const uint8_t *seq, *qual; const uint64_t *offsets; uint64_t n_records;
  orc_shard *shards; uint64_t n_shards; int shard_shift;
  volatile uint64_t next; /* record cursor */
  uint64_t windows;
} orc_par;

static inline uint64_t fx_hash(uint64_t k) { return k * 0x517cc1b727220a95ULL; } /* rustc-hash multiplicative */

static void orc_par_upsert(orc_par *P, uint64_t key) {
  uint64_t h = fx_hash(key);
  orc_shard *s = &P->shards[(h << 7) >> P->shard_shift]; /* dashmap: (hash << 7) >> shift */
  pthread_rwlock_wrlock(&s->lock);
  orc_map_add(&s->map, key, 1);
  pthread_rwlock_unlock(&s->lock);
}

static void *orc_par_worker(void *arg) {
  orc_par *P = (orc_par *)arg;
  uint64_t k = P->k, local_windows = 0;
  uint8_t thr = sat_add_u8(P->min_quality, 33);
  for (;;) {
    uint64_t r = __atomic_fetch_add(&P->next, 1, __ATOMIC_RELAXED);
    if (r >= P->n_records) break;
    const uint8_t *seq = P->seq + P->offsets[r];
    const uint8_t *qual = P->qual ? P->qual + P->offsets[r] : NULL;
    uint64_t len = P->offsets[r + 1] - P->offsets[r];
    if (len < k) continue;
    int have_thr = P->has_q && qual;
    uint64_t i = 0;
    while (i <= len - k) {
      if (have_thr) {
        int64_t bad = -1;
        for (uint64_t j = 0; j < k; j++) if (qual[i + j] < thr) { bad = (int64_t)j; break; }
        if (bad >= 0) { i += (uint64_t)bad + 1; continue; }
      }
      uint8_t *norm = (uint8_t *)malloc(k);                 /* Vec<u8> alloc in from_sub */
      int64_t pos = orc_from_sub(seq + i, k, norm, NULL);
      if (pos < 0) {
        int is_rc;
        uint64_t fwd_bits = orc_pack_bytes(norm, k);        /* pack() */
        (void)fwd_bits;
        uint64_t key = orc_canonical(norm, k, &is_rc);      /* canonical(): compare (+ repack) */
        if (is_rc) { uint8_t *rcbuf = (uint8_t *)malloc(k); memcpy(rcbuf, norm, k); free(rcbuf); } /* 2nd alloc */
        orc_par_upsert(P, key);
        local_windows++;
        i += 1;
      } else {
        i += (uint64_t)pos + 1;
      }
      free(norm);
    }
  }
  __atomic_fetch_add(&P->windows, local_windows, __ATOMIC_RELAXED);
  return NULL;
}

/* Counts a batch with n_threads workers.  Returns counted windows; *distinct_out = #keys.
 * If keys/counts are non-NULL (cap entries) the (unsorted) table is copied out. */
ORC_API uint64_t orc_reference_path_count(uint32_t k, int has_min_quality, uint8_t min_quality,
                                          const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets,
                                          uint64_t n_records, uint32_t n_threads, uint64_t *distinct_out,
                                          uint64_t *keys, uint64_t *counts, uint64_t cap) {
  orc_par P; memset(&P, 0, sizeof P);
  P.k = k; P.has_q = has_min_quality; P.min_quality = min_quality;
  P.seq = seq; P.qual = qual; P.offsets = offsets; P.n_records = n_records;
  if (n_threads < 1) n_threads = 1;
  uint64_t ns = 4; while (ns < (uint64_t)n_threads * 4) ns <<= 1;
  P.n_shards = ns; P.shard_shift = 64 - __builtin_ctzll(ns);
  P.shards = (orc_shard *)calloc(ns, sizeof(orc_shard));
  for (uint64_t i = 0; i < ns; i++) { pthread_rwlock_init(&P.shards[i].lock, NULL); orc_map_init(&P.shards[i].map, 1024); }
  pthread_t *th = (pthread_t *)malloc(n_threads * sizeof *th);
  for (uint32_t t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, orc_par_worker, &P);
  for (uint32_t t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  uint64_t d = 0, o = 0;
  for (uint64_t i = 0; i < ns; i++) {
    d += P.shards[i].map.n;
    if (keys && counts)
      for (uint64_t j = 0; j < P.shards[i].map.cap; j++)
        if (P.shards[i].map.used[j] && o < cap) { keys[o] = P.shards[i].map.keys[j]; counts[o] = P.shards[i].map.vals[j]; o++; }
    orc_map_free(&P.shards[i].map);
    pthread_rwlock_destroy(&P.shards[i].lock);
  }
  free(P.shards); free(th);
  if (distinct_out) *distinct_out = d;
  return P.windows;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic input generator shared by tests and bench (counter-based so any slice can be
 * produced independently): one splitmix64 draw per 32 bases, 2 bits per base, MSB first.
 * The device generator in the product library (kmg_synth_*) uses the same function; tests
 * check they agree.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
ORC_API void orc_synth_uniform(uint64_t seed, uint64_t first_base, uint64_t n, uint8_t *out) {
  static const uint8_t T[4] = {'A', 'C', 'G', 'T'};
  for (uint64_t i = 0; i < n; i++) {
    uint64_t g = first_base + i;
    uint64_t w = splitmix64(seed * 0x9E3779B97F4A7C15ULL + (g >> 5));
    out[i] = T[(w >> (62 - 2 * (g & 31))) & 3];
  }
}

/* ------------------------------------------------------------------------------------------
 * Synthetic READ generator (R20M for config C3, R200M for config C5), counter-based
 * like orc_synth_uniform so that the device generator of the product library
 * (kmg_synth_reads_device) can produce the identical bytes without shipping files; a test
 * compares the two.  Specification (all arithmetic on u64, GOLD = 0x9E3779B97F4A7C15):
 *   source genome = orc_synth_uniform(seed 42), SYN_GENOME bases; every read has SYN_READ_LEN bases.
 *   read r:  a = splitmix64(seed*GOLD + 2r), b = splitmix64(seed*GOLD + 2r + 1)
 *            start  = ((a >> 32) * (SYN_GENOME - SYN_READ_LEN + 1)) >> 32,  strand = b & 1
 *     profile 3 (R20M): 1 % of the reads ((b >> 1) % 100 == 0) carry a 10-base N run at (b >> 16) % 141
 *     profile 5 (R200M): 10 % of the reads ((b >> 1) % 10 == 0) come from a satellite set instead of the
 *            genome: unit u = (b >> 8) % 64 (u = 0: "A", u = 1: "AC", else 3 + u % 29 bases drawn from
 *            splitmix64(0x5A7E111E + u), 2 bits per base from the top), phase (b >> 16) % unit length
 *   base i:  h = splitmix64((seed ^ 0xABCDEF)*GOLD + 256r + i)
 *            code = genome (forward, or reverse complement when strand = 1) or satellite unit base
 *            substitution when (h & 0xFFFF) < S (profile 3: 655 = 1 %, profile 5: 328 = 0.5 %):
 *                code = (code + 1 + ((h >> 16) & 0xFF) % 3) & 3
 *            profile 3: 'N' when ((h >> 24) & 0xFFFF) < 328 (0.5 %) or inside the read's N run
 *            profile 3 quality: u = (h >> 40) & 0xFFFF, classes Phred {2, 11, 25, 37} with cumulative
 *                thresholds {1311, 6554, 19661} (2 %, 8 %, 20 %, 70 %); in the last 30 cycles the two
 *                lowest classes are 3x likelier: {3932, 19661, 32768} (6 %, 24 %, 20 %, 50 %)
 *            profile 5: quality 'I' (no filter applies)
 * ---------------------------------------------------------------------------------------- */
#define SYN_GOLD 0x9E3779B97F4A7C15ULL
#define SYN_GENOME 100000000ULL
#define SYN_READ_LEN 150u

static inline unsigned syn_genome_code(uint64_t g) {
  uint64_t w = splitmix64(42ULL * SYN_GOLD + (g >> 5));
  return (unsigned)((w >> (62 - 2 * (g & 31))) & 3);
}
static inline unsigned syn_unit_len(unsigned u) { return u == 0 ? 1u : u == 1 ? 2u : 3u + u % 29u; }
static inline unsigned syn_unit_code(unsigned u, unsigned j) {
  if (u == 0) return 0;
  if (u == 1) return j & 1u;  /* A, C */
  uint64_t w = splitmix64(0x5A7E111EULL + u);
  return (unsigned)((w >> (62 - 2 * j)) & 3);
}

ORC_API void orc_synth_reads(uint64_t seed, uint32_t profile, uint64_t first_read, uint64_t n_reads,
                             uint8_t *seq_out, uint8_t *qual_out) {
  static const uint8_t T[4] = {'A', 'C', 'G', 'T'};
  const unsigned L = SYN_READ_LEN;
  for (uint64_t rr = 0; rr < n_reads; rr++) {
    const uint64_t r = first_read + rr;
    const uint64_t a = splitmix64(seed * SYN_GOLD + 2 * r), b = splitmix64(seed * SYN_GOLD + 2 * r + 1);
    const uint64_t start = ((a >> 32) * (SYN_GENOME - L + 1)) >> 32;
    const unsigned strand = (unsigned)(b & 1);
    const int sat = profile == 5 && ((b >> 1) % 10) == 0;
    const unsigned u = (unsigned)((b >> 8) % 64), ulen = syn_unit_len(u), phase = (unsigned)((b >> 16) % ulen);
    const int nrun = profile == 3 && ((b >> 1) % 100) == 0;
    const unsigned npos = (unsigned)((b >> 16) % (L - 9));
    const unsigned sub_thr = profile == 3 ? 655u : 328u;
    for (unsigned i = 0; i < L; i++) {
      const uint64_t h = splitmix64((seed ^ 0xABCDEFULL) * SYN_GOLD + 256 * r + i);
      unsigned code;
      if (sat) code = syn_unit_code(u, (phase + i) % ulen);
      else code = strand ? 3u - syn_genome_code(start + (L - 1 - i)) : syn_genome_code(start + i);
      if ((h & 0xFFFF) < sub_thr) code = (code + 1 + (unsigned)((h >> 16) & 0xFF) % 3u) & 3u;
      uint8_t base = T[code], q = 'I';
      if (profile == 3) {
        if (((h >> 24) & 0xFFFF) < 328u || (nrun && i >= npos && i < npos + 10)) base = 'N';
        const unsigned uq = (unsigned)((h >> 40) & 0xFFFF);
        const int late = i >= L - 30;
        const unsigned t0 = late ? 3932u : 1311u, t1 = late ? 19661u : 6554u, t2 = late ? 32768u : 19661u;
        q = (uint8_t)(33 + (uq < t0 ? 2 : uq < t1 ? 11 : uq < t2 ? 25 : 37));
      }
      seq_out[rr * L + i] = base;
      if (qual_out) qual_out[rr * L + i] = q;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Multi-threaded form of oracle #2 for inputs of BASELINE size (1e8 .. 3e9 windows): worker threads scan
 * disjoint groups of records with the rolling predicate / canonical min of orc_add_rolling (same semantics,
 * run.rs:526-571), optionally keep only the keys with key % filter_mod == filter_rem (a design-independent
 * sample of the key space, used to check the 3.1 Gbp configuration shard-wise), then the keys are bucketed by
 * their top byte and every bucket is sorted + run-length encoded by a thread.  Output arrays are malloc'ed
 * here and released with orc_free.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t k; int has_q; uint8_t min_quality;
  const uint8_t *seq, *qual; const uint64_t *offsets; uint64_t n_records, grain;
  uint64_t filter_mod, filter_rem;
  volatile uint64_t next;
  uint64_t windows;
  uint64_t **bufs; uint64_t *lens; /* per thread */
} orc_mt;
typedef struct { orc_mt *M; uint32_t tid; } orc_mt_arg;

static void *orc_mt_scan(void *argp) {
  orc_mt_arg *A = (orc_mt_arg *)argp;
  orc_mt *M = A->M;
  const uint64_t k = M->k, mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
  const uint8_t thr = sat_add_u8(M->min_quality, 33);
  uint64_t *buf = NULL, n = 0, cap = 0, windows = 0;
  for (;;) {
    const uint64_t r0 = __atomic_fetch_add(&M->next, M->grain, __ATOMIC_RELAXED);
    if (r0 >= M->n_records) break;
    const uint64_t r1 = r0 + M->grain < M->n_records ? r0 + M->grain : M->n_records;
    for (uint64_t r = r0; r < r1; r++) {
      const uint8_t *s = M->seq + M->offsets[r];
      const uint8_t *q = M->qual ? M->qual + M->offsets[r] : NULL;
      const uint64_t len = M->offsets[r + 1] - M->offsets[r];
      const int have_thr = M->has_q && q != NULL;
      uint64_t fwd = 0, rc = 0, run = 0;
      for (uint64_t e = 0; e < len; e++) {
        const int code = orc_code(s[e]);
        if (code < 0 || (have_thr && q[e] < thr)) { run = 0; fwd = 0; rc = 0; continue; }
        fwd = ((fwd << 2) | (uint64_t)code) & mask;
        rc = (rc >> 2) | ((uint64_t)(3 - code) << (2 * (k - 1)));
        if (++run >= k) {
          const uint64_t key = fwd < rc ? fwd : rc;
          windows++;
          if (M->filter_mod && key % M->filter_mod != M->filter_rem) continue;
          if (n == cap) { cap = cap ? cap * 2 : (1u << 16); buf = (uint64_t *)realloc(buf, cap * 8); }
          buf[n++] = key;
        }
      }
    }
  }
  M->bufs[A->tid] = buf; M->lens[A->tid] = n;
  __atomic_fetch_add(&M->windows, windows, __ATOMIC_RELAXED);
  return NULL;
}

typedef struct {
  uint64_t *keys; const uint64_t *bstart; volatile uint64_t next; uint64_t *out_n; /* per bucket: distinct */
  uint64_t *counts;                                                                /* RLE written in place: keys[], counts[] */
} orc_bs;
static void *orc_bucket_sort(void *argp) {
  orc_bs *B = (orc_bs *)argp;
  for (;;) {
    const uint64_t b = __atomic_fetch_add(&B->next, 1, __ATOMIC_RELAXED);
    if (b >= 256) break;
    uint64_t *a = B->keys + B->bstart[b];
    const uint64_t n = B->bstart[b + 1] - B->bstart[b];
    uint64_t *c = B->counts + B->bstart[b];
    radix_sort_u64(a, n);
    uint64_t d = 0;
    for (uint64_t i = 0; i < n;) {
      uint64_t j = i;
      while (j < n && a[j] == a[i]) j++;
      a[d] = a[i]; c[d] = j - i; d++;
      i = j;
    }
    B->out_n[b] = d;
  }
  return NULL;
}

ORC_API void orc_free(void *p) { free(p); }

/* Returns the number of distinct (kept) keys; *keys_out / *counts_out ascending by key; *windows_out counts
 * ALL counted windows (before the key filter). */
ORC_API uint64_t orc_count_batch_mt(uint32_t k, int has_min_quality, uint8_t min_quality, const uint8_t *seq,
                                    const uint8_t *qual, const uint64_t *offsets, uint64_t n_records,
                                    uint32_t n_threads, uint64_t filter_mod, uint64_t filter_rem,
                                    uint64_t **keys_out, uint64_t **counts_out, uint64_t *windows_out) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  orc_mt M; memset(&M, 0, sizeof M);
  M.k = k; M.has_q = has_min_quality; M.min_quality = min_quality;
  M.seq = seq; M.qual = qual; M.offsets = offsets; M.n_records = n_records;
  M.grain = n_records > 4096ull * n_threads ? 1024 : 1;
  M.filter_mod = filter_mod; M.filter_rem = filter_rem;
  M.bufs = (uint64_t **)calloc(n_threads, sizeof(uint64_t *));
  M.lens = (uint64_t *)calloc(n_threads, 8);
  pthread_t *th = (pthread_t *)malloc(n_threads * sizeof *th);
  orc_mt_arg *args = (orc_mt_arg *)malloc(n_threads * sizeof *args);
  for (uint32_t t = 0; t < n_threads; t++) { args[t].M = &M; args[t].tid = t; pthread_create(&th[t], NULL, orc_mt_scan, &args[t]); }
  for (uint32_t t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  uint64_t total = 0;
  for (uint32_t t = 0; t < n_threads; t++) total += M.lens[t];
  /* bucket by the top byte of the 2k-bit key (for 2k < 8: by the whole key) */
  const int sh = 2 * (int)k > 8 ? 2 * (int)k - 8 : 0;
  uint64_t bstart[257]; memset(bstart, 0, sizeof bstart);
  for (uint32_t t = 0; t < n_threads; t++)
    for (uint64_t i = 0; i < M.lens[t]; i++) bstart[((M.bufs[t][i] >> sh) & 255) + 1]++;
  for (int b = 0; b < 256; b++) bstart[b + 1] += bstart[b];
  uint64_t *keys = (uint64_t *)malloc((total + 1) * 8), *counts = (uint64_t *)malloc((total + 1) * 8);
  uint64_t cur[256];
  for (int b = 0; b < 256; b++) cur[b] = bstart[b];
  for (uint32_t t = 0; t < n_threads; t++) {
    for (uint64_t i = 0; i < M.lens[t]; i++) { const uint64_t key = M.bufs[t][i]; keys[cur[(key >> sh) & 255]++] = key; }
    free(M.bufs[t]);
  }
  uint64_t out_n[256];
  orc_bs B; B.keys = keys; B.bstart = bstart; B.next = 0; B.out_n = out_n; B.counts = counts;
  for (uint32_t t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, orc_bucket_sort, &B);
  for (uint32_t t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  uint64_t d = 0;
  for (int b = 0; b < 256; b++) {  /* close the gaps between the buckets' run-length encoded prefixes */
    if (d != bstart[b]) { memmove(keys + d, keys + bstart[b], out_n[b] * 8); memmove(counts + d, counts + bstart[b], out_n[b] * 8); }
    d += out_n[b];
  }
  free(M.bufs); free(M.lens); free(th); free(args);
  *keys_out = keys; *counts_out = counts;
  if (windows_out) *windows_out = M.windows;
  return d;
}
