// Internal C++ interface between the C-ABI layer (kmg_api.cu) and the kernels (kmg_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmg {

// packed-stream view handed to the scan kernels; pointers include the zero lead-in words.
struct ScanInput {
  const uint64_t *bases;
  const uint32_t *valid;
  const uint32_t *start;  // nullptr: single record / no boundaries inside the stream
  uint64_t n_tiles;
  int k;
};

struct HashTable {
  uint64_t *slots;  // cap x {key, count}
  uint64_t cap;
};

// read-side view over either table kind
struct TableView {
  const uint64_t *slots;             // v1 hash table (key, occurrences-1) slots
  const unsigned long long *dense;   // direct path: 4^k counters
  const uint64_t *pair_keys;         // partitioned path: consolidated run, SoA
  const uint64_t *pair_counts;
  uint64_t n;                        // slots, 4^k or number of pairs
};

// ---- partitioned pipeline (phase B, kmg_consolidate.cu) --------------------------------------------
constexpr int CONS_THREADS = 256;
constexpr int CONS_CTAS_PER_SM = 4;
constexpr int CONS_INSERT_CHUNK = 8192;   // keys per insert ticket (32 per thread, 4 rounds of 8)
constexpr int CONS_COMPACT_CHUNK = 8192;  // table slots per compact ticket
constexpr int CONS_NBUF = 3;              // L2-resident table buffers in flight
constexpr int CONS_MAX_RUNS = 16;
constexpr int MAX_PARTS = 8192;

struct ConsRun {
  const uint64_t *keys;
  const uint64_t *counts;   // nullptr: every key counts 1
  const uint64_t *offsets;  // device, n_parts + 1 entries
};
struct ConsPhase {
  uint32_t first_ticket;
  uint32_t part_and_type;  // bit 31: 1 = compact, 0 = insert
};
struct ConsParams {
  uint32_t n_parts, R, total_tickets, preagg;
  ConsRun runs[CONS_MAX_RUNS];
  const ConsPhase *phases;         // schedule order, terminated by a sentinel with first_ticket = total_tickets
  const uint32_t *part_cap_log2;   // table capacity (log2 slots) of each partition
  const uint32_t *part_nI, *part_nC;  // tickets per phase
  const uint32_t *part_wait;       // previous NON-EMPTY partition that used the same table buffer (~0: none)
  uint64_t *tables;                // CONS_NBUF buffers of table_stride_slots (key, count-1) slots
  uint64_t table_stride_slots;
  uint64_t *out_keys, *out_counts;
  unsigned long long *out_base;    // n_parts + 1; [0] = 0, rest ~0 until published
  uint32_t *done_I, *done_C, *distinct, *out_cursor;  // n_parts each, zeroed
  uint32_t *ticket, *error_flag;
};

enum { CTR_WINDOWS = 0, CTR_DISTINCT = 1, CTR_FULL = 2, CTR_SCRATCH = 3, CTR_N = 8 };

constexpr int SCAN_CTAS_PER_SM = 4;
constexpr int DENSE_SMEM_MAX_K = 6;      // 4^6 u32 = 16 KiB privatised per CTA
constexpr int DENSE_MAX_K = 14;          // 4^14 u64 = 2 GiB
constexpr int DENSE_DEFAULT_MAX_K = 13;  // automatic choice of the direct path
constexpr int HIST_SMEM_BINS = 2048;
constexpr int HIST_DENSE_BINS = 65536;

cudaError_t launch_ingest(const uint8_t *d_seq, const uint8_t *d_qual, uint64_t n_bytes, uint32_t thr, uint64_t n_words_total,
                          uint64_t *d_bases, uint32_t *d_valid, cudaStream_t s);
cudaError_t launch_start_bits(const uint64_t *d_offsets, uint64_t n_records, uint64_t base_offset, uint64_t n_bytes,
                              uint64_t n_words_total, uint32_t *d_start, cudaStream_t s);
cudaError_t launch_synth_uniform(uint64_t seed, uint64_t first_base, uint64_t n, uint8_t *d_out, cudaStream_t s);
cudaError_t launch_scan_hash(const ScanInput &in, HashTable t, unsigned long long *counters, uint32_t flags, cudaStream_t s);
cudaError_t launch_scan_dense(const ScanInput &in, unsigned long long *dense, unsigned long long *counters, uint32_t flags,
                              cudaStream_t s);
// scatter == false: part_counts[p] += keys of partition p.  scatter == true: keys are written to
// out[part_start[p] + ...]; part_cursor[] (zeroed) hands out ranges inside each partition.
cudaError_t launch_scan_partition(const ScanInput &in, uint32_t n_parts, bool scatter, unsigned long long *part_counts,
                                  const unsigned long long *part_start, unsigned long long *part_cursor, uint64_t *out,
                                  unsigned long long *counters, cudaStream_t s);
cudaError_t launch_partition_keys(const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n, uint32_t n_parts, bool scatter,
                                  unsigned long long *part_counts, const unsigned long long *part_start,
                                  unsigned long long *part_cursor, uint64_t *out_keys, uint64_t *out_counts, int num_sms,
                                  cudaStream_t s);
cudaError_t launch_consolidate(const ConsParams &P, int num_sms, cudaStream_t s);
int num_sms();
cudaError_t launch_table_init(HashTable t, cudaStream_t s);
cudaError_t launch_insert_keys(HashTable t, const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n,
                               unsigned long long *counters, cudaStream_t s);
cudaError_t launch_insert_keys_dense(unsigned long long *dense, const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n,
                                     unsigned long long *counters, cudaStream_t s);
cudaError_t launch_rehash(HashTable from, HashTable to, unsigned long long *counters, cudaStream_t s);
cudaError_t launch_table_stats(const TableView &v, uint64_t min_count, unsigned long long *d_stats3, cudaStream_t s);
cudaError_t launch_compact(const TableView &v, uint64_t min_count, uint64_t *d_keys, uint64_t *d_counts, uint64_t cap_out,
                           unsigned long long *d_cursor, cudaStream_t s);
cudaError_t launch_histogram(const TableView &v, uint64_t min_count, unsigned long long *d_bins, uint64_t *d_overflow,
                             uint64_t overflow_cap, unsigned long long *d_overflow_n, cudaStream_t s);
uint64_t kernel_launches();  // number of kernels of this library launched so far (process-wide)
cudaError_t sort_pairs(uint64_t *d_keys, uint64_t *d_counts, uint64_t n, int key_bits, cudaStream_t s);

}  // namespace kmg
