// C-ABI layer of libkmerust_gpu (include/kmerust_gpu.h): context, device memory pool, the staging ring of the
// host feed, the choice between the three counting paths (direct 4^k array, one HBM table, partitioned pipeline) and
// its adaptation to a growing input (table growth, migration to the partitioned path, speculative layouts with
// exact fallback, re-split / multi-pass when the partition plan is outgrown), consolidation of the runs, result export.
// The counting itself happens in the sm_100a kernels of kmg_kernels.cu / kmg_partition.cu.  There is no CPU
// fallback: every entry point needs a CUDA device.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/kmerust_gpu.h"
#include "kmg_device.cuh"
#include "kmg_kernels.h"

using namespace kmg;

#define KMG_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_create_error;

constexpr double LOAD_TARGET = 0.60;  // sizing goal when memory allows
constexpr double LOAD_SOFT = 0.70;    // grow after a batch if exceeded and memory allows
constexpr double LOAD_HARD = 0.90;    // never start a batch that could exceed this
constexpr uint64_t DEFAULT_BATCH_BASES = 1ull << 28;  // per staging slot; every batch becomes one run of the partitioned pipeline
constexpr int N_STAGE = 4;                             // ASCII staging ring: copies run up to 3 chunks ahead of the kernels
constexpr uint64_t MIN_TABLE_SLOTS = 1ull << 16;

struct Staging {
  uint8_t *h_seq = nullptr, *h_qual = nullptr;  // pinned (lazily allocated; skipped for pinned callers)
  uint64_t *h_off = nullptr;
  uint8_t *d_seq = nullptr, *d_qual = nullptr;
  uint64_t *d_off = nullptr;
  uint64_t off_cap = 0;
  cudaEvent_t h2d_done = nullptr, compute_done = nullptr;
  bool h2d_pending = false, compute_pending = false;
  // pre-packed feed (kmg_acquire_batch)
  uint64_t *hp_bases = nullptr;
  uint32_t *hp_valid = nullptr, *hp_start = nullptr;
  uint64_t *dp_bases = nullptr;                      // the slot's packed stream on the device (with the zero lead-in words)
  uint32_t *dp_valid = nullptr, *dp_start = nullptr;
  uint64_t hp_used_words = 0;                        // words the last producer may have written
};

}  // namespace

// A run is a partition-indexed list of keys (phase A output, every key counts 1) or of (key, count)
// pairs (phase B output / weighted inserts).  All runs of a context share the same partition count;
// partition p of a run is entries [seg_start[p], seg_start[p] + seg_len[p]).
struct Run {
  uint64_t *d_keys = nullptr, *d_counts = nullptr, *d_seg_start = nullptr, *d_seg_len = nullptr;
  uint64_t n = 0;        // entries (a consolidated run may hold skipped filler entries, see count_partitions_smem_kernel)
  uint64_t n_valid = 0;  // consolidated runs: distinct keys
  bool padded = false;   // speculative layout: segments start on 16-byte boundaries and an EMPTY_MIX entry follows every segment of odd length
};

struct kmg_ctx {
  kmg_config cfg{};
  int device = 0;
  int k = 0;
  bool use_dense = false;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  bool own_stream = false;
  std::string err;

  // table
  HashTable table{nullptr, 0};
  unsigned long long *dense = nullptr;
  uint64_t dense_n = 0;
  uint64_t distinct_ub = 0;  // upper bound of distinct keys currently in the table
  uint32_t n_grows = 0;

  unsigned long long *d_counters = nullptr;  // CTR_N
  unsigned long long *h_counters = nullptr;  // pinned mirror
  unsigned long long *d_stats = nullptr;     // 3 + cursor + overflow_n

  // packed stream buffers (grown on demand)
  uint64_t *d_bases = nullptr;
  uint32_t *d_valid = nullptr, *d_start = nullptr;
  uint64_t packed_words = 0;  // capacity in words (excluding lead-in)

  uint64_t batch_bases = DEFAULT_BATCH_BASES;
  Staging st[N_STAGE];
  bool staging_ready = false, packed_feed_ready = false;
  uint64_t staging_cap = 0;
  uint32_t next_slot = 0;    // pre-packed feed: next ring slot to hand out
  int pending_slot = -1;     // pre-packed feed: batch whose copy is queued but which has not been scanned yet
  uint64_t pending_words = 0, dp_words = 0;
  uint32_t next_ascii = 0;   // ASCII staging ring

  // partitioned pipeline (v2)
  enum Mode { MODE_UNDECIDED, MODE_DENSE, MODE_TABLE, MODE_PARTITIONED } mode = MODE_UNDECIDED;
  uint32_t n_parts = 0, n_coarse = 0, n_sub = 0, scratch_log2 = 13;
  std::vector<Run> runs;      // pending, not yet consolidated
  Run result;                 // consolidated (key, count) run
  bool has_result = false;
  bool spec_coarse_ok = !getenv("KMG_NO_SPECULATION");  // same for the level-1 scatter
  bool spec_fine_ok = !getenv("KMG_NO_SPECULATION");  // level-2 scatter without a count pass until a partition overflows its share
  size_t total_mem = 0;                      // device memory size (cudaMemGetInfo is slow; asked once)
  void *d_scan_tmp = nullptr;                // CUB scan scratch for scan_tmp_items items
  size_t scan_tmp_bytes = 0;
  uint64_t scan_tmp_items = 0, fine_cursor_items = 0;
  bool in_resplit = false;
  struct ShardShm *shm = nullptr;     // sharded mode (kmg_shard_join): the group's shared segment
  uint32_t sh_world = 1, sh_rank = 0;
  uint64_t *sh_recv = nullptr;        // this rank's SEND buffer: two halves of sh_recv_entries keys (owners pull their regions over NVLink)
  uint64_t sh_recv_entries = 0;
  uint64_t *sh_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // every rank's send buffer as seen from here
  bool sh_peer_ipc[8] = {false, false, false, false, false, false, false, false};
  uint64_t sh_sent = 0, sh_recv_keys = 0, sh_rounds = 0, sh_exact_rounds = 0;
  std::vector<uint64_t> sh_hist_vals, sh_hist_freqs;  // merged histogram of the last kmg_shard_histogram
  int feed_kind = 0;        // partitioned path: 1 = runs binned by this context's own coarse function, 2 = blocks adopted from a sender
                            // that binned by (owner, bin) -- the two bin functions differ, so one context takes only one kind
  bool poisoned = false;    // a re-split failed half way: runs with different partitionings coexist; only kmg_reset / kmg_destroy are safe
  double dedup_ratio = 1.0;  // distinct keys per raw entry seen by the last consolidation (sizes the next consolidated run)
  double plan_scale = 1.0;  // kmg_count_ascii with a quality filter: (bases of the whole call) / (bases of its first chunk)
  int building_run = 0;  // > 0 while refine_to_run derives a run from the current plan: a nested consolidate must not re-split
  // count-of-counts of `result`, produced by phase B itself (consolidate)
  unsigned long long *d_hist = nullptr;      // HIST_DENSE_BINS bins + overflow counter
  uint64_t *d_hist_ov = nullptr;             // HIST_OVERFLOW_CAP counts >= HIST_DENSE_BINS
  bool fused_valid = false, fused_cached = false;
  std::vector<unsigned long long> h_bins;    // host copy, bins[1] filled in
  std::vector<uint64_t> h_ov;                // ascending
  uint64_t fused_sum = 0, fused_max = 0;
  uint64_t pending_bytes = 0;
  unsigned long long *d_part = nullptr;  // 3 * MAX_PARTS scratch: coarse counts, starts, cursors
  unsigned long long *d_fine_cursor = nullptr;  // n_parts
  uint32_t n_consolidations = 0;
  uint64_t sieve_redo = 0;  // partitions the sieve variant of phase B handed to the compacting variant (diagnostic)

  // device-memory pool for the partitioned pipeline's large, short-lived buffers: cudaMalloc/cudaFree of tens
  // of GB cost ~10 ms each and a counting job needs a dozen of them, so freed blocks are kept for reuse
  std::vector<std::pair<void *, size_t>> pool_idle;
  std::unordered_map<void *, size_t> pool_live;

  uint64_t n_records = 0, n_bases = 0, h2d_bytes = 0;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  double kernel_ms = 0.0;
  double cat_ms[3] = {0.0, 0.0, 0.0};  // 0: scan / partition kernels (includes 2), 1: consolidation kernel, 2: refine (A2) launches, nested inside 0
  struct Timer { cudaEvent_t a, b; int cat; };
  std::vector<Timer> pending_timers;
};

namespace {

void shard_release(kmg_ctx *c);  // defined with the shard group code below
kmg_status scan_pending_batch(kmg_ctx *c);  // pre-packed feed: scan the batch whose copy is in flight (defined with kmg_submit_batch)

kmg_status fail(kmg_ctx *ctx, kmg_status s, const std::string &msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return s;
}
kmg_status cuda_fail(kmg_ctx *ctx, cudaError_t e, const char *what) {
  kmg_status s = (e == cudaErrorMemoryAllocation) ? KMG_ERR_OOM : KMG_ERR_CUDA;
  return fail(ctx, s, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call)                                            \
  do {                                                           \
    cudaError_t _e = (call);                                     \
    if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call);     \
  } while (0)

inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

void pool_release_idle(kmg_ctx *c) {
  for (auto &b : c->pool_idle) cudaFree(b.first);
  c->pool_idle.clear();
}
cudaError_t pool_alloc(kmg_ctx *c, void **p, size_t bytes) {
  bytes = std::max<size_t>(bytes, 256);
  size_t best = SIZE_MAX, best_i = 0;
  for (size_t i = 0; i < c->pool_idle.size(); ++i) {
    const size_t sz = c->pool_idle[i].second;
    if (sz >= bytes && sz <= bytes + bytes / 8 + (1u << 20) && sz < best) { best = sz; best_i = i; }
  }
  if (best != SIZE_MAX) {
    *p = c->pool_idle[best_i].first;
    c->pool_live[*p] = best;
    c->pool_idle.erase(c->pool_idle.begin() + best_i);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {  // give idle blocks back to the driver and retry once
    cudaGetLastError();
    pool_release_idle(c);
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) c->pool_live[*p] = bytes;
  return e;
}
template <class T>
cudaError_t pool_alloc(kmg_ctx *c, T **p, size_t bytes) { return pool_alloc(c, reinterpret_cast<void **>(p), bytes); }
void pool_free(kmg_ctx *c, void *p) {
  if (!p) return;
  auto it = c->pool_live.find(p);
  if (it == c->pool_live.end()) { cudaFree(p); return; }
  c->pool_idle.emplace_back(p, it->second);
  c->pool_live.erase(it);
}

TableView view_of(const kmg_ctx *c) {
  TableView v{nullptr, nullptr, nullptr, nullptr, 0};
  if (c->mode == kmg_ctx::MODE_PARTITIONED) { v.pair_keys = c->result.d_keys; v.pair_counts = c->result.d_counts; v.n = c->has_result ? c->result.n : 0; }
  else if (c->use_dense) { v.dense = c->dense; v.n = c->dense_n; }
  else { v.slots = c->table.slots; v.n = c->table.slots ? c->table.cap : 0; }
  return v;
}

void free_run(kmg_ctx *c, Run &r) {
  pool_free(c, r.d_keys); pool_free(c, r.d_counts); pool_free(c, r.d_seg_start); pool_free(c, r.d_seg_len);
  r = Run();
}

kmg_status alloc_table(kmg_ctx *c, uint64_t cap, HashTable *out) {
  cap = std::max<uint64_t>(cap, MIN_TABLE_SLOTS);
  cap = round_up(cap, 2);
  uint64_t *p = nullptr;
  cudaError_t e = cudaMalloc(&p, cap * 16);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(c, KMG_ERR_OOM, "cudaMalloc(table) failed: " + std::string(cudaGetErrorString(e))); }
  out->slots = p; out->cap = cap;
  CU(c, launch_table_init(*out, c->stream));
  return KMG_OK;
}

// Try target loads from generous to tight until an allocation succeeds.
kmg_status alloc_table_for(kmg_ctx *c, uint64_t need_keys, HashTable *out) {
  const double loads[3] = {LOAD_TARGET, 0.75, LOAD_HARD};
  kmg_status s = KMG_ERR_OOM;
  for (double l : loads) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    uint64_t cap = (uint64_t)((double)need_keys / l) + 2;
    if (cap * 16 > (uint64_t)free_b) continue;
    s = alloc_table(c, cap, out);
    if (s == KMG_OK) return s;
  }
  return fail(c, KMG_ERR_TABLE_FULL, "cannot allocate a table for " + std::to_string(need_keys) + " keys");
}

kmg_status read_counters(kmg_ctx *c) {
  CU(c, cudaMemcpyAsync(c->h_counters, c->d_counters, CTR_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->h_counters[CTR_FULL]) return fail(c, KMG_ERR_TABLE_FULL, "hash table filled up during a batch (capacity planning bug)");
  return KMG_OK;
}

kmg_status grow_table(kmg_ctx *c, uint64_t need_keys) {
  HashTable nt{nullptr, 0};
  kmg_status s = alloc_table_for(c, need_keys, &nt);
  if (s != KMG_OK) return s;
  if (nt.cap <= c->table.cap) { cudaFree(nt.slots); return fail(c, KMG_ERR_TABLE_FULL, "table cannot grow any further"); }
  CU(c, launch_rehash(c->table, nt, c->d_counters, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->table.slots);
  c->table = nt;
  c->n_grows++;
  return KMG_OK;
}

// Make sure `incoming` more upserts cannot push the load above LOAD_HARD.  Returns how many of them
// may be issued now (>= 1 tile's worth or an error).
kmg_status reserve_capacity(kmg_ctx *c, uint64_t incoming, uint64_t *granted) {
  *granted = incoming;
  if (c->use_dense) return KMG_OK;
  auto room = [&]() -> uint64_t {
    uint64_t lim = (uint64_t)(LOAD_HARD * (double)c->table.cap);
    return lim > c->distinct_ub ? lim - c->distinct_ub : 0;
  };
  if (incoming <= room()) { c->distinct_ub += incoming; return KMG_OK; }
  kmg_status s = read_counters(c);  // exact distinct count
  if (s != KMG_OK) return s;
  c->distinct_ub = c->h_counters[CTR_DISTINCT];
  if (incoming <= room() && (double)c->distinct_ub <= LOAD_SOFT * (double)c->table.cap) { c->distinct_ub += incoming; return KMG_OK; }
  // grow: enough for what is there plus the incoming batch
  s = grow_table(c, c->distinct_ub + incoming);
  if (s != KMG_OK) {
    // could not grow: hand out whatever room is left (caller splits the batch)
    uint64_t r = room();
    if (r < (uint64_t)TILE_WORDS * 32) return s;
    c->err.clear();
    *granted = std::min(incoming, r);
  }
  c->distinct_ub += *granted;
  return KMG_OK;
}

kmg_status ensure_packed(kmg_ctx *c, uint64_t n_words_total) {
  if (n_words_total <= c->packed_words) return KMG_OK;
  CU(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_bases); cudaFree(c->d_valid); cudaFree(c->d_start);
  c->d_bases = nullptr; c->d_valid = c->d_start = nullptr; c->packed_words = 0;
  CU(c, cudaMalloc(&c->d_bases, (LEAD_BASE_WORDS + n_words_total) * 8));
  CU(c, cudaMalloc(&c->d_valid, (LEAD_MASK_WORDS + n_words_total) * 4));
  CU(c, cudaMalloc(&c->d_start, (LEAD_MASK_WORDS + n_words_total) * 4));
  CU(c, cudaMemsetAsync(c->d_bases, 0, LEAD_BASE_WORDS * 8, c->stream));
  CU(c, cudaMemsetAsync(c->d_valid, 0, LEAD_MASK_WORDS * 4, c->stream));
  CU(c, cudaMemsetAsync(c->d_start, 0, LEAD_MASK_WORDS * 4, c->stream));
  c->packed_words = n_words_total;
  return KMG_OK;
}

size_t timer_begin(kmg_ctx *c, int cat = 0) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, c->stream);
  c->pending_timers.push_back(kmg_ctx::Timer{a, b, cat});
  return c->pending_timers.size() - 1;
}
void timer_end(kmg_ctx *c, size_t idx) { cudaEventRecord(c->pending_timers[idx].b, c->stream); }
void timers_collect(kmg_ctx *c) {
  for (auto &p : c->pending_timers) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { if (p.cat != 2) c->kernel_ms += ms; c->cat_ms[p.cat] += ms; }
    cudaEventDestroy(p.a); cudaEventDestroy(p.b);
  }
  c->pending_timers.clear();
}


// =================================================================================================
// Partitioned pipeline: A1 coarse scatter, A2 refine to fine partitions, B per-CTA counting
// (kernels and rationale: kmg_partition.cu).
// =================================================================================================
constexpr uint64_t TARGET_KEYS_PER_PART = 3600;  // <= 4096 with 8 sigma to spare: one upsert round per partition; load ~0.44

// n_coarse / n_sub are set: allocate the partitioned pipeline's bookkeeping
kmg_status init_partitioned(kmg_ctx *c) {
  c->n_parts = c->n_coarse * c->n_sub;
  CU(c, cudaMalloc(&c->d_part, 3 * (size_t)MAX_PARTS * sizeof(unsigned long long)));
  CU(c, cudaMalloc(&c->d_fine_cursor, (size_t)c->n_parts * sizeof(unsigned long long)));
  c->fine_cursor_items = c->n_parts;
  return KMG_OK;
}

// Decide how this context counts.  Called at the first feeding call, when the input size is known.
kmg_status decide_mode(kmg_ctx *c, uint64_t first_call_windows) {
  if (c->mode != kmg_ctx::MODE_UNDECIDED) return KMG_OK;
  if (c->use_dense) { c->mode = kmg_ctx::MODE_DENSE; return KMG_OK; }
  const uint32_t f = c->cfg.flags;
  const uint64_t hint = c->cfg.expected_distinct ? c->cfg.expected_distinct : first_call_windows;  // an explicit hint is taken at its word
  const bool part = (f & KMG_FLAG_FORCE_PARTITIONED) || (!(f & KMG_FLAG_FORCE_HASH) && hint >= (1ull << 25));
  if (!part) {
    c->mode = kmg_ctx::MODE_TABLE;
    if (!c->table.slots) {
      kmg_status s = c->cfg.expected_distinct ? alloc_table_for(c, c->cfg.expected_distinct, &c->table) : alloc_table(c, 1ull << 22, &c->table);
      if (s != KMG_OK) return s;
    }
    return KMG_OK;
  }
  c->mode = kmg_ctx::MODE_PARTITIONED;
  if (c->cfg.parts_log2) {
    const uint32_t lg = std::min<uint32_t>(c->cfg.parts_log2, 20);
    c->n_coarse = 1u << (lg / 2);
    c->n_sub = 1u << (lg - lg / 2);
  } else {
    uint64_t target = TARGET_KEYS_PER_PART;
    if (const char *t = getenv("KMG_TARGET_KEYS")) target = std::max<uint64_t>(256, strtoull(t, nullptr, 10));  // tuning experiments only
    const uint64_t want = std::max<uint64_t>(4, (hint + target - 1) / target);
    uint64_t p1 = 1;
    while (p1 * p1 < want) ++p1;
    p1 = std::min<uint64_t>(p1, 2048);
    c->n_coarse = (uint32_t)p1;
    c->n_sub = (uint32_t)std::min<uint64_t>((want + p1 - 1) / p1, 2048);
  }
  return init_partitioned(c);
}

kmg_status consolidate(kmg_ctx *c, bool recompact = false);

kmg_status add_run(kmg_ctx *c, Run &&r) {
  if (r.n == 0) { free_run(c, r); return KMG_OK; }
  c->pending_bytes += r.n * (r.d_counts ? 16 : 8);
  c->runs.push_back(std::move(r));
  if (!c->total_mem) { size_t free_b = 0; cudaMemGetInfo(&free_b, &c->total_mem); }
  const size_t total_b = c->total_mem;
  // consolidate early when the pending runs get numerous or large (keeps streaming inputs bounded in memory)
  if (c->runs.size() + (c->has_result ? 1 : 0) >= (size_t)CONS_MAX_RUNS - 1 || c->pending_bytes > (uint64_t)total_b * 20 / 100)
    return consolidate(c);
  return KMG_OK;
}

// allocate with one retry after consolidating what is pending (frees the pending runs' buffers)
kmg_status alloc_or_consolidate(kmg_ctx *c, void **p, size_t bytes, const char *what) {
  cudaError_t e = pool_alloc(c, p, bytes ? bytes : 1);
  if (e == cudaSuccess) return KMG_OK;
  cudaGetLastError();
  if (c->in_resplit) return fail(c, KMG_ERR_OOM, std::string("cudaMalloc(") + what + ") failed while re-splitting the runs");
  kmg_status s = consolidate(c);
  if (s != KMG_OK) return s;
  e = pool_alloc(c, p, bytes ? bytes : 1);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(c, KMG_ERR_OOM, std::string("cudaMalloc(") + what + ") failed"); }
  return KMG_OK;
}

// A2: coarse-partitioned keys (+counts) -> fine-partitioned run.  Takes ownership of the coarse buffers when `owns`.
// `sync`: wait for the scatter before returning (needed when the input belongs to the caller).  Without it the coarse
// buffers go back to the pool while the kernels are still queued, which is safe because every pool block is only ever
// touched by work on c->stream (stream-ordered reuse).
// coarse_len == nullptr: coarse partition p is [coarse_off[p], coarse_off[p + 1]); otherwise [coarse_off[p], coarse_off[p] + len[p]).
// `split` != nullptr: the input is a fine-partitioned run (split->n_in partitions) that is re-split split->m ways; the new run is
// returned in *split->out instead of being added to the pending runs.
struct SplitPlan { uint32_t n_in, m, sub_old; Run *out; };
kmg_status refine_to_run(kmg_ctx *c, uint64_t *d_ckeys, uint64_t *d_ccounts, const std::vector<uint64_t> &coarse_off, bool owns = true,
                         bool sync = true, const std::vector<uint64_t> *coarse_len = nullptr, const SplitPlan *split = nullptr,
                         bool in_keys = false,    // in_keys: the input holds plain keys, not mixes (blocks adopted from another rank)
                         uint32_t in_group = 1,   // in_group g: coarse bin b arrives as the g input partitions b*g .. b*g+g-1 (sharded: one per source rank)
                         const uint64_t *const *src = nullptr) {  // src[s]: input partition p is read from src[p % in_group] (other ranks' send buffers)
  struct BuildGuard {
    kmg_ctx *c;
    explicit BuildGuard(kmg_ctx *c_) : c(c_) { ++c->building_run; }
    void release() { if (c) { --c->building_run; c = nullptr; } }
    ~BuildGuard() { release(); }
  } guard(c);
  const uint32_t P1 = split ? split->n_in : c->n_coarse * in_group;  // input partitions
  const uint32_t n_sub = split ? split->m : c->n_sub;
  const uint32_t P = split ? split->n_in * split->m : c->n_parts;
  auto len_of = [&](uint32_t p) { return coarse_len ? (*coarse_len)[p] : coarse_off[p + 1] - coarse_off[p]; };
  uint64_t n = 0;
  for (uint32_t p = 0; p < P1; ++p) n += len_of(p);
  Run r;
  uint64_t *d_cstart = nullptr, *d_clen = nullptr;
  uint32_t *d_tprefix = nullptr;
  auto cleanup = [&]() { if (owns) { pool_free(c, d_ckeys); pool_free(c, d_ccounts); } pool_free(c, d_cstart); pool_free(c, d_clen); pool_free(c, d_tprefix); };
  std::vector<uint32_t> tprefix(P1 + 1, 0);
  uint64_t tiles = 0;
  for (uint32_t p = 0; p < P1; ++p) { tprefix[p] = (uint32_t)tiles; tiles += (len_of(p) + REFINE_TILE - 1) / REFINE_TILE; }
  tprefix[P1] = (uint32_t)tiles;
  cudaError_t e = cudaSuccess;
  if (c->fine_cursor_items < P) {  // the partition count grew (re-split)
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_fine_cursor); c->d_fine_cursor = nullptr; c->fine_cursor_items = 0;
    CU(c, cudaMalloc(&c->d_fine_cursor, (size_t)P * sizeof(unsigned long long)));
    c->fine_cursor_items = P;
  }
  e = pool_alloc(c, &d_cstart, (P1 + 1) * 8);
  if (e == cudaSuccess && coarse_len) {
    e = pool_alloc(c, &d_clen, (size_t)P1 * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_clen, coarse_len->data(), (size_t)P1 * 8, cudaMemcpyHostToDevice, c->stream);
  }
  if (e == cudaSuccess) e = pool_alloc(c, &d_tprefix, (P1 + 1) * 4);
  if (e == cudaSuccess) e = pool_alloc(c, &r.d_seg_start, (size_t)P * 8);
  if (e == cudaSuccess) e = pool_alloc(c, &r.d_seg_len, (size_t)P * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_cstart, coarse_off.data(), (P1 + 1) * 8, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_tprefix, tprefix.data(), (P1 + 1) * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(r.d_seg_len, 0, (size_t)P * 8, c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(c->d_fine_cursor, 0, (size_t)P * 8, c->stream);
  if (e != cudaSuccess) { cleanup(); free_run(c, r); return cuda_fail(c, e, "refine setup"); }
  RefineParams rp{};
  rp.keys = d_ckeys; rp.counts = d_ccounts; rp.coarse_start = d_cstart; rp.coarse_len = d_clen; rp.tile_prefix = d_tprefix;
  rp.n_coarse = P1; rp.n_sub = n_sub; rp.n_tiles = (uint32_t)tiles;
  if (split) { rp.sub_total = split->sub_old * split->m; rp.sub_old = split->sub_old; }
  rp.in_keys = in_keys ? 1u : 0u;
  rp.in_group = in_group;
  if (src) { rp.n_src = in_group; for (uint32_t i = 0; i < in_group && i < 8; ++i) rp.src[i] = src[i]; }
  // Speculative layout first: hash partitions are Poisson-sized, so every fine partition gets mean + 7 sigma + 16 slots
  // and the count pass (a full read of the keys) is skipped.  Skewed input overflows a share: the kernel then raises a
  // flag, and this chunk -- and, sticky, the rest of the job -- takes the exact count + prefix + scatter route below.
  const double mu = (double)n / P;
  const uint64_t cap_f = ((uint64_t)(mu + 7.0 * std::sqrt(mu) + 16.0) + 7) & ~7ull;
  if (c->spec_fine_ok && mu >= 64.0 && cap_f * P < (1ull << 32) && refine_single_pass_available(n_sub, d_ccounts != nullptr)) {
    kmg_status s = alloc_or_consolidate(c, reinterpret_cast<void **>(&r.d_keys), cap_f * P * 8, "fine keys");
    if (s != KMG_OK) { cleanup(); free_run(c, r); return s; }
    rp.fine_cap = cap_f;
    rp.fine_cursor = reinterpret_cast<unsigned long long *>(r.d_seg_len);  // zeroed above; ends up as the partition sizes
    rp.overflow_flag = reinterpret_cast<uint32_t *>(c->d_stats + 5);
    rp.out_keys = r.d_keys;
    uint32_t h_flag = 0;
    e = cudaMemsetAsync(rp.overflow_flag, 0, 4, c->stream);
    if (e == cudaSuccess) e = launch_fill_strided(r.d_seg_start, P, cap_f, c->stream);
    const size_t tmr2 = timer_begin(c, 2);
    if (e == cudaSuccess) e = launch_refine(rp, true, c->stream);
    timer_end(c, tmr2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, rp.overflow_flag, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cleanup(); free_run(c, r); return cuda_fail(c, e, "refine scatter (speculative layout)"); }
    if (!h_flag) {
      cleanup();
      r.n = n;
      if (!(cap_f & 1ull) && launch_pad_segments(r.d_keys, r.d_seg_start, r.d_seg_len, P, c->stream) == cudaSuccess) r.padded = true;
      if (split) { *split->out = std::move(r); return KMG_OK; }
      guard.release();  // the run is complete: a consolidation triggered by add_run may re-split it with the others
      return add_run(c, std::move(r));
    }
    c->spec_fine_ok = false;
    pool_free(c, r.d_keys); r.d_keys = nullptr;
    rp.fine_cap = 0; rp.overflow_flag = nullptr; rp.out_keys = nullptr;
    e = cudaMemsetAsync(r.d_seg_len, 0, (size_t)P * 8, c->stream);
    if (e != cudaSuccess) { cleanup(); free_run(c, r); return cuda_fail(c, e, "refine setup"); }
  }
  rp.fine_counts = reinterpret_cast<unsigned long long *>(r.d_seg_len);
  rp.fine_start = reinterpret_cast<const unsigned long long *>(r.d_seg_start);
  rp.fine_cursor = c->d_fine_cursor;
  const size_t tmr3 = timer_begin(c, 2);
  e = launch_refine(rp, false, c->stream);
  timer_end(c, tmr3);
  if (e == cudaSuccess && (!c->d_scan_tmp || c->scan_tmp_items < P)) {
    cudaFree(c->d_scan_tmp); c->d_scan_tmp = nullptr;
    e = exclusive_sum_u64(nullptr, nullptr, P, nullptr, &c->scan_tmp_bytes, c->stream);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_scan_tmp, c->scan_tmp_bytes ? c->scan_tmp_bytes : 16);
    c->scan_tmp_items = P;
  }

  if (e == cudaSuccess) e = exclusive_sum_u64(r.d_seg_len, r.d_seg_start, P, c->d_scan_tmp, &c->scan_tmp_bytes, c->stream);
  if (e != cudaSuccess) { cleanup(); free_run(c, r); return cuda_fail(c, e, "refine count"); }
  kmg_status s = alloc_or_consolidate(c, reinterpret_cast<void **>(&r.d_keys), n * 8, "fine keys");
  if (s == KMG_OK && d_ccounts) s = alloc_or_consolidate(c, reinterpret_cast<void **>(&r.d_counts), n * 8, "fine counts");
  if (s != KMG_OK) { cleanup(); free_run(c, r); return s; }
  rp.out_keys = r.d_keys; rp.out_counts = r.d_counts;
  const size_t tmr4 = timer_begin(c, 2);
  e = launch_refine(rp, true, c->stream);
  timer_end(c, tmr4);
  if (e == cudaSuccess && sync) e = cudaStreamSynchronize(c->stream);
  cleanup();
  if (e != cudaSuccess) { free_run(c, r); return cuda_fail(c, e, "refine scatter"); }
  r.n = n;
  if (split) { *split->out = std::move(r); return KMG_OK; }
  guard.release();
  return add_run(c, std::move(r));
}

// A1 over the packed stream: coarse count pass, exact offsets, coarse scatter; then A2.
kmg_status scan_to_run(kmg_ctx *c, uint64_t n_words_total, bool has_start) {
  if (c->poisoned) return fail(c, KMG_ERR_STATE, "context is inconsistent after a failed re-split: kmg_reset it");
  if (c->feed_kind == 2) return fail(c, KMG_ERR_STATE, "context holds blocks adopted from a sender (kmg_adopt_coarse_device): it cannot also scan input itself");
  c->feed_kind = 1;
  const uint64_t n_tiles = n_words_total / TILE_WORDS;
  const uint64_t max_tiles = ((1ull << 32) - 1) / ((uint64_t)TILE_WORDS * 32);  // < 2^32 windows per launch
  const uint32_t P1 = c->n_coarse;
  unsigned long long *d_cnt = c->d_part, *d_start = c->d_part + MAX_PARTS, *d_cur = c->d_part + 2 * MAX_PARTS;
  for (uint64_t tile0 = 0; tile0 < n_tiles; tile0 += max_tiles) {
    ScanInput in;
    in.bases = c->d_bases + tile0 * TILE_WORDS;
    in.valid = c->d_valid + tile0 * TILE_WORDS;
    in.start = has_start ? c->d_start + tile0 * TILE_WORDS : nullptr;
    in.n_tiles = std::min(max_tiles, n_tiles - tile0);
    in.k = c->k;
    const size_t tmr = timer_begin(c, 0);
    CU(c, cudaMemsetAsync(c->d_part, 0, 3 * (size_t)MAX_PARTS * sizeof(unsigned long long), c->stream));
    // Speculative layout first (no count pass): hash bins are Poisson-sized, so every coarse bin gets an equal share of
    // (windows of this launch) / P1 + 7 sigma + 1024 slots.  A bin that outgrows its share (heavily repeated k-mers) raises
    // a flag; this launch -- and, sticky, the rest of the job -- then takes the exact count + prefix + scatter route.
    {
      const uint64_t n_ub = in.n_tiles * (uint64_t)TILE_WORDS * 32;
      const double mu = (double)n_ub / P1;
      const uint64_t cap_c = ((uint64_t)(mu + 7.0 * std::sqrt(mu) + 1024.0) + 15) & ~15ull;
      if (c->spec_coarse_ok && cap_c * P1 < (1ull << 32) && scan_scatter_supports_cap(P1)) {
        uint64_t *d_ckeys = nullptr;
        kmg_status s = alloc_or_consolidate(c, reinterpret_cast<void **>(&d_ckeys), cap_c * P1 * 8, "coarse keys");
        if (s != KMG_OK) return s;
        std::vector<uint64_t> off(P1 + 1, 0), lens(P1, 0);
        for (uint32_t p = 0; p <= P1; ++p) off[p] = (uint64_t)p * cap_c;
        ScanInput sin = in;
        sin.part_cap = cap_c;
        sin.overflow_flag = reinterpret_cast<uint32_t *>(c->d_stats + 5);
        uint32_t h_flag = 0;
        cudaError_t e = cudaMemsetAsync(sin.overflow_flag, 0, 4, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_start, off.data(), P1 * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = launch_scan_partition(sin, P1, true, d_cnt, d_start, d_cur, d_ckeys, c->d_counters, c->stream, /*mixed=*/true);
        if (e == cudaSuccess) e = cudaMemcpyAsync(lens.data(), d_cur, P1 * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, sin.overflow_flag, 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { pool_free(c, d_ckeys); return cuda_fail(c, e, "coarse scatter (speculative layout)"); }
        if (!h_flag) {
          uint64_t n = 0;
          for (uint32_t p = 0; p < P1; ++p) n += lens[p];
          if (n == 0) { pool_free(c, d_ckeys); timer_end(c, tmr); continue; }
          s = refine_to_run(c, d_ckeys, nullptr, off, /*owns=*/true, /*sync=*/false, &lens);
          timer_end(c, tmr);
          if (s != KMG_OK) return s;
          continue;
        }
        c->spec_coarse_ok = false;
        pool_free(c, d_ckeys);
        CU(c, cudaMemsetAsync(c->d_part, 0, 3 * (size_t)MAX_PARTS * sizeof(unsigned long long), c->stream));
      }
    }
    CU(c, launch_scan_partition(in, P1, false, d_cnt, d_start, d_cur, nullptr, c->d_counters, c->stream));
    std::vector<unsigned long long> counts(P1);
    CU(c, cudaMemcpyAsync(counts.data(), d_cnt, P1 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> off(P1 + 1, 0);
    for (uint32_t p = 0; p < P1; ++p) off[p + 1] = off[p] + counts[p];
    const uint64_t n = off[P1];
    if (n == 0) { timer_end(c, tmr); continue; }
    uint64_t *d_ckeys = nullptr;
    kmg_status s = alloc_or_consolidate(c, reinterpret_cast<void **>(&d_ckeys), n * 8, "coarse keys");
    if (s != KMG_OK) return s;
    CU(c, cudaMemcpyAsync(d_start, off.data(), P1 * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, launch_scan_partition(in, P1, true, d_cnt, d_start, d_cur, d_ckeys, c->d_counters, c->stream, /*mixed=*/true));
    s = refine_to_run(c, d_ckeys, nullptr, off, /*owns=*/true, /*sync=*/false);  // input is the context's own packed stream
    timer_end(c, tmr);
    if (s != KMG_OK) return s;
  }
  return KMG_OK;
}

// weighted keys (device) -> one run (the receive side of the multi-GPU exchange)
kmg_status keys_to_run(kmg_ctx *c, const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n) {
  if (c->poisoned) return fail(c, KMG_ERR_STATE, "context is inconsistent after a failed re-split: kmg_reset it");
  if (c->feed_kind == 2) return fail(c, KMG_ERR_STATE, "context holds blocks adopted from a sender (kmg_adopt_coarse_device): it cannot also take plain keys");
  c->feed_kind = 1;
  const uint32_t P1 = c->n_coarse;
  unsigned long long *d_cnt = c->d_part, *d_start = c->d_part + MAX_PARTS, *d_cur = c->d_part + 2 * MAX_PARTS;
  const uint64_t max_n = (1ull << 32) - 1;
  for (uint64_t i0 = 0; i0 < n; i0 += max_n) {
    const uint64_t m = std::min(max_n, n - i0);
    const size_t tmr = timer_begin(c, 0);
    CU(c, cudaMemsetAsync(c->d_part, 0, 3 * (size_t)MAX_PARTS * sizeof(unsigned long long), c->stream));
    CU(c, launch_keys_coarse(d_keys + i0, d_counts ? d_counts + i0 : nullptr, m, P1, false, d_cnt, d_start, d_cur, nullptr, nullptr, c->stream));
    std::vector<unsigned long long> counts(P1);
    CU(c, cudaMemcpyAsync(counts.data(), d_cnt, P1 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> off(P1 + 1, 0);
    for (uint32_t p = 0; p < P1; ++p) off[p + 1] = off[p] + counts[p];
    uint64_t *d_ckeys = nullptr, *d_ccounts = nullptr;
    kmg_status s = alloc_or_consolidate(c, reinterpret_cast<void **>(&d_ckeys), m * 8, "coarse keys");
    if (s == KMG_OK && d_counts) s = alloc_or_consolidate(c, reinterpret_cast<void **>(&d_ccounts), m * 8, "coarse counts");
    if (s != KMG_OK) { pool_free(c, d_ckeys); return s; }
    CU(c, cudaMemcpyAsync(d_start, off.data(), P1 * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, launch_keys_coarse(d_keys + i0, d_counts ? d_counts + i0 : nullptr, m, P1, true, d_cnt, d_start, d_cur, d_ckeys, d_ccounts, c->stream));
    s = refine_to_run(c, d_ckeys, d_ccounts, off);  // synchronises: the caller may reuse d_keys when we return
    timer_end(c, tmr);
    if (s != KMG_OK) return s;
  }
  return KMG_OK;
}

// The input outgrew the partition plan: split every fine partition of every run m ways (one streaming pass per run; the
// sub-bin function nests, so partition p becomes partitions p*m .. p*m + m-1).  Afterwards n_sub and n_parts are m times larger.
kmg_status resplit_runs(kmg_ctx *c, uint32_t m) {
  const uint32_t P_old = c->n_parts;
  std::vector<Run *> all;
  if (c->has_result && c->result.n) all.push_back(&c->result);
  for (auto &r : c->runs) all.push_back(&r);
  for (auto *r : all) if (r->n >= (1ull << 32) - 1) return KMG_OK;  // refine launches index with 32 bits: leave it to the multi-pass kernel
  c->in_resplit = true;
  kmg_status st = KMG_OK;
  std::vector<uint64_t> off(P_old + 1), lens(P_old);
  for (auto *r : all) {
    cudaError_t e = cudaMemcpyAsync(off.data(), r->d_seg_start, (size_t)P_old * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lens.data(), r->d_seg_len, (size_t)P_old * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { st = cuda_fail(c, e, "re-split (segment table)"); break; }
    off[P_old] = P_old ? off[P_old - 1] + lens[P_old - 1] : 0;
    Run nr;
    SplitPlan sp{P_old, m, c->n_sub, &nr};
    st = refine_to_run(c, r->d_keys, r->d_counts, off, /*owns=*/false, /*sync=*/true, &lens, &sp);
    if (st != KMG_OK) break;
    nr.n_valid = r->n_valid;
    free_run(c, *r);
    *r = std::move(nr);
  }
  c->in_resplit = false;
  if (st != KMG_OK) {  // runs already re-split and runs not yet re-split coexist now: refuse everything but a reset
    c->poisoned = true;
    return fail(c, st, c->err + " (while re-splitting the runs: the context must be reset)");
  }
  c->n_sub *= m;
  c->n_parts = P_old * m;
  return KMG_OK;
}

// phase B: merge the consolidated result (if any) and all pending runs into a new consolidated run
// recompact: rewrite the consolidated result alone (its entries are mostly skipped fillers after a direct-emit launch on
// duplicate-rich input) with the compacting kernel
kmg_status consolidate(kmg_ctx *c, bool recompact) {
  if (c->runs.empty() && !(recompact && c->has_result && c->result.n)) return KMG_OK;
  {  // partitions several times larger than planned: refine the plan before counting
    uint64_t entries = c->has_result ? c->result.n_valid : 0;
    for (auto &r : c->runs) entries += r.n;
    const uint64_t avg = entries / std::max<uint32_t>(c->n_parts, 1);
    // (not while refine_to_run is building a run from the current plan -- its allocations may land here -- : the run would be
    // finished with the old partition count and sub-bin function; partitions are counted in passes instead)
    if (avg > 2 * TARGET_KEYS_PER_PART && c->n_sub * 2 <= 2048 && !c->cfg.parts_log2 && !c->building_run) {  // an explicit partition count is respected
      const uint32_t m = (uint32_t)std::min<uint64_t>((avg + TARGET_KEYS_PER_PART - 1) / TARGET_KEYS_PER_PART, 2048 / c->n_sub);
      // Re-splitting costs one tile (~3.5 us of one SM) per input partition or per 8192 entries of every run plus a streaming
      // pass; counting in m passes instead costs ~7 ps per entry and extra pass, again at every later consolidation.  Many small
      // runs over a slightly outgrown plan are cheaper in passes, one big run over a badly outgrown plan is cheaper re-split.
      double resplit_ns = (double)entries * 0.006, passes_ns = (double)entries * (m - 1) * 0.007 * 2.0;
      auto tiles_of = [&](const Run &r) { return (double)std::max<uint64_t>(c->n_parts, r.n / REFINE_TILE); };
      if (c->has_result && c->result.n) resplit_ns += tiles_of(c->result) * 3500.0 / num_sms();
      for (auto &r : c->runs) resplit_ns += tiles_of(r) * 3500.0 / num_sms();
      if (m >= 2 && resplit_ns < passes_ns) { kmg_status rs = resplit_runs(c, m); if (rs != KMG_OK) return rs; }
    }
  }
  std::vector<Run *> in;
  if (c->has_result && c->result.n) in.push_back(&c->result);
  for (auto &r : c->runs) in.push_back(&r);
  const uint32_t P = c->n_parts, R = (uint32_t)in.size();
  if (R > (uint32_t)CONS_MAX_RUNS) return fail(c, KMG_ERR_STATE, "too many pending runs");
  uint64_t total = 0;
  for (auto *r : in) total += r->n;

  CountParams prm{};
  prm.n_parts = P; prm.R = R; prm.preagg = !(c->cfg.flags & KMG_FLAG_NO_PREAGG);
  for (uint32_t r = 0; r < R; ++r) prm.runs[r] = ConsRun{in[r]->d_keys, in[r]->d_counts, in[r]->d_seg_start, in[r]->d_seg_len};

  // processing order: partitions with many entries first (a partition is counted by ONE CTA).  Only when some partition is
  // heavy (> 8x the average) is the per-partition table fetched and sorted; otherwise partitions are taken as they come.
  unsigned long long *d_totals = nullptr;
  uint32_t *d_order = nullptr;
  unsigned long long h_max = 0;
  cudaError_t e = pool_alloc(c, &d_totals, (size_t)P * 8);
  if (e == cudaSuccess) e = launch_sum_lens(prm, d_totals, c->d_stats + 6, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_max, c->d_stats + 6, 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { pool_free(c, d_totals); return cuda_fail(c, e, "consolidate setup"); }
  const uint64_t max_total = h_max;
  const uint64_t heavy = 8 * std::max<uint64_t>(total / P, 1024);
  if (max_total > heavy) {
    std::vector<unsigned long long> totals(P);
    e = pool_alloc(c, &d_order, (size_t)P * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(totals.data(), d_totals, (size_t)P * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { pool_free(c, d_totals); pool_free(c, d_order); return cuda_fail(c, e, "consolidate setup"); }
    std::vector<uint32_t> order, big;
    order.reserve(P);
    for (uint32_t p = 0; p < P; ++p) if (totals[p] > heavy) big.push_back(p);
    std::sort(big.begin(), big.end(), [&](uint32_t a, uint32_t b) { return totals[a] > totals[b]; });
    order = big;
    for (uint32_t p = 0; p < P; ++p) if (totals[p] <= heavy) order.push_back(p);
    e = cudaMemcpy(d_order, order.data(), (size_t)P * 4, cudaMemcpyHostToDevice);  // synchronous: `order` dies with this scope
  }
  pool_free(c, d_totals);
  prm.order = d_order;

  // The output run holds one entry per DISTINCT key.  `total` (one per input entry) is always enough, but read sets repeat every
  // k-mer many times (C5: 26 G windows, 2.4 G distinct) and a run sized for the input would not fit next to the inputs.  So the
  // run is sized from the dedup ratio the previous consolidation observed (1.0 at first), and a launch that runs out of space
  // (nospace flag, nothing written past the capacity) is repeated with whatever memory allows.
  Run out;
  uint64_t raw_entries = 0;
  for (auto *r : in) if (!r->d_counts) raw_entries += r->n;
  // Direct emit (count_partitions_smem_kernel<false, true>): only unweighted runs of keys that are mostly distinct (a genome, not a
  // deep read set), partitions within the plan.  Its output holds one entry per input entry.
  static const bool no_direct = getenv("KMG_NO_DIRECT") != nullptr;  // ablation
  bool direct = !no_direct && !recompact && raw_entries == total && c->dedup_ratio >= 0.5 && max_total <= 0xffffu &&
                total / std::max<uint32_t>(P, 1) <= SMEM_TABLE_SLOTS / 2 && !(c->cfg.flags & KMG_FLAG_NO_PREAGG);
  // Sieve variant (count_partitions_sieve_kernel): same precondition; it leaves the partitions it cannot take (too many repeated keys,
  // more than one batch of entries) to the compacting variant, so no bound on the largest partition is needed.
  static const bool no_sieve = getenv("KMG_NO_SIEVE") != nullptr;  // ablation
  bool sieve = !no_sieve && !recompact && raw_entries == total && c->dedup_ratio >= 0.5 &&
               total / std::max<uint32_t>(P, 1) <= SIEVE_MAX_ENTRIES * 15 / 16 && !(c->cfg.flags & KMG_FLAG_NO_PREAGG);
  uint32_t *d_redo = nullptr;
  unsigned long long *d_redo_base = nullptr;
  if (sieve && (pool_alloc(c, &d_redo, (size_t)P * 8) != cudaSuccess || pool_alloc(c, &d_redo_base, (size_t)P * 8) != cudaSuccess)) {  // d_redo: partition numbers, then range lengths
    cudaGetLastError(); pool_free(c, d_redo); pool_free(c, d_redo_base); d_redo = nullptr; d_redo_base = nullptr; sieve = false;
  }
  bool sieve_padded = sieve && getenv("KMG_NO_SIEVE_TMA") == nullptr;
  for (auto *r : in) sieve_padded = sieve_padded && r->padded;
  uint64_t out_cap = std::max<uint64_t>(total, 1) + (sieve_padded ? (uint64_t)P * R : 0);  // padded segments: up to one filler entry per partition and run
  if (recompact) out_cap = std::min<uint64_t>(out_cap, c->result.n_valid + (1ull << 20));
  else if (total > (1ull << 28) && !direct && !sieve) {
    const uint64_t est = (total - raw_entries) + (uint64_t)((double)raw_entries * std::min(1.0, c->dedup_ratio * 1.25 + 0.02)) + (1ull << 24);
    out_cap = std::min(out_cap, est);
  }
  auto alloc_out = [&](uint64_t cap) -> cudaError_t {
    cudaError_t ae = pool_alloc(c, &out.d_keys, cap * 8);
    if (ae == cudaSuccess) ae = pool_alloc(c, &out.d_counts, cap * 8);
    if (ae != cudaSuccess) { pool_free(c, out.d_keys); pool_free(c, out.d_counts); out.d_keys = out.d_counts = nullptr; cudaGetLastError(); }
    return ae;
  };
  if (e == cudaSuccess) {
    e = alloc_out(out_cap);
    if (e != cudaSuccess && out_cap > (1ull << 26)) {  // not even the estimate fits: take what is there
      size_t free_b = 0, total_b = 0;
      pool_release_idle(c);
      cudaMemGetInfo(&free_b, &total_b);
      const uint64_t room = free_b > (3ull << 30) ? (free_b - (2ull << 30)) / 16 : 0;
      if (room >= (1ull << 26)) { out_cap = std::min(out_cap, room); e = alloc_out(out_cap); }
    }
  }
  if (e == cudaSuccess) e = pool_alloc(c, &out.d_seg_start, (size_t)P * 8);
  if (e == cudaSuccess) e = pool_alloc(c, &out.d_seg_len, (size_t)P * 8);
  if (e != cudaSuccess) { pool_free(c, d_order); pool_free(c, d_redo); pool_free(c, d_redo_base); free_run(c, out); return cuda_fail(c, e, "cudaMalloc(consolidated run)"); }
  prm.out_keys = out.d_keys; prm.out_counts = out.d_counts;
  prm.out_seg_start = out.d_seg_start; prm.out_seg_len = out.d_seg_len;
  prm.out_cap = out_cap;

  if (!c->d_hist) {
    e = cudaMalloc(&c->d_hist, (HIST_DENSE_BINS + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_hist_ov, HIST_OVERFLOW_CAP * 8);
    if (e != cudaSuccess) { pool_free(c, d_order); pool_free(c, d_redo); pool_free(c, d_redo_base); free_run(c, out); return cuda_fail(c, e, "cudaMalloc(histogram)"); }
  }
  prm.hist = c->d_hist; prm.hist_overflow = c->d_hist_ov; prm.hist_overflow_cap = HIST_OVERFLOW_CAP;
  c->fused_valid = c->fused_cached = false;

  const unsigned grid = (unsigned)std::min<uint64_t>(P, (uint64_t)num_sms() * COUNT_CTAS_PER_SM);
  kmg_status st = KMG_OK;
  unsigned long long n_out = 0;
  // attempt 0: shared-memory tables (primary).  If a partition holds more distinct keys than such a table
  // (inputs much larger than the partition plan), fall back to L2-resident scratch tables of growing size.
  // The shared-memory kernel counts partitions that outgrew the plan in several passes (split_log2, from the average
  // partition size); if a partition still overflows its table the split is refined twice before the L2-scratch variant
  // takes over.
  uint32_t split_log2 = 0;
  while (split_log2 < 8 && (3600ull << split_log2) * 5 / 4 < total / std::max<uint32_t>(P, 1)) ++split_log2;
  unsigned long long n_valid = 0;
  int smem_attempts = 0;
  for (int attempt = 0;; ++attempt) {
    const bool use_smem = smem_attempts < 3 && split_log2 <= 8;
    uint64_t *d_scratch = nullptr;
    unsigned long long *d_sync = nullptr;  // [0] out_cursor, [1] next (u32) | error (u32), [2] distinct keys, [3] no-space flag (u32)
    const uint64_t slots = use_smem ? 0 : (uint64_t)grid << c->scratch_log2;
    prm.split_log2 = split_log2;
    e = pool_alloc(c, &d_sync, 32);
    if (e == cudaSuccess && slots) e = pool_alloc(c, &d_scratch, slots * 16);
    if (e != cudaSuccess) { pool_free(c, d_scratch); pool_free(c, d_sync); st = cuda_fail(c, e, "cudaMalloc(count scratch)"); break; }
    prm.scratch = d_scratch; prm.scratch_log2 = c->scratch_log2;
    prm.out_cursor = d_sync;
    prm.out_distinct = d_sync + 2;
    prm.next = reinterpret_cast<uint32_t *>(d_sync + 1);
    prm.error_flag = reinterpret_cast<uint32_t *>(d_sync + 1) + 1;
    prm.nospace_flag = reinterpret_cast<uint32_t *>(d_sync + 3);
    e = cudaMemsetAsync(d_sync, 0, 32, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_hist, 0, (HIST_DENSE_BINS + 1) * sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess && slots) e = launch_table_init(HashTable{d_scratch, slots}, c->stream, EMPTY_MIX);
    const size_t tmr = timer_begin(c, 1);
    bool weighted = max_total > ((2ull * SMEM_COUNT_THREADS * 8) << split_log2);  // partitions oversized beyond the split want the pre-aggregating variant
    for (uint32_t r = 0; r < R; ++r) weighted |= in[r]->d_counts != nullptr;
    if (weighted || split_log2 || !use_smem) direct = sieve = false;
    unsigned long long h_sync[4] = {0, 0, 0, 0};
    if (sieve) {
      prm.redo_list = d_redo; prm.redo_count = reinterpret_cast<uint32_t *>(d_sync + 3) + 1;
      prm.redo_len = d_redo + P; prm.redo_base = d_redo_base;
      if (e == cudaSuccess) e = launch_count_partitions_sieve(prm, sieve_padded, c->stream);
      timer_end(c, tmr);
      if (e == cudaSuccess) e = cudaMemcpyAsync(h_sync, d_sync, 32, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      const uint32_t n_redo = (uint32_t)(h_sync[3] >> 32);
      if (e == cudaSuccess && (n_redo > sieve_redo_limit_host(P) || (uint32_t)h_sync[3])) {
        // duplicate-rich input (the reserved ranges of the partitions handed back are lost as fillers, the padded variant stops early):
        // the launch is repeated without the sieve
        pool_free(c, d_scratch); pool_free(c, d_sync);
        sieve = sieve_padded = direct = false;
        c->dedup_ratio = std::min(c->dedup_ratio, 0.49);
        continue;
      }
      if (e == cudaSuccess && n_redo) {
        // the partitions the sieve left alone: compacting variant, continuing the same output cursor, distinct count and histogram
        CountParams redo = prm;
        redo.order = d_redo; redo.n_parts = n_redo; redo.redo_list = nullptr; redo.redo_count = nullptr;
        redo.pre_base = d_redo_base; redo.pre_len = d_redo + P; redo.redo_base = nullptr; redo.redo_len = nullptr;
        e = cudaMemsetAsync(prm.next, 0, 4, c->stream);
        const size_t tmr2 = timer_begin(c, 1);
        if (e == cudaSuccess) e = launch_count_partitions_smem(redo, false, false, c->stream);
        timer_end(c, tmr2);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_sync, d_sync, 32, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        c->sieve_redo += n_redo;
      }
      static const bool dbg_sieve = getenv("KMG_DEBUG_SIEVE") != nullptr;
      if (dbg_sieve) fprintf(stderr, "[kmg] sieve: %u partitions, %u left to the compacting variant, %llu entries, %llu distinct\n", P, n_redo, (unsigned long long)h_sync[0], (unsigned long long)h_sync[2]);
    } else {
      if (e == cudaSuccess) e = use_smem ? launch_count_partitions_smem(prm, weighted, direct, c->stream) : launch_count_partitions(prm, grid, c->stream);
      timer_end(c, tmr);
      if (e == cudaSuccess) e = cudaMemcpyAsync(h_sync, d_sync, 32, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    pool_free(c, d_scratch); pool_free(c, d_sync);
    if (e != cudaSuccess) { st = cuda_fail(c, e, "count_partitions"); break; }
    if ((uint32_t)h_sync[3] && !(h_sync[1] >> 32)) {  // the run was too small for the distinct keys: enlarge it and repeat
      const uint64_t need = std::min<uint64_t>(total, h_sync[0] + (1ull << 20));  // out_cursor kept counting: the exact demand (plus filler entries)
      if (out_cap >= total || attempt >= 16) { st = fail(c, KMG_ERR_OOM, "the consolidated run does not fit in device memory"); break; }
      pool_free(c, out.d_keys); pool_free(c, out.d_counts); out.d_keys = out.d_counts = nullptr;
      out_cap = need;
      if (alloc_out(out_cap) != cudaSuccess) {
        pool_release_idle(c);
        if (alloc_out(out_cap) != cudaSuccess) { st = fail(c, KMG_ERR_OOM, "the consolidated run (" + std::to_string(need) + " entries) does not fit in device memory"); break; }
      }
      prm.out_keys = out.d_keys; prm.out_counts = out.d_counts; prm.out_cap = out_cap;
      c->dedup_ratio = 1.0;
      continue;
    }
    n_out = h_sync[0]; n_valid = h_sync[2];
    if (!(h_sync[1] >> 32)) break;  // no overflow
    if (sieve) { sieve = false; continue; }    // a left-over partition was beyond the compacting variant too: the whole launch is repeated without the sieve
    if (direct) { direct = false; continue; }  // a partition beyond the direct variant's limits: the compacting variant takes the launch
    if (attempt >= 16 || c->scratch_log2 >= 30) { st = fail(c, KMG_ERR_TABLE_FULL, "partition table overflow could not be resolved"); break; }
    if (use_smem) {  // finer split first (weights beyond 32 bits are not cured by it: the L2 variant follows after three tries)
      ++smem_attempts;
      split_log2 += 2;
      c->scratch_log2 = std::max<uint32_t>(c->scratch_log2, 14);
    } else ++c->scratch_log2;  // retry with larger tables
  }
  pool_free(c, d_order); pool_free(c, d_redo); pool_free(c, d_redo_base);
  if (st != KMG_OK) { free_run(c, out); return st; }
  out.n = n_out; out.n_valid = n_valid;
  if (raw_entries > (1ull << 24)) {  // distinct keys the raw runs added per raw entry (an upper estimate: keys already in the result count as new)
    const uint64_t before = total - raw_entries;
    c->dedup_ratio = std::min(1.0, (double)(n_valid > before / 2 ? n_valid - before / 2 : 0) / (double)raw_entries);
  }
  if (c->has_result) free_run(c, c->result);
  for (auto &r : c->runs) free_run(c, r);
  c->runs.clear();
  c->pending_bytes = 0;
  // give back the slack (upper bound was one slot per input entry) when it is worth a copy
  if (out.n && out.n < total / 2) {
    uint64_t *k2 = nullptr, *c2 = nullptr;
    if (pool_alloc(c, &k2, out.n * 8) == cudaSuccess && pool_alloc(c, &c2, out.n * 8) == cudaSuccess) {
      cudaMemcpyAsync(k2, out.d_keys, out.n * 8, cudaMemcpyDeviceToDevice, c->stream);
      cudaMemcpyAsync(c2, out.d_counts, out.n * 8, cudaMemcpyDeviceToDevice, c->stream);
      cudaStreamSynchronize(c->stream);
      pool_free(c, out.d_keys); pool_free(c, out.d_counts);
      out.d_keys = k2; out.d_counts = c2;
    } else { cudaGetLastError(); pool_free(c, k2); pool_free(c, c2); }
  }
  c->result = std::move(out);
  c->has_result = true;
  c->fused_valid = true;
  c->n_consolidations++;
  if ((direct || sieve) && c->result.n > (1ull << 20) && c->result.n_valid < c->result.n / 2) {
    // duplicate-rich input went through the direct variant (nothing was known about it yet): most entries are fillers.
    // Rewrite the run compactly once; dedup_ratio now steers later consolidations to the compacting variant.
    kmg_status rs = consolidate(c, /*recompact=*/true);
    if (rs != KMG_OK) return rs;
    c->n_consolidations--;
  }
  return KMG_OK;
}

// Host copy of the count-of-counts phase B left behind for `result`.  false: not available (no consolidated
// result, or more than HIST_OVERFLOW_CAP counts >= HIST_DENSE_BINS) -- the caller then runs the table kernels.
bool fetch_fused_hist(kmg_ctx *c) {
  if (c->mode != kmg_ctx::MODE_PARTITIONED || !c->has_result || !c->runs.empty() || !c->fused_valid) return false;
  if (c->fused_cached) return true;
  c->h_bins.assign(HIST_DENSE_BINS + 1, 0);
  if (cudaMemcpyAsync(c->h_bins.data(), c->d_hist, (HIST_DENSE_BINS + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess) { cudaGetLastError(); c->fused_valid = false; return false; }
  const uint64_t ov_n = c->h_bins[HIST_DENSE_BINS];
  if (ov_n > HIST_OVERFLOW_CAP) { c->fused_valid = false; return false; }
  c->h_ov.resize(ov_n);
  if (ov_n && cudaMemcpy(c->h_ov.data(), c->d_hist_ov, ov_n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); c->fused_valid = false; return false; }
  std::sort(c->h_ov.begin(), c->h_ov.end());
  c->h_bins.resize(HIST_DENSE_BINS);
  uint64_t others = ov_n, sum = 0, mx = 0;
  for (uint64_t b = 2; b < (uint64_t)HIST_DENSE_BINS; ++b) if (c->h_bins[b]) { others += c->h_bins[b]; sum += b * c->h_bins[b]; mx = b; }
  c->h_bins[0] = 0;
  c->h_bins[1] = c->result.n_valid - others;  // counts of 1 are not recorded by the kernel
  sum += c->h_bins[1];
  if (c->h_bins[1] && mx < 1) mx = 1;
  for (uint64_t v : c->h_ov) sum += v;
  if (ov_n) mx = c->h_ov.back();
  c->fused_sum = sum; c->fused_max = mx;
  c->fused_cached = true;
  return true;
}

// Run the counting scan over a packed stream of n_words_total words (already in d_bases/d_valid/d_start).
// A context that started small sits on the single HBM table (random atomics: fine while it fits L2, ~6x slower than the
// partitioned pipeline beyond).  Once the table plus the incoming batch pass MIGRATE_KEYS it is emptied into a weighted run
// and the context continues on the partitioned path, planned for 16x what it has seen (over-partitioning is cheap at this size;
// re-splitting / multi-pass counting covers streams that grow further).
constexpr uint64_t MIGRATE_KEYS = 1ull << 26;
kmg_status migrate_to_partitioned(kmg_ctx *c, uint64_t incoming) {
  kmg_status s = read_counters(c);
  if (s != KMG_OK) return s;
  const uint64_t n = c->h_counters[CTR_DISTINCT];
  uint64_t *dk = nullptr, *dc = nullptr;
  if (n) {
    cudaError_t e = pool_alloc(c, &dk, n * 8);
    if (e == cudaSuccess) e = pool_alloc(c, &dc, n * 8);
    if (e == cudaSuccess) e = launch_compact(view_of(c), 1, dk, dc, n, c->d_stats + 3, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { pool_free(c, dk); pool_free(c, dc); return cuda_fail(c, e, "migrate to the partitioned path"); }
  }
  cudaFree(c->table.slots);
  c->table = HashTable{nullptr, 0};
  c->mode = kmg_ctx::MODE_UNDECIDED;
  c->cfg.flags |= KMG_FLAG_FORCE_PARTITIONED;
  c->cfg.expected_distinct = std::max<uint64_t>(c->cfg.expected_distinct, 16 * (n + incoming));  // the hint was too small (or absent)
  s = decide_mode(c, c->cfg.expected_distinct);
  if (s == KMG_OK && n) s = keys_to_run(c, dk, dc, n);
  pool_free(c, dk); pool_free(c, dc);
  c->n_grows++;
  return s;
}

kmg_status scan_packed(kmg_ctx *c, uint64_t n_words_total, bool has_start) {
  const uint64_t n_tiles = n_words_total / TILE_WORDS;
  uint64_t incoming = n_words_total * 32;  // upper bound of the keys this stream adds
  bool incoming_exact = false;
  if (!c->use_dense && c->cfg.has_min_quality) {
    // a quality filter can remove almost every window (config C3 keeps ~3 %): plan the table / the partitions from the
    // countable windows (masks only, 0.25 B/base), not from the bases
    unsigned long long h_ok = 0;
    CU(c, launch_count_windows(c->d_valid, has_start ? c->d_start : nullptr, n_words_total, c->k, c->d_stats + 7, c->stream));
    CU(c, cudaMemcpyAsync(&h_ok, c->d_stats + 7, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    incoming = h_ok; incoming_exact = true;
  }
  uint64_t plan_windows = incoming;
  if (incoming_exact && c->plan_scale > 1.0) plan_windows = (uint64_t)((double)incoming * c->plan_scale * 1.25) + 1024;  // first chunk of a longer call
  c->plan_scale = 1.0;
  kmg_status ms = decide_mode(c, plan_windows);
  if (ms != KMG_OK) return ms;
  if (c->mode == kmg_ctx::MODE_TABLE && !(c->cfg.flags & KMG_FLAG_FORCE_HASH) && c->distinct_ub + incoming > MIGRATE_KEYS) {
    ms = migrate_to_partitioned(c, incoming);
    if (ms != KMG_OK) return ms;
  }
  if (c->mode == kmg_ctx::MODE_PARTITIONED) {
    static const bool no_sparse = getenv("KMG_NO_SPARSE") != nullptr;  // ablation
    if (incoming_exact && !no_sparse && !c->shm && incoming * 8 < n_words_total * 32 && incoming < (1ull << 32) - 1) {
      // sparse route: fewer than one window in eight survives the quality filter -- scan + compact the surviving keys, then partition
      // those (the scatter kernels would pay their per-sub-tile work for every tile of the input)
      if (incoming == 0) return KMG_OK;
      uint64_t *d_keys = nullptr;
      kmg_status st = alloc_or_consolidate(c, reinterpret_cast<void **>(&d_keys), incoming * 8, "sparse keys");
      if (st != KMG_OK) return st;
      const size_t tmr = timer_begin(c, 0);
      unsigned long long h_n = 0;
      cudaError_t e = cudaMemsetAsync(c->d_stats + 7, 0, 8, c->stream);
      for (uint64_t tile0 = 0; e == cudaSuccess && tile0 < n_tiles;) {
        const uint64_t tiles = std::min<uint64_t>(n_tiles - tile0, ((1ull << 32) - 1) / ((uint64_t)TILE_WORDS * 32));
        ScanInput in;
        in.bases = c->d_bases + tile0 * TILE_WORDS; in.valid = c->d_valid + tile0 * TILE_WORDS;
        in.start = has_start ? c->d_start + tile0 * TILE_WORDS : nullptr;
        in.n_tiles = tiles; in.k = c->k;
        e = launch_scan_emit_keys(in, d_keys, c->d_stats + 7, c->stream);
        tile0 += tiles;
      }
      if (e == cudaSuccess) e = cudaMemcpyAsync(&h_n, c->d_stats + 7, 8, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      timer_end(c, tmr);
      if (e != cudaSuccess) { pool_free(c, d_keys); return cuda_fail(c, e, "sparse scan"); }
      if (h_n != incoming) { pool_free(c, d_keys); return fail(c, KMG_ERR_STATE, "sparse scan: emitted keys differ from the counted windows"); }
      st = keys_to_run(c, d_keys, nullptr, incoming);
      pool_free(c, d_keys);
      return st;
    }
    return scan_to_run(c, n_words_total, has_start);
  }
  uint64_t tile0 = 0;
  const size_t tmr = timer_begin(c);
  while (tile0 < n_tiles) {
    uint64_t want = (n_tiles - tile0) * TILE_WORDS * 32, granted = 0;
    if (incoming_exact && tile0 == 0) want = std::max<uint64_t>(incoming, 1);  // the whole stream adds exactly this many keys
    kmg_status s = reserve_capacity(c, want, &granted);
    if (s != KMG_OK) return s;
    uint64_t tiles = granted >= want ? n_tiles - tile0 : granted / ((uint64_t)TILE_WORDS * 32);
    if (tiles == 0) return fail(c, KMG_ERR_TABLE_FULL, "no table capacity left for even one tile");
    ScanInput in;
    in.bases = c->d_bases + tile0 * TILE_WORDS;
    in.valid = c->d_valid + tile0 * TILE_WORDS;
    in.start = has_start ? c->d_start + tile0 * TILE_WORDS : nullptr;
    in.n_tiles = tiles;
    in.k = c->k;
    if (c->use_dense) CU(c, launch_scan_dense(in, c->dense, c->d_counters, c->cfg.flags, c->stream));
    else CU(c, launch_scan_hash(in, c->table, c->d_counters, c->cfg.flags, c->stream));
    tile0 += tiles;
  }
  timer_end(c, tmr);
  return KMG_OK;
}

// ingest raw ASCII already on the device into the packed buffers and scan it.
// d_starts[n_starts] are byte offsets (minus base_offset) of record starts inside this chunk.
kmg_status count_device_chunk(kmg_ctx *c, const uint8_t *d_seq, const uint8_t *d_qual, const uint64_t *d_starts,
                              uint64_t n_starts, uint64_t base_offset, uint64_t n_bytes) {
  if (n_bytes == 0) return KMG_OK;
  const uint64_t n_words_total = round_up((n_bytes + 31) / 32, TILE_WORDS);
  kmg_status s = ensure_packed(c, n_words_total);
  if (s != KMG_OK) return s;
  const bool use_q = c->cfg.has_min_quality && d_qual != nullptr;
  const uint32_t thr = std::min<uint32_t>(255u, (uint32_t)c->cfg.min_quality + 33u);  // saturating_add (src/run.rs:538)
  CU(c, launch_ingest(d_seq, use_q ? d_qual : nullptr, n_bytes, thr, n_words_total, c->d_bases, c->d_valid, c->stream));
  const bool has_start = d_starts != nullptr && n_starts > 0;
  if (has_start) CU(c, launch_start_bits(d_starts, n_starts, base_offset, n_bytes, n_words_total, c->d_start, c->stream));
  return scan_packed(c, n_words_total, has_start);
}

kmg_status ensure_staging(kmg_ctx *c, bool need_pinned, bool need_qual, uint64_t need_bytes) {
  // staging slots are sized for the largest chunk seen so far (at most batch_bases), not for batch_bases up front
  const uint64_t want = std::min<uint64_t>(c->batch_bases, std::max<uint64_t>(need_bytes, 1u << 20)) + 64;
  if (want > c->staging_cap) {
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    for (auto &s : c->st) {
      cudaFree(s.d_seq); cudaFree(s.d_qual); s.d_seq = s.d_qual = nullptr;
      if (s.h_seq) cudaFreeHost(s.h_seq);
      if (s.h_qual) cudaFreeHost(s.h_qual);
      s.h_seq = s.h_qual = nullptr;
      s.h2d_pending = s.compute_pending = false;
    }
    c->staging_cap = want;
  }
  for (auto &s : c->st) {
    if (!s.d_seq) CU(c, cudaMalloc(&s.d_seq, c->staging_cap));
    if (need_qual && !s.d_qual) CU(c, cudaMalloc(&s.d_qual, c->staging_cap));
    if (!s.h2d_done) {
      CU(c, cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
      CU(c, cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming));
    }
    if (need_pinned && !s.h_seq) CU(c, cudaHostAlloc(&s.h_seq, c->staging_cap, cudaHostAllocDefault));
    if (need_pinned && need_qual && !s.h_qual) CU(c, cudaHostAlloc(&s.h_qual, c->staging_cap, cudaHostAllocDefault));
  }
  c->staging_ready = true;
  return KMG_OK;
}

kmg_status ensure_offsets(kmg_ctx *c, Staging &s, uint64_t n) {
  if (n <= s.off_cap) return KMG_OK;
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaStreamSynchronize(c->copy_stream));
  if (s.h_off) cudaFreeHost(s.h_off);
  if (s.d_off) cudaFree(s.d_off);
  s.h_off = nullptr; s.d_off = nullptr; s.off_cap = 0;
  uint64_t cap = std::max<uint64_t>(n, 1024) * 2;
  CU(c, cudaHostAlloc(&s.h_off, cap * 8, cudaHostAllocDefault));
  CU(c, cudaMalloc(&s.d_off, cap * 8));
  s.off_cap = cap;
  return KMG_OK;
}

bool is_pinned_host(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// ---- CRC32 (IEEE, reflected 0xEDB88320; src/index.rs:404-431), slicing-by-8 ----------------------------
struct Crc32 {
  uint32_t t[8][256];
  Crc32() {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
  uint32_t update(uint32_t crc, const uint8_t *p, size_t n) const {
    while (n >= 8) {
      uint32_t a, b;
      memcpy(&a, p, 4); memcpy(&b, p + 4, 4);
      a ^= crc;
      crc = t[7][a & 0xFF] ^ t[6][(a >> 8) & 0xFF] ^ t[5][(a >> 16) & 0xFF] ^ t[4][a >> 24] ^
            t[3][b & 0xFF] ^ t[2][(b >> 8) & 0xFF] ^ t[1][(b >> 16) & 0xFF] ^ t[0][b >> 24];
      p += 8; n -= 8;
    }
    while (n--) crc = t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
    return crc;
  }
};
const Crc32 &crc_tables() { static Crc32 c; return c; }

}  // namespace

// =================================================================================================
KMG_EXPORT uint32_t kmg_abi_version(void) { return KMG_ABI_VERSION; }
KMG_EXPORT uint32_t kmg_ctx_k(const kmg_ctx *c) { return c ? (uint32_t)c->k : 0u; }

KMG_EXPORT const char *kmg_status_string(kmg_status s) {
  switch (s) {
    case KMG_OK: return "ok";
    case KMG_ERR_INVALID_K: return "invalid k-mer length (must be 1..=32)";
    case KMG_ERR_INVALID_ARG: return "invalid argument";
    case KMG_ERR_CUDA: return "CUDA error";
    case KMG_ERR_OOM: return "out of memory";
    case KMG_ERR_TABLE_FULL: return "k-mer table full";
    case KMG_ERR_STATE: return "invalid state";
    case KMG_ERR_IO: return "I/O error";
    case KMG_ERR_ABI: return "ABI version mismatch";
    case KMG_ERR_CAPACITY: return "output capacity too small";
    case KMG_ERR_PARSE: return "sequence parse error";
  }
  return "unknown";
}

KMG_EXPORT const char *kmg_last_error(const kmg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

KMG_EXPORT uint32_t kmg_owner_of(uint64_t canonical_key, uint32_t n_shards) { return n_shards <= 1 ? 0 : part_of(canonical_key, n_shards); }

KMG_EXPORT void kmg_destroy(kmg_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  timers_collect(c);
  if (c->shm || c->sh_recv) shard_release(c);
  for (auto &r : c->runs) free_run(c, r);
  if (c->has_result) free_run(c, c->result);
  pool_release_idle(c);
  for (auto &kv : c->pool_live) cudaFree(kv.first);
  cudaFree(c->d_part); cudaFree(c->d_fine_cursor);
  cudaFree(c->table.slots); cudaFree(c->dense); cudaFree(c->d_counters); cudaFree(c->d_stats); cudaFree(c->d_hist); cudaFree(c->d_hist_ov); cudaFree(c->d_scan_tmp);
  cudaFree(c->d_bases); cudaFree(c->d_valid); cudaFree(c->d_start);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  for (auto &s : c->st) {
    if (s.h_seq) cudaFreeHost(s.h_seq);
    if (s.h_qual) cudaFreeHost(s.h_qual);
    if (s.h_off) cudaFreeHost(s.h_off);
    if (s.hp_bases) cudaFreeHost(s.hp_bases);
    if (s.hp_valid) cudaFreeHost(s.hp_valid);
    if (s.hp_start) cudaFreeHost(s.hp_start);
    cudaFree(s.d_seq); cudaFree(s.d_qual); cudaFree(s.d_off);
    cudaFree(s.dp_bases); cudaFree(s.dp_valid); cudaFree(s.dp_start);
    if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    if (s.compute_done) cudaEventDestroy(s.compute_done);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

KMG_EXPORT kmg_status kmg_create(const kmg_config *cfg, kmg_ctx **out) {
  if (!cfg || !out) return fail(nullptr, KMG_ERR_INVALID_ARG, "cfg and out must not be NULL");
  *out = nullptr;
  if (cfg->abi_version != KMG_ABI_VERSION)
    return fail(nullptr, KMG_ERR_ABI, "abi_version " + std::to_string(cfg->abi_version) + " != " + std::to_string(KMG_ABI_VERSION));
  if (cfg->k < 1 || cfg->k > 32)  // KmerLengthError (src/kmer.rs:100-111)
    return fail(nullptr, KMG_ERR_INVALID_K, "k-mer length " + std::to_string(cfg->k) + " is out of range (must be 1-32)");
  if ((cfg->flags & KMG_FLAG_FORCE_DIRECT) && (cfg->flags & KMG_FLAG_FORCE_HASH))
    return fail(nullptr, KMG_ERR_INVALID_ARG, "FORCE_DIRECT and FORCE_HASH are mutually exclusive");
  if ((cfg->flags & KMG_FLAG_FORCE_DIRECT) && cfg->k > (uint32_t)DENSE_MAX_K)
    return fail(nullptr, KMG_ERR_INVALID_ARG, "direct-indexed path supports k <= 14");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, KMG_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  kmg_ctx *c = new kmg_ctx();
  c->cfg = *cfg;
  c->k = (int)cfg->k;
  if (cfg->device >= 0) c->device = cfg->device; else cudaGetDevice(&c->device);
  auto bail = [&](kmg_status s) { g_create_error = c->err; kmg_destroy(c); return s; };
#define CUC(call)                                                                  \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) { cuda_fail(c, _e, #call); return bail(_e == cudaErrorMemoryAllocation ? KMG_ERR_OOM : KMG_ERR_CUDA); } \
  } while (0)
  CUC(cudaSetDevice(c->device));
  if (const char *dbg = getenv("KMG_DEBUG")) set_debug((uint32_t)atoi(dbg));
  if (cfg->stream) c->stream = (cudaStream_t)cfg->stream;
  else { CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CUC(cudaMalloc(&c->d_counters, CTR_N * sizeof(unsigned long long)));
  CUC(cudaMemsetAsync(c->d_counters, 0, CTR_N * sizeof(unsigned long long), c->stream));
  CUC(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));
  CUC(cudaHostAlloc(&c->h_counters, 16 * sizeof(unsigned long long), cudaHostAllocDefault));
  if (cfg->batch_bases) c->batch_bases = std::max<uint64_t>(round_up(cfg->batch_bases, 32), 4096);
  c->use_dense = (cfg->flags & KMG_FLAG_FORCE_DIRECT) ||
                 (!(cfg->flags & (KMG_FLAG_FORCE_HASH | KMG_FLAG_FORCE_PARTITIONED)) && cfg->k <= (uint32_t)DENSE_DEFAULT_MAX_K);
  if (c->use_dense) {
    c->dense_n = 1ull << (2 * c->k);
    CUC(cudaMalloc(&c->dense, c->dense_n * 8));
    CUC(cudaMemsetAsync(c->dense, 0, c->dense_n * 8, c->stream));
  }
  // hash table / partition buffers are allocated by decide_mode() at the first feeding call, when the
  // input size is known (small inputs: one HBM table; large inputs: the partitioned pipeline)
  if ((cfg->flags & KMG_FLAG_FORCE_PARTITIONED) && (cfg->flags & (KMG_FLAG_FORCE_HASH | KMG_FLAG_FORCE_DIRECT))) {
    c->err = "FORCE_PARTITIONED excludes FORCE_HASH / FORCE_DIRECT";
    return bail(KMG_ERR_INVALID_ARG);
  }
  if (cfg->parts_log2 > 20) { c->err = "parts_log2 must be <= 20"; return bail(KMG_ERR_INVALID_ARG); }
  if (cfg->flags & KMG_FLAG_FORCE_PARTITIONED) c->use_dense = false;
  CUC(cudaStreamSynchronize(c->stream));
#undef CUC
  *out = c;
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_reset(kmg_ctx *c) {
  if (!c) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->copy_stream));
  c->pending_slot = -1;  // a submitted but unscanned pre-packed batch is forgotten with everything else
  if (c->use_dense) CU(c, cudaMemsetAsync(c->dense, 0, c->dense_n * 8, c->stream));
  else if (c->mode == kmg_ctx::MODE_TABLE) CU(c, launch_table_init(c->table, c->stream));
  else if (c->mode == kmg_ctx::MODE_PARTITIONED) {
    CU(c, cudaStreamSynchronize(c->stream));
    for (auto &r : c->runs) free_run(c, r);
    c->runs.clear();
    if (c->has_result) free_run(c, c->result);
    c->has_result = false; c->pending_bytes = 0; c->n_consolidations = 0; c->dedup_ratio = 1.0;
    c->feed_kind = 0; c->poisoned = false;
    c->fused_valid = c->fused_cached = false;
  }
  CU(c, cudaMemsetAsync(c->d_counters, 0, CTR_N * sizeof(unsigned long long), c->stream));
  c->distinct_ub = 0; c->n_records = c->n_bases = c->h2d_bytes = 0;
  CU(c, cudaStreamSynchronize(c->stream));
  timers_collect(c);
  c->kernel_ms = c->cat_ms[0] = c->cat_ms[1] = c->cat_ms[2] = 0.0;
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_count_ascii_device(kmg_ctx *c, const uint8_t *d_seq, const uint8_t *d_qual, const uint64_t *d_offsets,
                                             uint64_t n_records, uint64_t n_bytes) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (c->shm) return fail(c, KMG_ERR_STATE, "context belongs to a shard group: feed it with kmg_shard_count_ascii(_device)");
  if (n_bytes && !d_seq) return fail(c, KMG_ERR_INVALID_ARG, "d_seq is NULL");
  if (((uintptr_t)d_seq & 15) || ((uintptr_t)d_qual & 15)) return fail(c, KMG_ERR_INVALID_ARG, "d_seq / d_qual must be 16-byte aligned");
  CU(c, cudaSetDevice(c->device));
  // offsets[0] (== 0) needs no start bit; a single record needs none at all
  kmg_status s = count_device_chunk(c, d_seq, d_qual, n_records > 1 ? d_offsets : nullptr, n_records, 0, n_bytes);
  if (s != KMG_OK) return s;
  c->n_records += n_records; c->n_bases += n_bytes;
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_count_ascii(kmg_ctx *c, const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets, uint64_t n_records) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (c->shm) return fail(c, KMG_ERR_STATE, "context belongs to a shard group: feed it with kmg_shard_count_ascii(_device)");
  if (n_records == 0) return KMG_OK;
  if (!offsets || !seq) {
    if (offsets && offsets[n_records] == offsets[0]) { c->n_records += n_records; return KMG_OK; }
    return fail(c, KMG_ERR_INVALID_ARG, "seq / offsets must not be NULL");
  }
  for (uint64_t r = 0; r < n_records; ++r)
    if (offsets[r + 1] < offsets[r]) return fail(c, KMG_ERR_INVALID_ARG, "offsets must be non-decreasing");
  CU(c, cudaSetDevice(c->device));
  const uint64_t begin = offsets[0], end = offsets[n_records];
  const bool use_q = c->cfg.has_min_quality && qual != nullptr;
  const bool src_pinned = is_pinned_host(seq + begin) && (!use_q || is_pinned_host(qual + begin));
  kmg_status s = KMG_OK;
  if (use_q && !c->cfg.expected_distinct && c->mode == kmg_ctx::MODE_UNDECIDED && !c->use_dense)
    c->plan_scale = (double)(end - begin) / (double)std::min<uint64_t>(end - begin, c->batch_bases);  // decided after the first chunk's ingest (scan_packed)
  else s = decide_mode(c, end - begin);  // plan for the whole call, not for its first staging chunk
  if (s != KMG_OK) return s;
  s = ensure_staging(c, !src_pinned, use_q, end - begin);
  if (s != KMG_OK) return s;
  const uint64_t K1 = (uint64_t)c->k - 1;
  const uint64_t B = c->batch_bases;  // bytes per chunk including the k-1 overlap
  const uint64_t step = B - K1;
  // chunk plan first, then a software pipeline: the H2D copy of chunk i+1 is queued BEFORE chunk i is processed
  // (processing contains host syncs), so copies overlap the kernels of the previous chunk.
  struct Chunk { uint64_t pos, len, r_lo, nrec; };
  std::vector<Chunk> chunks;
  {
    uint64_t r_lo = 0;
    for (uint64_t pos = begin; pos < end; pos += step) {
      const uint64_t len = std::min<uint64_t>(B, end - pos);
      if (pos != begin && len <= K1) break;  // only the overlap is left: no new windows
      while (r_lo < n_records && offsets[r_lo] <= pos) ++r_lo;  // records with a start strictly inside (pos, pos+len)
      uint64_t r_hi = r_lo;
      while (r_hi < n_records && offsets[r_hi] < pos + len) ++r_hi;
      chunks.push_back(Chunk{pos, len, r_lo, r_hi - r_lo});
    }
  }
  auto stage_chunk = [&](const Chunk &ch, Staging &st) -> kmg_status {
    if (st.h2d_pending) { CU(c, cudaEventSynchronize(st.h2d_done)); st.h2d_pending = false; }
    if (st.compute_pending) { CU(c, cudaStreamWaitEvent(c->copy_stream, st.compute_done, 0)); }
    kmg_status es = ensure_offsets(c, st, ch.nrec);
    if (es != KMG_OK) return es;
    const uint8_t *src_seq = seq + ch.pos, *src_qual = use_q ? qual + ch.pos : nullptr;
    if (!src_pinned) {
      memcpy(st.h_seq, src_seq, ch.len); src_seq = st.h_seq;
      if (use_q) { memcpy(st.h_qual, src_qual, ch.len); src_qual = st.h_qual; }
    }
    CU(c, cudaMemcpyAsync(st.d_seq, src_seq, ch.len, cudaMemcpyHostToDevice, c->copy_stream));
    if (use_q) CU(c, cudaMemcpyAsync(st.d_qual, src_qual, ch.len, cudaMemcpyHostToDevice, c->copy_stream));
    if (ch.nrec) {
      memcpy(st.h_off, offsets + ch.r_lo, ch.nrec * 8);
      CU(c, cudaMemcpyAsync(st.d_off, st.h_off, ch.nrec * 8, cudaMemcpyHostToDevice, c->copy_stream));
    }
    c->h2d_bytes += ch.len * (use_q ? 2 : 1) + ch.nrec * 8;
    CU(c, cudaEventRecord(st.h2d_done, c->copy_stream));
    st.h2d_pending = true;
    return KMG_OK;
  };
  // ring of N_STAGE slots: chunk i lives in slot (slot0 + i) % N_STAGE, and chunks i+1 .. i+N_STAGE-1 are already
  // queued on the copy stream while chunk i is processed (slot reuse waits for the previous user's compute_done)
  const uint32_t slot0 = c->next_ascii;
  size_t staged = 0;
  for (size_t i = 0; i < chunks.size(); ++i) {
    for (; staged < chunks.size() && staged < i + N_STAGE; ++staged)
      if ((s = stage_chunk(chunks[staged], c->st[(slot0 + staged) % N_STAGE])) != KMG_OK) return s;
    Staging &st = c->st[(slot0 + i) % N_STAGE];
    const Chunk &ch = chunks[i];
    CU(c, cudaStreamWaitEvent(c->stream, st.h2d_done, 0));
    s = count_device_chunk(c, st.d_seq, use_q ? st.d_qual : nullptr, ch.nrec ? st.d_off : nullptr, ch.nrec, ch.pos, ch.len);
    if (s != KMG_OK) return s;
    CU(c, cudaEventRecord(st.compute_done, c->stream));
    st.compute_pending = true;
  }
  c->next_ascii = (slot0 + (uint32_t)chunks.size()) % N_STAGE;
  c->n_records += n_records; c->n_bases += end - begin;
  if (src_pinned)  // the copies read the CALLER's buffer: it must be reusable when this call returns
    for (auto &st : c->st)
      if (st.h2d_pending) { CU(c, cudaEventSynchronize(st.h2d_done)); st.h2d_pending = false; }
  return KMG_OK;
}


// FASTA / FASTQ file image (HOST bytes, e.g. an mmap'ed file: src/mmap.rs:33-71) -> counts, with the records found ON THE DEVICE:
// the raw bytes go over PCIe in chunks cut at line (FASTA) / record (FASTQ) boundaries, a few small kernels classify the lines,
// compact sequence (+ quality) bytes, mark the record starts, and the usual ingest + scan follow.  Replaces reader::read /
// read_with_quality (src/reader.rs:82-247: whole file -> Vec<Record> -> Vec<SequenceWithQuality>, two host copies) for
// well-formed input; multi-line FASTQ, or a line longer than a chunk, is refused with KMG_ERR_PARSE -- nothing has been
// counted then only if it happens in the first chunk, so callers fall back to kmg_parse_fastx + kmg_count_ascii after kmg_reset.
KMG_EXPORT kmg_status kmg_count_fastx(kmg_ctx *c, const uint8_t *buf, uint64_t len, int is_fastq, uint64_t *n_records_out) {
  if (!c || (len && !buf)) return KMG_ERR_INVALID_ARG;
  if (c->shm) return fail(c, KMG_ERR_STATE, "context belongs to a shard group: feed it with kmg_shard_count_ascii(_device)");
  if (n_records_out) *n_records_out = 0;
  if (len == 0) return KMG_OK;
  const uint8_t marker = is_fastq ? '@' : '>';
  if (buf[0] != marker) return fail(c, KMG_ERR_PARSE, is_fastq ? "FASTQ record does not start with '@' (record 0)" : "FASTA record does not start with '>' (record 0)");
  CU(c, cudaSetDevice(c->device));
  const bool use_q = is_fastq && c->cfg.has_min_quality;
  const uint64_t K1 = (uint64_t)c->k - 1;
  const uint64_t CB = std::min<uint64_t>(c->batch_bases, 1ull << 30);
  // ---- chunk plan: cuts at line starts (FASTA) / record starts (FASTQ)
  struct Chunk { uint64_t pos, len; };
  std::vector<Chunk> chunks;
  for (uint64_t pos = 0; pos < len;) {
    uint64_t end = std::min(len, pos + CB);
    if (end < len) {
      const void *nl = memrchr(buf + pos, '\n', (size_t)(end - pos));
      if (!nl) return fail(c, KMG_ERR_PARSE, "a line is longer than a staging chunk (raise batch_bases or use the host splitter)");
      end = (uint64_t)((const uint8_t *)nl - buf) + 1;
      if (is_fastq) {  // step back line by line to a line that opens a record: '@' first, and the line after next begins with '+'
        uint64_t cut = end;
        bool found = false;
        for (int tries = 0; tries < 64 && cut > pos; ++tries) {
          if (buf[cut] == '@') {
            const void *l1 = memchr(buf + cut, '\n', (size_t)(len - cut));
            const void *l2 = l1 ? memchr((const uint8_t *)l1 + 1, '\n', (size_t)(buf + len - ((const uint8_t *)l1 + 1))) : nullptr;
            if (l2 && (const uint8_t *)l2 + 1 < buf + len && ((const uint8_t *)l2)[1] == '+') { found = true; break; }
          }
          const void *prev = cut >= 2 ? memrchr(buf + pos, '\n', (size_t)(cut - 1 - pos)) : nullptr;  // start of the previous line
          if (!prev) break;
          cut = (uint64_t)((const uint8_t *)prev - buf) + 1;
        }
        if (!found || cut <= pos) return fail(c, KMG_ERR_PARSE, "no FASTQ record boundary found near a chunk end (multi-line FASTQ? use the host splitter)");
        end = cut;
      }
    }
    chunks.push_back(Chunk{pos, end - pos});
    pos = end;
  }
  uint64_t max_len = 0;
  for (auto &ch : chunks) max_len = std::max(max_len, ch.len);
  if (max_len >= (1ull << 32)) return fail(c, KMG_ERR_INVALID_ARG, "staging chunks must stay below 4 GiB");
  // ---- plan the table / partitions for the whole file
  kmg_status s = KMG_OK;
  if (use_q && !c->cfg.expected_distinct && c->mode == kmg_ctx::MODE_UNDECIDED && !c->use_dense) c->plan_scale = (double)len / (double)chunks[0].len;
  else s = decide_mode(c, is_fastq ? len / 2 : len);
  if (s != KMG_OK) return s;
  const bool src_pinned = is_pinned_host(buf);
  s = ensure_staging(c, !src_pinned, false, max_len);
  if (s != KMG_OK) return s;
  // ---- parse buffers (pooled)
  uint8_t *d_kind = nullptr, *d_keep = nullptr, *d_seq = nullptr, *d_qual = nullptr, *d_mark = nullptr, *d_carry = nullptr;
  uint32_t *d_lineno = nullptr, *d_pos_s = nullptr, *d_pos_q = nullptr, *d_misc = nullptr, *h_tot = nullptr;
  void *d_tmp = nullptr;
  const size_t tmp_bytes = fastx_scan_tmp_bytes(max_len);
  auto release = [&]() {
    pool_free(c, d_kind); pool_free(c, d_keep); pool_free(c, d_seq); pool_free(c, d_qual); pool_free(c, d_mark); pool_free(c, d_carry);
    pool_free(c, d_lineno); pool_free(c, d_pos_s); pool_free(c, d_pos_q); pool_free(c, d_misc); pool_free(c, d_tmp);
    if (h_tot) cudaFreeHost(h_tot);
  };
  cudaError_t e = pool_alloc(c, &d_kind, max_len);
  if (e == cudaSuccess) e = pool_alloc(c, &d_keep, max_len);
  if (e == cudaSuccess) e = pool_alloc(c, &d_pos_s, max_len * 4);
  if (e == cudaSuccess && is_fastq) e = pool_alloc(c, &d_lineno, max_len * 4);
  if (e == cudaSuccess && is_fastq) e = pool_alloc(c, &d_pos_q, max_len * 4);
  if (e == cudaSuccess) e = pool_alloc(c, &d_seq, max_len + K1 + 64);
  if (e == cudaSuccess && use_q) e = pool_alloc(c, &d_qual, max_len + K1 + 64);
  if (e == cudaSuccess) e = pool_alloc(c, &d_mark, max_len + K1 + 64);
  if (e == cudaSuccess) e = pool_alloc(c, &d_carry, 2 * 64);
  if (e == cudaSuccess) e = pool_alloc(c, &d_misc, 64);   // [0..1] n_records (u64), [2..3] totals, [4] error word
  if (e == cudaSuccess) e = pool_alloc(c, &d_tmp, tmp_bytes);
  if (e == cudaSuccess) e = cudaHostAlloc(&h_tot, 16, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_misc, 0, 64, c->stream);
  if (e != cudaSuccess) { release(); return cuda_fail(c, e, "cudaMalloc(fastx parse buffers)"); }
  unsigned long long *d_nrec = reinterpret_cast<unsigned long long *>(d_misc);
  uint32_t *d_err = d_misc + 4;
  const uint32_t thr = std::min<uint32_t>(255u, (uint32_t)c->cfg.min_quality + 33u);
  auto stage_chunk = [&](const Chunk &ch, Staging &st) -> kmg_status {
    if (st.h2d_pending) { CU(c, cudaEventSynchronize(st.h2d_done)); st.h2d_pending = false; }
    if (st.compute_pending) { CU(c, cudaStreamWaitEvent(c->copy_stream, st.compute_done, 0)); }
    const uint8_t *src = buf + ch.pos;
    if (!src_pinned) { memcpy(st.h_seq, src, ch.len); src = st.h_seq; }
    CU(c, cudaMemcpyAsync(st.d_seq, src, ch.len, cudaMemcpyHostToDevice, c->copy_stream));
    c->h2d_bytes += ch.len;
    CU(c, cudaEventRecord(st.h2d_done, c->copy_stream));
    st.h2d_pending = true;
    return KMG_OK;
  };
  const uint32_t slot0 = c->next_ascii;
  size_t staged = 0;
  uint64_t prev_n = 0, bases = 0;
  auto body = [&]() -> kmg_status {
    for (size_t i = 0; i < chunks.size(); ++i) {
      for (; staged < chunks.size() && staged < i + N_STAGE; ++staged) {
        kmg_status ss = stage_chunk(chunks[staged], c->st[(slot0 + staged) % N_STAGE]);
        if (ss != KMG_OK) return ss;
      }
      Staging &st = c->st[(slot0 + i) % N_STAGE];
      const Chunk &ch = chunks[i];
      CU(c, cudaStreamWaitEvent(c->stream, st.h2d_done, 0));
      // a FASTA chunk that does not open with a header continues the previous chunk's last record: its last k-1 bases (and their
      // record marks) are put in front, so that the windows spanning the cut are counted -- and none twice (k-1 bases hold no window)
      const uint64_t carry = (!is_fastq && i > 0 && buf[ch.pos] != '>') ? std::min<uint64_t>(K1, prev_n) : 0;
      CU(c, cudaMemsetAsync(d_mark, 0, ch.len + K1 + 64, c->stream));
      CU(c, cudaMemsetAsync(d_err, 0, 4, c->stream));
      if (carry) {
        CU(c, cudaMemcpyAsync(d_seq, d_carry + (K1 - carry), carry, cudaMemcpyDeviceToDevice, c->stream));
        CU(c, cudaMemcpyAsync(d_mark, d_carry + 64 + (K1 - carry), carry, cudaMemcpyDeviceToDevice, c->stream));
      }
      CU(c, launch_fastx_parse(st.d_seq, ch.len, is_fastq, d_kind, d_keep, d_lineno, d_pos_s, d_pos_q, d_tmp, tmp_bytes, carry, d_seq, d_qual, d_mark,
                               d_nrec, d_err, h_tot, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
      const uint32_t total_s = h_tot[0], total_q = h_tot[1], err = h_tot[2];
      if (err & 1u) return fail(c, KMG_ERR_PARSE, "FASTQ record does not start with '@' (multi-line FASTQ is not handled by the device parser)");
      if (err & 2u) return fail(c, KMG_ERR_PARSE, "FASTQ record without a '+' separator line where one is expected");
      if (is_fastq && total_s != total_q) return fail(c, KMG_ERR_PARSE, "sequence and quality lengths differ");
      // a record whose header closed the previous chunk starts with this chunk's first base (the scatter below may raise the flag anew)
      if (!is_fastq && total_s) CU(c, launch_fastx_apply_pending(d_err + 1, d_mark + carry, c->stream));
      CU(c, launch_fastx_scatter(st.d_seq, ch.len, is_fastq, d_kind, d_lineno, d_keep, d_pos_s, d_pos_q, carry, total_s, d_seq, use_q ? d_qual : nullptr,
                                 d_mark, d_nrec, d_err, c->stream));
      CU(c, cudaEventRecord(st.compute_done, c->stream));  // the raw chunk has been consumed
      st.compute_pending = true;
      const uint64_t n = carry + total_s;
      bases += total_s;
      if (n) {
        const uint64_t n_words_total = round_up((n + 31) / 32, TILE_WORDS);
        kmg_status ps = ensure_packed(c, n_words_total);
        if (ps != KMG_OK) return ps;
        CU(c, launch_ingest(d_seq, use_q ? d_qual : nullptr, n, thr, n_words_total, c->d_bases, c->d_valid, c->stream));
        CU(c, launch_marks_to_bits(d_mark, n, n_words_total, c->d_start, c->stream));
        if (!is_fastq) {  // keep the tail for the next chunk
          const uint64_t t = std::min<uint64_t>(K1, n);
          if (t) {
            CU(c, cudaMemcpyAsync(d_carry + (K1 - t), d_seq + (n - t), t, cudaMemcpyDeviceToDevice, c->stream));
            CU(c, cudaMemcpyAsync(d_carry + 64 + (K1 - t), d_mark + (n - t), t, cudaMemcpyDeviceToDevice, c->stream));
          }
        }
        ps = scan_packed(c, n_words_total, true);
        if (ps != KMG_OK) return ps;
        if (is_fastq && use_q) {  // a quality mismatch inside a record shows up at the next record start (checked in the scatter)
          uint32_t e2 = 0;
          CU(c, cudaMemcpyAsync(&e2, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
          CU(c, cudaStreamSynchronize(c->stream));
          if (e2 & 8u) return fail(c, KMG_ERR_PARSE, "sequence and quality lengths differ");
        }
      }
      prev_n = n;
    }
    return KMG_OK;
  };
  s = body();
  unsigned long long n_rec = 0;
  if (s == KMG_OK) {
    e = cudaMemcpyAsync(&n_rec, d_nrec, 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) s = cuda_fail(c, e, "fastx record count");
  } else {
    cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->copy_stream);
  }
  c->next_ascii = (slot0 + (uint32_t)chunks.size()) % N_STAGE;
  for (auto &st : c->st) if (st.h2d_pending && src_pinned) { cudaEventSynchronize(st.h2d_done); st.h2d_pending = false; }
  release();
  if (s != KMG_OK) return s;
  c->n_records += n_rec; c->n_bases += bases;
  if (n_records_out) *n_records_out = n_rec;
  return KMG_OK;
}

// ---- pre-packed, zero-copy feed (the Rust reader layer packs 2-bit words + masks straight into pinned memory) ------------------
// N_STAGE pinned slots, each with its own packed device buffers.  kmg_submit_batch queues the slot's H2D copy on the copy stream
// at once and then scans the batch submitted BEFORE it, so the copy of batch i overlaps the kernels of batch i-1 while the
// producer fills batch i+1; the last batch is scanned by whichever call needs the table next (finalize, export, ...).
namespace {
kmg_status scan_pending_batch(kmg_ctx *c) {
  if (c->pending_slot < 0) return KMG_OK;
  Staging &s = c->st[c->pending_slot];
  const uint64_t n_words_total = c->pending_words;
  c->pending_slot = -1;
  CU(c, cudaStreamWaitEvent(c->stream, s.h2d_done, 0));
  // the scan reads the context's packed stream: point it at this slot's buffers for the duration of the call
  uint64_t *sb = c->d_bases; uint32_t *sv = c->d_valid, *ss = c->d_start; const uint64_t sw = c->packed_words;
  c->d_bases = s.dp_bases; c->d_valid = s.dp_valid; c->d_start = s.dp_start; c->packed_words = c->dp_words;
  const kmg_status st = scan_packed(c, n_words_total, true);
  c->d_bases = sb; c->d_valid = sv; c->d_start = ss; c->packed_words = sw;
  if (st != KMG_OK) return st;
  CU(c, cudaEventRecord(s.compute_done, c->stream));
  s.compute_pending = true;
  return KMG_OK;
}
}  // namespace

KMG_EXPORT kmg_status kmg_acquire_batch(kmg_ctx *c, kmg_batch *b) {
  if (!c || !b) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  const uint64_t words = round_up((c->batch_bases + 31) / 32, TILE_WORDS);
  if (!c->packed_feed_ready) {
    for (int i = 0; i < N_STAGE; ++i) {
      Staging &s = c->st[i];
      CU(c, cudaHostAlloc(&s.hp_bases, words * 8, cudaHostAllocDefault));
      CU(c, cudaHostAlloc(&s.hp_valid, words * 4, cudaHostAllocDefault));
      CU(c, cudaHostAlloc(&s.hp_start, words * 4, cudaHostAllocDefault));
      memset(s.hp_bases, 0, words * 8); memset(s.hp_valid, 0, words * 4); memset(s.hp_start, 0, words * 4);
      CU(c, cudaMalloc(&s.dp_bases, (LEAD_BASE_WORDS + words) * 8));
      CU(c, cudaMalloc(&s.dp_valid, (LEAD_MASK_WORDS + words) * 4));
      CU(c, cudaMalloc(&s.dp_start, (LEAD_MASK_WORDS + words) * 4));
      CU(c, cudaMemsetAsync(s.dp_bases, 0, LEAD_BASE_WORDS * 8, c->stream));
      CU(c, cudaMemsetAsync(s.dp_valid, 0, LEAD_MASK_WORDS * 4, c->stream));
      CU(c, cudaMemsetAsync(s.dp_start, 0, LEAD_MASK_WORDS * 4, c->stream));
      if (!s.h2d_done) {
        CU(c, cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming));
      }
    }
    CU(c, cudaStreamSynchronize(c->stream));
    c->dp_words = words;
    c->packed_feed_ready = true;
  }
  if ((int)c->next_slot == c->pending_slot) {  // the ring is full: the oldest submitted batch must be scanned before its slot is handed out again
    kmg_status st = scan_pending_batch(c);
    if (st != KMG_OK) return st;
  }
  Staging &s = c->st[c->next_slot];
  if (s.h2d_pending) { CU(c, cudaEventSynchronize(s.h2d_done)); s.h2d_pending = false; }  // the pinned words have left the host
  if (s.hp_used_words) {  // hand out zeroed words (producers OR their bits in): only what the previous user touched
    memset(s.hp_bases, 0, s.hp_used_words * 8); memset(s.hp_valid, 0, s.hp_used_words * 4); memset(s.hp_start, 0, s.hp_used_words * 4);
    s.hp_used_words = 0;
  }
  b->bases2bit = s.hp_bases; b->valid_bits = s.hp_valid; b->start_bits = s.hp_start;
  b->capacity_bases = c->batch_bases; b->n_bases = 0; b->n_records = 0; b->slot = c->next_slot;
  c->next_slot = (c->next_slot + 1) % N_STAGE;
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_submit_batch(kmg_ctx *c, const kmg_batch *b) {
  if (!c || !b) return KMG_ERR_INVALID_ARG;
  if (c->shm) return fail(c, KMG_ERR_STATE, "context belongs to a shard group: feed it with kmg_shard_count_ascii(_device)");
  if (b->slot >= (uint32_t)N_STAGE || !c->packed_feed_ready || b->bases2bit != c->st[b->slot].hp_bases)
    return fail(c, KMG_ERR_STATE, "batch was not obtained from kmg_acquire_batch");
  if (b->n_bases > b->capacity_bases) return fail(c, KMG_ERR_INVALID_ARG, "n_bases exceeds the batch capacity");
  if (b->n_bases == 0) return KMG_OK;
  CU(c, cudaSetDevice(c->device));
  Staging &s = c->st[b->slot];
  const uint64_t n_words = (b->n_bases + 31) / 32, n_words_total = round_up(n_words, TILE_WORDS);
  s.hp_used_words = n_words_total;
  // the copy goes out now, on the copy stream, as soon as the slot's device buffers are free (its previous scan has finished)
  if (s.compute_pending) { CU(c, cudaStreamWaitEvent(c->copy_stream, s.compute_done, 0)); s.compute_pending = false; }
  CU(c, cudaMemcpyAsync(s.dp_bases + LEAD_BASE_WORDS, s.hp_bases, n_words_total * 8, cudaMemcpyHostToDevice, c->copy_stream));
  CU(c, cudaMemcpyAsync(s.dp_valid + LEAD_MASK_WORDS, s.hp_valid, n_words_total * 4, cudaMemcpyHostToDevice, c->copy_stream));
  CU(c, cudaMemcpyAsync(s.dp_start + LEAD_MASK_WORDS, s.hp_start, n_words_total * 4, cudaMemcpyHostToDevice, c->copy_stream));
  CU(c, cudaEventRecord(s.h2d_done, c->copy_stream));
  s.h2d_pending = true;
  c->h2d_bytes += n_words_total * 16;
  c->n_records += b->n_records; c->n_bases += b->n_bases;
  // ... and while it travels, the batch submitted before this one is scanned
  kmg_status st = scan_pending_batch(c);
  c->pending_slot = (int)b->slot; c->pending_words = n_words_total;
  return st;
}

KMG_EXPORT kmg_status kmg_insert_keys_device(kmg_ctx *c, const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (c->shm) return fail(c, KMG_ERR_STATE, "context belongs to a shard group: feed it with kmg_shard_count_ascii(_device)");
  if (n == 0) return KMG_OK;
  if (!d_keys) return fail(c, KMG_ERR_INVALID_ARG, "d_keys is NULL");
  CU(c, cudaSetDevice(c->device));
  kmg_status ms = decide_mode(c, n);
  if (ms != KMG_OK) return ms;
  if (c->use_dense) { CU(c, launch_insert_keys_dense(c->dense, d_keys, d_counts, n, c->d_counters, c->stream)); return KMG_OK; }
  if (c->mode == kmg_ctx::MODE_PARTITIONED) return keys_to_run(c, d_keys, d_counts, n);
  uint64_t done = 0;
  while (done < n) {
    uint64_t granted = 0;
    kmg_status s = reserve_capacity(c, n - done, &granted);
    if (s != KMG_OK) return s;
    CU(c, launch_insert_keys(c->table, d_keys + done, d_counts ? d_counts + done : nullptr, granted, c->d_counters, c->stream));
    done += granted;
  }
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_partition_plan(kmg_ctx *c, uint64_t expected_keys, uint32_t *n_coarse, uint32_t *n_sub) {
  if (!c) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  if (c->mode == kmg_ctx::MODE_UNDECIDED) {
    if (c->use_dense) return fail(c, KMG_ERR_STATE, "context uses the direct-indexed path; create it with KMG_FLAG_FORCE_PARTITIONED");
    c->cfg.flags |= KMG_FLAG_FORCE_PARTITIONED;
    c->cfg.flags &= ~(uint32_t)KMG_FLAG_FORCE_HASH;
    kmg_status s = decide_mode(c, expected_keys);
    if (s != KMG_OK) return s;
  }
  if (c->mode != kmg_ctx::MODE_PARTITIONED) return fail(c, KMG_ERR_STATE, "context is not on the partitioned path");
  if (n_coarse) *n_coarse = c->n_coarse;
  if (n_sub) *n_sub = c->n_sub;
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_adopt_coarse_device(kmg_ctx *c, const uint64_t *d_keys, const uint64_t *bin_counts, uint32_t n_bins, uint64_t n) {
  if (!c || !bin_counts) return KMG_ERR_INVALID_ARG;
  if (c->shm) return fail(c, KMG_ERR_STATE, "context belongs to a shard group: feed it with kmg_shard_count_ascii(_device)");
  if (c->mode != kmg_ctx::MODE_PARTITIONED) return fail(c, KMG_ERR_STATE, "call kmg_partition_plan first");
  if (n_bins != c->n_coarse) return fail(c, KMG_ERR_INVALID_ARG, "n_bins must equal the context's coarse partition count");
  if (n >= (1ull << 32)) return fail(c, KMG_ERR_INVALID_ARG, "at most 2^32-1 keys per call");
  std::vector<uint64_t> off(n_bins + 1, 0);
  for (uint32_t b = 0; b < n_bins; ++b) off[b + 1] = off[b] + bin_counts[b];
  if (off[n_bins] != n) return fail(c, KMG_ERR_INVALID_ARG, "bin_counts do not add up to n");
  if (n == 0) return KMG_OK;
  if (!d_keys) return fail(c, KMG_ERR_INVALID_ARG, "d_keys is NULL");
  if (c->poisoned) return fail(c, KMG_ERR_STATE, "context is inconsistent after a failed re-split: kmg_reset it");
  if (c->feed_kind == 1) return fail(c, KMG_ERR_STATE, "context already holds runs binned by its own coarse function: adopted blocks (binned by owner and bin) cannot be mixed in");
  c->feed_kind = 2;
  CU(c, cudaSetDevice(c->device));
  const size_t tmr = timer_begin(c, 0);
  kmg_status s = refine_to_run(c, const_cast<uint64_t *>(d_keys), nullptr, off, /*owns=*/false, /*sync=*/true, nullptr, nullptr, /*in_keys=*/true);
  timer_end(c, tmr);
  return s;
}

KMG_EXPORT kmg_status kmg_extract_keys_device(kmg_ctx *c, const uint8_t *d_seq, const uint8_t *d_qual, const uint64_t *d_offsets,
                                              uint64_t n_records, uint64_t n_bytes, uint32_t n_shards, uint64_t *d_keys_out,
                                              uint64_t cap, uint64_t *shard_counts_out) {
  if (!c || !shard_counts_out) return KMG_ERR_INVALID_ARG;
  if (n_shards < 1 || n_shards > (uint32_t)MAX_PARTS) return fail(c, KMG_ERR_INVALID_ARG, "n_shards must be in 1..=8192");
  if (((uintptr_t)d_seq & 15) || ((uintptr_t)d_qual & 15)) return fail(c, KMG_ERR_INVALID_ARG, "d_seq / d_qual must be 16-byte aligned");
  CU(c, cudaSetDevice(c->device));
  for (uint32_t i = 0; i < n_shards; ++i) shard_counts_out[i] = 0;
  if (n_bytes == 0) return KMG_OK;
  const uint64_t n_words_total = round_up((n_bytes + 31) / 32, TILE_WORDS);
  kmg_status s = ensure_packed(c, n_words_total);
  if (s != KMG_OK) return s;
  const bool use_q = c->cfg.has_min_quality && d_qual != nullptr;
  const uint32_t thr = std::min<uint32_t>(255u, (uint32_t)c->cfg.min_quality + 33u);
  CU(c, launch_ingest(d_seq, use_q ? d_qual : nullptr, n_bytes, thr, n_words_total, c->d_bases, c->d_valid, c->stream));
  const bool has_start = d_offsets != nullptr && n_records > 1;
  if (has_start) CU(c, launch_start_bits(d_offsets, n_records, 0, n_bytes, n_words_total, c->d_start, c->stream));
  if (n_bytes >= (1ull << 32)) return fail(c, KMG_ERR_INVALID_ARG, "kmg_extract_keys_device handles < 2^32 bytes per call");
  unsigned long long *d_pc = nullptr;
  CU(c, pool_alloc(c, &d_pc, 3 * (size_t)n_shards * 8 + 64));  // pooled: a cudaMalloc / cudaFree pair per call synchronises the device
  unsigned long long *d_cur = d_pc + n_shards;
  unsigned long long *d_ps = d_cur + n_shards;
  unsigned long long *d_ctr = d_ps + n_shards;  // scratch counters (windows of this call only)
  CU(c, cudaMemsetAsync(d_pc, 0, 3 * (size_t)n_shards * 8 + 64, c->stream));
  ScanInput in{c->d_bases, c->d_valid, has_start ? c->d_start : nullptr, n_words_total / TILE_WORDS, c->k};
  cudaError_t e = launch_scan_partition(in, n_shards, false, d_pc, d_ps, d_cur, nullptr, d_ctr, c->stream);
  std::vector<unsigned long long> h(n_shards);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_pc, n_shards * 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { pool_free(c, d_pc); return cuda_fail(c, e, "partition count pass"); }
  uint64_t total = 0;
  std::vector<unsigned long long> prefix(n_shards);
  for (uint32_t i = 0; i < n_shards; ++i) { prefix[i] = total; total += h[i]; shard_counts_out[i] = h[i]; }
  if (total > cap || (total && !d_keys_out)) { pool_free(c, d_pc); return fail(c, KMG_ERR_CAPACITY, "d_keys_out too small: need " + std::to_string(total)); }
  e = cudaMemcpyAsync(d_ps, prefix.data(), n_shards * 8, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = launch_scan_partition(in, n_shards, true, d_pc, d_ps, d_cur, d_keys_out, d_ctr, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  pool_free(c, d_pc);
  if (e != cudaSuccess) return cuda_fail(c, e, "partition scatter pass");
  c->n_records += n_records; c->n_bases += n_bytes;
  return KMG_OK;
}


// =================================================================================================
// Sharded (multi-GPU) counting: one context per GPU -- processes (torchrun, a Rust host with one process per device) or
// threads of one process -- joined into a group on ONE box.  The count table shards by hash: global coarse bin
// g = coarse_of_mix(mix, world * P1) belongs to rank g / P1.  Every rank scatters its keys into its own send buffer, one
// region per global bin; the exchange is FUSED into the owner's refine kernel (A2), whose tile loads read the owner's regions
// straight out of all ranks' send buffers through P2P-mapped pointers (NVLink loads, prefetched a tile ahead), so no key is
// ever copied: it crosses NVLink once, on its way into the kernel that partitions it further.  Only metadata crosses between
// the ranks on the host: a POSIX shared-memory segment carries the region sizes, flags, summaries, histograms and a
// sense-reversing barrier.  No NCCL, no staging copies, no count pass on the fast path.
// (Measured alternative, round 2: the scatter kernel writing into the owners' buffers with P2P STORES.  A sub-tile yields only
// ~8 keys per bin, and 64-byte NVLink writes cost more than they move: 4.45 ms per 310 M-base round against 2.05 ms for the
// local scatter; fewer bins made the stores longer but the owner's refine slower -- profiles/r2_summary.md.)
// =================================================================================================
constexpr uint32_t SHARD_MAX_WORLD = 8;
constexpr uint32_t SHARD_MAX_BINS = 2048;        // world * P1 (the single-scan rows kernel's limit)
constexpr uint32_t SHARD_HIST_MAX = 2 * 65536;   // (value, frequency) pairs a rank can post
constexpr uint32_t SHARD_MAGIC = 0x4b4d4753u;
constexpr double SHARD_TIMEOUT_S = 120.0;

struct ShardRankSlot {
  int32_t pid, device;
  cudaIpcMemHandle_t handle;
  uint64_t raw_ptr, recv_entries;
  uint64_t round_bytes, round_flag;
  uint64_t summary[8];
  uint64_t hist_n;
};
struct ShardShm {
  uint32_t ready, world;
  uint32_t bar_count, bar_gen, abort, spec_disabled;
  ShardRankSlot ranks[SHARD_MAX_WORLD];
  uint64_t lens[SHARD_MAX_WORLD][SHARD_MAX_BINS];      // [source rank][global bin]: keys source s holds for bin g this round
  uint64_t hist[SHARD_MAX_WORLD][2 * SHARD_HIST_MAX];  // a rank's (value, frequency) pairs
};

namespace {

double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

kmg_status shard_fail(kmg_ctx *c, kmg_status st, const std::string &msg) {
  if (c->shm) __atomic_store_n(&c->shm->abort, 1u, __ATOMIC_RELEASE);  // peers waiting in a barrier give up instead of hanging
  return fail(c, st, msg);
}

// sense-reversing barrier over the group's shared segment (processes or threads)
kmg_status shard_barrier(kmg_ctx *c) {
  ShardShm *S = c->shm;
  const uint32_t gen = __atomic_load_n(&S->bar_gen, __ATOMIC_ACQUIRE);
  if (__atomic_add_fetch(&S->bar_count, 1u, __ATOMIC_ACQ_REL) == S->world) {
    __atomic_store_n(&S->bar_count, 0u, __ATOMIC_RELAXED);
    __atomic_add_fetch(&S->bar_gen, 1u, __ATOMIC_RELEASE);
    return KMG_OK;
  }
  const double t0 = now_s();
  for (uint32_t spins = 0;; ++spins) {
    if (__atomic_load_n(&S->bar_gen, __ATOMIC_ACQUIRE) != gen) return KMG_OK;
    if (__atomic_load_n(&S->abort, __ATOMIC_ACQUIRE)) return fail(c, KMG_ERR_STATE, "a peer rank of the shard group failed");
    if (spins > 2000) {
      sched_yield();
      if ((spins & 1023) == 0 && now_s() - t0 > SHARD_TIMEOUT_S) return shard_fail(c, KMG_ERR_STATE, "shard group barrier timed out (a peer did not make the matching call)");
    }
  }
}
#define SB(c) do { kmg_status _s = shard_barrier(c); if (_s != KMG_OK) return _s; } while (0)
#define SCU(c, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cuda_fail(c, _e, #call); return shard_fail(c, _e == cudaErrorMemoryAllocation ? KMG_ERR_OOM : KMG_ERR_CUDA, c->err); } } while (0)

void shard_release(kmg_ctx *c) {
  for (uint32_t r = 0; r < SHARD_MAX_WORLD; ++r) {
    if (c->sh_peer[r] && c->sh_peer_ipc[r]) cudaIpcCloseMemHandle(c->sh_peer[r]);
    c->sh_peer[r] = nullptr; c->sh_peer_ipc[r] = false;
  }
  if (c->sh_recv) { cudaFree(c->sh_recv); c->sh_recv = nullptr; }
  if (c->shm) { munmap(c->shm, sizeof(ShardShm)); c->shm = nullptr; }
  c->sh_world = 1; c->sh_rank = 0;
}

// One round of the sharded pipeline over the context's packed stream (n_words_total may be 0: a rank that has run out of
// input still takes part).  A1 scatters this rank's keys into ITS OWN send buffer, one region per global bin (owner-major);
// after the barrier every owner runs A2 straight over the regions that belong to it in ALL ranks' send buffers -- the tile
// loads of the refine kernel are P2P loads, i.e. the NVLink transfer happens inside A2, tile by tile, prefetched one tile ahead.
// Send buffers are double buffered: round i+1 scatters into the other half while slower owners may still be pulling round i.
kmg_status shard_scan_round(kmg_ctx *c, uint64_t n_words_total, bool has_start) {
  ShardShm *S = c->shm;
  const uint32_t W = c->sh_world, me = c->sh_rank, P1 = c->n_coarse, G = W * P1;
  unsigned long long *d_cnt = c->d_part, *d_start = c->d_part + MAX_PARTS, *d_cur = c->d_part + 2 * MAX_PARTS;
  static const bool trace = getenv("KMG_SHARD_TIMING") != nullptr;  // per-round wall times (every phase ends in a sync anyway)
  const double tr0 = now_s();
  // 1. the round's size is known to all (and everybody has finished the round before the previous one: see the buffers' reuse)
  S->ranks[me].round_bytes = n_words_total * 32;
  S->ranks[me].round_flag = 0;
  SB(c);
  uint64_t max_w = 0;
  for (uint32_t r = 0; r < W; ++r) max_w = std::max<uint64_t>(max_w, S->ranks[r].round_bytes);
  if (max_w == 0) return KMG_OK;
  if (max_w >= (1ull << 32) || max_w > c->sh_recv_entries) return shard_fail(c, KMG_ERR_INVALID_ARG, "a sharded round handles at most batch_bases (< 2^32) bases per rank");
  const uint64_t parity = c->sh_rounds & 1;
  c->sh_rounds++;
  uint64_t *send = c->sh_recv + parity * c->sh_recv_entries;
  const uint64_t *src[SHARD_MAX_WORLD];
  for (uint32_t r = 0; r < W; ++r) src[r] = c->sh_peer[r] + parity * c->sh_recv_entries;
  ScanInput in;
  in.bases = c->d_bases; in.valid = c->d_valid; in.start = has_start ? c->d_start : nullptr;
  in.n_tiles = n_words_total / TILE_WORDS; in.k = c->k;
  std::vector<uint64_t> h_start(G), h_len(G);
  const size_t tmr = timer_begin(c, 0);
  bool done = false;
  double tr1 = tr0, tr2 = tr0;
  // 2. speculative layout: every global bin owns cap slots of the send buffer (cap from the largest rank's round, so that owners
  //    know where to read without being told); no count pass
  const double mu = (double)max_w / G;
  const uint64_t cap = ((uint64_t)(mu + 7.0 * std::sqrt(mu) + 64.0) + 15) & ~15ull;
  if (!__atomic_load_n(&S->spec_disabled, __ATOMIC_ACQUIRE) && cap * G <= c->sh_recv_entries && scan_scatter_supports_cap(G)) {
    for (uint32_t g = 0; g < G; ++g) h_start[g] = (uint64_t)g * cap;
    ScanInput sin = in;
    sin.part_cap = cap;
    sin.overflow_flag = reinterpret_cast<uint32_t *>(c->d_stats + 5);
    uint32_t h_flag = 0;
    SCU(c, cudaMemsetAsync(c->d_part, 0, 3 * (size_t)MAX_PARTS * sizeof(unsigned long long), c->stream));
    SCU(c, cudaMemsetAsync(sin.overflow_flag, 0, 4, c->stream));
    SCU(c, cudaMemcpyAsync(d_start, h_start.data(), G * 8, cudaMemcpyHostToDevice, c->stream));
    if (in.n_tiles) SCU(c, launch_scan_partition(sin, G, true, d_cnt, d_start, d_cur, send, c->d_counters, c->stream, /*mixed=*/true));
    SCU(c, cudaMemcpyAsync(h_len.data(), d_cur, G * 8, cudaMemcpyDeviceToHost, c->stream));
    SCU(c, cudaMemcpyAsync(&h_flag, sin.overflow_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    SCU(c, cudaStreamSynchronize(c->stream));
    tr1 = now_s();
    memcpy(S->lens[me], h_len.data(), G * 8);
    S->ranks[me].round_flag = h_flag;
    SB(c);  // every rank's regions are complete in its send buffer
    tr2 = now_s();
    bool any = false;
    for (uint32_t r = 0; r < W; ++r) any |= S->ranks[r].round_flag != 0;
    if (!any) {
      std::vector<uint64_t> off((size_t)P1 * W + 1), lens((size_t)P1 * W);
      uint64_t n_in = 0;
      for (uint32_t cl = 0; cl < P1; ++cl)
        for (uint32_t s2 = 0; s2 < W; ++s2) {
          const size_t j = (size_t)cl * W + s2;
          off[j] = (uint64_t)(me * P1 + cl) * cap; lens[j] = S->lens[s2][me * P1 + cl]; n_in += lens[j];
        }
      off[(size_t)P1 * W] = 0;  // unused: lengths are explicit
      for (uint32_t g = 0; g < G; ++g) if (g / P1 != me) c->sh_sent += h_len[g];
      c->sh_recv_keys += n_in;
      if (n_in) {
        kmg_status st = refine_to_run(c, nullptr, nullptr, off, /*owns=*/false, /*sync=*/true, &lens, nullptr, false, W, src);
        if (st != KMG_OK) return shard_fail(c, st, c->err);
      }
      done = true;
    } else {
      __atomic_store_n(&S->spec_disabled, 1u, __ATOMIC_RELEASE);  // sticky for the whole group: skewed input
    }
  }
  // 3. exact route (skewed input): count pass, exact prefix inside the own send buffer, scatter; owners derive every source's
  //    region offsets from the size matrix in the shared segment
  if (!done) {
    c->sh_exact_rounds++;
    SCU(c, cudaMemsetAsync(c->d_part, 0, 3 * (size_t)MAX_PARTS * sizeof(unsigned long long), c->stream));
    if (in.n_tiles) SCU(c, launch_scan_partition(in, G, false, d_cnt, d_start, d_cur, nullptr, c->d_counters, c->stream));
    SCU(c, cudaMemcpyAsync(h_len.data(), d_cnt, G * 8, cudaMemcpyDeviceToHost, c->stream));
    SCU(c, cudaStreamSynchronize(c->stream));
    uint64_t run = 0;
    for (uint32_t g = 0; g < G; ++g) { h_start[g] = run; run += h_len[g]; }
    if (run > c->sh_recv_entries) return shard_fail(c, KMG_ERR_STATE, "send buffer smaller than a round (internal sizing error)");
    SCU(c, cudaMemcpyAsync(d_start, h_start.data(), G * 8, cudaMemcpyHostToDevice, c->stream));
    if (in.n_tiles) SCU(c, launch_scan_partition(in, G, true, d_cnt, d_start, d_cur, send, c->d_counters, c->stream, /*mixed=*/true));
    SCU(c, cudaStreamSynchronize(c->stream));
    tr1 = now_s();
    memcpy(S->lens[me], h_len.data(), G * 8);
    SB(c);
    tr2 = now_s();
    std::vector<uint64_t> off((size_t)P1 * W + 1), lens((size_t)P1 * W);
    uint64_t n_in = 0;
    for (uint32_t s2 = 0; s2 < W; ++s2) {
      uint64_t pre = 0;
      for (uint32_t g = 0; g < G; ++g) {
        if (g / P1 == me) { const size_t j = (size_t)(g - me * P1) * W + s2; off[j] = pre; lens[j] = S->lens[s2][g]; n_in += lens[j]; }
        pre += S->lens[s2][g];
      }
    }
    off[(size_t)P1 * W] = 0;
    for (uint32_t g = 0; g < G; ++g) if (g / P1 != me) c->sh_sent += h_len[g];
    c->sh_recv_keys += n_in;
    if (n_in) {
      kmg_status st = refine_to_run(c, nullptr, nullptr, off, /*owns=*/false, /*sync=*/true, &lens, nullptr, false, W, src);
      if (st != KMG_OK) return shard_fail(c, st, c->err);
    }
  }
  timer_end(c, tmr);
  if (trace && me == 0)
    fprintf(stderr, "[shard round %llu] bases %llu  scatter %.2f ms  wait-for-peers %.2f ms  pull+refine %.2f ms  (G=%u, %s)\n",
            (unsigned long long)c->sh_rounds, (unsigned long long)(n_words_total * 32), (tr1 - tr0) * 1e3, (tr2 - tr1) * 1e3, (now_s() - tr2) * 1e3,
            G, done ? "speculative" : "exact");
  return KMG_OK;
}

}  // namespace

KMG_EXPORT kmg_status kmg_shard_join(kmg_ctx *c, uint32_t world, uint32_t rank, const char *group, uint64_t expected_keys_total) {
  if (!c || !group || !*group) return KMG_ERR_INVALID_ARG;
  if (world < 1 || world > SHARD_MAX_WORLD || rank >= world) return fail(c, KMG_ERR_INVALID_ARG, "world must be 1..=8 and rank < world");
  if (c->shm || c->sh_world > 1) return fail(c, KMG_ERR_STATE, "context already belongs to a shard group");
  if (c->mode != kmg_ctx::MODE_UNDECIDED) return fail(c, KMG_ERR_STATE, "kmg_shard_join must come before the first feeding call");
  if (world == 1) return KMG_OK;  // a group of one is the plain single-GPU path
  if (c->cfg.flags & (KMG_FLAG_FORCE_HASH | KMG_FLAG_FORCE_DIRECT)) return fail(c, KMG_ERR_INVALID_ARG, "the sharded path is the partitioned pipeline: FORCE_HASH / FORCE_DIRECT do not apply");
  CU(c, cudaSetDevice(c->device));
  // ---- partition plan: every rank derives the same one from the same arguments
  if (c->use_dense) { cudaFree(c->dense); c->dense = nullptr; c->dense_n = 0; c->use_dense = false; }
  c->cfg.flags |= KMG_FLAG_FORCE_PARTITIONED;
  c->mode = kmg_ctx::MODE_PARTITIONED;
  {
    const uint64_t per_rank = std::max<uint64_t>(expected_keys_total ? expected_keys_total : c->cfg.expected_distinct, 1) / world + 1;
    const uint64_t want = std::max<uint64_t>(4, (per_rank + TARGET_KEYS_PER_PART - 1) / TARGET_KEYS_PER_PART);
    uint64_t p1 = 1;
    while (p1 * p1 < want) ++p1;
    uint64_t g_max = SHARD_MAX_BINS;  // world * P1 global bins in one scatter (the single-scan rows kernel's limit)
    if (const char *e = getenv("KMG_SHARD_BINS")) g_max = std::min<uint64_t>(std::max<uint64_t>(strtoull(e, nullptr, 10), world), SHARD_MAX_BINS);
    p1 = std::min<uint64_t>(p1, g_max / world);
    c->n_coarse = (uint32_t)std::max<uint64_t>(p1, 1);
    c->n_sub = (uint32_t)std::min<uint64_t>((want + c->n_coarse - 1) / c->n_coarse, 2048);
    kmg_status st = init_partitioned(c);
    if (st != KMG_OK) return st;
  }
  // ---- send buffer (double buffered, one allocation so that one IPC handle covers both halves): a round scatters at most
  // batch_bases keys into it; the speculative layout needs mean + 7 sigma + 64 slots per global bin
  const uint64_t G = (uint64_t)world * c->n_coarse;
  const uint64_t round_max = round_up(c->batch_bases, (uint64_t)TILE_WORDS * 32);  // rounds are padded to whole tiles
  const double mu = (double)round_max / (double)G;
  c->sh_recv_entries = std::min<uint64_t>(round_max + G * (uint64_t)(7.0 * std::sqrt(mu) + 96.0), (1ull << 32) - 1);
  CU(c, cudaMalloc(&c->sh_recv, 2 * c->sh_recv_entries * 8));
  // ---- shared segment: rank 0 creates, the others attach
  const std::string name = std::string("/kmg-") + group;
  int fd = -1;
  const double t0 = now_s();
  if (rank == 0) {
    shm_unlink(name.c_str());
    fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)sizeof(ShardShm)) != 0) { if (fd >= 0) close(fd); shard_release(c); return fail(c, KMG_ERR_IO, "cannot create the shard group's shared segment " + name); }
  } else {
    while ((fd = shm_open(name.c_str(), O_RDWR, 0600)) < 0) {
      if (now_s() - t0 > SHARD_TIMEOUT_S) { shard_release(c); return fail(c, KMG_ERR_STATE, "shard group " + name + " did not appear (rank 0 missing?)"); }
      usleep(1000);
    }
    struct stat sb;
    while (fstat(fd, &sb) == 0 && (size_t)sb.st_size < sizeof(ShardShm)) {
      if (now_s() - t0 > SHARD_TIMEOUT_S) { close(fd); shard_release(c); return fail(c, KMG_ERR_STATE, "shard group segment was never sized"); }
      usleep(1000);
    }
  }
  void *mem = mmap(nullptr, sizeof(ShardShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (mem == MAP_FAILED) { shard_release(c); return fail(c, KMG_ERR_IO, "cannot map the shard group's shared segment"); }
  c->shm = static_cast<ShardShm *>(mem);
  c->sh_world = world; c->sh_rank = rank;
  ShardShm *S = c->shm;
  if (rank == 0) {  // ftruncate zero-filled the segment
    S->world = world;
    __atomic_store_n(&S->ready, SHARD_MAGIC, __ATOMIC_RELEASE);
  } else {
    while (__atomic_load_n(&S->ready, __ATOMIC_ACQUIRE) != SHARD_MAGIC) {
      if (now_s() - t0 > SHARD_TIMEOUT_S) { shard_release(c); return fail(c, KMG_ERR_STATE, "shard group segment was never initialised"); }
      usleep(200);
    }
    if (S->world != world) { shard_release(c); return fail(c, KMG_ERR_INVALID_ARG, "ranks disagree on the world size"); }
  }
  ShardRankSlot &mine = S->ranks[rank];
  mine.pid = (int32_t)getpid(); mine.device = c->device;
  mine.raw_ptr = (uint64_t)(uintptr_t)c->sh_recv; mine.recv_entries = c->sh_recv_entries;
  mine.summary[0] = c->n_coarse; mine.summary[1] = c->n_sub; mine.summary[2] = (uint64_t)c->k;
  cudaError_t e = cudaIpcGetMemHandle(&mine.handle, c->sh_recv);
  if (e != cudaSuccess) { cudaGetLastError(); memset(&mine.handle, 0, sizeof mine.handle); }  // same-process groups do not need it
  kmg_status st = shard_barrier(c);
  if (st != KMG_OK) { shard_release(c); return st; }
  std::string why;
  for (uint32_t r = 0; r < world && why.empty(); ++r) {
    const ShardRankSlot &o = S->ranks[r];
    if (o.summary[0] != c->n_coarse || o.summary[1] != c->n_sub || o.summary[2] != (uint64_t)c->k || o.recv_entries != c->sh_recv_entries) {
      why = "ranks disagree on k / the partition plan / batch_bases (pass identical arguments on every rank)";
    } else if (r == rank) {
      c->sh_peer[r] = c->sh_recv;
    } else if (o.pid == mine.pid) {  // threads of one process: plain pointers (+ peer access between different devices)
      if (o.device != c->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, c->device, o.device);
        if (!can) { why = "no peer access between devices " + std::to_string(c->device) + " and " + std::to_string(o.device); break; }
        cudaError_t pe = cudaDeviceEnablePeerAccess(o.device, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) why = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe);
        cudaGetLastError();
      }
      c->sh_peer[r] = reinterpret_cast<uint64_t *>((uintptr_t)o.raw_ptr);
    } else {
      void *pp = nullptr;
      cudaError_t pe = cudaIpcOpenMemHandle(&pp, o.handle, cudaIpcMemLazyEnablePeerAccess);
      if (pe != cudaSuccess) { cudaGetLastError(); why = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(pe); }
      else { c->sh_peer[r] = static_cast<uint64_t *>(pp); c->sh_peer_ipc[r] = true; }
    }
  }
  if (!why.empty()) __atomic_store_n(&S->abort, 1u, __ATOMIC_RELEASE);
  st = shard_barrier(c);  // every rank has opened its peers (or the group is aborted)
  if (rank == 0) shm_unlink(name.c_str());  // the mapping lives on; no name is left behind
  if (!why.empty() || st != KMG_OK) { shard_release(c); c->mode = kmg_ctx::MODE_PARTITIONED; return why.empty() ? st : fail(c, KMG_ERR_CUDA, why); }
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_shard_leave(kmg_ctx *c) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (!c->shm) return KMG_OK;
  CU(c, cudaSetDevice(c->device));
  cudaStreamSynchronize(c->stream);
  shard_barrier(c);  // nobody may still be writing into a buffer that is about to be unmapped / freed
  shard_release(c);
  return KMG_OK;
}

// Collective: every rank of the group calls it the same number of times (a rank without input passes n_bytes = 0).
// Input: THIS rank's slice of the records (device pointers), layout as kmg_count_ascii_device.
KMG_EXPORT kmg_status kmg_shard_count_ascii_device(kmg_ctx *c, const uint8_t *d_seq, const uint8_t *d_qual, const uint64_t *d_offsets,
                                                   uint64_t n_records, uint64_t n_bytes) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (!c->shm) return kmg_count_ascii_device(c, d_seq, d_qual, d_offsets, n_records, n_bytes);
  if (n_bytes && !d_seq) return shard_fail(c, KMG_ERR_INVALID_ARG, "d_seq is NULL");
  if (((uintptr_t)d_seq & 15) || ((uintptr_t)d_qual & 15)) return shard_fail(c, KMG_ERR_INVALID_ARG, "d_seq / d_qual must be 16-byte aligned");
  CU(c, cudaSetDevice(c->device));
  // rounds of at most batch_bases bases (k-1 overlap, records may span rounds); the ranks agree on the number of rounds
  // (round starts stay 16-byte aligned for the ingest kernel's vector loads: step is rounded down, a round covers step + k-1 bases)
  const uint64_t K1 = (uint64_t)c->k - 1, B = c->batch_bases, step = (B - K1) & ~15ull;
  if (step == 0) return shard_fail(c, KMG_ERR_INVALID_ARG, "batch_bases is too small for sharded rounds");
  uint64_t my_rounds = 0;
  for (uint64_t pos = 0; pos < n_bytes; pos += step) { if (pos && n_bytes - pos <= K1) break; ++my_rounds; }
  ShardShm *S = c->shm;
  S->ranks[c->sh_rank].summary[7] = my_rounds;
  SB(c);
  uint64_t rounds = 0;
  for (uint32_t r = 0; r < c->sh_world; ++r) rounds = std::max<uint64_t>(rounds, S->ranks[r].summary[7]);
  SB(c);
  const bool use_q = c->cfg.has_min_quality && d_qual != nullptr;
  const uint32_t thr = std::min<uint32_t>(255u, (uint32_t)c->cfg.min_quality + 33u);
  for (uint64_t i = 0; i < rounds; ++i) {
    const uint64_t pos = i * step;
    uint64_t len = 0;
    if (i < my_rounds) len = std::min<uint64_t>(step + K1, n_bytes - pos);
    uint64_t n_words_total = 0;
    bool has_start = false;
    if (len) {
      n_words_total = round_up((len + 31) / 32, TILE_WORDS);
      kmg_status st = ensure_packed(c, n_words_total);
      if (st != KMG_OK) return shard_fail(c, st, c->err);
      SCU(c, launch_ingest(d_seq + pos, use_q ? d_qual + pos : nullptr, len, thr, n_words_total, c->d_bases, c->d_valid, c->stream));
      has_start = d_offsets != nullptr && n_records > 1;
      if (has_start) SCU(c, launch_start_bits(d_offsets, n_records, pos, len, n_words_total, c->d_start, c->stream));
    }
    kmg_status st = shard_scan_round(c, n_words_total, has_start);
    if (st != KMG_OK) return st;
  }
  c->n_records += n_records; c->n_bases += n_bytes;
  return KMG_OK;
}

// The same with HOST pointers (this rank's slice): staged through the context's pinned ring like kmg_count_ascii.
KMG_EXPORT kmg_status kmg_shard_count_ascii(kmg_ctx *c, const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets, uint64_t n_records) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (!c->shm) return kmg_count_ascii(c, seq, qual, offsets, n_records);
  if (n_records && (!offsets || !seq)) return shard_fail(c, KMG_ERR_INVALID_ARG, "seq / offsets must not be NULL");
  CU(c, cudaSetDevice(c->device));
  const uint64_t begin = n_records ? offsets[0] : 0, end = n_records ? offsets[n_records] : 0;
  const bool use_q = c->cfg.has_min_quality && qual != nullptr;
  const bool src_pinned = end > begin && is_pinned_host(seq + begin) && (!use_q || is_pinned_host(qual + begin));
  kmg_status s = ensure_staging(c, !src_pinned, use_q, std::max<uint64_t>(end - begin, 1));
  if (s != KMG_OK) return shard_fail(c, s, c->err);
  const uint64_t K1 = (uint64_t)c->k - 1, B = c->batch_bases, step = B - K1;
  struct Chunk { uint64_t pos, len, r_lo, nrec; };
  std::vector<Chunk> chunks;
  {
    uint64_t r_lo = 0;
    for (uint64_t pos = begin; pos < end; pos += step) {
      const uint64_t len = std::min<uint64_t>(B, end - pos);
      if (pos != begin && len <= K1) break;
      while (r_lo < n_records && offsets[r_lo] <= pos) ++r_lo;
      uint64_t r_hi = r_lo;
      while (r_hi < n_records && offsets[r_hi] < pos + len) ++r_hi;
      chunks.push_back(Chunk{pos, len, r_lo, r_hi - r_lo});
    }
  }
  ShardShm *S = c->shm;
  S->ranks[c->sh_rank].summary[7] = chunks.size();
  SB(c);
  uint64_t rounds = 0;
  for (uint32_t r = 0; r < c->sh_world; ++r) rounds = std::max<uint64_t>(rounds, S->ranks[r].summary[7]);
  SB(c);
  auto stage_chunk = [&](const Chunk &ch, Staging &st) -> kmg_status {
    if (st.h2d_pending) { CU(c, cudaEventSynchronize(st.h2d_done)); st.h2d_pending = false; }
    if (st.compute_pending) { CU(c, cudaStreamWaitEvent(c->copy_stream, st.compute_done, 0)); }
    kmg_status es = ensure_offsets(c, st, ch.nrec);
    if (es != KMG_OK) return es;
    const uint8_t *src_seq = seq + ch.pos, *src_qual = use_q ? qual + ch.pos : nullptr;
    if (!src_pinned) {
      memcpy(st.h_seq, src_seq, ch.len); src_seq = st.h_seq;
      if (use_q) { memcpy(st.h_qual, src_qual, ch.len); src_qual = st.h_qual; }
    }
    CU(c, cudaMemcpyAsync(st.d_seq, src_seq, ch.len, cudaMemcpyHostToDevice, c->copy_stream));
    if (use_q) CU(c, cudaMemcpyAsync(st.d_qual, src_qual, ch.len, cudaMemcpyHostToDevice, c->copy_stream));
    if (ch.nrec) {
      memcpy(st.h_off, offsets + ch.r_lo, ch.nrec * 8);
      CU(c, cudaMemcpyAsync(st.d_off, st.h_off, ch.nrec * 8, cudaMemcpyHostToDevice, c->copy_stream));
    }
    c->h2d_bytes += ch.len * (use_q ? 2 : 1) + ch.nrec * 8;
    CU(c, cudaEventRecord(st.h2d_done, c->copy_stream));
    st.h2d_pending = true;
    return KMG_OK;
  };
  const uint32_t thr = std::min<uint32_t>(255u, (uint32_t)c->cfg.min_quality + 33u);
  const uint32_t slot0 = c->next_ascii;
  size_t staged = 0;
  for (uint64_t i = 0; i < rounds; ++i) {
    uint64_t n_words_total = 0;
    bool has_start = false;
    if (i < chunks.size()) {
      for (; staged < chunks.size() && staged < i + N_STAGE; ++staged)
        if ((s = stage_chunk(chunks[staged], c->st[(slot0 + staged) % N_STAGE])) != KMG_OK) return shard_fail(c, s, c->err);
      Staging &st = c->st[(slot0 + i) % N_STAGE];
      const Chunk &ch = chunks[i];
      SCU(c, cudaStreamWaitEvent(c->stream, st.h2d_done, 0));
      n_words_total = round_up((ch.len + 31) / 32, TILE_WORDS);
      if ((s = ensure_packed(c, n_words_total)) != KMG_OK) return shard_fail(c, s, c->err);
      SCU(c, launch_ingest(st.d_seq, use_q ? st.d_qual : nullptr, ch.len, thr, n_words_total, c->d_bases, c->d_valid, c->stream));
      has_start = ch.nrec > 0;
      if (has_start) SCU(c, launch_start_bits(st.d_off, ch.nrec, ch.pos, ch.len, n_words_total, c->d_start, c->stream));
      SCU(c, cudaEventRecord(st.compute_done, c->stream));  // the staging slot is free once ingest has packed it
      st.compute_pending = true;
    }
    s = shard_scan_round(c, n_words_total, has_start);
    if (s != KMG_OK) return s;
  }
  c->next_ascii = (slot0 + (uint32_t)chunks.size()) % N_STAGE;
  c->n_records += n_records; c->n_bases += end - begin;
  if (src_pinned)
    for (auto &st : c->st)
      if (st.h2d_pending) { CU(c, cudaEventSynchronize(st.h2d_done)); st.h2d_pending = false; }
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_finalize(kmg_ctx *c, kmg_summary *out) {
  if (!c) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  { kmg_status _fs = scan_pending_batch(c); if (_fs != KMG_OK) return _fs; }
  CU(c, cudaStreamSynchronize(c->copy_stream));
  kmg_status s = read_counters(c);
  if (s != KMG_OK) return s;
  for (auto &st : c->st) { st.h2d_pending = false; st.compute_pending = false; }
  if (c->mode == kmg_ctx::MODE_PARTITIONED && (s = consolidate(c)) != KMG_OK) return s;
  timers_collect(c);
  if (c->mode == kmg_ctx::MODE_TABLE) c->distinct_ub = c->h_counters[CTR_DISTINCT];
  if (!out) return KMG_OK;
  memset(out, 0, sizeof *out);
  unsigned long long h[3];
  if (fetch_fused_hist(c)) { h[0] = c->result.n_valid; h[1] = c->fused_max; h[2] = c->fused_sum; }
  else {
    CU(c, launch_table_stats(view_of(c), 1, c->d_stats, c->stream));
    CU(c, cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  out->n_records = c->n_records; out->n_bases = c->n_bases;
  out->n_windows = h[2];
  out->n_distinct = h[0]; out->max_count = h[1];
  out->table_capacity = c->use_dense ? c->dense_n : c->mode == kmg_ctx::MODE_PARTITIONED ? c->n_parts : c->table.cap;
  out->path = c->use_dense ? 1 : c->mode == kmg_ctx::MODE_PARTITIONED ? 2 : 0;
  out->n_grows = c->mode == kmg_ctx::MODE_PARTITIONED ? c->n_consolidations : c->n_grows;
  out->kernel_ns = (uint64_t)(c->kernel_ms * 1e6);
  out->scan_ns = (uint64_t)(c->cat_ms[0] * 1e6);
  out->consolidate_ns = (uint64_t)(c->cat_ms[1] * 1e6);
  out->h2d_bytes = c->h2d_bytes;
  return KMG_OK;
}

namespace {
kmg_status count_filtered(kmg_ctx *c, uint64_t min_count, uint64_t *n, uint64_t shard_mod = 0, uint64_t shard_rem = 0) {
  { kmg_status _fs = scan_pending_batch(c); if (_fs != KMG_OK) return _fs; }
  CU(c, cudaStreamSynchronize(c->copy_stream));
  if (c->mode == kmg_ctx::MODE_PARTITIONED) { kmg_status s = consolidate(c); if (s != KMG_OK) return s; }
  if (shard_mod <= 1 && fetch_fused_hist(c)) {
    uint64_t m = 0;
    for (uint64_t b = std::max<uint64_t>(min_count, 1); b < (uint64_t)HIST_DENSE_BINS; ++b) m += c->h_bins[b];
    m += c->h_ov.end() - std::lower_bound(c->h_ov.begin(), c->h_ov.end(), min_count);
    *n = m;
    return KMG_OK;
  }
  TableView v = view_of(c);
  v.shard_mod = shard_mod; v.shard_rem = shard_rem;
  CU(c, launch_table_stats(v, min_count, c->d_stats, c->stream));
  unsigned long long h[3];
  CU(c, cudaMemcpyAsync(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *n = h[0];
  return KMG_OK;
}

kmg_status export_device(kmg_ctx *c, uint64_t min_count, int sorted, uint64_t shard_mod, uint64_t shard_rem, uint64_t *d_keys,
                         uint64_t *d_counts, uint64_t cap, uint64_t *n_out) {
  if (!c || !n_out) return KMG_ERR_INVALID_ARG;
  if (shard_mod > 1 && shard_rem >= shard_mod) return fail(c, KMG_ERR_INVALID_ARG, "shard must be < n_shards");
  CU(c, cudaSetDevice(c->device));
  uint64_t n = 0;
  kmg_status s = count_filtered(c, min_count, &n, shard_mod, shard_rem);
  if (s != KMG_OK) return s;
  *n_out = n;
  if (!d_keys || !d_counts) return KMG_OK;
  if (cap < n) return fail(c, KMG_ERR_CAPACITY, "output arrays hold " + std::to_string(cap) + " entries, need " + std::to_string(n));
  TableView v = view_of(c);
  v.shard_mod = shard_mod; v.shard_rem = shard_rem;
  CU(c, launch_compact(v, min_count, d_keys, d_counts, cap, c->d_stats + 3, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (sorted) CU(c, sort_pairs(d_keys, d_counts, n, 2 * c->k, c->stream));
  return KMG_OK;
}

kmg_status export_host(kmg_ctx *c, uint64_t min_count, int sorted, uint64_t shard_mod, uint64_t shard_rem, uint64_t *keys,
                       uint64_t *counts, uint64_t cap, uint64_t *n_out) {
  if (!c || !n_out) return KMG_ERR_INVALID_ARG;
  if (shard_mod > 1 && shard_rem >= shard_mod) return fail(c, KMG_ERR_INVALID_ARG, "shard must be < n_shards");
  CU(c, cudaSetDevice(c->device));
  uint64_t n = 0;
  kmg_status s = count_filtered(c, min_count, &n, shard_mod, shard_rem);
  if (s != KMG_OK) return s;
  *n_out = n;
  if (!keys || !counts) return KMG_OK;
  if (cap < n) return fail(c, KMG_ERR_CAPACITY, "output arrays hold " + std::to_string(cap) + " entries, need " + std::to_string(n));
  if (n == 0) return KMG_OK;
  uint64_t *dk = nullptr, *dc = nullptr;
  CU(c, pool_alloc(c, &dk, n * 8));
  cudaError_t e = pool_alloc(c, &dc, n * 8);
  if (e != cudaSuccess) { pool_free(c, dk); return cuda_fail(c, e, "cudaMalloc(export)"); }
  uint64_t n2 = 0;
  s = export_device(c, min_count, sorted, shard_mod, shard_rem, dk, dc, n, &n2);
  if (s == KMG_OK) {
    e = cudaMemcpyAsync(keys, dk, n * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, dc, n * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) s = cuda_fail(c, e, "D2H export");
  }
  pool_free(c, dk); pool_free(c, dc);
  return s;
}
}  // namespace

KMG_EXPORT kmg_status kmg_export_counts_device(kmg_ctx *c, uint64_t min_count, int sorted, uint64_t *d_keys, uint64_t *d_counts,
                                               uint64_t cap, uint64_t *n_out) {
  return export_device(c, min_count, sorted, 0, 0, d_keys, d_counts, cap, n_out);
}

KMG_EXPORT kmg_status kmg_export_counts(kmg_ctx *c, uint64_t min_count, int sorted, uint64_t *keys, uint64_t *counts, uint64_t cap,
                                        uint64_t *n_out) {
  return export_host(c, min_count, sorted, 0, 0, keys, counts, cap, n_out);
}

KMG_EXPORT kmg_status kmg_export_shard(kmg_ctx *c, uint64_t min_count, int sorted, uint64_t n_shards, uint64_t shard, uint64_t *keys,
                                       uint64_t *counts, uint64_t cap, uint64_t *n_out) {
  return export_host(c, min_count, sorted, n_shards, shard, keys, counts, cap, n_out);
}

KMG_EXPORT kmg_status kmg_export_shard_device(kmg_ctx *c, uint64_t min_count, int sorted, uint64_t n_shards, uint64_t shard,
                                              uint64_t *d_keys, uint64_t *d_counts, uint64_t cap, uint64_t *n_out) {
  return export_device(c, min_count, sorted, n_shards, shard, d_keys, d_counts, cap, n_out);
}

KMG_EXPORT kmg_status kmg_histogram(kmg_ctx *c, uint64_t min_count, uint64_t *count_vals, uint64_t *freqs, uint64_t cap, uint64_t *n_out) {
  if (!c || !n_out) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  { kmg_status _fs = scan_pending_batch(c); if (_fs != KMG_OK) return _fs; }
  CU(c, cudaStreamSynchronize(c->copy_stream));
  kmg_status s = read_counters(c);
  if (s != KMG_OK) return s;
  if (c->mode == kmg_ctx::MODE_PARTITIONED && (s = consolidate(c)) != KMG_OK) return s;
  if (fetch_fused_hist(c)) {  // phase B already built it
    const uint64_t lo = std::max<uint64_t>(min_count, 1);
    const auto ov0 = std::lower_bound(c->h_ov.begin(), c->h_ov.end(), min_count);
    uint64_t n = 0;
    for (uint64_t b = lo; b < (uint64_t)HIST_DENSE_BINS; ++b) if (c->h_bins[b]) ++n;
    for (auto it = ov0; it != c->h_ov.end(); ++it) if (it == ov0 || *it != *(it - 1)) ++n;
    *n_out = n;
    if (!count_vals || !freqs) return KMG_OK;
    if (cap < n) return fail(c, KMG_ERR_CAPACITY, "histogram arrays hold " + std::to_string(cap) + " bins, need " + std::to_string(n));
    uint64_t o = 0;
    for (uint64_t b = lo; b < (uint64_t)HIST_DENSE_BINS; ++b) if (c->h_bins[b]) { count_vals[o] = b; freqs[o] = c->h_bins[b]; ++o; }
    for (auto it = ov0; it != c->h_ov.end();) {
      auto j = it;
      while (j != c->h_ov.end() && *j == *it) ++j;
      count_vals[o] = *it; freqs[o] = (uint64_t)(j - it); ++o;
      it = j;
    }
    return KMG_OK;
  }
  // counts >= HIST_DENSE_BINS: at most (sum of counts) / HIST_DENSE_BINS distinct keys can reach that
  CU(c, launch_table_stats(view_of(c), 1, c->d_stats, c->stream));
  unsigned long long hs[3];
  CU(c, cudaMemcpyAsync(hs, c->d_stats, sizeof hs, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  const uint64_t ov_cap = hs[2] / HIST_DENSE_BINS + 16;
  unsigned long long *d_bins = nullptr;
  uint64_t *d_ov = nullptr;
  CU(c, pool_alloc(c, &d_bins, HIST_DENSE_BINS * 8));
  cudaError_t e = pool_alloc(c, &d_ov, ov_cap * 8);
  if (e != cudaSuccess) { pool_free(c, d_bins); return cuda_fail(c, e, "cudaMalloc(histogram overflow)"); }
  std::vector<unsigned long long> bins(HIST_DENSE_BINS);
  std::vector<uint64_t> ov;
  unsigned long long ov_n = 0;
  e = launch_histogram(view_of(c), min_count, d_bins, d_ov, ov_cap, c->d_stats + 4, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bins.data(), d_bins, HIST_DENSE_BINS * 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&ov_n, c->d_stats + 4, 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess && ov_n > ov_cap) { pool_free(c, d_bins); pool_free(c, d_ov); return fail(c, KMG_ERR_STATE, "histogram overflow list exceeded its bound"); }
  if (e == cudaSuccess && ov_n) { ov.resize(ov_n); e = cudaMemcpy(ov.data(), d_ov, ov_n * 8, cudaMemcpyDeviceToHost); }
  pool_free(c, d_bins); pool_free(c, d_ov);
  if (e != cudaSuccess) return cuda_fail(c, e, "histogram");
  // assemble ascending (count, frequency) pairs: dense bins first, then the (tiny) overflow tail
  std::sort(ov.begin(), ov.end());
  uint64_t n = 0;
  for (uint64_t b = 0; b < HIST_DENSE_BINS; ++b) if (bins[b]) ++n;
  for (size_t i = 0; i < ov.size(); ++i) if (i == 0 || ov[i] != ov[i - 1]) ++n;
  *n_out = n;
  if (!count_vals || !freqs) return KMG_OK;
  if (cap < n) return fail(c, KMG_ERR_CAPACITY, "histogram arrays hold " + std::to_string(cap) + " bins, need " + std::to_string(n));
  uint64_t o = 0;
  for (uint64_t b = 0; b < HIST_DENSE_BINS; ++b) if (bins[b]) { count_vals[o] = b; freqs[o] = bins[b]; ++o; }
  for (size_t i = 0; i < ov.size();) {
    size_t j = i;
    while (j < ov.size() && ov[j] == ov[i]) ++j;
    count_vals[o] = ov[i]; freqs[o] = j - i; ++o;
    i = j;
  }
  return KMG_OK;
}

namespace {
// CRC-32 of a concatenation from the parts' CRCs (zlib's crc32_combine idea: advance crc1 over len2 zero bytes with
// GF(2) matrix squaring, then xor crc2) -- lets every GPU shard checksum its own records (src/index.rs:434-456 streams one CRC).
uint32_t gf2_times(const uint32_t *mat, uint32_t vec) {
  uint32_t sum = 0;
  for (; vec; vec >>= 1, ++mat) if (vec & 1) sum ^= *mat;
  return sum;
}
void gf2_square(uint32_t *sq, const uint32_t *mat) { for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]); }
uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;  // operator for one zero bit
  for (uint32_t n = 1, row = 1; n < 32; ++n, row <<= 1) odd[n] = row;
  gf2_square(even, odd);  // two zero bits
  gf2_square(odd, even);  // four zero bits
  do {
    gf2_square(even, odd);
    if (len2 & 1) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}

// CRC-32 of a large buffer on all host cores: the parts' CRCs are combined (a single core manages ~2 GB/s, an index is tens of GB)
uint32_t crc32_parallel(const uint8_t *p, size_t n) {
  const Crc32 &T = crc_tables();
  unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  if (n < (8u << 20) || nt == 1) return ~T.update(~0u, p, n);
  const size_t per = (n + nt - 1) / nt;
  std::vector<uint32_t> part(nt, 0);
  std::vector<size_t> len(nt, 0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    const size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
    len[t] = hi - lo;
    th.emplace_back([&, t, lo, hi] { part[t] = ~T.update(~0u, p + lo, hi - lo); });
  }
  for (auto &x : th) x.join();
  uint32_t crc = part[0];
  for (unsigned t = 1; t < nt; ++t) crc = crc32_combine(crc, part[t], len[t]);
  return crc;
}

// Walk the table in ascending key order WITHOUT a second copy of it: the key space is cut into ranges of at most ~piece_target
// entries (from a histogram over the top key bits); each range is compacted (count >= min_count), sorted, and handed to
// fn(d_keys, d_counts, m).  The pieces concatenate to the sorted result.
template <class F>
kmg_status for_each_sorted_piece(kmg_ctx *c, uint64_t min_count, uint64_t piece_target, F &&fn) {
  uint64_t n_all = 0;
  kmg_status s = count_filtered(c, 0, &n_all);
  if (s != KMG_OK || n_all == 0) return s;
  const int shift = 2 * c->k > 12 ? 2 * c->k - 12 : 0;
  unsigned long long *d_b = nullptr;
  CU(c, pool_alloc(c, &d_b, KEY_BUCKETS * 8));
  std::vector<unsigned long long> hb(KEY_BUCKETS);
  cudaError_t e = launch_key_buckets(view_of(c), shift, d_b, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hb.data(), d_b, KEY_BUCKETS * 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  pool_free(c, d_b);
  if (e != cudaSuccess) return cuda_fail(c, e, "key histogram");
  uint64_t piece_cap = std::min<uint64_t>(n_all, piece_target);
  for (auto v : hb) piece_cap = std::max<uint64_t>(piece_cap, v);
  uint64_t *dk = nullptr, *dc = nullptr;
  CU(c, pool_alloc(c, &dk, piece_cap * 8));
  e = pool_alloc(c, &dc, piece_cap * 8);
  if (e != cudaSuccess) { pool_free(c, dk); return cuda_fail(c, e, "cudaMalloc(sorted piece)"); }
  for (int b0 = 0; s == KMG_OK && b0 < KEY_BUCKETS;) {
    uint64_t ub = hb[b0];
    int b1 = b0 + 1;
    while (b1 < KEY_BUCKETS && ub + hb[b1] <= piece_cap) ub += hb[b1++];
    if (ub) {
      TableView v = view_of(c);
      v.range_lo = (uint64_t)b0 << shift;
      v.range_hi = b1 == KEY_BUCKETS ? ~0ull : ((uint64_t)b1 << shift) - 1;
      unsigned long long m = 0;
      e = launch_compact(v, min_count, dk, dc, piece_cap, c->d_stats + 3, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(&m, c->d_stats + 3, 8, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e == cudaSuccess && m > piece_cap) { s = fail(c, KMG_ERR_STATE, "sorted piece larger than its key histogram"); break; }
      if (e == cudaSuccess && m) e = sort_pairs(dk, dc, m, 2 * c->k, c->stream);
      if (e != cudaSuccess) { s = cuda_fail(c, e, "sorted piece"); break; }
      if (m) s = fn(dk, dc, (uint64_t)m);
    }
    b0 = b1;
  }
  pool_free(c, dk); pool_free(c, dc);
  return s;
}

// Device bytes -> sink through two pinned staging buffers (the copy of chunk i+1 overlaps the sink of chunk i).
struct HostStager {
  kmg_ctx *c;
  uint8_t *h[2] = {nullptr, nullptr};
  size_t cap = 0;
  explicit HostStager(kmg_ctx *c_, size_t cap_) : c(c_), cap(cap_) {}
  ~HostStager() { for (auto *p : h) if (p) cudaFreeHost(p); }
  template <class Sink>
  kmg_status drain(const uint8_t *d_src, size_t n, Sink &&sink) {
    for (int i = 0; i < 2; ++i)
      if (!h[i] && cudaHostAlloc(&h[i], cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(c, KMG_ERR_OOM, "pinned staging buffer"); }
    size_t issued = 0;
    int slot = 0;
    size_t len_prev = 0;
    if (n) { len_prev = std::min(cap, n); CU(c, cudaMemcpyAsync(h[0], d_src, len_prev, cudaMemcpyDeviceToHost, c->stream)); issued = len_prev; }
    while (len_prev) {
      CU(c, cudaStreamSynchronize(c->stream));
      const size_t len_next = std::min(cap, n - issued);
      if (len_next) { CU(c, cudaMemcpyAsync(h[slot ^ 1], d_src + issued, len_next, cudaMemcpyDeviceToHost, c->stream)); issued += len_next; }
      kmg_status s = sink(h[slot], len_prev);
      if (s != KMG_OK) { cudaStreamSynchronize(c->stream); return s; }
      slot ^= 1; len_prev = len_next;
    }
    return KMG_OK;
  }
};

// Stream all (key, count) records of the context to `f` in ascending key order (16-byte records interleaved on the device).
// *crc_out is the finished CRC-32 of the bytes written here (standard init / xor-out).
kmg_status write_kmix_records(kmg_ctx *c, FILE *f, uint32_t *crc_out, uint64_t *n_out) {
  uint64_t n = 0;
  kmg_status s = count_filtered(c, 0, &n);
  if (s != KMG_OK) return s;
  *n_out = n;
  uint32_t crc = 0;  // CRC of the empty string
  uint64_t written = 0;
  bool first = true;
  HostStager stage(c, 64u << 20);
  void *d_pairs = nullptr;
  uint64_t pairs_cap = 0;
  s = for_each_sorted_piece(c, 0, 1ull << 25, [&](const uint64_t *dk, const uint64_t *dc, uint64_t m) -> kmg_status {
    if (m > pairs_cap) {
      pool_free(c, d_pairs); d_pairs = nullptr;
      pairs_cap = std::max<uint64_t>(m, std::min<uint64_t>(n, 1ull << 25));
      CU(c, pool_alloc(c, &d_pairs, pairs_cap * 16));
    }
    CU(c, launch_interleave_pairs(dk, dc, m, d_pairs, c->stream));
    written += m;
    return stage.drain(static_cast<const uint8_t *>(d_pairs), m * 16, [&](const uint8_t *p, size_t len) -> kmg_status {
      if (fwrite(p, 1, len, f) != len) return fail(c, KMG_ERR_IO, "short write to the index file");
      const uint32_t part = crc32_parallel(p, len);
      crc = first ? part : crc32_combine(crc, part, len);
      first = false;
      return KMG_OK;
    });
  });
  pool_free(c, d_pairs);
  if (s != KMG_OK) return s;
  if (written != n) return fail(c, KMG_ERR_STATE, "kmix writer: pieces do not add up to the table");
  *crc_out = crc;
  return KMG_OK;
}

bool has_gz_suffix(const char *path) { const size_t l = strlen(path); return l >= 3 && strcmp(path + l - 3, ".gz") == 0; }
}  // namespace

KMG_EXPORT kmg_status kmg_save_kmix(kmg_ctx *c, const char *path) {
  if (!c || !path) return KMG_ERR_INVALID_ARG;
  // src/index.rs:159-168 gzips when the path ends in .gz; this writer does not compress -- refuse rather than write raw bytes
  // under a .gz name (the reference's load_index would reject that file)
  if (has_gz_suffix(path)) return fail(c, KMG_ERR_INVALID_ARG, "kmg_save_kmix writes uncompressed .kmix files; gzip on the host side for a .gz path");
  CU(c, cudaSetDevice(c->device));
  uint64_t n = 0;
  kmg_status s = count_filtered(c, 0, &n);
  if (s != KMG_OK) return s;
  FILE *f = fopen(path, "wb");
  if (!f) return fail(c, KMG_ERR_IO, std::string("cannot open ") + path + " for writing");
  const Crc32 &T = crc_tables();
  uint8_t hdr[14] = {'K', 'M', 'I', 'X', 1, (uint8_t)c->k};
  for (int b = 0; b < 8; ++b) hdr[6 + b] = (uint8_t)(n >> (8 * b));
  bool ok = fwrite(hdr, 1, 14, f) == 14;
  uint32_t crc = ~T.update(~0u, hdr, 14), body_crc = 0;
  uint64_t n2 = 0;
  s = write_kmix_records(c, f, &body_crc, &n2);
  if (s != KMG_OK) { fclose(f); return s; }
  crc = crc32_combine(crc, body_crc, n2 * 16);
  uint8_t tail[4] = {(uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24)};
  ok = ok && n2 == n && fwrite(tail, 1, 4, f) == 4;
  ok = (fclose(f) == 0) && ok;
  if (!ok) return fail(c, KMG_ERR_IO, std::string("short write to ") + path);
  return KMG_OK;
}

// .kmix from several shards (one context per GPU / process; src/index.rs:222-279 writes header, n records in arbitrary order,
// CRC).  Every shard writes its records into ITS byte range of the same file -- record_offset = number of records of the shards
// before it -- and reports their count and CRC; kmg_kmix_finish then writes the header (n = sum) and the trailing CRC combined
// from the shards' CRCs.  The file must exist (kmg_kmix_begin creates / truncates it).
KMG_EXPORT kmg_status kmg_kmix_begin(const char *path) {
  if (!path || has_gz_suffix(path)) return KMG_ERR_INVALID_ARG;
  FILE *f = fopen(path, "wb");
  if (!f) return KMG_ERR_IO;
  return fclose(f) == 0 ? KMG_OK : KMG_ERR_IO;
}
KMG_EXPORT kmg_status kmg_save_kmix_shard(kmg_ctx *c, const char *path, uint64_t record_offset, uint64_t *n_records_out, uint32_t *crc_out) {
  if (!c || !path || !n_records_out || !crc_out) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  FILE *f = fopen(path, "r+b");
  if (!f) return fail(c, KMG_ERR_IO, std::string("cannot open ") + path + " (call kmg_kmix_begin first)");
  if (fseeko(f, (off_t)(14 + 16 * record_offset), SEEK_SET) != 0) { fclose(f); return fail(c, KMG_ERR_IO, "seek failed"); }
  kmg_status s = write_kmix_records(c, f, crc_out, n_records_out);
  const bool ok = fclose(f) == 0;
  if (s != KMG_OK) return s;
  return ok ? KMG_OK : fail(c, KMG_ERR_IO, std::string("short write to ") + path);
}
KMG_EXPORT kmg_status kmg_kmix_finish(const char *path, uint32_t k, const uint64_t *shard_records, const uint32_t *shard_crcs, uint32_t n_shards) {
  if (!path || k < 1 || k > 32 || (n_shards && (!shard_records || !shard_crcs))) return KMG_ERR_INVALID_ARG;
  uint64_t n = 0;
  for (uint32_t i = 0; i < n_shards; ++i) n += shard_records[i];
  FILE *f = fopen(path, "r+b");
  if (!f) return KMG_ERR_IO;
  const Crc32 &T = crc_tables();
  uint8_t hdr[14] = {'K', 'M', 'I', 'X', 1, (uint8_t)k};
  for (int b = 0; b < 8; ++b) hdr[6 + b] = (uint8_t)(n >> (8 * b));
  uint32_t crc = ~T.update(~0u, hdr, 14);
  for (uint32_t i = 0; i < n_shards; ++i) crc = crc32_combine(crc, shard_crcs[i], shard_records[i] * 16);
  uint8_t tail[4] = {(uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24)};
  bool ok = fseeko(f, 0, SEEK_SET) == 0 && fwrite(hdr, 1, 14, f) == 14;
  ok = ok && fseeko(f, (off_t)(14 + 16 * n), SEEK_SET) == 0 && fwrite(tail, 1, 4, f) == 4;
  ok = (fclose(f) == 0) && ok;
  return ok ? KMG_OK : KMG_ERR_IO;
}


// ---- results of a shard group (collective calls: every rank makes the same sequence of them) ---------------------------
// Global summary: sums over the (disjoint) shards, maximum of max_count; the timing fields stay this rank's own.
KMG_EXPORT kmg_status kmg_shard_finalize(kmg_ctx *c, kmg_summary *out) {
  if (!c) return KMG_ERR_INVALID_ARG;
  kmg_summary loc;
  kmg_status s = kmg_finalize(c, &loc);
  if (!c->shm) { if (out) *out = loc; return s; }
  if (s != KMG_OK) return shard_fail(c, s, c->err);
  ShardShm *S = c->shm;
  uint64_t *m = S->ranks[c->sh_rank].summary;
  m[0] = loc.n_windows; m[1] = loc.n_distinct; m[2] = loc.n_records; m[3] = loc.n_bases; m[4] = loc.max_count;
  SB(c);
  if (out) {
    *out = loc;
    out->n_windows = out->n_distinct = out->n_records = out->n_bases = out->max_count = 0;
    for (uint32_t r = 0; r < c->sh_world; ++r) {
      const uint64_t *o = S->ranks[r].summary;
      out->n_windows += o[0]; out->n_distinct += o[1]; out->n_records += o[2]; out->n_bases += o[3];
      out->max_count = std::max<uint64_t>(out->max_count, o[4]);
    }
  }
  SB(c);
  return KMG_OK;
}

// Count-of-counts over all shards: the element-wise sum of the shard histograms (a k-mer lives in exactly one shard).
// Replaces compute_histogram_packed (src/histogram.rs:110-116) for the sharded table.  Two-call protocol like kmg_histogram;
// the size query does the merge (collective), the call with arrays returns that merged result.
KMG_EXPORT kmg_status kmg_shard_histogram(kmg_ctx *c, uint64_t min_count, uint64_t *count_vals, uint64_t *freqs, uint64_t cap, uint64_t *n_out) {
  if (!c || !n_out) return KMG_ERR_INVALID_ARG;
  if (!c->shm) return kmg_histogram(c, min_count, count_vals, freqs, cap, n_out);
  if (!count_vals || !freqs) {
    uint64_t n = 0;
    kmg_status s = kmg_histogram(c, min_count, nullptr, nullptr, 0, &n);
    if (s != KMG_OK) return shard_fail(c, s, c->err);
    if (n > SHARD_HIST_MAX) return shard_fail(c, KMG_ERR_CAPACITY, "shard histogram has too many distinct count values");
    ShardShm *S = c->shm;
    uint64_t *mine = S->hist[c->sh_rank];
    if (n && (s = kmg_histogram(c, min_count, mine, mine + SHARD_HIST_MAX, n, &n)) != KMG_OK) return shard_fail(c, s, c->err);
    S->ranks[c->sh_rank].hist_n = n;
    SB(c);
    std::vector<std::pair<uint64_t, uint64_t>> all;
    for (uint32_t r = 0; r < c->sh_world; ++r)
      for (uint64_t i = 0; i < S->ranks[r].hist_n; ++i) all.emplace_back(S->hist[r][i], S->hist[r][SHARD_HIST_MAX + i]);
    SB(c);
    std::sort(all.begin(), all.end());
    c->sh_hist_vals.clear(); c->sh_hist_freqs.clear();
    for (size_t i = 0; i < all.size();) {
      uint64_t f = 0;
      size_t j = i;
      for (; j < all.size() && all[j].first == all[i].first; ++j) f += all[j].second;
      c->sh_hist_vals.push_back(all[i].first); c->sh_hist_freqs.push_back(f);
      i = j;
    }
    *n_out = c->sh_hist_vals.size();
    return KMG_OK;
  }
  const uint64_t n = c->sh_hist_vals.size();
  *n_out = n;
  if (cap < n) return fail(c, KMG_ERR_CAPACITY, "histogram arrays hold " + std::to_string(cap) + " bins, need " + std::to_string(n));
  if (n) { memcpy(count_vals, c->sh_hist_vals.data(), n * 8); memcpy(freqs, c->sh_hist_freqs.data(), n * 8); }
  return KMG_OK;
}

// ONE .kmix index from all shards of the group (src/index.rs:222-279; records in shard order, each shard ascending): rank 0
// creates the file, every rank writes its records into its own byte range, rank 0 writes the header (n = sum) and the CRC of
// the whole file combined from the shards' CRCs.  *n_records_out: records of the whole index.
KMG_EXPORT kmg_status kmg_shard_save_kmix(kmg_ctx *c, const char *path, uint64_t *n_records_out) {
  if (!c || !path) return KMG_ERR_INVALID_ARG;
  if (!c->shm) {
    kmg_status s = kmg_save_kmix(c, path);
    if (s == KMG_OK && n_records_out) { uint64_t n = 0; s = count_filtered(c, 0, &n); *n_records_out = n; }
    return s;
  }
  if (has_gz_suffix(path)) return shard_fail(c, KMG_ERR_INVALID_ARG, "kmg_shard_save_kmix writes uncompressed .kmix files");
  CU(c, cudaSetDevice(c->device));
  ShardShm *S = c->shm;
  uint64_t n = 0;
  kmg_status s = count_filtered(c, 0, &n);
  if (s != KMG_OK) return shard_fail(c, s, c->err);
  uint64_t *m = S->ranks[c->sh_rank].summary;
  m[5] = n;
  if (c->sh_rank == 0 && kmg_kmix_begin(path) != KMG_OK) return shard_fail(c, KMG_ERR_IO, std::string("cannot create ") + path);
  SB(c);
  uint64_t before = 0;
  for (uint32_t r = 0; r < c->sh_rank; ++r) before += S->ranks[r].summary[5];
  uint64_t n2 = 0;
  uint32_t crc = 0;
  s = kmg_save_kmix_shard(c, path, before, &n2, &crc);
  if (s != KMG_OK) return shard_fail(c, s, c->err);
  if (n2 != n) return shard_fail(c, KMG_ERR_STATE, "shard index writer wrote an unexpected number of records");
  m[6] = crc;
  SB(c);
  uint64_t total = 0;
  for (uint32_t r = 0; r < c->sh_world; ++r) total += S->ranks[r].summary[5];
  kmg_status fs = KMG_OK;
  if (c->sh_rank == 0) {
    uint64_t recs[SHARD_MAX_WORLD];
    uint32_t crcs[SHARD_MAX_WORLD];
    for (uint32_t r = 0; r < c->sh_world; ++r) { recs[r] = S->ranks[r].summary[5]; crcs[r] = (uint32_t)S->ranks[r].summary[6]; }
    fs = kmg_kmix_finish(path, (uint32_t)c->k, recs, crcs, c->sh_world);
    if (fs != KMG_OK) return shard_fail(c, fs, std::string("cannot finish ") + path);
  }
  SB(c);
  if (n_records_out) *n_records_out = total;
  return KMG_OK;
}

// Diagnostics of the exchange: keys this rank wrote into other ranks' buffers, keys it received, rounds, exact-route rounds.
KMG_EXPORT kmg_status kmg_shard_stats(const kmg_ctx *c, uint64_t *sent_keys, uint64_t *recv_keys, uint64_t *rounds, uint64_t *exact_rounds) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (sent_keys) *sent_keys = c->sh_sent;
  if (recv_keys) *recv_keys = c->sh_recv_keys;
  if (rounds) *rounds = c->sh_rounds;
  if (exact_rounds) *exact_rounds = c->sh_exact_rounds;
  return KMG_OK;
}


// Text emitters of output_counts / count_to_writer (src/run.rs:452-470, src/builder.rs:399-442) for --format tsv and fasta:
// records with count >= min_count in ascending k-mer order (the reference prints HashMap order; parity is on the sorted text),
// formatted ON THE DEVICE (2-bit unpack + decimal count), delivered to `sink` in chunks of at most 64 MiB.
KMG_EXPORT kmg_status kmg_emit_text(kmg_ctx *c, uint64_t min_count, int format, kmg_text_sink sink, void *user, uint64_t *n_records_out,
                                    uint64_t *n_bytes_out) {
  if (!c || !sink) return KMG_ERR_INVALID_ARG;
  if (format != KMG_TEXT_FASTA && format != KMG_TEXT_TSV) return fail(c, KMG_ERR_INVALID_ARG, "format must be KMG_TEXT_FASTA or KMG_TEXT_TSV");
  CU(c, cudaSetDevice(c->device));
  if (min_count == 0) min_count = 1;
  const int fasta = format == KMG_TEXT_FASTA;
  uint64_t records = 0, bytes = 0;
  HostStager stage(c, 64u << 20);
  uint64_t *d_off = nullptr;
  uint8_t *d_text = nullptr;
  uint64_t off_cap = 0, text_cap = 0;
  kmg_status s = for_each_sorted_piece(c, min_count, 1ull << 24, [&](const uint64_t *dk, const uint64_t *dc, uint64_t m) -> kmg_status {
    if (m + 1 > off_cap) { pool_free(c, d_off); d_off = nullptr; off_cap = std::max<uint64_t>(m + 1, 1ull << 20); CU(c, pool_alloc(c, &d_off, off_cap * 8)); }
    CU(c, launch_text_len(dc, m, c->k, fasta, d_off, c->stream));
    CU(c, cudaMemsetAsync(d_off + m, 0, 8, c->stream));
    if (!c->d_scan_tmp || c->scan_tmp_items < m + 1) {
      CU(c, cudaStreamSynchronize(c->stream));
      cudaFree(c->d_scan_tmp); c->d_scan_tmp = nullptr;
      CU(c, exclusive_sum_u64(nullptr, nullptr, m + 1, nullptr, &c->scan_tmp_bytes, c->stream));
      CU(c, cudaMalloc(&c->d_scan_tmp, c->scan_tmp_bytes ? c->scan_tmp_bytes : 16));
      c->scan_tmp_items = m + 1;
    }
    CU(c, exclusive_sum_u64(d_off, d_off, m + 1, c->d_scan_tmp, &c->scan_tmp_bytes, c->stream));  // entry m = bytes of the piece
    uint64_t total = 0;
    CU(c, cudaMemcpyAsync(&total, d_off + m, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (total > text_cap) { pool_free(c, d_text); d_text = nullptr; text_cap = total + total / 8; CU(c, pool_alloc(c, &d_text, text_cap)); }
    CU(c, launch_text_write(dk, dc, d_off, m, c->k, fasta, d_text, c->stream));
    records += m; bytes += total;
    return stage.drain(d_text, total, [&](const uint8_t *p, size_t len) -> kmg_status {
      return sink(user, p, len) == 0 ? KMG_OK : fail(c, KMG_ERR_IO, "the text sink reported an error");
    });
  });
  pool_free(c, d_off); pool_free(c, d_text);
  if (n_records_out) *n_records_out = records;
  if (n_bytes_out) *n_bytes_out = bytes;
  return s;
}

namespace {
int file_sink(void *user, const uint8_t *p, size_t n) { return fwrite(p, 1, n, static_cast<FILE *>(user)) == n ? 0 : 1; }
}  // namespace

// The same into a file ("-" = stdout): what `kmerust <k> <path> --format tsv|fasta --min-count m > file` produces, sorted.
KMG_EXPORT kmg_status kmg_write_text(kmg_ctx *c, uint64_t min_count, int format, const char *path, uint64_t *n_records_out, uint64_t *n_bytes_out) {
  if (!c || !path) return KMG_ERR_INVALID_ARG;
  const bool to_stdout = strcmp(path, "-") == 0;
  FILE *f = to_stdout ? stdout : fopen(path, "wb");
  if (!f) return fail(c, KMG_ERR_IO, std::string("cannot open ") + path + " for writing");
  kmg_status s = kmg_emit_text(c, min_count, format, file_sink, f, n_records_out, n_bytes_out);
  const bool ok = to_stdout ? fflush(f) == 0 : fclose(f) == 0;
  if (s == KMG_OK && !ok) return fail(c, KMG_ERR_IO, std::string("short write to ") + path);
  return s;
}


// ---- index queries and loading (src/index.rs:127-131 KmerIndex::get, :199-216 load_index, :282-401 read_index; the consumer is
// the `query` subcommand, src/main.rs:233-281) --------------------------------------------------------------------------------
namespace {
kmg_status query_device_keys(kmg_ctx *c, const uint64_t *d_keys, uint64_t n, uint64_t *counts_out) {
  { kmg_status _fs = scan_pending_batch(c); if (_fs != KMG_OK) return _fs; }
  CU(c, cudaStreamSynchronize(c->copy_stream));
  if (c->mode == kmg_ctx::MODE_PARTITIONED) { kmg_status s = consolidate(c); if (s != KMG_OK) return s; }
  uint64_t *d_counts = nullptr;
  CU(c, pool_alloc(c, &d_counts, std::max<uint64_t>(n, 1) * 8));
  const bool part = c->mode == kmg_ctx::MODE_PARTITIONED;
  TableView v = view_of(c);
  cudaError_t e = cudaSuccess;
  if (part && !c->has_result) e = cudaMemsetAsync(d_counts, 0, n * 8, c->stream);  // nothing counted yet
  else e = launch_query(v, c->table, part ? c->result.d_seg_start : nullptr, part ? c->result.d_seg_len : nullptr, std::max<uint32_t>(c->n_coarse, 1),
                        std::max<uint32_t>(c->n_sub, 1), c->shm ? c->sh_world : 1, c->shm ? c->sh_rank : 0, d_keys, n, d_counts, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts_out, d_counts, n * 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  pool_free(c, d_counts);
  if (e != cudaSuccess) return cuda_fail(c, e, "query");
  return KMG_OK;
}
}  // namespace

// counts_out[i] = count of the canonical packed key keys[i] (0 when absent).  HOST arrays.
KMG_EXPORT kmg_status kmg_query_keys(kmg_ctx *c, const uint64_t *keys, uint64_t n, uint64_t *counts_out) {
  if (!c || (n && (!keys || !counts_out))) return KMG_ERR_INVALID_ARG;
  if (n == 0) return KMG_OK;
  CU(c, cudaSetDevice(c->device));
  uint64_t *d_keys = nullptr;
  CU(c, pool_alloc(c, &d_keys, n * 8));
  cudaError_t e = cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, c->stream);
  kmg_status s = e == cudaSuccess ? query_device_keys(c, d_keys, n, counts_out) : cuda_fail(c, e, "H2D query keys");
  pool_free(c, d_keys);
  return s;
}

// n k-mers as ASCII, k bytes each, back to back, any case: canonicalised on the device (upper-case, pack, min(fwd, rc) -- what the
// query subcommand does on the host) and looked up.  A k-mer with a byte outside ACGTacgt counts 0 and sets *n_invalid_out.
KMG_EXPORT kmg_status kmg_query_ascii(kmg_ctx *c, const uint8_t *kmers, uint64_t n, uint64_t *counts_out, uint64_t *n_invalid_out) {
  if (!c || (n && (!kmers || !counts_out))) return KMG_ERR_INVALID_ARG;
  if (n_invalid_out) *n_invalid_out = 0;
  if (n == 0) return KMG_OK;
  CU(c, cudaSetDevice(c->device));
  uint8_t *d_kmers = nullptr;
  uint64_t *d_keys = nullptr;
  CU(c, pool_alloc(c, &d_kmers, n * (uint64_t)c->k));
  cudaError_t e = pool_alloc(c, &d_keys, n * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_kmers, kmers, n * (uint64_t)c->k, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = launch_query_pack(d_kmers, n, c->k, d_keys, c->stream);
  kmg_status s = e == cudaSuccess ? query_device_keys(c, d_keys, n, counts_out) : cuda_fail(c, e, "query (pack)");
  if (s == KMG_OK && n_invalid_out) {
    std::vector<uint64_t> hk(n);
    e = cudaMemcpy(hk.data(), d_keys, n * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) s = cuda_fail(c, e, "query (invalid k-mers)");
    else for (uint64_t x : hk) *n_invalid_out += x == EMPTY_KEY;
  }
  pool_free(c, d_kmers); pool_free(c, d_keys);
  return s;
}

// Open a .kmix index as a ready-to-query context on `device`: the checks of read_index in the reference's order (size >= 18,
// magic, CRC, version, k, n * 16 == data size; src/index.rs:282-401), then the records are streamed to the device and
// upserted.  Errors: KMG_ERR_IO (cannot read) / KMG_ERR_PARSE (invalid index; message via kmg_last_error(NULL)).
KMG_EXPORT kmg_status kmg_index_open(const char *path, int32_t device, kmg_ctx **out) {
  if (!path || !out) return KMG_ERR_INVALID_ARG;
  *out = nullptr;
  if (has_gz_suffix(path)) return fail(nullptr, KMG_ERR_INVALID_ARG, "kmg_index_open reads uncompressed .kmix files; gunzip on the host side");
  FILE *f = fopen(path, "rb");
  if (!f) return fail(nullptr, KMG_ERR_IO, std::string("failed to read index from '") + path + "'");
  struct Closer { FILE *f; ~Closer() { if (f) fclose(f); } } closer{f};
  if (fseeko(f, 0, SEEK_END) != 0) return fail(nullptr, KMG_ERR_IO, "seek failed");
  const uint64_t size = (uint64_t)ftello(f);
  rewind(f);
  if (size < 18) return fail(nullptr, KMG_ERR_PARSE, "file too small");
  uint8_t hdr[14];
  if (fread(hdr, 1, 14, f) != 14) return fail(nullptr, KMG_ERR_IO, "short read");
  if (memcmp(hdr, "KMIX", 4) != 0) return fail(nullptr, KMG_ERR_PARSE, "invalid magic bytes (not a kmerust index file)");
  // CRC over everything before the last four bytes, streamed (the reference buffers the whole file)
  const Crc32 &T = crc_tables();
  uint32_t crc = ~T.update(~0u, hdr, 14);
  const uint64_t body = size - 18;
  std::vector<uint8_t> buf((size_t)std::min<uint64_t>(std::max<uint64_t>(body, 16), 256ull << 20));
  for (uint64_t done = 0; done < body;) {
    const size_t m = (size_t)std::min<uint64_t>(buf.size(), body - done);
    if (fread(buf.data(), 1, m, f) != m) return fail(nullptr, KMG_ERR_IO, "short read");
    crc = crc32_combine(crc, crc32_parallel(buf.data(), m), m);
    done += m;
  }
  uint8_t tail[4];
  if (fread(tail, 1, 4, f) != 4) return fail(nullptr, KMG_ERR_IO, "short read");
  const uint32_t stored = (uint32_t)tail[0] | (uint32_t)tail[1] << 8 | (uint32_t)tail[2] << 16 | (uint32_t)tail[3] << 24;
  if (stored != crc) { char m[96]; snprintf(m, sizeof m, "checksum mismatch (expected %#x, got %#x)", stored, crc); return fail(nullptr, KMG_ERR_PARSE, m); }
  if (hdr[4] != 1) return fail(nullptr, KMG_ERR_PARSE, "unsupported version " + std::to_string(hdr[4]));
  const uint32_t k = hdr[5];
  if (k < 1 || k > 32) return fail(nullptr, KMG_ERR_PARSE, "invalid k-mer length: k-mer length " + std::to_string(k) + " is out of range (must be 1-32)");
  uint64_t n = 0;
  for (int b = 0; b < 8; ++b) n |= (uint64_t)hdr[6 + b] << (8 * b);
  if (body != n * 16) return fail(nullptr, KMG_ERR_PARSE, "data size mismatch (expected " + std::to_string(n * 16) + " bytes, got " + std::to_string(body) + " bytes)");
  kmg_config cfg{};
  cfg.abi_version = KMG_ABI_VERSION; cfg.k = k; cfg.device = device; cfg.expected_distinct = std::max<uint64_t>(n, 1);
  kmg_ctx *c = nullptr;
  kmg_status s = kmg_create(&cfg, &c);
  if (s != KMG_OK) return s;
  auto bail = [&](kmg_status st) { g_create_error = c->err; kmg_destroy(c); return st; };
  if (n) {
    if (fseeko(f, 14, SEEK_SET) != 0) return bail(fail(c, KMG_ERR_IO, "seek failed"));
    const uint64_t CH = std::min<uint64_t>(n, 1ull << 24);  // records per piece (256 MiB)
    void *d_pairs = nullptr;
    uint64_t *dk = nullptr, *dc = nullptr;
    cudaError_t e = pool_alloc(c, &d_pairs, CH * 16);
    if (e == cudaSuccess) e = pool_alloc(c, &dk, CH * 8);
    if (e == cudaSuccess) e = pool_alloc(c, &dc, CH * 8);
    if (e != cudaSuccess) return bail(cuda_fail(c, e, "cudaMalloc(index load)"));
    buf.resize((size_t)(CH * 16));
    for (uint64_t done = 0; done < n && s == KMG_OK;) {
      const uint64_t m = std::min<uint64_t>(CH, n - done);
      if (fread(buf.data(), 16, m, f) != m) { s = fail(c, KMG_ERR_IO, "short read"); break; }
      e = cudaMemcpyAsync(d_pairs, buf.data(), m * 16, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = launch_deinterleave_pairs(d_pairs, m, dk, dc, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) { s = cuda_fail(c, e, "index load"); break; }
      s = kmg_insert_keys_device(c, dk, dc, m);
      done += m;
    }
    pool_free(c, d_pairs); pool_free(c, dk); pool_free(c, dc);
    if (s == KMG_OK) s = kmg_finalize(c, nullptr);
    if (s != KMG_OK) return bail(s);
  }
  *out = c;
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_progress(const kmg_ctx *c, uint64_t *records, uint64_t *bases) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (records) *records = c->n_records;
  if (bases) *bases = c->n_bases;
  return KMG_OK;
}

KMG_EXPORT uint64_t kmg_kernel_launches(void) { return kernel_launches(); }

KMG_EXPORT kmg_status kmg_phase_times(kmg_ctx *c, uint64_t *out_ns) {
  if (!c || !out_ns) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  timers_collect(c);
  const double a2 = c->cat_ms[2], a1 = std::max(0.0, c->cat_ms[0] - a2), b = c->cat_ms[1];
  out_ns[0] = (uint64_t)(a1 * 1e6); out_ns[1] = (uint64_t)(a2 * 1e6); out_ns[2] = (uint64_t)(b * 1e6);
  out_ns[3] = (uint64_t)(std::max(0.0, c->kernel_ms - c->cat_ms[0] - b) * 1e6);
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_synth_uniform_device(kmg_ctx *c, uint64_t seed, uint64_t first_base, uint64_t n, uint8_t *d_out) {
  if (!c) return KMG_ERR_INVALID_ARG;
  CU(c, cudaSetDevice(c->device));
  if (n) CU(c, launch_synth_uniform(seed, first_base, n, d_out, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return KMG_OK;
}

KMG_EXPORT kmg_status kmg_synth_reads_device(kmg_ctx *c, uint64_t seed, uint32_t profile, uint64_t first_read, uint64_t n_reads,
                                             uint8_t *d_seq, uint8_t *d_qual) {
  if (!c) return KMG_ERR_INVALID_ARG;
  if (profile != 3 && profile != 5) return fail(c, KMG_ERR_INVALID_ARG, "profile must be 3 (R20M shape) or 5 (R200M shape)");
  if (n_reads && !d_seq) return fail(c, KMG_ERR_INVALID_ARG, "d_seq is NULL");
  CU(c, cudaSetDevice(c->device));
  CU(c, launch_synth_reads(seed, profile, first_read, n_reads, d_seq, d_qual, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return KMG_OK;
}
