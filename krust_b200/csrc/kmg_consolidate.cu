// Phase B of the partitioned pipeline: count every hash partition in an L2-RESIDENT table.
//
// Why: an upsert into an HBM-resident table is bounded by ~20 G random atomics/s on a B200, but the
// same atomics run at 127-200 G/s when the table fits the 126 MB L2 (profiles/microbench_r1.jsonl).
// So keys are first scattered into P hash partitions with streaming writes (phase A,
// scan_partition_kernel), and here each partition (a few hundred thousand keys) is upserted into a
// small table that lives in L2, compacted into the output run and the table is handed to a later
// partition.  HBM sees only streaming traffic: 8 B/key in, 16 B/distinct key out.
//
// One persistent kernel, no grid-wide barriers: CTAs draw tickets from a global counter; the ticket
// order I(0) I(1) C(0) I(2) C(1) ... I(P-1) C(P-2) C(P-1) (I = insert chunk, C = compact chunk)
// interleaves three table buffers so that every dependency points at least two phases back:
//     I(p) needs C(p-3) finished (its buffer is free again),  C(p) needs I(p) finished.
// Dependencies only ever point to lower tickets, which are held by CTAs that are already running,
// so the scheme cannot deadlock and needs no cooperative launch.
//
// Replaces the DashMap upsert + iteration of src/run.rs:565-582 for large inputs; results are the
// same multiset of (canonical key, count) pairs.
#include <atomic>

#include "kmg_device.cuh"
#include "kmg_kernels.h"

namespace kmg {

namespace {

constexpr unsigned long long BASE_UNSET = ~0ull;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace

// Ticket life cycle (per CTA): warp 0 prepares (ticket, phase decode, dependency wait, input segments) ->
// barrier -> all warps work without further block barriers (reservations are per warp) -> fence + barrier ->
// thread 0 signals completion.  The next ticket is drawn early so its latency hides behind the work; the
// lowest unfinished ticket is always executing with all its dependencies done, so progress is guaranteed.
__global__ void __launch_bounds__(CONS_THREADS, CONS_CTAS_PER_SM) consolidate_kernel(ConsParams P) {
  __shared__ uint64_t seg_begin[CONS_MAX_RUNS];
  __shared__ uint32_t seg_prefix[CONS_MAX_RUNS + 1];
  __shared__ uint32_t s_ticket, s_phase;
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp-0/lane-0 private scheduling state
  uint32_t phase_idx = 0;                          // monotone: tickets only grow
  uint32_t next_ticket = 0;
  int64_t known_pub = 0;                           // out_base[0..known_pub] are known to be published
  int64_t known_done[CONS_NBUF];                   // per table buffer: last partition whose C is known complete
#pragma unroll
  for (int b = 0; b < CONS_NBUF; ++b) known_done[b] = -1;
  if (tid == 0) next_ticket = atomicAdd(P.ticket, 1u);

  for (;;) {
    if (warp == 0) {
      uint32_t t = 0, ph_idx = 0;
      if (lane == 0) {
        t = next_ticket;
        if (t < P.total_tickets) {
          while (P.phases[phase_idx + 1].first_ticket <= t) ++phase_idx;
          next_ticket = atomicAdd(P.ticket, 1u);  // consumed one iteration later
        }
        ph_idx = phase_idx;
      }
      t = __shfl_sync(0xffffffffu, t, 0);
      ph_idx = __shfl_sync(0xffffffffu, ph_idx, 0);
      if (t < P.total_tickets) {
        const ConsPhase ph = P.phases[ph_idx];
        const uint32_t p = ph.part_and_type & 0x7fffffffu;
        if (!(ph.part_and_type >> 31)) {
          // insert ticket: input segments of partition p across the runs, and the table buffer must be free
          uint32_t len = 0;
          uint64_t b = 0;
          if (lane < (int)P.R) { b = P.runs[lane].offsets[p]; len = (uint32_t)(P.runs[lane].offsets[p + 1] - b); }
          uint32_t incl = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
          if (lane < (int)P.R) { seg_begin[lane] = b; seg_prefix[lane] = incl - len; }
          if (lane == (int)P.R - 1) seg_prefix[P.R] = incl;
          if (lane == 0) {
            const uint32_t q = P.part_wait[p];  // previous non-empty user of this table buffer
            if (q != 0xffffffffu && (int64_t)q > known_done[p % CONS_NBUF]) {
              const uint32_t need = P.part_nC[q];
              while (ld_acquire_u32(P.done_C + q) < need) __nanosleep(100);
              known_done[p % CONS_NBUF] = q;
            }
          }
        } else if (lane == 0) {
          // compact ticket: partition p must be completely inserted and its output offset known
          if ((int64_t)p + 1 > known_pub) {
            while (ld_acquire_u64(P.out_base + p + 1) == BASE_UNSET) __nanosleep(100);
            known_pub = (int64_t)p + 1;
          }
          s_base = __ldcg(P.out_base + p);
        }
      }
      if (lane == 0) { s_ticket = t; s_phase = ph_idx; }
    }
    __syncthreads();
    const uint32_t ticket = s_ticket;
    if (ticket >= P.total_tickets) break;
    const ConsPhase ph = P.phases[s_phase];
    const uint32_t chunk = ticket - ph.first_ticket;
    const uint32_t p = ph.part_and_type & 0x7fffffffu;
    const bool is_compact = ph.part_and_type >> 31;
    const uint64_t mask = (1ull << P.part_cap_log2[p]) - 1;
    unsigned long long *table = reinterpret_cast<unsigned long long *>(P.tables) + (uint64_t)(p % CONS_NBUF) * P.table_stride_slots * 2;

    if (!is_compact) {
      // ------------------------------------------------------------------ I(p): upsert one chunk of partition p
      const uint32_t n_p = seg_prefix[P.R];
      const uint32_t lo = chunk * CONS_INSERT_CHUNK;
      constexpr int G = 8, ROUNDS = CONS_INSERT_CHUNK / (CONS_THREADS * G);
      uint32_t new_keys = 0;
#pragma unroll 1
      for (int rd = 0; rd < ROUNDS; ++rd) {
        if (lo + (uint32_t)rd * G * CONS_THREADS >= n_p) break;  // block-uniform
        uint64_t key[G], old[G], slot[G], w[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const uint32_t idx = lo + (rd * G + j) * CONS_THREADS + tid;
          w[j] = 0; key[j] = EMPTY_KEY;
          if (idx < n_p) {
            uint32_t r = 0;
            while (r + 1 < P.R && idx >= seg_prefix[r + 1]) ++r;
            const uint64_t src = seg_begin[r] + (idx - seg_prefix[r]);
            key[j] = __ldcs(P.runs[r].keys + src);
            w[j] = P.runs[r].counts ? __ldcs(P.runs[r].counts + src) : 1ull;
          }
        }
        if (P.preagg) {
          // Warp run-length pre-aggregation: phase A writes the keys of consecutive windows next to each other,
          // so homopolymer / tandem-repeat runs arrive as runs of equal keys in adjacent lanes.  The head lane of
          // each run upserts once with the run length; this bounds same-address atomic bursts on skewed inputs.
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const uint64_t kk = w[j] ? key[j] : EMPTY_KEY;
            const uint64_t kp = __shfl_up_sync(0xffffffffu, kk, 1);
            const bool head = lane == 0 || kp != kk;
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            if (__all_sync(0xffffffffu, w[j] <= 1ull)) {  // unit weights only (keys-runs); pair-runs are distinct per run
              const uint32_t above = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
              const uint32_t end = above ? (uint32_t)__ffs(above) - 1u : 32u;
              if (w[j]) w[j] = head ? (uint64_t)(end - lane) : 0ull;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {  // all first probes in flight together
          old[j] = 0;
          slot[j] = mix64(key[j]) & mask;  // low mix bits; the partition index used the high ones
          if (w[j]) old[j] = atomicCAS(table + 2 * slot[j], EMPTY_KEY, key[j]);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
          if (!w[j]) continue;
          uint64_t sl = slot[j], cur = old[j], probes = 0;
          while (cur != EMPTY_KEY && cur != key[j]) {  // linear probing inside the partition's table
            if (++probes > mask) { atomicExch(P.error_flag, 1u); break; }
            sl = (sl + 1) & mask;
            cur = *reinterpret_cast<volatile unsigned long long *>(table + 2 * sl);
            if (cur == EMPTY_KEY) cur = atomicCAS(table + 2 * sl, EMPTY_KEY, key[j]);
          }
          if (cur == EMPTY_KEY) { ++new_keys; if (w[j] > 1) atomicAdd(table + 2 * sl + 1, (unsigned long long)(w[j] - 1)); }
          else if (cur == key[j]) atomicAdd(table + 2 * sl + 1, (unsigned long long)w[j]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) new_keys += __shfl_xor_sync(0xffffffffu, new_keys, o);
      if (lane == 0 && new_keys) atomicAdd(P.distinct + p, new_keys);
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const uint32_t done = atomicAdd(P.done_I + p, 1u) + 1;
        if (done == P.part_nI[p]) {  // last inserter of p: its distinct count is final -> publish where p+1 starts
          __threadfence();
          unsigned long long b;
          while ((b = ld_acquire_u64(P.out_base + p)) == BASE_UNSET) __nanosleep(100);
          const uint32_t d = atomicAdd(P.distinct + p, 0u);
          st_release_u64(P.out_base + p + 1, b + d);
          if ((int64_t)p + 1 > known_pub) known_pub = (int64_t)p + 1;
        }
      }
    } else {
      // ------------------------------------------------------------------ C(p): drain one chunk of p's table
      const unsigned long long base = s_base;
      constexpr int G = 8, ROUNDS = CONS_COMPACT_CHUNK / (CONS_THREADS * G);
      const uint64_t lo = (uint64_t)chunk * CONS_COMPACT_CHUNK;
#pragma unroll 1
      for (int rd = 0; rd < ROUNDS; ++rd) {
        if (lo + (uint64_t)rd * G * CONS_THREADS > mask) break;  // block-uniform
        ulonglong2 sl[G];
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const uint64_t i = lo + (uint64_t)(rd * G + j) * CONS_THREADS + tid;
          sl[j] = make_ulonglong2(EMPTY_KEY, 0ull);
          if (i <= mask) sl[j] = __ldcg(reinterpret_cast<const ulonglong2 *>(table) + i);  // L2, never a stale L1 line
          mine += sl[j].x != EMPTY_KEY;
        }
        uint32_t incl = mine;  // warp-level reservation: one atomic per warp per round, no block barrier
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        uint32_t wbase = 0;
        if (lane == 31 && incl) wbase = atomicAdd(P.out_cursor + p, incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        uint64_t o = base + wbase + (incl - mine);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          if (sl[j].x == EMPTY_KEY) continue;
          __stcs(P.out_keys + o, sl[j].x);
          __stcs(P.out_counts + o, sl[j].y + 1);  // slots store occurrences - 1
          ++o;
          const uint64_t i = lo + (uint64_t)(rd * G + j) * CONS_THREADS + tid;
          reinterpret_cast<ulonglong2 *>(table)[i] = make_ulonglong2(EMPTY_KEY, 0ull);  // hand the slot back clean
        }
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(P.done_C + p, 1u);
    }
  }
}

namespace {
}  // namespace

// scatter already-extracted keys (optionally weighted) into partitions: the receive side of the
// multi-GPU exchange and re-partitioning of foreign runs.  pass 0 counts, pass 1 scatters.
template <bool SCATTER>
__global__ void __launch_bounds__(256) partition_keys_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ counts,
                                                             uint64_t n, uint32_t n_parts, unsigned long long *part_counts,
                                                             const unsigned long long *part_start, unsigned long long *part_cursor,
                                                             uint64_t *out_keys, uint64_t *out_counts) {
  extern __shared__ uint32_t sm[];
  uint32_t *hist = sm, *toff = sm + n_parts;
  constexpr uint32_t TILE = 8192;
  for (uint32_t p = threadIdx.x; p < n_parts; p += blockDim.x) hist[p] = 0;
  __syncthreads();
  for (uint64_t t0 = (uint64_t)blockIdx.x * TILE; t0 < n; t0 += (uint64_t)gridDim.x * TILE) {
    const uint32_t m = (uint32_t)(n - t0 < TILE ? n - t0 : TILE);
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) atomicAdd(hist + part_of(keys[t0 + i], n_parts), 1u);
    if (SCATTER) {
      __syncthreads();
      for (uint32_t p = threadIdx.x; p < n_parts; p += blockDim.x) {
        const uint32_t c = hist[p];
        toff[p] = c ? (uint32_t)atomicAdd(part_cursor + p, (unsigned long long)c) : 0;
        hist[p] = 0;
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const uint64_t key = keys[t0 + i];
        const uint32_t p = part_of(key, n_parts);
        const uint64_t o = __ldg(part_start + p) + toff[p] + atomicAdd(hist + p, 1u);
        out_keys[o] = key;
        if (out_counts) out_counts[o] = counts ? counts[t0 + i] : 1ull;
      }
      __syncthreads();
      for (uint32_t p = threadIdx.x; p < n_parts; p += blockDim.x) hist[p] = 0;
      __syncthreads();
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < n_parts; p += blockDim.x)
      if (hist[p]) atomicAdd(part_counts + p, (unsigned long long)hist[p]);
  }
}

extern std::atomic<uint64_t> g_launches;

cudaError_t launch_consolidate(const ConsParams &P, int num_sms, cudaStream_t s) {
  if (P.total_tickets == 0) return cudaSuccess;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  consolidate_kernel<<<num_sms * CONS_CTAS_PER_SM, CONS_THREADS, 0, s>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_partition_keys(const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n, uint32_t n_parts, bool scatter,
                                  unsigned long long *part_counts, const unsigned long long *part_start,
                                  unsigned long long *part_cursor, uint64_t *out_keys, uint64_t *out_counts, int num_sms,
                                  cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const size_t smem = 2 * (size_t)n_parts * sizeof(uint32_t);
  uint64_t want = (n + 8191) / 8192;
  unsigned grid = (unsigned)(want < (uint64_t)num_sms * 4 ? want : (uint64_t)num_sms * 4);
  cudaError_t e;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (scatter) {
    if ((e = cudaFuncSetAttribute(partition_keys_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    partition_keys_kernel<true><<<grid, 256, smem, s>>>(d_keys, d_counts, n, n_parts, part_counts, part_start, part_cursor, out_keys, out_counts);
  } else {
    if ((e = cudaFuncSetAttribute(partition_keys_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    partition_keys_kernel<false><<<grid, 256, smem, s>>>(d_keys, d_counts, n, n_parts, part_counts, part_start, part_cursor, out_keys, out_counts);
  }
  return cudaGetLastError();
}

}  // namespace kmg
