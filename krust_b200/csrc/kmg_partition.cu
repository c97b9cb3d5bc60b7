// Partitioned pipeline, phases A2 and B (phase A1, the scan-side coarse scatter, lives in kmg_kernels.cu).
//
// Why partition at all: an upsert into one big HBM-resident table is bounded by ~20 G random atomics/s
// on a B200 and moves 144 B of DRAM traffic per k-mer (profiles/r1_summary.md), while HBM streams at
// 6.5 TB/s.  So the keys are radix-partitioned by hash with streaming writes until a partition is so
// small (~4.4 K keys) that ONE CTA can count it in a private 128 KiB scratch table that never leaves
// L2, compact it into the output and move on -- no cross-CTA dependencies, no grid barriers, HBM sees
// only streams.
//
//   A1  scan -> canonical keys -> P1 coarse partitions           (partition_count/scatter_kernel)
//   A2  coarse partition -> P2 sub-bins each (P = P1*P2 fine)    (refine_kernel<false/true>, here)
//   B   one CTA per fine partition: upsert, compact, clean       (count_partitions_kernel, here)
//
// Two levels because a scatter needs long runs per (tile, bin) to write whole sectors: with ~850 bins per
// level a 16-32 K-key tile yields 20-40 key runs, while a single level with 700 K bins would not.
//
// Replaces the DashMap upsert + iteration of src/run.rs:565-582 for large inputs; results are the same
// multiset of (canonical key, count) pairs.
#include <algorithm>
#include <atomic>

#include "kmg_device.cuh"
#include "kmg_kernels.h"

namespace kmg {

extern std::atomic<uint64_t> g_launches;

// ---------------------------------------------------------------------------------------------------
// level 1 for keys that are already extracted (receive side of the multi-GPU exchange, weighted inserts)
// ---------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) keys_coarse_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ counts,
                                                          uint64_t n, uint32_t n_coarse, unsigned long long *coarse_counts,
                                                          const unsigned long long *coarse_start, unsigned long long *coarse_cursor,
                                                          uint64_t *out_keys, uint64_t *out_counts) {
  extern __shared__ uint32_t sm[];
  uint32_t *hist = sm;  // histogram, then ABSOLUTE output cursors (the atomic's return value is the destination)
  constexpr uint32_t TILE = 16384;
  for (uint32_t p = threadIdx.x; p < n_coarse; p += blockDim.x) hist[p] = 0;
  __syncthreads();
  constexpr int U = 8;  // keys in flight per thread: the loops are otherwise serialised on one global load each
  for (uint64_t t0 = (uint64_t)blockIdx.x * TILE; t0 < n; t0 += (uint64_t)gridDim.x * TILE) {
    const uint32_t m = (uint32_t)(n - t0 < TILE ? n - t0 : TILE);
    for (uint32_t i0 = 0; i0 < m; i0 += U * 256) {
      uint64_t key[U];
#pragma unroll
      for (int j = 0; j < U; ++j) { const uint32_t i = i0 + j * 256 + threadIdx.x; key[j] = i < m ? keys[t0 + i] : EMPTY_KEY; }
#pragma unroll
      for (int j = 0; j < U; ++j) if (i0 + j * 256 + threadIdx.x < m) atomicAdd(hist + coarse_of_mix(mix64(key[j]), n_coarse), 1u);
    }
    if (SCATTER) {
      __syncthreads();
      for (uint32_t p = threadIdx.x; p < n_coarse; p += blockDim.x) {
        const uint32_t c = hist[p];
        hist[p] = c ? (uint32_t)(coarse_start[p] + atomicAdd(coarse_cursor + p, (unsigned long long)c)) : 0u;
      }
      __syncthreads();
      for (uint32_t i0 = 0; i0 < m; i0 += U * 256) {
        uint64_t key[U], cnt[U];
        uint32_t p[U], o[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const uint32_t i = i0 + j * 256 + threadIdx.x;
          key[j] = i < m ? keys[t0 + i] : EMPTY_KEY;
          cnt[j] = (i < m && counts) ? counts[t0 + i] : 1ull;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) { p[j] = coarse_of_mix(mix64(key[j]), n_coarse); o[j] = 0; if (i0 + j * 256 + threadIdx.x < m) o[j] = atomicAdd(hist + p[j], 1u); }
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (i0 + j * 256 + threadIdx.x < m) {
            out_keys[o[j]] = key[j];
            if (out_counts) out_counts[o[j]] = cnt[j];
          }
      }
      __syncthreads();
      for (uint32_t p = threadIdx.x; p < n_coarse; p += blockDim.x) hist[p] = 0;
      __syncthreads();
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < n_coarse; p += blockDim.x)
      if (hist[p]) atomicAdd(coarse_counts + p, (unsigned long long)hist[p]);
  }
}

cudaError_t launch_keys_coarse(const uint64_t *d_keys, const uint64_t *d_counts, uint64_t n, uint32_t n_coarse, bool scatter,
                               unsigned long long *coarse_counts, const unsigned long long *coarse_start,
                               unsigned long long *coarse_cursor, uint64_t *out_keys, uint64_t *out_counts, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const size_t smem = (size_t)n_coarse * sizeof(uint32_t);
  const uint64_t want = (n + 16383) / 16384, cap = (uint64_t)num_sms() * 2;
  const unsigned grid = (unsigned)std::min(want, cap);
  cudaError_t e;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (scatter) {
    if ((e = cudaFuncSetAttribute(keys_coarse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    keys_coarse_kernel<true><<<grid, 256, smem, s>>>(d_keys, d_counts, n, n_coarse, coarse_counts, coarse_start, coarse_cursor, out_keys, out_counts);
  } else {
    if ((e = cudaFuncSetAttribute(keys_coarse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    keys_coarse_kernel<false><<<grid, 256, smem, s>>>(d_keys, d_counts, n, n_coarse, coarse_counts, coarse_start, coarse_cursor, out_keys, out_counts);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// level 2: split every coarse partition into n_sub sub-bins.  Work unit = a tile of REFINE_TILE keys that
// lies inside ONE coarse partition (tile_prefix[] maps a global tile number to its partition).
// SCATTER == false: fine_counts[c*n_sub + s] += ...      SCATTER == true: fine_start[] is the exclusive
// prefix of those counts and fine_cursor[] starts at zero.
// ---------------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(REFINE_THREADS) refine_kernel(RefineParams P) {
  extern __shared__ __align__(16) uint8_t rsm[];
  // count pass: one histogram.  scatter pass: the tile's keys are ranked into a shared-memory staging buffer in
  // sub-bin order and copied out by consecutive lanes (see partition_scatter_staged_kernel: scattered 8-byte
  // stores, one L2 transaction each, were 60 % of the unstaged kernels).
  uint64_t *staging = reinterpret_cast<uint64_t *>(rsm);                                // SCATTER only: REFINE_TILE keys
  uint64_t *staging_c = staging + REFINE_TILE;                                          // SCATTER with counts only
  uint32_t *hist = reinterpret_cast<uint32_t *>(rsm + (SCATTER ? (size_t)REFINE_TILE * 8 * (P.counts ? 2 : 1) : 0));
  uint32_t *s_off = hist + P.n_sub, *g_base = s_off + P.n_sub;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ uint32_t s_c, s_scan[REFINE_THREADS / 32 + 1];
  for (uint32_t s = tid; s < P.n_sub; s += REFINE_THREADS) hist[s] = 0;
  __syncthreads();
  for (uint32_t g = blockIdx.x; g < P.n_tiles; g += gridDim.x) {
    if (tid == 0) {  // which coarse partition owns tile g: largest c with tile_prefix[c] <= g
      uint32_t lo = 0, hi = P.n_coarse;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.tile_prefix[mid] <= g) lo = mid; else hi = mid; }
      s_c = lo;
    }
    __syncthreads();
    const uint32_t c = s_c;
    const uint64_t begin = P.coarse_start[c] + (uint64_t)(g - P.tile_prefix[c]) * REFINE_TILE;
    const uint64_t end_c = P.coarse_start[c + 1];
    const uint32_t m = (uint32_t)(end_c - begin < (uint64_t)REFINE_TILE ? end_c - begin : (uint64_t)REFINE_TILE);
    constexpr int U = 8;  // keys in flight per thread
    for (uint32_t i0 = 0; i0 < m; i0 += U * REFINE_THREADS) {
      uint64_t key[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const uint32_t i = i0 + j * REFINE_THREADS + tid;
        key[j] = i < m ? (SCATTER ? P.keys[begin + i] : __ldcs(P.keys + begin + i)) : EMPTY_KEY;
      }
#pragma unroll
      for (int j = 0; j < U; ++j) if (i0 + j * REFINE_THREADS + tid < m) atomicAdd(hist + sub_of_mix(mix64(key[j]), P.n_sub), 1u);
    }
    __syncthreads();
    const uint64_t f0 = (uint64_t)c * P.n_sub;
    if (!SCATTER) {
      for (uint32_t s = tid; s < P.n_sub; s += REFINE_THREADS) {
        const uint32_t h = hist[s];
        if (h) { atomicAdd(P.fine_counts + f0 + s, (unsigned long long)h); hist[s] = 0; }
      }
    } else {
      // exclusive prefix over the sub-bins (staging offsets) + one global reservation per sub-bin
      const uint32_t per = (P.n_sub + REFINE_THREADS - 1) / REFINE_THREADS;
      const uint32_t b0 = tid * per, b1 = b0 + per < P.n_sub ? b0 + per : P.n_sub;
      uint32_t mine = 0;
      for (uint32_t s = b0; s < b1; ++s) mine += hist[s];
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) s_scan[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        uint32_t v = lane < REFINE_THREADS / 32 ? s_scan[lane] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane < REFINE_THREADS / 32) s_scan[lane] = inc - v;
      }
      __syncthreads();
      uint32_t run = s_scan[warp] + (incl - mine);
      for (uint32_t s = b0; s < b1; ++s) {
        const uint32_t h = hist[s];
        s_off[s] = run;
        g_base[s] = h ? (uint32_t)(P.fine_start[f0 + s] + atomicAdd(P.fine_cursor + f0 + s, (unsigned long long)h)) : 0u;
        hist[s] = run;  // becomes the staging cursor
        run += h;
      }
      __syncthreads();
      for (uint32_t i0 = 0; i0 < m; i0 += U * REFINE_THREADS) {  // second read of the tile comes from L2
        uint64_t key[U], cnt[U];
        uint32_t sb[U], o[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const uint32_t i = i0 + j * REFINE_THREADS + tid;
          key[j] = i < m ? __ldcs(P.keys + begin + i) : EMPTY_KEY;  // last use of this tile
          cnt[j] = (i < m && P.counts) ? __ldcs(P.counts + begin + i) : 1ull;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) { sb[j] = sub_of_mix(mix64(key[j]), P.n_sub); o[j] = 0; if (i0 + j * REFINE_THREADS + tid < m) o[j] = atomicAdd(hist + sb[j], 1u); }
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (i0 + j * REFINE_THREADS + tid < m) { staging[o[j]] = key[j]; if (P.counts) staging_c[o[j]] = cnt[j]; }
      }
      __syncthreads();
      for (uint32_t i = tid; i < m; i += REFINE_THREADS) {  // coalesced copy-out; destination recomputed from the key
        const uint64_t key = staging[i];
        const uint32_t sbin = sub_of_mix(mix64(key), P.n_sub);
        const uint64_t dst = (uint64_t)g_base[sbin] + (i - s_off[sbin]);
        P.out_keys[dst] = key;
        if (P.out_counts) P.out_counts[dst] = P.counts ? staging_c[i] : 1ull;
      }
      __syncthreads();
      for (uint32_t s = tid; s < P.n_sub; s += REFINE_THREADS) hist[s] = 0;
    }
    __syncthreads();
  }
}

cudaError_t launch_refine(const RefineParams &P, bool scatter, cudaStream_t s) {
  if (P.n_tiles == 0) return cudaSuccess;
  const size_t smem = scatter ? (size_t)REFINE_TILE * 8 * (P.counts ? 2 : 1) + 3 * (size_t)P.n_sub * sizeof(uint32_t) : (size_t)P.n_sub * sizeof(uint32_t);
  const unsigned grid = (unsigned)std::min<uint64_t>(P.n_tiles, (uint64_t)num_sms() * ((scatter && smem > 110 * 1024) ? 1 : 2));
  cudaError_t e;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (scatter) {
    if ((e = cudaFuncSetAttribute(refine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    refine_kernel<true><<<grid, REFINE_THREADS, smem, s>>>(P);
  } else {
    if ((e = cudaFuncSetAttribute(refine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    refine_kernel<false><<<grid, REFINE_THREADS, smem, s>>>(P);
  }
  return cudaGetLastError();
}

// per-partition totals over all input runs (for load-balanced ordering on the host)
__global__ void sum_lens_kernel(CountParams P, unsigned long long *totals) {
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n_parts; p += gridDim.x * blockDim.x) {
    unsigned long long t = 0;
    for (uint32_t r = 0; r < P.R; ++r) t += P.runs[r].seg_len[p];
    totals[p] = t;
  }
}
cudaError_t launch_sum_lens(const CountParams &P, unsigned long long *d_totals, cudaStream_t s) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  sum_lens_kernel<<<(unsigned)std::min<uint64_t>((P.n_parts + 255) / 256, (uint64_t)num_sms() * 8), 256, 0, s>>>(P, d_totals);
  return cudaGetLastError();
}

// Count-of-counts while compacting (src/histogram.rs:88-116 computes it from the finished map; here it falls out
// of phase B for free).  Must be called by all 32 lanes together.  Counts of 1 -- the bulk on most inputs -- are
// not recorded at all: hist[1] = distinct - everything else.  Equal counts within the warp are merged first, so
// a popular count value costs one atomic per warp; the lowest bins live in shared memory until the CTA retires.
__device__ __forceinline__ void hist_note(bool ok, unsigned long long cnt, uint32_t *s_hist, const CountParams &P, int lane) {
  const bool agg = ok && cnt > 1 && cnt < (unsigned long long)HIST_DENSE_BINS;
  uint32_t pending = __ballot_sync(0xffffffffu, agg);
  while (pending) {
    const int leader = __ffs(pending) - 1;
    const uint32_t v = __shfl_sync(0xffffffffu, (uint32_t)cnt, leader);
    const uint32_t same = __ballot_sync(0xffffffffu, agg && (uint32_t)cnt == v);
    if (lane == leader) {
      if (v < (uint32_t)HIST_CTA_BINS) atomicAdd(&s_hist[v], (uint32_t)__popc(same));
      else atomicAdd(P.hist + v, (unsigned long long)__popc(same));
    }
    pending &= ~same;
  }
  if (ok && cnt >= (unsigned long long)HIST_DENSE_BINS) {
    const unsigned long long o = atomicAdd(P.hist + HIST_DENSE_BINS, 1ull);
    if (o < P.hist_overflow_cap) P.hist_overflow[o] = cnt;
  }
}
__device__ __forceinline__ void hist_flush(const uint32_t *s_hist, const CountParams &P, int tid) {
  if (tid < HIST_CTA_BINS && s_hist[tid]) atomicAdd(P.hist + tid, (unsigned long long)s_hist[tid]);
}

// ---------------------------------------------------------------------------------------------------
// phase B: one CTA counts one fine partition at a time in its private scratch table (L2-resident: the whole
// grid's scratch is ~39 MB), then compacts it into the output run and hands the slots back clean.
// Slots are (key, occurrences-1): a new key costs one CAS, duplicates one more RED.
// The table cannot overflow by construction (capacity >= 1.6 x the partition's ENTRIES when small, and for
// oversized -- i.e. high-multiplicity -- partitions the number of DISTINCT keys is bounded by the hash-uniform
// share of a partition); should that bound ever be violated the kernel raises error_flag and the host retries
// with larger scratch tables.  Nothing is dropped silently.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(COUNT_THREADS, COUNT_CTAS_PER_SM) count_partitions_kernel(CountParams P) {
  __shared__ uint64_t seg_begin[CONS_MAX_RUNS];
  __shared__ uint64_t seg_prefix[CONS_MAX_RUNS + 1];
  __shared__ uint32_t s_work, s_warp[COUNT_THREADS / 32 + 1], s_hist[HIST_CTA_BINS];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < HIST_CTA_BINS) s_hist[tid] = 0;
  unsigned long long *table = reinterpret_cast<unsigned long long *>(P.scratch) + (uint64_t)blockIdx.x * (2ull << P.scratch_log2);
  uint32_t next_work = 0;
  if (tid == 0) next_work = atomicAdd(P.next, 1u);

  for (;;) {
    // ---- fetch a partition (the next one is drawn early so the atomic's latency hides behind this one's work)
    if (warp == 0) {
      uint32_t w = 0;
      if (lane == 0) { w = next_work; if (w < P.n_parts) next_work = atomicAdd(P.next, 1u); }
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w < P.n_parts) {
        const uint32_t p = P.order[w];
        uint64_t b = 0, len = 0;
        if (lane < (int)P.R) { b = P.runs[lane].seg_start[p]; len = P.runs[lane].seg_len[p]; }
        uint64_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane < (int)P.R) { seg_begin[lane] = b; seg_prefix[lane] = incl - len; }
        if (lane == (int)P.R - 1) seg_prefix[P.R] = incl;
      }
      if (lane == 0) s_work = w;
    }
    __syncthreads();
    const uint32_t work = s_work;
    if (work >= P.n_parts) break;
    const uint32_t p = P.order[work];
    const uint64_t n_p = seg_prefix[P.R];
    if (n_p == 0) {  // block-uniform
      if (tid == 0) { P.out_seg_start[p] = 0; P.out_seg_len[p] = 0; }
      __syncthreads();
      continue;
    }
    // capacity: 1.6x the entries, at least 256 slots, at most the scratch table
    uint32_t cap_log2 = 8;
    while (cap_log2 < P.scratch_log2 && (1ull << cap_log2) * 5 < n_p * 8) ++cap_log2;
    const uint64_t mask = (1ull << cap_log2) - 1;

    // ---- upsert all entries of partition p
    constexpr int G = 4;
    uint32_t new_keys = 0;
    uint32_t r = 0;  // run holding the thread's current entry: entries are visited in ascending order
#pragma unroll 1
    for (uint64_t base = 0; base < n_p; base += (uint64_t)COUNT_THREADS * G) {
      uint64_t key[G], w[G], slot[G], cur[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const uint64_t idx = base + (uint64_t)j * COUNT_THREADS + tid;
        w[j] = 0; key[j] = EMPTY_KEY;
        if (idx < n_p) {
          while (r + 1 < P.R && idx >= seg_prefix[r + 1]) ++r;
          const uint64_t src = seg_begin[r] + (idx - seg_prefix[r]);
          key[j] = __ldcs(P.runs[r].keys + src);
          w[j] = P.runs[r].counts ? __ldcs(P.runs[r].counts + src) : 1ull;
        }
      }
      if (P.preagg) {
        // Warp run-length pre-aggregation: the scatter passes keep the keys of consecutive windows close together,
        // so homopolymer / tandem-repeat runs tend to arrive as runs of equal keys in adjacent lanes.  The head lane
        // of each run upserts once with the run length; this bounds same-address atomic bursts on skewed inputs.
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
      uint32_t pend = 0;
#pragma unroll
      for (int j = 0; j < G; ++j) {  // first probes of all G keys in flight together
        slot[j] = mix64(key[j]) & mask;  // lowest mix bits; coarse / sub-bin used the top bits of each half
        cur[j] = 0;
        if (w[j]) cur[j] = atomicCAS(table + 2 * slot[j], EMPTY_KEY, key[j]);
      }
#pragma unroll
      for (int j = 0; j < G; ++j) {
        if (!w[j]) continue;
        if (cur[j] == EMPTY_KEY) { ++new_keys; if (w[j] > 1) atomicAdd(table + 2 * slot[j] + 1, (unsigned long long)(w[j] - 1)); }
        else if (cur[j] == key[j]) atomicAdd(table + 2 * slot[j] + 1, (unsigned long long)w[j]);
        else pend |= 1u << j;
      }
      uint64_t tries = 0;
      while (pend) {  // linear probing; all still-pending keys of the thread advance together (keeps MLP)
        if (++tries > mask) { atomicExch(P.error_flag, 1u); break; }
#pragma unroll
        for (int j = 0; j < G; ++j)
          if (pend >> j & 1u) { slot[j] = (slot[j] + 1) & mask; cur[j] = *reinterpret_cast<volatile unsigned long long *>(table + 2 * slot[j]); }
#pragma unroll
        for (int j = 0; j < G; ++j)
          if ((pend >> j & 1u) && cur[j] == EMPTY_KEY) cur[j] = atomicCAS(table + 2 * slot[j], EMPTY_KEY, key[j]);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          if (!(pend >> j & 1u)) continue;
          if (cur[j] == EMPTY_KEY) { ++new_keys; if (w[j] > 1) atomicAdd(table + 2 * slot[j] + 1, (unsigned long long)(w[j] - 1)); pend &= ~(1u << j); }
          else if (cur[j] == key[j]) { atomicAdd(table + 2 * slot[j] + 1, (unsigned long long)w[j]); pend &= ~(1u << j); }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) new_keys += __shfl_xor_sync(0xffffffffu, new_keys, o);
    if (lane == 0) s_warp[warp] = new_keys;
    __threadfence_block();
    __syncthreads();  // all upserts of this CTA are done (they are L2 atomics issued by this CTA only)
    if (tid == 0) {
      uint32_t d = 0;
      for (int w2 = 0; w2 < COUNT_THREADS / 32; ++w2) d += s_warp[w2];
      const unsigned long long b = atomicAdd(P.out_cursor, (unsigned long long)d);  // the partition's contiguous output range
      P.out_seg_start[p] = b; P.out_seg_len[p] = d;
      s_base = b;
    }
    __syncthreads();

    // ---- compact the table into the output run and clean it
    const unsigned long long out0 = s_base;
    uint32_t mine = 0;
    for (uint64_t i = tid; i <= mask; i += COUNT_THREADS) mine += __ldcg(table + 2 * i) != EMPTY_KEY;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t v = lane < COUNT_THREADS / 32 ? s_warp[lane] : 0, inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      if (lane < COUNT_THREADS / 32) s_warp[lane] = inc - v;
    }
    __syncthreads();
    uint64_t o = out0 + s_warp[warp] + (incl - mine);
    const ulonglong2 *tab2 = reinterpret_cast<const ulonglong2 *>(table);
    for (uint64_t i = tid; i <= mask; i += COUNT_THREADS) {  // trip count is warp-uniform (mask + 1 is a multiple of 256)
      const ulonglong2 sl = __ldcg(tab2 + i);
      const bool used = sl.x != EMPTY_KEY;
      if (used) {
        __stcs(P.out_keys + o, sl.x);
        __stcs(P.out_counts + o, sl.y + 1);  // slots store occurrences - 1
        ++o;
        reinterpret_cast<ulonglong2 *>(table)[i] = make_ulonglong2(EMPTY_KEY, 0ull);
      }
      if (P.hist) hist_note(used, sl.y + 1, s_hist, P, lane);
    }
    __syncthreads();  // table clean (same-CTA visibility) before the next partition's upserts
  }
  if (P.hist) hist_flush(s_hist, P, tid);
}


// ---------------------------------------------------------------------------------------------------
// phase B, primary variant: the partition's table lives in SHARED memory.  A probe step then costs tens
// of cycles instead of an L2 round trip, which is what gated the L2-scratch variant above (its CTA-wide
// time per partition was set by the longest probe chain).  Layout: SoA, 8192 x u64 keys + 8192 x u32
// (occurrences - 1) = 96 KiB, so TWO 512-thread CTAs fit per SM and one CTA's latency bubbles (metadata,
// key loads, output reservation) overlap with the other's work.  Output is compacted per warp (ballot +
// popc) so that a warp writes one contiguous run.
// If a partition holds more distinct keys than the table, or a count does not fit 32 bits, error_flag is
// raised and the host re-runs phase B with the L2-scratch variant (u64 counts, larger tables).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SMEM_COUNT_THREADS, 2) count_partitions_smem_kernel(CountParams P) {
  extern __shared__ __align__(16) unsigned long long skeys[];  // SMEM_TABLE_SLOTS keys, u32 counts, u16 occupied-slot list
  uint32_t *scnt = reinterpret_cast<uint32_t *>(skeys + SMEM_TABLE_SLOTS);
  uint16_t *slist = reinterpret_cast<uint16_t *>(scnt + SMEM_TABLE_SLOTS);  // slots claimed for this partition, in claim order
  __shared__ uint32_t s_list_n;
  __shared__ uint64_t seg_begin[CONS_MAX_RUNS];
  __shared__ uint64_t seg_prefix[CONS_MAX_RUNS + 1];
  __shared__ uint32_t s_work, s_hist[HIST_CTA_BINS];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t SLOTS = SMEM_TABLE_SLOTS;
  if (tid < HIST_CTA_BINS) s_hist[tid] = 0;
  for (uint32_t i = tid; i < SLOTS; i += SMEM_COUNT_THREADS) { skeys[i] = EMPTY_KEY; scnt[i] = 0; }
  uint32_t next_work = 0;
  if (tid == 0) { next_work = atomicAdd(P.next, 1u); s_list_n = 0; }
  __syncthreads();

  for (;;) {
    if (warp == 0) {
      uint32_t w = 0;
      if (lane == 0) { w = next_work; if (w < P.n_parts) next_work = atomicAdd(P.next, 1u); }
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w < P.n_parts) {
        const uint32_t p = P.order[w];
        uint64_t b = 0, len = 0;
        if (lane < (int)P.R) { b = P.runs[lane].seg_start[p]; len = P.runs[lane].seg_len[p]; }
        uint64_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane < (int)P.R) { seg_begin[lane] = b; seg_prefix[lane] = incl - len; }
        if (lane == (int)P.R - 1) seg_prefix[P.R] = incl;
      }
      if (lane == 0) s_work = w;
    }
    __syncthreads();
    const uint32_t work = s_work;
    if (work >= P.n_parts) break;
    const uint32_t p = P.order[work];
    const uint64_t n_p = seg_prefix[P.R];
    if (n_p == 0) {
      if (tid == 0) { P.out_seg_start[p] = 0; P.out_seg_len[p] = 0; }
      __syncthreads();
      continue;
    }
    uint32_t cap_log2 = 8;  // small partitions use a prefix of the table: less to compact
    while ((1u << cap_log2) < SLOTS && (1ull << cap_log2) * 5 < n_p * 8) ++cap_log2;
    const uint32_t mask = (1u << cap_log2) - 1;

    // 8 keys per thread are loaded up front (one exposed global-load latency per 4096 entries instead of two),
    // then upserted in two batches of 4 whose first probes are in flight together.
    constexpr int G = 8, H = 4;
    uint32_t r = 0;  // run holding the thread's current entry: entries are visited in ascending order
#pragma unroll 1
    for (uint64_t base = 0; base < n_p; base += (uint64_t)SMEM_COUNT_THREADS * G) {
      uint64_t key[G];
      uint32_t w[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const uint64_t idx = base + (uint64_t)j * SMEM_COUNT_THREADS + tid;
        w[j] = 0; key[j] = EMPTY_KEY;
        if (idx < n_p) {
          while (r + 1 < P.R && idx >= seg_prefix[r + 1]) ++r;
          const uint64_t src = seg_begin[r] + (idx - seg_prefix[r]);
          key[j] = __ldcs(P.runs[r].keys + src);
          uint64_t w64 = 1;
          if (P.runs[r].counts) w64 = __ldcs(P.runs[r].counts + src);
          if (w64 > 0xffffffffull) { atomicExch(P.error_flag, 1u); w64 = 0; }  // needs the u64 (L2-scratch) variant
          w[j] = (uint32_t)w64;
        }
      }
      if (P.preagg && n_p > 2 * SMEM_COUNT_THREADS * G) {  // warp run-length pre-aggregation, only for oversized (= skewed) partitions
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const uint64_t kk = w[j] ? key[j] : EMPTY_KEY;
          const uint64_t kp = __shfl_up_sync(0xffffffffu, kk, 1);
          const bool head = lane == 0 || kp != kk;
          const uint32_t heads = __ballot_sync(0xffffffffu, head);
          if (__all_sync(0xffffffffu, w[j] <= 1u)) {
            const uint32_t above = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
            const uint32_t end = above ? (uint32_t)__ffs(above) - 1u : 32u;
            if (w[j]) w[j] = head ? end - lane : 0u;
          }
        }
      }
      // First probes of all G keys, H at a time in flight; whatever collides is left in `pend`.
      uint32_t pend = 0, newm = 0;  // bit j: key j still to place / key j claimed a fresh slot
      uint64_t fs_lo = 0, fs_hi = 0;  // slot of key j, 16 bits each (keys 0-3 / 4-7), meaningful for the keys in newm
#pragma unroll
      for (int h0 = 0; h0 < G; h0 += H) {
        uint32_t sl[H];
        unsigned long long cur[H];
#pragma unroll
        for (int j = 0; j < H; ++j) { sl[j] = (uint32_t)mix64(key[h0 + j]) & mask; cur[j] = w[h0 + j] ? skeys[sl[j]] : 0ull; }
#pragma unroll
        for (int j = 0; j < H; ++j) if (w[h0 + j] && cur[j] == EMPTY_KEY) cur[j] = atomicCAS(&skeys[sl[j]], EMPTY_KEY, key[h0 + j]);
#pragma unroll
        for (int j = 0; j < H; ++j) {
          const uint32_t wj = w[h0 + j];
          if (!wj) continue;
          const bool is_new = cur[j] == EMPTY_KEY;
          if (is_new || cur[j] == key[h0 + j]) {
            const uint32_t add = wj - (is_new ? 1u : 0u);  // slots store occurrences - 1
            if (add) { const uint32_t old = atomicAdd(&scnt[sl[j]], add); if (old > 0xffffffffu - add) atomicExch(P.error_flag, 1u); }
            if (is_new) { newm |= 1u << (h0 + j); (h0 + j < 4 ? fs_lo : fs_hi) |= (uint64_t)sl[j] << (16 * ((h0 + j) & 3)); }
          } else pend |= 1u << (h0 + j);
        }
      }
      // Collisions: every lane works through ITS pending keys on its own (no warp-wide rendezvous per key, so the
      // warp runs for the longest per-lane total, not for the sum of the per-key maxima).  Double hashing -- an odd
      // stride from the high mix bits -- keeps the chains short; lookups only ever happen through this same sequence.
      while (pend) {
        const int j = __ffs(pend) - 1;
        uint64_t kj = key[0];
        uint32_t wj = w[0];
#pragma unroll
        for (int q = 1; q < G; ++q) if (j == q) { kj = key[q]; wj = w[q]; }
        const uint64_t m = mix64(kj);
        const uint32_t step = (uint32_t)(m >> 40) | 1u;
        uint32_t s2 = (uint32_t)m & mask;
        unsigned long long c2;
        for (uint32_t tries = 0;; ++tries) {
          s2 = (s2 + step) & mask;
          c2 = skeys[s2];
          if (c2 == EMPTY_KEY) c2 = atomicCAS(&skeys[s2], EMPTY_KEY, kj);
          if (c2 == EMPTY_KEY || c2 == kj) break;
          if (tries > mask) { atomicExch(P.error_flag, 1u); break; }  // table full: the host re-runs with the L2 variant
        }
        const bool is_new = c2 == EMPTY_KEY;
        if (is_new || c2 == kj) {
          const uint32_t add = wj - (is_new ? 1u : 0u);
          if (add) { const uint32_t old = atomicAdd(&scnt[s2], add); if (old > 0xffffffffu - add) atomicExch(P.error_flag, 1u); }
          if (is_new) {
            newm |= 1u << j;
            const uint64_t f = (uint64_t)s2 << (16 * (j & 3));
            if (j < 4) fs_lo |= f; else fs_hi |= f;
          }
        }
        pend &= pend - 1;
      }
      // remember which slots this partition claimed -- ONE warp-aggregated append per batch; compaction visits only those
      {
        const uint32_t mine = (uint32_t)__popc(newm);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {  // warp-uniform
          uint32_t lb = 0;
          if (lane == 31) lb = atomicAdd(&s_list_n, total);
          lb = __shfl_sync(0xffffffffu, lb, 31) + incl - mine;
#pragma unroll
          for (int q = 0; q < G; ++q)
            if (newm >> q & 1u) slist[lb++] = (uint16_t)((q < 4 ? fs_lo : fs_hi) >> (16 * (q & 3)));
        }
      }
    }
    __syncthreads();  // all upserts done; s_list_n = number of distinct keys of this partition
    // ---- compact: reserve the partition's output range with one global atomic, then walk the claimed-slot list:
    // entry i goes to out[base + i] (perfectly coalesced) and its slot is handed back clean.
    if (tid == 0) {
      const uint32_t d = s_list_n;
      const unsigned long long b = atomicAdd(P.out_cursor, (unsigned long long)d);
      P.out_seg_start[p] = b; P.out_seg_len[p] = d;
      s_base = b;
    }
    __syncthreads();
    {
      const unsigned long long out0 = s_base;
      const uint32_t d = s_list_n;
      for (uint32_t i0 = 0; i0 < d; i0 += SMEM_COUNT_THREADS) {
        const uint32_t i = i0 + tid;
        const bool ok = i < d;
        unsigned long long cnt = 0;
        if (ok) {
          const uint32_t slot = slist[i];
          cnt = (unsigned long long)scnt[slot] + 1;  // slots store occurrences - 1
          __stcs(P.out_keys + out0 + i, (uint64_t)skeys[slot]);
          __stcs(P.out_counts + out0 + i, (uint64_t)cnt);
          skeys[slot] = EMPTY_KEY; scnt[slot] = 0;
        }
        if (P.hist) hist_note(ok, cnt, s_hist, P, lane);
      }
    }
    __syncthreads();  // table clean before the next partition
    if (tid == 0) s_list_n = 0;
  }
  if (P.hist) hist_flush(s_hist, P, tid);
}

cudaError_t launch_count_partitions_smem(const CountParams &P, cudaStream_t s) {
  if (P.n_parts == 0) return cudaSuccess;
  const size_t smem = (size_t)SMEM_TABLE_SLOTS * 14;  // u64 keys + u32 counts + u16 slot list
  cudaError_t e = cudaFuncSetAttribute(count_partitions_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const unsigned grid = (unsigned)std::min<uint64_t>(P.n_parts, (uint64_t)num_sms() * 2);
  count_partitions_smem_kernel<<<grid, SMEM_COUNT_THREADS, smem, s>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_count_partitions(const CountParams &P, unsigned grid, cudaStream_t s) {
  if (P.n_parts == 0) return cudaSuccess;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  count_partitions_kernel<<<grid, COUNT_THREADS, 0, s>>>(P);
  return cudaGetLastError();
}

}  // namespace kmg
