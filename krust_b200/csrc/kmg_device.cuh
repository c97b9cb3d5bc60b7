// Device-side building blocks shared by the sm_100a kernels of libkmerust_gpu.
// Nothing in here is derived from the reference's code; citations name the reference
// behaviour (paths relative to the kmerust repository) each helper must reproduce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmg {

constexpr uint64_t EMPTY_KEY = ~0ull;  // never a canonical key for any k (k<32: bits above 2k set;
                                       // k=32: TTT..T, whose reverse complement AAA..A = 0 is smaller)

// ---- packed-stream layout -------------------------------------------------------------------
// bases : u64 words, base j at word j/32, shift 62-2*(j%32)  (MSB first == src/kmer.rs:467-471 fold)
// valid : u32 words, base j at word j/32, bit 31-(j%32)
// start : same layout as valid; 1 on the first base of each record
// All three arrays carry a zero-filled lead-in so that "word -1" can be read unconditionally and
// TMA bulk copies stay 16-byte aligned.
constexpr int LEAD_BASE_WORDS = 2;  // 16 bytes
constexpr int LEAD_MASK_WORDS = 4;  // 16 bytes
constexpr int TILE_WORDS = 1024;    // 32768 bases per tile
constexpr int SCAN_THREADS = 256;
constexpr int WORDS_PER_THREAD = TILE_WORDS / SCAN_THREADS;

__host__ __device__ __forceinline__ uint64_t kmer_mask(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// murmur3 finaliser: bijective 64-bit mix.  Low bits pick the table slot, high bits the owner shard.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// inverse of mix64 (every step of the finaliser is a bijection: x ^= x >> 33 is an involution, the multipliers are odd)
__host__ __device__ __forceinline__ uint64_t unmix64(uint64_t x) {
  x ^= x >> 33; x *= 0x9cb4b2f8129337dbull;  // inverse of 0xc4ceb9fe1a85ec53 mod 2^64
  x ^= x >> 33; x *= 0x4f74430c22a54005ull;  // inverse of 0xff51afd7ed558ccd mod 2^64
  x ^= x >> 33;
  return x;
}
// The partitioned pipeline carries MIXED keys (v = mix64(key)) through its runs: every stage needs hash bits, none needs the
// key itself, and the mix is a bijection, so equality of mixes is equality of keys; the key is recovered with unmix64 when
// a result leaves the table (export / compaction).  EMPTY_MIX = mix64(EMPTY_KEY) can therefore never be a stored value.
constexpr uint64_t EMPTY_MIX = 0x64b5720b4b825f21ull;

// Multiply-shift range reduction of a 32-bit hash to [0, n).  On the device this MUST be the __umulhi
// intrinsic: nvcc 12.9 miscompiled the equivalent 64-bit expression when it indexed a shared-memory
// atomic (the IMAD.HI term vanished and every key landed in bin 0; tools/test_partkeys.cu).
__host__ __device__ __forceinline__ uint32_t reduce32(uint32_t h, uint32_t n) {
#ifdef __CUDA_ARCH__
  return __umulhi(h, n);
#else
  return (uint32_t)(((uint64_t)h * (uint64_t)n) >> 32);
#endif
}

// owner shard / coarse partition of a canonical key among n parts: HIGH half of the mix.
__host__ __device__ __forceinline__ uint32_t part_of(uint64_t key, uint32_t n_parts) {
  return reduce32((uint32_t)(mix64(key) >> 32), n_parts);
}
// Two-level partition index used by the partitioned pipeline: coarse from the high half of the mix, sub-bin
// from the top bits of the LOW half (the per-partition table slot uses the lowest bits, so all three are
// independent).  fine = coarse * n_sub + sub.
__host__ __device__ __forceinline__ uint32_t coarse_of_mix(uint64_t m, uint32_t n_coarse) { return reduce32((uint32_t)(m >> 32), n_coarse); }
__host__ __device__ __forceinline__ uint32_t sub_of_mix(uint64_t m, uint32_t n_sub) { return reduce32((uint32_t)m, n_sub); }

#ifdef __CUDACC__
// slot in [0, cap) for arbitrary (not power-of-two) capacities: multiply-shift range reduction.
__device__ __forceinline__ uint64_t slot_of(uint64_t key, uint64_t cap) {
  return __umul64hi(mix64(key) * 0x9E3779B97F4A7C15ull, cap);
}

// reverse complement of the low 2k bits of x (complement = 3 - code, src/kmer.rs:36-47).
__device__ __forceinline__ uint64_t revcomp(uint64_t x, int k) {
  uint64_t y = __brevll(~x);                                                  // reverses bits, swaps within pairs
  y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);  // undo the in-pair swap
  return y >> (64 - 2 * k);
}

// bit b of the result is set iff bits b .. b+len-1 of v are all set (len in 1..32, bits above 63 read as 0).
__device__ __forceinline__ uint64_t run_and(uint64_t v, int len) {
  uint64_t r = v;
  int n = 1;
  for (int bit = 30 - __clz(len); bit >= 0; --bit) {  // bits of len below its top bit
    r &= r >> n; n <<= 1;
    if ((len >> bit) & 1) { r &= v >> n; n += 1; }
  }
  return r;
}

// 32-bit mask (bit 31-e <-> position e of the current word) of windows ENDING at position e that
// may be counted: all k bases valid (ACGT + quality, src/run.rs:543-560) and no record start
// strictly inside the window (records are processed separately, src/run.rs:500-503).
__device__ __forceinline__ uint32_t window_ok_mask(uint32_t vprev, uint32_t vcur, uint32_t sprev, uint32_t scur,
                                                   int k, bool has_start) {
  uint64_t v = ((uint64_t)vprev << 32) | vcur;
  uint64_t ok = run_and(v, k);
  if (has_start && k > 1) {
    uint64_t ns = ~(((uint64_t)sprev << 32) | scur);
    ok &= run_and(ns, k - 1);
  }
  return (uint32_t)ok;
}

// ---- TMA (bulk async copy) + mbarrier, raw PTX -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- shared-memory accesses by 32-bit shared-space address ----------------------------------------
// The scatter kernels' inner loops address their shared-memory rows / counters through addresses computed ONCE per kernel
// (smem_u32).  Through generic pointers derived from the extern array the compiler re-derived the CTA's shared window
// (S2R SR_CgaCtaId + LEA) inside every predicated region, i.e. per key, and wrapped every conditional access into a
// BSSY / BRA / BSYNC region; the forms below are single predicated instructions.
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// if (pred) shared[addr] = v
__device__ __forceinline__ void sts64_if(bool pred, uint32_t addr, uint64_t v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u64 [%0], %1; }" ::"r"(addr), "l"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void sts32_if(bool pred, uint32_t addr, uint32_t v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u32 [%0], %1; }" ::"r"(addr), "r"(v), "r"((uint32_t)pred) : "memory");
}
// if (a < b) *dst = shared[addr]   (one predicated 8-byte load + store; t is set first so that it is not live around the caller's loop; STREAM: st.global.cs, the line is not re-read soon)
template <bool STREAM>
__device__ __forceinline__ void copy64_if_lt(uint32_t a, uint32_t b, uint64_t *dst, uint32_t addr) {
  if (STREAM)
    asm volatile("{ .reg .pred p; .reg .b64 t; setp.lt.u32 p, %2, %3; mov.b64 t, 0; @p ld.shared.b64 t, [%0]; @p st.global.cs.b64 [%1], t; }" ::"r"(addr), "l"(dst),
                 "r"(a), "r"(b)
                 : "memory");
  else
    asm volatile("{ .reg .pred p; .reg .b64 t; setp.lt.u32 p, %2, %3; mov.b64 t, 0; @p ld.shared.b64 t, [%0]; @p st.global.b64 [%1], t; }" ::"r"(addr), "l"(dst),
                 "r"(a), "r"(b)
                 : "memory");
}
#endif  // __CUDACC__

}  // namespace kmg
