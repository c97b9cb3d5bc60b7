// Design-guidance microbenchmarks for the B200 (run under gpurun; results go to profiles/).
//   1. random 64-bit atomics (CAS key + RED count on one 16-byte slot) vs table size: the honest
//      "atomic ceiling" next to the HBM-bytes roofline
//   2. random RED.ADD.u64 only, random 16-byte loads
//   3. scattered 8-byte stores into P append streams (the partitioning pass of a 2-phase design)
//   4. shared-memory atomic histogram rate
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}

template <int MODE, int MLP>  // 0: CAS+RED, 1: RED only, 2: 16B load, 3: CAS only
__global__ void rand_access(unsigned long long *tab, uint64_t cap, uint64_t n, uint64_t seed, unsigned long long *sink) {
  uint64_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride * MLP) {
    uint64_t slot[MLP], key[MLP], old[MLP];
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      key[j] = mix64(seed + i + j * stride);
      slot[j] = __umul64hi(key[j], cap);
      key[j] |= 1;
    }
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      if (i + j * stride >= n) { old[j] = 0; continue; }
      if (MODE == 0 || MODE == 3) old[j] = atomicCAS(tab + 2 * slot[j], ~0ull, key[j]);
      else if (MODE == 1) { atomicAdd(tab + 2 * slot[j] + 1, 1ull); old[j] = 0; }
      else { ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(tab + 2 * slot[j]); old[j] = v.x + v.y; }
    }
#pragma unroll
    for (int j = 0; j < MLP; ++j) {
      if (i + j * stride >= n) continue;
      if (MODE == 0) atomicAdd(tab + 2 * slot[j] + 1, 1ull);
      acc += old[j];
    }
  }
  if (acc == 0x1234567) *sink = acc;
}

__global__ void fill(unsigned long long *p, uint64_t n, unsigned long long v) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// scatter: each CTA processes tiles of TILE keys; per tile histogram in smem, reserve, scatter
template <int TILE>
__global__ void scatter_parts(const uint64_t *__restrict__ keys, uint64_t n, uint32_t P, unsigned long long *cursor,
                              uint64_t *out, uint64_t part_cap) {
  extern __shared__ uint32_t sm[];
  uint32_t *hist = sm;
  uint64_t *base = reinterpret_cast<uint64_t *>(sm + P + (P & 1));
  for (uint64_t t0 = (uint64_t)blockIdx.x * TILE; t0 < n; t0 += (uint64_t)gridDim.x * TILE) {
    for (uint32_t p = threadIdx.x; p < P; p += blockDim.x) hist[p] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < TILE && t0 + i < n; i += blockDim.x) {
      uint32_t p = (uint32_t)(((mix64(keys[t0 + i]) >> 32) * P) >> 32);
      atomicAdd(hist + p, 1u);
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < P; p += blockDim.x) {
      uint32_t c = hist[p];
      base[p] = (uint64_t)p * part_cap + (c ? atomicAdd(cursor + p, (unsigned long long)c) : 0);
      hist[p] = 0;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < TILE && t0 + i < n; i += blockDim.x) {
      uint64_t k = keys[t0 + i];
      uint32_t p = (uint32_t)(((mix64(k) >> 32) * P) >> 32);
      uint32_t o = atomicAdd(hist + p, 1u);
      out[base[p] + o] = k;
    }
    __syncthreads();
  }
}

__global__ void gen_keys(uint64_t *keys, uint64_t n, uint64_t seed) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) keys[i] = mix64(seed + i);
}

__global__ void smem_hist(const uint64_t *__restrict__ keys, uint64_t n, uint32_t bins, unsigned long long *sink) {
  extern __shared__ uint32_t sm[];
  for (uint32_t p = threadIdx.x; p < bins; p += blockDim.x) sm[p] = 0;
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(sm + (uint32_t)(((keys[i] >> 32) * bins) >> 32), 1u);
  __syncthreads();
  if (threadIdx.x == 0 && sm[0] == 0xffffffffu) *sink = 1;
}

__global__ void copy64(const uint64_t *__restrict__ a, uint64_t *__restrict__ b, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main(int argc, char **argv) {
  const bool only_scatter = argc > 1 && argv[1][0] == 's';
  const int nlog = argc > 2 ? atoi(argv[2]) : 28;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d}\n", prop.name, sms, prop.l2CacheSize);
  unsigned long long *sink; CK(cudaMalloc(&sink, 8));
  const uint64_t N = 1ull << 29;
  const int grid = sms * 8, threads = 256;
  const double sizes_mb[] = {16, 32, 48, 64, 96, 128, 256, 1024, 8192, 65536};
  for (double mb : sizes_mb) {
    if (only_scatter) break;
    uint64_t cap = (uint64_t)(mb * 1048576.0 / 16.0);
    unsigned long long *tab;
    if (cudaMalloc(&tab, cap * 16) != cudaSuccess) { printf("{\"skip_mb\": %.0f}\n", mb); cudaGetLastError(); continue; }
    auto reset = [&]() { fill<<<grid, threads>>>(tab, cap * 2, ~0ull); };
    // MODE 0: first-touch inserts (table mostly empty -> CAS succeeds): n = min(N, cap/2) so load stays <= 0.5
    uint64_t n_ins = cap / 2 < N ? cap / 2 : N;
    reset(); CK(cudaDeviceSynchronize());
    float ms = time_ms([&]() { rand_access<0, 8><<<grid, threads>>>(tab, cap, n_ins, 1, sink); }, 1);
    printf("{\"test\": \"cas+red first-touch\", \"table_mb\": %.0f, \"ops\": %llu, \"ms\": %.3f, \"gops\": %.2f}\n", mb, (unsigned long long)n_ins, ms, n_ins / ms / 1e6);
    float ms1 = time_ms([&]() { rand_access<1, 8><<<grid, threads>>>(tab, cap, N, 7, sink); });
    printf("{\"test\": \"red.add.u64 random\", \"table_mb\": %.0f, \"ops\": %llu, \"ms\": %.3f, \"gops\": %.2f}\n", mb, (unsigned long long)N, ms1, N / ms1 / 1e6);
    float ms3 = time_ms([&]() { rand_access<3, 8><<<grid, threads>>>(tab, cap, N, 9, sink); });
    printf("{\"test\": \"cas random (mostly occupied)\", \"table_mb\": %.0f, \"ops\": %llu, \"ms\": %.3f, \"gops\": %.2f}\n", mb, (unsigned long long)N, ms3, N / ms3 / 1e6);
    float ms2 = time_ms([&]() { rand_access<2, 8><<<grid, threads>>>(tab, cap, N, 11, sink); });
    printf("{\"test\": \"ld.16B random\", \"table_mb\": %.0f, \"ops\": %llu, \"ms\": %.3f, \"gops\": %.2f}\n", mb, (unsigned long long)N, ms2, N / ms2 / 1e6);
    float ms4 = time_ms([&]() { rand_access<0, 8><<<grid, threads>>>(tab, cap, N, 13, sink); });
    printf("{\"test\": \"cas+red random (steady)\", \"table_mb\": %.0f, \"ops\": %llu, \"ms\": %.3f, \"gops\": %.2f}\n", mb, (unsigned long long)N, ms4, N / ms4 / 1e6);
    fflush(stdout);
    cudaFree(tab);
  }
  // scatter into P streams
  {
    const uint64_t n = 1ull << nlog;
    uint64_t *keys, *out;
    CK(cudaMalloc(&keys, n * 8));
    gen_keys<<<grid, threads>>>(keys, n, 99);
    float msc = 0;
    {
      uint64_t *tmp; CK(cudaMalloc(&tmp, n * 8));
      msc = time_ms([&]() { copy64<<<grid, threads>>>(keys, tmp, n); });
      printf("{\"test\": \"copy 8B/key\", \"keys\": %llu, \"ms\": %.3f, \"gkeys\": %.2f, \"gbs\": %.1f}\n", (unsigned long long)n, msc, n / msc / 1e6, 16.0 * n / msc / 1e6);
      cudaFree(tmp);
    }
    for (uint32_t P : {8u, 64u, 512u, 2048u, 4096u, 8192u}) {
      uint64_t part_cap = (uint64_t)((double)n / P * 1.2) + 4096;
      CK(cudaMalloc(&out, P * part_cap * 8));
      unsigned long long *cursor; CK(cudaMalloc(&cursor, P * 8));
      size_t smem = (P + (P & 1)) * 4 + P * 8;
      CK(cudaFuncSetAttribute(scatter_parts<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaFuncSetAttribute(scatter_parts<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_ms([&]() { cudaMemset(cursor, 0, P * 8); scatter_parts<8192><<<sms * 4, 512, smem>>>(keys, n, P, cursor, out, part_cap); });
      float msb = time_ms([&]() { cudaMemset(cursor, 0, P * 8); scatter_parts<32768><<<sms * 4, 512, smem>>>(keys, n, P, cursor, out, part_cap); });
      printf("{\"test\": \"scatter\", \"P\": %u, \"keys\": %llu, \"ms_tile8k\": %.3f, \"gkeys_tile8k\": %.2f, \"ms_tile32k\": %.3f, \"gkeys_tile32k\": %.2f}\n",
             P, (unsigned long long)n, ms, n / ms / 1e6, msb, n / msb / 1e6);
      fflush(stdout);
      cudaFree(out); cudaFree(cursor);
    }
    for (uint32_t bins : {8u, 512u, 4096u}) {
      float ms = time_ms([&]() { smem_hist<<<sms * 8, 256, bins * 4>>>(keys, n, bins, sink); });
      printf("{\"test\": \"smem atomic hist\", \"bins\": %u, \"keys\": %llu, \"ms\": %.3f, \"gkeys\": %.2f}\n", bins, (unsigned long long)n, ms, n / ms / 1e6);
    }
    cudaFree(keys);
  }
  return 0;
}
