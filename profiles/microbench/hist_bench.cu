// Micro-benchmark: ways to histogram four 8-bit digit places of 1e8 random 64-bit keys
// (the primitive inside pack_keys_kernel and the onesweep ranking).  Build + run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/hist_bench profiles/microbench/hist_bench.cu && /tmp/hist_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int P = 4;  // digit places
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ u32 match8(u32 d) {
  u32 peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) { bool bit = (d >> b) & 1u; u32 m = __ballot_sync(0xffffffffu, bit); peers &= bit ? m : ~m; }
  return peers;
}
template <int MODE>
__global__ void __launch_bounds__(256) hist_kernel(const u64* __restrict__ keys, u64 n, u64* __restrict__ ghist) {
  // MODE 0: ballot match + leader plain add, per-warp tables
  // MODE 1: __match_any_sync + leader plain add, per-warp tables
  // MODE 2: shared atomicAdd, per-warp tables
  // MODE 3: shared atomicAdd, one table per block
  // MODE 4: no histogram (read-only roofline)
  extern __shared__ u32 tabs[];
  const int ntab = (MODE == 3) ? 1 : 8;
  for (int i = threadIdx.x; i < ntab * P * 256; i += blockDim.x) tabs[i] = 0;
  __syncthreads();
  u32* wtab = tabs + ((MODE == 3) ? 0 : (threadIdx.x >> 5) * P * 256);
  const u32 lt = lanemask_lt();
  u64 acc = 0;
  const u64 per_block = 256ull * 8ull;
  for (u64 base = (u64)blockIdx.x * per_block; base < n; base += (u64)gridDim.x * per_block) {
    u64 k[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { u64 i = base + (u64)j * 256 + threadIdx.x; k[j] = i < n ? keys[i] : 0; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 4) { acc += k[j]; continue; }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const u32 d = (u32)(k[j] >> (8 * p + 16)) & 255u;
        if (MODE == 0) { u32 peers = match8(d); if ((peers & lt) == 0) wtab[p * 256 + d] += __popc(peers); }
        if (MODE == 1) { u32 peers = __match_any_sync(0xffffffffu, d); if ((peers & lt) == 0) wtab[p * 256 + d] += __popc(peers); }
        if (MODE == 2 || MODE == 3) atomicAdd(&wtab[p * 256 + d], 1u);
      }
      if (MODE <= 1) __syncwarp();
    }
  }
  __syncthreads();
  if (MODE == 4) { if (acc == 0x1234567) ghist[0] = acc; return; }
  for (int i = threadIdx.x; i < P * 256; i += blockDim.x) {
    u32 c = 0;
    for (int w = 0; w < ntab; ++w) c += tabs[w * P * 256 + i];
    if (c) atomicAdd(&ghist[i], (u64)c);
  }
}
__global__ void fill(u64* k, u64 n, int sorted) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 x = i * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
  k[i] = sorted ? (i << 20) : x;  // sorted: high digits constant across a warp (worst case for atomics)
}
template <int MODE> void run(const char* name, const u64* keys, u64 n, u64* gh) {
  size_t smem = ((MODE == 3) ? 1 : 8) * P * 256 * 4;
  cudaFuncSetAttribute(hist_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaMemset(gh, 0, P * 256 * 8);
    cudaEventRecord(a);
    hist_kernel<MODE><<<148 * 8, 256, smem>>>(keys, n, gh);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  u64 h[4]; cudaMemcpy(h, gh, 32, cudaMemcpyDeviceToHost);
  printf("%-40s %.3f ms  %.0f GB/s  (check %llu)\n", name, best, n * 8 / best / 1e6, h[0] + h[1] + h[2] + h[3]);
}
int main() {
  const u64 n = 100000000ull;
  u64 *keys, *gh; cudaMalloc(&keys, n * 8); cudaMalloc(&gh, P * 256 * 8);
  for (int sorted = 0; sorted < 2; ++sorted) {
    fill<<<(unsigned)((n + 255) / 256), 256>>>(keys, n, sorted);
    printf("--- %s keys, %d digit places ---\n", sorted ? "sorted" : "random", P);
    run<4>("read only", keys, n, gh);
    run<0>("ballot match, per-warp tables", keys, n, gh);
    run<1>("match.any, per-warp tables", keys, n, gh);
    run<2>("smem atomics, per-warp tables", keys, n, gh);
    run<3>("smem atomics, per-block table", keys, n, gh);
  }
  return 0;
}
