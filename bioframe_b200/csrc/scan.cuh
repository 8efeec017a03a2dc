// Device-wide scans as reduce -> spine scan -> downsweep (no spin-waits):
//   exclusive sum of u32 counts into u64 offsets (sizes the variable-length
//   outputs; replaces np.repeat/cumsum in arrops.py:122-127 and np.cumsum in
//   arrops.py:472), and an inclusive running max of u64 (np.maximum.accumulate,
//   arrops.py:462).  A tile is 2048 elements; the producer kernel usually
//   writes the per-tile partials itself, so the inputs are read once more only
//   by the downsweep.
#pragma once
#include "common.cuh"

namespace bf {

constexpr int SC_THREADS = 256;
constexpr int SC_IPT = 8;
constexpr int SC_TILE = SC_THREADS * SC_IPT;  // 2048

static inline i64 scan_tiles(u64 n) { return ceil_div((i64)n, SC_TILE); }

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    u64 t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// Block-wide reductions for SC_THREADS threads; result valid in every thread.
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64* s_tmp /*[8]*/) {
  v = warp_sum_u64(v);
  if (lane_id() == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 t = 0;
#pragma unroll
  for (int w = 0; w < SC_THREADS / 32; ++w) t += s_tmp[w];
  __syncthreads();
  return t;
}
__device__ __forceinline__ u64 block_max_u64(u64 v, u64* s_tmp) {
  v = warp_max_u64(v);
  if (lane_id() == 0) s_tmp[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 t = 0;
#pragma unroll
  for (int w = 0; w < SC_THREADS / 32; ++w) t = s_tmp[w] > t ? s_tmp[w] : t;
  __syncthreads();
  return t;
}

// Exclusive prefix of per-thread totals across the block (sum). Returns the
// exclusive prefix for this thread; *block_total gets the block sum.
__device__ __forceinline__ u64 block_excl_sum_u64(u64 v, u64* s_tmp, u64* block_total) {
  const u32 lane = lane_id(), w = threadIdx.x >> 5;
  u64 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u64 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (u32)o) inc += t;
  }
  if (lane == 31) s_tmp[w] = inc;
  __syncthreads();
  u64 pre = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SC_THREADS / 32; ++i) {
    u64 x = s_tmp[i];
    if ((u32)i < w) pre += x;
    tot += x;
  }
  __syncthreads();
  *block_total = tot;
  return pre + inc - v;
}
__device__ __forceinline__ u64 block_excl_max_u64(u64 v, u64* s_tmp) {
  const u32 lane = lane_id(), w = threadIdx.x >> 5;
  u64 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u64 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (u32)o) inc = t > inc ? t : inc;
  }
  if (lane == 31) s_tmp[w] = inc;
  __syncthreads();
  u64 pre = 0;
#pragma unroll
  for (int i = 0; i < SC_THREADS / 32; ++i) {
    u64 x = s_tmp[i];
    if ((u32)i < w) pre = x > pre ? x : pre;
  }
  __syncthreads();
  u64 up = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) up = 0;
  return up > pre ? up : pre;  // max over all earlier threads (0 if none)
}

// ---- spine: one block turns per-tile partials into exclusive prefixes ----
// Each of the 32 warps owns a contiguous range and walks it 32 entries at a time (coalesced, no block
// barrier inside the loops): range totals -> scan of the 32 totals -> ranges rewritten as exclusive
// prefixes with a running carry.
template <bool IS_MAX>
static __global__ void __launch_bounds__(1024) spine_scan_kernel(u64* __restrict__ spine, u64 ntiles,
                                                                 u64* __restrict__ total_out) {
  __shared__ u64 s_warp[32];
  const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const u64 per = ((ntiles + 31) / 32 + 31) / 32 * 32;  // entries per warp, a multiple of 32
  const u64 lo = (u64)w * per;
  const u64 hi = lo + per < ntiles ? lo + per : ntiles;
  auto op = [](u64 a, u64 b) { return IS_MAX ? (a > b ? a : b) : a + b; };
  u64 acc = 0;
  for (u64 i = lo + lane; i < hi; i += 32) acc = op(acc, spine[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = op(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) s_warp[w] = acc;
  __syncthreads();
  if (w == 0) {
    const u64 x = s_warp[lane];
    u64 y = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= (u32)o) y = op(y, t);
    }
    u64 ex = __shfl_up_sync(0xffffffffu, y, 1);
    if (lane == 0) ex = 0;
    s_warp[lane] = ex;  // exclusive prefix of the warps
    if (lane == 31 && total_out) *total_out = y;
  }
  __syncthreads();
  u64 carry = s_warp[w];
  for (u64 base = lo; base < hi; base += 32) {
    const u64 i = base + lane;
    const u64 v = i < hi ? spine[i] : 0ull;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc = op(inc, t);
    }
    u64 up = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) up = 0;
    if (i < hi) spine[i] = op(carry, up);
    carry = op(carry, __shfl_sync(0xffffffffu, inc, 31));
  }
}

// ---- per-tile sums of a u32 array (when no producer wrote the spine) ----
static __global__ void __launch_bounds__(SC_THREADS) tile_sum_u32_kernel(const u32* __restrict__ in, u64 n,
                                                                  u64* __restrict__ spine) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    u64 i = base + (u64)j * SC_THREADS + threadIdx.x;
    if (i < n) acc += in[i];
  }
  acc = block_sum_u64(acc, s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = acc;
}

// ---- downsweep: out[i] = spine[tile] + sum(in[tile_start .. i)), out[n] = total ----
static __global__ void __launch_bounds__(SC_THREADS) excl_sum_u32_kernel(const u32* __restrict__ in, u64 n,
                                                                  const u64* __restrict__ spine,
                                                                  u64* __restrict__ out) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE + (u64)threadIdx.x * SC_IPT;
  u32 v[SC_IPT];
  u64 mine = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    u64 i = base + j;
    v[j] = i < n ? in[i] : 0u;
    mine += v[j];
  }
  u64 tot;
  u64 run = spine[blockIdx.x] + block_excl_sum_u64(mine, s_tmp, &tot);
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    u64 i = base + j;
    if (i < n) out[i] = run;
    run += v[j];
    if (i + 1 == n) out[n] = run;
  }
}

static int exclusive_sum_u32(const u32* counts, u64 n, u64* spine, bool spine_ready, u64* out,
                             cudaStream_t stream) {
  if (n == 0) {
    BF_CUDA(cudaMemsetAsync(out, 0, sizeof(u64), stream));
    return BF_OK;
  }
  const u64 tiles = (u64)scan_tiles(n);
  if (!spine_ready) {
    BF_LAUNCH(PC_SCAN, 0, stream, tile_sum_u32_kernel<<<(unsigned)tiles, SC_THREADS, 0, stream>>>(counts, n, spine));
  }
  BF_LAUNCH(PC_SCAN, 0, stream, spine_scan_kernel<false><<<1, 1024, 0, stream>>>(spine, tiles, nullptr));
  BF_LAUNCH(PC_SCAN, 0, stream, excl_sum_u32_kernel<<<(unsigned)tiles, SC_THREADS, 0, stream>>>(counts, n, spine, out));
  return BF_OK;
}

}  // namespace bf
