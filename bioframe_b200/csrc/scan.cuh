// Block-level scan helpers and the spine scan.  Device-wide prefix sums / running maxima are done as
//   producer kernel: in-tile prefix + per-tile total  ->  spine_scan_kernel (one block)  ->  consumers add
//   spine[tile] to the in-tile value
// (sizes the variable-length outputs: np.repeat/cumsum of arrops.py:122-127, np.cumsum of arrops.py:472,
// np.maximum.accumulate of arrops.py:462).  A tile is 2048 elements.  No kernel spins on another one,
// except the radix pass (radix_sort.cuh) and two small closest kernels (lookback.cuh).
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

}  // namespace bf
