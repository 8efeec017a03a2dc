// Single-pass device-wide prefix sums (decoupled look-back) used inside the
// producing kernels, so a scan costs no extra pass over the data:
// every block takes a tile by ticket, computes its tile total, publishes it,
// and warp 0 sums the totals of the preceding tiles (32 status words per probe)
// until it meets a tile whose inclusive prefix is already known.
// status word = flag (2 high bits: 1 = tile total, 2 = inclusive prefix) | value.
// The status array and the ticket must be zeroed before the launch.
#pragma once
#include "common.cuh"

namespace bf {

constexpr u64 LB_VAL_MASK = (1ull << 62) - 1;
constexpr u64 LB_AGG = 1ull << 62;
constexpr u64 LB_INC = 2ull << 62;

struct ScanState {
  u64* status;  // [tiles]
  u32* ticket;  // [1]
};
static inline ScanState take_scan_state(Arena& a, u64 tiles) {
  ScanState s;
  s.status = a.take<u64>(tiles + 1);
  s.ticket = a.take<u32>(1);
  return s;
}
static inline int reset_scan_state(const ScanState& s, u64 tiles, cudaStream_t st) {
  BF_CUDA(cudaMemsetAsync(s.status, 0, (tiles + 1) * sizeof(u64), st));
  BF_CUDA(cudaMemsetAsync(s.ticket, 0, sizeof(u32), st));
  return BF_OK;
}

// Block-uniform ticket; call from all threads. Contains a __syncthreads().
__device__ __forceinline__ u32 take_ticket(u32* ticket, u32* s_slot) {
  if (threadIdx.x == 0) *s_slot = atomicAdd(ticket, 1u);
  __syncthreads();
  return *s_slot;
}

// Called by ALL lanes of warp 0 with the tile total in `aggregate`; returns the
// sum of the totals of tiles [0, tile).
__device__ __forceinline__ u64 lookback_exclusive(u64* status, u64 tile, u64 aggregate) {
  const u32 lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(status, LB_INC | aggregate);
    return 0;
  }
  if (lane == 0) st_relaxed_u64(status + tile, LB_AGG | aggregate);
  u64 excl = 0;
  i64 base = (i64)tile - 1;
  for (;;) {
    const i64 t = base - (i64)lane;
    u64 w = LB_INC;  // tiles before 0: inclusive prefix 0
    if (t >= 0) w = ld_relaxed_u64(status + t);
    const u32 flag = (u32)(w >> 62);
    const u32 inc_mask = __ballot_sync(0xffffffffu, flag == 2);
    const u32 wait_mask = __ballot_sync(0xffffffffu, flag == 0);
    const u32 first_inc = inc_mask ? (u32)(__ffs(inc_mask) - 1) : 31u;
    const u32 needed = first_inc >= 31u ? 0xffffffffu : ((2u << first_inc) - 1u);
    if (wait_mask & needed) continue;  // a needed predecessor has not published yet
    u64 v = (lane <= first_inc) ? (w & LB_VAL_MASK) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (inc_mask) break;
    base -= 32;
  }
  if (lane == 0) st_relaxed_u64(status + tile, LB_INC | (excl + aggregate));
  return excl;
}

// Same protocol with max as the operator (running maximum across tiles).
__device__ __forceinline__ u64 lookback_exclusive_max(u64* status, u64 tile, u64 aggregate) {
  const u32 lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(status, LB_INC | aggregate);
    return 0;
  }
  if (lane == 0) st_relaxed_u64(status + tile, LB_AGG | aggregate);
  u64 excl = 0;
  i64 base = (i64)tile - 1;
  for (;;) {
    const i64 t = base - (i64)lane;
    u64 w = LB_INC;
    if (t >= 0) w = ld_relaxed_u64(status + t);
    const u32 flag = (u32)(w >> 62);
    const u32 inc_mask = __ballot_sync(0xffffffffu, flag == 2);
    const u32 wait_mask = __ballot_sync(0xffffffffu, flag == 0);
    const u32 first_inc = inc_mask ? (u32)(__ffs(inc_mask) - 1) : 31u;
    const u32 needed = first_inc >= 31u ? 0xffffffffu : ((2u << first_inc) - 1u);
    if (wait_mask & needed) continue;
    u64 v = (lane <= first_inc) ? (w & LB_VAL_MASK) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const u64 x = __shfl_xor_sync(0xffffffffu, v, o);
      v = x > v ? x : v;
    }
    excl = v > excl ? v : excl;
    if (inc_mask) break;
    base -= 32;
  }
  const u64 inc = excl > aggregate ? excl : aggregate;
  if (lane == 0) st_relaxed_u64(status + tile, LB_INC | inc);
  return excl;
}

}  // namespace bf
