// Sorting one interval set into the reference's order and the scratch it needs.
//   sort_rows(): keys (keys.cuh) -> stable onesweep radix argsort on the
//   (group, start) bits -> tie fix-up so rows sharing (group, start) come in
//   (end', row) order == np.lexsort([ends, starts]) per group
//   (reference src/bioframe/core/arrops.py:332-335, 459-460).
#pragma once
#include "common.cuh"
#include "keys.cuh"
#include "lookback.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace bf {

// Buffers of one sorted set. tmp8/tmp4a/tmp4b are general per-row scratch that
// the tie fix-up borrows and later stages reuse (scan offsets, lo positions...).
struct SortBufs {
  u64 *keyA, *keyB;
  u32 *idA, *idB;
  u64 *tmp8;   // [n+1]
  u32 *tmp4a;  // [n]
  u32 *tmp4b;  // [n]
  ScanState ss;
  u64 *nT;     // device-side count of tied rows
};
static inline void take_sort_bufs(Arena& a, i64 n, SortBufs* b) {
  b->keyA = a.take<u64>(n);
  b->keyB = a.take<u64>(n);
  b->idA = a.take<u32>(n);
  b->idB = a.take<u32>(n);
  b->tmp8 = a.take<u64>(n + 1);
  b->tmp4a = a.take<u32>(n);
  b->tmp4b = a.take<u32>(n);
  b->ss = take_scan_state(a, (u64)scan_tiles((u64)n));
  b->nT = a.take<u64>(1);
}

// ---- tie fix-up: rows sharing (group, start) into (end', row) order ----
// After the stable sort on cs the rows of a tie run are in row order; the
// reference wants them by (end', row).  Tied rows are rare, so: count them,
// compact (key, id, slot), sort that subset by the FULL key (stable, so equal
// keys keep row order) and write it back into the same slots in order.
// One pass: stage the tile's keys (+ one halo key each side) in shared memory,
// flag tied rows, scan the flags (tile scan + look-back) and write the subset.
static __global__ void __launch_bounds__(SC_THREADS)
    tie_compact_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n, KeyLayout L, ScanState ss,
                       u64* __restrict__ tkeys, u32* __restrict__ tids, u32* __restrict__ tpos,
                       u64* __restrict__ nT) {
  __shared__ u64 s_k[SC_TILE + 2];
  __shared__ u64 s_tmp[8];
  __shared__ u64 s_prefix;
  __shared__ u32 s_ticket;
  const u32 tid = threadIdx.x;
  const u64 tile = take_ticket(ss.ticket, &s_ticket);
  const u64 base = tile * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  for (u32 i = tid; i < rows + 2; i += SC_THREADS) {
    const i64 p = (i64)base + i - 1;
    // halo outside the array: a key whose cs differs from every neighbour's
    s_k[i] = (p >= 0 && (u64)p < n) ? keys[p] : ~0ull;
  }
  __syncthreads();
  const bool lo_edge = base == 0, hi_edge = base + rows == n;
  u32 mask = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u32 r = tid * SC_IPT + j;  // row inside the tile, s_k index r + 1
    if (r < rows) {
      const u64 c = key_cs(s_k[r + 1], L);
      const bool prev_ok = !(lo_edge && r == 0);
      const bool next_ok = !(hi_edge && r + 1 == rows);
      if ((prev_ok && key_cs(s_k[r], L) == c) || (next_ok && key_cs(s_k[r + 2], L) == c)) mask |= 1u << j;
    }
  }
  u64 tot;
  u64 at = block_excl_sum_u64((u64)__popc(mask), s_tmp, &tot);
  if (tid < 32) {
    const u64 pre = lookback_exclusive(ss.status, tile, tot);
    if (tid == 0) s_prefix = pre;
  }
  __syncthreads();
  at += s_prefix;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    if (mask & (1u << j)) {
      const u32 r = tid * SC_IPT + j;
      tkeys[at] = s_k[r + 1];
      tids[at] = ids[base + r];
      tpos[at] = (u32)(base + r);
      ++at;
    }
  }
  if (hi_edge && tid == 0) *nT = s_prefix + tot;
}
static __global__ void __launch_bounds__(256) tie_scatter_kernel(const u64* __restrict__ tkeys, const u32* __restrict__ tids,
                                                          const u32* __restrict__ tpos, const u64* __restrict__ nT,
                                                          u64* __restrict__ keys, u32* __restrict__ ids) {
  const u64 n = *nT;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
    const u32 p = tpos[t];
    keys[p] = tkeys[t];
    ids[p] = tids[t];
  }
}

// ---------------------------------------------------------------- host side

// Sort rows (grp may be null for a single group) into res->keys / res->vals.
static int sort_rows(const i32* grp, const i64* s, const i64* e, i64 n, const KeyLayout& L, int point_adjust,
                     const SortBufs& ws, const RadixScratch& sc, cudaStream_t st, SortResult* res) {
  res->keys = ws.keyA;
  res->vals = ws.idA;
  res->keys_alt = ws.keyB;
  res->vals_alt = ws.idB;
  if (n == 0) return BF_OK;
  // main sort: stable, on the (group, start) bits only
  const RadixPlan rp = make_radix_plan(L.LB, L.total_bits);
  BF_CUDA(cudaMemsetAsync(sc.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
  BF_TRY(launch_pack_keys(grp, s, e, (u64)n, L, point_adjust, rp, ws.keyA, sc.hist, st));
  BF_TRY(radix_sort_passes<true>(ws.keyA, ws.keyB, ws.idA, ws.idB, true, (u64)n, nullptr, rp, sc, st, res));
  if (point_adjust & KEY_NO_TIE_FIX) return BF_OK;
  // tie fix-up on the subset of rows that share (group, start); scratch = the
  // sort's spare buffers and the per-row temporaries
  const unsigned tiles = (unsigned)scan_tiles((u64)n);
  u64* tkeyA = res->keys_alt;
  u32* tidA = res->vals_alt;
  u64* tkeyB = ws.tmp8;
  u32* tidB = ws.tmp4a;
  u32* tpos = ws.tmp4b;
  BF_TRY(reset_scan_state(ws.ss, tiles, st));
  BF_LAUNCH(PC_TIEFIX, 8 * n, st,
            tie_compact_kernel<<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, L, ws.ss, tkeyA, tidA, tpos, ws.nT));
  const RadixPlan full = make_radix_plan(0, L.total_bits);
  BF_CUDA(cudaMemsetAsync(sc.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
  BF_TRY(launch_radix_hist(tkeyA, (u64)n, ws.nT, full, sc.hist, st, PC_TIEFIX));
  SortResult tr;
  BF_TRY(radix_sort_passes<true>(tkeyA, tkeyB, tidA, tidB, false, (u64)n, ws.nT, full, sc, st, &tr, PC_TIEFIX));
  BF_LAUNCH(PC_TIEFIX, 0, st,
            tie_scatter_kernel<<<stream_grid((u64)n, 256 * 64), 256, 0, st>>>(tr.keys, tr.vals, tpos, ws.nT, res->keys, res->vals));
  return BF_OK;
}

// seg[c] = first sorted position whose group rank is >= c   (c = 0..nb, seg[nb] = n)
static __global__ void segment_table_kernel(const u64* __restrict__ keys, i64 n, KeyLayout L, i32 nb,
                                     i64* __restrict__ seg) {
  i32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nb) return;
  if (c == nb) {
    seg[c] = n;
    return;
  }
  const u64 first_cs = L.goff[c];  // rows of groups >= c start at this linear coordinate
  seg[c] = lower_bound_by(0, n, [&](i64 m) { return key_cs(keys[m], L) < first_cs; });
}

}  // namespace bf
