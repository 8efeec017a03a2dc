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
  u64 *spine;  // [scan tiles + 1] per-tile counts / offsets of rows in long tie runs
  u64 *nT;     // device-side count of rows in long tie runs
};
static inline void take_sort_bufs(Arena& a, i64 n, SortBufs* b) {
  b->keyA = a.take<u64>(n);
  b->keyB = a.take<u64>(n);
  b->idA = a.take<u32>(n);
  b->idB = a.take<u32>(n);
  b->tmp8 = a.take<u64>(n + 1);
  b->tmp4a = a.take<u32>(n);
  b->tmp4b = a.take<u32>(n);
  b->spine = a.take<u64>(scan_tiles((u64)n) + 2);
  b->nT = a.take<u64>(1);
}

// ---- tie fix-up: rows sharing (group, start) into (end', row) order ----
// After the stable sort on cs the rows of a tie run are in row order; the
// reference wants them by (end', row).  One pass over the sorted keys: a tile's
// keys (+ TIE_HALO keys each side) are staged in shared memory and every tied
// row measures its run.  Runs of up to TIE_SMALL rows (nearly all of them: two
// or three intervals that happen to start at the same base) are insertion-
// sorted in place by the thread that owns the run's first row -- stable, so
// equal keys keep row order.  Rows of longer runs (heavily duplicated input)
// are compacted as (key, id, slot), sorted by the FULL key with the stable
// radix sort and written back into the same slots in order.
constexpr int TIE_SMALL = 32;
constexpr int TIE_HALO = TIE_SMALL + 1;

// Stage a tile's keys (+ halo), measure the run of every tied row; with FIX,
// sort short runs in place.  Returns the mask of this thread's rows that belong
// to long runs.
// `e_raw` is null unless the layout clamps ends (keys.cuh measure_layout): then rows whose keys are
// equal may still differ in their original ends, which decide their order.
template <bool FIX>
__device__ __forceinline__ u32 tie_scan_tile(u64* __restrict__ keys, u32* __restrict__ ids, u64 n, const KeyLayout& L,
                                             u64 base, u64 rows, u64* s_k, const i64* __restrict__ e_raw) {
  auto after = [&](u64 k1, u32 v1, u64 k2, u32 v2) {  // (k1, row v1) sorts strictly after (k2, row v2)
    if (k1 != k2) return k1 > k2;
    return e_raw ? e_raw[v1] > e_raw[v2] : false;
  };
  const u32 tid = threadIdx.x;
  // s_k[i] = keys[base - TIE_HALO + i]; all loads of a thread in flight before its first store
  {
    u64 v[SC_IPT], hv = ~0ull;
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
      const u64 p = base + (u64)j * SC_THREADS + tid;
      v[j] = p < n ? keys[p] : ~0ull;
    }
    if (tid < 2 * TIE_HALO) {  // halo: TIE_HALO keys before the tile, TIE_HALO after
      const i64 p = tid < TIE_HALO ? (i64)base - TIE_HALO + tid : (i64)base + SC_TILE + (tid - TIE_HALO);
      if (p >= 0 && (u64)p < n) hv = keys[p];
      s_k[tid < TIE_HALO ? tid : SC_TILE + tid] = hv;
    }
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) s_k[TIE_HALO + j * SC_THREADS + tid] = v[j];
  }
  __syncthreads();
  u32 mask = 0;
#pragma unroll 1
  for (int j = 0; j < SC_IPT; ++j) {
    // row inside the tile (s_k index r + TIE_HALO).  The fix pass only counts, so its threads take
    // strided rows (conflict-free shared-memory reads); the gather pass needs consecutive rows per
    // thread for the ordered compaction.
    const u32 r = FIX ? (u32)j * SC_THREADS + tid : tid * SC_IPT + j;
    if (r >= rows) {
      if (FIX) continue;
      break;
    }
    const u32 at = r + TIE_HALO;
    const u64 c = key_cs(s_k[at], L);
    const bool prev_same = (base + r > 0) && key_cs(s_k[at - 1], L) == c;
    const bool next_same = (base + r + 1 < n) && key_cs(s_k[at + 1], L) == c;
    if (!prev_same && !next_same) continue;
    u32 back = 0, fwd = 0;
    while (back < TIE_SMALL && base + r > back && key_cs(s_k[at - back - 1], L) == c) ++back;
    while (fwd < TIE_SMALL && base + r + fwd + 1 < n && key_cs(s_k[at + fwd + 1], L) == c) ++fwd;
    const u32 m = back + fwd + 1;
    if (m > TIE_SMALL) {
      mask |= 1u << j;  // long run: general path
    } else if (FIX && back == 0) {
      // owner of a short run: stable insertion sort of (key, id) by key, in place
      const u64 p0 = base + r;
      if (m == 2) {
        bool swap = s_k[at] > s_k[at + 1];
        if (e_raw && s_k[at] == s_k[at + 1]) swap = after(s_k[at], ids[p0], s_k[at + 1], ids[p0 + 1]);
        if (swap) {
          const u32 v0 = ids[p0], v1 = ids[p0 + 1];
          keys[p0] = s_k[at + 1];
          keys[p0 + 1] = s_k[at];
          ids[p0] = v1;
          ids[p0 + 1] = v0;
        }
      } else {
        u64 k[TIE_SMALL];
        u32 v[TIE_SMALL];
        for (u32 x = 0; x < m; ++x) {
          const u64 kx = s_k[at + x];
          const u32 vx = ids[p0 + x];
          u32 y = x;
          while (y > 0 && after(k[y - 1], v[y - 1], kx, vx)) {
            k[y] = k[y - 1];
            v[y] = v[y - 1];
            --y;
          }
          k[y] = kx;
          v[y] = vx;
        }
        for (u32 x = 0; x < m; ++x) {
          if (k[x] != s_k[at + x]) keys[p0 + x] = k[x];
          ids[p0 + x] = v[x];  // equal keys may still have traded places with each other's neighbours
        }
      }
    }
  }
  return mask;
}

// pass 1: fix short runs in place, count rows of long runs per tile
static __global__ void __launch_bounds__(SC_THREADS)
    tie_fix_kernel(u64* __restrict__ keys, u32* __restrict__ ids, u64 n, KeyLayout L, const i64* __restrict__ e_raw,
                   u64* __restrict__ spine) {
  __shared__ u64 s_k[SC_TILE + 2 * TIE_HALO];
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = tie_scan_tile<true>(keys, ids, n, L, base, rows, s_k, e_raw);
  const u64 c = block_sum_u64((u64)__popc(mask), s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = c;
}
// pass 2 (after the spine scan): tiles that hold rows of long runs compact them as (key, id, slot)
static __global__ void __launch_bounds__(SC_THREADS)
    tie_gather_kernel(u64* __restrict__ keys, u32* __restrict__ ids, u64 n, KeyLayout L, const u64* __restrict__ spine,
                      u64 ntiles, const u64* __restrict__ nT, u64* __restrict__ tkeys, u32* __restrict__ tids,
                      u32* __restrict__ tpos) {
  __shared__ u64 s_k[SC_TILE + 2 * TIE_HALO];
  __shared__ u64 s_tmp[8];
  const u64 tile = blockIdx.x;
  const u64 first = spine[tile];
  const u64 next = tile + 1 < ntiles ? spine[tile + 1] : *nT;
  if (next == first) return;  // nearly every tile of real data
  const u64 base = tile * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = tie_scan_tile<false>(keys, ids, n, L, base, rows, s_k, nullptr);
  u64 tot;
  u64 at_out = first + block_excl_sum_u64((u64)__popc(mask), s_tmp, &tot);
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    if (mask & (1u << j)) {
      const u32 r = threadIdx.x * SC_IPT + j;
      tkeys[at_out] = s_k[r + TIE_HALO];
      tids[at_out] = ids[base + r];
      tpos[at_out] = (u32)(base + r);
      ++at_out;
    }
  }
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

// ---- clamp mode: (group, start, true end', row) by two stable sorts ----
static __global__ void __launch_bounds__(256) true_end_keys_kernel(const i64* __restrict__ s, const i64* __restrict__ e,
                                                                   u64 n, int point_adjust, u64* __restrict__ keys) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    i64 b = e[i];
    if ((point_adjust & KEY_POINT_ADJUST) && s[i] == b && b != I64_MAX) b += 1;
    keys[i] = (u64)b ^ 0x8000000000000000ull;  // order-preserving image of the signed value
  }
}
// packed keys of the rows in the order given by `order` (+ histograms of the digit places of the next sort)
static __global__ void __launch_bounds__(HIST_THREADS)
    pack_keys_gather_kernel(const i32* __restrict__ grp, const i64* __restrict__ s, const i64* __restrict__ e,
                            const u32* __restrict__ order, u64 n, KeyLayout L, int point_adjust, RadixPlan rp,
                            u64* __restrict__ keys, u64* __restrict__ ghist) {
  extern __shared__ u32 hist_tabs[];
  hist_clear(hist_tabs, rp);
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u32 r = order[i];
    i64 a = s[r], b = e[r];
    if ((point_adjust & KEY_POINT_ADJUST) && a == b) b += 1;
    b = min(b, L.endcap);
    const u64 off = grp ? L.goff[grp[r]] : 0ull;
    const u64 at = (point_adjust & KEY_BY_END) ? (u64)b : (u64)a;
    const u64 k = ((off + (at - L.cmin)) << L.LB) | ((u64)b - (u64)a);
    keys[i] = k;
    hist_add(hist_tabs, rp, k, true);
  }
  __syncthreads();
  hist_flush(hist_tabs, rp, ghist);
}

static int sort_rows_two_keys(const i32* grp, const i64* s, const i64* e, i64 n, const KeyLayout& L, int point_adjust,
                              const SortBufs& ws, const RadixScratch& sc, cudaStream_t st, SortResult* res) {
  // 1. rows by true end' (64-bit keys, row number as payload)
  const RadixPlan all64 = make_radix_plan(0, 64);
  BF_LAUNCH(PC_TIEFIX, 16 * n, st,
            true_end_keys_kernel<<<stream_grid((u64)n, 1024), 256, 0, st>>>(s, e, (u64)n, point_adjust, ws.keyA));
  BF_CUDA(cudaMemsetAsync(sc.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
  BF_TRY(launch_radix_hist(ws.keyA, (u64)n, nullptr, all64, sc.hist, st, PC_TIEFIX));
  SortResult by_end;
  BF_TRY(radix_sort_passes<true>(ws.keyA, ws.keyB, ws.idA, ws.idB, true, (u64)n, nullptr, all64, sc, st, &by_end, PC_TIEFIX));
  // 2. packed keys in that order, stable sort on the (group, start) field, payload = row
  const RadixPlan rp = make_radix_plan(L.LB, L.total_bits);
  const size_t smem = hist_smem_bytes(rp);
  BF_CUDA(cudaMemsetAsync(sc.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
  BF_CUDA(cudaFuncSetAttribute(pack_keys_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BF_LAUNCH(PC_PACK, 32 * n, st,
            pack_keys_gather_kernel<<<stream_grid((u64)n, 2048), HIST_THREADS, smem, st>>>(
                grp, s, e, by_end.vals, (u64)n, L, point_adjust, rp, by_end.keys, sc.hist));
  BF_TRY(radix_sort_passes<true>(by_end.keys, by_end.keys_alt, by_end.vals, by_end.vals_alt, false, (u64)n, nullptr, rp,
                                 sc, st, res));
  return BF_OK;
}

// Sort rows (grp may be null for a single group) into res->keys / res->vals.
static int sort_rows(const i32* grp, const i64* s, const i64* e, i64 n, const KeyLayout& L, int point_adjust,
                     const SortBufs& ws, const RadixScratch& sc, cudaStream_t st, SortResult* res) {
  res->keys = ws.keyA;
  res->vals = ws.idA;
  res->keys_alt = ws.keyB;
  res->vals_alt = ws.idB;
  if (n == 0) return BF_OK;
  if (L.endcap != I64_MAX && !(point_adjust & KEY_NO_TIE_FIX)) {
    // Clamped ends (INT64_MAX-style inputs, keys.cuh measure_layout): the packed length no longer orders
    // rows that share a start and reach beyond the clamp.  Rare and usually tiny (views, complement
    // bounds), so do the plain two-key LSD sort instead of start-sort + tie fix: stable sort by the TRUE
    // end', then stable sort by (group, start) -> (group, start, end', row) exactly.
    return sort_rows_two_keys(grp, s, e, n, L, point_adjust, ws, sc, st, res);
  }
  // main sort: stable, on the (group, start) bits only -- or nothing at all when the extent pass found the
  // rows already in that order (the tie fix-up below still puts equal starts in (end', row) order)
  const bool presorted = (point_adjust & KEY_PRESORTED) && !(point_adjust & KEY_BY_END);
  const RadixPlan rp = presorted ? make_radix_plan(0, 0) : make_radix_plan(L.LB, L.total_bits);
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
  BF_LAUNCH(PC_TIEFIX, 8 * n, st, tie_fix_kernel<<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, L,
                                                                             L.endcap != I64_MAX ? e : nullptr, ws.spine));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(ws.spine, (u64)tiles, ws.nT));
  BF_LAUNCH(PC_TIEFIX, 0, st,
            tie_gather_kernel<<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, L, ws.spine, (u64)tiles, ws.nT,
                                                            tkeyA, tidA, tpos));
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
