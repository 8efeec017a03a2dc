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

// ---- finishing pass: local order inside runs of rows that share their high key bits ----
// The radix passes only sort the key bits from `run_shift` up (a stable LSD sort): all of the (group, start)
// field when rows are sparse, but for dense sets -- 1e8 rows on a 2^32 axis are 30 positions apart -- its
// lowest bits are left out, which saves a whole pass over the data (24 B/row).  Afterwards the rows of a
// "run" (equal key >> run_shift) sit next to each other in input order, a handful at a time, and this one pass
// finishes the job: every row counts the rows of its run that must precede it -- smaller (key >> cmp_shift),
// or equal and earlier -- and is written to that place in the other buffer.  cmp_shift = 0 orders by the full
// key, i.e. (group, start, end', row) = np.lexsort([ends, starts]) of the reference (arrops.py:332-335,
// 459-460); cmp_shift = LB orders by (group, start) and keeps equal starts in row order (KEY_NO_TIE_FIX).
// With run_shift = LB it is the former tie fix-up (rows sharing (group, start) into (end', row) order).
// A tile's keys (+ RUN_MAX keys each side) are staged in shared memory; one thread per row, no owner thread,
// no atomics, no dependence between blocks.  Rows of runs longer than RUN_MAX (heavily duplicated or
// clustered input) are copied through unchanged, compacted as (key, id, slot), sorted by the remaining
// key bits with the stable radix sort and written back into the same slots.
constexpr int RUN_MAX = 64;
constexpr int RUN_HALO = RUN_MAX;
constexpr int RUN_STAGED = SC_TILE + 2 * RUN_HALO;  // rows a tile stages: its own and RUN_HALO either side

// Stage a tile and classify / place its rows, IN PLACE.
//   s_k[i] = keys[base - RUN_HALO + i]; rows are compared by s_k >> cmp_shift; two neighbours belong to one
//   run when they agree at and above bit run_shift, i.e. (a ^ b) < 2^run_shift.
// Nearly every row of sparse data is a run of its own: two comparisons with its neighbours and it is done.  A
// row with an equal neighbour walks to both ends of its run (at most RUN_MAX rows, else the run is "long") and
// counts on the way the rows that must precede it.
// PLACE: every short run is put in order by the ONE tile that holds its first row -- so the rows of a run are
// read (from the staged snapshot) and written by a single block: in place without a race; the rows of the run
// may reach up to RUN_MAX - 1 rows into the next tile's range, which skips them.  A row that already is where it
// belongs is not written at all, so the pass reads 8 B per row and writes next to nothing.  Rows that move
// are parked in a list in shared memory (slot, destination, row id) and written after a barrier, when every
// row id that moves has been read.  Another tile may stage rows of a run while its owner rewrites them: it
// only looks at their run prefix, which every version of the row shares.
// Returns the mask of this thread's OWN tile rows that belong to long runs; bit j is tile row
// j * SC_THREADS + tid (PLACE) or tid * SC_IPT + j (!PLACE: the compaction needs consecutive rows per thread).
struct RunMove {
  u32 id;
  u16 from, to;  // staged indices
};
constexpr int RUN_MOVES = SC_TILE + RUN_HALO;  // every row a tile can own

template <bool PLACE>
__device__ __forceinline__ u32 run_scan_tile(u64* __restrict__ keys, u32* __restrict__ ids, u64 n, int run_shift,
                                             int cmp_shift, u64 base, u64 rows, u64* s_k, RunMove* s_move,
                                             u32* s_nmove) {
  const u32 tid = threadIdx.x;
  // staged indices [v_lo, v_hi) hold rows of the array (32-bit from here on: the kernel is bound by its
  // instruction count, and 64-bit index arithmetic doubles it)
  const u32 v_lo = base >= (u64)RUN_HALO ? 0u : (u32)(RUN_HALO - base);
  const u64 left = n - (base + v_lo - RUN_HALO);  // rows from staged index v_lo to the end of the array
  const u32 v_hi = left >= (u64)(RUN_STAGED - v_lo) ? (u32)RUN_STAGED : v_lo + (u32)left;
  const u64* kp = keys + base - RUN_HALO;  // kp[i] = staged row i (only dereferenced for v_lo <= i < v_hi)
  {  // all loads of a thread in flight before its first store
    u64 v[SC_IPT], hv = ~0ull;
    const u32 hi = tid < RUN_HALO ? tid : SC_TILE + tid;  // this thread's halo row (threads 0 .. 2 * RUN_HALO - 1)
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
      const u32 i = RUN_HALO + j * SC_THREADS + tid;
      v[j] = i < v_hi ? kp[i] : ~0ull;  // tile rows are never below v_lo
    }
    if (tid < 2 * RUN_HALO) {
      if (hi >= v_lo && hi < v_hi) hv = kp[hi];
      s_k[hi] = hv;
    }
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) s_k[RUN_HALO + j * SC_THREADS + tid] = v[j];
    if (PLACE && tid == 0) *s_nmove = 0;
  }
  __syncthreads();
  // run_shift = 64: the whole set is one run (a tiny set sorted by this pass alone)
  const u64 run_lim = run_shift < 64 ? 1ull << run_shift : ~0ull;
  const bool one_run = run_shift >= 64;
  auto same_run = [&](u64 a, u64 b) { return one_run || (a ^ b) < run_lim; };
  u32 mask = 0;
  const u32 nrows = (u32)rows;
  // rows of this thread: SC_IPT tile rows, and for RUN_HALO threads one row of the forward halo (a run that
  // starts in this tile may end there)
#pragma unroll 1
  for (int j = 0; j <= SC_IPT; ++j) {
    u32 r;
    if (j < SC_IPT) {
      r = PLACE ? (u32)j * SC_THREADS + tid : tid * SC_IPT + j;
      if (r >= nrows) continue;
    } else {
      if (!PLACE || tid >= RUN_HALO) break;
      r = SC_TILE + tid;
      if (r + RUN_HALO >= v_hi) break;
    }
    const u32 i = r + RUN_HALO;
    const u64 k = s_k[i];
    const bool prev_same = i > v_lo && same_run(k, s_k[i - 1]);
    const bool next_same = i + 1 < v_hi && same_run(k, s_k[i + 1]);
    if (!prev_same && !next_same) continue;  // a run of its own
    // walk to both ends of the run; rows before me count when they compare <= me, rows after me when < me
    const u64 mine = k >> cmp_shift;
    u32 back = 0, fwd = 0, rank = 0;
    while (back < RUN_MAX && i - back > v_lo) {
      const u64 o = s_k[i - back - 1];
      if (!same_run(k, o)) break;
      rank += (o >> cmp_shift) <= mine;
      ++back;
    }
    while (fwd < RUN_MAX && i + fwd + 1 < v_hi) {
      const u64 o = s_k[i + fwd + 1];
      if (!same_run(k, o)) break;
      rank += (o >> cmp_shift) < mine;
      ++fwd;
    }
    if (back + fwd + 1 > RUN_MAX) {  // also when the walk ran into the edge of the staged rows
      if (j < SC_IPT) mask |= 1u << j;
      continue;
    }
    if (!PLACE) continue;  // classification only
    const u32 start = i - back;
    if (start < RUN_HALO || start >= RUN_HALO + SC_TILE) continue;  // the run belongs to another tile
    if (rank != back) {  // the row moves
      const u32 slot = atomicAdd(s_nmove, 1u);
      BF_DEVICE_CHECK(slot < (u32)RUN_MOVES && start + rank < (u32)RUN_STAGED);
      RunMove m;
      m.id = ids[base + r];
      m.from = (u16)i;
      m.to = (u16)(start + rank);
      s_move[slot] = m;
    }
  }
  if (PLACE) {
    __syncthreads();  // every row id that moves has been read before the first one is overwritten
    const u32 nm = *s_nmove;
    for (u32 x = tid; x < nm; x += SC_THREADS) {
      const RunMove m = s_move[x];
      const u64 at = base + m.to - RUN_HALO;  // >= base: the run starts in this tile
      BF_DEVICE_CHECK(at >= base && at < n && m.from < RUN_STAGED);
      keys[at] = s_k[m.from];
      ids[at] = m.id;
    }
  }
  return mask;
}

// pass 1: order the short runs in place, count rows of long runs per tile
static __global__ void __launch_bounds__(SC_THREADS)
    local_order_kernel(u64* __restrict__ keys, u32* __restrict__ ids, u64 n, int run_shift, int cmp_shift,
                       u64* __restrict__ spine) {
  __shared__ u64 s_k[RUN_STAGED];
  __shared__ RunMove s_move[RUN_MOVES];
  __shared__ u32 s_nmove;
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = run_scan_tile<true>(keys, ids, n, run_shift, cmp_shift, base, rows, s_k, s_move, &s_nmove);
  const u64 c = block_sum_u64((u64)__popc(mask), s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = c;
}
// pass 2 (after the spine scan): tiles that hold rows of long runs compact them as (key, id, slot)
static __global__ void __launch_bounds__(SC_THREADS)
    long_run_gather_kernel(u64* __restrict__ keys, u32* __restrict__ ids, u64 n, int run_shift,
                           const u64* __restrict__ spine, u64 ntiles, const u64* __restrict__ nT,
                           u64* __restrict__ tkeys, u32* __restrict__ tids, u32* __restrict__ tpos) {
  __shared__ u64 s_k[RUN_STAGED];
  __shared__ u64 s_tmp[8];
  const u64 tile = blockIdx.x;
  const u64 first = spine[tile];
  const u64 next = tile + 1 < ntiles ? spine[tile + 1] : *nT;
  if (next == first) return;  // nearly every tile of real data
  const u64 base = tile * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = run_scan_tile<false>(keys, ids, n, run_shift, 0, base, rows, s_k, nullptr, nullptr);
  u64 tot;
  u64 at_out = first + block_excl_sum_u64((u64)__popc(mask), s_tmp, &tot);
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    if (mask & (1u << j)) {
      const u32 r = threadIdx.x * SC_IPT + j;
      tkeys[at_out] = s_k[r + RUN_HALO];
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
    keys[p] = tkeys[t];  // p was recorded from a valid slot by long_run_gather_kernel
    ids[p] = tids[t];
  }
}

// ---------------------------------------------------------------- host side

// How many low bits of the B-bit (group, start) field the radix passes may leave to the finishing pass.
// Every dropped byte saves a pass over the data (24 B/row), but the finishing pass pays ~6 instructions per
// pair of rows that share a run: measured at 1e8 rows on the 32-bit hg38 axis (profiles/NOTES.md, r2), three
// passes + runs of ~8 rows cost 2.31 + 1.47 ms against 3.08 + 0.53 ms for four passes + runs of ~1 row -- a
// loss -- while sets whose runs stay around one row (1e7 rows: 0.8 per 256 positions) just lose a pass.  So
// bits are dropped only while the expected run of evenly spread rows, n * 2^drop / 2^B, stays below 1.5.
static inline int choose_drop_bits(int B, i64 n) {
  if (n <= 0) return 0;
  int best = 0;
  for (int passes = (B + 7) / 8; passes >= 0; --passes) {
    const int sorted_bits = passes * 8 < B ? passes * 8 : B;
    const int drop = B - sorted_bits;
    // expected rows per run = n / 2^(B - drop) <= 1.5
    const double buckets = sorted_bits >= 63 ? 9.2e18 : (double)(1ull << sorted_bits);
    if ((double)n > 1.5 * buckets) break;
    best = drop;
  }
  return best;
}

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
  // main sort: stable LSD passes on the upper bits of the (group, start) field -- or none at all when the
  // extent pass found the rows already in that order -- then one finishing pass that orders the rows inside
  // every run of equal sorted bits (local_order_kernel)
  const bool presorted = (point_adjust & KEY_PRESORTED) && !(point_adjust & KEY_BY_END);
  const bool keep_row_order = (point_adjust & KEY_NO_TIE_FIX) != 0;  // equal (group, start): rows stay in row order
  const int drop = presorted ? 0 : choose_drop_bits(L.B, n);
  const int run_shift = L.LB + drop;
  const int cmp_shift = keep_row_order ? L.LB : 0;
  const RadixPlan rp = presorted ? make_radix_plan(0, 0) : make_radix_plan(run_shift, L.total_bits);
  BF_CUDA(cudaMemsetAsync(sc.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
  BF_TRY(launch_pack_keys(grp, s, e, (u64)n, L, point_adjust, rp, ws.keyA, sc.hist, st));
  BF_TRY(radix_sort_passes<true>(ws.keyA, ws.keyB, ws.idA, ws.idB, true, (u64)n, nullptr, rp, sc, st, res));
  if (keep_row_order && drop == 0) return BF_OK;  // runs = equal (group, start), already in row order
  const unsigned tiles = (unsigned)scan_tiles((u64)n);
  BF_LAUNCH(PC_TIEFIX, 8 * n, st,
            local_order_kernel<<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, run_shift, cmp_shift, ws.spine));
  // rows of long runs (rare): compact, sort by the bits below the run prefix, write back into their slots
  u64* tkeyA = res->keys_alt;
  u32* tidA = res->vals_alt;
  u64* tkeyB = ws.tmp8;
  u32* tidB = ws.tmp4a;
  u32* tpos = ws.tmp4b;
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(ws.spine, (u64)tiles, ws.nT));
  BF_LAUNCH(PC_TIEFIX, 0, st,
            long_run_gather_kernel<<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, run_shift, ws.spine, (u64)tiles,
                                                                 ws.nT, tkeyA, tidA, tpos));
  // the compacted rows are still in run order and, inside a run, in input order: a stable sort on ALL compared
  // bits puts them into the order of the whole array restricted to those slots
  const RadixPlan full = make_radix_plan(cmp_shift, L.total_bits);
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
