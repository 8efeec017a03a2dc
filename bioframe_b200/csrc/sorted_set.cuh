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

// Stage a tile and classify / place its rows.  PLACE: write every row of a short run to its final position in
// (kout, vout), copy rows of long runs through.  Returns the mask of this thread's rows that belong to long
// runs; row j of the mask is tile row  j * SC_THREADS + tid  (PLACE)  or  tid * SC_IPT + j  (!PLACE: the
// compaction needs consecutive rows per thread).
template <bool PLACE>
__device__ __forceinline__ u32 run_scan_tile(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n,
                                             int run_shift, int cmp_shift, u64 base, u64 rows, u64* s_k,
                                             u64* __restrict__ kout, u32* __restrict__ vout) {
  const u32 tid = threadIdx.x;
  u32 idv[SC_IPT];
  {  // s_k[i] = keys[base - RUN_HALO + i]; all loads of a thread in flight before its first store
    u64 v[SC_IPT], hv = ~0ull;
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
      const u64 p = base + (u64)j * SC_THREADS + tid;
      v[j] = p < n ? keys[p] : ~0ull;
      if (PLACE) idv[j] = p < n ? ids[p] : 0u;
    }
    if (tid < 2 * RUN_HALO) {
      const i64 p = tid < RUN_HALO ? (i64)base - RUN_HALO + tid : (i64)base + SC_TILE + (tid - RUN_HALO);
      if (p >= 0 && (u64)p < n) hv = keys[p];
      s_k[tid < RUN_HALO ? tid : SC_TILE + tid] = hv;
    }
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) s_k[RUN_HALO + j * SC_THREADS + tid] = v[j];
  }
  __syncthreads();
  u32 mask = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u32 r = PLACE ? (u32)j * SC_THREADS + tid : tid * SC_IPT + j;
    if (r >= rows) continue;
    const u32 at = r + RUN_HALO;
    const u64 pos = base + r;
    const u64 k = s_k[at];
    // run_shift can be 64 (no bit sorted: a tiny set is one run); cmp_shift < 64 always
    const u64 run = run_shift < 64 ? k >> run_shift : 0ull, mine = k >> cmp_shift;
    // rows of my run before me: all of them count if they compare <= me (equal ones are earlier rows);
    // rows after me count if they compare < me
    u32 back = 0, fwd = 0, rank = 0;
    while (back < RUN_MAX && pos > back) {
      const u64 o = s_k[at - back - 1];
      if ((run_shift < 64 ? o >> run_shift : 0ull) != run) break;
      rank += (o >> cmp_shift) <= mine;
      ++back;
    }
    while (fwd < RUN_MAX && pos + fwd + 1 < n) {
      const u64 o = s_k[at + fwd + 1];
      if ((run_shift < 64 ? o >> run_shift : 0ull) != run) break;
      rank += (o >> cmp_shift) < mine;
      ++fwd;
    }
    if (back + fwd + 1 > RUN_MAX) {
      mask |= 1u << j;
      if (PLACE) {
        kout[pos] = k;
        vout[pos] = idv[j];
      }
    } else if (PLACE) {
      const u64 dst = pos - back + rank;
      kout[dst] = k;
      vout[dst] = idv[j];
    }
  }
  return mask;
}

// pass 1: place the rows of short runs, count rows of long runs per tile
static __global__ void __launch_bounds__(SC_THREADS)
    local_order_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n, int run_shift, int cmp_shift,
                       u64* __restrict__ kout, u32* __restrict__ vout, u64* __restrict__ spine) {
  __shared__ u64 s_k[SC_TILE + 2 * RUN_HALO];
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = run_scan_tile<true>(keys, ids, n, run_shift, cmp_shift, base, rows, s_k, kout, vout);
  const u64 c = block_sum_u64((u64)__popc(mask), s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = c;
}
// pass 2 (after the spine scan): tiles that hold rows of long runs compact them as (key, id, slot)
static __global__ void __launch_bounds__(SC_THREADS)
    long_run_gather_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n, int run_shift,
                           const u64* __restrict__ spine, u64 ntiles, const u64* __restrict__ nT,
                           u64* __restrict__ tkeys, u32* __restrict__ tids, u32* __restrict__ tpos) {
  __shared__ u64 s_k[SC_TILE + 2 * RUN_HALO];
  __shared__ u64 s_tmp[8];
  const u64 tile = blockIdx.x;
  const u64 first = spine[tile];
  const u64 next = tile + 1 < ntiles ? spine[tile + 1] : *nT;
  if (next == first) return;  // nearly every tile of real data
  const u64 base = tile * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = run_scan_tile<false>(keys, nullptr, n, run_shift, 0, base, rows, s_k, nullptr, nullptr);
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
    keys[p] = tkeys[t];
    ids[p] = tids[t];
  }
}

// ---------------------------------------------------------------- host side

// How many low bits of the B-bit (group, start) field the radix passes may leave to the finishing pass: a run
// of 2^drop positions should hold about 8-16 rows of evenly spread input, and the sorted bits fill whole 8-bit
// passes.  1e8 rows on the 32-bit hg38 axis: drop = 8, three passes instead of four.
static inline int choose_drop_bits(int B, i64 n) {
  if (n <= 0) return 0;
  int drop = B - bit_width((u64)n) + 3;
  if (drop <= 0) return 0;
  if (drop > B) drop = B;
  const int passes = (B - drop + 7) / 8;
  const int sorted_bits = passes * 8 < B ? passes * 8 : B;
  return B - sorted_bits;
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
  BF_LAUNCH(PC_TIEFIX, 24 * n, st,
            local_order_kernel<<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, run_shift, cmp_shift, res->keys_alt,
                                                            res->vals_alt, ws.spine));
  {  // the ordered rows are in the other buffer pair now
    u64* tk = res->keys;
    res->keys = res->keys_alt;
    res->keys_alt = tk;
    u32* tv = res->vals;
    res->vals = res->vals_alt;
    res->vals_alt = tv;
  }
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
