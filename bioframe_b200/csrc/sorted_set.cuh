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
constexpr int RUN_STAGED = SC_TILE + 2 * RUN_HALO;  // 2176 staged rows = 68 words of run-head bits
constexpr int RUN_WORDS = RUN_STAGED / 32;
static_assert(RUN_STAGED % 32 == 0 && RUN_HALO % 32 == 0, "head bit words must tile the staged rows");

// Stage a tile and classify / place its rows.  PLACE: write every row of a short run to its final position in
// (kout, vout), copy rows of long runs through.  Returns the mask of this thread's rows that belong to long
// runs; bit j of the mask is tile row  j * SC_THREADS + tid  (PLACE)  or  tid * SC_IPT + j  (!PLACE: the
// compaction needs consecutive rows per thread).
//   s_c[i]   = keys[base - RUN_HALO + i] >> cmp_shift   (what rows are compared by)
//   s_head   = one bit per staged row: the row starts a run (its key >> run_shift differs from its
//              predecessor's).  A row finds its run as [last head at or before it, next head after it) with a
//              few bit operations on at most three words either way -- no walk, no divergence; a run whose
//              bounds are not inside that reach is longer than RUN_MAX.
//   SUB32: run_shift - cmp_shift <= 32, so rows of one run differ only in the low 32 bits of s_c: the rank
//              loop then reads 4-byte words (s_lo).
template <bool PLACE, bool SUB32>
__device__ __forceinline__ u32 run_scan_tile(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n,
                                             int run_shift, int cmp_shift, u64 base, u64 rows, u64* s_c, u32* s_lo,
                                             u32* s_head, u64* __restrict__ kout, u32* __restrict__ vout) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64 kv[SC_IPT];
  u32 idv[SC_IPT];
  {  // all loads of a thread in flight before its first store
    u64 hv = ~0ull;
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
      const u64 p = base + (u64)j * SC_THREADS + tid;
      kv[j] = p < n ? keys[p] : ~0ull;
      if (PLACE) idv[j] = p < n ? ids[p] : 0u;
    }
    if (tid < 2 * RUN_HALO) {
      const i64 p = tid < RUN_HALO ? (i64)base - RUN_HALO + tid : (i64)base + SC_TILE + (tid - RUN_HALO);
      if (p >= 0 && (u64)p < n) hv = keys[p];
      const u32 i = tid < RUN_HALO ? tid : SC_TILE + tid;
      s_c[i] = hv >> cmp_shift;
      if (SUB32) s_lo[i] = (u32)(hv >> cmp_shift);
    }
#pragma unroll
    for (int j = 0; j < SC_IPT; ++j) {
      const u32 i = RUN_HALO + j * SC_THREADS + tid;
      s_c[i] = kv[j] >> cmp_shift;
      if (SUB32) s_lo[i] = (u32)(kv[j] >> cmp_shift);
    }
  }
  __syncthreads();
  const int pshift = run_shift - cmp_shift;  // run prefix of a staged value; 64 = every row in one run
  for (u32 c = warp; c < RUN_WORDS; c += SC_THREADS / 32) {
    const u32 i = c * 32 + lane;
    const i64 p = (i64)base - RUN_HALO + (i64)i;
    bool head = true;  // rows outside the array and the first staged row close / open a run
    if (p > 0 && (u64)p < n && i > 0) {
      const u64 a = s_c[i], b = s_c[i - 1];
      head = pshift < 64 && (a >> pshift) != (b >> pshift);
    }
    const u32 m = __ballot_sync(0xffffffffu, head);
    if (lane == 0) s_head[c] = m;
  }
  __syncthreads();
  u32 mask = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u32 r = PLACE ? (u32)j * SC_THREADS + tid : tid * SC_IPT + j;
    if (r >= rows) continue;
    const u32 i = r + RUN_HALO, c = i >> 5, l = i & 31;
    // first row of my run: the last head at or before me
    i32 start = -1;
    {
      const u32 w0 = s_head[c] & (0xffffffffu >> (31 - l));
      if (w0) {
        start = (i32)(c << 5) + 31 - __clz(w0);
      } else {
        const u32 w1 = s_head[c - 1];
        if (w1) {
          start = (i32)((c - 1) << 5) + 31 - __clz(w1);
        } else {
          const u32 w2 = s_head[c - 2];
          if (w2) start = (i32)((c - 2) << 5) + 31 - __clz(w2);
        }
      }
    }
    // one past its last row: the next head after me
    i32 end = -1;
    {
      const u32 w0 = l == 31 ? 0u : s_head[c] & (0xffffffffu << (l + 1));
      if (w0) {
        end = (i32)(c << 5) + __ffs(w0) - 1;
      } else {
        const u32 w1 = s_head[c + 1];
        if (w1) {
          end = (i32)((c + 1) << 5) + __ffs(w1) - 1;
        } else {
          const u32 w2 = c + 2 < RUN_WORDS ? s_head[c + 2] : 0u;
          if (w2) end = (i32)((c + 2) << 5) + __ffs(w2) - 1;
        }
      }
    }
    const u64 pos = base + r;
    const bool is_long = start < 0 || end < 0 || end - start > RUN_MAX || start == 0;  // staged row 0: bound unknown
    if (is_long) {
      mask |= 1u << j;
      if (PLACE) {
        kout[pos] = kv[j];
        vout[pos] = idv[j];
      }
    } else if (PLACE) {
      // rows of my run that precede me: earlier ones that compare <= me, later ones that compare < me
      u32 rank = 0;
      if (SUB32) {
        const u32 mine = s_lo[i];
        for (i32 x = start; x < (i32)i; ++x) rank += s_lo[x] <= mine;
        for (i32 x = (i32)i + 1; x < end; ++x) rank += s_lo[x] < mine;
      } else {
        const u64 mine = s_c[i];
        for (i32 x = start; x < (i32)i; ++x) rank += s_c[x] <= mine;
        for (i32 x = (i32)i + 1; x < end; ++x) rank += s_c[x] < mine;
      }
      const u64 dst = pos - (u64)((i32)i - start) + rank;
      kout[dst] = kv[j];
      vout[dst] = idv[j];
    }
  }
  return mask;
}

// pass 1: place the rows of short runs, count rows of long runs per tile
template <bool SUB32>
static __global__ void __launch_bounds__(SC_THREADS)
    local_order_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n, int run_shift, int cmp_shift,
                       u64* __restrict__ kout, u32* __restrict__ vout, u64* __restrict__ spine) {
  __shared__ u64 s_c[RUN_STAGED];
  __shared__ u32 s_lo[SUB32 ? RUN_STAGED : 1];
  __shared__ u32 s_head[RUN_WORDS];
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = run_scan_tile<true, SUB32>(keys, ids, n, run_shift, cmp_shift, base, rows, s_c, s_lo, s_head, kout, vout);
  const u64 c = block_sum_u64((u64)__popc(mask), s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = c;
}
// pass 2 (after the spine scan): tiles that hold rows of long runs compact them as (key, id, slot)
static __global__ void __launch_bounds__(SC_THREADS)
    long_run_gather_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n, int run_shift,
                           const u64* __restrict__ spine, u64 ntiles, const u64* __restrict__ nT,
                           u64* __restrict__ tkeys, u32* __restrict__ tids, u32* __restrict__ tpos) {
  __shared__ u64 s_c[RUN_STAGED];
  __shared__ u32 s_head[RUN_WORDS];
  __shared__ u64 s_tmp[8];
  const u64 tile = blockIdx.x;
  const u64 first = spine[tile];
  const u64 next = tile + 1 < ntiles ? spine[tile + 1] : *nT;
  if (next == first) return;  // nearly every tile of real data
  const u64 base = tile * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
  const u32 mask = run_scan_tile<false, false>(keys, nullptr, n, run_shift, 0, base, rows, s_c, nullptr, s_head, nullptr, nullptr);
  u64 tot;
  u64 at_out = first + block_excl_sum_u64((u64)__popc(mask), s_tmp, &tot);
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    if (mask & (1u << j)) {
      const u32 r = threadIdx.x * SC_IPT + j;
      tkeys[at_out] = s_c[r + RUN_HALO];  // cmp_shift = 0 here: the staged value is the key itself
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
  if (run_shift - cmp_shift <= 32) {
    BF_LAUNCH(PC_TIEFIX, 24 * n, st,
              local_order_kernel<true><<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, run_shift, cmp_shift,
                                                                    res->keys_alt, res->vals_alt, ws.spine));
  } else {
    BF_LAUNCH(PC_TIEFIX, 24 * n, st,
              local_order_kernel<false><<<tiles, SC_THREADS, 0, st>>>(res->keys, res->vals, (u64)n, run_shift, cmp_shift,
                                                                     res->keys_alt, res->vals_alt, ws.spine));
  }
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
