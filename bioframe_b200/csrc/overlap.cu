// bf_overlap_count / bf_overlap_emit: the whole of ops._overlap_intidxs
// (src/bioframe/ops.py:242-338) for all groups in one set of launches.
//
//   extent  -> pack keys + digit histograms -> onesweep radix argsort (both sets)
//   -> group segment table -> [running max of end' for outer joins]
//   -> search + count (A side per sorted query, B side per sorted target)
//   -> exclusive scans -> per-group offset table -> (host reads the row count)
//   -> emit A pairs, emit B pairs, emit unpaired rows.
//
// Output order is the reference's: per group (in group-rank order) the A block
// (per sorted row of set 1 the targets whose start is in [s1, e1')), the B
// block (per sorted row of set 2 the queries whose start is in (s2, e2')),
// then unpaired rows of set 1 ascending, then of set 2 ascending
// (arrops.py:338-368, ops.py:319-334).
#include "common.cuh"
#include "keys.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "lookback.cuh"
#include "sorted_set.cuh"
#include "../../include/bioframe_b200.h"
#include <vector>

namespace bf {

constexpr i64 EMIT_CHUNK_BLOCKS = 1 << 20;  // emit blocks per partition chunk (fixed-size scratch)
constexpr i64 SEARCH_MIN_ROWS = 256;        // == SR_MIN_ROWS (smallest block of rows of the search kernels)

// ---------------------------------------------------------------- workspace
struct SetWS {
  u32 *lo;       // [n] first partner position per sorted row
  u32 *loc;      // [n] exclusive output offset inside the row's search tile
  u64 *tilepre;  // [search tiles + 1] exclusive output offset of every search tile; last = total
  int shift;     // log2(rows per search tile) of this side (host copy, set by launch_search)
  u64 *spine;    // [tiles]
  i64 *bracket;  // [tiles+1] window of the other set per block of sorted rows
  u64 *pmax;   // [n] running max of ce (only for outer joins on the other side)
  u8 *rowflag; // [n] 1 = row has no partner (row order)
  u64 *ukeyA, *ukeyB;  // [n] (rank << 32 | row) of unpaired rows
  u64 *uspine;
  u64 *ucnt;   // [n_groups] unpaired rows per group
  i64 *seg;    // [n_groups+1] first sorted position of each group
};

struct OverlapWS {
  SortBufs sb[2];  // sort double buffers + per-row temporaries (one-shot calls only: prepared sets bring their own)
  SetWS s[2];
  RadixScratch radix;
  LayoutScratch layout;
  // per-group tables, n_groups+1 entries each
  i64 *cumA, *cumB;      // group boundaries in A-space / B-space (scan values)
  i64 *deltaA, *deltaB;  // output row = scan-space index + delta[group]
  i64 *ubase1, *ubase2;  // first output row of the group's unpaired blocks
  i64 *ustart1, *ustart2;  // first index of the group in the partitioned unpaired lists
  i64 *partA, *partB;    // merge-path partition of the emit blocks
  i32 *pgrpA, *pgrpB;    // group of the first output of every emit block
  i64 *header;           // n_rows, n_pairs, nU1, nU2
  u64 *nU;               // [2] device-side unpaired totals (sort sizes)
  i64 *ownpos;           // [4] owned sorted-position ranges of a query-sharded call: set 1 [0,1), set 2 [2,3)
};

// Everything the count / emit stages need besides the sorted sets themselves.
static void layout_count(Arena& a, i64 n1, i64 n2, i32 nb, i32 how, OverlapWS* w) {
  const i64 n[2] = {n1, n2};
  const bool need[2] = {(how & 1) != 0, (how & 2) != 0};
  for (int k = 0; k < 2; ++k) {
    SetWS& s = w->s[k];
    s.lo = a.take<u32>(n[k]);
    s.loc = a.take<u32>(n[k]);
    s.tilepre = a.take<u64>(ceil_div(n[k], SEARCH_MIN_ROWS) + 2);
    s.shift = 0;
    s.spine = a.take<u64>(scan_tiles(n[k]) + 1);
    s.bracket = a.take<i64>(ceil_div(n[k], SEARCH_MIN_ROWS) + 2);
    s.seg = a.take<i64>(nb + 1);
    s.ucnt = a.take<u64>(nb + 1);
    // set k needs flags / unpaired lists when it is on the kept side; the OTHER
    // set then needs the running max of its end'.
    s.pmax = need[1 - k] ? a.take<u64>(n[k]) : nullptr;
    if (need[k]) {
      s.rowflag = a.take<u8>(n[k]);
      s.ukeyA = a.take<u64>(n[k]);
      s.ukeyB = a.take<u64>(n[k]);
      s.uspine = a.take<u64>(scan_tiles(n[k]) + 1);
    } else {
      s.rowflag = nullptr;
      s.ukeyA = s.ukeyB = s.uspine = nullptr;
    }
  }
  w->radix = take_radix_scratch(a, (u64)(n1 > n2 ? n1 : n2));
  w->cumA = a.take<i64>(nb + 1);
  w->cumB = a.take<i64>(nb + 1);
  w->deltaA = a.take<i64>(nb + 1);
  w->deltaB = a.take<i64>(nb + 1);
  w->ubase1 = a.take<i64>(nb + 1);
  w->ubase2 = a.take<i64>(nb + 1);
  w->ustart1 = a.take<i64>(nb + 1);
  w->ustart2 = a.take<i64>(nb + 1);
  w->partA = a.take<i64>(EMIT_CHUNK_BLOCKS + 1);
  w->partB = a.take<i64>(EMIT_CHUNK_BLOCKS + 1);
  w->pgrpA = a.take<i32>(EMIT_CHUNK_BLOCKS + 1);
  w->pgrpB = a.take<i32>(EMIT_CHUNK_BLOCKS + 1);
  w->header = a.take<i64>(8);
  w->nU = a.take<u64>(2);
  w->ownpos = a.take<i64>(4);
}

// One-shot call (bf_overlap_count): the sort buffers of both sets live in the same workspace.
static void layout_overlap(Arena& a, i64 n1, i64 n2, i32 nb, i32 how, OverlapWS* w) {
  take_sort_bufs(a, n1, &w->sb[0]);
  take_sort_bufs(a, n2, &w->sb[1]);
  w->layout = take_layout_scratch(a, nb);
  layout_count(a, n1, n2, nb, how, w);
}

struct OverlapPlan {  // lives in the caller's bf_plan
  u64 magic;
  i64 n1, n2;
  i32 nb, how, closed, ready;
  KeyLayout L;
  i64 n_rows, n_pairs, nU1, nU2;
  i64 totA, totB;
  i32 shift1, shift2;  // log2(rows per search tile) of the two sides
  i64 row_off1, row_off2;  // added to every reported row number of set 1 / set 2 (query-sharded callers)
  i32 prepared;            // 1: produced by bf_overlap_count_prepared (workspace without sort buffers)
  u32 n_own1;              // rows of set 1 at and above this local row number are halo rows ...
  const i64* halo_ids;     // ... whose global row numbers are halo_ids[row - n_own1] (device; query-sharded callers)
  // sorted results (pointers into the workspace)
  u64 *key1, *key2;
  u32 *id1, *id2;
  u64 *ukeys1, *ukeys2;
};
static_assert(sizeof(OverlapPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 OVERLAP_MAGIC = 0x62664f7665726c70ull;

// ---------------------------------------------------------------- kernels

// inclusive running max of ce(key) in sorted order (np.maximum.accumulate over
// end', group resets come free because the group rank is the high field)
__global__ void __launch_bounds__(SC_THREADS) tile_max_ce_kernel(const u64* __restrict__ keys, u64 n,
                                                                 KeyLayout L, u64* __restrict__ spine) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    u64 i = base + (u64)j * SC_THREADS + threadIdx.x;
    if (i < n) {
      u64 v = key_ce(keys[i], L);
      acc = v > acc ? v : acc;
    }
  }
  acc = block_max_u64(acc, s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(SC_THREADS) running_max_ce_kernel(const u64* __restrict__ keys, u64 n,
                                                                    KeyLayout L, const u64* __restrict__ spine,
                                                                    u64* __restrict__ pmax) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE + (u64)threadIdx.x * SC_IPT;
  u64 v[SC_IPT];
  u64 mine = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    u64 i = base + j;
    v[j] = i < n ? key_ce(keys[i], L) : 0ull;
    mine = v[j] > mine ? v[j] : mine;
  }
  u64 before = block_excl_max_u64(mine, s_tmp);
  u64 run = spine[blockIdx.x];
  run = before > run ? before : run;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    u64 i = base + j;
    run = v[j] > run ? v[j] : run;
    if (i < n) pmax[i] = run;
  }
}

// Output offsets of a side, two levels: offset of the row's search tile (u64,
// scanned by one small kernel) + offset of the row inside its tile (u32, block
// scan in the search kernel).  No device-wide scan ever touches the rows again,
// and no block waits on another (a fused look-back made every tile wait for
// the slowest of its predecessors: 15-20 % of the search time, profiles/).
struct ScanView {
  const u64* tilepre;
  const u32* loc;
  i64 n;
  int shift;
  __device__ __forceinline__ i64 at(i64 p) const {
    return p >= n ? (i64)tilepre[(n + (1ll << shift) - 1) >> shift] : (i64)(tilepre[p >> shift] + loc[p]);
  }
};

// Search + count for one side.
//   B_SIDE == false: rows of set 1 against starts of set 2:
//       lo = #(cs2 <  cs1), hi = #(cs2 <  ce1)   (closed: <=)     arrops.py:338-339
//   B_SIDE == true : rows of set 2 against starts of set 1:
//       lo = #(cs1 <= cs2), hi = #(cs1 <  ce2)   (closed: <=)     arrops.py:341-342
// Both arrays are sorted, so the lo positions of a block of consecutive rows
// form a window [bracket[b], bracket[b+1]] of the other set; a tiny partition
// kernel finds all windows up front (one full binary search per block boundary,
// all in parallel).  The rows per block (256 * IPT) are chosen on the host from
// the size ratio of the two sets so that the expected window is a few thousand
// entries.  A block stages its window (+ slack for the hi searches) in shared
// memory as 32-bit offsets; every thread owns IPT consecutive rows: one
// bit-step search for the first row's lo, a short linear advance for the next
// rows (lo is monotone), and a branch-free bit-step search for every hi (same
// trip count in every lane: no divergence).  Counts stay in registers for the
// fused exclusive scan (tile scan + look-back over tile totals).
constexpr int SR_STAGE = 6144;  // staged composite starts of the other set, 32-bit relative (24 KB)
constexpr int SR_SLACK = 512;
constexpr int SR_MIN_ROWS = 256;  // smallest block of rows (IPT = 1)

static inline i64 search_tiles(i64 n, int ipt) { return ceil_div(n, (i64)SC_THREADS * ipt); }
// rows per thread for a side with n rows searching m entries
static inline int pick_search_ipt(i64 n, i64 m) {
  if (n <= 0) return 8;
  const double ratio = (double)m / (double)n;
  int ipt = 4;  // 8 rows per thread spill at 40 registers (6 blocks/SM): 1.52 -> 1.45 ms for 1e8 rows; 2 rows: 1.82 ms
  while (ipt > 1 && (double)(SC_THREADS * ipt) * ratio > 3072.0) ipt >>= 1;
  return ipt;
}

template <bool B_SIDE>
__global__ void __launch_bounds__(256) search_partition_kernel(const u64* __restrict__ mine, i64 n,
                                                               const u64* __restrict__ other, i64 m, KeyLayout L,
                                                               i64 tile_rows, i64 ntiles, i64* __restrict__ bracket) {
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b > ntiles) return;
  i64 row = b * tile_rows;
  if (row > n - 1) row = n - 1;
  const u64 x = key_cs(mine[row], L);
  bracket[b] = B_SIDE ? lower_bound_by(0, m, [&](i64 q) { return key_cs(other[q], L) <= x; })
                      : lower_bound_by(0, m, [&](i64 q) { return key_cs(other[q], L) < x; });
}

// Staged values are (cs - base_cs + 1) clamped to 32 bits, probes likewise, so
// a probe below the window maps to 0 and one far above it to 0xffffffff; the
// window is only staged when its span fits, which keeps every comparison exact.
__device__ __forceinline__ u32 rel32(u64 x, u64 base_cs) {
  if (x < base_cs) return 0u;
  const u64 r = x - base_cs + 1;
  return r > 0xffffffffull ? 0xffffffffu : (u32)r;
}

// hi on the full keys, starting from position `from` (gallop, then bisect).  Rare path (a row whose
// partners run past the staged entries, or a block whose window could not be staged): kept out of line
// so that the unrolled search loop stays small enough for the instruction cache.
__device__ __noinline__ i64 gallop_hi(const u64* __restrict__ other, i64 from, i64 m, u64 xe, int closed,
                                         const KeyLayout& L) {
  i64 ga = from, gb = from, gstep = 1;
  for (;;) {
    gb = ga + gstep;
    if (gb >= m) {
      gb = m;
      break;
    }
    const u64 v = key_cs(other[gb - 1], L);
    if (!(closed ? (v <= xe) : (v < xe))) break;
    ga = gb;
    gstep <<= 1;
  }
  return closed ? lower_bound_by(ga, gb, [&](i64 q) { return key_cs(other[q], L) <= xe; })
                : lower_bound_by(ga, gb, [&](i64 q) { return key_cs(other[q], L) < xe; });
}

template <bool B_SIDE>
__device__ __noinline__ i64 unstaged_lo(const u64* __restrict__ other, i64 br_lo, i64 br_hi, u64 xs, const KeyLayout& L) {
  return B_SIDE ? lower_bound_by(br_lo, br_hi, [&](i64 q) { return key_cs(other[q], L) <= xs; })
                : lower_bound_by(br_lo, br_hi, [&](i64 q) { return key_cs(other[q], L) < xs; });
}

template <bool B_SIDE, bool FLAGS, int IPT>
__global__ void __launch_bounds__(SC_THREADS, 6)
    search_count_kernel(const u64* __restrict__ mine, const u32* __restrict__ mine_id, i64 n,
                        const u64* __restrict__ other, i64 m, const u64* __restrict__ other_pmax,
                        const i64* __restrict__ bracket, KeyLayout L, int closed, u32* __restrict__ lo_out,
                        u32* __restrict__ loc_out, u64* __restrict__ tile_tot, i64* __restrict__ overflow,
                        u8* __restrict__ rowflag, const i64* __restrict__ ownpos) {
  constexpr int TILE = SC_THREADS * IPT;
  __shared__ u32 s_rel[SR_STAGE];
  __shared__ u64 s_tmp[8];
  const u32 tid = threadIdx.x;
  const i64 tile = blockIdx.x;
  if (ownpos) {
    // query-sharded call: a tile without a single owned row has nothing to count (its offsets and its tile
    // total were zeroed by the host wrapper) -- 7 of 8 tiles of the target side on 8 ranks
    const i64 own_lo = ownpos[B_SIDE ? 2 : 0], own_hi = ownpos[B_SIDE ? 3 : 1];
    if (tile * TILE >= own_hi || (tile + 1) * TILE <= own_lo) return;
  }
  const i64 base = tile * TILE + (i64)tid * IPT;  // this thread's first row
  // own rows: IPT consecutive keys per thread, 16-byte loads
  u64 key[IPT];
  if (base + IPT <= n) {
    if (IPT >= 2) {
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(mine + base);
#pragma unroll
      for (int j = 0; j < IPT / 2; ++j) {
        const ulonglong2 v = src[j];
        key[2 * j] = v.x;
        key[2 * j + 1] = v.y;
      }
    } else {
      key[0] = mine[base];
    }
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) key[j] = base + j < n ? mine[base + j] : 0ull;
  }
  const i64 br_lo = bracket[tile], br_hi = bracket[tile + 1];
  const u32 W = (u32)(br_hi - br_lo);  // lo candidates: other[br_lo .. br_hi)
  u32 nst = 0;                         // staged entries: other[br_lo .. br_lo + nst)
  u64 base_cs = 0;
  if ((i64)W + 1 <= SR_STAGE && br_lo < m) {
    i64 want = (i64)W + 1 + SR_SLACK;
    if (want > SR_STAGE) want = SR_STAGE;
    if (want > m - br_lo) want = m - br_lo;
    const u64 first = key_cs(other[br_lo], L), last = key_cs(other[br_lo + want - 1], L);
    if (last - first < 0xfffffff0ull) {
      nst = (u32)want;
      base_cs = first;
    }
  }
#pragma unroll 4
  for (u32 i = tid; i < nst; i += SC_THREADS) s_rel[i] = (u32)(key_cs(other[br_lo + i], L) - base_cs + 1);
  // largest powers of two <= W and <= nst (0 for 0); the shift loops these replace were 13 % of the kernel's
  // executed instructions (ncu source page, r1 capture)
  const u32 pow2w = W ? 1u << (31 - __clz(W)) : 0u;
  const u32 pow2n = nst ? 1u << (31 - __clz(nst)) : 0u;
  __syncthreads();

  u32 cnt[IPT];
  u32 lo32[IPT];
  const u32 rel_base = smem_addr(s_rel);
  u32 pos = 0;  // lo of the previous row, relative to br_lo (staged path)
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    const i64 p = base + j;
    cnt[j] = 0;
    lo32[j] = 0;
    if (p < n) {
      const u64 k = key[j];
      const u64 xs = key_cs(k, L), xe = key_ce(k, L);
      i64 lo, hi;
      if (nst > 0) {
        const u32 xs_rel = rel32(xs, base_cs), xe_rel = rel32(xe, base_cs);
        auto rel_at = [&](u32 i) { return lds_u32(rel_base + 4u * i); };  // s_rel[i]
        bool searched = false;
        if (j > 0) {  // lo is monotone over consecutive rows: short linear advance
          const u32 lim = pos + 6 < W ? pos + 6 : W;
          while (pos < lim && (B_SIDE ? (rel_at(pos) <= xs_rel) : (rel_at(pos) < xs_rel))) ++pos;
          searched = pos < lim || pos == W || !(B_SIDE ? (rel_at(pos) <= xs_rel) : (rel_at(pos) < xs_rel));
        }
        if (!searched) {  // bit-step lower bound over the W candidates
          u32 q = 0;
          for (u32 step = pow2w; step > 0; step >>= 1) {
            const u32 nx = q + step;
            if (nx <= W) {
              const u32 v = rel_at(nx - 1);
              if (B_SIDE ? (v <= xs_rel) : (v < xs_rel)) q = nx;
            }
          }
          pos = q;
        }
        lo = br_lo + pos;
        // hi >= lo.  Nearly every row has fewer than 64 partners: probe the entry 64 ahead, then six
        // bit-steps inside that window; rows with more partners (or near the end of the staged
        // entries) take the full bit-step search from lo.
        u32 h = pos;
        const u32 probe = pos + 64;
        bool near = false;
        if (probe <= nst) {
          const u32 v = rel_at(probe - 1);
          near = !(closed ? (v <= xe_rel) : (v < xe_rel));
        }
        if (near) {
#pragma unroll
          for (u32 step = 32; step > 0; step >>= 1) {
            const u32 v = rel_at(h + step - 1);
            if (closed ? (v <= xe_rel) : (v < xe_rel)) h += step;
          }
        } else {
          for (u32 step = pow2n; step > 0; step >>= 1) {
            const u32 nx = h + step;
            if (nx <= nst) {
              const u32 v = rel_at(nx - 1);
              if (closed ? (v <= xe_rel) : (v < xe_rel)) h = nx;
            }
          }
        }
        hi = br_lo + h;
        if (h == nst && hi < m) hi = gallop_hi(other, hi, m, xe, closed, L);  // ran off the staged window
      } else {
        lo = unstaged_lo<B_SIDE>(other, br_lo, br_hi, xs, L);
        hi = gallop_hi(other, lo, m, xe, closed, L);
      }
      BF_DEVICE_CHECK(lo >= 0 && lo <= hi && hi <= m);
      cnt[j] = (u32)(hi - lo);
      lo32[j] = (u32)lo;
    }
  }
  if (ownpos) {
    // query-sharded call: only rows inside this shard's owned range of sorted positions produce output --
    // queries outside it are halo rows (they only serve as partners in the B block), targets outside it
    // have their B block written by the shard that owns them
    const i64 own_lo = ownpos[B_SIDE ? 2 : 0], own_hi = ownpos[B_SIDE ? 3 : 1];
#pragma unroll
    for (int j = 0; j < IPT; ++j)
      if (base + j < own_lo || base + j >= own_hi) {
        cnt[j] = 0;
        lo32[j] = 0xffffffffu;  // marks "not owned" for the unpaired-row check below
      }
  }
  if (FLAGS) {
    // rows without a pair of their own still have a partner if an earlier interval of the other set
    // reaches past their start (running max of end').  Done after the search loop
    // so that the loads of all IPT rows are in flight together instead of one per dependent iteration.
    u64 pm[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const bool ask = base + j < n && cnt[j] == 0 && lo32[j] > 0 && lo32[j] != 0xffffffffu;
      pm[j] = ask ? other_pmax[lo32[j] - 1] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const i64 p = base + j;
      if (p < n && cnt[j] == 0 && lo32[j] != 0xffffffffu) {
        const u64 xs = key_cs(key[j], L);
        const bool partner = lo32[j] > 0 && (closed ? (pm[j] >= xs) : (pm[j] > xs));
        if (!partner) rowflag[mine_id[p]] = 1;  // per-group counts come later from the compacted list
      }
    }
  }
  // ---- exclusive scan of the counts inside the tile; the tile total goes to the spine ----
  u64 mine_sum = 0;
#pragma unroll
  for (int j = 0; j < IPT; ++j) mine_sum += cnt[j];
  u64 tot;
  u32 run = (u32)block_excl_sum_u64(mine_sum, s_tmp, &tot);  // < 2^32: TILE rows x < 2^31 partners, checked below
  if (tid == 0) {
    tile_tot[tile] = tot;
    if (tot >> 32) *overflow = 1;  // in-tile offsets are 32-bit
  }
  if (base + IPT <= n && IPT >= 4) {
    // 16-byte stores: the thread's IPT offsets and IPT positions are contiguous
    u32 sc[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      sc[j] = run;
      run += cnt[j];
    }
    uint4* so = reinterpret_cast<uint4*>(loc_out + base);
    uint4* lo4 = reinterpret_cast<uint4*>(lo_out + base);
#pragma unroll
    for (int j = 0; j < IPT / 4; ++j) {
      so[j] = make_uint4(sc[4 * j], sc[4 * j + 1], sc[4 * j + 2], sc[4 * j + 3]);
      lo4[j] = make_uint4(lo32[4 * j], lo32[4 * j + 1], lo32[4 * j + 2], lo32[4 * j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const i64 p = base + j;
      if (p < n) {
        loc_out[p] = run;
        lo_out[p] = lo32[j];
      }
      run += cnt[j];
    }
  }
}

// Per-group output offsets; single thread (n_groups is tens, not millions).
__global__ void offset_table_kernel(OverlapWS w, i64 n1, i64 n2, i32 nb, i32 how) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool need1 = how & 1, need2 = how & 2;
  const ScanView scanA = {w.s[0].tilepre, w.s[0].loc, n1, w.s[0].shift};
  const ScanView scanB = {w.s[1].tilepre, w.s[1].loc, n2, w.s[1].shift};
  i64 base = 0, runU1 = 0, runU2 = 0, pairs = 0;
  for (i32 c = 0; c < nb; ++c) {
    i64 a0 = scanA.at(w.s[0].seg[c]), a1 = scanA.at(w.s[0].seg[c + 1]);
    i64 b0 = scanB.at(w.s[1].seg[c]), b1 = scanB.at(w.s[1].seg[c + 1]);
    i64 U1 = need1 ? (i64)w.s[0].ucnt[c] : 0;
    i64 U2 = need2 ? (i64)w.s[1].ucnt[c] : 0;
    w.cumA[c] = a0;
    w.cumB[c] = b0;
    w.deltaA[c] = base - a0;
    w.deltaB[c] = base + (a1 - a0) - b0;
    w.ubase1[c] = base + (a1 - a0) + (b1 - b0);
    w.ubase2[c] = w.ubase1[c] + U1;
    w.ustart1[c] = runU1;
    w.ustart2[c] = runU2;
    base += (a1 - a0) + (b1 - b0) + U1 + U2;
    pairs += (a1 - a0) + (b1 - b0);
    runU1 += U1;
    runU2 += U2;
  }
  w.cumA[nb] = scanA.at(n1);
  w.cumB[nb] = scanB.at(n2);
  w.deltaA[nb] = w.deltaB[nb] = 0;
  w.ubase1[nb] = w.ubase2[nb] = base;
  w.ustart1[nb] = runU1;
  w.ustart2[nb] = runU2;
  w.header[0] = base;
  w.header[1] = pairs;
  w.header[2] = runU1;
  w.header[3] = runU2;
  w.header[4] = scanA.at(n1);
  w.header[5] = scanB.at(n2);
}

// Emit one block (A or B) of pairs:  A: (id1[p], id2[lo[p] + k])   B: (id1[lo[q] + k], id2[q]).
// Load-balanced over the merged sequence "owner row, its outputs, next owner
// row, ..." (merge path): a block takes 2048 consecutive positions of that
// sequence, so it handles at most 2048 owner rows AND at most 2048 outputs no
// matter how skewed the fan-out is.  Two binary searches per block find its
// rows; their scan values, lo and ids are staged in shared memory; each owner
// marks the slot of its first output and a block-wide running max turns the
// marks into "owner of every output slot" (the np.repeat of arrops.py:359-366
// without a per-output search).  Stores are 8-byte, coalesced across the warp.
constexpr int EM_THREADS = 256;
constexpr int EM_IPT = 8;
constexpr int EM_TILE = EM_THREADS * EM_IPT;
constexpr int EM_AHEAD = 148 * 8;  // blocks resident at a time (8 per SM)

// part[b] = owner rows consumed before diagonal b * EM_TILE = #{p : p + scan[p] < D};
// pgrp[b] = group of the first output of block b.  One thread per block
// boundary, so the ~27-step searches of all blocks run side by side instead of
// at the head of every emit block.
__global__ void __launch_bounds__(256) emit_partition_kernel(const ScanView scan, i64 n_owner, i64 total,
                                                             i64 block0, i64 nparts, const i64* __restrict__ cum,
                                                             i32 nb, i64* __restrict__ part, i32* __restrict__ pgrp) {
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nparts) return;
  const i64 W = n_owner + total;
  i64 D = (block0 + b) * EM_TILE;
  if (D > W) D = W;
  const i64 p = lower_bound_by(0, n_owner, [&](i64 q) { return q + scan.at(q) < D; });
  part[b] = p;
  const i64 t = D - p;  // outputs before the diagonal
  pgrp[b] = (i32)(lower_bound_by(0, nb, [&](i64 c) { return cum[c] <= t; }) - 1);  // last c with cum[c] <= t
}

struct EmitItem {  // one owner row staged in shared memory
  u32 src;  // lo - (first output slot of the row in this block), wrapping: partner = other_id[src + slot]
  u32 id;   // row id of the owner
};

template <bool B_SIDE>
__global__ void __launch_bounds__(EM_THREADS, 8)
    emit_pairs_kernel(const ScanView scan, const u32* __restrict__ lo, i64 n_owner,
                      const u32* __restrict__ owner_id, const u32* __restrict__ other_id, i64 total,
                      i64 block0, const i64* __restrict__ part, const i32* __restrict__ pgrp,
                      const i64* __restrict__ cum, const i64* __restrict__ delta, i32 nb, i64 off1, i64 off2,
                      u32 n_own1, const i64* __restrict__ halo_ids, i64 n_other, i64 n_rows, i64* __restrict__ ev1,
                      i64* __restrict__ ev2) {
  __shared__ EmitItem s_item[EM_TILE + 2];
  // global row number of a set-1 row listed as a partner (B side only: owners of the A side are never halo rows)
  auto row1 = [&](u32 id) -> i64 { return id < n_own1 ? (i64)id + off1 : halo_ids[id - n_own1]; };
  __shared__ __align__(16) u16 s_owner[EM_TILE];  // owner (index into s_item) of every output slot
  __shared__ u32 s_wmax[EM_THREADS / 32];
  const u32 tid = threadIdx.x;
  const i64 W = n_owner + total;
  const i64 D0 = (block0 + (i64)blockIdx.x) * EM_TILE;
  const i64 D1 = D0 + EM_TILE < W ? D0 + EM_TILE : W;
  // every thread reads the block's partition entries itself (broadcast loads, no barrier)
  const i64 p0 = part[blockIdx.x], p1 = part[blockIdx.x + 1];
  const i32 c_lo = pgrp[blockIdx.x], c_hi = pgrp[blockIdx.x + 1];  // c_hi: upper bound of the last output's group
  const i64 delta_lo = delta[c_lo];
  const i64 t0 = D0 - p0, t1 = D1 - p1;
  const i32 nout = (i32)(t1 - t0);
  // the block that will take over this SM slot about one wave from now: its partition entries are read
  // here and its owner rows are pulled into L2 further down, off this block's critical path
  const bool ahead = blockIdx.x + EM_AHEAD < gridDim.x;
  i64 q0 = 0, q1 = 0;
  if (ahead) {
    q0 = part[blockIdx.x + EM_AHEAD];
    q1 = part[blockIdx.x + EM_AHEAD + 1];
  }
  if (nout <= 0) return;
  const i64 first = p0 > 0 ? p0 - 1 : 0;  // owner of the first output may have started earlier
  const i32 nitems = (i32)(p1 - first);
  // scan[row] - t0 clamped to 32 bits: only offsets in [0, nout] are ever used exactly; the first owner
  // may have started up to 2^31 - 1 outputs earlier (a row has < 2^31 partners), later owners may start
  // far ahead
  auto rel_at = [&](i64 p) -> i32 {
    const i64 rel64 = scan.at(p) - t0;
    return rel64 > 0x7fffffffLL ? 0x7fffffff : (i32)rel64;
  };
  // An owner row marks the slot of its first output; its successor's offset (same cache line, read by
  // the thread itself) tells whether it has any.  The first round of loads is in flight across the barrier
  // that separates the clearing of the slots from the marks.
  i32 i = (i32)tid;
  i32 r = 0, r_next = 0;
  EmitItem it = {0u, 0u};
  auto load_row = [&]() {
    r = rel_at(first + i);
    r_next = rel_at(first + i + 1);
    it.src = lo[first + i] - (u32)r;
    it.id = owner_id[first + i];
  };
  if (i < nitems) load_row();
#define EM_PAD(x) (x)
  reinterpret_cast<uint4*>(s_owner)[tid] = make_uint4(0u, 0u, 0u, 0u);  // 8 slots per thread
  __syncthreads();
  if (ahead) {
    const i64 f = (q0 > 0 ? q0 - 1 : 0) & ~(i64)31;  // 128-byte lines of 32 rows
    const i64 row = f + (i64)tid * 32;
    if (row <= q1 && row < n_owner) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(scan.loc + row));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(lo + row));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(owner_id + row));
    }
  }
  while (i < nitems) {
    s_item[i] = it;
    if (r_next > r && r >= 0 && r < nout) s_owner[r] = (u16)i;
    i += EM_THREADS;
    if (i < nitems) load_row();
  }
  __syncthreads();
  {  // inclusive running max over the 2048 slots, 8 consecutive slots (one 16-byte word) per thread
    const uint4 q = reinterpret_cast<const uint4*>(s_owner)[tid];
    const u32 word[4] = {q.x, q.y, q.z, q.w};
    u32 v[EM_IPT];
    u32 run = 0;
#pragma unroll
    for (int j = 0; j < EM_IPT; ++j) {
      const u32 x = (word[j >> 1] >> (16 * (j & 1))) & 0xffffu;
      run = x > run ? x : run;
      v[j] = run;
    }
    u32 inc = run;
    const u32 lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc = up > inc ? up : inc;
    }
    if (lane == 31) s_wmax[w] = inc;
    u32 before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) before = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < EM_THREADS / 32; ++i)
      if ((u32)i < w) before = s_wmax[i] > before ? s_wmax[i] : before;
    u32 out[4];
#pragma unroll
    for (int j = 0; j < EM_IPT; j += 2) {
      const u32 a = v[j] > before ? v[j] : before, b = v[j + 1] > before ? v[j + 1] : before;
      out[j >> 1] = a | (b << 16);
    }
    reinterpret_cast<uint4*>(s_owner)[tid] = make_uint4(out[0], out[1], out[2], out[3]);
  }
  __syncthreads();
  auto group_of = [&](i64 t, i64 a, i64 b) {  // last c with cum[c] <= t
    return lower_bound_by(a, b, [&](i64 c) { return cum[c] <= t; }) - 1;
  };
  // outputs of one group (nearly every block): row = t0 + delta[group] + slot
  const bool one_group = c_lo == c_hi;
  if (one_group) {
    // Two adjacent output rows per thread and step, written with 16-byte stores.  The pairs are aligned
    // to even output rows (16-byte aligned addresses): when the block's first row is odd, slot 0 is a
    // single that thread 0 writes on its own.
    const i64 row0 = t0 + delta_lo;
    BF_DEVICE_CHECK(row0 >= 0 && row0 + nout <= n_rows);
    i64* const out1 = ev1 + row0;
    i64* const out2 = ev2 + row0;
    const i32 par = (i32)(row0 & 1);
    auto slot = [&](i32 x, u32& own, u32& oth) {  // x clamped to a valid slot by the caller
      const EmitItem it = s_item[s_owner[EM_PAD(x)]];
      own = it.id;
      BF_DEVICE_CHECK((i64)(u32)(it.src + (u32)x) < n_other && (i64)s_owner[EM_PAD(x)] < nitems);
      oth = other_id[it.src + (u32)x];
    };
    if (par && tid == 0) {
      u32 own, oth;
      slot(0, own, oth);
      out1[0] = B_SIDE ? row1(oth) : (i64)own + off1;
      out2[0] = (B_SIDE ? (i64)own : (i64)oth) + off2;
    }
    const i32 last = nout - 1;
#pragma unroll 1
    for (int h = 0; h < EM_IPT / 2; h += 2) {
      u32 own[4], oth[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const i32 x = par + 2 * ((h + j) * EM_THREADS + (i32)tid);
        slot(min(x, last), own[2 * j], oth[2 * j]);
        slot(min(x + 1, last), own[2 * j + 1], oth[2 * j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const i32 x = par + 2 * ((h + j) * EM_THREADS + (i32)tid);
        if (x > last) continue;
        const i64 a0 = B_SIDE ? row1(oth[2 * j]) : (i64)own[2 * j] + off1;
        const i64 b0 = (B_SIDE ? (i64)own[2 * j] : (i64)oth[2 * j]) + off2;
        if (x + 1 <= last) {
          const i64 a1 = B_SIDE ? row1(oth[2 * j + 1]) : (i64)own[2 * j + 1] + off1;
          const i64 b1 = (B_SIDE ? (i64)own[2 * j + 1] : (i64)oth[2 * j + 1]) + off2;
          *reinterpret_cast<longlong2*>(out1 + x) = make_longlong2(a0, a1);
          *reinterpret_cast<longlong2*>(out2 + x) = make_longlong2(b0, b1);
        } else {
          out1[x] = a0;
          out2[x] = b0;
        }
      }
    }
  } else {
    // a block whose outputs span a group boundary: the row offset changes inside the block
#pragma unroll 1
    for (int h = 0; h < EM_IPT; ++h) {
      const i32 x = h * EM_THREADS + (i32)tid;
      if (x < nout) {
        const EmitItem it = s_item[s_owner[EM_PAD(x)]];
        BF_DEVICE_CHECK((i64)(u32)(it.src + (u32)x) < n_other);
        const u32 own = it.id, oth = other_id[it.src + (u32)x];
        const i64 t = t0 + x;
        const i64 row = t + delta[group_of(t, c_lo, (i64)c_hi + 1)];
        BF_DEVICE_CHECK(row >= 0 && row < n_rows);
        ev1[row] = B_SIDE ? row1(oth) : (i64)own + off1;
        ev2[row] = (B_SIDE ? (i64)own : (i64)oth) + off2;
      }
    }
  }
#undef EM_PAD
}

// ---- unpaired rows: compaction in row order -> stable partition by group ----
// Two launches around a spine scan (count per tile, then write), no look-back: a
// single-pass chained scan over 1-byte flags is bound by the latency of the
// chain (one L2 round trip per 32 tiles), not by the 0.1 GB it reads.
constexpr int UP_BYTES = 64;                       // flag bytes per thread (4 x 16-byte loads)
constexpr int UP_TILE = SC_THREADS * UP_BYTES;     // 16384 rows per block
static inline i64 unpaired_tiles(i64 n) { return ceil_div(n, UP_TILE); }

__device__ __forceinline__ void load_flags64(const u8* __restrict__ flag, u64 base, u64 n, u64 (&f)[8]) {
  if (base + UP_BYTES <= n) {
    const uint4* src = reinterpret_cast<const uint4*>(flag + base);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = src[q];
      f[2 * q] = (u64)v.x | ((u64)v.y << 32);
      f[2 * q + 1] = (u64)v.z | ((u64)v.w << 32);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      f[q] = 0;
      for (int j = 0; j < 8; ++j) {
        const u64 i = base + q * 8 + j;
        if (i < n) f[q] |= (u64)flag[i] << (8 * j);
      }
    }
  }
}
__global__ void __launch_bounds__(SC_THREADS)
    unpaired_count_kernel(const u8* __restrict__ flag, u64 n, u64* __restrict__ spine) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * UP_TILE + (u64)threadIdx.x * UP_BYTES;
  u64 f[8];
  load_flags64(flag, base, n, f);
  u64 c = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) c += (u64)__popcll(f[q]);
  c = block_sum_u64(c, s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = c;
}
__global__ void __launch_bounds__(SC_THREADS)
    unpaired_write_kernel(const u8* __restrict__ flag, const i32* __restrict__ grp, u64 n,
                          const u64* __restrict__ spine, u64* __restrict__ ukeys) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * UP_TILE + (u64)threadIdx.x * UP_BYTES;
  u64 f[8];
  load_flags64(flag, base, n, f);
  u64 c = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) c += (u64)__popcll(f[q]);
  u64 tot;
  u64 at = spine[blockIdx.x] + block_excl_sum_u64(c, s_tmp, &tot);
  if (c == 0) return;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    u64 bits = f[q];
    while (bits) {
      const int j = (__ffsll((long long)bits) - 1) >> 3;  // flags are 0/1 bytes
      bits &= ~(0xffull << (8 * j));
      const u64 i = base + q * 8 + j;
      const u64 g = grp ? (u64)(u32)grp[i] : 0ull;
      ukeys[at++] = (g << 32) | i;
    }
  }
}
// unpaired rows per group from the list sorted by (group, row): ucnt[c] = #{keys with group c}
__global__ void __launch_bounds__(128) unpaired_group_counts_kernel(const u64* __restrict__ ukeys,
                                                                    const u64* __restrict__ nU, i32 nb,
                                                                    u64* __restrict__ ucnt) {
  const i32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb) return;
  const i64 n = (i64)*nU;
  const i64 a = lower_bound_by(0, n, [&](i64 m) { return (ukeys[m] >> 32) < (u64)c; });
  const i64 b = lower_bound_by(a, n, [&](i64 m) { return (ukeys[m] >> 32) <= (u64)c; });
  ucnt[c] = (u64)(b - a);
}

template <bool SECOND>
__global__ void __launch_bounds__(256) emit_unpaired_kernel(const u64* __restrict__ ukeys, i64 nU,
                                                            const i64* __restrict__ ustart,
                                                            const i64* __restrict__ ubase, i64 row_off,
                                                            i64* __restrict__ ev1, i64* __restrict__ ev2) {
  i64 u = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nU) return;
  u64 k = ukeys[u];
  u32 c = (u32)(k >> 32);
  i64 row = (i64)(k & 0xffffffffull);
  i64 out = ubase[c] + (u - ustart[c]);
  BF_DEVICE_CHECK(out >= 0);
  ev1[out] = SECOND ? -1 : row + row_off;
  ev2[out] = SECOND ? row + row_off : -1;
}

template <bool B_SIDE, bool FLAGS, int IPT>
static int launch_search_ipt(const SortResult& mine, i64 n, const SortResult& other, i64 m, SetWS& sw,
                             const u64* other_pmax, const KeyLayout& L, int closed, i64* overflow, const i64* ownpos,
                             cudaStream_t st) {
  const i64 tiles = search_tiles(n, IPT);
  if (ownpos) {  // tiles outside the owned range return at once: their outputs are these zeros
    BF_CUDA(cudaMemsetAsync(sw.loc, 0, (size_t)n * sizeof(u32), st));
    BF_CUDA(cudaMemsetAsync(sw.tilepre, 0, ((size_t)tiles + 1) * sizeof(u64), st));
  }
  BF_LAUNCH(PC_SEARCH, 0, st, search_partition_kernel<B_SIDE><<<(unsigned)ceil_div(tiles + 1, 256), 256, 0, st>>>(
      mine.keys, n, other.keys, m, L, (i64)SC_THREADS * IPT, tiles, sw.bracket));
  BF_LAUNCH(PC_SEARCH, 16 * n, st, (search_count_kernel<B_SIDE, FLAGS, IPT><<<(unsigned)tiles, SC_THREADS, 0, st>>>(
      mine.keys, mine.vals, n, other.keys, m, other_pmax, sw.bracket, L, closed, sw.lo, sw.loc, sw.tilepre, overflow,
      FLAGS ? sw.rowflag : nullptr, ownpos)));
  // tile totals -> exclusive tile offsets, grand total in tilepre[tiles]
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(sw.tilepre, (u64)tiles, sw.tilepre + tiles));
  int shift = 0;
  while ((1 << shift) < SC_THREADS * IPT) ++shift;
  sw.shift = shift;
  return BF_OK;
}
template <bool B_SIDE>
static int launch_search(const SortResult& mine, i64 n, const SortResult& other, i64 m, bool flags, SetWS& sw,
                         const u64* other_pmax, const KeyLayout& L, int closed, i64* overflow, const i64* ownpos,
                         cudaStream_t st) {
  if (n == 0) {
    BF_CUDA(cudaMemsetAsync(sw.tilepre, 0, 2 * sizeof(u64), st));
    sw.shift = 8;
    return BF_OK;
  }
  const int ipt = pick_search_ipt(n, m);
#define BF_SEARCH_CASE(I)                                                                                          \
  case I:                                                                                                          \
    return flags ? launch_search_ipt<B_SIDE, true, I>(mine, n, other, m, sw, other_pmax, L, closed, overflow, ownpos, st)  \
                 : launch_search_ipt<B_SIDE, false, I>(mine, n, other, m, sw, other_pmax, L, closed, overflow, ownpos, st);
  switch (ipt) {
    BF_SEARCH_CASE(8)
    BF_SEARCH_CASE(4)
    BF_SEARCH_CASE(2)
    default:
      BF_SEARCH_CASE(1)
  }
#undef BF_SEARCH_CASE
}

// Sorted-position ranges a query-sharded call owns: rows whose (group, start) lies in
// [(g_lo, s_lo), (g_hi, s_hi)) in lexicographic order.  One thread; the groups lie end to end on the linear
// axis, so the bounds become two linear coordinates and four binary searches.
__global__ void own_positions_kernel(const u64* __restrict__ key1, i64 n1, const u64* __restrict__ key2, i64 n2,
                                     KeyLayout L, i64 g_lo, i64 s_lo, i64 g_hi, i64 s_hi, i64* __restrict__ ownpos) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto linear = [&](i64 g, i64 s) -> u64 {
    if (g >= L.nb) return L.goff[L.nb];
    if (g < 0) return 0ull;
    const u64 lo = L.goff[g], hi = L.goff[g + 1];
    if (s <= (i64)L.cmin) return lo;
    const u64 off = (u64)s - L.cmin;
    return off >= hi - lo ? hi : lo + off;
  };
  const u64 c_lo = linear(g_lo, s_lo), c_hi = linear(g_hi, s_hi);
  ownpos[0] = lower_bound_by(0, n1, [&](i64 q) { return key_cs(key1[q], L) < c_lo; });
  ownpos[1] = lower_bound_by(0, n1, [&](i64 q) { return key_cs(key1[q], L) < c_hi; });
  ownpos[2] = lower_bound_by(0, n2, [&](i64 q) { return key_cs(key2[q], L) < c_lo; });
  ownpos[3] = lower_bound_by(0, n2, [&](i64 q) { return key_cs(key2[q], L) < c_hi; });
}

// Everything after the two sorts: segments, running maxima, search + count, unpaired lists, offset table.
// `own` (host, 4 words or null): (g_lo, s_lo, g_hi, s_hi) of a query-sharded call.
static int count_sorted(OverlapWS& w, const KeyLayout& L, const SortResult& r1, i64 n1, const SortResult& r2, i64 n2,
                        const i32* grp1, const i32* grp2, i32 nb, i32 how, i32 closed, const i64* own,
                        OverlapPlan* plan, cudaStream_t st, const u64* pmax1_ready = nullptr,
                        const u64* pmax2_ready = nullptr) {
  const bool need1 = how & 1, need2 = how & 2;
  plan->L = L;
  plan->key1 = r1.keys;
  plan->id1 = r1.vals;
  plan->key2 = r2.keys;
  plan->id2 = r2.vals;

  // ---- group segments ----
  const unsigned seg_grid = (unsigned)ceil_div(nb + 1, 128);
  BF_LAUNCH(PC_SEGMENT, 0, st, segment_table_kernel<<<seg_grid, 128, 0, st>>>(r1.keys, n1, L, nb, w.s[0].seg));
  BF_LAUNCH(PC_SEGMENT, 0, st, segment_table_kernel<<<seg_grid, 128, 0, st>>>(r2.keys, n2, L, nb, w.s[1].seg));
  const i64* ownpos = nullptr;
  if (own) {
    BF_LAUNCH(PC_SEGMENT, 0, st, own_positions_kernel<<<1, 32, 0, st>>>(r1.keys, n1, r2.keys, n2, L, own[0], own[1], own[2],
                                                                       own[3], w.ownpos));
    ownpos = w.ownpos;
  }

  // ---- running max of end' (only where an outer side needs it) ----
  const SortResult* rs[2] = {&r1, &r2};
  const i64 nn[2] = {n1, n2};
  const u64* pmax_ready[2] = {pmax1_ready, pmax2_ready};
  for (int k = 0; k < 2; ++k) {
    if (w.s[k].pmax && pmax_ready[k]) {  // a prepared set brought its running maximum along (bf_sorted_set_runmax)
      w.s[k].pmax = const_cast<u64*>(pmax_ready[k]);
    } else if (w.s[k].pmax && nn[k] > 0) {
      const unsigned tiles = (unsigned)scan_tiles((u64)nn[k]);
      BF_LAUNCH(PC_RUNMAX, 0, st, tile_max_ce_kernel<<<tiles, SC_THREADS, 0, st>>>(rs[k]->keys, (u64)nn[k], L, w.s[k].spine));
      BF_LAUNCH(PC_RUNMAX, 0, st, spine_scan_kernel<true><<<1, 1024, 0, st>>>(w.s[k].spine, tiles, nullptr));
      BF_LAUNCH(PC_RUNMAX, 0, st, running_max_ce_kernel<<<tiles, SC_THREADS, 0, st>>>(rs[k]->keys, (u64)nn[k], L, w.s[k].spine, w.s[k].pmax));
    }
  }

  // ---- search + count ----
  if (need1) {
    BF_CUDA(cudaMemsetAsync(w.s[0].ucnt, 0, (nb + 1) * sizeof(u64), st));
    if (n1 > 0) BF_CUDA(cudaMemsetAsync(w.s[0].rowflag, 0, (size_t)n1, st));
  }
  if (need2) {
    BF_CUDA(cudaMemsetAsync(w.s[1].ucnt, 0, (nb + 1) * sizeof(u64), st));
    if (n2 > 0) BF_CUDA(cudaMemsetAsync(w.s[1].rowflag, 0, (size_t)n2, st));
  }
  // algorithmic bytes: key read (8) + lo (4) + scan offset (8) written per row
  BF_CUDA(cudaMemsetAsync(w.header, 0, 8 * sizeof(i64), st));
  BF_TRY((launch_search<false>(r1, n1, r2, n2, need1, w.s[0], w.s[1].pmax, L, closed, w.header + 6, ownpos, st)));
  BF_TRY((launch_search<true>(r2, n2, r1, n1, need2, w.s[1], w.s[0].pmax, L, closed, w.header + 6, ownpos, st)));
  plan->shift1 = w.s[0].shift;
  plan->shift2 = w.s[1].shift;

  // ---- unpaired lists: compact in row order, partition by group ----
  BF_CUDA(cudaMemsetAsync(w.nU, 0, 2 * sizeof(u64), st));
  const i32* grps[2] = {nb > 1 ? grp1 : nullptr, nb > 1 ? grp2 : nullptr};
  u64* ukeys_sorted[2] = {nullptr, nullptr};
  for (int k = 0; k < 2; ++k) {
    const bool need = k == 0 ? need1 : need2;
    if (!need || nn[k] == 0) continue;
    const unsigned tiles = (unsigned)unpaired_tiles(nn[k]);
    BF_LAUNCH(PC_UNPAIRED, nn[k], st, unpaired_count_kernel<<<tiles, SC_THREADS, 0, st>>>(w.s[k].rowflag, (u64)nn[k], w.s[k].uspine));
    BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.s[k].uspine, (u64)tiles, w.nU + k));
    BF_LAUNCH(PC_UNPAIRED, nn[k], st, unpaired_write_kernel<<<tiles, SC_THREADS, 0, st>>>(w.s[k].rowflag, grps[k], (u64)nn[k], w.s[k].uspine, w.s[k].ukeyA));
    ukeys_sorted[k] = w.s[k].ukeyA;
    if (L.rank_bits > 0) {
      const RadixPlan urp = make_radix_plan(32, 32 + L.rank_bits);
      BF_CUDA(cudaMemsetAsync(w.radix.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
      BF_TRY(launch_radix_hist(w.s[k].ukeyA, (u64)nn[k], w.nU + k, urp, w.radix.hist, st));
      SortResult ur;
      BF_TRY(radix_sort_passes<false>(w.s[k].ukeyA, w.s[k].ukeyB, nullptr, nullptr, false, (u64)nn[k], w.nU + k,
                                      urp, w.radix, st, &ur, PC_UNPAIRED));
      ukeys_sorted[k] = ur.keys;
    }
    BF_LAUNCH(PC_UNPAIRED, 0, st,
              unpaired_group_counts_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(ukeys_sorted[k], w.nU + k, nb,
                                                                                       w.s[k].ucnt));
  }
  plan->ukeys1 = ukeys_sorted[0];
  plan->ukeys2 = ukeys_sorted[1];

  // ---- per-group offsets and totals ----
  BF_LAUNCH(PC_TABLE, 0, st, offset_table_kernel<<<1, 1, 0, st>>>(w, n1, n2, nb, how));
  i64 header[8];
  BF_CUDA(cudaMemcpyAsync(header, w.header, sizeof(header), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  plan->n_rows = header[0];
  plan->n_pairs = header[1];
  plan->nU1 = header[2];
  plan->nU2 = header[3];
  plan->totA = header[4];
  plan->totB = header[5];
  if (header[6]) {
    set_error("bf_overlap_count: more than 2^32-1 pairs inside one block of %d consecutive sorted rows", SC_TILE);
    return BF_ERR_RANGE;
  }
  plan->magic = OVERLAP_MAGIC;
  plan->ready = 1;
  return BF_OK;
}

static int check_count_args(const char* who, i64 n1, i64 n2, i32 nb, i32 how, const OverlapPlan* plan) {
  if (n1 < 0 || n2 < 0 || nb < 1 || how < 0 || how > 3 || !plan) {
    set_error("%s: bad argument (n1=%lld n2=%lld n_groups=%d how=%d)", who, (long long)n1, (long long)n2, nb, how);
    return BF_ERR_ARG;
  }
  if (n1 >= (1ll << 31) || n2 >= (1ll << 31)) {
    set_error("%s: more than 2^31-1 rows in one set", who);
    return BF_ERR_RANGE;
  }
  return BF_OK;
}
static void init_plan(OverlapPlan* plan, i64 n1, i64 n2, i32 nb, i32 how, i32 closed, i32 prepared) {
  plan->magic = 0;
  plan->n1 = n1;
  plan->n2 = n2;
  plan->nb = nb;
  plan->how = how;
  plan->closed = closed;
  plan->ready = 0;
  plan->row_off1 = plan->row_off2 = 0;
  plan->prepared = prepared;
  plan->n_own1 = 0xffffffffu;
  plan->halo_ids = nullptr;
}

static int overlap_count_impl(const i32* grp1, const i64* s1, const i64* e1, i64 n1, const i32* grp2,
                              const i64* s2, const i64* e2, i64 n2, i32 nb, i32 how, i32 closed,
                              void* ws_ptr, size_t ws_bytes, OverlapPlan* plan, cudaStream_t st) {
  BF_TRY(check_count_args("bf_overlap_count", n1, n2, nb, how, plan));
  if ((n1 > 0 && (!s1 || !e1)) || (n2 > 0 && (!s2 || !e2)) || (nb > 1 && ((n1 > 0 && !grp1) || (n2 > 0 && !grp2)))) {
    set_error("bf_overlap_count: null input pointer");
    return BF_ERR_ARG;
  }
  Arena measure(nullptr);
  OverlapWS w;
  layout_overlap(measure, n1, n2, nb, how, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_overlap_count: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_overlap(arena, n1, n2, nb, how, &w);
  init_plan(plan, n1, n2, nb, how, closed, 0);

  // ---- coordinate extent, group code check, key layout (one small D2H) ----
  const RowSet sets[2] = {{grp1, s1, e1, n1}, {grp2, s2, e2, n2}};
  KeyLayout L;
  BF_TRY(measure_layout(sets, 2, nb, 1, 1, w.layout, st, &L, "bf_overlap_count"));
  // ---- sort both sets ----
  SortResult r1, r2;
  BF_TRY(sort_rows(nb > 1 ? grp1 : nullptr, s1, e1, n1, L, KEY_POINT_ADJUST | presorted_mode(L, 0), w.sb[0], w.radix, st, &r1));
  BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_POINT_ADJUST | presorted_mode(L, 1), w.sb[1], w.radix, st, &r2));
  return count_sorted(w, L, r1, n1, r2, n2, grp1, grp2, nb, how, closed, nullptr, plan, st);
}

// ---------------------------------------------------------------- prepared (pre-sorted) sets
// A set sorted once and used by several count/emit calls, or sorted on one rank and adopted on the
// others after a broadcast of its 8-byte keys and 4-byte row ids.
struct SetStore {
  SortBufs sb;
  RadixScratch radix;
  LayoutScratch layout;
  u64 *pmax, *pmax_spine;
};
static void layout_set_store(Arena& a, i64 n, i32 nb, bool adopt, SetStore* s) {
  if (adopt) {
    s->sb.keyA = a.take<u64>(n);
    s->sb.idA = a.take<u32>(n);
    s->sb.keyB = nullptr;
    s->sb.idB = nullptr;
    s->sb.tmp8 = nullptr;
    s->sb.tmp4a = s->sb.tmp4b = nullptr;
    s->sb.spine = s->sb.nT = nullptr;
    s->layout = take_layout_scratch(a, nb);
    s->pmax = a.take<u64>(n);
    s->pmax_spine = a.take<u64>(scan_tiles((u64)n) + 1);
    return;
  }
  take_sort_bufs(a, n, &s->sb);
  s->radix = take_radix_scratch(a, (u64)n);
  s->layout = take_layout_scratch(a, nb);
  s->pmax = a.take<u64>(n);
  s->pmax_spine = a.take<u64>(scan_tiles((u64)n) + 1);
}
struct SortedSet {  // lives in the caller's bf_sorted_set
  u64 magic;
  i64 n;
  i32 nb, adopted;
  KeyLayout L;  // goff points into the set's own store
  u64* keys;
  u32* ids;
  u64* pmax;        // [n] running max of end' in sorted order, valid once pmax_ready (bf_sorted_set_runmax)
  u64* pmax_spine;  // [scan tiles + 1]
  i32 pmax_ready, pad;
};
static_assert(sizeof(SortedSet) <= sizeof(bf_sorted_set), "bf_sorted_set too small");
constexpr u64 SET_MAGIC = 0x6266536f72746564ull;

// layout from the merged extent words, goff uploaded into the store
static int set_layout(const i64* words, i32 nb, const SetStore& store, cudaStream_t st, KeyLayout* L) {
  std::vector<u64> goff((size_t)nb + 2);
  BF_TRY(layout_from_words(words, nb, 1, L, goff.data()));
  BF_CUDA(cudaMemcpyAsync(store.layout.goff, goff.data(), ((size_t)nb + 1) * sizeof(u64), cudaMemcpyHostToDevice, st));
  BF_CUDA(cudaStreamSynchronize(st));  // the host vector goes out of scope
  L->goff = store.layout.goff;
  return BF_OK;
}

}  // namespace bf

extern "C" {

size_t bf_extent_words(int32_t n_groups) { return n_groups < 1 ? 0 : (size_t)bf::EXT_HDR + (size_t)n_groups; }
size_t bf_extent_workspace_bytes(int32_t n_groups) {
  if (n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::take_layout_scratch(a, n_groups);
  return a.off;
}
int bf_extent_measure(const int32_t* grp, const int64_t* start, const int64_t* end, int64_t n, int32_t n_groups,
                      void* ws, size_t ws_bytes, void* stream, int64_t* words_out, int32_t* presorted_out) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  if (n < 0 || n_groups < 1 || !words_out || (n > 0 && (!start || !end)) || (n > 0 && n_groups > 1 && !grp)) {
    set_error("bf_extent_measure: bad argument");
    return BF_ERR_ARG;
  }
  if (!ws || ws_bytes < bf_extent_workspace_bytes(n_groups)) {
    set_error("bf_extent_measure: workspace too small");
    return BF_ERR_WORKSPACE;
  }
  Arena a(ws);
  LayoutScratch sc = take_layout_scratch(a, n_groups);
  const i32 nb = n_groups;
  BF_LAUNCH(PC_EXTENT, 0, st, extent_init_kernel<<<(unsigned)ceil_div(nb + 1, 256), 256, 0, st>>>(sc.extent, sc.gmax, nb));
  const size_t smem = nb <= EXTENT_SMEM_GROUPS ? (size_t)nb * sizeof(i64) : 0;
  if (n > 0)
    BF_LAUNCH(PC_EXTENT, 20 * n, st,
              extent_kernel<<<stream_grid((u64)n, 1024), 256, smem, st>>>(nb > 1 ? (const i32*)grp : nullptr, (const i64*)start,
                                                                         (const i64*)end, (u64)n, 1, nb, 1, sc.extent, sc.gmax));
  Extent ext;
  BF_CUDA(cudaMemcpyAsync(&ext, sc.extent, sizeof(Extent), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  if (n > 1 && !ext.unsorted) {  // the sampled check found no disorder: exact check
    BF_LAUNCH(PC_EXTENT, 12 * n, st,
              order_check_kernel<<<stream_grid((u64)n, 2048), 256, 0, st>>>(nb > 1 ? (const i32*)grp : nullptr, (const i64*)start,
                                                                           (u64)n, 1, sc.extent));
    BF_CUDA(cudaMemcpyAsync(&ext, sc.extent, sizeof(Extent), cudaMemcpyDeviceToHost, st));
    BF_CUDA(cudaStreamSynchronize(st));
  }
  if (n > 0 && ext.bad) {
    set_error("interval with end < start: not a valid bedframe interval");
    return BF_ERR_ARG;
  }
  if (n > 0 && nb > 1 && (ext.gmin < 0 || ext.gmax >= nb)) {
    set_error("bf_extent_measure: group code outside [0,%d): min %d max %d", nb, ext.gmin, ext.gmax);
    return BF_ERR_GROUP;
  }
  std::vector<i64> gmax((size_t)nb);
  BF_CUDA(cudaMemcpyAsync(gmax.data(), sc.gmax, (size_t)nb * sizeof(i64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  extent_to_words(ext, gmax.data(), nb, n == 0, (i64*)words_out);
  if (presorted_out) *presorted_out = (n > 0 && !ext.unsorted) ? 1 : 0;
  return BF_OK;
}

size_t bf_sorted_set_bytes(int64_t n, int32_t n_groups, int32_t adopt) {
  if (n < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::SetStore s;
  bf::layout_set_store(a, n, n_groups, adopt != 0, &s);
  return a.off;
}

int bf_overlap_prepare_set(const int64_t* extent_words, const int32_t* grp, const int64_t* start, const int64_t* end,
                           int64_t n, int32_t n_groups, int32_t presorted, void* store, size_t store_bytes,
                           void* stream, bf_sorted_set* set_out) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  SortedSet* set = (SortedSet*)set_out;
  if (!extent_words || !set || n < 0 || n_groups < 1 || n >= (1ll << 31) || (n > 0 && (!start || !end)) ||
      (n > 0 && n_groups > 1 && !grp)) {
    set_error("bf_overlap_prepare_set: bad argument");
    return BF_ERR_ARG;
  }
  if (!store || store_bytes < bf_sorted_set_bytes(n, n_groups, 0)) {
    set_error("bf_overlap_prepare_set: set storage too small");
    return BF_ERR_WORKSPACE;
  }
  Arena a(store);
  SetStore ss;
  layout_set_store(a, n, n_groups, false, &ss);
  set->magic = 0;
  BF_TRY(set_layout((const i64*)extent_words, n_groups, ss, st, &set->L));
  SortResult r;
  BF_TRY(sort_rows(n_groups > 1 ? (const i32*)grp : nullptr, (const i64*)start, (const i64*)end, n, set->L,
                   KEY_POINT_ADJUST | (presorted ? KEY_PRESORTED : 0), ss.sb, ss.radix, st, &r));
  set->n = n;
  set->nb = n_groups;
  set->adopted = 0;
  set->keys = r.keys;
  set->ids = r.vals;
  set->pmax = ss.pmax;
  set->pmax_spine = ss.pmax_spine;
  set->pmax_ready = 0;
  set->magic = SET_MAGIC;
  return BF_OK;
}

int bf_overlap_adopt_set(const int64_t* extent_words, int64_t n, int32_t n_groups, void* store, size_t store_bytes,
                         void* stream, bf_sorted_set* set_out) {
  using namespace bf;
  SortedSet* set = (SortedSet*)set_out;
  if (!extent_words || !set || n < 0 || n_groups < 1 || n >= (1ll << 31)) {
    set_error("bf_overlap_adopt_set: bad argument");
    return BF_ERR_ARG;
  }
  if (!store || store_bytes < bf_sorted_set_bytes(n, n_groups, 1)) {
    set_error("bf_overlap_adopt_set: set storage too small");
    return BF_ERR_WORKSPACE;
  }
  Arena a(store);
  SetStore ss;
  layout_set_store(a, n, n_groups, true, &ss);
  set->magic = 0;
  BF_TRY(set_layout((const i64*)extent_words, n_groups, ss, (cudaStream_t)stream, &set->L));
  set->n = n;
  set->nb = n_groups;
  set->adopted = 1;
  set->keys = ss.sb.keyA;
  set->ids = ss.sb.idA;
  set->pmax = ss.pmax;
  set->pmax_spine = ss.pmax_spine;
  set->pmax_ready = 0;
  set->magic = SET_MAGIC;
  return BF_OK;
}

int bf_sorted_set_buffers(const bf_sorted_set* set_in, void** keys_out, void** ids_out, int64_t* n_out) {
  const bf::SortedSet* set = (const bf::SortedSet*)set_in;
  if (!set || set->magic != bf::SET_MAGIC) {
    bf::set_error("bf_sorted_set_buffers: not a prepared set");
    return BF_ERR_STATE;
  }
  if (keys_out) *keys_out = set->keys;
  if (ids_out) *ids_out = set->ids;
  if (n_out) *n_out = set->n;
  return BF_OK;
}

int bf_sorted_set_runmax(bf_sorted_set* set_io, void* stream) {
  using namespace bf;
  SortedSet* set = (SortedSet*)set_io;
  cudaStream_t st = (cudaStream_t)stream;
  if (!set || set->magic != SET_MAGIC) {
    set_error("bf_sorted_set_runmax: not a prepared set");
    return BF_ERR_STATE;
  }
  if (set->n > 0) {
    const unsigned tiles = (unsigned)scan_tiles((u64)set->n);
    BF_LAUNCH(PC_RUNMAX, 0, st, tile_max_ce_kernel<<<tiles, SC_THREADS, 0, st>>>(set->keys, (u64)set->n, set->L, set->pmax_spine));
    BF_LAUNCH(PC_RUNMAX, 0, st, spine_scan_kernel<true><<<1, 1024, 0, st>>>(set->pmax_spine, tiles, nullptr));
    BF_LAUNCH(PC_RUNMAX, 16 * set->n, st,
              running_max_ce_kernel<<<tiles, SC_THREADS, 0, st>>>(set->keys, (u64)set->n, set->L, set->pmax_spine, set->pmax));
  }
  set->pmax_ready = 1;
  return BF_OK;
}

size_t bf_overlap_prepared_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups, int32_t how) {
  if (n1 < 0 || n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::OverlapWS w;
  bf::layout_count(a, n1, n2, n_groups, how, &w);
  return a.off;
}

int bf_overlap_count_prepared(const bf_sorted_set* set1_in, const bf_sorted_set* set2_in, const int32_t* grp1,
                              const int32_t* grp2, int32_t how, int32_t closed, const int64_t* own_range, void* ws,
                              size_t ws_bytes, bf_plan* plan_out, void* stream, int64_t* n_rows_out,
                              int64_t* n_pairs_out) {
  using namespace bf;
  const SortedSet* a = (const SortedSet*)set1_in;
  const SortedSet* b = (const SortedSet*)set2_in;
  OverlapPlan* plan = (OverlapPlan*)plan_out;
  if (!a || !b || a->magic != SET_MAGIC || b->magic != SET_MAGIC) {
    set_error("bf_overlap_count_prepared: sets were not produced by bf_overlap_prepare_set / bf_overlap_adopt_set");
    return BF_ERR_STATE;
  }
  if (a->nb != b->nb || a->L.B != b->L.B || a->L.LB != b->L.LB || a->L.cmin != b->L.cmin || a->L.endcap != b->L.endcap) {
    set_error("bf_overlap_count_prepared: the two sets were prepared with different extents");
    return BF_ERR_STATE;
  }
  const i32 nb = a->nb;
  BF_TRY(check_count_args("bf_overlap_count_prepared", a->n, b->n, nb, how, plan));
  if (nb > 1 && (((how & 1) && a->n > 0 && !grp1) || ((how & 2) && b->n > 0 && !grp2))) {
    set_error("bf_overlap_count_prepared: the group codes of a kept side are needed for its unpaired rows");
    return BF_ERR_ARG;
  }
  if (own_range && (how & 2)) {
    set_error("bf_overlap_count_prepared: a query-sharded call supports how=inner/left (unpaired targets need a reduction)");
    return BF_ERR_ARG;
  }
  if (!ws || ws_bytes < bf_overlap_prepared_workspace_bytes(a->n, b->n, nb, how)) {
    set_error("bf_overlap_count_prepared: workspace too small");
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws);
  OverlapWS w;
  layout_count(arena, a->n, b->n, nb, how, &w);
  init_plan(plan, a->n, b->n, nb, how, closed, 1);
  SortResult r1 = {a->keys, a->ids, nullptr, nullptr}, r2 = {b->keys, b->ids, nullptr, nullptr};
  int rc = count_sorted(w, a->L, r1, a->n, r2, b->n, (const i32*)grp1, (const i32*)grp2, nb, how, closed,
                        (const i64*)own_range, plan, (cudaStream_t)stream, a->pmax_ready ? a->pmax : nullptr,
                        b->pmax_ready ? b->pmax : nullptr);
  if (rc == BF_OK) {
    if (n_rows_out) *n_rows_out = plan->n_rows;
    if (n_pairs_out) *n_pairs_out = plan->n_pairs;
  }
  return rc;
}

int bf_overlap_set_halo_rows(bf_plan* plan_in, int64_t n_own1, const int64_t* halo_ids) {
  bf::OverlapPlan* p = (bf::OverlapPlan*)plan_in;
  if (!p || p->magic != bf::OVERLAP_MAGIC || !p->ready) {
    bf::set_error("bf_overlap_set_halo_rows: plan was not produced by a successful count call");
    return BF_ERR_STATE;
  }
  if (n_own1 < 0 || n_own1 > p->n1 || (n_own1 < p->n1 && !halo_ids)) {
    bf::set_error("bf_overlap_set_halo_rows: bad argument");
    return BF_ERR_ARG;
  }
  p->n_own1 = (u32)n_own1;
  p->halo_ids = (const i64*)halo_ids;
  return BF_OK;
}

int bf_overlap_group_counts(const bf_plan* plan_in, void* ws, int64_t* counts_out, void* stream) {
  using namespace bf;
  const OverlapPlan* p = (const OverlapPlan*)plan_in;
  cudaStream_t st = (cudaStream_t)stream;
  if (!p || p->magic != OVERLAP_MAGIC || !p->ready || !counts_out) {
    set_error("bf_overlap_group_counts: plan was not produced by a successful count call");
    return BF_ERR_STATE;
  }
  Arena arena(ws);
  OverlapWS w;
  if (p->prepared)
    layout_count(arena, p->n1, p->n2, p->nb, p->how, &w);
  else
    layout_overlap(arena, p->n1, p->n2, p->nb, p->how, &w);
  const i32 nb = p->nb;
  std::vector<i64> cum((size_t)2 * (nb + 1));
  std::vector<u64> uc((size_t)2 * (nb + 1), 0ull);
  BF_CUDA(cudaMemcpyAsync(cum.data(), w.cumA, (size_t)(nb + 1) * sizeof(i64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(cum.data() + nb + 1, w.cumB, (size_t)(nb + 1) * sizeof(i64), cudaMemcpyDeviceToHost, st));
  if (p->how & 1) BF_CUDA(cudaMemcpyAsync(uc.data(), w.s[0].ucnt, (size_t)(nb + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
  if (p->how & 2) BF_CUDA(cudaMemcpyAsync(uc.data() + nb + 1, w.s[1].ucnt, (size_t)(nb + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  for (i32 c = 0; c < nb; ++c) {
    counts_out[4 * c + 0] = cum[c + 1] - cum[c];
    counts_out[4 * c + 1] = cum[nb + 1 + c + 1] - cum[nb + 1 + c];
    counts_out[4 * c + 2] = (i64)uc[c];
    counts_out[4 * c + 3] = (i64)uc[nb + 1 + c];
  }
  return BF_OK;
}

}  // extern "C"

namespace bf {

static int overlap_emit_impl(const OverlapPlan* plan, void* ws_ptr, i64* ev1, i64* ev2, cudaStream_t st) {
  if (!plan || plan->magic != OVERLAP_MAGIC || !plan->ready) {
    set_error("bf_overlap_emit: plan was not produced by a successful bf_overlap_count");
    return BF_ERR_STATE;
  }
  if (plan->n_rows > 0 && (!ev1 || !ev2)) {
    set_error("bf_overlap_emit: null output pointer");
    return BF_ERR_ARG;
  }
  Arena arena(ws_ptr);
  OverlapWS w;
  if (plan->prepared)
    layout_count(arena, plan->n1, plan->n2, plan->nb, plan->how, &w);
  else
    layout_overlap(arena, plan->n1, plan->n2, plan->nb, plan->how, &w);
  // pairs: chunks of EMIT_CHUNK_BLOCKS blocks, each preceded by its partition kernel
  for (int side = 0; side < 2; ++side) {
    const i64 total = side ? plan->totB : plan->totA;
    if (total <= 0) continue;
    const i64 n_owner = side ? plan->n2 : plan->n1;
    const SetWS& sw = w.s[side];
    const ScanView scan = {sw.tilepre, sw.loc, n_owner, side ? plan->shift2 : plan->shift1};
    i64* part = side ? w.partB : w.partA;
    i32* pgrp = side ? w.pgrpB : w.pgrpA;
    const i64* cum = side ? w.cumB : w.cumA;
    const i64* delta = side ? w.deltaB : w.deltaA;
    const i64 blocks = ceil_div(n_owner + total, EM_TILE);
    for (i64 b0 = 0; b0 < blocks; b0 += EMIT_CHUNK_BLOCKS) {
      const i64 nb_chunk = blocks - b0 < EMIT_CHUNK_BLOCKS ? blocks - b0 : EMIT_CHUNK_BLOCKS;
      BF_LAUNCH(PC_EMIT, 0, st,
                emit_partition_kernel<<<(unsigned)ceil_div(nb_chunk + 1, 256), 256, 0, st>>>(scan, n_owner, total, b0,
                                                                                            nb_chunk + 1, cum, plan->nb, part, pgrp));
      const u64 bytes = (u64)(20.0 * (double)total * (double)nb_chunk / (double)blocks);
      if (side == 0) {
        BF_LAUNCH(PC_EMIT, bytes, st,
                  emit_pairs_kernel<false><<<(unsigned)nb_chunk, EM_THREADS, 0, st>>>(
                      scan, sw.lo, n_owner, plan->id1, plan->id2, total, b0, part, pgrp, cum, delta, plan->nb,
                      plan->row_off1, plan->row_off2, 0xffffffffu, nullptr, plan->n2, plan->n_rows, ev1, ev2));
      } else {
        BF_LAUNCH(PC_EMIT, bytes, st,
                  emit_pairs_kernel<true><<<(unsigned)nb_chunk, EM_THREADS, 0, st>>>(
                      scan, sw.lo, n_owner, plan->id2, plan->id1, total, b0, part, pgrp, cum, delta, plan->nb,
                      plan->row_off1, plan->row_off2, plan->halo_ids ? plan->n_own1 : 0xffffffffu, plan->halo_ids,
                      plan->n1, plan->n_rows, ev1, ev2));
      }
    }
  }
  if (plan->nU1 > 0) {
    BF_LAUNCH(PC_UNPAIRED, 0, st, emit_unpaired_kernel<false><<<(unsigned)ceil_div(plan->nU1, 256), 256, 0, st>>>(plan->ukeys1, plan->nU1,
                                                                                    w.ustart1, w.ubase1, plan->row_off1, ev1, ev2));
  }
  if (plan->nU2 > 0) {
    BF_LAUNCH(PC_UNPAIRED, 0, st, emit_unpaired_kernel<true><<<(unsigned)ceil_div(plan->nU2, 256), 256, 0, st>>>(plan->ukeys2, plan->nU2,
                                                                                   w.ustart2, w.ubase2, plan->row_off2, ev1, ev2));
  }
  return BF_OK;
}

}  // namespace bf

extern "C" {

size_t bf_overlap_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups, int32_t how) {
  if (n1 < 0 || n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::OverlapWS w;
  bf::layout_overlap(a, n1, n2, n_groups, how, &w);
  return a.off;
}

int bf_overlap_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                     const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                     int32_t n_groups, int32_t how, int32_t closed, void* ws, size_t ws_bytes, bf_plan* plan,
                     void* stream, int64_t* n_rows_out, int64_t* n_pairs_out) {
  bf::OverlapPlan* p = (bf::OverlapPlan*)plan;
  int rc = bf::overlap_count_impl((const i32*)grp1, (const i64*)start1, (const i64*)end1, n1, (const i32*)grp2,
                                  (const i64*)start2, (const i64*)end2, n2, n_groups, how, closed, ws, ws_bytes, p,
                                  (cudaStream_t)stream);
  if (rc == BF_OK) {
    if (n_rows_out) *n_rows_out = p->n_rows;
    if (n_pairs_out) *n_pairs_out = p->n_pairs;
  }
  return rc;
}

int bf_overlap_set_row_offsets(bf_plan* plan, int64_t offset1, int64_t offset2) {
  bf::OverlapPlan* p = (bf::OverlapPlan*)plan;
  if (!p || p->magic != bf::OVERLAP_MAGIC || !p->ready) {
    bf::set_error("bf_overlap_set_row_offsets: plan was not produced by a successful bf_overlap_count");
    return BF_ERR_STATE;
  }
  p->row_off1 = offset1;
  p->row_off2 = offset2;
  return BF_OK;
}

int bf_overlap_emit(const bf_plan* plan, void* ws, int64_t* events1_out, int64_t* events2_out, void* stream) {
  return bf::overlap_emit_impl((const bf::OverlapPlan*)plan, ws, (i64*)events1_out, (i64*)events2_out,
                               (cudaStream_t)stream);
}

}  // extern "C"
