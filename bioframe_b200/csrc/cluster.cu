// bf_cluster_* / bf_coverage / bf_complement_*: merge_intervals
// (src/bioframe/core/arrops.py:415-479) for all groups in one pass, and the two
// operations the reference builds on it (complement_intervals arrops.py:482-503,
// coverage ops.py:842-916).
//
//   extent -> pack keys (raw end, no point adjustment) -> radix argsort + tie fix
//   -> ONE fused kernel over the sorted rows:
//        running max of end   (tile scan + max look-back;   np.maximum.accumulate, arrops.py:462)
//        cluster borders      (start > runmax_prev + min_dist, or >= when min_dist is None; arrops.py:467-470)
//        cluster ids          (tile scan + sum look-back;    np.cumsum, arrops.py:472)
//        scatter ids to input order, cluster start at borders, cluster end before
//        the next border (arrops.py:473-477), first row of each cluster (-> n_intervals, ops.py:653)
//   Groups are visited in group-rank order, ids run on across groups exactly as
//   `cluster_ids_group += max_cluster_id + 1` does (ops.py:654-656).
#include "common.cuh"
#include "keys.cuh"
#include "lookback.cuh"
#include "lut.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "sorted_set.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

// Cluster ids in input row order.  Two ways, chosen on the host once the cluster count is known:
//   few clusters (C < n / 16, e.g. deep coverage where everything merges): every row looks its cluster up by
//   its start -- the last cluster starting at or before it (cluster starts are strictly increasing when
//   min_dist is a number) -- with coalesced reads and writes and a small, cache-resident table;
//   otherwise: scatter of the per-sorted-row ids through the sort permutation (one random 8-byte write per row,
//   the slowest access pattern of the whole library: 4.3 ms per 1e8 rows).
constexpr i64 CLUSTER_GATHER_RATIO = 16;

struct ClusterWS {
  SortBufs sb;
  RadixScratch radix;
  LayoutScratch layout;
  u64 *spine_max, *spine_cnt;  // [tiles + 1]
  u64* cend_true;  // [n+1] clamp mode only: true cluster ends (sign-flipped image, 0 = none)
  u32* cid_sorted; // [n] cluster id of every sorted row (input of the two id kernels below)
  u64* cs_lin;     // [n/16 + 2] linear starts of the clusters (gather path of the ids, few clusters only)
  u32* id_lut;     // direct-address table over cs_lin
  i64* header;  // [0] = number of clusters
};
static void layout_cluster(Arena& a, i64 n, i32 nb, ClusterWS* w) {
  take_sort_bufs(a, n, &w->sb);
  w->radix = take_radix_scratch(a, (u64)n);
  w->layout = take_layout_scratch(a, nb);
  w->spine_max = a.take<u64>(scan_tiles((u64)n) + 2);
  w->spine_cnt = a.take<u64>(scan_tiles((u64)n) + 2);
  w->cend_true = a.take<u64>(n + 1);
  w->cid_sorted = a.take<u32>(n + 1);
  w->cs_lin = a.take<u64>(n / CLUSTER_GATHER_RATIO + 2);
  w->id_lut = a.take<u32>(lut_words(n / CLUSTER_GATHER_RATIO + 1));
  w->header = a.take<i64>(4);
}

// cluster table in the workspace (n entries of room; header[0] are valid)
struct ClusterTable {
  i64* cstart;  // raw coordinates
  i64* cend;
  u32* cfirst;  // sorted position of the cluster's first row
  u32* cgrp;    // group rank
  i64* n_clusters;  // device scalar
};

struct ClusterPlan {
  u64 magic;
  i64 n;
  i32 nb, ready;
  i64 n_clusters;
  ClusterTable tab;
};
static_assert(sizeof(ClusterPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 CLUSTER_MAGIC = 0x6266436c75737472ull;

// group of the next row given the group of the previous one (rows are sorted, groups only grow)
__device__ __forceinline__ u32 advance_rank(u32 rank, u64 cs, const KeyLayout& L) {
  while ((i32)rank + 1 < L.nb && L.goff[rank + 1] <= cs) ++rank;
  return rank;
}
// border test between two consecutive sorted rows of the SAME group: arrops.py:468 / 470
__device__ __forceinline__ bool gap_opens_cluster(u64 cs, u64 runmax_prev, u64 min_dist, int has_min_dist) {
  return has_min_dist ? (cs > runmax_prev + min_dist) : (cs >= runmax_prev);
}

// Three short passes instead of one pass with two chained look-backs (a chain of ~5e4 tiles costs more
// than re-reading 8 B/row twice, see profiles/NOTES.md):
//   cluster_tile_max_kernel  -> per-tile max of linear end            -> spine scan (exclusive max)
//   cluster_pass_kernel<2>   -> running max, borders, per-tile count  -> spine scan (exclusive sum)
//   cluster_pass_kernel<3>   -> the same again + ids, cluster table
__global__ void __launch_bounds__(SC_THREADS)
    cluster_tile_max_kernel(const u64* __restrict__ keys, u64 n, KeyLayout L, u64* __restrict__ spine_max) {
  __shared__ u64 s_tmp[8];
  const u64 base = (u64)blockIdx.x * SC_TILE;
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u64 i = base + (u64)j * SC_THREADS + threadIdx.x;
    if (i < n) {
      const u64 v = key_ce(keys[i], L);
      acc = v > acc ? v : acc;
    }
  }
  acc = block_max_u64(acc, s_tmp);
  if (threadIdx.x == 0) spine_max[blockIdx.x] = acc;
}

template <int PHASE>
__global__ void __launch_bounds__(SC_THREADS, 6)
    cluster_pass_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids, u64 n, KeyLayout L, u64 min_dist,
                        int has_min_dist, const u64* __restrict__ spine_max, u64* __restrict__ spine_cnt,
                        u32* __restrict__ cid_sorted, ClusterTable tab, const i64* __restrict__ e_raw,
                        u64* __restrict__ cend_true) {
  // [0] row before the tile, [1..rows] tile, [rows+1] row after; one pad slot every 8 entries so that
  // threads reading 8 consecutive rows each do not collide on the shared-memory banks
  __shared__ u64 s_kp[SC_TILE + 2 + (SC_TILE + 2) / 8 + 1];
#define s_k(i) s_kp[(i) + ((i) >> 3)]
  __shared__ u64 s_tmp[8];
  const u32 tid = threadIdx.x;
  const u64 tile = blockIdx.x;
  const u64 base = tile * SC_TILE;
  const u64 rows = n - base < (u64)SC_TILE ? n - base : (u64)SC_TILE;
#pragma unroll 9
  for (u32 i = tid; i < rows + 2; i += SC_THREADS) {
    const i64 p = (i64)base + i - 1;
    s_k(i) = (p >= 0 && (u64)p < n) ? keys[p] : 0ull;
  }
  __syncthreads();

  // ---- running max of linear (group, end) over the sorted rows ----
  const u32 r0 = tid * SC_IPT;
  u64 rm[SC_IPT];  // inclusive running max inside the thread
  u64 tmax = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u32 r = r0 + j;
    const u64 v = r < rows ? key_ce(s_k(r + 1), L) : 0ull;
    tmax = v > tmax ? v : tmax;
    rm[j] = tmax;
  }
  u64 before = block_excl_max_u64(tmax, s_tmp);
  const u64 tile_pre = spine_max[tile];
  before = before > tile_pre ? before : tile_pre;  // running max over every earlier row

  // ---- group of each row (binary search once per thread, then advance), borders ----
  u32 rk[SC_IPT + 1];  // [0] row before the thread's first, [1..IPT] own rows
  rk[0] = r0 < rows ? rank_of_cs(key_cs(s_k(r0), L), L) : 0u;
  u32 bmask = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u32 r = r0 + j;
    rk[j + 1] = rk[j];
    if (r < rows) {
      const u64 cs = key_cs(s_k(r + 1), L);
      rk[j + 1] = advance_rank(rk[j], cs, L);
      const u64 prev_rm = j == 0 ? before : (rm[j - 1] > before ? rm[j - 1] : before);
      const bool first_row = base + r == 0;
      if (first_row || rk[j + 1] != rk[j] || gap_opens_cluster(cs, prev_rm, min_dist, has_min_dist)) bmask |= 1u << j;
    }
  }
  u64 tot;
  u64 nborders_before = block_excl_sum_u64((u64)__popc(bmask), s_tmp, &tot);
  if (PHASE == 2) {
    if (tid == 0) spine_cnt[tile] = tot;
    return;
  }
  nborders_before += spine_cnt[tile];

  // ---- ids, cluster table ----
  u64 seen = nborders_before;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u32 r = r0 + j;
    if (r < rows) {
      const u64 k = base + r;
      const u64 key = s_k(r + 1);
      if (bmask & (1u << j)) ++seen;
      const u64 id = seen - 1;
      BF_DEVICE_CHECK(seen >= 1 && id <= k);  // a row's cluster ordinal never exceeds its sorted position
      const u32 rank = rk[j + 1];
      const u64 gbase = L.goff[rank];
      const u64 run = rm[j] > before ? rm[j] : before;  // running max including this row
      if (cid_sorted) cid_sorted[k] = (u32)id;
      if (bmask & (1u << j)) {
        tab.cstart[id] = (i64)(key_cs(key, L) - gbase + L.cmin);
        tab.cfirst[id] = (u32)k;
        tab.cgrp[id] = rank;
      }
      bool closes = k + 1 == n;
      if (!closes) {
        const u64 ncs = key_cs(s_k(r + 2), L);
        closes = advance_rank(rank, ncs, L) != rank || gap_opens_cluster(ncs, run, min_dist, has_min_dist);
      }
      if (closes) tab.cend[id] = (i64)(run - gbase + L.cmin);
      if (e_raw) {  // clamp mode: the packed end of this row may be cut off; remember the true one
        const i64 te = e_raw[ids[k]];
        if (te > L.endcap) atomicMax(&cend_true[id], (u64)te ^ 0x8000000000000000ull);
      }
    }
  }
#undef s_k
}

__global__ void __launch_bounds__(256)
    cluster_ids_scatter_kernel(const u32* __restrict__ cid_sorted, const u32* __restrict__ ids, i64 n,
                               i64* __restrict__ ids_out) {
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (i64)gridDim.x * blockDim.x)
    ids_out[ids[k]] = (i64)cid_sorted[k];
}
__global__ void __launch_bounds__(256)
    cluster_starts_lin_kernel(ClusterTable tab, i64 C, KeyLayout L, u64* __restrict__ cs_lin) {
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) cs_lin[c] = L.goff[tab.cgrp[c]] + ((u64)tab.cstart[c] - L.cmin);
}
__global__ void __launch_bounds__(256)
    cluster_ids_gather_kernel(const i32* __restrict__ grp, const i64* __restrict__ s, i64 n, KeyLayout L, Lut lut,
                              const u64* __restrict__ cs_lin, i64* __restrict__ ids_out) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const u64 x = (grp ? L.goff[grp[r]] : 0ull) + ((u64)s[r] - L.cmin);
  ids_out[r] = lut_count_below<true>(cs_lin, lut, x) - 1;  // last cluster starting at or before the row
}

// clamp mode: cluster ends that were cut off by the clamp get their true value back
__global__ void __launch_bounds__(256) cluster_fix_ends_kernel(ClusterTable tab, const u64* __restrict__ cend_true) {
  const i64 C = *tab.n_clusters;
  for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < C; c += (i64)gridDim.x * blockDim.x) {
    const u64 t = cend_true[c];
    if (t) {
      const i64 te = (i64)(t ^ 0x8000000000000000ull);
      if (te > tab.cend[c]) tab.cend[c] = te;
    }
  }
}

__global__ void __launch_bounds__(256)
    cluster_emit_kernel(ClusterTable tab, i64 n_clusters, i64 n, i32* __restrict__ cgrp_out,
                        i64* __restrict__ cstart_out, i64* __restrict__ cend_out, i64* __restrict__ ccount_out) {
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clusters) return;
  if (cgrp_out) cgrp_out[c] = (i32)tab.cgrp[c];
  if (cstart_out) cstart_out[c] = tab.cstart[c];
  if (cend_out) cend_out[c] = tab.cend[c];
  if (ccount_out) ccount_out[c] = (c + 1 < n_clusters ? (i64)tab.cfirst[c + 1] : n) - (i64)tab.cfirst[c];
}
__global__ void __launch_bounds__(256)
    cluster_rows_kernel(const i64* __restrict__ ids, i64 n, ClusterTable tab, i64* __restrict__ row_start,
                        i64* __restrict__ row_end) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const i64 c = ids[r];
  if (row_start) row_start[r] = tab.cstart[c];
  if (row_end) row_end[r] = tab.cend[c];
}

// Sort + fused scan; leaves the cluster table in the workspace. No host sync.
static int cluster_run(const i32* grp, const i64* s, const i64* e, i64 n, i32 nb, i64 min_dist, int has_min_dist,
                       ClusterWS& w, bool want_ids, cudaStream_t st, ClusterTable* tab, KeyLayout* L_out,
                       int allow_clamp = 1, SortResult* sorted_out = nullptr) {
  const RowSet sets[1] = {{grp, s, e, n}};
  KeyLayout L;
  // ends beyond (largest start + 1) may be clamped (INT64_MAX ends, e.g. the last gap of a complement):
  // the borders do not change (such an end reaches past every start either way) and the true cluster
  // ends are restored afterwards (cluster_fix_ends_kernel)
  BF_TRY(measure_layout(sets, 1, nb, 0, allow_clamp, w.layout, st, &L, "bf_cluster"));
  const bool clamped = L.endcap != I64_MAX;
  if (clamped) BF_CUDA(cudaMemsetAsync(w.cend_true, 0, (size_t)(n + 1) * sizeof(u64), st));
  if (L.B > 62) {
    set_error("coordinate span too wide for the running-max scan (%d bits)", L.B);
    return BF_ERR_RANGE;
  }
  SortResult r;
  BF_TRY(sort_rows(nb > 1 ? grp : nullptr, s, e, n, L, presorted_mode(L, 0), w.sb, w.radix, st, &r));
  // table scratch: buffers that are free once the sort is done
  tab->cstart = (i64*)r.keys_alt;
  tab->cend = (i64*)w.sb.tmp8;
  tab->cfirst = w.sb.tmp4a;
  tab->cgrp = w.sb.tmp4b;
  tab->n_clusters = w.header;
  const unsigned tiles = (unsigned)scan_tiles((u64)n);
  u64 md = 0;
  if (has_min_dist) md = min_dist < 0 ? 0ull : (u64)min_dist;
  if (md > (1ull << 62)) md = 1ull << 62;
  BF_LAUNCH(PC_CLUSTER, 8 * n, st, cluster_tile_max_kernel<<<tiles, SC_THREADS, 0, st>>>(r.keys, (u64)n, L, w.spine_max));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<true><<<1, 1024, 0, st>>>(w.spine_max, (u64)tiles, nullptr));
  BF_LAUNCH(PC_CLUSTER, 8 * n, st,
            cluster_pass_kernel<2><<<tiles, SC_THREADS, 0, st>>>(r.keys, r.vals, (u64)n, L, md, has_min_dist, w.spine_max,
                                                                 w.spine_cnt, nullptr, *tab, nullptr, nullptr));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.spine_cnt, (u64)tiles, (u64*)w.header));
  BF_LAUNCH(PC_CLUSTER, (u64)(12 + (want_ids ? 4 : 0)) * (u64)n, st,
            cluster_pass_kernel<3><<<tiles, SC_THREADS, 0, st>>>(r.keys, r.vals, (u64)n, L, md, has_min_dist, w.spine_max,
                                                                 w.spine_cnt, want_ids ? w.cid_sorted : nullptr, *tab,
                                                                 clamped ? e : nullptr, w.cend_true));
  if (clamped) {
    BF_LAUNCH(PC_CLUSTER, 0, st, cluster_fix_ends_kernel<<<stream_grid((u64)n, 1024), 256, 0, st>>>(*tab, w.cend_true));
  }
  if (L_out) *L_out = L;
  if (sorted_out) *sorted_out = r;
  return BF_OK;
}

static int cluster_count_impl(const i32* grp, const i64* s, const i64* e, i64 n, i32 nb, i64 min_dist,
                              int has_min_dist, void* ws_ptr, size_t ws_bytes, ClusterPlan* plan, cudaStream_t st,
                              i64* ids_out) {
  if (n < 0 || nb < 1 || !plan) {
    set_error("bf_cluster_count: bad argument (n=%lld n_groups=%d)", (long long)n, nb);
    return BF_ERR_ARG;
  }
  if (n >= (1ll << 31)) {
    set_error("bf_cluster_count: more than 2^31-1 rows");
    return BF_ERR_RANGE;
  }
  if (n > 0 && (!s || !e || (nb > 1 && !grp))) {
    set_error("bf_cluster_count: null input pointer");
    return BF_ERR_ARG;
  }
  Arena measure(nullptr);
  ClusterWS w;
  layout_cluster(measure, n, nb, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_cluster_count: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_cluster(arena, n, nb, &w);
  plan->magic = 0;
  plan->n = n;
  plan->nb = nb;
  plan->ready = 0;
  plan->n_clusters = 0;
  if (n > 0) {
    KeyLayout L;
    SortResult sorted;
    BF_TRY(cluster_run(grp, s, e, n, nb, min_dist, has_min_dist, w, ids_out != nullptr, st, &plan->tab, &L, 1, &sorted));
    i64 c = 0;
    BF_CUDA(cudaMemcpyAsync(&c, w.header, sizeof(i64), cudaMemcpyDeviceToHost, st));
    BF_CUDA(cudaStreamSynchronize(st));
    plan->n_clusters = c;
    if (ids_out) {
      if (has_min_dist && c * CLUSTER_GATHER_RATIO < n) {
        const Lut lut = make_lut(w.id_lut, c, 0, 1ull << L.B, 0);
        BF_LAUNCH(PC_CLUSTER, 0, st, cluster_starts_lin_kernel<<<(unsigned)ceil_div(c, 256), 256, 0, st>>>(plan->tab, c, L, w.cs_lin));
        BF_TRY(launch_lut_build(w.cs_lin, c, nullptr, lut, w.id_lut, st, PC_CLUSTER));
        BF_LAUNCH(PC_CLUSTER, 20 * n, st,
                  cluster_ids_gather_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(nb > 1 ? grp : nullptr, s, n, L, lut,
                                                                                       w.cs_lin, ids_out));
      } else {
        BF_LAUNCH(PC_CLUSTER, 16 * n, st,
                  cluster_ids_scatter_kernel<<<stream_grid((u64)n, 1024), 256, 0, st>>>(w.cid_sorted, sorted.vals, n, ids_out));
      }
    }
  }
  plan->magic = CLUSTER_MAGIC;
  plan->ready = 1;
  return BF_OK;
}

static int cluster_emit_impl(const ClusterPlan* plan, i32* cgrp_out, i64* cstart_out, i64* cend_out, i64* ccount_out,
                             const i64* ids, i64* row_start, i64* row_end, cudaStream_t st) {
  if (!plan || plan->magic != CLUSTER_MAGIC || !plan->ready) {
    set_error("bf_cluster_emit: plan was not produced by a successful bf_cluster_count");
    return BF_ERR_STATE;
  }
  if (plan->n_clusters > 0) {
    BF_LAUNCH(PC_CLUSTER, 28 * plan->n_clusters, st,
              cluster_emit_kernel<<<(unsigned)ceil_div(plan->n_clusters, 256), 256, 0, st>>>(
                  plan->tab, plan->n_clusters, plan->n, cgrp_out, cstart_out, cend_out, ccount_out));
  }
  if (plan->n > 0 && ids && (row_start || row_end)) {
    BF_LAUNCH(PC_CLUSTER, 24 * plan->n, st,
              cluster_rows_kernel<<<(unsigned)ceil_div(plan->n, 256), 256, 0, st>>>(ids, plan->n, plan->tab, row_start,
                                                                                   row_end));
  }
  return BF_OK;
}

// ------------------------------------------------------------------ coverage
// coverage[i] = sum over the MERGED targets of max(0, min(e1, E) - max(s1, S))
// (ops.py:888-912: merge(df2), left overlap, clip, per-query sum).  The merged
// targets are disjoint and sorted, so with PL = exclusive prefix sum of their
// lengths the covered part of (-inf, x) in a group is
//     F(x) = PL[j-1] + min(x, E[j-1]) - S[j-1],   j = #{clusters of the group with S < x}
// and coverage = F(end) - F(start): two binary searches per query, no pairs.
struct CoverageWS {
  ClusterWS cl;
  ulonglong4* rec;   // [n2+1] merged intervals as 32-byte records: x = linear start, y = linear end,
                     // z = exclusive prefix of merged lengths (one sector answers a whole F(x) evaluation)
  u64* spine;        // [tiles + 1]
  u32* lut;          // direct-address table over slin
};
static void layout_coverage(Arena& a, i64 n2, i32 nb, CoverageWS* w) {
  layout_cluster(a, n2, nb, &w->cl);
  w->rec = a.take<ulonglong4>(n2 + 1);
  w->spine = a.take<u64>(scan_tiles((u64)n2) + 2);
  w->lut = a.take<u32>(lut_words(n2));
}

// per-tile sums of the merged lengths (n lives on the device)
__global__ void __launch_bounds__(SC_THREADS) merged_length_tiles_kernel(ClusterTable tab, i64 endcap,
                                                                          u64* __restrict__ spine) {
  __shared__ u64 s_tmp[8];
  const u64 n = (u64)*tab.n_clusters;
  const u64 base = (u64)blockIdx.x * SC_TILE;
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u64 c = base + (u64)j * SC_THREADS + threadIdx.x;
    if (c < n) acc += (u64)(min(tab.cend[c], endcap) - tab.cstart[c]);
  }
  acc = block_sum_u64(acc, s_tmp);
  if (threadIdx.x == 0) spine[blockIdx.x] = acc;
}
// linear start / end and exclusive length prefix of every merged interval
__global__ void __launch_bounds__(SC_THREADS)
    merged_table_kernel(ClusterTable tab, KeyLayout L, const u64* __restrict__ spine, ulonglong4* __restrict__ rec) {
  __shared__ u64 s_tmp[8];
  const u64 n = (u64)*tab.n_clusters;
  const u64 base = (u64)blockIdx.x * SC_TILE + (u64)threadIdx.x * SC_IPT;
  if ((u64)blockIdx.x * SC_TILE > n) return;  // block-uniform
  u64 len[SC_IPT];
  u64 mine = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u64 c = base + j;
    len[j] = c < n ? (u64)(min(tab.cend[c], L.endcap) - tab.cstart[c]) : 0ull;
    mine += len[j];
  }
  u64 tot;
  u64 run = spine[blockIdx.x] + block_excl_sum_u64(mine, s_tmp, &tot);
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const u64 c = base + j;
    if (c < n) {
      const u64 off = L.goff[tab.cgrp[c]];
      rec[c] = make_ulonglong4(off + ((u64)tab.cstart[c] - L.cmin), off + ((u64)min(tab.cend[c], L.endcap) - L.cmin), run,
                               0ull);
    }
    run += len[j];
  }
}

// Covered length below a linear coordinate x:  F(x) = plen[j-1] + min(x, E[j-1]) - S[j-1],  j = #{S < x}
// (0 when j == 0).  On the linear axis no group test is needed: merged intervals of earlier groups end
// before the group's offset, so they count in full in both F(end) and F(start) and cancel.
__global__ void __launch_bounds__(256)
    coverage_kernel(const i32* __restrict__ grp1, const i64* __restrict__ s1, const i64* __restrict__ e1, i64 n1,
                    const i64* __restrict__ n_clusters, KeyLayout L, Lut lut, const ulonglong4* __restrict__ rec,
                    i64* __restrict__ cov) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n1) return;
  const i64 C = *n_clusters;
  const u32 g = grp1 ? (u32)grp1[r] : 0u;
  if (g >= (u32)L.nb) {  // group code outside the call's groups: nothing can cover it
    cov[r] = 0;
    return;
  }
  // query coordinates clamped into the group's stretch of the axis (the layout was measured on the
  // targets only; nothing of a group lies outside its stretch, so clamping cannot change the coverage)
  const u64 g0 = L.goff[g], span = L.goff[g + 1] - g0;
  auto lin = [&](i64 x) -> u64 {
    if ((u64)x - L.cmin > (1ull << 63)) return g0;  // x below the origin (wrapped difference)
    const u64 d = (u64)x - L.cmin;
    return g0 + (d < span ? d : span);
  };
  const u64* slin = reinterpret_cast<const u64*>(rec);  // entry m at word 4 * m (lut.stride == 4)
  auto below = [&](i64 j, u64 x) -> u64 {
    if (j == 0) return 0ull;
    const ulonglong4 r1 = rec[j - 1];
    return r1.z + ((x < r1.y ? x : r1.y) - r1.x);
  };
  const u64 xs = lin(s1[r]), xe = lin(e1[r]);
  const i64 js = lut_count_below<false>(slin, lut, xs);  // = #{S < xs}: one table sector + a search inside the bucket
  // the end is rarely more than a few merged intervals further on: gallop from js
  i64 lo = js, hi = js, step = 1;
  for (;;) {
    hi = lo + step;
    if (hi >= C) {
      hi = C;
      break;
    }
    if (!(slin[4 * (hi - 1)] < xe)) break;
    lo = hi;
    step <<= 1;
  }
  const i64 je = lower_bound_by(lo, hi, [&](i64 m) { return slin[4 * m] < xe; });
  const u64 fe = below(je, xe), fs = below(js, xs);
  cov[r] = fe > fs ? (i64)(fe - fs) : 0;
}

static int coverage_impl(const i32* grp1, const i64* s1, const i64* e1, i64 n1, const i32* grp2, const i64* s2,
                         const i64* e2, i64 n2, i32 nb, void* ws_ptr, size_t ws_bytes, cudaStream_t st, i64* cov) {
  if (n1 < 0 || n2 < 0 || nb < 1) {
    set_error("bf_coverage: bad argument");
    return BF_ERR_ARG;
  }
  if (n1 >= (1ll << 31) || n2 >= (1ll << 31)) {
    set_error("bf_coverage: more than 2^31-1 rows in one set");
    return BF_ERR_RANGE;
  }
  if (n1 == 0) return BF_OK;
  if (!cov || !s1 || !e1 || (nb > 1 && !grp1)) {
    set_error("bf_coverage: null pointer");
    return BF_ERR_ARG;
  }
  if (n2 == 0) {
    BF_CUDA(cudaMemsetAsync(cov, 0, (size_t)n1 * sizeof(i64), st));
    return BF_OK;
  }
  Arena measure(nullptr);
  CoverageWS w;
  layout_coverage(measure, n2, nb, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_coverage: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_coverage(arena, n2, nb, &w);
  ClusterTable tab;
  KeyLayout L;
  // no clamping here: coverage does arithmetic on the ends themselves, so targets whose ends do not fit
  // the key (INT64_MAX) are refused (BF_ERR_RANGE) rather than answered approximately
  BF_TRY(cluster_run(grp2, s2, e2, n2, nb, 0, 1, w.cl, false, st, &tab, &L, 0));
  const unsigned tiles = (unsigned)scan_tiles((u64)n2);
  BF_LAUNCH(PC_COVERAGE, 16 * n2, st, merged_length_tiles_kernel<<<tiles, SC_THREADS, 0, st>>>(tab, L.endcap, w.spine));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.spine, (u64)tiles, nullptr));
  BF_LAUNCH(PC_COVERAGE, 44 * n2, st, merged_table_kernel<<<tiles, SC_THREADS, 0, st>>>(tab, L, w.spine, w.rec));
  const Lut lut = make_lut(w.lut, n2, 0, 1ull << L.B, 0, 4);
  BF_TRY(launch_lut_build(reinterpret_cast<const u64*>(w.rec), 0, tab.n_clusters, lut, w.lut, st, PC_COVERAGE));
  BF_LAUNCH(PC_COVERAGE, 28 * n1, st,
            coverage_kernel<<<(unsigned)ceil_div(n1, 256), 256, 0, st>>>(nb > 1 ? grp1 : nullptr, s1, e1, n1, tab.n_clusters, L,
                                                                        lut, w.rec, cov));
  return BF_OK;
}

// ---------------------------------------------------------------- complement
// Per view region r (group rank = position of the region in sorted name order,
// ops.py:1645-1648) with bounds (b0, b1):  merged = merge(min_dist=0);
// lo = #{merged end <= b0}, hi = #{merged start < b1}  (arrops.py:489-490);
// gaps = (b0, S[lo]), (E[lo], S[lo+1]), ..., (E[hi-1], b1) minus a leading /
// trailing empty gap (arrops.py:496-501).
struct ComplementWS {
  ClusterWS cl;
  i64 *r_lo, *r_hi, *r_off;  // [n_regions + 1]
  i32* r_first;              // [n_regions] 1 if the leading gap is kept
  i64* total;                // [1]
};
static void layout_complement(Arena& a, i64 n, i32 nr, ComplementWS* w) {
  layout_cluster(a, n, nr, &w->cl);
  w->r_lo = a.take<i64>(nr + 1);
  w->r_hi = a.take<i64>(nr + 1);
  w->r_off = a.take<i64>(nr + 1);
  w->r_first = a.take<i32>(nr + 1);
  w->total = a.take<i64>(2);
}
struct ComplementPlan {
  u64 magic;
  i64 n;
  i32 nr, ready;
  i64 n_gaps, n_clusters;
  ClusterTable tab;
  const i64 *b0, *b1;
};
static_assert(sizeof(ComplementPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 COMPLEMENT_MAGIC = 0x6266436f6d706c6dull;

__global__ void __launch_bounds__(256)
    complement_regions_kernel(ClusterTable tab, i64 have_table, const i64* __restrict__ b0s, const i64* __restrict__ b1s,
                              i32 nr, i64* __restrict__ r_lo, i64* __restrict__ r_hi, i64* __restrict__ r_cnt,
                              i32* __restrict__ r_first) {
  const i32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr) return;
  const i64 C = have_table ? *tab.n_clusters : 0;
  const i64 b0 = b0s[r], b1 = b1s[r];
  const i64 c0 = lower_bound_by(0, C, [&](i64 m) { return tab.cgrp[m] < (u32)r; });
  const i64 c1 = lower_bound_by(c0, C, [&](i64 m) { return tab.cgrp[m] <= (u32)r; });
  const i64 lo = lower_bound_by(c0, c1, [&](i64 m) { return tab.cend[m] <= b0; });
  i64 hi = lower_bound_by(c0, c1, [&](i64 m) { return tab.cstart[m] < b1; });
  if (hi < lo) hi = lo;
  const i64 m = hi - lo;
  i64 cnt;
  i32 first;
  if (m == 0) {
    // a region without any interval yields the whole region (ops.py:1652-1657);
    // one whose merged intervals all fall outside the bounds yields it only if non-empty
    first = (c0 == c1) ? 1 : (b0 < b1);
    cnt = first;
  } else {
    first = b0 < tab.cstart[lo];
    const i32 last = tab.cend[hi - 1] < b1;
    cnt = (m - 1) + first + last;
  }
  r_lo[r] = lo;
  r_hi[r] = hi;
  r_cnt[r] = cnt;
  r_first[r] = first;
}
__global__ void complement_offsets_kernel(i64* __restrict__ r_cnt_to_off, i32 nr, i64* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  i64 run = 0;
  for (i32 r = 0; r < nr; ++r) {
    const i64 c = r_cnt_to_off[r];
    r_cnt_to_off[r] = run;
    run += c;
  }
  r_cnt_to_off[nr] = run;
  *total = run;
}
__global__ void __launch_bounds__(256)
    complement_edges_kernel(ClusterTable tab, const i64* __restrict__ b0s, const i64* __restrict__ b1s, i32 nr,
                            const i64* __restrict__ r_lo, const i64* __restrict__ r_hi, const i64* __restrict__ r_off,
                            const i32* __restrict__ r_first, i32* __restrict__ reg_out, i64* __restrict__ s_out,
                            i64* __restrict__ e_out) {
  const i32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr) return;
  const i64 lo = r_lo[r], hi = r_hi[r], off = r_off[r], cnt = r_off[r + 1] - off;
  const i64 m = hi - lo;
  if (cnt == 0) return;
  if (m == 0) {
    reg_out[off] = r;
    s_out[off] = b0s[r];
    e_out[off] = b1s[r];
    return;
  }
  if (r_first[r]) {
    reg_out[off] = r;
    s_out[off] = b0s[r];
    e_out[off] = tab.cstart[lo];
  }
  if (tab.cend[hi - 1] < b1s[r]) {
    const i64 at = off + cnt - 1;
    reg_out[at] = r;
    s_out[at] = tab.cend[hi - 1];
    e_out[at] = b1s[r];
  }
}
__global__ void __launch_bounds__(256)
    complement_inner_kernel(ClusterTable tab, i64 n_clusters, const i64* __restrict__ r_lo,
                            const i64* __restrict__ r_hi, const i64* __restrict__ r_off,
                            const i32* __restrict__ r_first, i32* __restrict__ reg_out, i64* __restrict__ s_out,
                            i64* __restrict__ e_out) {
  const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_clusters) return;
  const u32 r = tab.cgrp[j];
  if (j < r_lo[r] || j + 1 >= r_hi[r]) return;  // gap between cluster j and j+1 of the same region
  const i64 at = r_off[r] + r_first[r] + (j - r_lo[r]);
  reg_out[at] = (i32)r;
  s_out[at] = tab.cend[j];
  e_out[at] = tab.cstart[j + 1];
}

static int complement_count_impl(const i32* region, const i64* s, const i64* e, i64 n, const i64* b0, const i64* b1,
                                 i32 nr, void* ws_ptr, size_t ws_bytes, ComplementPlan* plan, cudaStream_t st) {
  if (n < 0 || nr < 1 || !plan || !b0 || !b1) {
    set_error("bf_complement_count: bad argument (n=%lld n_regions=%d)", (long long)n, nr);
    return BF_ERR_ARG;
  }
  if (n >= (1ll << 31)) {
    set_error("bf_complement_count: more than 2^31-1 rows");
    return BF_ERR_RANGE;
  }
  Arena measure(nullptr);
  ComplementWS w;
  layout_complement(measure, n, nr, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_complement_count: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_complement(arena, n, nr, &w);
  plan->magic = 0;
  plan->n = n;
  plan->nr = nr;
  plan->ready = 0;
  plan->b0 = b0;
  plan->b1 = b1;
  ClusterTable tab = {};
  if (n > 0) BF_TRY(cluster_run(region, s, e, n, nr, 0, 1, w.cl, false, st, &tab, nullptr));
  tab.n_clusters = w.cl.header;
  if (n == 0) BF_CUDA(cudaMemsetAsync(w.cl.header, 0, sizeof(i64), st));
  plan->tab = tab;
  BF_LAUNCH(PC_CLUSTER, 0, st,
            complement_regions_kernel<<<(unsigned)ceil_div(nr, 256), 256, 0, st>>>(tab, n > 0, b0, b1, nr, w.r_lo, w.r_hi,
                                                                                  w.r_off, w.r_first));
  BF_LAUNCH(PC_CLUSTER, 0, st, complement_offsets_kernel<<<1, 1, 0, st>>>(w.r_off, nr, w.total));
  i64 host[2] = {0, 0};
  BF_CUDA(cudaMemcpyAsync(&host[0], w.total, sizeof(i64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaMemcpyAsync(&host[1], w.cl.header, sizeof(i64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  plan->n_gaps = host[0];
  plan->n_clusters = host[1];
  plan->magic = COMPLEMENT_MAGIC;
  plan->ready = 1;
  return BF_OK;
}

static int complement_emit_impl(const ComplementPlan* plan, void* ws_ptr, i32* reg_out, i64* s_out, i64* e_out,
                                cudaStream_t st) {
  if (!plan || plan->magic != COMPLEMENT_MAGIC || !plan->ready) {
    set_error("bf_complement_emit: plan was not produced by a successful bf_complement_count");
    return BF_ERR_STATE;
  }
  if (plan->n_gaps == 0) return BF_OK;
  if (!reg_out || !s_out || !e_out) {
    set_error("bf_complement_emit: null output pointer");
    return BF_ERR_ARG;
  }
  Arena arena(ws_ptr);
  ComplementWS w;
  layout_complement(arena, plan->n, plan->nr, &w);
  BF_LAUNCH(PC_CLUSTER, 0, st,
            complement_edges_kernel<<<(unsigned)ceil_div(plan->nr, 256), 256, 0, st>>>(
                plan->tab, plan->b0, plan->b1, plan->nr, w.r_lo, w.r_hi, w.r_off, w.r_first, reg_out, s_out, e_out));
  if (plan->n_clusters > 1) {
    BF_LAUNCH(PC_CLUSTER, 20 * plan->n_clusters, st,
              complement_inner_kernel<<<(unsigned)ceil_div(plan->n_clusters, 256), 256, 0, st>>>(
                  plan->tab, plan->n_clusters, w.r_lo, w.r_hi, w.r_off, w.r_first, reg_out, s_out, e_out));
  }
  return BF_OK;
}

}  // namespace bf

#include "subtract.cuh"

extern "C" {

size_t bf_subtract_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups) {
  if (n1 < 0 || n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::SubtractWS w;
  bf::layout_subtract(a, n1, n2, n_groups, &w);
  return a.off;
}
int bf_subtract_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1, const int32_t* grp2,
                      const int64_t* start2, const int64_t* end2, int64_t n2, int32_t n_groups, void* ws,
                      size_t ws_bytes, bf_plan* plan, void* stream, int64_t* n_rows_out) {
  bf::SubtractPlan* p = (bf::SubtractPlan*)plan;
  int rc = bf::subtract_count_impl((const i32*)grp1, (const i64*)start1, (const i64*)end1, n1, (const i32*)grp2,
                                   (const i64*)start2, (const i64*)end2, n2, n_groups, ws, ws_bytes, p,
                                   (cudaStream_t)stream);
  if (rc == BF_OK && n_rows_out) *n_rows_out = p->n_rows;
  return rc;
}
int bf_subtract_emit(const bf_plan* plan, void* ws, int64_t* row_out, int64_t* start_out, int64_t* end_out,
                     int64_t* sub_index_out, void* stream) {
  return bf::subtract_emit_impl((const bf::SubtractPlan*)plan, ws, (i64*)row_out, (i64*)start_out, (i64*)end_out,
                                (i64*)sub_index_out, (cudaStream_t)stream);
}

size_t bf_cluster_workspace_bytes(int64_t n, int32_t n_groups) {
  if (n < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::ClusterWS w;
  bf::layout_cluster(a, n, n_groups, &w);
  return a.off;
}
int bf_cluster_count(const int32_t* grp, const int64_t* start, const int64_t* end, int64_t n, int32_t n_groups,
                     int64_t min_dist, int32_t has_min_dist, void* ws, size_t ws_bytes, bf_plan* plan, void* stream,
                     int64_t* cluster_ids_out, int64_t* n_clusters_out) {
  bf::ClusterPlan* p = (bf::ClusterPlan*)plan;
  int rc = bf::cluster_count_impl((const i32*)grp, (const i64*)start, (const i64*)end, n, n_groups, min_dist,
                                  has_min_dist, ws, ws_bytes, p, (cudaStream_t)stream, (i64*)cluster_ids_out);
  if (rc == BF_OK && n_clusters_out) *n_clusters_out = p->n_clusters;
  return rc;
}
int bf_cluster_emit(const bf_plan* plan, void* ws, int32_t* cgrp_out, int64_t* cstart_out, int64_t* cend_out,
                    int64_t* ccount_out, const int64_t* cluster_ids, int64_t* row_start_out, int64_t* row_end_out,
                    void* stream) {
  (void)ws;
  return bf::cluster_emit_impl((const bf::ClusterPlan*)plan, (i32*)cgrp_out, (i64*)cstart_out, (i64*)cend_out,
                               (i64*)ccount_out, (const i64*)cluster_ids, (i64*)row_start_out, (i64*)row_end_out,
                               (cudaStream_t)stream);
}

size_t bf_coverage_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups) {
  (void)n1;
  if (n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::CoverageWS w;
  bf::layout_coverage(a, n2, n_groups, &w);
  return a.off;
}
int bf_coverage(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1, const int32_t* grp2,
                const int64_t* start2, const int64_t* end2, int64_t n2, int32_t n_groups, void* ws, size_t ws_bytes,
                void* stream, int64_t* coverage_out) {
  return bf::coverage_impl((const i32*)grp1, (const i64*)start1, (const i64*)end1, n1, (const i32*)grp2,
                           (const i64*)start2, (const i64*)end2, n2, n_groups, ws, ws_bytes, (cudaStream_t)stream,
                           (i64*)coverage_out);
}

size_t bf_complement_workspace_bytes(int64_t n, int32_t n_regions) {
  if (n < 0 || n_regions < 1) return 0;
  bf::Arena a(nullptr);
  bf::ComplementWS w;
  bf::layout_complement(a, n, n_regions, &w);
  return a.off;
}
int bf_complement_count(const int32_t* region, const int64_t* start, const int64_t* end, int64_t n,
                        const int64_t* bounds_lo, const int64_t* bounds_hi, int32_t n_regions, void* ws, size_t ws_bytes,
                        bf_plan* plan, void* stream, int64_t* n_gaps_out) {
  bf::ComplementPlan* p = (bf::ComplementPlan*)plan;
  int rc = bf::complement_count_impl((const i32*)region, (const i64*)start, (const i64*)end, n, (const i64*)bounds_lo,
                                     (const i64*)bounds_hi, n_regions, ws, ws_bytes, p, (cudaStream_t)stream);
  if (rc == BF_OK && n_gaps_out) *n_gaps_out = p->n_gaps;
  return rc;
}
int bf_complement_emit(const bf_plan* plan, void* ws, int32_t* region_out, int64_t* start_out, int64_t* end_out,
                       void* stream) {
  return bf::complement_emit_impl((const bf::ComplementPlan*)plan, ws, (i32*)region_out, (i64*)start_out,
                                  (i64*)end_out, (cudaStream_t)stream);
}

}  // extern "C"
