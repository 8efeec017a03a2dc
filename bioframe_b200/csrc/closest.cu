// bf_closest_count / bf_closest_emit: the whole of ops._closest_intidxs
// (src/bioframe/ops.py:919-1040) for all groups in one set of launches, i.e. the
// per-group loop around arrops.closest_intervals (arrops.py:601-754) and
// arrops._closest_intervals_nooverlap (arrops.py:506-598).
//
// The reference materialises every candidate event (all overlaps, k left, k right
// neighbours of every query) and lexsorts the lot by (query, distance, target)
// -- 97 % of its run time.  Here nothing global is materialised: the targets
// are sorted twice (by linear start and by linear end, stable, so ties are in
// row order exactly like a stable argsort), and every query is one thread that
//   * finds its four positions by binary search (arrops.py:338-339, 569, 587),
//   * walks its candidates -- overlaps (distance 0), the last k targets ending
//     at or before its start (distance start - end + 1, arrops.py:724), the
//     first k targets starting at or after its end (arrops.py:725) -- and
//   * keeps the k smallest by (distance, target row) with an insertion into its
//     own slice of the output (arrops.py:737-754).
// Candidates form a multiset exactly as in the reference (a point query
// reports a target starting at the same position both as an overlap and as a
// right neighbour).
//
// Output order (ops.py:997-1035): per group in rank order, the matched rows of
// the group's queries in ascending row order, then (row, -1) for its unmatched
// queries in ascending row order.  Queries are put in (group, row) order by one
// stable radix pass over the group rank; two fused exclusive scans (rows per
// query, unmatched flags) and a per-group table place every slice.
#include "common.cuh"
#include "keys.cuh"
#include "lookback.cuh"
#include "lut.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "sorted_set.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

struct ClosestWS {
  SortBufs by_start, by_end;  // the two target orders
  RadixScratch radix;
  LayoutScratch layout;
  u64* pmax;        // [n2] running max of linear end' over the start order
  u64 *bmax1, *bmax2, *bmax3;  // max of linear end' over blocks of 32 / 1024 / 32768 consecutive sorted targets
  i64* seg;         // [nb+1] first sorted target of each group (same for both orders)
  u64 *qkeyA, *qkeyB;  // [n1] (rank << 32 | row) of the queries
  i64* qseg;        // [nb+1] first query of each group in (group, row) order
  uint4* qpos;      // [n1] (loA, hiA, stop, begin) of every query, in (group, row) order
  u64 *rows, *lone; // [n1+1] exclusive scans: output rows, unmatched queries
  ScanState ss_rows, ss_lone, ss_tmp;
  i64 *obase, *ubase;  // [nb+1] first output row of a group's matched / unmatched block
  i64* header;      // n_rows, n_matched_rows
  u64* nq;          // [1] device count for the query sort
  u32 *lutS, *lutE; // direct-address tables over the two target orders
};
static void layout_closest(Arena& a, i64 n1, i64 n2, i32 nb, ClosestWS* w) {
  take_sort_bufs(a, n2, &w->by_start);
  take_sort_bufs(a, n2, &w->by_end);
  w->radix = take_radix_scratch(a, (u64)(n1 > n2 ? n1 : n2));
  w->layout = take_layout_scratch(a, nb);
  w->pmax = a.take<u64>(n2 + 1);
  w->bmax1 = a.take<u64>((n2 >> 5) + 2);
  w->bmax2 = a.take<u64>((n2 >> 10) + 2);
  w->bmax3 = a.take<u64>((n2 >> 15) + 2);
  w->seg = a.take<i64>(nb + 2);
  w->qkeyA = a.take<u64>(n1 + 1);
  w->qkeyB = a.take<u64>(n1 + 1);
  w->qseg = a.take<i64>(nb + 2);
  w->qpos = a.take<uint4>(n1 + 1);
  w->rows = a.take<u64>(n1 + 2);
  w->lone = a.take<u64>(n1 + 2);
  const u64 tiles = (u64)scan_tiles((u64)(n1 > n2 ? n1 : n2));
  w->ss_rows = take_scan_state(a, tiles);
  w->ss_lone = take_scan_state(a, tiles);
  w->ss_tmp = take_scan_state(a, tiles);
  w->obase = a.take<i64>(nb + 2);
  w->ubase = a.take<i64>(nb + 2);
  w->header = a.take<i64>(4);
  w->nq = a.take<u64>(1);
  w->lutS = a.take<u32>(lut_words(n2));
  w->lutE = a.take<u32>(lut_words(n2));
}

struct ClosestArgs {  // everything a query thread needs (kernel parameter)
  const i32* grp1;
  const i64 *s1, *e1;
  const u8* along;   // may be null: every query looks along the axis
  const u64 *kS, *kE;  // target keys in start order / end order
  const u32 *idS, *idE;
  const u64* pmax;
  const u64 *bmax1, *bmax2, *bmax3;
  const i64 *seg, *qseg;
  const u64* qkey;   // (rank << 32 | row) in (group, row) order; null: identity (one group)
  KeyLayout L;
  Lut lutS, lutE;
  i64 n1, n2;
  i32 k, self, ignore_overlaps, k_up, k_dn;
};

struct ClosestPlan {
  u64 magic;
  i32 ready, pad;
  i64 n_rows, n_matched;
  ClosestArgs args;
};
static_assert(sizeof(ClosestPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 CLOSEST_MAGIC = 0x6266436c6f736573ull;

// ---------------------------------------------------------------- kernels

// running max of linear end' (raw-length keys: end' = end + (len == 0)) in start order
__device__ __forceinline__ u64 key_ce_adj(u64 k, const KeyLayout& L) {
  const u64 len = k & L.lmask;
  return (k >> L.LB) + len + (len == 0);
}
__global__ void __launch_bounds__(SC_THREADS)
    runmax_adj_kernel(const u64* __restrict__ keys, u64 n, KeyLayout L, ScanState ss, u64* __restrict__ pmax) {
  __shared__ u64 s_tmp[8];
  __shared__ u64 s_pre;
  __shared__ u32 s_ticket;
  const u32 tid = threadIdx.x;
  const u64 tile = take_ticket(ss.ticket, &s_ticket);
  const u64 base = tile * SC_TILE + (u64)tid * SC_IPT;
  u64 v[SC_IPT];
  u64 mine = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    v[j] = base + j < n ? key_ce_adj(keys[base + j], L) : 0ull;
    mine = v[j] > mine ? v[j] : mine;
  }
  u64 before = block_excl_max_u64(mine, s_tmp);
  const u64 tile_max = block_max_u64(mine, s_tmp);
  if (tid < 32) {
    const u64 pre = lookback_exclusive_max(ss.status, tile, tile_max);
    if (tid == 0) s_pre = pre;
  }
  __syncthreads();
  u64 run = before > s_pre ? before : s_pre;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    run = v[j] > run ? v[j] : run;
    if (base + j < n) pmax[base + j] = run;
  }
}

// bmax[b] = max over entries [b << 5, (b << 5) + 32) of the level below (level 0: linear end' of the keys)
__global__ void __launch_bounds__(256)
    block_max_kernel(const u64* __restrict__ keys, const u64* __restrict__ below, i64 n_below, KeyLayout L,
                     u64* __restrict__ out) {
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if ((b << 5) >= n_below) return;
  u64 mx = 0;
  for (i64 i = b << 5; i < n_below && i < (b << 5) + 32; ++i) {
    const u64 v = keys ? key_ce_adj(keys[i], L) : below[i];
    mx = v > mx ? v : mx;
  }
  out[b] = mx;
}

// Walk the targets that start before the query (positions < from, downwards) and still reach past its
// start xs: fn(m) for every such position.  The running max of end' ends the walk as soon as nothing
// earlier can reach xs; whole blocks of 32 / 1024 / 32768 targets whose max end' does not reach xs are
// skipped, so one very long early target does not make every query scan everything in between
// (cost ~ matches + blocks between them, not the distance to the long target).
template <typename F>
__device__ __forceinline__ void walk_stabbing(const ClosestArgs& A, i64 from, u64 xs, F fn);

__global__ void __launch_bounds__(256) query_keys_kernel(const i32* __restrict__ grp, i64 n, u64* __restrict__ qkey) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) qkey[i] = ((u64)(u32)grp[i] << 32) | (u64)i;
}
__global__ void query_segments_kernel(const u64* __restrict__ qkey, i64 n, i32 nb, i64* __restrict__ qseg) {
  const i32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nb) return;
  if (!qkey) {  // one group, identity order
    qseg[c] = c == 0 ? 0 : n;
    return;
  }
  qseg[c] = lower_bound_by(0, n, [&](i64 m) { return (qkey[m] >> 32) < (u64)c; });
}

template <typename F>
__device__ __forceinline__ void walk_stabbing(const ClosestArgs& A, i64 from, u64 xs, F fn) {
  const KeyLayout& L = A.L;
  i64 m = from - 1;
  while (m >= 0 && A.pmax[m] > xs) {
    if ((m & 31) == 31 && A.bmax1[m >> 5] <= xs) {  // at the top of a block of 32 that holds no match
      if ((m & 1023) == 1023 && A.bmax2[m >> 10] <= xs) {
        if ((m & 32767) == 32767 && A.bmax3[m >> 15] <= xs)
          m -= 32768;
        else
          m -= 1024;
      } else {
        m -= 32;
      }
      continue;
    }
    if (key_ce_adj(A.kS[m], L) > xs && !fn(m)) return;
    --m;
  }
}

struct Probe {
  i64 row;        // query row
  u32 rank;
  i64 t0, t1;     // the group's targets: [t0, t1) in both orders
  i64 loA, hiA;   // start order: targets starting inside [s1, e1')
  i64 stop;       // end order: targets ending at or before s1 are [t0, stop)
  i64 begin;      // start order: targets starting at or after e1 are [begin, t1)
  u64 xs, xe;     // linear start / raw end of the query
  i32 k_left, k_right;
};

__device__ __forceinline__ void closest_probe(const ClosestArgs& A, i64 t, Probe* P) {
  const KeyLayout& L = A.L;
  i64 row = t;
  u32 rank = 0;
  if (A.qkey) {
    const u64 q = A.qkey[t];
    row = (i64)(q & 0xffffffffull);
    rank = (u32)(q >> 32);
  }
  P->row = row;
  P->rank = rank;
  const i64 t0 = A.seg[rank], t1 = A.seg[rank + 1];
  P->t0 = t0;
  P->t1 = t1;
  const i64 a = A.s1[row], b = min(A.e1[row], L.endcap);  // ends are clamped like the packed ones (keys.cuh)
  const u64 off = L.goff[rank];
  const u64 xs = off + ((u64)a - L.cmin), xe = off + ((u64)b - L.cmin);
  const u64 xe_adj = xe + (a == b);
  P->xs = xs;
  P->xe = xe;
  // positions on the whole sorted arrays: the query's coordinates lie inside its group's stretch of the
  // linear axis, so targets of other groups fall entirely below or above every probe
  P->loA = lut_count_below<false>(A.kS, A.lutS, xs);
  P->begin = lut_count_below<false>(A.kS, A.lutS, xe);
  P->hiA = (a == b) ? lut_count_below<false>(A.kS, A.lutS, xe_adj) : P->begin;
  P->stop = lut_count_below<true>(A.kE, A.lutE, xs);
  const bool fwd = A.along ? A.along[row] != 0 : true;
  P->k_left = fwd ? A.k_up : A.k_dn;   // arrops.py:667-706: "-" rows look right for upstream
  P->k_right = fwd ? A.k_dn : A.k_up;
}

// number of overlapping targets of the query, counted only up to `need`
__device__ __forceinline__ i64 count_overlaps_upto(const ClosestArgs& A, const Probe& P, i64 need) {
  const KeyLayout& L = A.L;
  i64 c = P.hiA - P.loA;
  if (A.self && c > 0) {  // drop the query itself if it sits in its own A range
    for (i64 m = P.loA; m < P.hiA; ++m)
      if ((i64)A.idS[m] == P.row) {
        --c;
        break;
      }
  }
  if (c < need) {
    walk_stabbing(A, P.loA, P.xs, [&](i64 m) {
      if (!(A.self && (i64)A.idS[m] == P.row)) ++c;
      return c < need;  // stop as soon as enough are known
    });
  }
  return c;
}

// rows this query contributes: min(k, #overlaps + #left + #right)
__global__ void __launch_bounds__(SC_THREADS)
    closest_count_kernel(ClosestArgs A, ScanState ss_rows, ScanState ss_lone, uint4* __restrict__ qpos,
                         u64* __restrict__ rows_out, u64* __restrict__ lone_out) {
  __shared__ u64 s_tmp[8];
  __shared__ u64 s_pre[2];
  __shared__ u32 s_ticket;
  const u32 tid = threadIdx.x;
  const u64 tile = take_ticket(ss_rows.ticket, &s_ticket);
  const i64 base = (i64)tile * SC_TILE + (i64)tid * SC_IPT;
  u32 cnt[SC_IPT];
  u64 sum = 0, nlone = 0;
#pragma unroll 1
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 t = base + j;
    cnt[j] = 0;
    if (t < A.n1) {
      Probe P;
      closest_probe(A, t, &P);
      i64 nl = P.stop - P.t0, nr = P.t1 - P.begin;
      if (nl > P.k_left) nl = P.k_left;
      if (nr > P.k_right) nr = P.k_right;
      i64 c = nl + nr;
      if (c < A.k && !A.ignore_overlaps) c += count_overlaps_upto(A, P, (i64)A.k - c);
      if (c > A.k) c = A.k;
      cnt[j] = (u32)c;
      qpos[t] = make_uint4((u32)P.loA, (u32)P.hiA, (u32)P.stop, (u32)P.begin);
      sum += (u64)c;
      nlone += c == 0;
    }
  }
  u64 tot_rows, tot_lone;
  u64 run = block_excl_sum_u64(sum, s_tmp, &tot_rows);
  u64 lrun = block_excl_sum_u64(nlone, s_tmp, &tot_lone);
  if (tid < 32) {
    const u64 p0 = lookback_exclusive(ss_rows.status, tile, tot_rows);
    const u64 p1 = lookback_exclusive(ss_lone.status, tile, tot_lone);
    if (tid == 0) {
      s_pre[0] = p0;
      s_pre[1] = p1;
    }
  }
  __syncthreads();
  run += s_pre[0];
  lrun += s_pre[1];
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 t = base + j;
    if (t < A.n1) {
      rows_out[t] = run;
      lone_out[t] = lrun;
    }
    run += cnt[j];
    lrun += (t < A.n1 && cnt[j] == 0);
    if (t + 1 == A.n1) {
      rows_out[A.n1] = run;
      lone_out[A.n1] = lrun;
    }
  }
}

// per-group output bases; single thread (n_groups is tens)
__global__ void closest_table_kernel(const u64* __restrict__ rows, const u64* __restrict__ lone,
                                     const i64* __restrict__ qseg, i32 nb, i64* __restrict__ obase,
                                     i64* __restrict__ ubase, i64* __restrict__ header) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  i64 base = 0, matched = 0;
  for (i32 c = 0; c < nb; ++c) {
    const i64 m = (i64)(rows[qseg[c + 1]] - rows[qseg[c]]);
    const i64 u = (i64)(lone[qseg[c + 1]] - lone[qseg[c]]);
    obase[c] = base - (i64)rows[qseg[c]];       // output row of a matched slice = rows[t] + obase[group]
    ubase[c] = base + m - (i64)lone[qseg[c]];   // output row of an unmatched query = lone[t] + ubase[group]
    base += m + u;
    matched += m;
  }
  header[0] = base;
  header[1] = matched;
}

// insert (d, j) into the slice sorted by (d, j), keeping at most cap entries
__device__ __forceinline__ void slice_insert(i64* __restrict__ dist, i64* __restrict__ tgt, i32& len, i32 cap, i64 d,
                                             i64 j) {
  if (len == cap) {
    if (cap == 0) return;
    const i64 ld = dist[len - 1], lj = tgt[len - 1];
    if (d > ld || (d == ld && j >= lj)) return;  // not better than the current worst (equal keys: duplicates
                                                  // are indistinguishable, dropping the newcomer is the same)
    --len;
  }
  i32 at = len;
  while (at > 0) {
    const i64 pd = dist[at - 1], pj = tgt[at - 1];
    if (pd < d || (pd == d && pj <= j)) break;
    dist[at] = pd;
    tgt[at] = pj;
    --at;
  }
  dist[at] = d;
  tgt[at] = j;
  ++len;
}

__global__ void __launch_bounds__(256)
    closest_emit_kernel(ClosestArgs A, const uint4* __restrict__ qpos, const u64* __restrict__ rows,
                        const u64* __restrict__ lone, const i64* __restrict__ obase, const i64* __restrict__ ubase,
                        i64* __restrict__ ev1, i64* __restrict__ ev2) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n1) return;
  const KeyLayout& L = A.L;
  i64 row = t;
  u32 rank = 0;
  if (A.qkey) {
    const u64 q = A.qkey[t];
    row = (i64)(q & 0xffffffffull);
    rank = (u32)(q >> 32);
  }
  const i32 cap = (i32)(rows[t + 1] - rows[t]);
  if (cap == 0) {
    const i64 out = (i64)lone[t] + ubase[rank];
    ev1[out] = row;
    ev2[out] = -1;
    return;
  }
  const i64 out = (i64)rows[t] + obase[rank];
  // working slice: thread-local arrays for the usual small k (the insertion shifts stay in L1 / registers),
  // otherwise the query's slice of the output itself (ev1 holds the distances until the very end)
  constexpr int K_LOCAL = 32;
  i64 loc_d[K_LOCAL], loc_j[K_LOCAL];
  const bool local = cap <= K_LOCAL;
  i64* dist = local ? loc_d : ev1 + out;
  i64* tgt = local ? loc_j : ev2 + out;
  i32 len = 0;
  const uint4 qp = qpos[t];
  const i64 loA = qp.x, hiA = qp.y, stop = qp.z, begin = qp.w;
  const i64 t0 = A.seg[rank], t1 = A.seg[rank + 1];
  const i64 a = A.s1[row], b = min(A.e1[row], L.endcap);
  const u64 off = L.goff[rank];
  const u64 xs = off + ((u64)a - L.cmin), xe = off + ((u64)b - L.cmin);
  const bool fwd = A.along ? A.along[row] != 0 : true;
  const i32 k_left = fwd ? A.k_up : A.k_dn, k_right = fwd ? A.k_dn : A.k_up;
  if (!A.ignore_overlaps) {  // distance 0 (arrops.py:650-659)
    for (i64 m = loA; m < hiA; ++m) {
      const i64 j = A.idS[m];
      if (!(A.self && j == row)) slice_insert(dist, tgt, len, cap, 0, j);
    }
    walk_stabbing(A, loA, xs, [&](i64 m) {
      const i64 j = A.idS[m];
      if (!(A.self && j == row)) slice_insert(dist, tgt, len, cap, 0, j);
      return true;
    });
  }
  {  // left neighbours: last k_left of the targets ending at or before the start (arrops.py:560-576, 724)
    i64 from = stop - k_left;
    if (from < t0) from = t0;
    for (i64 m = stop - 1; m >= from; --m)
      slice_insert(dist, tgt, len, cap, (i64)(xs - key_cs(A.kE[m], L)) + 1, (i64)A.idE[m]);
  }
  {  // right neighbours: first k_right of the targets starting at or after the end (arrops.py:578-596, 725)
    i64 to = begin + k_right;
    if (to > t1) to = t1;
    for (i64 m = begin; m < to; ++m)
      slice_insert(dist, tgt, len, cap, (i64)(key_cs(A.kS[m], L) - xe) + 1, (i64)A.idS[m]);
  }
  // len == cap by construction of the count
  if (local) {
    for (i32 x = 0; x < cap; ++x) {
      ev1[out + x] = row;
      ev2[out + x] = loc_j[x];
    }
  } else {
    for (i32 x = 0; x < cap; ++x) dist[x] = row;
  }
}

// ---------------------------------------------------------------- host side

static int closest_count_impl(const i32* grp1, const i64* s1, const i64* e1, i64 n1, const i32* grp2, const i64* s2,
                              const i64* e2, i64 n2, i32 nb, i32 k, i32 flags, const u8* along, void* ws_ptr,
                              size_t ws_bytes, ClosestPlan* plan, cudaStream_t st) {
  if (n1 < 0 || n2 < 0 || nb < 1 || k < 1 || !plan) {
    set_error("bf_closest_count: bad argument (n1=%lld n2=%lld n_groups=%d k=%d)", (long long)n1, (long long)n2, nb, k);
    return BF_ERR_ARG;
  }
  if (n1 >= (1ll << 31) || n2 >= (1ll << 31)) {
    set_error("bf_closest_count: more than 2^31-1 rows in one set");
    return BF_ERR_RANGE;
  }
  if ((n1 > 0 && (!s1 || !e1)) || (n2 > 0 && (!s2 || !e2)) || (nb > 1 && ((n1 > 0 && !grp1) || (n2 > 0 && !grp2)))) {
    set_error("bf_closest_count: null input pointer");
    return BF_ERR_ARG;
  }
  Arena measure(nullptr);
  ClosestWS w;
  layout_closest(measure, n1, n2, nb, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_closest_count: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_closest(arena, n1, n2, nb, &w);
  plan->magic = 0;
  plan->ready = 0;
  plan->n_rows = plan->n_matched = 0;

  // layout over both sets (point-adjusted extent so that end' fits), raw lengths in the keys
  const RowSet sets[2] = {{grp1, s1, e1, n1}, {grp2, s2, e2, n2}};
  KeyLayout L;
  BF_TRY(measure_layout(sets, 2, nb, 1, 1, w.layout, st, &L, "bf_closest_count"));
  if (L.B > 62) {
    set_error("bf_closest_count: coordinate span too wide (%d bits)", L.B);
    return BF_ERR_RANGE;
  }
  SortResult rS, rE;
  BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_NO_TIE_FIX | presorted_mode(L, 1), w.by_start, w.radix, st, &rS));
  BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_NO_TIE_FIX | KEY_BY_END, w.by_end, w.radix, st, &rE));
  BF_LAUNCH(PC_SEGMENT, 0, st,
            segment_table_kernel<<<(unsigned)ceil_div(nb + 1, 128), 128, 0, st>>>(rS.keys, n2, L, nb, w.seg));
  if (n2 > 0) {
    const unsigned tiles = (unsigned)scan_tiles((u64)n2);
    BF_TRY(reset_scan_state(w.ss_tmp, tiles, st));
    BF_LAUNCH(PC_RUNMAX, 16 * n2, st, runmax_adj_kernel<<<tiles, SC_THREADS, 0, st>>>(rS.keys, (u64)n2, L, w.ss_tmp, w.pmax));
    const i64 nb1 = (n2 + 31) >> 5, nb2 = (nb1 + 31) >> 5, nb3 = (nb2 + 31) >> 5;
    BF_LAUNCH(PC_RUNMAX, 8 * n2, st, block_max_kernel<<<(unsigned)ceil_div(nb1, 256), 256, 0, st>>>(rS.keys, nullptr, n2, L, w.bmax1));
    BF_LAUNCH(PC_RUNMAX, 0, st, block_max_kernel<<<(unsigned)ceil_div(nb2, 256), 256, 0, st>>>(nullptr, w.bmax1, nb1, L, w.bmax2));
    BF_LAUNCH(PC_RUNMAX, 0, st, block_max_kernel<<<(unsigned)ceil_div(nb3, 256), 256, 0, st>>>(nullptr, w.bmax2, nb2, L, w.bmax3));
  }
  // queries in (group, row) order
  const u64* qkey = nullptr;
  if (nb > 1 && n1 > 0) {
    BF_LAUNCH(PC_CLOSEST, 12 * n1, st, query_keys_kernel<<<(unsigned)ceil_div(n1, 256), 256, 0, st>>>(grp1, n1, w.qkeyA));
    const RadixPlan qp = make_radix_plan(32, 32 + L.rank_bits);
    BF_CUDA(cudaMemsetAsync(w.radix.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
    BF_TRY(launch_radix_hist(w.qkeyA, (u64)n1, nullptr, qp, w.radix.hist, st, PC_CLOSEST));
    SortResult qr;
    BF_TRY(radix_sort_passes<false>(w.qkeyA, w.qkeyB, nullptr, nullptr, false, (u64)n1, nullptr, qp, w.radix, st, &qr,
                                    PC_CLOSEST));
    qkey = qr.keys;
  }
  BF_LAUNCH(PC_SEGMENT, 0, st,
            query_segments_kernel<<<(unsigned)ceil_div(nb + 1, 128), 128, 0, st>>>(qkey, n1, nb, w.qseg));

  ClosestArgs& A = plan->args;
  A.grp1 = grp1;
  A.s1 = s1;
  A.e1 = e1;
  A.along = along;
  A.kS = rS.keys;
  A.kE = rE.keys;
  A.idS = rS.vals;
  A.idE = rE.vals;
  A.pmax = w.pmax;
  A.bmax1 = w.bmax1;
  A.bmax2 = w.bmax2;
  A.bmax3 = w.bmax3;
  A.seg = w.seg;
  A.qseg = w.qseg;
  A.qkey = qkey;
  A.L = L;
  A.lutS = make_lut(w.lutS, n2, 0, 1ull << L.B, L.LB);
  A.lutE = make_lut(w.lutE, n2, 0, 1ull << L.B, L.LB);
  BF_TRY(launch_lut_build(rS.keys, n2, nullptr, A.lutS, w.lutS, st, PC_CLOSEST));
  BF_TRY(launch_lut_build(rE.keys, n2, nullptr, A.lutE, w.lutE, st, PC_CLOSEST));
  A.n1 = n1;
  A.n2 = n2;
  A.k = k;
  A.self = (flags & BF_CLOSEST_SELF) ? 1 : 0;
  A.ignore_overlaps = (flags & BF_CLOSEST_IGNORE_OVERLAPS) ? 1 : 0;
  A.k_up = (flags & BF_CLOSEST_IGNORE_UPSTREAM) ? 0 : k;
  A.k_dn = (flags & BF_CLOSEST_IGNORE_DOWNSTREAM) ? 0 : k;

  if (n1 > 0) {
    const unsigned tiles = (unsigned)scan_tiles((u64)n1);
    BF_TRY(reset_scan_state(w.ss_rows, tiles, st));
    BF_TRY(reset_scan_state(w.ss_lone, tiles, st));
    BF_LAUNCH(PC_CLOSEST, 52 * n1, st,
              closest_count_kernel<<<tiles, SC_THREADS, 0, st>>>(A, w.ss_rows, w.ss_lone, w.qpos, w.rows, w.lone));
  } else {
    BF_CUDA(cudaMemsetAsync(w.rows, 0, sizeof(u64), st));
    BF_CUDA(cudaMemsetAsync(w.lone, 0, sizeof(u64), st));
  }
  BF_LAUNCH(PC_TABLE, 0, st, closest_table_kernel<<<1, 1, 0, st>>>(w.rows, w.lone, w.qseg, nb, w.obase, w.ubase, w.header));
  i64 header[2];
  BF_CUDA(cudaMemcpyAsync(header, w.header, sizeof(header), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  plan->n_rows = header[0];
  plan->n_matched = header[1];
  plan->magic = CLOSEST_MAGIC;
  plan->ready = 1;
  return BF_OK;
}

static int closest_emit_impl(const ClosestPlan* plan, void* ws_ptr, i64* ev1, i64* ev2, cudaStream_t st) {
  if (!plan || plan->magic != CLOSEST_MAGIC || !plan->ready) {
    set_error("bf_closest_emit: plan was not produced by a successful bf_closest_count");
    return BF_ERR_STATE;
  }
  if (plan->n_rows == 0) return BF_OK;
  if (!ev1 || !ev2) {
    set_error("bf_closest_emit: null output pointer");
    return BF_ERR_ARG;
  }
  const ClosestArgs& A = plan->args;
  Arena arena(ws_ptr);
  ClosestWS w;
  layout_closest(arena, A.n1, A.n2, A.L.nb, &w);
  BF_LAUNCH(PC_CLOSEST, 40 * A.n1 + 16 * plan->n_rows, st,
            closest_emit_kernel<<<(unsigned)ceil_div(A.n1, 256), 256, 0, st>>>(A, w.qpos, w.rows, w.lone, w.obase, w.ubase,
                                                                              ev1, ev2));
  return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// bf_overlap_ordered_*: overlap(how='left') with the rows in (row of set 1, row of set 2) order -- what
// bioframe.overlap(how='left') returns by default (keep_order=True, ops.py:455-456, 549-550) when the indexes
// order like the rows.  The reference sorts the assembled frame; sorting 8.76e8 pairs on the device costs 4x
// the overlap itself (profiles/NOTES.md).  Here the pairs are BORN in that order: the queries are visited in
// row order and never sorted, only the targets are (twice, like count_overlaps):
//   count:  partners = #{start_t < end'_q} - #{end'_t <= start_q}        (two table-assisted probes)
//   emit:   targets starting inside the query are a contiguous range of the start order; those that started
//           earlier and still reach the query are found by the running-max walk of closest (walk_stabbing);
//           the query's few partner rows are sorted by row number in registers and written at its offset
//           (a query without partner writes (row, -1)).
struct OrderedWS {
  SortBufs by_start, by_end;
  RadixScratch radix;
  LayoutScratch layout;
  u64* pmax;
  u64 *bmax1, *bmax2, *bmax3;
  u32 *lutS, *lutE;
  u32 *cnt, *loc;  // [n1] partners of every query, output offset inside its tile
  u32* started;    // [n1] #{targets starting before the query's end'} (end of its range in the start order)
  // axes up to 2^32 (any real genome): everything a query touches per target packed next to each other
  uint2* rec;      // [n2] start order: x = linear end', y = row number
  u32* starts32;   // [n2] start order: linear start
  u32* ends32;     // [n2] end order: linear end'
  u64* spine;      // [tiles(n1) + 2]
  ScanState ss_tmp;
  i64* hdr;        // [0] rows, [1] overflow flag, [2] largest partner count
};
static void layout_ordered(Arena& a, i64 n1, i64 n2, i32 nb, OrderedWS* w) {
  take_sort_bufs(a, n2, &w->by_start);
  take_sort_bufs(a, n2, &w->by_end);
  w->radix = take_radix_scratch(a, (u64)n2);
  w->layout = take_layout_scratch(a, nb);
  w->pmax = a.take<u64>(n2 + 1);
  w->bmax1 = a.take<u64>((n2 >> 5) + 2);
  w->bmax2 = a.take<u64>((n2 >> 10) + 2);
  w->bmax3 = a.take<u64>((n2 >> 15) + 2);
  w->lutS = a.take<u32>(lut_words(n2));
  w->lutE = a.take<u32>(lut_words(n2));
  w->cnt = a.take<u32>(n1 + 1);
  w->loc = a.take<u32>(n1 + 1);
  w->started = a.take<u32>(n1 + 1);
  w->rec = a.take<uint2>(n2 + 1);
  w->starts32 = a.take<u32>(n2 + 1);
  w->ends32 = a.take<u32>(n2 + 1);
  w->spine = a.take<u64>(scan_tiles((u64)(n1 > 0 ? n1 : 1)) + 2);
  w->ss_tmp = take_scan_state(a, (u64)scan_tiles((u64)(n2 > 0 ? n2 : 1)));
  w->hdr = a.take<i64>(4);
}
struct OrderedPlan {
  u64 magic;
  i32 ready, pad;
  i64 n_rows, n_pairs, max_partners;
  ClosestArgs args;
};
static_assert(sizeof(OrderedPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 ORDERED_MAGIC = 0x62664f7264657265ull;
constexpr int ORDERED_SMALL = 32;  // partner lists up to this size are sorted in registers / local memory

// linear start and end' of query r (a point behaves as [x, x + 1), arrops.py:271-287; ends clamped like the packed ones)
__device__ __forceinline__ void ordered_query(const ClosestArgs& A, i64 r, u64* xs, u64* xe) {
  const u32 g = A.grp1 ? (u32)A.grp1[r] : 0u;
  const i64 a = A.s1[r];
  i64 b = A.e1[r];
  if (a == b) b += 1;
  b = min(b, A.L.endcap);
  const u64 off = A.L.goff[g];
  *xs = off + ((u64)a - A.L.cmin);
  *xe = off + ((u64)b - A.L.cmin);
}

// Queries arrive in arbitrary order, so every array they touch is a random DRAM access (the target arrays of
// the benchmark, 180 MB as 64-bit keys + ids + ends, do not fit the L2).  On axes up to 2^32 the targets are
// repacked: 8-byte records in start order (end', row: what the scan over a query's neighbourhood reads, a few
// adjacent cache lines; 16-byte records with the start included read 26 GB per 1e8 queries, twice this) and
// 4-byte arrays of the starts (start order) and the ends (end order) for the probes.
struct Ordered32 {
  const uint2* rec;   // null: the axis needs more than 32 bits, use the 64-bit arrays
  const u32* starts32;
  const u32* ends32;
};
__global__ void __launch_bounds__(256)
    ordered_pack32_kernel(const u64* __restrict__ kS, const u32* __restrict__ idS, const u64* __restrict__ kE, i64 n,
                          KeyLayout L, uint2* __restrict__ rec, u32* __restrict__ starts32, u32* __restrict__ ends32) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 k = kS[i];
  rec[i] = make_uint2((u32)key_ce_adj(k, L), idS[i]);
  starts32[i] = (u32)key_cs(k, L);
  ends32[i] = (u32)(kE[i] >> L.LB);
}
// #{records with start < x} / #{ends <= x}: one table sector, then a search inside the bucket
__device__ __forceinline__ i64 ordered_starts_below(const u32* __restrict__ starts32, const Lut& t, u64 x) {
  u64 b = x < t.base ? 0ull : (x - t.base) >> t.shift;
  if (b >= t.nbuckets) b = t.nbuckets - 1;
  i64 lo = t.tab[b], hi = t.tab[b + 1];
  if (x < t.base) lo = 0;
  while (lo < hi) {
    const i64 mid = lo + ((hi - lo) >> 1);
    if ((u64)starts32[mid] < x)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
__device__ __forceinline__ i64 ordered_ends_upto(const u32* __restrict__ ends32, const Lut& t, u64 x) {
  u64 b = x < t.base ? 0ull : (x - t.base) >> t.shift;
  if (b >= t.nbuckets) b = t.nbuckets - 1;
  i64 lo = t.tab[b], hi = t.tab[b + 1];
  if (x < t.base) lo = 0;
  while (lo < hi) {
    const i64 mid = lo + ((hi - lo) >> 1);
    if ((u64)ends32[mid] <= x)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(SC_THREADS)
    ordered_count_kernel(ClosestArgs A, Ordered32 P, u32* __restrict__ cnt_out, u32* __restrict__ started_out,
                         u32* __restrict__ loc_out, u64* __restrict__ tile_tot, i64* __restrict__ hdr) {
  __shared__ u64 s_tmp[8];
  const i64 tile0 = (i64)blockIdx.x * SC_TILE;
  u32 biggest = 0;
  // probes: rows strided over the block (coalesced reads of the query columns, coalesced writes of the counts)
#pragma unroll 1
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 r = tile0 + (i64)j * SC_THREADS + threadIdx.x;
    if (r >= A.n1) break;
    u32 c = 0, st = 0;
    if (A.n2 > 0) {
      u64 xs, xe;
      ordered_query(A, r, &xs, &xe);
      const i64 started = P.rec ? ordered_starts_below(P.starts32, A.lutS, xe) : lut_count_below<false>(A.kS, A.lutS, xe);
      const i64 ended = P.rec ? ordered_ends_upto(P.ends32, A.lutE, xs) : lut_count_below<true>(A.kE, A.lutE, xs);
      c = started > ended ? (u32)(started - ended) : 0u;
      st = (u32)started;
    }
    cnt_out[r] = c;
    started_out[r] = st;
    biggest = c > biggest ? c : biggest;
  }
  __syncthreads();  // the counts of the whole tile are visible to the block
  // offsets: every thread scans 8 consecutive rows (how='left': a query without partner still yields one row)
  const i64 base = tile0 + (i64)threadIdx.x * SC_IPT;
  u32 rows[SC_IPT];
  u64 mine = 0;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 r = base + j;
    const u32 c = r < A.n1 ? cnt_out[r] : 0u;
    rows[j] = r < A.n1 ? (c > 0 ? c : 1u) : 0u;
    mine += rows[j];
  }
  u64 tot;
  u64 run = block_excl_sum_u64(mine, s_tmp, &tot);
  const u64 block_biggest = block_max_u64((u64)biggest, s_tmp);
  if (threadIdx.x == 0) {
    tile_tot[blockIdx.x] = tot;
    if (tot >> 32) hdr[1] = 1;  // in-tile offsets are 32-bit
    if (block_biggest > 0) atomicMax(reinterpret_cast<unsigned long long*>(hdr + 2), (unsigned long long)block_biggest);
  }
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 r = base + j;
    if (r < A.n1) loc_out[r] = (u32)run;
    run += rows[j];
  }
}

// in-place heap sort of a long partner list (one thread; lists this long are rare)
template <typename T>
__device__ void ordered_heapsort(T* v, i64 n) {
  auto sift = [&](i64 root, i64 end) {
    for (;;) {
      i64 child = 2 * root + 1;
      if (child >= end) return;
      if (child + 1 < end && v[child] < v[child + 1]) ++child;
      if (v[root] >= v[child]) return;
      const T t = v[root];
      v[root] = v[child];
      v[child] = t;
      root = child;
    }
  };
  for (i64 i = n / 2 - 1; i >= 0; --i) sift(i, n);
  for (i64 end = n - 1; end > 0; --end) {
    const T t = v[0];
    v[0] = v[end];
    v[end] = t;
    sift(0, end);
  }
}

// One block = 256 consecutive queries, one per thread.  Every thread collects its partners -- the range of
// the start order inside the query, then the walk over earlier targets that still reach it -- into its
// stretch of a shared-memory staging area, sorted by row number; the block then writes its output rows
// coalesced (a thread's own rows are only ~9 x 8 bytes: written from the threads directly, every 32-byte
// sector of the result would be touched by four different store instructions).  A block whose queries have
// more rows than the staging area holds writes from the threads.
constexpr int OE_THREADS = 256;
constexpr int OE_CAP = 4096;          // staged output rows per block (avg 2240 for the benchmark workload)
constexpr u32 OE_NEAR_STEPS = 192;    // plain backward scan before the block-skipping walk takes over

// One block = 256 consecutive queries.
//   gather: one query per thread collects its partners UNSORTED into its stretch of a shared staging area --
//           the range of the start order inside the query, then a backward scan over earlier targets that
//           still reach it (their number is known, so the scan ends when all are found; a query whose
//           partners lie far back hands over to the block-skipping walk of closest);
//   rank:   every staged row, one per thread and step, counts the smaller rows of its own query: its rank is
//           its place.  The work of ordering, sum of c^2, is the same as an insertion sort per query, but
//           it is spread evenly over the block instead of leaving 255 threads waiting for the one query
//           with 40 partners (36 % of the stall samples at the barrier, 30 % in the insertion loop before);
//   write:  the block's rows leave coalesced.
// A block whose queries have more rows than the staging area holds writes and sorts per thread in global memory.
__global__ void __launch_bounds__(OE_THREADS)
    ordered_emit_kernel(ClosestArgs A, Ordered32 P, const u32* __restrict__ cnt_in,
                        const u32* __restrict__ started_in, const u32* __restrict__ loc_in,
                        const u64* __restrict__ spine, i64* __restrict__ ev1, i64* __restrict__ ev2) {
  __shared__ u32 s_raw[OE_CAP];             // partner rows as gathered, 0xffffffff = none
  __shared__ u32 s_sorted[OE_CAP];          // ... in row order within every query
  __shared__ u8 s_own[OE_CAP];              // query of the row, relative to the block's first query
  __shared__ u32 s_rel[OE_THREADS + 1];     // first output row of every query, relative to the block's first
  const u32 tid = threadIdx.x;
  const i64 r0 = (i64)blockIdx.x * OE_THREADS;
  const i64 r = r0 + tid;
  const i64 r_last = min(r0 + OE_THREADS, A.n1) - 1;
  const u32 nq = (u32)(r_last - r0 + 1);
  // 256 consecutive queries lie inside one scan tile (2048 rows): offsets relative to the block's first row
  const u32 loc0 = loc_in[r0];
  const u32 c_last = cnt_in[r_last];
  const u32 block_rows = loc_in[r_last] + (c_last > 0 ? c_last : 1u) - loc0;
  const i64 out0 = (i64)spine[r0 / SC_TILE] + loc0;
  // this thread's query: partner count, place, and where its partners are in the start order
  u32 c = 0, rel = block_rows;
  u64 xs = 0;
  i64 lo = 0;
  u32 n_inside = 0;
  if (r < A.n1) {
    c = cnt_in[r];
    rel = loc_in[r] - loc0;
    if (c > 0) {
      u64 xe;
      ordered_query(A, r, &xs, &xe);
      lo = P.rec ? ordered_starts_below(P.starts32, A.lutS, xs) : lut_count_below<false>(A.kS, A.lutS, xs);
      const i64 hi = (i64)started_in[r];
      n_inside = hi > lo ? (u32)(hi - lo) : 0u;
    }
  }
  s_rel[tid] = rel;
  if (tid == 0) s_rel[OE_THREADS] = block_rows;
  __syncthreads();
  // all partners of this thread's query, unsorted, handed to `put` one by one
  auto gather = [&](auto put, u32& k) {
    i64 from = lo;
    if (P.rec) {
      for (u32 i = 0; i < n_inside && k < c; ++i) put(P.rec[lo + i].y);
      const u32 xs32 = (u32)xs;
      i64 m = lo - 1;
      for (u32 step = 0; k < c && m >= 0 && step < OE_NEAR_STEPS; step += 4, m -= 4) {
        uint2 t[4];  // four records in flight: the loop is a chain of dependent round trips otherwise
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = P.rec[m - q >= 0 ? m - q : 0];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (m - q >= 0 && k < c && t[q].x > xs32) put(t[q].y);
      }
      from = m + 1 > 0 ? m + 1 : 0;
    } else {
      for (u32 i = 0; i < n_inside && k < c; ++i) put(A.idS[lo + i]);
    }
    if (k < c)
      walk_stabbing(A, from, xs, [&](i64 mm) {
        put(A.idS[mm]);
        return k < c;
      });
  };
  // The block's queries are taken in stretches whose rows fit the staging area.  Queries in arbitrary order
  // average 9 partners (2240 rows per block: one stretch); queries that arrive SORTED put a whole hotspot of
  // targets into one block (256 neighbours with 30+ partners each), which used to push the block onto the
  // per-thread path below -- 24 ms for sorted input against 18.7 ms for shuffled input (profiles/NOTES.md, r2).
  u32 qb = 0;
  while (qb < nq) {  // block-uniform
    const u32 first_row = s_rel[qb];
    u32 qe = qb + 1;
    {  // largest qe with rows(qb .. qe) <= OE_CAP (s_rel is non-decreasing; s_rel[nq .. 256] = block_rows)
      u32 a = qb + 1, b = nq;
      while (a < b) {
        const u32 mid = (a + b + 1) >> 1;
        if (s_rel[mid] - first_row <= (u32)OE_CAP)
          a = mid;
        else
          b = mid - 1;
      }
      qe = a;
    }
    const u32 seg_rows = s_rel[qe] - first_row;
    if (seg_rows > (u32)OE_CAP) {
      // one query with more rows than the staging area holds: written and sorted in global memory by its thread
      if (tid == qb) {
        const i64 at = out0 + rel;
        u32 k = 0;
        gather([&](u32 id) { ev2[at + k++] = id; }, k);
        ordered_heapsort(ev2 + at, (i64)k);
        for (u32 i = 0; i < k; ++i) ev1[at + i] = r;
      }
      qb = qe;
      continue;
    }
    if (tid >= qb && tid < qe) {
      const u32 at = rel - first_row;
      if (c == 0) {
        s_raw[at] = 0xffffffffu;
        s_own[at] = (u8)tid;
      } else {
        u32 k = 0;
        gather([&](u32 id) { s_raw[at + k++] = id; }, k);
        BF_DEVICE_CHECK(k == c && at + c <= (u32)OE_CAP);  // the walk found exactly the partners the count phase counted
        for (u32 i = 0; i < c; ++i) s_own[at + i] = (u8)tid;
      }
    }
    __syncthreads();
    // rank: every staged row counts the smaller rows of its own query; its rank is its place
    for (u32 x = tid; x < seg_rows; x += OE_THREADS) {
      const u32 o = s_own[x];
      const u32 a = s_rel[o] - first_row, b = s_rel[o + 1] - first_row;  // the query's rows: [a, b)
      BF_DEVICE_CHECK(o >= qb && o < qe && a <= x && x < b && b <= seg_rows);
      const u32 v = s_raw[x];
      u32 rank = 0;
      for (u32 j = a; j < b; ++j) rank += s_raw[j] < v;
      s_sorted[a + rank] = v;
    }
    __syncthreads();
    const i64 seg0 = out0 + first_row;
    for (u32 x = tid; x < seg_rows; x += OE_THREADS) {
      const u32 id = s_sorted[x];
      ev1[seg0 + x] = r0 + s_own[x];
      ev2[seg0 + x] = id == 0xffffffffu ? -1ll : (i64)id;
    }
    __syncthreads();  // the staging area is reused by the next stretch
    qb = qe;
  }
}

static int ordered_count_impl(const i32* grp1, const i64* s1, const i64* e1, i64 n1, const i32* grp2, const i64* s2,
                              const i64* e2, i64 n2, i32 nb, void* ws_ptr, size_t ws_bytes, OrderedPlan* plan,
                              cudaStream_t st) {
  if (n1 < 0 || n2 < 0 || nb < 1 || !plan) {
    set_error("bf_overlap_ordered_count: bad argument (n1=%lld n2=%lld n_groups=%d)", (long long)n1, (long long)n2, nb);
    return BF_ERR_ARG;
  }
  if (n1 >= (1ll << 31) || n2 >= (1ll << 31)) {
    set_error("bf_overlap_ordered_count: more than 2^31-1 rows in one set");
    return BF_ERR_RANGE;
  }
  plan->magic = 0;
  plan->ready = 0;
  plan->n_rows = plan->n_pairs = plan->max_partners = 0;
  if (n1 == 0) {
    plan->args = ClosestArgs{};
    plan->magic = ORDERED_MAGIC;
    plan->ready = 1;
    return BF_OK;
  }
  if (!s1 || !e1 || (n2 > 0 && (!s2 || !e2)) || (nb > 1 && (!grp1 || (n2 > 0 && !grp2)))) {
    set_error("bf_overlap_ordered_count: null input pointer");
    return BF_ERR_ARG;
  }
  Arena measure(nullptr);
  OrderedWS w;
  layout_ordered(measure, n1, n2, nb, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_overlap_ordered_count: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_ordered(arena, n1, n2, nb, &w);
  const RowSet sets[2] = {{grp1, s1, e1, n1}, {grp2, s2, e2, n2}};
  KeyLayout L;
  BF_TRY(measure_layout(sets, 2, nb, 1, 1, w.layout, st, &L, "bf_overlap_ordered_count"));
  if (L.B > 62) {
    set_error("bf_overlap_ordered_count: coordinate span too wide (%d bits)", L.B);
    return BF_ERR_RANGE;
  }
  ClosestArgs A = {};
  A.grp1 = nb > 1 ? grp1 : nullptr;
  A.s1 = s1;
  A.e1 = e1;
  A.L = L;
  A.n1 = n1;
  A.n2 = n2;
  if (n2 > 0) {
    SortResult rS, rE;
    BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_POINT_ADJUST | KEY_NO_TIE_FIX | presorted_mode(L, 1), w.by_start,
                     w.radix, st, &rS));
    BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_POINT_ADJUST | KEY_NO_TIE_FIX | KEY_BY_END, w.by_end, w.radix,
                     st, &rE));
    const unsigned tiles = (unsigned)scan_tiles((u64)n2);
    BF_TRY(reset_scan_state(w.ss_tmp, tiles, st));
    BF_LAUNCH(PC_RUNMAX, 16 * n2, st, runmax_adj_kernel<<<tiles, SC_THREADS, 0, st>>>(rS.keys, (u64)n2, L, w.ss_tmp, w.pmax));
    const i64 nb1 = (n2 + 31) >> 5, nb2 = (nb1 + 31) >> 5, nb3 = (nb2 + 31) >> 5;
    BF_LAUNCH(PC_RUNMAX, 8 * n2, st, block_max_kernel<<<(unsigned)ceil_div(nb1, 256), 256, 0, st>>>(rS.keys, nullptr, n2, L, w.bmax1));
    BF_LAUNCH(PC_RUNMAX, 0, st, block_max_kernel<<<(unsigned)ceil_div(nb2, 256), 256, 0, st>>>(nullptr, w.bmax1, nb1, L, w.bmax2));
    BF_LAUNCH(PC_RUNMAX, 0, st, block_max_kernel<<<(unsigned)ceil_div(nb3, 256), 256, 0, st>>>(nullptr, w.bmax2, nb2, L, w.bmax3));
    A.kS = rS.keys;
    A.kE = rE.keys;
    A.idS = rS.vals;
    A.idE = rE.vals;
    A.pmax = w.pmax;
    A.bmax1 = w.bmax1;
    A.bmax2 = w.bmax2;
    A.bmax3 = w.bmax3;
    A.lutS = make_lut(w.lutS, n2, 0, 1ull << L.B, L.LB);
    A.lutE = make_lut(w.lutE, n2, 0, 1ull << L.B, L.LB);
    BF_TRY(launch_lut_build(rS.keys, n2, nullptr, A.lutS, w.lutS, st, PC_COVERAGE));
    BF_TRY(launch_lut_build(rE.keys, n2, nullptr, A.lutE, w.lutE, st, PC_COVERAGE));
    if (L.B <= 32)
      BF_LAUNCH(PC_RUNMAX, 40 * n2, st,
                ordered_pack32_kernel<<<(unsigned)ceil_div(n2, 256), 256, 0, st>>>(rS.keys, rS.vals, rE.keys, n2, L, w.rec, w.starts32, w.ends32));
  }
  BF_CUDA(cudaMemsetAsync(w.hdr, 0, 4 * sizeof(i64), st));
  const Ordered32 P32 = {(n2 > 0 && L.B <= 32) ? w.rec : nullptr, w.starts32, w.ends32};
  const unsigned qtiles = (unsigned)scan_tiles((u64)n1);
  BF_LAUNCH(PC_SEARCH, 36 * n1, st, ordered_count_kernel<<<qtiles, SC_THREADS, 0, st>>>(A, P32, w.cnt, w.started, w.loc, w.spine, w.hdr));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.spine, (u64)qtiles, reinterpret_cast<u64*>(w.hdr)));
  i64 hdr[3];
  BF_CUDA(cudaMemcpyAsync(hdr, w.hdr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  if (hdr[1]) {
    set_error("bf_overlap_ordered_count: more than 2^32-1 output rows for 2048 consecutive queries");
    return BF_ERR_RANGE;
  }
  plan->n_rows = hdr[0];
  plan->max_partners = hdr[2];
  plan->args = A;
  plan->magic = ORDERED_MAGIC;
  plan->ready = 1;
  return BF_OK;
}

static int ordered_emit_impl(const OrderedPlan* plan, void* ws_ptr, i64* ev1, i64* ev2, cudaStream_t st) {
  if (!plan || plan->magic != ORDERED_MAGIC || !plan->ready) {
    set_error("bf_overlap_ordered_emit: plan was not produced by a successful bf_overlap_ordered_count");
    return BF_ERR_STATE;
  }
  if (plan->n_rows == 0) return BF_OK;
  if (!ev1 || !ev2) {
    set_error("bf_overlap_ordered_emit: null output pointer");
    return BF_ERR_ARG;
  }
  const ClosestArgs& A = plan->args;
  Arena arena(ws_ptr);
  OrderedWS w;
  layout_ordered(arena, A.n1, A.n2, A.L.nb, &w);
  BF_LAUNCH(PC_EMIT, 28 * A.n1 + 20 * plan->n_rows, st,
            ordered_emit_kernel<<<(unsigned)ceil_div(A.n1, OE_THREADS), OE_THREADS, 0, st>>>(
                A, Ordered32{(A.n2 > 0 && A.L.B <= 32) ? w.rec : nullptr, w.starts32, w.ends32}, w.cnt, w.started, w.loc, w.spine, ev1, ev2));
  return BF_OK;
}

}  // namespace bf

extern "C" {

size_t bf_overlap_ordered_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups) {
  if (n1 < 0 || n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::OrderedWS w;
  bf::layout_ordered(a, n1, n2, n_groups, &w);
  return a.off;
}
int bf_overlap_ordered_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                             const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                             int32_t n_groups, void* ws, size_t ws_bytes, bf_plan* plan, void* stream,
                             int64_t* n_rows_out, int64_t* max_partners_out) {
  bf::OrderedPlan* p = (bf::OrderedPlan*)plan;
  int rc = bf::ordered_count_impl((const i32*)grp1, (const i64*)start1, (const i64*)end1, n1, (const i32*)grp2,
                                  (const i64*)start2, (const i64*)end2, n2, n_groups, ws, ws_bytes, p,
                                  (cudaStream_t)stream);
  if (rc == BF_OK) {
    if (n_rows_out) *n_rows_out = p->n_rows;
    if (max_partners_out) *max_partners_out = p->max_partners;
  }
  return rc;
}
int bf_overlap_ordered_emit(const bf_plan* plan, void* ws, int64_t* events1_out, int64_t* events2_out, void* stream) {
  return bf::ordered_emit_impl((const bf::OrderedPlan*)plan, ws, (i64*)events1_out, (i64*)events2_out,
                               (cudaStream_t)stream);
}

size_t bf_closest_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups) {
  if (n1 < 0 || n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::ClosestWS w;
  bf::layout_closest(a, n1, n2, n_groups, &w);
  return a.off;
}
int bf_closest_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1, const int32_t* grp2,
                     const int64_t* start2, const int64_t* end2, int64_t n2, int32_t n_groups, int32_t k, int32_t flags,
                     const uint8_t* along, void* ws, size_t ws_bytes, bf_plan* plan, void* stream, int64_t* n_rows_out,
                     int64_t* n_matched_out) {
  bf::ClosestPlan* p = (bf::ClosestPlan*)plan;
  int rc = bf::closest_count_impl((const i32*)grp1, (const i64*)start1, (const i64*)end1, n1, (const i32*)grp2,
                                  (const i64*)start2, (const i64*)end2, n2, n_groups, k, flags, (const u8*)along, ws,
                                  ws_bytes, p, (cudaStream_t)stream);
  if (rc == BF_OK) {
    if (n_rows_out) *n_rows_out = p->n_rows;
    if (n_matched_out) *n_matched_out = p->n_matched;
  }
  return rc;
}
int bf_closest_emit(const bf_plan* plan, void* ws, int64_t* events1_out, int64_t* events2_out, void* stream) {
  return bf::closest_emit_impl((const bf::ClosestPlan*)plan, ws, (i64*)events1_out, (i64*)events2_out,
                               (cudaStream_t)stream);
}

}  // extern "C"
