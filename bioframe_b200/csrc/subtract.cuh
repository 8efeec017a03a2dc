// bf_subtract_*: df1 minus the union of df2 (ops.subtract, src/bioframe/ops.py:1243-1330), fused.
//
// The reference composes three frame operations: complement(df2) over [0, INT64_MAX) per chromosome
// (ops.py:1305-1309 -> clip to the view ops.py:1614-1622, merge with min_dist=0 and take the gaps
// arrops.py:482-503), a left overlap of df1 with those gaps returning the clipped coordinates
// (ops.py:1305-1316, 480-497), and a filter of the rows without a partner (ops.py:1318).  Here:
//
//   clip targets to start >= 0, drop those ending at or below 0   (one pass, 40 B per target)
//   merge them (cluster_run, min_dist = 0): disjoint sorted intervals (S_k, E_k) per group
//   gap j of a group with C merged intervals (j = 0..C):  [E_{j-1}, S_j)  with E_{-1} = 0, S_C = INT64_MAX
//   a query (s, e), e' = e + (e == s), overlaps the gaps with  S_j > s  and  E_{j-1} < e'  (arrops.py:338-342):
//       j in [ #{S_k <= s},  [0 < e'] + #{E_k < e'} )          -- two probes of the merged starts
//       (#{E_k < x} = a - 1 + [E_{a-1} < x] with a = #{S_k < x}, because the intervals are disjoint)
//   piece = (max(s, E_{j-1}), min(e, S_j))
//
// Queries are visited in input row order (keep_order=True of the reference, ops.py:549-550 on a RangeIndex),
// so they are never sorted: counts -> tile offsets + spine scan -> every query writes its pieces.
// Probes go through a direct-address table on the linear axis (lut.cuh) like coverage.
#pragma once

namespace bf {

struct SubtractWS {
  ClusterWS cl;
  i32* g2c;  // [n2] clipped targets (group nb = dropped)
  i64 *s2c, *e2c;
  ulonglong2* rec;   // [n2 + 1] merged intervals: x = linear start, y = linear end
  u32* lut;          // direct-address table over the linear starts
  i64* cg_first;     // [nb + 2] first merged interval of every group
  u32 *jlo, *cnt, *loc;  // [n1] first gap, number of pieces, output offset inside the tile
  u64* spine;        // [tiles(n1) + 2]
  i64* hdr;          // [0] = overflow flag, [1] = number of output rows
};
static void layout_subtract(Arena& a, i64 n1, i64 n2, i32 nb, SubtractWS* w) {
  layout_cluster(a, n2, nb + 1, &w->cl);
  w->g2c = a.take<i32>(n2 + 1);
  w->s2c = a.take<i64>(n2 + 1);
  w->e2c = a.take<i64>(n2 + 1);
  w->rec = a.take<ulonglong2>(n2 + 1);
  w->lut = a.take<u32>(lut_words(n2));
  w->cg_first = a.take<i64>(nb + 3);
  w->jlo = a.take<u32>(n1 + 1);
  w->cnt = a.take<u32>(n1 + 1);
  w->loc = a.take<u32>(n1 + 1);
  w->spine = a.take<u64>(scan_tiles((u64)n1) + 2);
  w->hdr = a.take<i64>(4);
}

struct SubtractView {  // what a query needs to find its gaps
  const ulonglong2* rec;
  const i64* cg_first;  // null: no targets at all
  Lut lut;
  KeyLayout L;
  i32 nb;  // groups of the call (the layout has one more: the dropped targets)
};
struct SubtractPlan {
  u64 magic;
  i64 n1, n2;
  i32 nb, ready;
  i64 n_rows;
  const i32* grp1;
  const i64 *s1, *e1;
  SubtractView view;
};
static_assert(sizeof(SubtractPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 SUBTRACT_MAGIC = 0x6266537562747263ull;

// the view [0, INT64_MAX) of ops.py:1307: starts below 0 move to 0, intervals ending at or below 0 leave
__global__ void __launch_bounds__(256)
    subtract_clip_kernel(const i32* __restrict__ g2, const i64* __restrict__ s2, const i64* __restrict__ e2, i64 n2,
                         i32 nb, i32* __restrict__ g2c, i64* __restrict__ s2c, i64* __restrict__ e2c) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const i64 s = s2[i], e = e2[i];
  const bool keep = (e > 0) || (e == 0 && s == 0);  // end' > 0 (a point at 0 counts as [0, 1))
  g2c[i] = keep ? (g2 ? g2[i] : 0) : nb;
  s2c[i] = keep ? (s > 0 ? s : 0) : 0;
  e2c[i] = keep ? e : 0;
}

__global__ void __launch_bounds__(256)
    subtract_table_kernel(ClusterTable tab, KeyLayout L, ulonglong2* __restrict__ rec) {
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= *tab.n_clusters) return;
  const u64 off = L.goff[tab.cgrp[c]];
  rec[c] = make_ulonglong2(off + ((u64)tab.cstart[c] - L.cmin), off + ((u64)tab.cend[c] - L.cmin));
}
__global__ void __launch_bounds__(256) subtract_groups_kernel(ClusterTable tab, i32 nb, i64* __restrict__ cg_first) {
  const i32 g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > nb + 1) return;
  const i64 C = *tab.n_clusters;
  cg_first[g] = lower_bound_by(0, C, [&](i64 m) { return tab.cgrp[m] < (u32)g; });
}

struct GapRange {
  i64 c0, C;     // merged intervals of the query's group: rec[c0 .. c0 + C)
  i64 jlo, jhi;  // candidate gaps
  u64 g0;        // linear offset of the group
};
__device__ __forceinline__ i64 sub_raw(u64 lin, u64 g0, const KeyLayout& L) { return (i64)(lin - g0 + L.cmin); }

__device__ __forceinline__ GapRange subtract_range(const SubtractView& v, u32 g, i64 s, i64 e) {
  GapRange r;
  r.c0 = 0;
  r.C = 0;
  r.g0 = 0;
  const i64 ep = e + (e == s && e != I64_MAX ? 1 : 0);
  u64 span = 0;
  if (v.cg_first) {
    r.c0 = v.cg_first[g];
    r.C = v.cg_first[g + 1] - r.c0;
    r.g0 = v.L.goff[g];
    span = v.L.goff[g + 1] - r.g0;
  }
  const u64* slin = reinterpret_cast<const u64*>(v.rec);  // entry m at word 2 * m
  i64 below_s = 0, below_e = 0;  // #{S_k <= s}, #{S_k < e'} inside the group
  if (r.C > 0) {
    const i64 cmin = (i64)v.L.cmin;
    if (s >= cmin) {
      const u64 d = (u64)s - v.L.cmin;
      below_s = d >= span ? r.C : lut_count_below<true>(slin, v.lut, r.g0 + d) - r.c0;
    }
    if (ep >= cmin) {
      const u64 d = (u64)ep - v.L.cmin;
      below_e = d >= span ? r.C : lut_count_below<false>(slin, v.lut, r.g0 + d) - r.c0;
    }
  }
  r.jlo = below_s;
  i64 ended = 0;  // #{E_k < e'}
  if (below_e > 0) ended = below_e - 1 + (sub_raw(v.rec[r.c0 + below_e - 1].y, r.g0, v.L) < ep ? 1 : 0);
  r.jhi = (ep > 0 ? 1 : 0) + ended;
  return r;
}
// gap j of the group: (E_{j-1}, S_j) in raw coordinates
__device__ __forceinline__ void subtract_gap(const SubtractView& v, const GapRange& r, i64 j, i64* cs, i64* ce) {
  *cs = j == 0 ? 0 : sub_raw(v.rec[r.c0 + j - 1].y, r.g0, v.L);
  *ce = j == r.C ? I64_MAX : sub_raw(v.rec[r.c0 + j].x, r.g0, v.L);
}

__global__ void __launch_bounds__(SC_THREADS)
    subtract_count_kernel(const i32* __restrict__ grp1, const i64* __restrict__ s1, const i64* __restrict__ e1, i64 n1,
                          SubtractView v, u32* __restrict__ jlo_out, u32* __restrict__ cnt_out,
                          u32* __restrict__ loc_out, u64* __restrict__ tile_tot, i64* __restrict__ overflow) {
  __shared__ u64 s_tmp[8];
  const i64 base = (i64)blockIdx.x * SC_TILE + (i64)threadIdx.x * SC_IPT;
  u64 mine = 0;
  // one row at a time (the probe code is not replicated SC_IPT times: registers, and with them occupancy,
  // decide the speed of these random probes); counts go straight to memory and are read back below
#pragma unroll 1
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 r = base + j;
    if (r >= n1) break;
    u32 cnt = 0, jl = 0;
    const i32 g = grp1 ? grp1[r] : 0;
    if (g >= 0 && g < v.nb) {
      const GapRange q = subtract_range(v, (u32)g, s1[r], e1[r]);
      i64 c = q.jhi - q.jlo;
      if (c > 0) {
        // only the outermost gaps of a group can be empty: before a merged interval starting at 0,
        // after one reaching INT64_MAX
        if (q.C > 0 && q.jlo == 0 && sub_raw(v.rec[q.c0].x, q.g0, v.L) <= 0) --c;
        if (q.C > 0 && q.jhi - 1 == q.C && sub_raw(v.rec[q.c0 + q.C - 1].y, q.g0, v.L) == I64_MAX) --c;
        if (c >> 31) {
          *overflow = 1;
          c = 0;
        }
        cnt = (u32)c;
        jl = (u32)q.jlo;
      }
    }
    jlo_out[r] = jl;
    cnt_out[r] = cnt;
    mine += cnt;
  }
  u64 tot;
  u64 run = block_excl_sum_u64(mine, s_tmp, &tot);
  if (threadIdx.x == 0) {
    tile_tot[blockIdx.x] = tot;
    if (tot >> 32) *overflow = 1;  // in-tile offsets are 32-bit
  }
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 r = base + j;
    if (r < n1) {
      loc_out[r] = (u32)run;
      run += cnt_out[r];  // written by this thread above
    }
  }
}

__global__ void __launch_bounds__(256)
    subtract_emit_kernel(const i32* __restrict__ grp1, const i64* __restrict__ s1, const i64* __restrict__ e1, i64 n1,
                         SubtractView v, const u32* __restrict__ jlo_in, const u32* __restrict__ cnt_in,
                         const u32* __restrict__ loc_in, const u64* __restrict__ spine, i64* __restrict__ row_out,
                         i64* __restrict__ start_out, i64* __restrict__ end_out, i64* __restrict__ sub_out) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n1) return;
  const u32 cnt = cnt_in[r];
  if (cnt == 0) return;
  const i64 s = s1[r], e = e1[r];
  const u32 g = grp1 ? (u32)grp1[r] : 0u;
  GapRange q;
  q.c0 = 0;
  q.C = 0;
  q.g0 = 0;
  if (v.cg_first) {
    q.c0 = v.cg_first[g];
    q.C = v.cg_first[g + 1] - q.c0;
    q.g0 = v.L.goff[g];
  }
  i64 at = (i64)spine[r / SC_TILE] + loc_in[r];
  u32 k = 0;
  for (i64 j = jlo_in[r]; k < cnt; ++j) {
    i64 cs, ce;
    subtract_gap(v, q, j, &cs, &ce);
    if (cs >= ce) continue;  // empty outermost gap
    if (row_out) row_out[at] = r;
    start_out[at] = s > cs ? s : cs;
    end_out[at] = e < ce ? e : ce;
    if (sub_out) sub_out[at] = k;
    ++at;
    ++k;
  }
}

static int subtract_count_impl(const i32* grp1, const i64* s1, const i64* e1, i64 n1, const i32* grp2, const i64* s2,
                               const i64* e2, i64 n2, i32 nb, void* ws_ptr, size_t ws_bytes, SubtractPlan* plan,
                               cudaStream_t st) {
  if (n1 < 0 || n2 < 0 || nb < 1 || !plan) {
    set_error("bf_subtract_count: bad argument (n1=%lld n2=%lld n_groups=%d)", (long long)n1, (long long)n2, nb);
    return BF_ERR_ARG;
  }
  if (n1 >= (1ll << 31) || n2 >= (1ll << 31)) {
    set_error("bf_subtract_count: more than 2^31-1 rows in one set");
    return BF_ERR_RANGE;
  }
  plan->magic = 0;
  plan->ready = 0;
  plan->n1 = n1;
  plan->n2 = n2;
  plan->nb = nb;
  plan->n_rows = 0;
  plan->grp1 = nb > 1 ? grp1 : nullptr;
  plan->s1 = s1;
  plan->e1 = e1;
  plan->view = SubtractView{};
  if (n1 == 0) {
    plan->magic = SUBTRACT_MAGIC;
    plan->ready = 1;
    return BF_OK;
  }
  if (!s1 || !e1 || (nb > 1 && !grp1) || (n2 > 0 && (!s2 || !e2 || (nb > 1 && !grp2)))) {
    set_error("bf_subtract_count: null pointer");
    return BF_ERR_ARG;
  }
  Arena measure(nullptr);
  SubtractWS w;
  layout_subtract(measure, n1, n2, nb, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_subtract_count: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_subtract(arena, n1, n2, nb, &w);
  BF_CUDA(cudaMemsetAsync(w.hdr, 0, 2 * sizeof(i64), st));
  SubtractView view = {};
  view.nb = nb;
  if (n2 > 0) {
    BF_LAUNCH(PC_GATHER, 40 * n2, st,
              subtract_clip_kernel<<<(unsigned)ceil_div(n2, 256), 256, 0, st>>>(nb > 1 ? grp2 : nullptr, s2, e2, n2, nb,
                                                                               w.g2c, w.s2c, w.e2c));
    ClusterTable tab;
    KeyLayout L;
    // no clamping (as in coverage): the pieces are built from the merged ends themselves
    BF_TRY(cluster_run(w.g2c, w.s2c, w.e2c, n2, nb + 1, 0, 1, w.cl, false, st, &tab, &L, 0));
    BF_LAUNCH(PC_COVERAGE, 36 * n2, st,
              subtract_table_kernel<<<(unsigned)ceil_div(n2, 256), 256, 0, st>>>(tab, L, w.rec));
    BF_LAUNCH(PC_TABLE, 0, st, subtract_groups_kernel<<<(unsigned)ceil_div(nb + 2, 256), 256, 0, st>>>(tab, nb, w.cg_first));
    const Lut lut = make_lut(w.lut, n2, 0, 1ull << L.B, 0, 2);
    BF_TRY(launch_lut_build(reinterpret_cast<const u64*>(w.rec), 0, tab.n_clusters, lut, w.lut, st, PC_COVERAGE));
    view.rec = w.rec;
    view.cg_first = w.cg_first;
    view.lut = lut;
    view.L = L;
  }
  const unsigned tiles = (unsigned)scan_tiles((u64)n1);
  BF_LAUNCH(PC_COVERAGE, 32 * n1, st,
            subtract_count_kernel<<<tiles, SC_THREADS, 0, st>>>(plan->grp1, s1, e1, n1, view, w.jlo, w.cnt, w.loc, w.spine,
                                                                w.hdr));
  BF_LAUNCH(PC_SCAN, 0, st,
            spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.spine, (u64)tiles, reinterpret_cast<u64*>(w.hdr + 1)));
  i64 host[2] = {0, 0};
  BF_CUDA(cudaMemcpyAsync(host, w.hdr, 2 * sizeof(i64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  if (host[0]) {
    set_error("bf_subtract_count: more than 2^31-1 pieces for one row or 2^32-1 for 2048 consecutive rows");
    return BF_ERR_RANGE;
  }
  plan->n_rows = host[1];
  plan->view = view;
  plan->magic = SUBTRACT_MAGIC;
  plan->ready = 1;
  return BF_OK;
}

static int subtract_emit_impl(const SubtractPlan* plan, void* ws_ptr, i64* row_out, i64* start_out, i64* end_out,
                              i64* sub_out, cudaStream_t st) {
  if (!plan || plan->magic != SUBTRACT_MAGIC || !plan->ready) {
    set_error("bf_subtract_emit: plan was not produced by a successful bf_subtract_count");
    return BF_ERR_STATE;
  }
  if (plan->n_rows == 0) return BF_OK;
  if (!start_out || !end_out) {
    set_error("bf_subtract_emit: null output pointer");
    return BF_ERR_ARG;
  }
  Arena arena(ws_ptr);
  SubtractWS w;
  layout_subtract(arena, plan->n1, plan->n2, plan->nb, &w);
  BF_LAUNCH(PC_EMIT, 20 * plan->n1 + 32 * plan->n_rows, st,
            subtract_emit_kernel<<<(unsigned)ceil_div(plan->n1, 256), 256, 0, st>>>(
                plan->grp1, plan->s1, plan->e1, plan->n1, plan->view, w.jlo, w.cnt, w.loc, w.spine, row_out, start_out,
                end_out, sub_out));
  return BF_OK;
}

}  // namespace bf
