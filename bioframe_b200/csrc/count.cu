// bf_count_overlaps: number of set-2 intervals overlapping every set-1 interval, in input row order.
//
// Replaces bioframe.count_overlaps (src/bioframe/ops.py:1371-1438: left overlap with keep_order, then a
// groupby-count over the pair list) and is what setdiff (ops.py:1333-1368) needs (count == 0).  No pair is
// enumerated: with the targets sorted once by linear start and once by linear end'
//     count(q) = #{t : s_t < e'_q}  -  #{t : e'_t <= s_q}
// (every target that ends at or before the query's start also starts before the query's end, so the
// difference is exactly #{t : s_t < e'_q and e'_t > s_q}, the overlap test of arrops.py:338-342).  On the
// linear (group, coordinate) axis the targets of earlier groups are counted in both terms and cancel.
// Two binary searches per query, any query order.
#include "common.cuh"
#include "keys.cuh"
#include "lut.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "sorted_set.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

struct CountWS {
  SortBufs by_start, by_end;
  RadixScratch radix;
  LayoutScratch layout;
  u32 *lutS, *lutE;  // direct-address tables over the two target orders
};
static void layout_count(Arena& a, i64 n2, i32 nb, CountWS* w) {
  take_sort_bufs(a, n2, &w->by_start);
  take_sort_bufs(a, n2, &w->by_end);
  w->radix = take_radix_scratch(a, (u64)n2);
  w->layout = take_layout_scratch(a, nb);
  w->lutS = a.take<u32>(lut_words(n2));
  w->lutE = a.take<u32>(lut_words(n2));
}

__global__ void __launch_bounds__(256)
    count_overlaps_kernel(const i32* __restrict__ grp1, const i64* __restrict__ s1, const i64* __restrict__ e1, i64 n1,
                          const u64* __restrict__ kS, const u64* __restrict__ kE, i64 n2, KeyLayout L, Lut lutS,
                          Lut lutE, int closed, i64* __restrict__ counts) {
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n1) return;
  const u32 g = grp1 ? (u32)grp1[r] : 0u;
  const i64 a = s1[r];
  i64 b = e1[r];
  if (a == b) b += 1;  // a point behaves as [x, x+1) (arrops.py:271-287)
  b = min(b, L.endcap);
  const u64 off = L.goff[g];
  const u64 xs = off + ((u64)a - L.cmin), xe = off + ((u64)b - L.cmin);
  // closed=True also counts intervals that only touch (arrops.py:339,342 with side="right")
  const i64 started = closed ? lut_count_below<true>(kS, lutS, xe) : lut_count_below<false>(kS, lutS, xe);
  const i64 ended = closed ? lut_count_below<false>(kE, lutE, xs) : lut_count_below<true>(kE, lutE, xs);
  counts[r] = started - ended;
}

static int count_overlaps_impl(const i32* grp1, const i64* s1, const i64* e1, i64 n1, const i32* grp2, const i64* s2,
                               const i64* e2, i64 n2, i32 nb, i32 closed, void* ws_ptr, size_t ws_bytes,
                               cudaStream_t st, i64* counts) {
  if (n1 < 0 || n2 < 0 || nb < 1) {
    set_error("bf_count_overlaps: bad argument (n1=%lld n2=%lld n_groups=%d)", (long long)n1, (long long)n2, nb);
    return BF_ERR_ARG;
  }
  if (n1 >= (1ll << 31) || n2 >= (1ll << 31)) {
    set_error("bf_count_overlaps: more than 2^31-1 rows in one set");
    return BF_ERR_RANGE;
  }
  if (n1 == 0) return BF_OK;
  if (!counts || !s1 || !e1 || (nb > 1 && !grp1) || (n2 > 0 && (!s2 || !e2 || (nb > 1 && !grp2)))) {
    set_error("bf_count_overlaps: null pointer");
    return BF_ERR_ARG;
  }
  if (n2 == 0) {
    BF_CUDA(cudaMemsetAsync(counts, 0, (size_t)n1 * sizeof(i64), st));
    return BF_OK;
  }
  Arena measure(nullptr);
  CountWS w;
  layout_count(measure, n2, nb, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_count_overlaps: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_count(arena, n2, nb, &w);
  const RowSet sets[2] = {{grp1, s1, e1, n1}, {grp2, s2, e2, n2}};
  KeyLayout L;
  BF_TRY(measure_layout(sets, 2, nb, 1, 1, w.layout, st, &L, "bf_count_overlaps"));
  SortResult rS, rE;
  BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_POINT_ADJUST | KEY_NO_TIE_FIX | presorted_mode(L, 1), w.by_start, w.radix, st, &rS));
  BF_TRY(sort_rows(nb > 1 ? grp2 : nullptr, s2, e2, n2, L, KEY_POINT_ADJUST | KEY_NO_TIE_FIX | KEY_BY_END, w.by_end,
                   w.radix, st, &rE));
  const Lut lutS = make_lut(w.lutS, n2, 0, 1ull << L.B, L.LB), lutE = make_lut(w.lutE, n2, 0, 1ull << L.B, L.LB);
  BF_TRY(launch_lut_build(rS.keys, n2, nullptr, lutS, w.lutS, st, PC_COVERAGE));
  BF_TRY(launch_lut_build(rE.keys, n2, nullptr, lutE, w.lutE, st, PC_COVERAGE));
  BF_LAUNCH(PC_COVERAGE, 28 * n1, st,
            count_overlaps_kernel<<<(unsigned)ceil_div(n1, 256), 256, 0, st>>>(nb > 1 ? grp1 : nullptr, s1, e1, n1, rS.keys,
                                                                              rE.keys, n2, L, lutS, lutE, closed, counts));
  return BF_OK;
}

}  // namespace bf

extern "C" {
size_t bf_count_overlaps_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups) {
  (void)n1;
  if (n2 < 0 || n_groups < 1) return 0;
  bf::Arena a(nullptr);
  bf::CountWS w;
  bf::layout_count(a, n2, n_groups, &w);
  return a.off;
}
int bf_count_overlaps(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1, const int32_t* grp2,
                      const int64_t* start2, const int64_t* end2, int64_t n2, int32_t n_groups, int32_t closed,
                      void* ws, size_t ws_bytes, void* stream, int64_t* counts_out) {
  return bf::count_overlaps_impl((const i32*)grp1, (const i64*)start1, (const i64*)end1, n1, (const i32*)grp2,
                                 (const i64*)start2, (const i64*)end2, n2, n_groups, closed, ws, ws_bytes,
                                 (cudaStream_t)stream, (i64*)counts_out);
}
}
