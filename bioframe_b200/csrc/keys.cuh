// Sort-key construction shared by overlap / cluster / closest / coverage.
//
// Every row becomes ONE 64-bit key
//     key = ( goff[group] + (start - cmin) ) << LB  |  (end' - start)
// where goff[g] is the exclusive prefix sum of the coordinate spans of the
// groups before g (span_g = max coordinate seen in group g - cmin + 1, over all
// sets of the call), i.e. the groups are laid end to end on one linear axis
// -- for hg38 that axis is 3.09e9 long and fits 32 bits, one radix pass less
// than a (group rank, start) bit field.  LB = bit width of the longest
// interval; end' is the point-adjusted end (end + (end == start),
// arrops.py:271-287) for overlap or the raw end for merge/closest.
//
// The reference orders each chromosome with np.lexsort([ends, starts])
// (arrops.py:332-333, 459): by start, ties by end, ties by row.  Here ONE stable
// LSD radix sort over the HIGH field only -- cs = linear (group, start) --
// orders all chromosomes at once; rows that share (group, start) (a few
// percent of real data) are then put in (end', row) order by a fix-up
// (sorted_set.cuh).  The length bits ride along in the key, so after the sort
// the key alone yields
//     cs(key) = key >> LB                 linear (group, start)
//     ce(key) = cs(key) + (key & lmask)   linear (group, end')
// and the searches and running maxima never touch the input columns again.
// ce of a row never reaches goff[g+1], so rows of different groups never
// interact in any comparison of cs / ce values.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace bf {

struct KeyLayout {
  int B;          // bits of the linear coordinate field
  int LB;         // bits of the length field
  int rank_bits;  // bits of a group rank (for sorting lists keyed by rank)
  int total_bits; // B + LB
  i32 nb;         // number of groups
  u64 cmin;       // coordinate origin (two's complement image of the int64 minimum)
  u64 lmask;      // (1 << LB) - 1
  i64 endcap;     // ends are clamped to this value before packing (I64_MAX: no clamp), see KEY_CLAMP_END
  u32 presorted;  // bit k: set k of the call already is in (group rank, start) order (no radix passes needed)
  const u64* goff;  // device [nb+1]: linear offset of every group, goff[nb] = total span
};

__host__ __device__ __forceinline__ u64 key_cs(u64 k, const KeyLayout& L) { return k >> L.LB; }
__host__ __device__ __forceinline__ u64 key_ce(u64 k, const KeyLayout& L) { return (k >> L.LB) + (k & L.lmask); }
// group of a linear coordinate: the last g with goff[g] <= cs (empty groups have zero span)
__device__ __forceinline__ u32 rank_of_cs(u64 cs, const KeyLayout& L) {
  if (L.nb <= 1) return 0u;
  i32 lo = 1, hi = L.nb;  // count g in [1, nb) with goff[g] <= cs
  while (lo < hi) {
    const i32 mid = (lo + hi) >> 1;
    if (L.goff[mid] <= cs)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (u32)(lo - 1);
}
__device__ __forceinline__ u32 key_rank(u64 k, const KeyLayout& L) { return rank_of_cs(key_cs(k, L), L); }

struct Extent {  // device-side reduction target
  i64 cmin, cmax;
  i64 smax;        // largest start
  u64 lmax;        // largest end' - start (exact in unsigned arithmetic for any int64 pair with end >= start)
  i32 gmin, gmax;
  i32 bad;         // some row has end < start
  i32 unsorted;    // bit k: set k is NOT in (group code, start) order
  i64 endcap;      // in: clamp for the ends when measuring spans (I64_MAX: none); set by the host wrapper
  u64 total;       // sum of the group spans (written by layout_offsets_kernel)
};
struct LayoutScratch {
  Extent* extent;
  i64* gmax;  // [nb] max coordinate per group
  u64* goff;  // [nb+1]
};
static inline LayoutScratch take_layout_scratch(Arena& a, i32 nb) {
  LayoutScratch s;
  s.extent = a.take<Extent>(1);
  s.gmax = a.take<i64>(nb + 1);
  s.goff = a.take<u64>(nb + 2);
  return s;
}
constexpr i64 I64_MIN = (i64)0x8000000000000000ULL;
constexpr i64 I64_MAX = 0x7fffffffffffffffLL;
constexpr int EXTENT_SMEM_GROUPS = 2048;  // per-group maxima staged in shared memory up to this many groups

static __global__ void extent_init_kernel(Extent* x, i64* gmax, i32 nb) {
  const i32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    x->cmin = I64_MAX;
    x->cmax = I64_MIN;
    x->smax = I64_MIN;
    x->lmax = 0;
    x->bad = 0;
    x->unsorted = 0;
    x->gmin = 0x7fffffff;
    x->gmax = (i32)0x80000000;
    x->total = 0;
  }
  if (i < nb) gmax[i] = I64_MIN;
}

// min/max of coordinates (end' when point_adjust), of lengths and of group
// codes, plus the max coordinate of every group.  The per-group table is read
// with a plain load and only touched with an atomic when a row raises it, which
// after the first few rows of a block almost never happens.
static __global__ void __launch_bounds__(256) extent_kernel(const i32* __restrict__ grp, const i64* __restrict__ s,
                                                            const i64* __restrict__ e, u64 n, int point_adjust, i32 nb,
                                                            int set_bit, Extent* __restrict__ out,
                                                            i64* __restrict__ gmax_out) {
  extern __shared__ i64 s_gmax[];
  const bool staged = nb <= EXTENT_SMEM_GROUPS;
  volatile i64* table = staged ? s_gmax : gmax_out;
  if (staged) {
    for (i32 i = threadIdx.x; i < nb; i += blockDim.x) s_gmax[i] = I64_MIN;
    __syncthreads();
  }
  i64 lo = I64_MAX, hi = I64_MIN;
  i64 shi = I64_MIN;
  u64 lhi = 0;
  bool bad = false;
  i32 glo = 0x7fffffff, ghi = (i32)0x80000000;
  bool disorder = false;
  // row i is out of order if (group, start) of row i-1 is larger (sorted input -- BED files usually are --
  // skips the radix passes altogether); the neighbour's line is in L1 from the previous lane's load
  auto check_order = [&](u64 i, i64 a, i32 g) {
    if (i == 0) return;
    const i64 pa = s[i - 1];
    const i32 pg = grp ? grp[i - 1] : 0;
    disorder |= pg > g || (pg == g && pa > a);
  };
  auto take = [&](i64 a, i64 b, i32 g) {
    if (point_adjust && a == b) b += 1;
    const i64 top = max(a, b);
    lo = min(lo, min(a, b));
    hi = max(hi, top);
    shi = max(shi, a);
    bad |= b < a;
    const u64 len = (u64)b - (u64)a;  // exact for b >= a, whatever the magnitudes
    lhi = b < a ? lhi : max(lhi, len);
    if (grp) {
      glo = min(glo, g);
      ghi = max(ghi, g);
    }
    if ((u32)g < (u32)nb && table[g] < top) atomicMax((i64*)&table[g], top);
  };
  // four rows per thread and iteration, all twelve loads issued before the first use
  const u64 stride = (u64)gridDim.x * blockDim.x;
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    i64 a[4], b[4];
    i32 g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = s[i + u * stride];
      b[u] = e[i + u * stride];
      g[u] = grp ? grp[i + u * stride] : 0;
    }
    // order check on one row in four only: random input fails it at once, and a set that passes is
    // checked in full by order_check_kernel (measure_layout) -- so unsorted input pays almost nothing
    check_order(i, a[0], g[0]);
#pragma unroll
    for (int u = 0; u < 4; ++u) take(a[u], b[u], g[u]);
  }
  for (; i < n; i += stride) {
    const i64 a = s[i];
    const i32 g = grp ? grp[i] : 0;
    check_order(i, a, g);
    take(a, e[i], g);
  }
  if (__any_sync(0xffffffffu, disorder) && lane_id() == 0) atomicOr(&out->unsorted, set_bit);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    shi = max(shi, __shfl_xor_sync(0xffffffffu, shi, o));
    lhi = max(lhi, __shfl_xor_sync(0xffffffffu, lhi, o));
    glo = min(glo, __shfl_xor_sync(0xffffffffu, glo, o));
    ghi = max(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
  }
  if (__any_sync(0xffffffffu, bad) && lane_id() == 0) out->bad = 1;
  if (lane_id() == 0) {
    atomicMin(&out->cmin, lo);
    atomicMax(&out->cmax, hi);
    atomicMax(&out->smax, shi);
    atomicMax(&out->lmax, lhi);
    if (grp) {
      atomicMin(&out->gmin, glo);
      atomicMax(&out->gmax, ghi);
    }
  }
  if (staged) {
    __syncthreads();
    for (i32 i = threadIdx.x; i < nb; i += blockDim.x)
      if (s_gmax[i] != I64_MIN) atomicMax(&gmax_out[i], s_gmax[i]);
  }
}

// exact order check of one set (only run for sets whose sampled check in extent_kernel found no disorder)
static __global__ void __launch_bounds__(256) order_check_kernel(const i32* __restrict__ grp, const i64* __restrict__ s,
                                                                 u64 n, int set_bit, Extent* __restrict__ out) {
  bool disorder = false;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (u64)gridDim.x * blockDim.x) {
    const i32 g = grp ? grp[i] : 0, pg = grp ? grp[i - 1] : 0;
    disorder |= pg > g || (pg == g && s[i - 1] > s[i]);
  }
  if (__any_sync(0xffffffffu, disorder) && lane_id() == 0) atomicOr(&out->unsorted, set_bit);
}

// goff[g] = sum of spans of the groups before g; one block, sequential over a
// few dozen (at most a few thousand) groups per thread chunk.
static __global__ void __launch_bounds__(1024) layout_offsets_kernel(Extent* __restrict__ x, const i64* __restrict__ gmax,
                                                                     i32 nb, int clamp_ends, u64* __restrict__ goff) {
  __shared__ u64 s_part[1024];
  const u32 tid = threadIdx.x;
  const u64 cmin = (u64)x->cmin;
  // with clamped ends (overlap / closest) no coordinate above max start + 1 is ever packed
  const i64 cap = (clamp_ends && x->smax < I64_MAX) ? x->smax + 1 : I64_MAX;
  const i32 per = (nb + 1023) / 1024;
  const i32 lo = tid * per, hi = min(nb, lo + per);
  auto span = [&](i32 g) -> u64 { return gmax[g] == I64_MIN ? 0ull : (u64)min(gmax[g], cap) - cmin + 1ull; };
  u64 acc = 0;
  for (i32 g = lo; g < hi; ++g) acc += span(g);
  s_part[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    u64 run = 0;
    for (int i = 0; i < 1024; ++i) {
      const u64 v = s_part[i];
      s_part[i] = run;
      run += v;
    }
    goff[nb] = run;
    x->total = run;
    x->endcap = cap;
  }
  __syncthreads();
  u64 run = s_part[tid];
  for (i32 g = lo; g < hi; ++g) {
    goff[g] = run;
    run += span(g);
  }
}

static inline int make_key_layout(const Extent& x, i32 n_groups, const u64* goff, KeyLayout* L) {
  if (x.bad) {
    set_error("interval with end < start: not a valid bedframe interval");
    return BF_ERR_ARG;
  }
  // total span = sum over groups of (group max - cmin + 1); 0 when there are no rows
  L->B = bit_width(x.total > 0 ? x.total - 1 : 0);
  if (L->B == 0) L->B = 1;
  // longest packed length: bounded by the longest interval and, with clamped ends, by the clamped span
  u64 lmax = x.lmax;
  if (x.endcap != I64_MAX && (u64)x.endcap - (u64)x.cmin < lmax) lmax = (u64)x.endcap - (u64)x.cmin;
  L->LB = bit_width(lmax);
  if (L->LB == 0) L->LB = 1;
  L->endcap = x.endcap;
  L->rank_bits = n_groups > 1 ? bit_width((u64)(n_groups - 1)) : 0;
  L->total_bits = L->B + L->LB;
  L->nb = n_groups;
  L->cmin = (u64)x.cmin;
  L->lmask = (1ull << L->LB) - 1ull;
  L->goff = goff;
  // a group span that wrapped (coordinates spread over more than 2^63) shows up as a huge total
  if (L->total_bits > 64 || L->LB > 62 || x.total > (1ull << 63)) {
    set_error("coordinates need %d position + %d length bits = %d > 64-bit key "
              "(wide-coordinate path not available for this call)",
              L->B, L->LB, L->total_bits);
    return BF_ERR_RANGE;
  }
  return BF_OK;
}

struct RowSet {  // one input interval set of a call
  const i32* grp;
  const i64* s;
  const i64* e;
  i64 n;
};
static inline unsigned stream_grid(u64 n, u64 per_block) {
  u64 g = (n + per_block - 1) / per_block;
  const u64 cap = 148ull * 8ull;  // a few resident blocks per SM, grid-stride beyond that
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

// Coordinate extent + per-group spans of all sets of a call -> key layout.
// One small D2H copy and one stream synchronisation.
// clamp_ends (allowed for overlap and closest): when the plain layout does not fit 64 key bits, ends
// above (largest start + 1) are packed as that value.  An end is only ever compared with starts, so all
// ends beyond the largest start behave alike; this keeps the INT64_MAX ends that complement / subtract /
// trim inject (ops.py:1309,1499,1604) packable.  The one thing the clamp can change is the (end', row)
// order among rows that share a start AND both reach beyond the largest start; the tie fix-up compares
// the original ends of such rows (sorted_set.cuh).
static int measure_layout(const RowSet* sets, int n_sets, i32 nb, int point_adjust, int clamp_ends,
                          const LayoutScratch& sc, cudaStream_t st, KeyLayout* L, const char* who) {
  BF_LAUNCH(PC_EXTENT, 0, st, extent_init_kernel<<<(unsigned)ceil_div(nb + 1, 256), 256, 0, st>>>(sc.extent, sc.gmax, nb));
  const size_t smem = nb <= EXTENT_SMEM_GROUPS ? (size_t)nb * sizeof(i64) : 0;
  i64 total_rows = 0;
  for (int k = 0; k < n_sets; ++k) {
    const RowSet& r = sets[k];
    total_rows += r.n;
    if (r.n <= 0) continue;
    BF_LAUNCH(PC_EXTENT, 20 * r.n, st,
              extent_kernel<<<stream_grid((u64)r.n, 1024), 256, smem, st>>>(nb > 1 ? r.grp : nullptr, r.s, r.e, (u64)r.n,
                                                                           point_adjust, nb, 1 << k, sc.extent, sc.gmax));
  }
  // first without clamping (the packed lengths then order tied starts exactly as the reference does);
  // only coordinates that do not fit 64 key bits otherwise -- INT64_MAX ends -- get the clamp
  u32 presorted = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    BF_LAUNCH(PC_EXTENT, 0, st, layout_offsets_kernel<<<1, 1024, 0, st>>>(sc.extent, sc.gmax, nb, attempt, sc.goff));
    Extent ext;
    BF_CUDA(cudaMemcpyAsync(&ext, sc.extent, sizeof(Extent), cudaMemcpyDeviceToHost, st));
    BF_CUDA(cudaStreamSynchronize(st));
    if (total_rows == 0) {
      ext.cmin = ext.cmax = 0;
      ext.smax = 0;
      ext.lmax = 0;
      ext.bad = 0;
      ext.endcap = I64_MAX;
      ext.total = 0;
    }
    if (nb > 1 && total_rows > 0 && (ext.gmin < 0 || ext.gmax >= nb)) {
      set_error("%s: group code outside [0,%d): min %d max %d", who, nb, ext.gmin, ext.gmax);
      return BF_ERR_GROUP;
    }
    if (attempt == 0) {
      // sets that passed the sampled order check get the exact one (12 B/row, only for sorted-looking input)
      const u32 maybe = ~(u32)ext.unsorted & ((1u << n_sets) - 1u);
      bool asked = false;
      for (int k = 0; k < n_sets; ++k) {
        if (!((maybe >> k) & 1u) || sets[k].n < 2) continue;
        BF_LAUNCH(PC_EXTENT, 12 * sets[k].n, st,
                  order_check_kernel<<<stream_grid((u64)sets[k].n, 2048), 256, 0, st>>>(
                      nb > 1 ? sets[k].grp : nullptr, sets[k].s, (u64)sets[k].n, 1 << k, sc.extent));
        asked = true;
      }
      if (asked) {
        BF_CUDA(cudaMemcpyAsync(&ext, sc.extent, sizeof(Extent), cudaMemcpyDeviceToHost, st));
        BF_CUDA(cudaStreamSynchronize(st));
      }
      presorted = ~(u32)ext.unsorted & ((1u << n_sets) - 1u);
    }
    const int rc = make_key_layout(ext, nb, sc.goff, L);
    L->presorted = presorted;
    if (rc != BF_ERR_RANGE || !clamp_ends || attempt == 1) return rc;
  }
  return BF_ERR_RANGE;
}

// ---- extent as a host-side record that merges across sets and ranks ----
// Sets that are sorted apart from each other -- the targets on rank 0, the query shard of every rank
//  -- must still be packed with ONE key layout.  The extent of a set is therefore also
// available as a small host array of int64 words that merges by elementwise MAX (np.maximum on the
// host, all_reduce(MAX) across ranks):
//   w[0] = ~cmin   w[1] = cmax   w[2] = largest start   w[3] = largest end' - start   w[4..7] = 0
//   w[8 + g] = largest coordinate of group g (INT64_MIN: the group has no rows)
constexpr int EXT_HDR = 8;
static inline void extent_to_words(const Extent& x, const i64* gmax_host, i32 nb, bool empty, i64* w) {
  for (int i = 0; i < EXT_HDR; ++i) w[i] = 0;
  w[0] = empty ? I64_MIN : ~x.cmin;
  w[1] = empty ? I64_MIN : x.cmax;
  w[2] = empty ? I64_MIN : x.smax;
  w[3] = empty ? 0 : (x.lmax > (u64)I64_MAX ? I64_MAX : (i64)x.lmax);
  for (i32 g = 0; g < nb; ++g) w[EXT_HDR + g] = empty ? I64_MIN : gmax_host[g];
}
// Layout from merged words; the same arithmetic as layout_offsets_kernel + make_key_layout.  goff_host gets
// nb + 1 entries (the caller uploads them and sets L->goff).
static inline int layout_from_words(const i64* w, i32 nb, int clamp_ends, KeyLayout* L, u64* goff_host) {
  Extent x;
  const bool empty = w[1] == I64_MIN;
  x.cmin = empty ? 0 : ~w[0];
  x.cmax = empty ? 0 : w[1];
  x.smax = empty ? 0 : w[2];
  x.lmax = (u64)w[3];
  x.bad = 0;
  x.unsorted = 0;
  x.gmin = 0;
  x.gmax = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const i64 cap = (attempt && !empty && x.smax < I64_MAX) ? x.smax + 1 : I64_MAX;
    u64 run = 0;
    for (i32 g = 0; g < nb; ++g) {
      goff_host[g] = run;
      const i64 gm = w[EXT_HDR + g];
      if (!empty && gm != I64_MIN) run += (u64)(gm < cap ? gm : cap) - (u64)x.cmin + 1ull;
    }
    goff_host[nb] = run;
    x.total = run;
    x.endcap = cap;
    const int rc = make_key_layout(x, nb, nullptr, L);
    L->presorted = 0;
    if (rc != BF_ERR_RANGE || !clamp_ends || attempt == 1) return rc;
  }
  return BF_ERR_RANGE;
}

// key modes (the `point_adjust` argument of pack_keys / sort_rows is a bit set)
constexpr int KEY_POINT_ADJUST = 1;  // end' = end + (end == start)
constexpr int KEY_BY_END = 2;        // high field = linear END coordinate (closest: left neighbours)
constexpr int KEY_NO_TIE_FIX = 4;    // keep rows that share the high field in row order
constexpr int KEY_PRESORTED = 8;     // the rows already are in (group, start) order: no radix passes
__host__ __device__ __forceinline__ int presorted_mode(const KeyLayout& L, int set_index) {
  return (L.presorted >> set_index) & 1u ? KEY_PRESORTED : 0;
}

// Build keys and, in the same pass over the inputs, the histograms of every
// radix digit place (so the sort never re-reads the keys to count them).
static __global__ void __launch_bounds__(HIST_THREADS) pack_keys_kernel(const i32* __restrict__ grp, const i64* __restrict__ s,
                                                                 const i64* __restrict__ e, u64 n, KeyLayout L,
                                                                 int point_adjust, RadixPlan rp,
                                                                 u64* __restrict__ keys, u64* __restrict__ ghist) {
  extern __shared__ u32 hist_tabs[];
  hist_clear(hist_tabs, rp);
  __syncthreads();
  u32* wtab = hist_tabs;
  const u64 per_block = (u64)HIST_THREADS * 8ull;
  for (u64 base = (u64)blockIdx.x * per_block; base < n; base += (u64)gridDim.x * per_block) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      u64 i = base + (u64)j * HIST_THREADS + threadIdx.x;
      bool ok = i < n;
      u64 k = 0;
      if (ok) {
        i64 a = s[i], b = e[i];
        if ((point_adjust & KEY_POINT_ADJUST) && a == b) b += 1;
        b = min(b, L.endcap);
        const u64 off = grp ? L.goff[grp[i]] : 0ull;
        const u64 at = (point_adjust & KEY_BY_END) ? (u64)b : (u64)a;  // field the rows are ordered by
        k = ((off + (at - L.cmin)) << L.LB) | ((u64)b - (u64)a);
        keys[i] = k;
      }
      hist_add(wtab, rp, k, ok);
    }
  }
  __syncthreads();
  hist_flush(hist_tabs, rp, ghist);
}

static int launch_pack_keys(const i32* grp, const i64* s, const i64* e, u64 n, const KeyLayout& L, int point_adjust,
                            const RadixPlan& rp, u64* keys, u64* ghist, cudaStream_t st) {
  const size_t smem = hist_smem_bytes(rp);
  BF_CUDA(cudaFuncSetAttribute(pack_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BF_LAUNCH(PC_PACK, 28 * n, st,
            pack_keys_kernel<<<stream_grid(n, 2048), HIST_THREADS, smem, st>>>(grp, s, e, n, L, point_adjust, rp, keys, ghist));
  return BF_OK;
}
static int launch_radix_hist(const u64* keys, u64 n_max, const u64* n_dev, const RadixPlan& rp, u64* ghist,
                             cudaStream_t st, int prof_cls = PC_SORT_MISC) {
  const size_t smem = hist_smem_bytes(rp);
  BF_CUDA(cudaFuncSetAttribute(radix_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BF_LAUNCH(prof_cls, 0, st,
            radix_hist_kernel<<<stream_grid(n_max, 4096), HIST_THREADS, smem, st>>>(keys, n_max, n_dev, rp, ghist));
  return BF_OK;
}

}  // namespace bf
