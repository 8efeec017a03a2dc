// Sort-key construction shared by overlap / cluster / closest / coverage.
//
// Every row becomes ONE 64-bit key
//     key = ( group_rank << B | (start - cmin) ) << LB  |  (end' - start)
// with B = bit width of the coordinate span of the whole call, LB = bit width
// of the longest interval, and end' the point-adjusted end (end + (end ==
// start), arrops.py:271-287) for overlap or the raw end for merge.
//
// The reference orders each chromosome with np.lexsort([ends, starts])
// (arrops.py:332-333, 459): by start, ties by end, ties by row.  Here ONE stable
// LSD radix sort over the HIGH field only -- cs = (group, start), 5 + 28 = 33
// bits for hg38 -- orders all chromosomes at once; rows that share (group,
// start) (a few percent of real data) are then put in (end', row) order by a
// fix-up that extracts just those rows, sorts them by the full key and writes
// them back into the same slots (tie_fix in overlap.cu).  The length bits ride
// along in the key, so after the sort the key alone yields
//     cs(key) = key >> LB                 (group, start)
//     ce(key) = cs(key) + (key & lmask)   (group, end')
// and the searches and running maxima never touch the input columns again.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace bf {

struct KeyLayout {
  int B;          // bits of the coordinate field
  int LB;         // bits of the length field
  int rank_bits;  // bits of the group rank
  int total_bits; // rank_bits + B + LB
  u64 cmin;       // coordinate origin (two's complement image of the int64 minimum)
  u64 lmask;      // (1 << LB) - 1
};

__host__ __device__ __forceinline__ u64 key_cs(u64 k, const KeyLayout& L) { return k >> L.LB; }
__host__ __device__ __forceinline__ u64 key_ce(u64 k, const KeyLayout& L) { return (k >> L.LB) + (k & L.lmask); }
__host__ __device__ __forceinline__ u32 key_rank(u64 k, const KeyLayout& L) {
  return L.rank_bits ? (u32)(k >> (L.LB + L.B)) : 0u;
}

struct Extent {  // device-side reduction target
  i64 cmin, cmax;
  i64 lmin, lmax;  // min / max of end' - start
  i32 gmin, gmax;
};

static __global__ void extent_init_kernel(Extent* x) {
  x->cmin = 0x7fffffffffffffffLL;
  x->cmax = (i64)0x8000000000000000ULL;
  x->lmin = 0x7fffffffffffffffLL;
  x->lmax = (i64)0x8000000000000000ULL;
  x->gmin = 0x7fffffff;
  x->gmax = (i32)0x80000000;
}

// min/max of coordinates (end' when point_adjust), of lengths and of group codes
static __global__ void __launch_bounds__(256) extent_kernel(const i32* __restrict__ grp, const i64* __restrict__ s,
                                                     const i64* __restrict__ e, u64 n, int point_adjust,
                                                     Extent* __restrict__ out) {
  i64 lo = 0x7fffffffffffffffLL, hi = (i64)0x8000000000000000ULL;
  i64 llo = 0x7fffffffffffffffLL, lhi = (i64)0x8000000000000000ULL;
  i32 glo = 0x7fffffff, ghi = (i32)0x80000000;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    i64 a = s[i], b = e[i];
    if (point_adjust && a == b) b += 1;
    lo = min(lo, min(a, b));
    hi = max(hi, max(a, b));
    // length in wrapping arithmetic; a span that overflows int64 shows up as negative
    const i64 len = (i64)((u64)b - (u64)a);
    llo = min(llo, len);
    lhi = max(lhi, len);
    if (grp) {
      i32 g = grp[i];
      glo = min(glo, g);
      ghi = max(ghi, g);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    llo = min(llo, __shfl_xor_sync(0xffffffffu, llo, o));
    lhi = max(lhi, __shfl_xor_sync(0xffffffffu, lhi, o));
    glo = min(glo, __shfl_xor_sync(0xffffffffu, glo, o));
    ghi = max(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
  }
  if (lane_id() == 0) {
    atomicMin(&out->cmin, lo);
    atomicMax(&out->cmax, hi);
    atomicMin(&out->lmin, llo);
    atomicMax(&out->lmax, lhi);
    if (grp) {
      atomicMin(&out->gmin, glo);
      atomicMax(&out->gmax, ghi);
    }
  }
}

static inline int make_key_layout(const Extent& x, i32 n_groups, KeyLayout* L) {
  if (x.lmin < 0) {
    set_error("interval with end < start (or a span beyond int64): not a valid bedframe interval");
    return BF_ERR_ARG;
  }
  u64 span = (x.cmax >= x.cmin) ? (u64)x.cmax - (u64)x.cmin : 0;
  L->B = bit_width(span);
  if (L->B == 0) L->B = 1;
  L->LB = bit_width((u64)x.lmax);
  if (L->LB == 0) L->LB = 1;
  L->rank_bits = n_groups > 1 ? bit_width((u64)(n_groups - 1)) : 0;
  L->total_bits = L->rank_bits + L->B + L->LB;
  L->cmin = (u64)x.cmin;
  L->lmask = (1ull << L->LB) - 1ull;
  if (L->total_bits > 64 || L->LB > 62) {
    set_error("coordinates need %d group + %d start + %d length bits = %d > 64-bit key "
              "(wide-coordinate path not available for this call)",
              L->rank_bits, L->B, L->LB, L->total_bits);
    return BF_ERR_RANGE;
  }
  return BF_OK;
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
  u32* wtab = hist_tabs + (threadIdx.x >> 5) * rp.passes * RS_BINS;
  const u64 per_block = (u64)HIST_THREADS * 8ull;
  for (u64 base = (u64)blockIdx.x * per_block; base < n; base += (u64)gridDim.x * per_block) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      u64 i = base + (u64)j * HIST_THREADS + threadIdx.x;
      bool ok = i < n;
      u64 k = 0;
      if (ok) {
        i64 a = s[i], b = e[i];
        if (point_adjust && a == b) b += 1;
        u64 g = grp ? (u64)(u32)grp[i] : 0ull;
        k = ((((L.rank_bits ? (g << L.B) : 0ull) | ((u64)a - L.cmin))) << L.LB) | ((u64)b - (u64)a);
        keys[i] = k;
      }
      hist_add(wtab, rp, k, ok);
    }
  }
  __syncthreads();
  hist_flush(hist_tabs, rp, ghist);
}

static int launch_pack_keys(const i32* grp, const i64* s, const i64* e, u64 n, const KeyLayout& L, int point_adjust,
                            const RadixPlan& rp, u64* keys, u64* ghist, cudaStream_t st);

static inline unsigned stream_grid(u64 n, u64 per_block) {
  u64 g = (n + per_block - 1) / per_block;
  const u64 cap = 148ull * 8ull;  // a few resident blocks per SM, grid-stride beyond that
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
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
