// Sort-key construction shared by overlap / cluster / closest / coverage.
//
// Every row becomes ONE 64-bit key
//     key = group_rank << (2B) | (start - cmin) << B | (end' - cmin)
// where B = bit width of the coordinate span of the whole call and end' is the
// point-adjusted end (end + (end == start), arrops.py:271-287) for overlap or
// the raw end for merge.  One stable LSD radix sort of these keys with the row
// number as payload gives, for all chromosomes at once, exactly the order
// np.lexsort([ends, starts]) gives per chromosome (arrops.py:332-333, 459) with
// the chromosomes laid out in group-rank order.  For hg38 coordinates and 24
// chromosomes that is 5 + 28 + 28 = 61 bits.  After the sort
// the key alone yields the composite search values
//     cs(key) = key >> B                        (group, start)
//     ce(key) = (key >> 2B) << B | key & mask   (group, end')
// so the binary searches and running maxima never touch the input columns again.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace bf {

struct KeyLayout {
  int B;          // bits per coordinate field
  int rank_bits;  // bits of the group rank
  int total_bits; // rank_bits + 2B
  u64 cmin;       // coordinate origin (two's complement image of the int64 minimum)
  u64 fmask;      // (1 << B) - 1
};

__host__ __device__ __forceinline__ u64 key_cs(u64 k, const KeyLayout& L) { return k >> L.B; }
__host__ __device__ __forceinline__ u64 key_ce(u64 k, const KeyLayout& L) {
  return ((k >> (2 * L.B)) << L.B) | (k & L.fmask);
}
__host__ __device__ __forceinline__ u32 key_rank(u64 k, const KeyLayout& L) { return (u32)(k >> (2 * L.B)); }
__host__ __device__ __forceinline__ u64 key_soff(u64 k, const KeyLayout& L) { return (k >> L.B) & L.fmask; }
__host__ __device__ __forceinline__ u64 key_eoff(u64 k, const KeyLayout& L) { return k & L.fmask; }

struct Extent {  // device-side reduction target
  i64 cmin, cmax;
  i32 gmin, gmax;
};

__global__ void extent_init_kernel(Extent* x) {
  x->cmin = 0x7fffffffffffffffLL;
  x->cmax = (i64)0x8000000000000000ULL;
  x->gmin = 0x7fffffff;
  x->gmax = (i32)0x80000000;
}

// min/max of coordinates (end' when point_adjust) and of group codes
__global__ void __launch_bounds__(256) extent_kernel(const i32* __restrict__ grp, const i64* __restrict__ s,
                                                     const i64* __restrict__ e, u64 n, int point_adjust,
                                                     Extent* __restrict__ out) {
  i64 lo = 0x7fffffffffffffffLL, hi = (i64)0x8000000000000000ULL;
  i32 glo = 0x7fffffff, ghi = (i32)0x80000000;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    i64 a = s[i], b = e[i];
    if (point_adjust && a == b) b += 1;
    lo = min(lo, min(a, b));
    hi = max(hi, max(a, b));
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
    glo = min(glo, __shfl_xor_sync(0xffffffffu, glo, o));
    ghi = max(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
  }
  if (lane_id() == 0) {
    atomicMin(&out->cmin, lo);
    atomicMax(&out->cmax, hi);
    if (grp) {
      atomicMin(&out->gmin, glo);
      atomicMax(&out->gmax, ghi);
    }
  }
}

static inline int make_key_layout(const Extent& x, i32 n_groups, KeyLayout* L) {
  u64 span = (x.cmax >= x.cmin) ? (u64)x.cmax - (u64)x.cmin : 0;
  L->B = bit_width(span);
  if (L->B == 0) L->B = 1;
  L->rank_bits = n_groups > 1 ? bit_width((u64)(n_groups - 1)) : 0;
  L->total_bits = L->rank_bits + 2 * L->B;
  L->cmin = (u64)x.cmin;
  L->fmask = L->B >= 64 ? ~0ull : ((1ull << L->B) - 1ull);
  if (L->total_bits > 64 || L->B > 31) {
    set_error("coordinate span needs %d bits per field and %d group bits: %d > 64-bit key "
              "(wide-coordinate path not available for this call)",
              L->B, L->rank_bits, L->total_bits);
    return BF_ERR_RANGE;
  }
  return BF_OK;
}

// Build keys and, in the same pass over the inputs, the histograms of every
// radix digit place (so the sort never re-reads the keys to count them).
__global__ void __launch_bounds__(HIST_THREADS) pack_keys_kernel(const i32* __restrict__ grp, const i64* __restrict__ s,
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
        k = (L.rank_bits ? (g << (2 * L.B)) : 0ull) | (((u64)a - L.cmin) << L.B) | ((u64)b - L.cmin);
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
                             cudaStream_t st) {
  const size_t smem = hist_smem_bytes(rp);
  BF_CUDA(cudaFuncSetAttribute(radix_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BF_LAUNCH(PC_SORT_MISC, 0, st,
            radix_hist_kernel<<<stream_grid(n_max, 4096), HIST_THREADS, smem, st>>>(keys, n_max, n_dev, rp, ghist));
  return BF_OK;
}

}  // namespace bf
