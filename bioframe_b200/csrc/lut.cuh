// Direct-address acceleration of lower/upper bounds over a sorted array of linear coordinates.
//
// Queries that arrive in arbitrary order (coverage, count_overlaps, closest work in input row order) make
// every step of a plain binary search a random L2 sector load: ~14 sectors per search at 1e7 targets, and
// the kernels were bound by L2 sector traffic (profiles/NOTES.md).  The linear axis is cut into 2^k equal
// buckets (about one per entry); tab[b] = number of entries below the bucket's first coordinate.  A probe
// reads tab[b], tab[b+1] (one sector) and finishes with a search over the handful of entries of its bucket.
#pragma once
#include "common.cuh"

namespace bf {

struct Lut {
  const u32* tab;  // [nbuckets + 1]
  u64 base;        // coordinate of bucket 0
  int shift;       // log2(bucket width)
  u32 nbuckets;
  int key_shift;   // entries are keys[m * stride] >> key_shift (packed keys: LB; plain coordinates: 0)
  int stride;      // distance between entries in 8-byte words (1: plain array; 4: first word of 32-byte records)
};

static inline u32 lut_buckets_for(i64 n) {  // power of two, one bucket per 2-4 entries (a 32-byte sector holds
  u32 b = 1024;                              // 4 keys, and a small table stays resident in L2), at least 1024
  while ((i64)b * 4 < n && b < (1u << 30)) b <<= 1;
  return b;
}
static inline size_t lut_words(i64 n) { return (size_t)lut_buckets_for(n) + 2; }

// span = number of coordinates covered (probes outside [base, base + span) are clamped to the end buckets)
static inline Lut make_lut(u32* tab, i64 n, u64 base, u64 span, int key_shift, int stride = 1) {
  Lut t;
  t.tab = tab;
  t.base = base;
  t.nbuckets = lut_buckets_for(n);
  int shift = 0;
  while (shift < 63 && (span >> shift) > (u64)t.nbuckets) ++shift;  // nbuckets << shift >= span
  if (((u64)t.nbuckets << shift) < span) ++shift;
  t.shift = shift;
  t.key_shift = key_shift;
  t.stride = stride;
  return t;
}

static __global__ void __launch_bounds__(256) lut_build_kernel(const u64* __restrict__ keys, i64 n_host,
                                                               const i64* __restrict__ n_dev, Lut t,
                                                               u32* __restrict__ tab) {
  const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > t.nbuckets) return;
  const i64 n = n_dev ? *n_dev : n_host;
  const u64 x = t.base + ((u64)b << t.shift);  // first coordinate of bucket b (b == nbuckets: one past the end)
  const bool past = b > 0 && x <= t.base;      // wrapped around the 64-bit axis
  tab[b] = past ? (u32)n : (u32)lower_bound_by(0, n, [&](i64 m) { return (keys[m * t.stride] >> t.key_shift) < x; });
}

// #{entries < x} (strict) or #{entries <= x}
template <bool INCLUSIVE>
__device__ __forceinline__ i64 lut_count_below(const u64* __restrict__ keys, const Lut& t, u64 x) {
  u64 b = x < t.base ? 0ull : (x - t.base) >> t.shift;
  if (b >= t.nbuckets) b = t.nbuckets - 1;
  i64 lo = t.tab[b], hi = t.tab[b + 1];
  // a probe clamped into an end bucket may lie outside it: widen to the array ends on that side
  if (x < t.base) lo = 0;
  while (lo < hi) {
    const i64 mid = lo + ((hi - lo) >> 1);
    const u64 v = keys[mid * t.stride] >> t.key_shift;
    if (INCLUSIVE ? (v <= x) : (v < x))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

static int launch_lut_build(const u64* keys, i64 n, const i64* n_dev, const Lut& t, u32* tab, cudaStream_t st,
                            int prof_cls) {
  BF_LAUNCH(prof_cls, 0, st,
            lut_build_kernel<<<(unsigned)ceil_div((i64)t.nbuckets + 1, 256), 256, 0, st>>>(keys, n, n_dev, t, tab));
  return BF_OK;
}

}  // namespace bf
