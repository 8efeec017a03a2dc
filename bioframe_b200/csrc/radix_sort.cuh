// Stable LSD radix sort of 64-bit keys (+ optional 32-bit row-id payload),
// "onesweep" style: one up-front histogram of every digit place, then ONE
// kernel per digit place that reads each key once and writes it once
// (12 B in + 12 B out per row with payload).  Per pass and tile:
//   warp-striped 8-byte loads -> warp-level digit ranking (warp-private shared-memory
//   counters bumped with atomics, match.any among colliding lanes) -> tile digit offsets ->
//   decoupled look-back over per-tile digit counts (one 64-bit status word per
//   (tile, digit): tag | flag | value, so no fence and no per-pass memset) ->
//   keys/payload staged in shared memory in sorted order -> coalesced runs out.
// Tiles take their index from an atomic ticket so a tile never waits on a tile
// that has not started.  Stability (ties keep input order) is what makes the
// result equal np.lexsort([end, start]) tie-broken by row (arrops.py:332-333).
//
// Replaces: np.lexsort / np.argsort call sites of the reference
// (src/bioframe/core/arrops.py:332-333, 459, 562, 580, 738).
#pragma once
#include "common.cuh"

namespace bf {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BLOCKS_PER_SM = 3;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 4096 keys per tile
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 8;
constexpr int RS_LOOK = 4;  // look-back predecessors fetched per round trip
constexpr u64 RS_PERSISTENT_BLOCKS = 148 * RS_BLOCKS_PER_SM;  // one wave of resident blocks, each loops over tiles

struct RadixPlan {
  int passes;
  int shift[RS_MAX_PASSES];
  int bits[RS_MAX_PASSES];
};

// Split key bits [begin_bit, end_bit) into ceil(nbits/8) digit places of
// near-equal width (<= 8 bits each).
static inline RadixPlan make_radix_plan(int begin_bit, int end_bit) {
  RadixPlan rp;
  int nbits = end_bit - begin_bit;
  if (nbits < 0) nbits = 0;
  rp.passes = (nbits + 7) / 8;
  int at = begin_bit;
  for (int i = 0; i < RS_MAX_PASSES; ++i) {
    rp.shift[i] = 0;
    rp.bits[i] = 0;
  }
  for (int i = 0; i < rp.passes; ++i) {
    int w = nbits / rp.passes + (i < nbits % rp.passes ? 1 : 0);
    rp.shift[i] = at;
    rp.bits[i] = w;
    at += w;
  }
  return rp;
}

static inline size_t radix_status_words(u64 n) { return (size_t)ceil_div((i64)n, RS_TILE) * RS_BINS; }

// Scratch a sort needs besides its double buffers.
struct RadixScratch {
  u64* hist;     // [RS_MAX_PASSES][256] counts, then exclusive digit bases
  u64* status;   // [tiles][256] look-back words
  u32* tickets;  // [RS_MAX_PASSES] tile tickets
};
static inline RadixScratch take_radix_scratch(Arena& a, u64 n_max) {
  RadixScratch s;
  s.hist = a.take<u64>(RS_MAX_PASSES * RS_BINS);
  s.status = a.take<u64>(radix_status_words(n_max) + RS_BINS);
  s.tickets = a.take<u32>(RS_MAX_PASSES);
  return s;
}

// ---- histogram of all digit places, callable from any key-producing kernel ----
// One [passes][256] table per block in shared memory, updated with shared-memory
// atomics.  Measured on B200 (profiles/microbench/hist_bench.cu, 1e8 keys, four
// digit places): 0.18 ms with atomics against 1.28 ms for ballot-based
// warp aggregation and 2.75 ms for match.any -- even when every lane of a warp
// hits the same counter (sorted input), so there is no reason to aggregate by hand.
constexpr int HIST_THREADS = 256;
static inline size_t hist_smem_bytes(const struct RadixPlan& rp);

__device__ __forceinline__ void hist_add(u32* tab, const RadixPlan& rp, u64 key, bool valid) {
#pragma unroll
  for (int i = 0; i < RS_MAX_PASSES; ++i) {
    if (i < rp.passes && valid) {
      const u32 d = (u32)(key >> rp.shift[i]) & ((1u << rp.bits[i]) - 1u);
      atomicAdd(&tab[i * RS_BINS + d], 1u);
    }
  }
}
__device__ __forceinline__ void hist_clear(u32* tabs, const RadixPlan& rp) {
  for (int i = threadIdx.x; i < rp.passes * RS_BINS; i += blockDim.x) tabs[i] = 0;
}
__device__ __forceinline__ void hist_flush(const u32* tabs, const RadixPlan& rp, u64* ghist) {
  for (int i = threadIdx.x; i < rp.passes * RS_BINS; i += blockDim.x) {
    const u32 c = tabs[i];
    if (c) atomicAdd(&ghist[i], (u64)c);
  }
}
static inline size_t hist_smem_bytes(const RadixPlan& rp) {
  return (size_t)(rp.passes > 0 ? rp.passes : 1) * RS_BINS * sizeof(u32);
}

// Generic histogram over an existing key array (n may live on the device).
// Dynamic shared memory: hist_smem_bytes(rp).
static __global__ void __launch_bounds__(HIST_THREADS) radix_hist_kernel(const u64* __restrict__ keys, u64 n_host,
                                                                   const u64* __restrict__ n_dev, RadixPlan rp,
                                                                   u64* __restrict__ ghist) {
  extern __shared__ u32 hist_tabs[];
  const u64 n = n_dev ? *n_dev : n_host;
  hist_clear(hist_tabs, rp);
  __syncthreads();
  u32* wtab = hist_tabs;
  const u64 per_block = (u64)HIST_THREADS * 16ull;
  for (u64 base = (u64)blockIdx.x * per_block; base < n; base += (u64)gridDim.x * per_block) {
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      u64 idx = base + (u64)j * HIST_THREADS + threadIdx.x;
      bool ok = idx < n;
      u64 k = ok ? keys[idx] : 0;
      hist_add(wtab, rp, k, ok);
    }
  }
  __syncthreads();
  hist_flush(hist_tabs, rp, ghist);
}

// counts -> exclusive digit bases, one block per digit place
static __global__ void __launch_bounds__(RS_BINS) radix_bases_kernel(u64* __restrict__ ghist) {
  __shared__ u64 warp_tot[RS_BINS / 32];
  u64* row = ghist + (size_t)blockIdx.x * RS_BINS;
  const u32 d = threadIdx.x, lane = d & 31, w = d >> 5;
  u64 c = row[d];
  u64 inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u64 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (u32)o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  u64 pre = 0;
  for (u32 i = 0; i < w; ++i) pre += warp_tot[i];
  row[d] = pre + inc - c;
}

constexpr u64 RS_VAL_MASK = (1ull << 56) - 1;
constexpr u64 RS_FLAG_AGG = 1ull << 56;
constexpr u64 RS_FLAG_INC = 2ull << 56;

constexpr size_t rs_smem_bytes(bool has_val) {
  return (size_t)RS_TILE * 8 + (has_val ? (size_t)RS_TILE * 4 : 0) + (size_t)RS_WARPS * RS_BINS * 4 +
         (size_t)RS_BINS * 4 + (size_t)RS_BINS * 8;
}

template <bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS, RS_BLOCKS_PER_SM)
    onesweep_pass_kernel(const u64* __restrict__ kin, u64* __restrict__ kout, const u32* __restrict__ vin,
                         u32* __restrict__ vout, u64 n_host, const u64* __restrict__ n_dev, int shift,
                         u32 dmask, u64 tag, const u64* __restrict__ digit_base, u64* __restrict__ status,
                         u32* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  u64* s_keys = (u64*)rs_smem;
  u32* s_vals = (u32*)(s_keys + RS_TILE);
  u32* s_wcnt = s_vals + (HAS_VAL ? RS_TILE : 0);  // [RS_WARPS][256]
  u32* s_tbase = s_wcnt + RS_WARPS * RS_BINS;      // [256] exclusive digit offsets inside the tile
  u64* s_gofs = (u64*)(s_tbase + RS_BINS);         // [256] global offset minus s_tbase
  __shared__ u32 s_tile;
  __shared__ u32 s_wtot[RS_BINS / 32];

  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool dthread = tid < RS_BINS;  // threads 0..255 each own one digit in the per-tile bookkeeping
  const u32 d_own = tid & (RS_BINS - 1);
  const u64 n = n_dev ? *n_dev : n_host;
  // persistent block: take tiles by ticket until the input is exhausted
  for (;;) {
  __syncthreads();  // previous tile fully written out before shared memory is reused
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) s_wcnt[i] = 0;
  __syncthreads();
  const u64 tile = s_tile;
  const u64 tile_start = tile * RS_TILE;
  if (tile_start >= n) return;
  const u32 valid = (u32)((n - tile_start < (u64)RS_TILE) ? (n - tile_start) : (u64)RS_TILE);

  // ---- load (warp-striped: lane l of warp w owns tile items w*(32*IPT) + j*32 + l) ----
  u64 key[RS_IPT];
  const u32 wbase = warp * (RS_IPT * 32) + lane;
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j) {
    u32 li = wbase + j * 32;
    key[j] = (li < valid) ? kin[tile_start + li] : ~0ull;  // pads sort last inside the tile
  }

  // ---- rank inside the warp, digit by digit ----
  // Warp-private digit counters, bumped with shared-memory atomics (fast on B200, see hist_add).  A lane
  // reads its counter before and after the warp's adds: the difference is the number of lanes holding
  // the same digit.  Alone (the common case for 8-bit digits) its rank is the old count; otherwise the
  // few lanes that collided order themselves by lane with one match.any over just those lanes, which
  // keeps the pass stable.  Replaces 8 ballots + mask logic per key.
  u16 lpos[RS_IPT];
  u32* my_cnt = s_wcnt + warp * RS_BINS;
  const u32 lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j) {
    const u32 d = (u32)(key[j] >> shift) & dmask;
    const u32 before = my_cnt[d];
    __syncwarp();
    atomicAdd(&my_cnt[d], 1u);
    __syncwarp();
    const u32 same = my_cnt[d] - before;  // lanes of this round with digit d (>= 1)
    u32 r = before;
    const u32 crowd = __ballot_sync(0xffffffffu, same > 1);
    if (same > 1) r += __popc(__match_any_sync(crowd, d) & lt);
    lpos[j] = (u16)r;
  }
  __syncthreads();

  // ---- thread d: exclusive offsets of digit d across warps, tile digit count; publish the tile's
  //      digit counts right away so that later tiles can start summing them ----
  u64 cnt_valid = 0;
  u32 tbase_own = 0;
  {
    u32 tile_cnt = 0;
    if (dthread) {
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        u32 c = s_wcnt[w * RS_BINS + d_own];
        s_wcnt[w * RS_BINS + d_own] = tile_cnt;
        tile_cnt += c;
      }
    }
    // exclusive scan of tile_cnt over the 256 digits (warps 0..7)
    u32 inc = tile_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc += t;
    }
    if (dthread && lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    if (dthread) {
      u32 pre = 0;
#pragma unroll
      for (int w = 0; w < RS_BINS / 32; ++w)
        if ((u32)w < warp) pre += s_wtot[w];
      tbase_own = pre + inc - tile_cnt;
      s_tbase[d_own] = tbase_own;
      // pads all carry digit dmask and sit at the very end of the tile
      cnt_valid = tile_cnt;
      if (d_own == dmask) cnt_valid -= (u64)(RS_TILE - valid);
      st_relaxed_u64(status + tile * RS_BINS + d_own, tag | (tile == 0 ? RS_FLAG_INC : RS_FLAG_AGG) | cnt_valid);
    }
  }
  __syncthreads();

  // ---- stage keys in shared memory in sorted order (needs only tile-local offsets; the
  //      predecessors get this much more time to publish before the look-back polls them) ----
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j) {
    u32 d = (u32)(key[j] >> shift) & dmask;
    u32 p = s_tbase[d] + my_cnt[d] + lpos[j];
    BF_DEVICE_CHECK(p < (u32)RS_TILE);
    lpos[j] = (u16)p;
    s_keys[p] = key[j];
  }
  if (HAS_VAL) {
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
      u32 li = wbase + j * 32;
      u32 v = 0;
      if (li < valid) v = vin ? vin[tile_start + li] : (u32)(tile_start + li);
      s_vals[lpos[j]] = v;
    }
  }

  // ---- decoupled look-back: digits before this tile ----
  // RS_LOOK predecessors are fetched per round trip (independent loads).
  if (dthread) {
    u64 excl = 0;
    if (tile > 0) {
      i64 pt = (i64)tile - 1;  // nearest predecessor not yet accounted for
      for (;;) {
        u64 w[RS_LOOK];
#pragma unroll
        for (int b = 0; b < RS_LOOK; ++b)
          w[b] = pt - b >= 0 ? ld_relaxed_u64(status + (u64)(pt - b) * RS_BINS + d_own) : (tag | RS_FLAG_INC);
        bool finished = false, stalled = false;
        int consumed = 0;
#pragma unroll
        for (int b = 0; b < RS_LOOK; ++b) {
          if (!finished && !stalled) {
            const u64 flag = w[b] & (3ull << 56);
            // a word of an earlier pass (other tag) or an unpublished one: poll again from here
            if ((w[b] & ~(RS_VAL_MASK | (3ull << 56))) != tag || flag == 0) {
              stalled = true;
            } else {
              excl += w[b] & RS_VAL_MASK;
              ++consumed;
              finished = flag == RS_FLAG_INC;
            }
          }
        }
        if (finished) break;
        pt -= consumed;
        if (consumed == 0) __nanosleep(40);  // predecessor not published yet: leave the issue slots to others
      }
      st_relaxed_u64(status + tile * RS_BINS + d_own, tag | RS_FLAG_INC | (excl + cnt_valid));
    }
    s_gofs[d_own] = digit_base[d_own] + excl - (u64)tbase_own;
  }
  __syncthreads();

  // ---- coalesced write-out: consecutive slots of one digit are consecutive in memory ----
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    u32 idx = i * RS_THREADS + tid;
    if (idx < valid) {
      u64 k = s_keys[idx];
      u32 d = (u32)(k >> shift) & dmask;
      u64 dst = s_gofs[d] + idx;
      BF_DEVICE_CHECK(dst < n);
      kout[dst] = k;
      if (HAS_VAL) vout[dst] = s_vals[idx];
    }
  }
  }  // next tile
}

static __global__ void iota_u32_kernel(u32* __restrict__ out, u64 n) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (u32)i;
}

// Result location of a sort: which of the two buffers holds the sorted data.
struct SortResult {
  u64* keys;
  u32* vals;
  u64* keys_alt;
  u32* vals_alt;
};

// The histogram (scratch.hist, counts per digit place) must already be filled,
// either by radix_hist_kernel or by the kernel that produced the keys.
// vals_in == nullptr with HAS_VAL means "payload is the row number" (iota).
template <bool HAS_VAL>
static int radix_sort_passes(u64* kA, u64* kB, u32* vA, u32* vB, bool iota_vals, u64 n_max,
                             const u64* n_dev, const RadixPlan& rp, const RadixScratch& sc,
                             cudaStream_t stream, SortResult* out, int prof_cls = PC_SORT_PASS) {
  BF_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<HAS_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)rs_smem_bytes(HAS_VAL)));
  if (HAS_VAL && iota_vals && n_max > 0 && rp.passes == 0) {  // nothing to sort on: identity order
    BF_LAUNCH(PC_SORT_MISC, 0, stream, iota_u32_kernel<<<(unsigned)ceil_div((i64)n_max, 256), 256, 0, stream>>>(vA, n_max));
  }
  u64 *kin = kA, *kout = kB;
  u32 *vin = vA, *vout = vB;
  if (n_max > 0 && rp.passes > 0) {
    u64 tiles = (u64)ceil_div((i64)n_max, RS_TILE);
    if (tiles > RS_PERSISTENT_BLOCKS) tiles = RS_PERSISTENT_BLOCKS;
    BF_LAUNCH(PC_SORT_MISC, 0, stream, radix_bases_kernel<<<rp.passes, RS_BINS, 0, stream>>>(sc.hist));
    BF_CUDA(cudaMemsetAsync(sc.status, 0, radix_status_words(n_max) * sizeof(u64), stream));
    BF_CUDA(cudaMemsetAsync(sc.tickets, 0, RS_MAX_PASSES * sizeof(u32), stream));
    for (int p = 0; p < rp.passes; ++p) {
      const u32* vsrc = (HAS_VAL && !(iota_vals && p == 0)) ? vin : nullptr;
      BF_LAUNCH(prof_cls, n_dev ? 0ull : (HAS_VAL ? 24ull : 16ull) * n_max - ((HAS_VAL && !vsrc) ? 4ull * n_max : 0ull), stream, onesweep_pass_kernel<HAS_VAL><<<(unsigned)tiles, RS_THREADS, rs_smem_bytes(HAS_VAL), stream>>>(
          kin, kout, vsrc, vout, n_max, n_dev, rp.shift[p], (1u << rp.bits[p]) - 1u, (u64)(p + 1) << 58,
          sc.hist + (size_t)p * RS_BINS, sc.status, sc.tickets + p));
      u64* tk = kin;
      kin = kout;
      kout = tk;
      u32* tv = vin;
      vin = vout;
      vout = tv;
    }
  }
  out->keys = kin;
  out->vals = vin;
  out->keys_alt = kout;
  out->vals_alt = vout;
  return BF_OK;
}

}  // namespace bf
