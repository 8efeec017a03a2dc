// bf_sort_pairs: order result pairs by (row of set 1, row of set 2) on the device.
//
// bioframe.overlap(how='left') defaults to keep_order=True, i.e.
// `out_df.sort_values([index1, index2])` over the whole result frame (ops.py:549-550) -- with pandas that
// one call costs several times the join itself (18.6 s vs 4.0 s at 1e7 x 1e6).  When the frame
// indexes are monotonic, ordering by labels is ordering by row positions, so the glue can sort the
// (events1, events2) arrays here before it assembles the frame and skip the pandas sort.
//   pack (e1, e2) -> one 64-bit key (e1' << bits2 | e2'), -1 mapped to the largest value of its field
//   (pandas puts NA last) -> stable onesweep radix sort over the used bits -> unpack in place.
#include "common.cuh"
#include "keys.cuh"
#include "radix_sort.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

struct PairsWS {
  u64 *keyA, *keyB;
  RadixScratch radix;
};
static void layout_pairs(Arena& a, i64 n, PairsWS* w) {
  w->keyA = a.take<u64>(n + 1);
  w->keyB = a.take<u64>(n + 1);
  w->radix = take_radix_scratch(a, (u64)n);
}

__global__ void __launch_bounds__(HIST_THREADS)
    pack_pairs_kernel(const i64* __restrict__ e1, const i64* __restrict__ e2, u64 n, u64 na1, u64 na2, int bits2,
                      RadixPlan rp, u64* __restrict__ keys, u64* __restrict__ ghist) {
  extern __shared__ u32 hist_tabs[];
  hist_clear(hist_tabs, rp);
  __syncthreads();
  const u64 per_block = (u64)HIST_THREADS * 8ull;
  for (u64 base = (u64)blockIdx.x * per_block; base < n; base += (u64)gridDim.x * per_block) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u64 i = base + (u64)j * HIST_THREADS + threadIdx.x;
      const bool ok = i < n;
      u64 k = 0;
      if (ok) {
        const i64 a = e1[i], b = e2[i];
        k = ((a < 0 ? na1 : (u64)a) << bits2) | (b < 0 ? na2 : (u64)b);
        keys[i] = k;
      }
      hist_add(hist_tabs, rp, k, ok);
    }
  }
  __syncthreads();
  hist_flush(hist_tabs, rp, ghist);
}
__global__ void __launch_bounds__(256)
    unpack_pairs_kernel(const u64* __restrict__ keys, u64 n, u64 na1, u64 na2, int bits2, i64* __restrict__ e1,
                        i64* __restrict__ e2) {
  const u64 mask2 = (1ull << bits2) - 1ull;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    const u64 a = k >> bits2, b = k & mask2;
    e1[i] = a == na1 ? -1 : (i64)a;
    e2[i] = b == na2 ? -1 : (i64)b;
  }
}

static int sort_pairs_impl(i64* e1, i64* e2, i64 n, i64 max1, i64 max2, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
  if (n < 0 || max1 < 0 || max2 < 0) {
    set_error("bf_sort_pairs: bad argument (n=%lld max1=%lld max2=%lld)", (long long)n, (long long)max1, (long long)max2);
    return BF_ERR_ARG;
  }
  if (n == 0) return BF_OK;
  if (!e1 || !e2) {
    set_error("bf_sort_pairs: null pointer");
    return BF_ERR_ARG;
  }
  const u64 na1 = (u64)max1 + 1ull, na2 = (u64)max2 + 1ull;  // the value -1 is packed as
  const int bits1 = bit_width(na1), bits2 = bit_width(na2);
  if (bits1 + bits2 > 64) {
    set_error("bf_sort_pairs: row numbers need %d + %d > 64 bits", bits1, bits2);
    return BF_ERR_RANGE;
  }
  Arena measure(nullptr);
  PairsWS w;
  layout_pairs(measure, n, &w);
  if (!ws_ptr || ws_bytes < measure.off) {
    set_error("bf_sort_pairs: workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws_ptr);
  layout_pairs(arena, n, &w);
  const RadixPlan rp = make_radix_plan(0, bits1 + bits2);
  const size_t smem = hist_smem_bytes(rp);
  BF_CUDA(cudaMemsetAsync(w.radix.hist, 0, RS_MAX_PASSES * RS_BINS * sizeof(u64), st));
  BF_CUDA(cudaFuncSetAttribute(pack_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BF_LAUNCH(PC_GATHER, 24 * n, st,
            pack_pairs_kernel<<<stream_grid((u64)n, 2048), HIST_THREADS, smem, st>>>(e1, e2, (u64)n, na1, na2, bits2, rp,
                                                                                    w.keyA, w.radix.hist));
  SortResult r;
  BF_TRY(radix_sort_passes<false>(w.keyA, w.keyB, nullptr, nullptr, false, (u64)n, nullptr, rp, w.radix, st, &r,
                                  PC_SORT_PASS));
  BF_LAUNCH(PC_GATHER, 24 * n, st,
            unpack_pairs_kernel<<<stream_grid((u64)n, 1024), 256, 0, st>>>(r.keys, (u64)n, na1, na2, bits2, e1, e2));
  return BF_OK;
}

// Coordinates of the reported pairs: the four gathers + max / min of return_overlap=True
// (ops.py:480-497, closest ops.py:1187-1194) and closest's distance (ops.py:1207-1221), one pass over
// the pair list while it still is on the device.  Rows with a missing partner (-1) get 0; the frame glue
// masks them like the reference does (ops.py:541-544, 1205, 1222).
__global__ void __launch_bounds__(256)
    pair_coords_kernel(const i64* __restrict__ ev1, const i64* __restrict__ ev2, i64 n, const i64* __restrict__ s1,
                       const i64* __restrict__ e1, const i64* __restrict__ s2, const i64* __restrict__ e2,
                       i64* __restrict__ ostart, i64* __restrict__ oend, i64* __restrict__ dist) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const i64 a = ev1[i], b = ev2[i];
    i64 os = 0, oe = 0, d = 0;
    if (a >= 0 && b >= 0) {
      const i64 as = s1[a], ae = e1[a], bs = s2[b], be = e2[b];
      os = as > bs ? as : bs;
      oe = ae < be ? ae : be;
      const i64 left = as - be, right = bs - ae;  // gap when the partner lies to the left / right
      d = left > right ? left : right;
      if (d < 0) d = 0;
    }
    if (ostart) ostart[i] = os;
    if (oend) oend[i] = oe;
    if (dist) dist[i] = d;
  }
}

// Column gather: out[i] = src[idx[i]] for a fixed-width column (4- or 8-byte values), 0 where idx[i] is -1.
// The df.iloc[events] of the frame assembly (ops.py:499-508, 1224-1234) for numeric columns, done where the
// pair list lives; `valid` (optional) is the Arrow-style validity byte per row.
template <typename T>
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const T* __restrict__ src, const i64* __restrict__ idx, i64 n, T* __restrict__ out,
                       u8* __restrict__ valid) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const i64 r = idx[i];
    out[i] = r >= 0 ? src[r] : T(0);
    if (valid) valid[i] = r >= 0;
  }
}

static int gather_rows_impl(const void* src, i32 itemsize, const i64* idx, i64 n, void* out, u8* valid, cudaStream_t st) {
  if (n < 0 || (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8)) {
    set_error("bf_gather_rows: bad argument (n_rows=%lld itemsize=%d; 1-, 2-, 4- and 8-byte columns are supported)",
              (long long)n, itemsize);
    return BF_ERR_ARG;
  }
  if (n == 0) return BF_OK;
  if (!src || !idx || !out) {
    set_error("bf_gather_rows: null pointer");
    return BF_ERR_ARG;
  }
  const unsigned grid = stream_grid((u64)n, 1024);
  if (itemsize == 8) {
    BF_LAUNCH(PC_GATHER, 24 * n, st,
              gather_rows_kernel<u64><<<grid, 256, 0, st>>>((const u64*)src, idx, n, (u64*)out, valid));
  } else if (itemsize == 4) {
    BF_LAUNCH(PC_GATHER, 16 * n, st,
              gather_rows_kernel<u32><<<grid, 256, 0, st>>>((const u32*)src, idx, n, (u32*)out, valid));
  } else if (itemsize == 2) {  // categorical codes of up to 32767 categories
    BF_LAUNCH(PC_GATHER, 12 * n, st,
              gather_rows_kernel<u16><<<grid, 256, 0, st>>>((const u16*)src, idx, n, (u16*)out, valid));
  } else {  // int8 categorical codes (chromosomes), bool columns, the mask bytes of nullable columns
    BF_LAUNCH(PC_GATHER, 10 * n, st,
              gather_rows_kernel<u8><<<grid, 256, 0, st>>>((const u8*)src, idx, n, (u8*)out, valid));
  }
  return BF_OK;
}

static int pair_coords_impl(const i64* ev1, const i64* ev2, i64 n, const i64* s1, const i64* e1, const i64* s2,
                            const i64* e2, i64* ostart, i64* oend, i64* dist, cudaStream_t st) {
  if (n < 0) {
    set_error("bf_pair_coords: bad argument (n_rows=%lld)", (long long)n);
    return BF_ERR_ARG;
  }
  if (n == 0 || (!ostart && !oend && !dist)) return BF_OK;
  if (!ev1 || !ev2 || !s1 || !e1 || !s2 || !e2) {
    set_error("bf_pair_coords: null pointer");
    return BF_ERR_ARG;
  }
  const int outs = (ostart != nullptr) + (oend != nullptr) + (dist != nullptr);
  BF_LAUNCH(PC_GATHER, (16 + 32 + 8 * outs) * n, st,
            pair_coords_kernel<<<stream_grid((u64)n, 1024), 256, 0, st>>>(ev1, ev2, n, s1, e1, s2, e2, ostart, oend, dist));
  return BF_OK;
}

}  // namespace bf

extern "C" {
int bf_gather_rows(const void* src, int32_t itemsize, const int64_t* row_idx, int64_t n_rows, void* out,
                   uint8_t* valid_out, void* stream) {
  return bf::gather_rows_impl(src, itemsize, (const i64*)row_idx, n_rows, out, (u8*)valid_out, (cudaStream_t)stream);
}
int bf_pair_coords(const int64_t* events1, const int64_t* events2, int64_t n_rows, const int64_t* start1,
                   const int64_t* end1, const int64_t* start2, const int64_t* end2, int64_t* overlap_start_out,
                   int64_t* overlap_end_out, int64_t* distance_out, void* stream) {
  return bf::pair_coords_impl((const i64*)events1, (const i64*)events2, n_rows, (const i64*)start1, (const i64*)end1,
                              (const i64*)start2, (const i64*)end2, (i64*)overlap_start_out, (i64*)overlap_end_out,
                              (i64*)distance_out, (cudaStream_t)stream);
}
size_t bf_sort_pairs_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return 0;
  bf::Arena a(nullptr);
  bf::PairsWS w;
  bf::layout_pairs(a, n_rows, &w);
  return a.off;
}
int bf_sort_pairs(int64_t* events1, int64_t* events2, int64_t n_rows, int64_t max1, int64_t max2, void* ws,
                  size_t ws_bytes, void* stream) {
  return bf::sort_pairs_impl((i64*)events1, (i64*)events2, n_rows, max1, max2, ws, ws_bytes, (cudaStream_t)stream);
}
}
