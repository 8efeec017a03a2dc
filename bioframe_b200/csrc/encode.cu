// Chromosome (group) encoding on the device: the a14 row of SURVEY.md 8a.
//
// The reference finds its groups with a hash groupby per call (df.groupby(...).indices / .groups,
// src/bioframe/ops.py:287-289, 631, 772, 983-985): 2-15 % of a call, more for object columns
// (docs/guide-performance.ipynb:497-498 recommends categoricals for that reason).  A pandas categorical already
// IS a dictionary encoding -- one small integer code per row, -1 for a null -- so here the column crosses PCIe
// as those 1- or 2-byte codes and two passes over them do the rest:
//   bf_group_codes_present  which categories occur at all (the "observed" keys of the groupby)
//   bf_group_codes_apply    code -> group rank through a lookup table the glue builds from the observed keys of
//                           both frames in the reference's visiting order
// 1e8 rows: 0.1 GB read + 0.4 GB written, against ~1.3 s of numpy passes on the host.
#include "common.cuh"
#include "keys.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

constexpr int ENC_MAX_CATEGORIES = 32767;  // int16 codes

template <typename T>
__global__ void __launch_bounds__(256) codes_present_kernel(const T* __restrict__ codes, i64 n, i32 n_cat,
                                                            i32* __restrict__ present) {
  extern __shared__ u32 s_seen[];  // one bit per slot: slot 0 = null (-1), slot c + 1 = category c
  const i32 words = (n_cat + 1 + 31) / 32;
  for (i32 i = threadIdx.x; i < words; i += blockDim.x) s_seen[i] = 0;
  __syncthreads();
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const i32 slot = (i32)codes[i] + 1;
    if (slot >= 0 && slot <= n_cat) {
      const u32 bit = 1u << (slot & 31);
      if (!(s_seen[slot >> 5] & bit)) atomicOr(&s_seen[slot >> 5], bit);  // read first: after the first rows nothing is new
    } else {
      present[n_cat + 1] = 1;  // a code outside [-1, n_cat): reported as an error by the host wrapper
    }
  }
  __syncthreads();
  for (i32 s = threadIdx.x; s <= n_cat; s += blockDim.x)
    if ((s_seen[s >> 5] >> (s & 31)) & 1u) present[s] = 1;
}

template <typename T>
__global__ void __launch_bounds__(256) codes_apply_kernel(const T* __restrict__ codes, i64 n, const i32* __restrict__ lut,
                                                          i32 n_cat, i32* __restrict__ out) {
  extern __shared__ i32 s_lut[];
  for (i32 i = threadIdx.x; i <= n_cat; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const i32 slot = (i32)codes[i] + 1;
    out[i] = (slot >= 0 && slot <= n_cat) ? s_lut[slot] : -1;
  }
}

static int check_codes_args(const char* who, const void* codes, i32 itemsize, i64 n, i32 n_cat) {
  if (n < 0 || n_cat < 0 || n_cat > ENC_MAX_CATEGORIES || (itemsize != 1 && itemsize != 2) || (n > 0 && !codes)) {
    set_error("%s: bad argument (n=%lld n_categories=%d itemsize=%d; int8 / int16 codes, at most %d categories)", who,
              (long long)n, n_cat, itemsize, ENC_MAX_CATEGORIES);
    return BF_ERR_ARG;
  }
  return BF_OK;
}

}  // namespace bf

extern "C" {

int bf_group_codes_present(const void* codes, int32_t itemsize, int64_t n, int32_t n_categories, int32_t* present_out,
                           void* stream) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  BF_TRY(check_codes_args("bf_group_codes_present", codes, itemsize, n, n_categories));
  if (!present_out) {
    set_error("bf_group_codes_present: null output");
    return BF_ERR_ARG;
  }
  BF_CUDA(cudaMemsetAsync(present_out, 0, ((size_t)n_categories + 2) * sizeof(i32), st));
  if (n == 0) return BF_OK;
  const size_t smem = (size_t)((n_categories + 1 + 31) / 32) * sizeof(u32);
  const unsigned grid = stream_grid((u64)n, 256 * 32);
  if (itemsize == 1) {
    BF_LAUNCH(PC_GATHER, n, st, codes_present_kernel<signed char><<<grid, 256, smem, st>>>((const signed char*)codes, n, n_categories, present_out));
  } else {
    BF_LAUNCH(PC_GATHER, 2 * n, st, codes_present_kernel<short><<<grid, 256, smem, st>>>((const short*)codes, n, n_categories, present_out));
  }
  return BF_OK;
}

int bf_group_codes_apply(const void* codes, int32_t itemsize, int64_t n, const int32_t* lut, int32_t n_categories,
                         int32_t* grp_out, void* stream) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  BF_TRY(check_codes_args("bf_group_codes_apply", codes, itemsize, n, n_categories));
  if (n == 0) return BF_OK;
  if (!lut || !grp_out) {
    set_error("bf_group_codes_apply: null pointer");
    return BF_ERR_ARG;
  }
  const size_t smem = ((size_t)n_categories + 1) * sizeof(i32);
  if (smem > 48 * 1024) {
    BF_CUDA(cudaFuncSetAttribute(codes_apply_kernel<signed char>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BF_CUDA(cudaFuncSetAttribute(codes_apply_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const unsigned grid = stream_grid((u64)n, 256 * 16);
  if (itemsize == 1) {
    BF_LAUNCH(PC_GATHER, 5 * n, st, codes_apply_kernel<signed char><<<grid, 256, smem, st>>>((const signed char*)codes, n, (const i32*)lut, n_categories, (i32*)grp_out));
  } else {
    BF_LAUNCH(PC_GATHER, 6 * n, st, codes_apply_kernel<short><<<grid, 256, smem, st>>>((const short*)codes, n, (const i32*)lut, n_categories, (i32*)grp_out));
  }
  return BF_OK;
}

}  // extern "C"
