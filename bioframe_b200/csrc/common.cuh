// Shared device/host helpers for the bioframe_b200 CUDA library (sm_100a).
// Everything on this path is int64/uint64 integer work bound by HBM bandwidth:
// no tensor cores, coalesced 8/16-byte accesses, shared-memory staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef int i32;
typedef unsigned short u16;
typedef unsigned char u8;

#define BF_OK 0
#define BF_ERR_ARG -1       // bad argument (null pointer, negative size, bad enum)
#define BF_ERR_WORKSPACE -2 // workspace too small
#define BF_ERR_RANGE -3     // coordinate range / group count does not fit the key layouts
#define BF_ERR_CUDA -4      // CUDA runtime error (see bf_last_error)
#define BF_ERR_GROUP -5     // group code outside [0, n_groups)
#define BF_ERR_STATE -6     // emit called on a plan that did not complete count

// Checked build (python -m bioframe_b200.build --checked -> libbioframe_b200_checked.so): every index the kernels
// compute for a global load or store outside plain "i < n" loops is range-checked on the device and the kernel
// traps on a violation (the launch then fails with a CUDA error and the test with it).  compute-sanitizer is
// closed on the GPU pool; tests/test_gpu_checked_build.py runs the fuzz / edge / sharding cases on this build.
#ifdef BF_CHECKED
#define BF_DEVICE_CHECK(cond)              \
  do {                                     \
    if (!(cond)) {                         \
      printf("BF_DEVICE_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      __trap();                            \
    }                                      \
  } while (0)
#else
#define BF_DEVICE_CHECK(cond) ((void)0)
#endif

namespace bf {

void set_error(const char* fmt, ...);

#define BF_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      bf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return BF_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define BF_LAUNCH_CHECK() BF_CUDA(cudaGetLastError())

// Kernel classes for the launch counter / CUDA-event profiler (bf_prof_*).
enum ProfClass {
  PC_EXTENT = 0,   // coordinate extent + group-code check
  PC_PACK,         // key packing + digit histograms
  PC_SORT_PASS,    // onesweep radix pass (the dominant kernel)
  PC_SORT_MISC,    // histogram scans, generic histograms, iota
  PC_TIEFIX,       // (end', row) order among rows sharing (group, start)
  PC_SEGMENT,      // group segment table
  PC_RUNMAX,       // running max of end
  PC_SEARCH,       // binary search + count
  PC_SCAN,         // exclusive scans (spine + downsweep)
  PC_TABLE,        // per-group offset tables
  PC_EMIT,         // pair emission
  PC_UNPAIRED,     // unpaired-row compaction / emission
  PC_CLUSTER,      // cluster/merge/complement kernels
  PC_CLOSEST,      // closest-k kernels
  PC_COVERAGE,     // coverage kernels
  PC_GATHER,       // misc gathers / fills
  PC_NUM
};
void prof_pre(int cls, cudaStream_t st);
void prof_post(int cls, cudaStream_t st, u64 bytes);

// Every kernel launch of the library goes through this macro: it counts the
// launch and, when profiling is enabled, brackets it with CUDA events on the
// launching stream. `bytes` = algorithmic bytes the launch moves.
#define BF_LAUNCH(cls, bytes, st, ...)            \
  do {                                            \
    bf::prof_pre((cls), (st));                    \
    __VA_ARGS__;                                  \
    BF_LAUNCH_CHECK();                            \
    bf::prof_post((cls), (st), (u64)(bytes));     \
  } while (0)

#define BF_TRY(expr)           \
  do {                         \
    int _s = (expr);           \
    if (_s != BF_OK) return _s; \
  } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }
static inline int bit_width(u64 x) {
  int b = 0;
  while (x) {
    ++b;
    x >>= 1;
  }
  return b;
}

// Bump allocator over a caller-owned workspace. With base == nullptr it only
// measures (used by the *_workspace_bytes entry points).
struct Arena {
  char* base;
  size_t off;
  explicit Arena(void* b) : base((char*)b), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t at = off;
    off = align_up(off + count * sizeof(T));
    return base ? (T*)(base + at) : (T*)nullptr;
  }
};

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Shared-memory reads through an explicit 32-bit shared address.  With array indexing the compiler
// rebuilds the (cluster-aware) shared window address in front of many accesses -- an S2UR of
// SR_CgaCtaId plus a ULEA each time, 41 of them in the unrolled search loop -- and the S2UR latency sits
// on the dependent chain of every search step.  Computing the address once and using ld.shared keeps
// the loop to one LDS per step.
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 lds_u32(u32 addr) {
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// first index in [lo, hi) with !(a[idx] < x)   (np.searchsorted side="left")
template <typename F>
__device__ __forceinline__ i64 lower_bound_by(i64 lo, i64 hi, F less_than_x) {
  while (lo < hi) {
    i64 mid = lo + ((hi - lo) >> 1);
    if (less_than_x(mid))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

}  // namespace bf
