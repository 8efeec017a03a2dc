// bf_download_rows: bring an int64 device array of row numbers (the events1 / events2 of
// ops._overlap_intidxs, ops.py:242-338, or any other result whose values lie in [-1, 2^31)) into an int64 host
// array.  A call through host buffers is bound by PCIe (profiles/NOTES.md: 14 GB of pairs at ~55 GB/s), and
// row numbers of sets below 2^31 rows carry 4 bytes of information: they cross the bus as int32 and are
// widened on the host by a few threads with streaming stores while the next chunk is in flight.
//
//   per chunk:  narrow kernel (int64 -> int32, checks the range)  ->  cudaMemcpyAsync to pinned staging  ->
//               event;  host: wait for the previous chunk's event, widen it into the destination
//   two device and two pinned staging chunks: the bus carries chunk k + 1 while the host widens chunk k.
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "common.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

__global__ void __launch_bounds__(256)
    narrow_rows_kernel(const i64* __restrict__ src, i64 n, i32* __restrict__ dst, i64* __restrict__ bad) {
  const i64 stride = (i64)gridDim.x * blockDim.x * 4;
  for (i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const longlong2 a = *reinterpret_cast<const longlong2*>(src + i);
      const longlong2 b = *reinterpret_cast<const longlong2*>(src + i + 2);
      if (((a.x + 1) | (a.y + 1) | (b.x + 1) | (b.y + 1)) >> 31) *bad = 1;  // outside [-1, 2^31 - 1)
      *reinterpret_cast<int4*>(dst + i) = make_int4((i32)a.x, (i32)a.y, (i32)b.x, (i32)b.y);
    } else {
      for (i64 j = i; j < n; ++j) {
        const i64 v = src[j];
        if ((v + 1) >> 31) *bad = 1;
        dst[j] = (i32)v;
      }
    }
  }
}

static void widen_plain(const i32* src, i64* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}
#if defined(__x86_64__)
// streaming stores: the destination is written once and not read here, so it should not pull its lines
// into the cache first (2x the memory traffic; profiles/microbench/widen_bench.cpp)
__attribute__((target("avx2"))) static void widen_stream(const i32* src, i64* dst, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) {
    dst[i] = src[i];
    ++i;
  }
  for (; i + 8 <= n; i += 8) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_cvtepi32_epi64(a));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 4), _mm256_cvtepi32_epi64(b));
  }
  for (; i < n; ++i) dst[i] = src[i];
  _mm_sfence();
}
#endif

static void widen_range(const i32* src, i64* dst, size_t n) {
#if defined(__x86_64__)
  static const bool stream = __builtin_cpu_supports("avx2");
  if (stream) {
    widen_stream(src, dst, n);
    return;
  }
#endif
  widen_plain(src, dst, n);
}

// Worker threads of one download call.  They live for the whole call and take every chunk's slices as they are
// posted (spinning in between: chunks arrive every few hundred microseconds), so that small chunks -- which
// stay in the last-level cache between the DMA write and the widening read -- cost no thread start each.
struct WidenPool {
  std::vector<std::thread> threads;
  std::atomic<i64> posted{0};   // number of chunks handed out so far
  std::atomic<i64> finished{0}; // worker-slices completed (threads.size() per chunk)
  std::atomic<bool> stop{false};
  const i32* src = nullptr;
  i64* dst = nullptr;
  size_t n = 0;
  int parts = 1;

  explicit WidenPool(int nthreads) : parts(nthreads < 1 ? 1 : nthreads) {
    for (int k = 1; k < parts; ++k) threads.emplace_back([this, k] { work(k); });
  }
  ~WidenPool() {
    stop.store(true, std::memory_order_release);
    for (auto& t : threads) t.join();
  }
  void slice(int k) const { widen_range(src + n * k / parts, dst + n * k / parts, n * (k + 1) / parts - n * k / parts); }
  void work(int k) {
    i64 seen = 0;
    for (;;) {
      // a short spin (the next chunk usually is a few microseconds away), then yield the core: while the
      // PCIe copy is the bottleneck the workers must not fight the copy / driver threads for cores
      for (unsigned spins = 0; posted.load(std::memory_order_acquire) == seen; ++spins) {
        if (stop.load(std::memory_order_acquire)) return;
        if (spins < 2000) {
#if defined(__x86_64__)
          _mm_pause();
#endif
        } else {
          std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
      }
      ++seen;
      slice(k);
      finished.fetch_add(1, std::memory_order_release);
    }
  }
  // widen one chunk with all threads; returns when it is complete
  void run(const i32* s, i64* d, size_t count) {
    src = s;
    dst = d;
    n = count;
    const i64 target = finished.load(std::memory_order_relaxed) + (i64)threads.size();
    posted.fetch_add(1, std::memory_order_release);
    slice(0);
    for (unsigned spins = 0; finished.load(std::memory_order_acquire) != target; ++spins) {
      if (spins < 20000) {
#if defined(__x86_64__)
        _mm_pause();
#endif
      } else {
        std::this_thread::yield();
      }
    }
  }
};

static int download_rows_impl(const i64* dev_src, i64 n, i64* host_dst, void* dev_stage, size_t dev_bytes, void* pinned,
                              size_t pinned_bytes, int threads, cudaStream_t st) {
  if (n < 0 || threads < 1) {
    set_error("bf_download_rows: bad argument (n=%lld n_threads=%d)", (long long)n, threads);
    return BF_ERR_ARG;
  }
  if (n == 0) return BF_OK;
  if (!dev_src || !host_dst || !dev_stage || !pinned) {
    set_error("bf_download_rows: null pointer");
    return BF_ERR_ARG;
  }
  if (reinterpret_cast<uintptr_t>(dev_src) & 15) {
    set_error("bf_download_rows: dev_src must be 16-byte aligned");
    return BF_ERR_ARG;
  }
  // staging: [flag (256 B)] [chunk 0] [chunk 1] on the device, [chunk 0] [chunk 1] pinned on the host
  const size_t usable = (dev_bytes > 256 ? dev_bytes - 256 : 0) < pinned_bytes ? (dev_bytes > 256 ? dev_bytes - 256 : 0) : pinned_bytes;
  const i64 chunk = (i64)((usable / 2 / sizeof(i32)) & ~(size_t)63);
  if (chunk < 64) {
    set_error("bf_download_rows: staging too small (device %zu bytes, pinned %zu bytes)", dev_bytes, pinned_bytes);
    return BF_ERR_WORKSPACE;
  }
  i64* bad = reinterpret_cast<i64*>(dev_stage);
  i32* dstage[2] = {reinterpret_cast<i32*>((char*)dev_stage + 256), reinterpret_cast<i32*>((char*)dev_stage + 256) + chunk};
  i32* hstage[2] = {reinterpret_cast<i32*>(pinned), reinterpret_cast<i32*>(pinned) + chunk};
  cudaEvent_t ev[2];
  BF_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  {
    const cudaError_t e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    if (e != cudaSuccess) {  // do not leave the first event behind
      cudaEventDestroy(ev[0]);
      set_error("bf_download_rows: cudaEventCreate: %s", cudaGetErrorString(e));
      return BF_ERR_CUDA;
    }
  }
  int rc = BF_OK;
  auto fail = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == BF_OK) {
      set_error("bf_download_rows: %s: %s", what, cudaGetErrorString(e));
      rc = BF_ERR_CUDA;
    }
  };
  fail(cudaMemsetAsync(bad, 0, sizeof(i64), st), "memset");
  const i64 nchunks = ceil_div(n, chunk);
  auto enqueue = [&](i64 k) {
    const i64 a = k * chunk, m = n - a < chunk ? n - a : chunk;
    prof_pre(PC_GATHER, st);
    narrow_rows_kernel<<<(unsigned)ceil_div(m, 256 * 4 * 4), 256, 0, st>>>(dev_src + a, m, dstage[k & 1], bad);
    fail(cudaGetLastError(), "narrow kernel");
    prof_post(PC_GATHER, st, (u64)(12 * m));
    fail(cudaMemcpyAsync(hstage[k & 1], dstage[k & 1], (size_t)m * sizeof(i32), cudaMemcpyDeviceToHost, st), "copy");
    fail(cudaEventRecord(ev[k & 1], st), "event");
  };
  WidenPool pool(n < (1 << 16) ? 1 : threads);
  enqueue(0);
  for (i64 k = 0; k < nchunks && rc == BF_OK; ++k) {
    if (k + 1 < nchunks) enqueue(k + 1);  // in flight while chunk k is widened
    fail(cudaEventSynchronize(ev[k & 1]), "wait");
    const i64 a = k * chunk, m = n - a < chunk ? n - a : chunk;
    if (rc == BF_OK) pool.run(hstage[k & 1], host_dst + a, (size_t)m);
  }
  i64 flag = 0;
  fail(cudaMemcpyAsync(&flag, bad, sizeof(i64), cudaMemcpyDeviceToHost, st), "flag");
  fail(cudaStreamSynchronize(st), "sync");
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  if (rc == BF_OK && flag) {
    set_error("bf_download_rows: a value lies outside [-1, 2^31 - 1): not a row number of a set below 2^31 rows");
    return BF_ERR_RANGE;
  }
  return rc;
}

}  // namespace bf

extern "C" {
size_t bf_download_stage_bytes(int64_t chunk_rows) {
  if (chunk_rows < 64) chunk_rows = 64;
  return 256 + 2 * (size_t)chunk_rows * sizeof(int32_t);
}
int bf_download_rows(const int64_t* dev_src, int64_t n, int64_t* host_dst, void* dev_stage, size_t dev_stage_bytes,
                     void* pinned_stage, size_t pinned_bytes, int32_t n_threads, void* stream) {
  return bf::download_rows_impl((const i64*)dev_src, n, (i64*)host_dst, dev_stage, dev_stage_bytes, pinned_stage,
                                pinned_bytes, n_threads, (cudaStream_t)stream);
}
}
