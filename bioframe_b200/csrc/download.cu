// bf_download_rows: bring an int64 device array of row numbers (the events1 / events2 of
// ops._overlap_intidxs, ops.py:242-338, or any other result whose values lie in [-1, 2^31)) into an int64 host
// array.  A call through host buffers is bound by PCIe (profiles/NOTES.md: 14 GB of pairs at ~55 GB/s), and
// row numbers of sets below 2^31 rows carry 4 bytes of information: they cross the bus as int32 and are
// widened on the host by a few threads with streaming stores while the next chunk is in flight.
//
//   per chunk:  narrow kernel (int64 -> int32, checks the range)  ->  cudaMemcpyAsync to pinned staging  ->
//               event;  host: wait for the previous chunk's event, widen it into the destination
//   two device and two pinned staging chunks: the bus carries chunk k + 1 while the host widens chunk k.
//
// Hosts differ: on the boxes of the pool the widening of a chunk takes between 1.1x and 5x the time of its copy
// (profiles/NOTES.md: 191 ... 699 ms for the same 1.75e9 row numbers).  When the destination is page-locked the
// call therefore measures both on its first chunks and, if the host threads are the slower side, hands the TAIL of
// the array to a second lane: plain int64 copies straight into the destination on a second stream (8 bytes per
// row over the bus, no host work), sized so that both lanes end together.
#include <atomic>
#include <chrono>
#include <cstdlib>
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

// BF_DOWNLOAD_DIRECT=0: never use the direct lane; =1: always split (tests); unset: decide from the measured times
enum DirectMode { DIRECT_ADAPTIVE, DIRECT_OFF, DIRECT_FORCE };
static DirectMode direct_mode() {
  const char* e = std::getenv("BF_DOWNLOAD_DIRECT");
  if (!e || !*e) return DIRECT_ADAPTIVE;
  return e[0] == '0' ? DIRECT_OFF : e[0] == '1' ? DIRECT_FORCE : DIRECT_ADAPTIVE;
}

static bool page_locked(const void* p, size_t bytes) {
  cudaPointerAttributes first{}, last{};
  const bool ok = cudaPointerGetAttributes(&first, p) == cudaSuccess &&
                  cudaPointerGetAttributes(&last, static_cast<const char*>(p) + bytes - 1) == cudaSuccess;
  if (!ok) cudaGetLastError();  // an unknown pointer is reported as an error by older drivers: not sticky
  return ok && first.type == cudaMemoryTypeHost && last.type == cudaMemoryTypeHost;
}

static thread_local i64 g_direct_rows = 0;  // rows of this thread's last call that took the direct lane

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
  // ev[]: chunk landed in the pinned stage; tb[]: recorded in front of the chunk's kernel, so that
  // elapsed(tb, ev) = what the chunk costs on the device + bus side
  cudaEvent_t ev[2] = {nullptr, nullptr}, tb[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaEventCreate(&ev[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&tb[i]);
    if (e != cudaSuccess) {  // leave no event behind
      for (int j = 0; j < 2; ++j) {
        if (ev[j]) cudaEventDestroy(ev[j]);
        if (tb[j]) cudaEventDestroy(tb[j]);
      }
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
  // Rows [0, next) have been handed to the narrow lane, rows [tail, n) to the direct lane; the lanes meet somewhere.
  i64 next = 0, tail = n;
  i64 first_row[2] = {0, 0}, rows_in[2] = {0, 0};
  auto enqueue = [&](int slot) {
    const i64 a = next, m = tail - a < chunk ? tail - a : chunk;
    first_row[slot] = a;
    rows_in[slot] = m;
    next = a + m;
    fail(cudaEventRecord(tb[slot], st), "event");
    prof_pre(PC_GATHER, st);
    narrow_rows_kernel<<<(unsigned)ceil_div(m, 256 * 4 * 4), 256, 0, st>>>(dev_src + a, m, dstage[slot], bad);
    fail(cudaGetLastError(), "narrow kernel");
    prof_post(PC_GATHER, st, (u64)(12 * m));
    fail(cudaMemcpyAsync(hstage[slot], dstage[slot], (size_t)m * sizeof(i32), cudaMemcpyDeviceToHost, st), "copy");
    fail(cudaEventRecord(ev[slot], st), "event");
  };
  // The direct lane: needs a page-locked destination and enough chunks to measure on.  Per narrow chunk the host
  // needs widen_ms and the bus copy_ms; what the narrow lane leaves of the bus carries direct slices of `chunk` rows
  // at 8 instead of 4 bytes per row: (widen / copy - 1) / 2 slices per narrow chunk.  The slices are handed out a
  // few per chunk, not all at once, so that the lanes share the bus whichever way the copy engines order them.
  const DirectMode mode = direct_mode();
  constexpr int kFirstSample = 1, kSamples = 3, kMinChunks = 8;
  bool undecided = mode != DIRECT_OFF && ceil_div(n, chunk) >= kMinChunks && page_locked(host_dst, (size_t)n * sizeof(i64));
  cudaStream_t st2 = nullptr;
  double widen_ms = 0, copy_ms = 0, quota = 0;
  auto claim_direct = [&](double slices) {
    quota += slices;
    while (quota >= 1.0 && tail > next && rc == BF_OK) {
      const i64 lo = tail - chunk > next ? tail - chunk : next;
      fail(cudaMemcpyAsync(host_dst + lo, dev_src + lo, (size_t)(tail - lo) * sizeof(i64), cudaMemcpyDeviceToHost, st2),
           "direct copy");
      tail = lo;
      quota -= 1.0;
    }
  };

  WidenPool pool(n < (1 << 16) ? 1 : threads);
  enqueue(0);
  for (i64 k = 0; rc == BF_OK; ++k) {
    const int slot = (int)(k & 1);
    const bool more = next < tail;
    if (more) enqueue(slot ^ 1);  // in flight while chunk k is widened
    fail(cudaEventSynchronize(ev[slot]), "wait");
    if (rc != BF_OK) break;
    float on_device = 0.f;
    if (undecided && k >= kFirstSample) fail(cudaEventElapsedTime(&on_device, tb[slot], ev[slot]), "elapsed");
    const auto w0 = std::chrono::steady_clock::now();
    pool.run(hstage[slot], host_dst + first_row[slot], (size_t)rows_in[slot]);
    const double widened = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    if (st2) {
      // the share of the bus the narrow lane leaves, from its latest widening time
      const double spare = copy_ms > 0 ? 0.5 * (widened / copy_ms - 1.0) : 1.0;
      claim_direct(spare < 0 ? 0.0 : spare > 8 ? 8.0 : spare);
    } else if (undecided && k >= kFirstSample && rc == BF_OK) {
      widen_ms += widened;
      copy_ms += on_device;
      if (k + 1 == kFirstSample + kSamples) {
        undecided = false;
        widen_ms /= kSamples;
        copy_ms /= kSamples;
        // dev_src is complete: this thread has synchronised with events recorded on `st` behind its producer
        if (mode == DIRECT_FORCE || widen_ms > 1.25 * copy_ms) {
          fail(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking), "stream");
          if (mode == DIRECT_FORCE && widen_ms <= 1.25 * copy_ms) copy_ms = 0;  // tests: one slice per chunk
        }
      }
    }
    if (!more) break;
  }
  if (st2) {
    fail(cudaStreamSynchronize(st2), "direct lane");
    cudaStreamDestroy(st2);
  }
  g_direct_rows = n - tail;
  i64 flag = 0;
  fail(cudaMemcpyAsync(&flag, bad, sizeof(i64), cudaMemcpyDeviceToHost, st), "flag");
  fail(cudaStreamSynchronize(st), "sync");
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(ev[i]);
    cudaEventDestroy(tb[i]);
  }
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
int64_t bf_download_direct_rows(void) { return bf::g_direct_rows; }
int bf_download_rows(const int64_t* dev_src, int64_t n, int64_t* host_dst, void* dev_stage, size_t dev_stage_bytes,
                     void* pinned_stage, size_t pinned_bytes, int32_t n_threads, void* stream) {
  return bf::download_rows_impl((const i64*)dev_src, n, (i64*)host_dst, dev_stage, dev_stage_bytes, pinned_stage,
                                pinned_bytes, n_threads, (cudaStream_t)stream);
}
}
