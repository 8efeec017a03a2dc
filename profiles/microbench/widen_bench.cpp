// Host-side cost of widening int32 row numbers to the int64 arrays the reference returns
// (the e2e path transfers 4-byte row numbers over PCIe and widens them on the host while the next chunk is in
// flight).  g++ -O3 -pthread widen_bench.cpp -o widen_bench && ./widen_bench [threads...]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

static void widen_plain(const int32_t* src, int64_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void widen_stream(const int32_t* src, int64_t* dst, size_t n) {
  size_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; ++i; }
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

int main(int argc, char** argv) {
  const size_t n = 400u * 1000 * 1000;
  int32_t* src = static_cast<int32_t*>(malloc(n * 4));
  int64_t* dst = static_cast<int64_t*>(malloc(n * 8));
  for (size_t i = 0; i < n; ++i) src[i] = (int32_t)i - 1;
  memset(dst, 0, n * 8);
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  std::vector<int> counts;
  for (int i = 1; i < argc; ++i) counts.push_back(atoi(argv[i]));
  if (counts.empty()) counts = {1, 4, 8, 16, 32, 64};
  for (int mode = 0; mode < 2; ++mode)
    for (int t : counts) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int k = 0; k < t; ++k) {
          const size_t a = n * k / t, b = n * (k + 1) / t;
          th.emplace_back([=] {
#if defined(__x86_64__)
            if (mode == 1 && __builtin_cpu_supports("avx2")) { widen_stream(src + a, dst + a, b - a); return; }
#endif
            widen_plain(src + a, dst + a, b - a);
          });
        }
        for (auto& x : th) x.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s < best) best = s;
      }
      printf("%s threads %3d: %.1f ms for %zu values -> %.1f GB/s written\n", mode ? "stream" : "plain ", t, best * 1e3, n,
             n * 8 / best / 1e9);
    }
  long long chk = 0;
  for (size_t i = 0; i < n; i += 1000003) chk += dst[i];
  printf("check %lld\n", chk);
  return 0;
}
