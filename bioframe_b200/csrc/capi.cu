// Error reporting, version and the launch counter / CUDA-event profiler of the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include <vector>

#include "common.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// ---- profiler: per kernel class, launches always counted; durations only
// while enabled (two cudaEventRecord per launch on the launching stream).
struct Pending {
  int cls;
  cudaEvent_t a, b;
  u64 bytes;
};
static bool g_prof_on = false;
static std::vector<Pending> g_pending;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t g_open = nullptr;
static long long g_launches[PC_NUM];
static long long g_bytes[PC_NUM];
static double g_ms[PC_NUM];
static long long g_timed[PC_NUM];
static const char* const g_names[PC_NUM] = {"extent",  "pack_keys", "sort_pass", "sort_misc", "tie_fix", "segment_table",
                                            "running_max", "search_count", "scan", "offset_table", "emit_pairs",
                                            "unpaired", "cluster", "closest", "coverage", "gather"};

static cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void prof_pre(int cls, cudaStream_t st) {
  g_launches[cls]++;
  if (!g_prof_on) return;
  g_open = take_event();
  cudaEventRecord(g_open, st);
}
void prof_post(int cls, cudaStream_t st, u64 bytes) {
  g_bytes[cls] += (long long)bytes;
  if (!g_prof_on || !g_open) return;
  Pending p;
  p.cls = cls;
  p.a = g_open;
  p.b = take_event();
  p.bytes = bytes;
  cudaEventRecord(p.b, st);
  g_open = nullptr;
  g_pending.push_back(p);
}
static void drain() {
  for (Pending& p : g_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_ms[p.cls] += ms;
      g_timed[p.cls]++;
    }
    g_pool.push_back(p.a);
    g_pool.push_back(p.b);
  }
  g_pending.clear();
}
}  // namespace bf

extern "C" {
const char* bf_last_error(void) { return bf::g_error; }
int bf_version(void) { return 200; }
int bf_checked_build(void) {
#ifdef BF_CHECKED
  return 1;
#else
  return 0;
#endif
}

int bf_prof_enable(int on) {
  bf::drain();
  bf::g_prof_on = on != 0;
  return BF_OK;
}
int bf_prof_reset(void) {
  bf::drain();
  for (int i = 0; i < bf::PC_NUM; ++i) {
    bf::g_launches[i] = 0;
    bf::g_bytes[i] = 0;
    bf::g_ms[i] = 0.0;
    bf::g_timed[i] = 0;
  }
  return BF_OK;
}
int bf_prof_num_classes(void) { return bf::PC_NUM; }
const char* bf_prof_class_name(int cls) { return (cls >= 0 && cls < bf::PC_NUM) ? bf::g_names[cls] : ""; }
int bf_prof_read(int cls, double* ms_total, int64_t* timed_launches, int64_t* launches, int64_t* bytes) {
  if (cls < 0 || cls >= bf::PC_NUM) return BF_ERR_ARG;
  bf::drain();
  if (ms_total) *ms_total = bf::g_ms[cls];
  if (timed_launches) *timed_launches = bf::g_timed[cls];
  if (launches) *launches = bf::g_launches[cls];
  if (bytes) *bytes = bf::g_bytes[cls];
  return BF_OK;
}
}
