// Error reporting and version entry points of the C ABI.
#include <stdarg.h>
#include <stdio.h>

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
}  // namespace bf

extern "C" {
const char* bf_last_error(void) { return bf::g_error; }
int bf_version(void) { return 100; }
}
