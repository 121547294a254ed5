// C-ABI plumbing shared by all translation units: version + thread-local error text.
#include <stdarg.h>

#include "dj_common.cuh"

static thread_local char g_dj_error[512] = "";

void dj_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_dj_error, sizeof(g_dj_error), fmt, ap);
  va_end(ap);
}

extern "C" int dj_version(void) { return 100; }
extern "C" const char* dj_last_error(void) { return g_dj_error; }
