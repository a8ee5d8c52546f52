// Thread-local last-error string behind ndt1_last_error().
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[1024] = "";
thread_local long long g_ndt1_launches = 0;

void ndt1_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ndt1_last_error(void) { return g_err; }
