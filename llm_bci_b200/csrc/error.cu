// Thread-local last-error string behind ndt1_last_error().
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";
thread_local long long g_ndt1_launches = 0;

bool ndt1_pdl_enabled() {
  static const bool on = !(getenv("NDT1_PDL") && atoi(getenv("NDT1_PDL")) == 0);
  return on;
}

void ndt1_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ndt1_last_error(void) { return g_err; }
