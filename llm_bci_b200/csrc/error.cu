// Thread-local last-error string behind ndt1_last_error().
#include "common.cuh"
#include "../../include/ndt1_b200.h"
#include <cxxabi.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>

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

int ndt1_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return dev;
}
int ndt1_num_sms() {
  static int sms[NDT1_MAX_DEVICES] = {};
  const int dev = ndt1_current_device() % NDT1_MAX_DEVICES;
  if (!sms[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    sms[dev] = n;
  }
  return sms[dev];
}

// ---------------------------------------------------------------------------
// Launch profiler (see common.cuh).  Events are pooled; nothing is allocated once the pool has grown to a step's launches.
// ---------------------------------------------------------------------------
bool g_ndt1_prof_on = false;
namespace {
struct ProfRec { cudaEvent_t e0, e1; const void* func; double flops, bytes; };
std::vector<ProfRec> g_recs;
size_t g_recs_used = 0;
std::mutex g_prof_mu;
thread_local double g_note_flops = 0, g_note_bytes = 0;
}  // namespace

void ndt1_prof_note(double flops, double bytes) { g_note_flops = flops; g_note_bytes = bytes; }

int ndt1_prof_before(const void* func, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_recs_used == g_recs.size()) {
    ProfRec r; r.func = nullptr; r.flops = r.bytes = 0;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
    g_recs.push_back(r);
  }
  const int i = (int)g_recs_used++;
  g_recs[i].func = func; g_recs[i].flops = g_note_flops; g_recs[i].bytes = g_note_bytes;
  g_note_flops = g_note_bytes = 0;
  cudaEventRecord(g_recs[i].e0, stream);
  return i;
}
void ndt1_prof_after(int rec, cudaStream_t stream) {
  if (rec < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_recs[rec].e1, stream);
}

extern "C" int ndt1_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_recs_used = 0; g_ndt1_prof_on = true;
  return 0;
}

extern "C" int ndt1_profile_end(ndt1_profile_entry* out, int capacity, int* n_out) {
  NDT1_REQUIRE(n_out && (out || capacity == 0), "profile_end: null argument");
  g_ndt1_prof_on = false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Acc { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<const void*, Acc> by_func;
  std::vector<const void*> order;
  for (size_t i = 0; i < g_recs_used; ++i) {
    NDT1_CUDA_CHECK(cudaEventSynchronize(g_recs[i].e1));
    float m = 0;
    NDT1_CUDA_CHECK(cudaEventElapsedTime(&m, g_recs[i].e0, g_recs[i].e1));
    if (!by_func.count(g_recs[i].func)) order.push_back(g_recs[i].func);
    Acc& a = by_func[g_recs[i].func];
    a.n += 1; a.ms += m; a.flops += g_recs[i].flops; a.bytes += g_recs[i].bytes;
  }
  int n = 0;
  for (const void* f : order) {
    if (n >= capacity) break;
    const Acc& a = by_func[f];
    ndt1_profile_entry& e = out[n++];
    memset(&e, 0, sizeof(e));
    const char* mangled = nullptr;
    if (cudaFuncGetName(&mangled, f) != cudaSuccess || !mangled) { cudaGetLastError(); mangled = "?"; }
    int st = 0;
    char* dem = abi::__cxa_demangle(mangled, nullptr, nullptr, &st);
    const char* nm = (st == 0 && dem) ? dem : mangled;
    // keep the kernel name and its template arguments; drop "void ", the anonymous-namespace qualifier and the parameter list
    const char* start = nm;
    if (strncmp(start, "void ", 5) == 0) start += 5;
    static const char kAnon[] = "(anonymous namespace)::";
    if (strncmp(start, kAnon, sizeof(kAnon) - 1) == 0) start += sizeof(kAnon) - 1;
    size_t len = strlen(start);
    int depth = 0;
    for (size_t k = 0; k < len; ++k) {
      if (start[k] == '<') ++depth;
      else if (start[k] == '>') --depth;
      else if (start[k] == '(' && depth == 0) { len = k; break; }
    }
    if (len >= sizeof(e.name)) len = sizeof(e.name) - 1;
    memcpy(e.name, start, len);
    free(dem);
    e.launches = a.n; e.ms = a.ms; e.flops = a.flops; e.bytes = a.bytes;
  }
  *n_out = n;
  g_recs_used = 0;
  return 0;
}
