// Shared device/host helpers for the NDT1 sm_100a kernels.
// Everything here is internal; the public surface is include/ndt1_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// Error plumbing: every C-ABI entry returns 0 on success; the message of the
// last failure on the calling thread is kept for ndt1_last_error().
// ---------------------------------------------------------------------------
void ndt1_set_error(const char* fmt, ...);

#define NDT1_CUDA_CHECK(expr)                                                   \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ndt1_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                     cudaGetErrorString(_e));                                   \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

// every kernel launch goes through this macro, so the counter is exact
extern thread_local long long g_ndt1_launches;
#define NDT1_CHECK_LAUNCH()                                                     \
  do {                                                                          \
    ++g_ndt1_launches;                                                          \
    NDT1_CUDA_CHECK(cudaGetLastError());                                        \
  } while (0)

#define NDT1_REQUIRE(cond, ...)                                                 \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ndt1_set_error(__VA_ARGS__);                                              \
      return 2;                                                                 \
    }                                                                           \
  } while (0)

#define NDT1_TRY(expr)                                                          \
  do {                                                                          \
    int _rc = (expr);                                                           \
    if (_rc != 0) return _rc;                                                   \
  } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the library starts with pdl_grid_sync() (before its first global
// access; the tensor-core GEMM after its barrier / tensor-memory set-up) and every launch goes through ndt1_launch(),
// which sets cudaLaunchAttributeProgrammaticStreamSerialization: the CTAs of kernel n+1 are scheduled -- and run their
// prologue -- on SMs that kernel n's last wave has left idle, then block in griddepcontrol.wait until kernel n has
// completed and flushed.  Safe transitively because EVERY CTA of every kernel waits before it touches global memory:
// kernel n+1 cannot complete before kernel n, so kernel n+2 (which waits on n+1) never runs ahead of n either.
// NDT1_PDL=0 in the environment turns the attribute off (plain stream order).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool ndt1_pdl_enabled();
// Measurement (bench.py `roofline` / `roofline_more`): while ndt1_profile_begin() .. ndt1_profile_end() is active every launch
// of the library is bracketed by CUDA events on ITS OWN stream and grouped by kernel; a launcher may attach the algorithmic
// FLOPs / bytes of the launch it is about to make with ndt1_prof_note (consumed by the next launch on this thread).
extern bool g_ndt1_prof_on;
int ndt1_prof_before(const void* func, cudaStream_t stream);
void ndt1_prof_after(int rec, cudaStream_t stream);
void ndt1_prof_note(double flops, double bytes);
template <typename... KArgs, typename... Args>
static inline cudaError_t ndt1_launch_cluster(void (*kernel)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (ndt1_pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  if (!g_ndt1_prof_on) return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  const int rec = ndt1_prof_before((const void*)kernel, stream);
  const cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  ndt1_prof_after(rec, stream);
  return rc;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t ndt1_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  return ndt1_launch_cluster(kernel, 1, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

static inline int ndt1_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Per-DEVICE host state (several GPUs in one process: one engine per device): the current device's ordinal and SM count, and a
// helper for "set this kernel attribute once per device" (function attributes are per device, not per process).
constexpr int NDT1_MAX_DEVICES = 64;
int ndt1_current_device();
int ndt1_num_sms();
struct Ndt1PerDeviceFlag {
  bool done[NDT1_MAX_DEVICES] = {};
  bool& here() { return done[ndt1_current_device() % NDT1_MAX_DEVICES]; }
};
struct Ndt1PerDeviceSize {
  size_t v[NDT1_MAX_DEVICES] = {};
  size_t& here() { return v[ndt1_current_device() % NDT1_MAX_DEVICES]; }
};

// ---------------------------------------------------------------------------
// Type helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_as_f32(const void* p, long long i, int is_bf16) {
  return is_bf16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void store_from_f32(void* p, long long i, int is_bf16, float v) {
  if (is_bf16) ((bf16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

// ---------------------------------------------------------------------------
// Warp / block reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// Philox4x32, counter based: (seed, counter, stream) -> 4 uniform u32.  SEVEN rounds: the fewest for which the
// Random123 authors report no BigCrush failure (their default of 10 is a safety margin); dropout masks and additive
// noise need statistical quality only, and the generator sits in the epilogue of GEMM and attention kernels.
// Forward and backward kernels index by ELEMENT, never by thread, so the mask
// is reproducible whatever the launch geometry.
// ---------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };
constexpr int kPhiloxRounds = 7;

__host__ __device__ __forceinline__ Philox4 philox4x32(uint64_t seed, uint64_t ctr, uint64_t stream) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32);
  uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
#pragma unroll
  for (int r = 0; r < kPhiloxRounds; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// The per-step Philox key, by value or read from device memory: a CUDA-graph replay of a step changes the key by updating the
// device word, without re-capturing (every stochastic kernel of the path takes a SeedRef).
struct SeedRef {
  const unsigned long long* ptr; unsigned long long val;
  __host__ __device__ SeedRef() : ptr(nullptr), val(0) {}
  __host__ __device__ SeedRef(unsigned long long v) : ptr(nullptr), val(v) {}
  __host__ __device__ static SeedRef at(const unsigned long long* p) { SeedRef r; r.ptr = p; return r; }
  __device__ __forceinline__ unsigned long long get() const { return ptr ? *ptr : val; }
};

// Dropout draws use 16 bits per element: one Philox block (4 x u32 = 8 x u16) covers the 8
// consecutive elements [8*ctr, 8*ctr+7].  An element is kept iff its u16 >= drop_threshold(p),
// i.e. P(drop) = round(p * 65536) / 65536 (|error| < 8e-6).
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 65536.0 + 0.5;
  if (t <= 0.0) return 0u;
  if (t >= 65535.0) return 65535u;
  return (uint32_t)t;
}
__device__ __forceinline__ uint32_t philox_u16(const Philox4& r, int lane) {   // lane in [0, 8)
  const uint32_t w = (lane >> 1) == 0 ? r.x : (lane >> 1) == 1 ? r.y : (lane >> 1) == 2 ? r.z : r.w;
  return (lane & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// keep-scale (0 or 1/(1-p)) of one element of a dropout site
__device__ __forceinline__ float drop_scale_1(uint64_t seed, uint64_t stream, uint64_t elem, uint32_t thr, float inv_keep) {
  const Philox4 r = philox4x32(seed, elem >> 3, stream);
  return philox_u16(r, (int)(elem & 7)) >= thr ? inv_keep : 0.f;
}
// keep-scales of the 4 consecutive elements starting at elem (elem % 4 == 0)
__device__ __forceinline__ void drop_scale_4(uint64_t seed, uint64_t stream, uint64_t elem, uint32_t thr, float ik, float* o) {
  const Philox4 r = philox4x32(seed, elem >> 3, stream);
  const uint32_t w0 = (elem & 4) ? r.z : r.x, w1 = (elem & 4) ? r.w : r.y;
  o[0] = (w0 & 0xFFFFu) >= thr ? ik : 0.f; o[1] = (w0 >> 16) >= thr ? ik : 0.f;
  o[2] = (w1 & 0xFFFFu) >= thr ? ik : 0.f; o[3] = (w1 >> 16) >= thr ? ik : 0.f;
}
// keep-scales of the 8 consecutive elements starting at elem (elem % 8 == 0)
__device__ __forceinline__ void drop_scale_8(uint64_t seed, uint64_t stream, uint64_t elem, uint32_t thr, float ik, float* o) {
  const Philox4 r = philox4x32(seed, elem >> 3, stream);
  o[0] = (r.x & 0xFFFFu) >= thr ? ik : 0.f; o[1] = (r.x >> 16) >= thr ? ik : 0.f;
  o[2] = (r.y & 0xFFFFu) >= thr ? ik : 0.f; o[3] = (r.y >> 16) >= thr ? ik : 0.f;
  o[4] = (r.z & 0xFFFFu) >= thr ? ik : 0.f; o[5] = (r.z >> 16) >= thr ? ik : 0.f;
  o[6] = (r.w & 0xFFFFu) >= thr ? ik : 0.f; o[7] = (r.w >> 16) >= thr ? ik : 0.f;
}

// Box-Muller on two u32 -> two N(0,1)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = (float)b * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  n0 = r * c; n1 = r * s;
}

// ---------------------------------------------------------------------------
// Activations (forward and derivative)
// ---------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_SOFTSIGN = 1, ACT_GELU = 2, ACT_RELU = 3 };
// derivative selectors for backward epilogues
enum { DACT_NONE = 0, DACT_SOFTSIGN_FROM_OUT = 1, DACT_GELU_FROM_IN = 2, DACT_RELU_FROM_OUT = 3,
       DACT_SAVED = 4 };     // the saved tensor IS act'(pre-activation), written by the forward epilogue (GemmEpilogue::out2_deriv)

__device__ __forceinline__ float act_apply(int act, float v) {
  switch (act) {
    case ACT_SOFTSIGN: return v / (1.0f + fabsf(v));
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    case ACT_RELU: return fmaxf(v, 0.f);
    default: return v;
  }
}
__device__ __forceinline__ float dact_apply(int dact, float saved) {
  switch (dact) {
    case DACT_SOFTSIGN_FROM_OUT: { float t = 1.0f - fabsf(saved); return t * t; }
    case DACT_GELU_FROM_IN: {
      const float cdf = 0.5f * (1.0f + erff(saved * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * expf(-0.5f * saved * saved);
      return cdf + saved * pdf;
    }
    case DACT_RELU_FROM_OUT: return saved > 0.f ? 1.f : 0.f;
    case DACT_SAVED: return saved;
    default: return 1.f;
  }
}
