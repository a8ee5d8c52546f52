// HBM-bound kernels of the NDT1 path: smoothing/noise prologue, masker,
// device-side pad/pack collate, casts, bias-gradient column sums, positional
// embedding gradient scatter, Poisson-NLL / MSE masked losses, AdamW.
// All are coalesced along the innermost (channel / hidden) dimension.
#include "kernels.cuh"

// ===========================================================================
// Prologue: depthwise Gaussian smoothing along time + training noise.
// Reference: SmoothAndNoise.forward, models/ndt1.py:92-107.
//   out[b,t,n] = sum_i w[i] * x[b, t + i - (K-1)/2, n]   (zero "same" padding)
//              + white_sd * N(0,1)[b,t,n] + offset_sd * N(0,1)[b,n]
// Each thread produces TT consecutive time bins of one channel so the K-tap
// window slides through registers (≈(TT+K-1)/TT loads per output).
// ===========================================================================
namespace {
constexpr int SM_TT = 8;
constexpr int SM_MAXK = 64;

struct SmoothParams {
  const float* x; float* out;            // out: fp32 result or null
  bf16* out_bf16; int ld_bf16;           // and / or the bf16 GEMM operand of the channel embedding, written directly (row stride ld_bf16)
  int B, T, N, K;
  float w[SM_MAXK];
  float white_sd, offset_sd;
  const float* white; const float* offset;   // injected draws or null
  int use_philox; SeedRef seed;
};

// White noise: ONE Philox block per four consecutive elements (all four words used): element e takes lane e % 2 of the Box-Muller
// pair formed from words (x, y) for e % 4 < 2 and (z, w) otherwise.
__device__ __forceinline__ float white_one(unsigned long long seed, unsigned long long e) {
  const Philox4 r = philox4x32(seed, e >> 2, 0x77686974ULL);
  float a, c;
  if (e & 2) box_muller(r.z, r.w, a, c); else box_muller(r.x, r.y, a, c);
  return (e & 1) ? c : a;
}
__global__ void smooth_noise_kernel(const SmoothParams p) { pdl_grid_sync();
  const unsigned long long seed = p.use_philox ? p.seed.get() : 0ull;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = blockIdx.y * SM_TT;
  const int b = blockIdx.z;
  if (n >= p.N) return;
  const float* xb = p.x + (long long)b * p.T * p.N + n;
  float* ob = p.out ? p.out + (long long)b * p.T * p.N + n : nullptr;
  bf16* ob16 = p.out_bf16 ? p.out_bf16 + (long long)b * p.T * p.ld_bf16 + n : nullptr;
  float off = 0.f;
  if (p.offset_sd != 0.f) {
    if (p.offset) off = p.offset_sd * p.offset[(long long)b * p.N + n];
    else if (p.use_philox) {
      const unsigned long long e = (unsigned long long)b * p.N + n;
      Philox4 r = philox4x32(seed, e, 0x6f666673ULL);
      float a, c; box_muller(r.x, r.y, a, c);
      off = p.offset_sd * a;
    }
  }
  if (p.K > 0) {
    const int half = (p.K - 1) / 2;
    float win[SM_TT + SM_MAXK];
    for (int i = 0; i < SM_TT + p.K - 1; ++i) {
      const int t = t0 + i - half;
      win[i] = (t >= 0 && t < p.T) ? xb[(long long)t * p.N] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < SM_TT; ++j) {
      const int t = t0 + j;
      if (t >= p.T) break;
      float acc = 0.f;
      for (int i = 0; i < p.K; ++i) acc = fmaf(p.w[i], win[j + i], acc);
      float v = acc + off;
      if (p.white_sd != 0.f) {
        const unsigned long long e = ((unsigned long long)b * p.T + t) * p.N + n;
        if (p.white) v += p.white_sd * p.white[e];
        else if (p.use_philox) v += p.white_sd * white_one(seed, e);
      }
      if (ob) ob[(long long)t * p.N] = v;
      if (ob16) ob16[(long long)t * p.ld_bf16] = __float2bfloat16_rn(v);
    }
  } else {
    for (int j = 0; j < SM_TT; ++j) {
      const int t = t0 + j;
      if (t >= p.T) break;
      float v = xb[(long long)t * p.N] + off;
      if (p.white_sd != 0.f) {
        const unsigned long long e = ((unsigned long long)b * p.T + t) * p.N + n;
        if (p.white) v += p.white_sd * p.white[e];
        else if (p.use_philox) v += p.white_sd * white_one(seed, e);
      }
      if (ob) ob[(long long)t * p.N] = v;
      if (ob16) ob16[(long long)t * p.ld_bf16] = __float2bfloat16_rn(v);
    }
  }
}
// Fast path (N % 4 == 0, compile-time tap count): one thread owns FOUR adjacent channels (128-bit loads/stores,
// a warp reads 512 contiguous bytes) and a strip of SMV_TT bins, sliding a K-deep register window along time; one
// Philox call per bin yields the four N(0,1) draws of its four channels (same element -> draw mapping as the
// generic kernel: counter = element / 2 for consecutive element pairs... see white_pair()).
constexpr int SMV_TT = 13;     // one trip of the K = 13 window per thread: 77 strips per 1000-bin trial -> 4+ CTAs per SM at the benchmark shape
__device__ __forceinline__ void white_quad(unsigned long long seed, unsigned long long e0, float* nz) {
  // elements e0..e0+3 (e0 % 4 == 0): one Philox block, the same element -> draw mapping as white_one()
  const Philox4 r = philox4x32(seed, e0 >> 2, 0x77686974ULL);
  box_muller(r.x, r.y, nz[0], nz[1]);
  box_muller(r.z, r.w, nz[2], nz[3]);
}
// Two adjacent channels per thread (64-bit loads / stores, a warp reads 256 contiguous bytes of a bin) and a strip of SMV_TT
// bins: 2 x K window registers per thread keep the kernel at three 256-thread CTAs per SM (the four-channel version held 104
// window registers and ran ONE CTA per SM, i.e. latency-bound at 20 % of the HBM rate).
template <int K>
__global__ void __launch_bounds__(256, 3) smooth_noise_vec_kernel(const SmoothParams p) { pdl_grid_sync();
  const unsigned long long seed = p.use_philox ? p.seed.get() : 0ull;
  const int n2 = blockIdx.x * 128 + (threadIdx.x & 127);        // channel pair
  const int strip = blockIdx.y * 2 + (threadIdx.x >> 7);
  const int b = blockIdx.z;
  const int t0 = strip * SMV_TT;
  if (n2 * 2 >= p.N || t0 >= p.T) return;
  constexpr int half = (K - 1) / 2;
  const float2* xb = (const float2*)(p.x + (long long)b * p.T * p.N) + n2;
  float2* ob = p.out ? (float2*)(p.out + (long long)b * p.T * p.N) + n2 : nullptr;
  bf16* ob16 = p.out_bf16 ? p.out_bf16 + (long long)b * p.T * p.ld_bf16 + n2 * 2 : nullptr;
  const int ld2 = p.N / 2;
  float2 off = make_float2(0.f, 0.f);
  if (p.offset_sd != 0.f) {
    float o[2] = {0.f, 0.f};
    for (int c = 0; c < 2; ++c) {
      const unsigned long long e = (unsigned long long)b * p.N + n2 * 2 + c;
      if (p.offset) o[c] = p.offset_sd * p.offset[e];
      else if (p.use_philox) {
        Philox4 r = philox4x32(seed, e, 0x6f666673ULL);
        float a, d; box_muller(r.x, r.y, a, d);
        o[c] = p.offset_sd * a;
      }
    }
    off = make_float2(o[0], o[1]);
  }
  float w[K];
#pragma unroll
  for (int i = 0; i < K; ++i) w[i] = p.w[i];
  float2 win[K];                                   // win[i] = x[t + i - half]
#pragma unroll
  for (int i = 0; i < K - 1; ++i) {
    const int t = t0 + i - half;
    win[i + 1] = (t >= 0 && t < p.T) ? __ldg(xb + (long long)t * ld2) : make_float2(0.f, 0.f);
  }
  float2 nxt[K];                                   // the K rows this strip appends: all loads issued before the first use
#pragma unroll
  for (int jj = 0; jj < K; ++jj) {
    const int tn = t0 + jj + half;
    nxt[jj] = (tn < p.T && jj < SMV_TT) ? __ldg(xb + (long long)tn * ld2) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int jj = 0; jj < SMV_TT; ++jj) {
    const int t = t0 + jj;
#pragma unroll
    for (int i = 0; i < K - 1; ++i) win[i] = win[i + 1];       // (compile-time rotation: the loop is fully unrolled)
    win[K - 1] = nxt[jj];
    if (t < p.T) {
      float2 acc = off;
#pragma unroll
      for (int i = 0; i < K; ++i) { acc.x = fmaf(w[i], win[i].x, acc.x); acc.y = fmaf(w[i], win[i].y, acc.y); }
      if (p.white_sd != 0.f) {
        const unsigned long long e = ((unsigned long long)b * p.T + t) * p.N + n2 * 2;
        if (p.white) {
          const float2 wn = __ldg((const float2*)(p.white + e));
          acc.x = fmaf(p.white_sd, wn.x, acc.x); acc.y = fmaf(p.white_sd, wn.y, acc.y);
        } else if (p.use_philox) {           // this pair's half of the Philox block of its four-element group (white_one's mapping)
          const Philox4 r = philox4x32(seed, e >> 2, 0x77686974ULL);
          float z0, z1;
          if (e & 2) box_muller(r.z, r.w, z0, z1); else box_muller(r.x, r.y, z0, z1);
          acc.x = fmaf(p.white_sd, z0, acc.x); acc.y = fmaf(p.white_sd, z1, acc.y);
        }
      }
      if (ob) ob[(long long)t * ld2] = acc;
      if (ob16) { const __nv_bfloat162 o2 = __floats2bfloat162_rn(acc.x, acc.y); *(uint32_t*)(ob16 + (long long)t * p.ld_bf16) = *(const uint32_t*)&o2; }
    }
  }
}
}  // namespace

int k_smooth_noise(const float* x, float* out, int B, int T, int N, const float* taps, int K, float white_sd, float offset_sd,
                   const float* white, const float* offset, int use_philox, SeedRef seed, cudaStream_t stream, bf16* out_bf16, int ld_bf16) {
  if ((long long)B * T * N == 0) return 0;
  NDT1_REQUIRE(out || out_bf16, "smooth: no output buffer");
  NDT1_REQUIRE(!out_bf16 || ld_bf16 >= N, "smooth: bf16 row stride %d < %d channels", ld_bf16, N);
  NDT1_REQUIRE(K >= 0 && K <= SM_MAXK - 1, "smooth: %d taps unsupported (max %d)", K, SM_MAXK - 1);
  NDT1_REQUIRE(K == 0 || (K % 2) == 1, "smooth: kernel length must be odd ('same' padding), got %d", K);
  if (B * T * N == 0) return 0;
  SmoothParams p;
  p.x = x; p.out = out; p.out_bf16 = out_bf16; p.ld_bf16 = ld_bf16; p.B = B; p.T = T; p.N = N; p.K = K;
  for (int i = 0; i < K; ++i) p.w[i] = taps[i];
  p.white_sd = white_sd; p.offset_sd = offset_sd; p.white = white; p.offset = offset;
  p.use_philox = use_philox; p.seed = seed;
  if (K == 13 && N % 2 == 0 && ((uintptr_t)x & 7) == 0 && ((uintptr_t)out & 7) == 0 && (!white || ((uintptr_t)white & 7) == 0) &&
      (!out_bf16 || (((uintptr_t)out_bf16 & 3) == 0 && ld_bf16 % 2 == 0))) {
    dim3 grid(ndt1_cdiv(N / 2, 128), ndt1_cdiv(ndt1_cdiv(T, SMV_TT), 2), B);     // the reference's default: gaussian(1 + 6 sd, sd = 2)
    if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)B * T * N * (4 + (out ? 4 : 0) + (out_bf16 ? 2 : 0)));     // fp32 in; fp32 and / or bf16 out
    ndt1_launch(smooth_noise_vec_kernel<13>, grid, 256, 0, stream, p);
    NDT1_CHECK_LAUNCH();
    return 0;
  }
  dim3 block(128), grid(ndt1_cdiv(N, 128), ndt1_cdiv(T, SM_TT), B);
  ndt1_launch(smooth_noise_kernel, grid, block, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ===========================================================================
// Masker (models/masker.py:44-104), bit exact given the draws.
//  pass 1: mask (mode broadcast + temporal dilation), zero replacement,
//          global max of the tensor AFTER zeroing (masker.py:100-102)
//  pass 2: random replacement max*rand, write mask (int64) and OR into targets_mask
// ===========================================================================
namespace {

struct MaskerParams {
  float* spikes;                 // (B,T,N) in place
  int B, T, N;
  int mode;                      // 0 temporal (B,T) 1 neuron/region (B,N) 2 random (B,T,N) 3 co-smooth (N)
  int timespan;
  const unsigned char* mask_draw; const unsigned char* zero_draw; const unsigned char* random_draw;
  const float* rand;
  long long* mask_out;           // (B,T,N) int64 or null
  long long* targets_mask;       // (B,T,N) int64, OR-ed, or null
  unsigned int* max_bits;        // ordered-int encoding of the running max
};

__device__ __forceinline__ unsigned int f32_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ bool masker_mask_at(const MaskerParams& p, int b, int t, int n) {
  switch (p.mode) {
    case 0: {
      if (p.timespan <= 1) return p.mask_draw[(long long)b * p.T + t] != 0;
      const int left = (p.timespan - 1) / 2;
      bool m = false;
      for (int k = 0; k < p.timespan; ++k) {
        const int tt = t - left + k;
        if (tt >= 0 && tt < p.T) m |= p.mask_draw[(long long)b * p.T + tt] != 0;
      }
      return m;
    }
    case 1: return p.mask_draw[(long long)b * p.N + n] != 0;
    case 2: return p.mask_draw[((long long)b * p.T + t) * p.N + n] != 0;
    default: return p.mask_draw[n] != 0;
  }
}

__global__ void masker_pass1(const MaskerParams p) { pdl_grid_sync();
  const long long total = (long long)p.B * p.T * p.N;
  float lmax = -INFINITY;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % p.N);
    const long long bt = e / p.N;
    const int t = (int)(bt % p.T), b = (int)(bt / p.T);
    float v = p.spikes[e];
    if (masker_mask_at(p, b, t, n) && p.zero_draw[e]) { v = 0.f; p.spikes[e] = 0.f; }
    lmax = fmaxf(lmax, v);
  }
  lmax = warp_max(lmax);
  __shared__ float sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = lmax;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : -INFINITY;
    v = warp_max(v);
    if (threadIdx.x == 0) atomicMax(p.max_bits, f32_to_ordered(v));
  }
}

__global__ void masker_pass2(const MaskerParams p) { pdl_grid_sync();
  const long long total = (long long)p.B * p.T * p.N;
  const float mx = ordered_to_f32(*p.max_bits);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % p.N);
    const long long bt = e / p.N;
    const int t = (int)(bt % p.T), b = (int)(bt / p.T);
    const bool m = masker_mask_at(p, b, t, n);
    if (m && !p.zero_draw[e] && p.random_draw[e]) p.spikes[e] = mx * p.rand[e];
    if (p.mask_out) p.mask_out[e] = m ? 1 : 0;
    if (p.targets_mask && m) p.targets_mask[e] = 1;
  }
}
}  // namespace

int k_masker_apply(float* spikes, int B, int T, int N, int mode, int timespan, const unsigned char* mask_draw,
                   const unsigned char* zero_draw, const unsigned char* random_draw, const float* rand, long long* mask_out,
                   long long* targets_mask, unsigned int* scratch, cudaStream_t stream) {
  NDT1_REQUIRE(mode >= 0 && mode <= 3, "masker: unknown mode %d", mode);
  NDT1_REQUIRE(timespan >= 1, "masker: timespan must be >= 1");
  const long long total = (long long)B * T * N;
  if (total == 0) return 0;
  MaskerParams p;
  p.spikes = spikes; p.B = B; p.T = T; p.N = N; p.mode = mode; p.timespan = timespan;
  p.mask_draw = mask_draw; p.zero_draw = zero_draw; p.random_draw = random_draw; p.rand = rand;
  p.mask_out = mask_out; p.targets_mask = targets_mask; p.max_bits = scratch;
  NDT1_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(unsigned int), stream));  // ordered(0) < ordered(-inf)
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  ndt1_launch(masker_pass1, blocks, 256, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  ndt1_launch(masker_pass2, blocks, 256, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// Device Bernoulli / uniform draws for the masker's fast path (own Philox stream).
namespace {
__global__ void bernoulli_u8_kernel(unsigned char* out, long long n, float prob, unsigned long long seed, unsigned long long stream) { pdl_grid_sync();
  const double tt = (double)prob * 4294967296.0;
  const uint32_t thr = tt <= 0.0 ? 0u : (tt >= 4294967295.0 ? 4294967295u : (uint32_t)tt);  // P(u < thr) = prob
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += (long long)gridDim.x * blockDim.x) {
    Philox4 r = philox4x32(seed, i, stream);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
    for (int k = 0; k < 4; ++k)
      if (i * 4 + k < n) out[i * 4 + k] = (prob >= 1.f) ? 1 : (u[k] < thr ? 1 : 0);
  }
}
__global__ void uniform_f32_kernel(float* out, long long n, unsigned long long seed, unsigned long long stream) { pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += (long long)gridDim.x * blockDim.x) {
    Philox4 r = philox4x32(seed, i, stream);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
    for (int k = 0; k < 4; ++k)
      if (i * 4 + k < n) out[i * 4 + k] = (float)(u[k] >> 8) * (1.0f / 16777216.0f);
  }
}
}  // namespace

int k_bernoulli_u8(unsigned char* out, long long n, float prob, unsigned long long seed, unsigned long long stream_id, cudaStream_t stream) {
  if (n == 0) return 0;
  const int blocks = (int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 + 1 : 148 * 8);
  ndt1_launch(bernoulli_u8_kernel, blocks, 256, 0, stream, out, n, prob, seed, stream_id);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int k_uniform_f32(float* out, long long n, unsigned long long seed, unsigned long long stream_id, cudaStream_t stream) {
  if (n == 0) return 0;
  const int blocks = (int)((n / 4 + 255) / 256 < 148 * 8 ? (n / 4 + 255) / 256 + 1 : 148 * 8);
  ndt1_launch(uniform_f32_kernel, blocks, 256, 0, stream, out, n, seed, stream_id);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ===========================================================================
// Device pad/pack collate (data_utils/datasets.py:191-221 semantics):
// rows of a ragged batch, stored back to back, are scattered into a
// (B, P, inner) tensor padded with `value` on `side`, then truncated to the
// FIRST `P` entries of the padded row (truncate keeps [0:truncate]).
//   elem_size in bytes (4 or 8); rows given by offsets[b] (in rows of `inner`)
// ===========================================================================
namespace {
template <typename E>
__global__ void pad_pack_kernel(const E* src, const long long* offsets, E* dst, int B, int P, int inner, int side_left, int full, E value) { pdl_grid_sync();
  // full = padded length before truncation (max(max_len, min_length)); P = min(truncate, full)
  const long long total = (long long)B * P * inner;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % inner);
    const long long bp = e / inner;
    const int pos = (int)(bp % P), b = (int)(bp / P);
    const long long r0 = offsets[b];
    const int len = (int)(offsets[b + 1] - r0);
    const int padn = full > len ? full - len : 0;
    const int srow = side_left ? pos - padn : pos;
    E v = value;
    if (srow >= 0 && srow < len) v = src[(r0 + srow) * inner + c];
    dst[e] = v;
  }
}
}  // namespace

int k_pad_pack(const void* src, const long long* offsets, void* dst, int B, int P, int inner, int elem_size, int side_left, int full,
               double value, cudaStream_t stream) {
  const long long total = (long long)B * P * inner;
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  if (elem_size == 4)
    ndt1_launch(pad_pack_kernel<float>, blocks, 256, 0, stream, (const float*)src, offsets, (float*)dst, B, P, inner, side_left, full, (float)value);
  else if (elem_size == 8)
    ndt1_launch(pad_pack_kernel<long long>, blocks, 256, 0, stream, (const long long*)src, offsets, (long long*)dst, B, P, inner, side_left, full,
                                                           (long long)value);
  else
    NDT1_REQUIRE(false, "pad_pack: element size %d unsupported (4 = float32, 8 = int64)", elem_size);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ===========================================================================
// Casts and small utilities
// ===========================================================================
namespace {
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long rows, int cols, long long ld_in, long long ld_out) { pdl_grid_sync();
  const long long total4 = rows * (cols / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (cols / 4);
    const int c = (int)(i % (cols / 4)) * 4;
    const float4 v = *(const float4*)(in + r * ld_in + c);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b;
    *(uint2*)(out + r * ld_out + c) = o;
  }
}
__global__ void cast_f32_bf16_scalar_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long rows, int cols, long long ld_in, long long ld_out) { pdl_grid_sync();
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols; const int c = (int)(i % cols);
    out[r * ld_out + c] = __float2bfloat16_rn(in[r * ld_in + c]);
  }
}
}  // namespace

int k_cast_f32_bf16(const float* in, bf16* out, long long rows, int cols, long long ld_in, long long ld_out, cudaStream_t stream) {
  if (rows * cols == 0) return 0;
  const bool vec = (cols % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)out & 7) == 0);
  const long long work = vec ? rows * (cols / 4) : rows * cols;
  const int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  if (vec) ndt1_launch(cast_f32_bf16_kernel, blocks, 256, 0, stream, in, out, rows, cols, ld_in, ld_out);
  else ndt1_launch(cast_f32_bf16_scalar_kernel, blocks, 256, 0, stream, in, out, rows, cols, ld_in, ld_out);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// All the bf16 weight copies of a forward in ONE launch: up to CAST_MAX_SEGS contiguous fp32 -> bf16 segments
// (element counts multiples of 8), flattened into 2048-element chunks; a block finds its segment by a short scan.
namespace {
__global__ void __launch_bounds__(256) cast_multi_kernel(const CastSegs segs) { pdl_grid_sync();
  for (long long chunk = blockIdx.x; chunk < segs.total_chunks; chunk += gridDim.x) {
    int k = 0;
    while (k + 1 < segs.n && chunk >= segs.first_chunk[k + 1]) ++k;
    const long long e0 = (chunk - segs.first_chunk[k]) * 2048 + (long long)threadIdx.x * 8;
    if (e0 < segs.count[k]) {
      const float4 a = *(const float4*)(segs.src[k] + e0), b = *(const float4*)(segs.src[k] + e0 + 4);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      uint4 o; o.x = *(uint32_t*)&p0; o.y = *(uint32_t*)&p1; o.z = *(uint32_t*)&p2; o.w = *(uint32_t*)&p3;
      *(uint4*)(segs.dst[k] + e0) = o;
    }
  }
}
}  // namespace

int k_cast_multi(CastSegs& segs, cudaStream_t stream) {
  if (segs.n == 0) return 0;
  long long chunks = 0;
  for (int i = 0; i < segs.n; ++i) {
    NDT1_REQUIRE(segs.count[i] % 8 == 0 && ((uintptr_t)segs.src[i] & 15) == 0 && ((uintptr_t)segs.dst[i] & 15) == 0,
                 "cast_multi: segment %d is not 8-element / 16-byte aligned", i);
    segs.first_chunk[i] = chunks;
    chunks += (segs.count[i] + 2047) / 2048;
  }
  segs.total_chunks = chunks;
  const int blocks = (int)(chunks < 148 * 16 ? chunks : 148 * 16);
  ndt1_launch(cast_multi_kernel, blocks, 256, 0, stream, segs);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ===========================================================================
// Column sums (bias gradients): out[c] += sum_r in[r, c]
// grid (cols/32, row-chunks); 32x8 threads; each warp-row strides the rows.
// ===========================================================================
namespace {
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ in, float* __restrict__ out, long long rows, int cols, long long ld) { pdl_grid_sync();
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols) {
    const long long per = (rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f32(in[r * ld + c]);
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][threadIdx.x];
    atomicAdd(out + c, s);
  }
}
}  // namespace

namespace {
// 8 columns per thread (one 16-byte load of bf16, two of fp32); grid (cols/256, row-chunks); 32x8 threads
template <typename T>
__global__ void colsum8_kernel(const T* __restrict__ in, float* __restrict__ out, long long rows, int cols, long long ld) { pdl_grid_sync();
  __shared__ float sm[8][32][9];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
    const long long per = (rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      if (sizeof(T) == 2) {
        const uint4 raw = *(const uint4*)((const bf16*)in + r * ld + c);
        const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * i] += __low2float(h[i]); acc[2 * i + 1] += __high2float(h[i]); }
      } else {
        const float4 a = *(const float4*)((const float*)in + r * ld + c), b = *(const float4*)((const float*)in + r * ld + c + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;        // 256 threads <-> 256 columns of the block
  const int cc = blockIdx.x * 256 + t;
  if (cc < cols) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) v += sm[y][t >> 3][t & 7];
    atomicAdd(out + cc, v);
  }
}
}  // namespace

template <typename T>
int k_colsum(const T* in, float* out, long long rows, int cols, long long ld, cudaStream_t stream) {
  if (rows * cols == 0) return 0;
  dim3 block(32, 8);
  if (cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)in & 15) == 0) {
    int chunks8 = (int)(rows / 128); if (chunks8 < 1) chunks8 = 1;
    const int gx = ndt1_cdiv(cols, 256);
    while (chunks8 > 1 && (long long)chunks8 * gx > 148 * 4) --chunks8;
    if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)rows * cols * sizeof(T));
    ndt1_launch(colsum8_kernel<T>, dim3(gx, chunks8), block, 0, stream, in, out, rows, cols, ld);
    NDT1_CHECK_LAUNCH();
    return 0;
  }
  int chunks = (int)(rows / 256); if (chunks < 1) chunks = 1; if (chunks > 64) chunks = 64;
  dim3 grid(ndt1_cdiv(cols, 32), chunks);
  ndt1_launch(colsum_kernel<T>, grid, block, 0, stream, in, out, rows, cols, ld);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_colsum<float>(const float*, float*, long long, int, long long, cudaStream_t);
template int k_colsum<bf16>(const bf16*, float*, long long, int, long long, cudaStream_t);

// ===========================================================================
// Gradient hand-off between the fp32 residual stream and the next GEMM:
//   out[r,c] = T( g[r,c] * dropscale(site, r*cols + c) )
// and (optionally) the positional-table scatter  dtab[idx[r], c] += same value.
// ===========================================================================
namespace {
template <typename T>
__global__ void grad_prep_kernel(const float* __restrict__ g, T* __restrict__ out, long long rows, int cols, float drop_p,
                                 SeedRef seed_ref, unsigned long long stream_id, float* dtab, const long long* idx, int tab_ld,
                                 int rows_per_b, long long idx_stride, int prefix) { pdl_grid_sync();
  const unsigned long long seed = drop_p > 0.f ? seed_ref.get() : 0ull;
  const long long total4 = rows * (cols / 4);
  const uint32_t thr = drop_threshold(drop_p);
  const float ik = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (cols / 4);
    const int c = (int)(i % (cols / 4)) * 4;
    float4 v = *(const float4*)(g + r * cols + c);
    if (drop_p > 0.f) {
      float ds[4];
      drop_scale_4(seed, stream_id, (unsigned long long)(r * cols + c), thr, ik, ds);
      v.x *= ds[0]; v.y *= ds[1]; v.z *= ds[2]; v.w *= ds[3];
    }
    if (out) {
      out[r * cols + c + 0] = from_f32<T>(v.x); out[r * cols + c + 1] = from_f32<T>(v.y);
      out[r * cols + c + 2] = from_f32<T>(v.z); out[r * cols + c + 3] = from_f32<T>(v.w);
    }
    if (dtab) {
      const int rb = (int)(r % rows_per_b);
      if (rb < prefix) continue;
      float* d = dtab + idx[(r / rows_per_b) * idx_stride + (rb - prefix)] * tab_ld + c;
      atomicAdd(d + 0, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
    }
  }
}
}  // namespace

template <typename T>
int k_grad_prep(const float* g, T* out, long long rows, int cols, float drop_p, SeedRef seed, unsigned long long stream_id,
                float* dtab, const long long* idx, int tab_ld, int rows_per_b, long long idx_stride, int prefix, cudaStream_t stream) {
  NDT1_REQUIRE(cols % 4 == 0, "grad_prep: cols %d must be a multiple of 4", cols);
  if (rows * cols == 0) return 0;
  const long long work = rows * (cols / 4);
  const int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)rows * cols * (4 + sizeof(T)));
  ndt1_launch(grad_prep_kernel<T>, blocks, 256, 0, stream, g, out, rows, cols, drop_p, seed, stream_id, dtab, idx, tab_ld, rows_per_b > 0 ? rows_per_b : 1,
                                                  idx_stride, prefix);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_grad_prep<float>(const float*, float*, long long, int, float, SeedRef, unsigned long long, float*, const long long*, int, int, long long, int, cudaStream_t);
template int k_grad_prep<bf16>(const float*, bf16*, long long, int, float, SeedRef, unsigned long long, float*, const long long*, int, int, long long, int, cudaStream_t);

// ===========================================================================
// Masked reconstruction loss (mlm / autoregressive): NDT1.forward
// models/ndt1.py:548-578 with nn.PoissonNLLLoss(log_input) / nn.MSELoss.
//   loss = sum_{b,t,n} w[b,t,n] * l(pred, target),   dpred = dloss * w * l'
//   kind 1: poisson log_input   l = exp(x) - t*x
//   kind 2: poisson rate input  l = x - t*log(x + 1e-8)
//   kind 3: mse                 l = (x-t)^2
// weight w = tmask[b,t,n] & pmask[b,t]  (mlm)   or pmask[b,t] (autoregressive,
// with the prediction at t scored against the target at t+1).
// ===========================================================================
namespace {
struct ReconParams {
  const float* pred; const float* target; float* dpred;
  const long long* tmask; const long long* pmask;
  int B, T, N, kind, shift, relu_out;
  float* loss; long long* count; const float* dloss;
};

__global__ void recon_loss_kernel(const ReconParams p) { pdl_grid_sync();
  const long long total = (long long)p.B * p.T * p.N;
  float lsum = 0.f; long long cnt = 0;
  const float gs = p.dloss ? *p.dloss : 1.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e % p.N);
    const long long bt = e / p.N;
    const int t = (int)(bt % p.T), b = (int)(bt / p.T);
    bool w;
    float tg;
    if (p.shift) {
      w = (t < p.T - 1) && p.pmask[(long long)b * p.T + t] != 0;
      tg = w ? p.target[e + p.N] : 0.f;
    } else {
      w = p.tmask[e] != 0 && p.pmask[(long long)b * p.T + t] != 0;
      tg = p.target[e];
    }
    const float x = p.pred[e];
    float l, dl;
    if (p.kind == 1) { const float ex = expf(x); l = ex - tg * x; dl = ex - tg; }
    else if (p.kind == 2) { l = x - tg * logf(x + 1e-8f); dl = 1.f - tg / (x + 1e-8f); }
    else { const float d = x - tg; l = d * d; dl = 2.f * d; }
    if (p.relu_out && x <= 0.f) dl = 0.f;   // decoder ReLU (models/ndt1.py:496-497)
    if (w) { lsum += l; cnt += 1; }
    if (p.dpred) p.dpred[e] = w ? dl * gs : 0.f;
  }
  lsum = warp_sum(lsum);
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  __shared__ float sl[32]; __shared__ long long sc[32];
  if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = lsum; sc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f; long long c = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { a += sl[i]; c += sc[i]; }
    if (p.loss) atomicAdd(p.loss, a);
    if (p.count) atomicAdd((unsigned long long*)p.count, (unsigned long long)c);
  }
}
}  // namespace

int k_recon_loss(const float* pred, const float* target, float* dpred, const long long* tmask, const long long* pmask, int B, int T,
                 int N, int kind, int shift, int relu_out, float* loss, long long* count, const float* dloss, cudaStream_t stream) {
  NDT1_REQUIRE(kind >= 1 && kind <= 3, "recon_loss: unknown loss kind %d", kind);
  const long long total = (long long)B * T * N;
  if (total == 0) return 0;
  ReconParams p{pred, target, dpred, tmask, pmask, B, T, N, kind, shift, relu_out, loss, count, dloss};
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  ndt1_launch(recon_loss_kernel, blocks, 256, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ===========================================================================
// AdamW (torch.optim.AdamW semantics, models/trainer.py:229): one flat buffer.
// ===========================================================================
namespace {
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float gscale) { pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}
}  // namespace

int k_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd, int step,
            float gscale, cudaStream_t stream) {
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  ndt1_launch(adamw_kernel, blocks, 256, 0, stream, p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2, gscale);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// One pass over the flat arenas, four parameters per thread: AdamW + bf16 shadow of the new weights + gradient reset.
namespace {
__global__ void __launch_bounds__(256) adamw_fused_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                          float4* __restrict__ v, long long n4, float lr, float b1, float b2, float eps,
                                                          float wd, float bc1, float rbc2, float gscale, uint2* __restrict__ shadow, int zero_grad) { pdl_grid_sync();
  const float decay = 1.f - lr * wd, step = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = g[i], m4 = m[i], v4 = v[i];
    float4 p4 = p[i];
    float4 mo, vo;
#define NDT1_ADAM1(c)                                                       \
    {                                                                       \
      const float gi = g4.c * gscale;                                       \
      mo.c = b1 * m4.c + (1.f - b1) * gi;                                   \
      vo.c = b2 * v4.c + (1.f - b2) * gi * gi;                              \
      p4.c = p4.c * decay - step * (mo.c / (sqrtf(vo.c) * rbc2 + eps));     \
    }
    NDT1_ADAM1(x) NDT1_ADAM1(y) NDT1_ADAM1(z) NDT1_ADAM1(w)
#undef NDT1_ADAM1
    m[i] = mo; v[i] = vo; p[i] = p4;
    if (shadow) {
      __nv_bfloat162 a = __floats2bfloat162_rn(p4.x, p4.y), b = __floats2bfloat162_rn(p4.z, p4.w);
      uint2 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b;
      shadow[i] = o;
    }
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
}  // namespace

int k_adamw_fused(float* p, float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd, int step,
                  float gscale, bf16* shadow, int zero_grad, cudaStream_t stream) {
  if (n == 0) return 0;
  NDT1_REQUIRE(n % 4 == 0 && (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)shadow & 7) == 0,
               "adamw_fused: the arenas must be 16-byte aligned and a multiple of 4 elements long");
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  const long long n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  // parameter, gradient, two moments read; parameter, two moments (+ cleared gradient, + bf16 shadow) written
  if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)n * (16 + 12 + (zero_grad ? 4 : 0) + (shadow ? 2 : 0)));
  ndt1_launch(adamw_fused_kernel, blocks, 256, 0, stream, (float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, lr, b1, b2, eps, wd, bc1,
                                                 1.0f / sqrtf(bc2), gscale, (uint2*)shadow, zero_grad);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ===========================================================================
// Small index kernels of the embedding layer
// ===========================================================================
namespace {
// mask'[b,r] = AND_{k<size} mask[b, r*stride + k]   (models/ndt1.py:182-183)
__global__ void stack_mask_kernel(const long long* mask, long long* out, int B, int T, int Tp, int size, int stride, int n_prefix) { pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = Tp + n_prefix;
  if (i >= B * L) return;
  const int b = i / L, r = i % L;
  if (r < n_prefix) { out[i] = 1; return; }
  long long m = 1;
  if (size > 0) {
    for (int k = 0; k < size; ++k) m *= mask[(long long)b * T + (r - n_prefix) * stride + k];
  } else {
    m = mask[(long long)b * T + (r - n_prefix)];
  }
  out[i] = m;
}
// x[b, slot, :] = table[idx[b], :]  (block / day tokens, models/ndt1.py:192-201)
__global__ void token_rows_kernel(const float* table, const long long* idx, float* x, int B, int L, int H, int slot) { pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, c = i % H;
  x[((long long)b * L + slot) * H + c] = table[idx[b] * H + c];
}
template <typename T>
__global__ void token_rows_grad_kernel(float* dtable, const long long* idx, const T* dx, int B, int L, int H, int slot) { pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, c = i % H;
  atomicAdd(dtable + idx[b] * H + c, to_f32(dx[((long long)b * L + slot) * H + c]));
}
}  // namespace

int k_stack_mask(const long long* mask, long long* out, int B, int T, int Tp, int size, int stride, int n_prefix, cudaStream_t stream) {
  const int n = B * (Tp + n_prefix);
  if (n == 0) return 0;
  ndt1_launch(stack_mask_kernel, ndt1_cdiv(n, 256), 256, 0, stream, mask, out, B, T, Tp, size, stride, n_prefix);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int k_token_rows(const float* table, const long long* idx, float* x, int B, int L, int H, int slot, cudaStream_t stream) {
  ndt1_launch(token_rows_kernel, ndt1_cdiv(B * H, 256), 256, 0, stream, table, idx, x, B, L, H, slot);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template <typename T>
int k_token_rows_grad(float* dtable, const long long* idx, const T* dx, int B, int L, int H, int slot, cudaStream_t stream) {
  ndt1_launch(token_rows_grad_kernel<T>, ndt1_cdiv(B * H, 256), 256, 0, stream, dtable, idx, dx, B, L, H, slot);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_token_rows_grad<float>(float*, const long long*, const float*, int, int, int, int, cudaStream_t);
template int k_token_rows_grad<bf16>(float*, const long long*, const bf16*, int, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------
// out[r, c] = T(in[r, c] * (*scale)) for c < cols, 0 for cols <= c < ld_out   (dlogits hand-off)
// lens'[b] = trunc(1 + (len - size) / stride)  (models/ndt1.py:207-208);  v[0] = value
// ---------------------------------------------------------------------------
namespace {
template <typename T>
__global__ void scale_cast_pad_kernel(const float* __restrict__ in, T* __restrict__ out, long long rows, int cols, int ld_out, const float* scale) { pdl_grid_sync();
  const float s = scale ? *scale : 1.f;
  const long long total = rows * ld_out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_out; const int c = (int)(i % ld_out);
    out[i] = from_f32<T>(c < cols ? in[r * cols + c] * s : 0.f);
  }
}
__global__ void stacked_lens_kernel(const long long* lens, long long* out, int B, int stack, int size, int stride) { pdl_grid_sync();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out[b] = stack ? (long long)(1.0f + (float)(lens[b] - size) / (float)stride) : lens[b];
}
__global__ void set_i64_kernel(long long* p, long long v) { pdl_grid_sync(); *p = v; }
__global__ void relu_inplace_kernel(float* x, long long n) { pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = fmaxf(x[i], 0.f);
}
__global__ void and_mask_kernel(const long long* tmask, const long long* pmask, long long* out, int B, int T, int N) { pdl_grid_sync();
  const long long total = (long long)B * T * N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    out[e] = tmask[e] & pmask[e / N];
}
}  // namespace

template <typename T>
int k_scale_cast_pad(const float* in, T* out, long long rows, int cols, int ld_out, const float* scale, cudaStream_t stream) {
  const long long total = rows * ld_out;
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  ndt1_launch(scale_cast_pad_kernel<T>, blocks, 256, 0, stream, in, out, rows, cols, ld_out, scale);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_scale_cast_pad<float>(const float*, float*, long long, int, int, const float*, cudaStream_t);
template int k_scale_cast_pad<bf16>(const float*, bf16*, long long, int, int, const float*, cudaStream_t);

int k_stacked_lens(const long long* lens, long long* out, int B, int stack, int size, int stride, cudaStream_t stream) {
  if (B == 0) return 0;
  ndt1_launch(stacked_lens_kernel, ndt1_cdiv(B, 128), 128, 0, stream, lens, out, B, stack, size, stride);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int k_set_i64(long long* p, long long v, cudaStream_t stream) {
  ndt1_launch(set_i64_kernel, 1, 1, 0, stream, p, v);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int k_relu_inplace(float* x, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  ndt1_launch(relu_inplace_kernel, blocks, 256, 0, stream, x, n);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int k_and_mask(const long long* tmask, const long long* pmask, long long* out, int B, int T, int N, cudaStream_t stream) {
  const long long total = (long long)B * T * N;
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  ndt1_launch(and_mask_kernel, blocks, 256, 0, stream, tmask, pmask, out, B, T, N);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------
// Options of the path that configs/ndt1.yaml leaves off: rotary positions, dropout in front of the factors
// projection, per-day channel embedding.  Plain coalesced element-wise kernels (none of them is on the benchmark path).
// ---------------------------------------------------------------------------
namespace {

// apply_rotary_pos_emb (models/ndt1.py:52-71, 285-286) in place on the q and k sections of the packed (B*L, 3H) buffer:
//   x'[i] = x[i] cos[pos][i] - x[i + hd/2] sin[pos][i],   x'[i + hd/2] = x[i + hd/2] cos[pos][i] + x[i] sin[pos][i]
// (cos[pos][i] == cos[pos][i + hd/2]: the table is cat(freqs, freqs)).  inverse = 1 applies the transpose (the backward).
template <typename T>
__global__ void rope_kernel(T* __restrict__ qkv, const long long* __restrict__ ts, long long ts_stride, const float* __restrict__ cs,
                            const float* __restrict__ sn, long long rows, int L, int H, int nh, int hd, int max_F, int inverse) { pdl_grid_sync();
  const int half = hd >> 1;
  const long long per_row = 2LL * nh * half;                  // (section, head, i)
  const long long total = rows * per_row;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long m = e / per_row;
    const int w = (int)(e - m * per_row);
    const int i = w % half, h = (w / half) % nh, sec = w / (half * nh);
    long long pos = ts[(m / L) * ts_stride + (m % L)];
    pos = pos < 0 ? 0 : (pos >= max_F ? max_F - 1 : pos);
    const float c = cs[pos * hd + i];
    const float s = inverse ? -sn[pos * hd + i] : sn[pos * hd + i];
    T* v = qkv + m * 3LL * H + (long long)sec * H + (long long)h * hd + i;
    const float x1 = to_f32(v[0]), x2 = to_f32(v[half]);
    v[0] = from_f32<T>(x1 * c - x2 * s);
    v[half] = from_f32<T>(x2 * c + x1 * s);
  }
}

// x *= keep-scale of dropout site `stream_id` (element index = flat index), in place; the same call on a gradient is the backward
template <typename T>
__global__ void dropout_inplace_kernel(T* __restrict__ x, long long n, float p, SeedRef seed_ref, unsigned long long stream_id) { pdl_grid_sync();
  const unsigned long long seed = seed_ref.get();
  const uint32_t thr = drop_threshold(p);
  const float ik = 1.0f / (1.0f - p);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    x[e] = from_f32<T>(to_f32(x[e]) * drop_scale_1(seed, stream_id, (unsigned long long)e, thr, ik));
}

// out[sel[b] * out_stride + c] += sum over the rows of trial b of in[b, r, c]   (per-day embedding bias gradient)
template <typename T>
__global__ void colsum_sel_kernel(const T* __restrict__ in, float* __restrict__ out, const long long* __restrict__ sel, int n_sel,
                                  long long out_stride, int rows_per_b, int cols) { pdl_grid_sync();
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const T* p = in + (long long)b * rows_per_b * cols + c;
  float acc = 0.f;
  for (int r = 0; r < rows_per_b; ++r) acc += to_f32(p[(long long)r * cols]);
  long long d = sel[b];
  d = d < 0 ? 0 : (d >= n_sel ? n_sel - 1 : d);
  atomicAdd(out + d * out_stride + c, acc);
}

template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, long long n) { pdl_grid_sync();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) out[e] = to_f32(in[e]);
}

// g *= act'(saved): gradient through an activation whose GEMM epilogue is not on the path (encoder-only backward with factors)
template <typename T>
__global__ void dact_inplace_kernel(T* __restrict__ g, const T* __restrict__ saved, long long n, int dact) { pdl_grid_sync();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    g[e] = from_f32<T>(to_f32(g[e]) * dact_apply(dact, to_f32(saved[e])));
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) { pdl_grid_sync();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) dst[e] += src[e];
}

inline int ew_blocks(long long n) { return (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8); }

}  // namespace

template <typename T>
int k_rope(T* qkv, const long long* ts, long long ts_stride, const float* cs, const float* sn, long long rows, int L, int H, int nh,
           int max_F, int inverse, cudaStream_t stream) {
  if (rows == 0) return 0;
  const int hd = H / nh;
  NDT1_REQUIRE(hd % 2 == 0, "rope: head size %d must be even", hd);
  ndt1_launch(rope_kernel<T>, ew_blocks(rows * nh * hd), 256, 0, stream, qkv, ts, ts_stride, cs, sn, rows, L, H, nh, hd, max_F, inverse);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_rope<float>(float*, const long long*, long long, const float*, const float*, long long, int, int, int, int, int, cudaStream_t);
template int k_rope<bf16>(bf16*, const long long*, long long, const float*, const float*, long long, int, int, int, int, int, cudaStream_t);

template <typename T>
int k_dropout_inplace(T* x, long long n, float p, SeedRef seed, unsigned long long stream_id, cudaStream_t stream) {
  if (n == 0 || p <= 0.f) return 0;
  ndt1_launch(dropout_inplace_kernel<T>, ew_blocks(n), 256, 0, stream, x, n, p, seed, stream_id);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_dropout_inplace<float>(float*, long long, float, SeedRef, unsigned long long, cudaStream_t);
template int k_dropout_inplace<bf16>(bf16*, long long, float, SeedRef, unsigned long long, cudaStream_t);

template <typename T>
int k_colsum_sel(const T* in, float* out, const long long* sel, int n_sel, long long out_stride, int B, int rows_per_b, int cols,
                 cudaStream_t stream) {
  if (B == 0 || rows_per_b == 0) return 0;
  ndt1_launch(colsum_sel_kernel<T>, dim3(ndt1_cdiv(cols, 128), B), 128, 0, stream, in, out, sel, n_sel, out_stride, rows_per_b, cols);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_colsum_sel<float>(const float*, float*, const long long*, int, long long, int, int, int, cudaStream_t);
template int k_colsum_sel<bf16>(const bf16*, float*, const long long*, int, long long, int, int, int, cudaStream_t);

int k_add_inplace(float* dst, const float* src, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  ndt1_launch(add_inplace_kernel, ew_blocks(n), 256, 0, stream, dst, src, n);
  NDT1_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int k_cast_to_f32(const T* in, float* out, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  ndt1_launch(cast_to_f32_kernel<T>, ew_blocks(n), 256, 0, stream, in, out, n);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_cast_to_f32<float>(const float*, float*, long long, cudaStream_t);
template int k_cast_to_f32<bf16>(const bf16*, float*, long long, cudaStream_t);

template <typename T>
int k_dact_inplace(T* g, const T* saved, long long n, int dact, cudaStream_t stream) {
  if (n == 0 || dact == DACT_NONE) return 0;
  ndt1_launch(dact_inplace_kernel<T>, ew_blocks(n), 256, 0, stream, g, saved, n, dact);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_dact_inplace<float>(float*, const float*, long long, int, cudaStream_t);
template int k_dact_inplace<bf16>(bf16*, const bf16*, long long, int, cudaStream_t);
