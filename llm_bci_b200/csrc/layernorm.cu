// LayerNorm forward / backward over the fp32 residual stream (eps inside the
// sqrt, biased variance: torch.nn.LayerNorm as used at models/ndt1.py:309-311,402).
// One warp per row, 128-bit loads, warp-shuffle reductions, statistics saved
// for the backward.  The backward also folds in the residual-gradient add and
// emits the (dropout-masked, low-precision) operand of the next GEMM so the
// fp32 gradient is read exactly once.
// Bytes per row (H=1024): fwd 4 KB in + 2 KB (bf16) out; bwd 2 KB + 4 KB + 4 KB in, 4 KB + 2 KB out.
#include "kernels.cuh"

namespace {

constexpr int LN_MAXV = 8;        // float4 per lane -> H <= 1024
constexpr int LN_WARPS = 4;
constexpr int LN_MAX_BLOCKS = 148 * 4;

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) { *(float4*)p = make_float4(a, b, c, d); }
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 x = __floats2bfloat162_rn(a, b), y = __floats2bfloat162_rn(c, d);
  uint2 o; o.x = *(uint32_t*)&x; o.y = *(uint32_t*)&y;
  *(uint2*)p = o;
}
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *(const float4*)p; }
template <>
__device__ __forceinline__ float4 load4<bf16>(const bf16* p) {
  const uint2 r = *(const uint2*)p;
  const __nv_bfloat162 x = *(const __nv_bfloat162*)&r.x, y = *(const __nv_bfloat162*)&r.y;
  return make_float4(__low2float(x), __high2float(x), __low2float(y), __high2float(y));
}

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean,
                                                              float* __restrict__ rstd, long long rows, int H, float eps) { pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int nv = H / 4;
  for (long long r = warp0; r < rows; r += (long long)gridDim.x * LN_WARPS) {
    const float* xr = x + r * H;
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) { v[i] = *(const float4*)(xr + c * 4); s += v[i].x + v[i].y + v[i].z + v[i].w; }
    }
    const float mu = warp_sum(s) / H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / H + eps);
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = *(const float4*)(gamma + c * 4), b = *(const float4*)(beta + c * 4);
        store4<T>(y + r * H + c * 4, (v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y, (v[i].z - mu) * rs * g.z + b.z,
                  (v[i].w - mu) * rs * g.w + b.w);
      }
    }
  }
}

// Backward.  One warp per row, 8 warps per CTA, one CTA per SM (persistent over rows): every lane issues the
// whole row's loads (dy, x AND the residual gradient) before the first reduction, so a warp keeps 10 KB in flight
// and the 8 warps of an SM cover the HBM latency.  The affine gradients (and the column sums of the emitted
// operand = the next layer's bias gradient) are accumulated in registers over the rows of the warp, folded across
// the CTA's warps through shared memory (plain stores + a tree, no atomics) and added to the outputs with red.global.
constexpr int LNB_WARPS = 8;    // two warps per SM sub-partition: up to 255 registers per thread

template <typename T> struct Raw4;
template <> struct Raw4<float> { typedef float4 type; };
template <> struct Raw4<bf16> { typedef uint2 type; };
__device__ __forceinline__ float4 unpack4(float4 r) { return r; }
__device__ __forceinline__ float4 unpack4(uint2 r) {
  const __nv_bfloat162 x = *(const __nv_bfloat162*)&r.x, y = *(const __nv_bfloat162*)&r.y;
  return make_float4(__low2float(x), __high2float(x), __low2float(y), __high2float(y));
}

template <typename T>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
ln_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, float* __restrict__ dres, T* __restrict__ out_lp, float drop_p, SeedRef seed_ref,
              unsigned long long stream_id, long long rows, int H, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ colsum_out) { pdl_grid_sync();
  const unsigned long long seed = drop_p > 0.f ? seed_ref.get() : 0ull;
  extern __shared__ float sm[];   // [warps][3][H]: per-warp partial sums of dgamma | dbeta | colsum(out_lp)
  typedef typename Raw4<T>::type raw_t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * LNB_WARPS + warp;
  const int nv = H / 4;
  float4 dg[LN_MAXV], db[LN_MAXV], cs[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) { dg[i] = make_float4(0.f, 0.f, 0.f, 0.f); db[i] = dg[i]; cs[i] = dg[i]; }
  const uint32_t thr = drop_threshold(drop_p);
  const float ik = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
  for (long long r = warp0; r < rows; r += (long long)gridDim.x * LNB_WARPS) {
    raw_t dr[LN_MAXV]; float4 xv[LN_MAXV], o[LN_MAXV];
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        dr[i] = *(const raw_t*)(dy + r * H + c * 4);
        xv[i] = *(const float4*)(x + r * H + c * 4);
        o[i] = *(const float4*)(dres + r * H + c * 4);
      }
    }
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 gm = __ldg((const float4*)(gamma + c * 4));
        const float4 d = unpack4(dr[i]);
        xv[i] = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);   // x-hat
        dg[i].x += d.x * xv[i].x; dg[i].y += d.y * xv[i].y; dg[i].z += d.z * xv[i].z; dg[i].w += d.w * xv[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
        const float4 g = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);                                // dy * gamma
        s1 += g.x + g.y + g.z + g.w;
        s2 += g.x * xv[i].x + g.y * xv[i].y + g.z * xv[i].z + g.w * xv[i].w;
      }
    }
    const float c1 = warp_sum(s1) / H, c2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 gm = __ldg((const float4*)(gamma + c * 4));
        const float4 d = unpack4(dr[i]);
        o[i].x += rs * (d.x * gm.x - c1 - xv[i].x * c2); o[i].y += rs * (d.y * gm.y - c1 - xv[i].y * c2);
        o[i].z += rs * (d.z * gm.z - c1 - xv[i].z * c2); o[i].w += rs * (d.w * gm.w - c1 - xv[i].w * c2);
        *(float4*)(dres + r * H + c * 4) = o[i];
        if (out_lp) {
          float4 q = o[i];
          if (drop_p > 0.f) {
            float ds[4];
            drop_scale_4(seed, stream_id, (unsigned long long)(r * H + c * 4), thr, ik, ds);
            q.x *= ds[0]; q.y *= ds[1]; q.z *= ds[2]; q.w *= ds[3];
          }
          store4<T>(out_lp + r * H + c * 4, q.x, q.y, q.z, q.w);
          cs[i].x += q.x; cs[i].y += q.y; cs[i].z += q.z; cs[i].w += q.w;   // bias gradient of the layer that consumes out_lp
        }
      }
    }
  }
  // fold across the CTA's warps: plain stores, one barrier, column-parallel sums, one red.global per column per CTA
  float* mine = sm + (size_t)warp * 3 * H;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      *(float4*)(mine + c * 4) = dg[i];
      *(float4*)(mine + H + c * 4) = db[i];
      *(float4*)(mine + 2 * H + c * 4) = cs[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
    float* dst = i < H ? dgamma : (i < 2 * H ? dbeta : colsum_out);
    if (!dst) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) v += sm[(size_t)w * 3 * H + i];
    atomicAdd(dst + (i < H ? i : (i < 2 * H ? i - H : i - 2 * H)), v);
  }
}

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------
// Row-group kernels (H % 128 == 0): a CTA of H/4 threads owns FOUR rows at a time and thread c the 4 columns
// [4c, 4c+4) of each of them.  Compared with one warp per row this keeps the per-thread state tiny (the affine /
// bias-gradient accumulators cover 4 columns, not 32), so several CTAs fit an SM and their loads overlap; the row
// statistics cross the warps through a double-buffered shared-memory table (one barrier per reduction).
// ---------------------------------------------------------------------------
constexpr int LNG_ROWS = 4;

// Sum EIGHT per-thread partials over the CTA.  Within a warp the 8 x 32 values are folded with 9 shuffles (each
// xor step halves the number of values a lane carries instead of reducing them one by one = 40 shuffles); the
// warps meet in a double-buffered shared-memory table (one barrier); on return lane l of every warp holds the
// CTA total of partial (l % 8) -- callers fetch total i with __shfl_sync(.., i).
__device__ __forceinline__ float cta_sum8(const float* v, float* red, int buf, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[4], u[2], t;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = h16 ? v[j] : v[j + 4], keep = h16 ? v[j + 4] : v[j];
    w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = h8 ? w[j] : w[j + 2], keep = h8 ? w[j + 2] : w[j];
    u[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = h4 ? u[0] : u[1], keep = h4 ? u[1] : u[0];
    t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  float* tab = red + buf * 32 * 8;
  if ((lane & 3) == 0) tab[warp * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = t;
  __syncthreads();
  float s = 0.f;
  for (int e = lane; e < nwarps * 8; e += 32) s += tab[e];      // entry e belongs to partial e % 8 == lane % 8
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  return s;
}

static_assert(LNG_ROWS == 4, "cta_sum8 folds exactly eight partials");
constexpr int LNF_ROWS = 8;   // forward: 8 rows per group (only 4 registers per row: twice the loads in flight, half the barriers)
template <typename T, int NT>
__global__ void __launch_bounds__(NT, NT <= 256 ? 4 : 1) ln_fwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean,
                                                           float* __restrict__ rstd, long long rows, int H, float eps) { pdl_grid_sync();
  __shared__ float red[2 * 32 * 8];
  const int c = threadIdx.x, nwarps = blockDim.x >> 5;
  const float4 g = __ldg((const float4*)gamma + c), bt = __ldg((const float4*)beta + c);
  const float inv_h = 1.0f / H;
  for (long long r0 = (long long)blockIdx.x * LNF_ROWS; r0 < rows; r0 += (long long)gridDim.x * LNF_ROWS) {
    float4 v[LNF_ROWS]; float part[LNF_ROWS], mu[LNF_ROWS], var[LNF_ROWS];
#pragma unroll
    for (int r = 0; r < LNF_ROWS; ++r) {
      v[r] = (r0 + r < rows) ? *((const float4*)(x + (r0 + r) * H) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      part[r] = v[r].x + v[r].y + v[r].z + v[r].w;
    }
    const float tot1 = cta_sum8(part, red, 0, nwarps) * inv_h;          // lane l: mean of row l % 8
#pragma unroll
    for (int r = 0; r < LNF_ROWS; ++r) {
      mu[r] = __shfl_sync(0xffffffffu, tot1, r);
      const float a = v[r].x - mu[r], b = v[r].y - mu[r], cc = v[r].z - mu[r], d = v[r].w - mu[r];
      part[r] = a * a + b * b + cc * cc + d * d;
    }
    const float rs_l = 1.0f / sqrtf(cta_sum8(part, red, 1, nwarps) * inv_h + eps);   // computed once per warp, for row l % 8
    (void)var;
#pragma unroll
    for (int r = 0; r < LNF_ROWS; ++r) {
      const float rs = __shfl_sync(0xffffffffu, rs_l, r);
      if (r0 + r >= rows) continue;
      if (c == 0) { mean[r0 + r] = mu[r]; rstd[r0 + r] = rs; }
      store4<T>(y + (r0 + r) * H + c * 4, (v[r].x - mu[r]) * rs * g.x + bt.x, (v[r].y - mu[r]) * rs * g.y + bt.y,
                (v[r].z - mu[r]) * rs * g.z + bt.z, (v[r].w - mu[r]) * rs * g.w + bt.w);
    }
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1)
ln_bwd_rows_kernel(const T* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
                   const float* __restrict__ rstd, float* __restrict__ dres, T* __restrict__ out_lp, float drop_p, SeedRef seed_ref,
                   unsigned long long stream_id, long long rows, int H, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ colsum_out) { pdl_grid_sync();
  const unsigned long long seed = drop_p > 0.f ? seed_ref.get() : 0ull;
  __shared__ float red[2 * 32 * 8];
  typedef typename Raw4<T>::type raw_t;
  const int c = threadIdx.x, lane = c & 31, nwarps = blockDim.x >> 5;
  const float4 gm = __ldg((const float4*)gamma + c);
  const float inv_h = 1.0f / H;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg, cs = dg;
  const uint32_t thr = drop_threshold(drop_p);
  const float ik = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
  int buf = 0;
  for (long long r0 = (long long)blockIdx.x * LNG_ROWS; r0 < rows; r0 += (long long)gridDim.x * LNG_ROWS) {
    raw_t dr[LNG_ROWS]; float4 xv[LNG_ROWS], o[LNG_ROWS]; float mu[LNG_ROWS], rs[LNG_ROWS];
#pragma unroll
    for (int r = 0; r < LNG_ROWS; ++r) {
      const bool in = r0 + r < rows;
      const long long e = (in ? r0 + r : r0) * H + c * 4;
      dr[r] = *(const raw_t*)(dy + e);
      xv[r] = *(const float4*)(x + e);
      o[r] = *(const float4*)(dres + e);
      mu[r] = mean[in ? r0 + r : r0]; rs[r] = in ? rstd[r0 + r] : 0.f;     // rs = 0 neutralises a row past the end
    }
    float part[2 * LNG_ROWS], sums[2 * LNG_ROWS];
    float4 gd[LNG_ROWS];
#pragma unroll
    for (int r = 0; r < LNG_ROWS; ++r) {
      float4 d = unpack4(dr[r]);
      if (r0 + r >= rows) d = make_float4(0.f, 0.f, 0.f, 0.f);
      xv[r] = make_float4((xv[r].x - mu[r]) * rs[r], (xv[r].y - mu[r]) * rs[r], (xv[r].z - mu[r]) * rs[r], (xv[r].w - mu[r]) * rs[r]);   // x-hat
      dg.x += d.x * xv[r].x; dg.y += d.y * xv[r].y; dg.z += d.z * xv[r].z; dg.w += d.w * xv[r].w;
      db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
      gd[r] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);                                                              // dy * gamma
      part[2 * r] = gd[r].x + gd[r].y + gd[r].z + gd[r].w;
      part[2 * r + 1] = gd[r].x * xv[r].x + gd[r].y * xv[r].y + gd[r].z * xv[r].z + gd[r].w * xv[r].w;
    }
    const float tot = cta_sum8(part, red, buf, nwarps) * inv_h;         // lane l: partial l % 8 of the group
    buf ^= 1;
#pragma unroll
    for (int i = 0; i < 2 * LNG_ROWS; ++i) sums[i] = __shfl_sync(0xffffffffu, tot, i);
    // dropout keep-scales of out_lp: one Philox block = 8 elements = the columns of a lane pair; the even lane draws the
    // block of rows 0 / 2, the odd lane of rows 1 / 3, and they swap halves (one Philox per 8 elements)
    float keep[LNG_ROWS][4];
    if (out_lp && drop_p > 0.f) {
      const bool odd = lane & 1;
#pragma unroll
      for (int k = 0; k < LNG_ROWS / 2; ++k) {
        const long long row = r0 + 2 * k + (odd ? 1 : 0);
        const unsigned long long elem = (unsigned long long)row * H + (unsigned)((c * 4) & ~7);
        const Philox4 ph = philox4x32(seed, elem >> 3, stream_id);
        const uint32_t s0 = odd ? ph.x : ph.z, s1 = odd ? ph.y : ph.w;
        const uint32_t q0 = __shfl_xor_sync(0xffffffffu, s0, 1), q1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        const uint32_t a0 = odd ? q0 : ph.x, a1 = odd ? q1 : ph.y;     // row 2k
        const uint32_t b0 = odd ? ph.z : q0, b1 = odd ? ph.w : q1;     // row 2k+1
        keep[2 * k][0] = (a0 & 0xFFFFu) >= thr ? ik : 0.f; keep[2 * k][1] = (a0 >> 16) >= thr ? ik : 0.f;
        keep[2 * k][2] = (a1 & 0xFFFFu) >= thr ? ik : 0.f; keep[2 * k][3] = (a1 >> 16) >= thr ? ik : 0.f;
        keep[2 * k + 1][0] = (b0 & 0xFFFFu) >= thr ? ik : 0.f; keep[2 * k + 1][1] = (b0 >> 16) >= thr ? ik : 0.f;
        keep[2 * k + 1][2] = (b1 & 0xFFFFu) >= thr ? ik : 0.f; keep[2 * k + 1][3] = (b1 >> 16) >= thr ? ik : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < LNG_ROWS; ++r) {
      if (r0 + r >= rows) continue;
      const float c1 = sums[2 * r], c2 = sums[2 * r + 1];
      o[r].x += rs[r] * (gd[r].x - c1 - xv[r].x * c2); o[r].y += rs[r] * (gd[r].y - c1 - xv[r].y * c2);
      o[r].z += rs[r] * (gd[r].z - c1 - xv[r].z * c2); o[r].w += rs[r] * (gd[r].w - c1 - xv[r].w * c2);
      const long long e = (r0 + r) * H + c * 4;
      *(float4*)(dres + e) = o[r];
      if (out_lp) {
        float4 q = o[r];
        if (drop_p > 0.f) { q.x *= keep[r][0]; q.y *= keep[r][1]; q.z *= keep[r][2]; q.w *= keep[r][3]; }
        store4<T>(out_lp + e, q.x, q.y, q.z, q.w);
        cs.x += q.x; cs.y += q.y; cs.z += q.z; cs.w += q.w;     // bias gradient of the layer that consumes out_lp
      }
    }
  }
  // this thread's accumulators already hold the CTA's sums for its four columns
  if (dgamma) red_add4(dgamma + c * 4, dg);
  if (dbeta) red_add4(dbeta + c * 4, db);
  if (colsum_out && out_lp) red_add4(colsum_out + c * 4, cs);
}

int g_sms() { return ndt1_num_sms(); }
int ln_blocks(long long rows) {
  long long b = (rows + LN_WARPS - 1) / LN_WARPS;
  return (int)(b < LN_MAX_BLOCKS ? b : LN_MAX_BLOCKS);
}

}  // namespace

size_t k_layernorm_bwd_partials_bytes(int H) { return (size_t)LN_MAX_BLOCKS * 2 * H * sizeof(float); }

template <typename T>
int k_layernorm_fwd(const float* x, const float* gamma, const float* beta, T* y, float* mean, float* rstd, long long rows, int H,
                    float eps, cudaStream_t stream) {
  NDT1_REQUIRE(H % 4 == 0 && H <= 128 * LN_MAXV, "layernorm: hidden size %d unsupported (multiple of 4, <= %d)", H, 128 * LN_MAXV);
  if (rows == 0) return 0;
  if (H % 128 == 0 && H <= 4096) {
    const int threads = H / 4;
    long long nb = (rows + LNF_ROWS - 1) / LNF_ROWS;
    const long long cap = (long long)g_sms() * (threads <= 256 ? 4 : 1);
    if (nb > cap) nb = cap;
    if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)rows * H * (4 + sizeof(T)));     // x in (fp32), y out
    if (threads <= 256) ndt1_launch(ln_fwd_rows_kernel<T, 256>, (int)nb, threads, 0, stream, x, gamma, beta, y, mean, rstd, rows, H, eps);
    else ndt1_launch(ln_fwd_rows_kernel<T, 1024>, (int)nb, threads, 0, stream, x, gamma, beta, y, mean, rstd, rows, H, eps);
    NDT1_CHECK_LAUNCH();
    return 0;
  }
  ndt1_launch(ln_fwd_kernel<T>, ln_blocks(rows), LN_WARPS * 32, 0, stream, x, gamma, beta, y, mean, rstd, rows, H, eps);
  NDT1_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int k_layernorm_bwd(const T* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dres, float* dgamma,
                    float* dbeta, T* out_lp, float drop_p, SeedRef seed, unsigned long long stream_id, long long rows, int H,
                    float* partials, cudaStream_t stream, float* colsum_out) {
  NDT1_REQUIRE(H % 4 == 0 && H <= 128 * LN_MAXV, "layernorm: hidden size %d unsupported (multiple of 4, <= %d)", H, 128 * LN_MAXV);
  if (rows == 0) return 0;
  (void)partials;
  const int sms = g_sms();
  if (H % 128 == 0 && H <= 4096) {
    const int threads = H / 4;
    long long nbr = (rows + LNG_ROWS - 1) / LNG_ROWS;
    const long long cap = (long long)sms * (threads <= 256 ? 2 : 1);
    if (nbr > cap) nbr = cap;
    // dy in, x in (fp32), residual gradient in + out (fp32), the next GEMM pair's operand out
    if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)rows * H * (sizeof(T) + 4 + 8 + (out_lp ? sizeof(T) : 0)));
    if (threads <= 256)
      ndt1_launch(ln_bwd_rows_kernel<T, 256>, (int)nbr, threads, 0, stream, dy, x, gamma, mean, rstd, dres, out_lp, drop_p, seed, stream_id, rows, H,
                                                                    dgamma, dbeta, out_lp ? colsum_out : nullptr);
    else
      ndt1_launch(ln_bwd_rows_kernel<T, 1024>, (int)nbr, threads, 0, stream, dy, x, gamma, mean, rstd, dres, out_lp, drop_p, seed, stream_id, rows, H,
                                                                     dgamma, dbeta, out_lp ? colsum_out : nullptr);
    NDT1_CHECK_LAUNCH();
    return 0;
  }
  long long nb = (rows + LNB_WARPS - 1) / LNB_WARPS;
  if (nb > sms) nb = sms;
  const size_t smem = (size_t)LNB_WARPS * 3 * H * sizeof(float);
  static Ndt1PerDeviceSize attr;
  if (smem > 48 * 1024 && smem > attr.here()) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.here() = smem;
  }
  ndt1_launch(ln_bwd_kernel<T>, (int)nb, LNB_WARPS * 32, smem, stream, dy, x, gamma, mean, rstd, dres, out_lp, drop_p, seed,
                                                                               stream_id, rows, H, dgamma, dbeta, out_lp ? colsum_out : nullptr);
  NDT1_CHECK_LAUNCH();
  return 0;
}

template int k_layernorm_fwd<float>(const float*, const float*, const float*, float*, float*, float*, long long, int, float, cudaStream_t);
template int k_layernorm_fwd<bf16>(const float*, const float*, const float*, bf16*, float*, float*, long long, int, float, cudaStream_t);
template int k_layernorm_bwd<float>(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*, float,
                                    SeedRef, unsigned long long, long long, int, float*, cudaStream_t, float*);
template int k_layernorm_bwd<bf16>(const bf16*, const float*, const float*, const float*, const float*, float*, float*, float*, bf16*, float,
                                   SeedRef, unsigned long long, long long, int, float*, cudaStream_t, float*);
