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
                                                              float* __restrict__ rstd, long long rows, int H, float eps) {
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

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean,
              const float* __restrict__ rstd, float* __restrict__ dres, T* __restrict__ out_lp, float drop_p, unsigned long long seed,
              unsigned long long stream_id, long long rows, int H, float* __restrict__ partials) {
  extern __shared__ float sm[];   // [2][H]
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int nv = H / 4;
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float4 dg[LN_MAXV], db[LN_MAXV], gm[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f); db[i] = dg[i];
    const int c = lane + 32 * i;
    gm[i] = c < nv ? *(const float4*)(gamma + c * 4) : dg[i];
  }
  const uint32_t thr = drop_threshold(drop_p);
  const float ik = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
  for (long long r = warp0; r < rows; r += (long long)gridDim.x * LN_WARPS) {
    const float mu = mean[r], rs = rstd[r];
    float4 g[LN_MAXV], xh[LN_MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 d = load4<T>(dy + r * H + c * 4);
        const float4 xv = *(const float4*)(x + r * H + c * 4);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      }
    }
    const float c1 = warp_sum(s1) / H, c2 = warp_sum(s2) / H;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float* dr = dres + r * H + c * 4;
        float4 o = *(float4*)dr;
        o.x += rs * (g[i].x - c1 - xh[i].x * c2); o.y += rs * (g[i].y - c1 - xh[i].y * c2);
        o.z += rs * (g[i].z - c1 - xh[i].z * c2); o.w += rs * (g[i].w - c1 - xh[i].w * c2);
        *(float4*)dr = o;
        if (out_lp) {
          if (drop_p > 0.f) {
            float ds[4];
            drop_scale_4(seed, stream_id, (unsigned long long)(r * H + c * 4), thr, ik, ds);
            o.x *= ds[0]; o.y *= ds[1]; o.z *= ds[2]; o.w *= ds[3];
          }
          store4<T>(out_lp + r * H + c * 4, o.x, o.y, o.z, o.w);
        }
      }
    }
  }
  // block reduction of the affine gradients, then one partial row per block
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      atomicAdd(&sm[c * 4 + 0], dg[i].x); atomicAdd(&sm[c * 4 + 1], dg[i].y); atomicAdd(&sm[c * 4 + 2], dg[i].z); atomicAdd(&sm[c * 4 + 3], dg[i].w);
      atomicAdd(&sm[H + c * 4 + 0], db[i].x); atomicAdd(&sm[H + c * 4 + 1], db[i].y); atomicAdd(&sm[H + c * 4 + 2], db[i].z); atomicAdd(&sm[H + c * 4 + 3], db[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) partials[(long long)blockIdx.x * 2 * H + i] = sm[i];
}

__global__ void ln_bwd_reduce_kernel(const float* __restrict__ partials, int nblocks, int H, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * H) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partials[(long long)b * 2 * H + c];
  if (c < H) dgamma[c] += s;
  else dbeta[c - H] += s;
}

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
  ln_fwd_kernel<T><<<ln_blocks(rows), LN_WARPS * 32, 0, stream>>>(x, gamma, beta, y, mean, rstd, rows, H, eps);
  NDT1_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int k_layernorm_bwd(const T* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dres, float* dgamma,
                    float* dbeta, T* out_lp, float drop_p, unsigned long long seed, unsigned long long stream_id, long long rows, int H,
                    float* partials, cudaStream_t stream) {
  NDT1_REQUIRE(H % 4 == 0 && H <= 128 * LN_MAXV, "layernorm: hidden size %d unsupported (multiple of 4, <= %d)", H, 128 * LN_MAXV);
  if (rows == 0) return 0;
  const int nb = ln_blocks(rows);
  ln_bwd_kernel<T><<<nb, LN_WARPS * 32, 2 * H * sizeof(float), stream>>>(dy, x, gamma, mean, rstd, dres, out_lp, drop_p, seed, stream_id,
                                                                          rows, H, partials);
  NDT1_CHECK_LAUNCH();
  ln_bwd_reduce_kernel<<<ndt1_cdiv(2 * H, 128), 128, 0, stream>>>(partials, nb, H, dgamma, dbeta);
  NDT1_CHECK_LAUNCH();
  return 0;
}

template int k_layernorm_fwd<float>(const float*, const float*, const float*, float*, float*, float*, long long, int, float, cudaStream_t);
template int k_layernorm_fwd<bf16>(const float*, const float*, const float*, bf16*, float*, float*, long long, int, float, cudaStream_t);
template int k_layernorm_bwd<float>(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*, float,
                                    unsigned long long, unsigned long long, long long, int, float*, cudaStream_t);
template int k_layernorm_bwd<bf16>(const bf16*, const float*, const float*, const float*, const float*, float*, float*, float*, bf16*, float,
                                   unsigned long long, unsigned long long, long long, int, float*, cudaStream_t);
