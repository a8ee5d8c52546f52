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
              const float* __restrict__ rstd, float* __restrict__ dres, T* __restrict__ out_lp, float drop_p, unsigned long long seed,
              unsigned long long stream_id, long long rows, int H, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ colsum_out) {
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
                    float* partials, cudaStream_t stream, float* colsum_out) {
  NDT1_REQUIRE(H % 4 == 0 && H <= 128 * LN_MAXV, "layernorm: hidden size %d unsupported (multiple of 4, <= %d)", H, 128 * LN_MAXV);
  if (rows == 0) return 0;
  (void)partials;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  long long nb = (rows + LNB_WARPS - 1) / LNB_WARPS;
  if (nb > sms) nb = sms;
  const size_t smem = (size_t)LNB_WARPS * 3 * H * sizeof(float);
  static size_t attr[2] = {0, 0};
  if (smem > 48 * 1024 && smem > attr[sizeof(T) == 2]) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[sizeof(T) == 2] = smem;
  }
  ln_bwd_kernel<T><<<(int)nb, LNB_WARPS * 32, smem, stream>>>(dy, x, gamma, mean, rstd, dres, out_lp, drop_p, seed,
                                                                               stream_id, rows, H, dgamma, dbeta, out_lp ? colsum_out : nullptr);
  NDT1_CHECK_LAUNCH();
  return 0;
}

template int k_layernorm_fwd<float>(const float*, const float*, const float*, float*, float*, float*, long long, int, float, cudaStream_t);
template int k_layernorm_fwd<bf16>(const float*, const float*, const float*, bf16*, float*, float*, long long, int, float, cudaStream_t);
template int k_layernorm_bwd<float>(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*, float,
                                    unsigned long long, unsigned long long, long long, int, float*, cudaStream_t, float*);
template int k_layernorm_bwd<bf16>(const bf16*, const float*, const float*, const float*, const float*, float*, float*, float*, bf16*, float,
                                   unsigned long long, unsigned long long, long long, int, float*, cudaStream_t, float*);
