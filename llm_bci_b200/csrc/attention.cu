// Context-window / padding masked multi-head self-attention, forward and
// backward, flash style (scores never reach HBM), CUDA-core arithmetic in
// fp32.  Reference: NeuralAttention.forward models/ndt1.py:266-292 with the
// mask of models/ndt1.py:30-41,435-437, evaluated as a predicate
//     allowed(b,i,j) = (i == j) || (j <= i + F && j >= i - Bk && key_valid[b][j])
// instead of the (B,L,L) int64 tensor the reference materialises.
// Dropout on the probabilities (p_attn) and on the merged output (p_out) uses
// the element-indexed Philox streams of common.cuh, regenerated in backward.
//
// Tiles: 32 queries x 32 keys per step, 128 threads; thread (i = tid/4,
// quarter = tid%4) owns 8 score columns and hd/4 output columns of query i.
#include "kernels.cuh"
#include <limits.h>

namespace {

constexpr int TQ = 32, TK = 32, AT_THREADS = 128, MAX_HD = 128;

__device__ __forceinline__ bool allowed(const AttnParams& p, const long long* kv, int i, int j) {
  if (i == j) return true;
  return (j <= i + p.ctx_fwd) && (j >= i - p.ctx_bwd) && kv[j] != 0;
}

template <typename T>
__device__ __forceinline__ void load_tile(float (*dst)[MAX_HD + 1], const T* base, int row0, int L, int ld, int hd) {
  // 32 rows x hd columns, zero beyond L
  for (int e = threadIdx.x; e < TQ * hd; e += AT_THREADS) {
    const int r = e / hd, d = e % hd;
    dst[r][d] = (row0 + r < L) ? to_f32(base[(long long)(row0 + r) * ld + d]) : 0.f;
  }
}

template <typename T, int DPER>
__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const AttnParams p) { pdl_grid_sync();
  extern __shared__ float smem_f[];
  float (*Qs)[MAX_HD + 1] = (float (*)[MAX_HD + 1])smem_f;
  float (*Ks)[MAX_HD + 1] = Qs + TQ;
  float (*Vs)[MAX_HD + 1] = Ks + TK;
  float (*Ps)[TK + 1] = (float (*)[TK + 1])(Vs + TK);

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int hd = p.hd, L = p.L, ld = 3 * p.H;
  const T* qkv = (const T*)p.qkv + (long long)b * L * ld + h * hd;
  const long long* kv = p.key_valid + (long long)b * L;
  const int i = threadIdx.x >> 2, qt = threadIdx.x & 3;
  const int qi = q0 + i;
  constexpr int dper = DPER;          // output columns per thread
  const int d0 = qt * dper;

  load_tile<T>(Qs, qkv, q0, L, ld, hd);
  float m = -INFINITY, l = 0.f;
  float acc[DPER];
#pragma unroll
  for (int d = 0; d < DPER; ++d) acc[d] = 0.f;
  const uint32_t thr = drop_threshold(p.p_attn);
  const float ik = p.p_attn > 0.f ? 1.0f / (1.0f - p.p_attn) : 1.f;

  for (int k0 = 0; k0 < L; k0 += TK) {
    __syncthreads();
    load_tile<T>(Ks, qkv + p.H, k0, L, ld, hd);
    load_tile<T>(Vs, qkv + 2 * p.H, k0, L, ld, hd);
    __syncthreads();
    float s[8];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = qt * 8 + jj, kj = k0 + j;
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(Qs[i][d], Ks[j][d], a);
      const bool ok = qi < L && kj < L && allowed(p, kv, qi, kj);
      s[jj] = ok ? a * p.scale : -INFINITY;
      tmax = fmaxf(tmax, s[jj]);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m, tmax);
    const float corr = (m_new == -INFINITY) ? 1.f : expf(m - m_new);
    float psum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = qt * 8 + jj;
      float pv = (s[jj] == -INFINITY) ? 0.f : expf(s[jj] - m_new);
      psum += pv;
      if (p.p_attn > 0.f && pv != 0.f) {
        const unsigned long long e = (((unsigned long long)b * p.nh + h) * L + qi) * (unsigned long long)L + (k0 + j);
        pv *= drop_scale_1(p.seed.get(), p.stream_attn, e, thr, ik);
      }
      Ps[i][j] = pv;
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l = l * corr + psum;
    m = m_new;
    __syncwarp();
    _Pragma("unroll") for (int d = 0; d < dper; ++d) acc[d] *= corr;
    for (int j = 0; j < TK; ++j) {
      const float pv = Ps[i][j];
      _Pragma("unroll") for (int d = 0; d < dper; ++d) acc[d] = fmaf(pv, Vs[j][d0 + d], acc[d]);
    }
  }
  if (qi < L) {
    const float inv = 1.f / l;
    T* o = (T*)p.out + ((long long)b * L + qi) * p.H + h * hd + d0;
    T* od = (T*)p.out_drop + ((long long)b * L + qi) * p.H + h * hd + d0;
    const uint32_t thr_o = drop_threshold(p.p_out);
    const float iko = p.p_out > 0.f ? 1.0f / (1.0f - p.p_out) : 1.f;
    _Pragma("unroll") for (int d = 0; d < dper; ++d) {
      const float v = acc[d] * inv;
      o[d] = from_f32<T>(v);
      if (p.p_out > 0.f) {
        const unsigned long long e = ((unsigned long long)b * L + qi) * (unsigned long long)p.H + h * hd + d0 + d;
        od[d] = from_f32<T>(v * drop_scale_1(p.seed.get(), p.stream_out, e, thr_o, iko));
      } else if (od != o) {
        od[d] = from_f32<T>(v);
      }
    }
    if (qt == 0) p.lse[((long long)b * p.nh + h) * L + qi] = m + logf(l);
  }
}

// delta[b,h,i] = sum_d dO[i,d] * O[i,d]   (dO already carries the output-dropout mask)
template <typename T>
__global__ void attn_delta_kernel(const AttnParams p) { pdl_grid_sync();
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (b*L + i)*nh + h
  const int lane = threadIdx.x & 31;
  const long long total = (long long)p.B * p.L * p.nh;
  if (row >= total) return;
  const int h = (int)(row % p.nh);
  const long long bi = row / p.nh;
  const T* o = (const T*)p.out + bi * p.H + h * p.hd;
  const T* g = (const T*)p.dout + bi * p.H + h * p.hd;
  float s = 0.f;
  if (sizeof(T) == 2 && (p.hd & 3) == 0 && (p.H & 3) == 0) {      // 8-byte loads of 4 bf16
    for (int d = lane * 4; d < p.hd; d += 128) {
      const uint2 a = *(const uint2*)((const bf16*)o + d), c = *(const uint2*)((const bf16*)g + d);
      const __nv_bfloat162 a0 = *(const __nv_bfloat162*)&a.x, a1 = *(const __nv_bfloat162*)&a.y;
      const __nv_bfloat162 c0 = *(const __nv_bfloat162*)&c.x, c1 = *(const __nv_bfloat162*)&c.y;
      s += __low2float(a0) * __low2float(c0) + __high2float(a0) * __high2float(c0) + __low2float(a1) * __low2float(c1) +
           __high2float(a1) * __high2float(c1);
    }
  } else {
    for (int d = lane; d < p.hd; d += 32) s += to_f32(o[d]) * to_f32(g[d]);
  }
  s = warp_sum(s);
  const int b = (int)(bi / p.L), i = (int)(bi % p.L);
  if (lane == 0) p.delta[((long long)b * p.nh + h) * p.L + i] = s;
}

// One CTA per key tile: accumulates dK, dV over all query tiles.
template <typename T, int DPER>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_kv_kernel(const AttnParams p) { pdl_grid_sync();
  extern __shared__ float smem_f[];
  float (*Ks)[MAX_HD + 1] = (float (*)[MAX_HD + 1])smem_f;
  float (*Vs)[MAX_HD + 1] = Ks + TK;
  float (*Qs)[MAX_HD + 1] = Vs + TK;
  float (*Gs)[MAX_HD + 1] = Qs + TQ;                       // dO tile
  float (*Ps)[TK + 1] = (float (*)[TK + 1])(Gs + TQ);      // dropped probabilities
  float (*Ds)[TK + 1] = Ps + TQ;                           // dS

  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * TK;
  const int hd = p.hd, L = p.L, ld = 3 * p.H;
  const T* qkv = (const T*)p.qkv + (long long)b * L * ld + h * hd;
  const T* dout = (const T*)p.dout + (long long)b * L * p.H + h * hd;
  const long long* kv = p.key_valid + (long long)b * L;
  const float* lse = p.lse + ((long long)b * p.nh + h) * L;
  const float* delta = p.delta + ((long long)b * p.nh + h) * L;
  const int i = threadIdx.x >> 2, qt = threadIdx.x & 3;    // (query i, key quarter) while scoring; (key i, column quarter) while accumulating
  constexpr int dper = DPER; const int d0 = qt * dper;
  const uint32_t thr = drop_threshold(p.p_attn);
  const float ik = p.p_attn > 0.f ? 1.0f / (1.0f - p.p_attn) : 1.f;

  load_tile<T>(Ks, qkv + p.H, k0, L, ld, hd);
  load_tile<T>(Vs, qkv + 2 * p.H, k0, L, ld, hd);
  float dk[DPER], dv[DPER];
#pragma unroll
  for (int d = 0; d < DPER; ++d) { dk[d] = 0.f; dv[d] = 0.f; }

  for (int q0 = 0; q0 < L; q0 += TQ) {
    __syncthreads();
    load_tile<T>(Qs, qkv, q0, L, ld, hd);
    load_tile<T>(Gs, dout, q0, L, p.H, hd);
    __syncthreads();
    const int qi = q0 + i;
    const float lse_i = qi < L ? lse[qi] : 0.f;
    const float del_i = qi < L ? delta[qi] : 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = qt * 8 + jj, kj = k0 + j;
      float sv = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) { sv = fmaf(Qs[i][d], Ks[j][d], sv); dp = fmaf(Gs[i][d], Vs[j][d], dp); }
      const bool ok = qi < L && kj < L && allowed(p, kv, qi, kj);
      float pv = ok ? expf(sv * p.scale - lse_i) : 0.f;
      float dm = 1.f;
      if (p.p_attn > 0.f && ok) {
        const unsigned long long e = (((unsigned long long)b * p.nh + h) * L + qi) * (unsigned long long)L + kj;
        dm = drop_scale_1(p.seed.get(), p.stream_attn, e, thr, ik);
      }
      Ps[i][j] = pv * dm;
      Ds[i][j] = pv * (dp * dm - del_i) * p.scale;
    }
    __syncthreads();
    // accumulate for key j = i (this thread's key row), columns d0..d0+dper
    for (int qq = 0; qq < TQ; ++qq) {
      const float pd = Ps[qq][i], ds = Ds[qq][i];
      _Pragma("unroll") for (int d = 0; d < dper; ++d) {
        dv[d] = fmaf(pd, Gs[qq][d0 + d], dv[d]);
        dk[d] = fmaf(ds, Qs[qq][d0 + d], dk[d]);
      }
    }
  }
  const int kj = k0 + i;
  if (kj < L) {
    T* o = (T*)p.dqkv + ((long long)b * L + kj) * ld + h * hd + d0;
    _Pragma("unroll") for (int d = 0; d < dper; ++d) { o[p.H + d] = from_f32<T>(dk[d]); o[2 * p.H + d] = from_f32<T>(dv[d]); }
  }
}

// One CTA per query tile: dQ.
template <typename T, int DPER>
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_q_kernel(const AttnParams p) { pdl_grid_sync();
  extern __shared__ float smem_f[];
  float (*Qs)[MAX_HD + 1] = (float (*)[MAX_HD + 1])smem_f;
  float (*Gs)[MAX_HD + 1] = Qs + TQ;
  float (*Ks)[MAX_HD + 1] = Gs + TQ;
  float (*Vs)[MAX_HD + 1] = Ks + TK;
  float (*Ds)[TK + 1] = (float (*)[TK + 1])(Vs + TK);

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int hd = p.hd, L = p.L, ld = 3 * p.H;
  const T* qkv = (const T*)p.qkv + (long long)b * L * ld + h * hd;
  const T* dout = (const T*)p.dout + (long long)b * L * p.H + h * hd;
  const long long* kv = p.key_valid + (long long)b * L;
  const int i = threadIdx.x >> 2, qt = threadIdx.x & 3;
  const int qi = q0 + i;
  constexpr int dper = DPER; const int d0 = qt * dper;
  const uint32_t thr = drop_threshold(p.p_attn);
  const float ik = p.p_attn > 0.f ? 1.0f / (1.0f - p.p_attn) : 1.f;
  const float lse_i = qi < L ? p.lse[((long long)b * p.nh + h) * L + qi] : 0.f;
  const float del_i = qi < L ? p.delta[((long long)b * p.nh + h) * L + qi] : 0.f;

  load_tile<T>(Qs, qkv, q0, L, ld, hd);
  load_tile<T>(Gs, dout, q0, L, p.H, hd);
  float dq[DPER];
#pragma unroll
  for (int d = 0; d < DPER; ++d) dq[d] = 0.f;

  for (int k0 = 0; k0 < L; k0 += TK) {
    __syncthreads();
    load_tile<T>(Ks, qkv + p.H, k0, L, ld, hd);
    load_tile<T>(Vs, qkv + 2 * p.H, k0, L, ld, hd);
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = qt * 8 + jj, kj = k0 + j;
      float sv = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) { sv = fmaf(Qs[i][d], Ks[j][d], sv); dp = fmaf(Gs[i][d], Vs[j][d], dp); }
      const bool ok = qi < L && kj < L && allowed(p, kv, qi, kj);
      const float pv = ok ? expf(sv * p.scale - lse_i) : 0.f;
      float dm = 1.f;
      if (p.p_attn > 0.f && ok) {
        const unsigned long long e = (((unsigned long long)b * p.nh + h) * L + qi) * (unsigned long long)L + kj;
        dm = drop_scale_1(p.seed.get(), p.stream_attn, e, thr, ik);
      }
      Ds[i][j] = pv * (dp * dm - del_i) * p.scale;
    }
    __syncwarp();
    for (int j = 0; j < TK; ++j) {
      const float ds = Ds[i][j];
      _Pragma("unroll") for (int d = 0; d < dper; ++d) dq[d] = fmaf(ds, Ks[j][d0 + d], dq[d]);
    }
  }
  if (qi < L) {
    T* o = (T*)p.dqkv + ((long long)b * L + qi) * ld + h * hd + d0;
    _Pragma("unroll") for (int d = 0; d < dper; ++d) o[d] = from_f32<T>(dq[d]);
  }
}

int check(const AttnParams& p) {
  NDT1_REQUIRE(p.hd == 16 || p.hd == 32 || p.hd == 64 || p.hd == 96 || p.hd == 128, "attention: head size %d unsupported (16, 32, 64, 96 or 128)", p.hd);
  NDT1_REQUIRE(p.nh * p.hd == p.H, "attention: hidden %d != heads %d x head size %d", p.H, p.nh, p.hd);
  return 0;
}

}  // namespace

template <typename T>
int k_attention_fwd(const AttnParams& p, cudaStream_t stream) {
  NDT1_TRY(check(p));
  if (p.B * p.L == 0) return 0;
  const size_t smem = (size_t)(TQ + 2 * TK) * (MAX_HD + 1) * 4 + (size_t)TQ * (TK + 1) * 4;
  dim3 grid(ndt1_cdiv(p.L, TQ), p.nh, p.B);
#define NDT1_ATT_FWD(D)                                                                                                          \
  {                                                                                                                              \
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    ndt1_launch(attn_fwd_kernel<T, D>, grid, AT_THREADS, smem, stream, p);                                                                \
  }
  switch (p.hd) { case 16: NDT1_ATT_FWD(4) break; case 32: NDT1_ATT_FWD(8) break; case 64: NDT1_ATT_FWD(16) break; case 96: NDT1_ATT_FWD(24) break; default: NDT1_ATT_FWD(32) }
#undef NDT1_ATT_FWD
  NDT1_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int k_attention_bwd(const AttnParams& p, cudaStream_t stream) {
  NDT1_TRY(check(p));
  if (p.B * p.L == 0) return 0;
  const long long rows = (long long)p.B * p.L * p.nh;
  ndt1_launch(attn_delta_kernel<T>, ndt1_cdiv(rows, 8), 256, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  const size_t smem_kv = (size_t)(2 * TQ + 2 * TK) * (MAX_HD + 1) * 4 + (size_t)2 * TQ * (TK + 1) * 4;
  const size_t smem_q = (size_t)(2 * TQ + 2 * TK) * (MAX_HD + 1) * 4 + (size_t)TQ * (TK + 1) * 4;
  dim3 grid(ndt1_cdiv(p.L, TK), p.nh, p.B);
#define NDT1_ATT_BWD(D)                                                                                                          \
  {                                                                                                                              \
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kv_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv)); \
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_q_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));   \
    ndt1_launch(attn_bwd_kv_kernel<T, D>, grid, AT_THREADS, smem_kv, stream, p);                                                          \
    ndt1_launch(attn_bwd_q_kernel<T, D>, grid, AT_THREADS, smem_q, stream, p);                                                            \
  }
  switch (p.hd) { case 16: NDT1_ATT_BWD(4) break; case 32: NDT1_ATT_BWD(8) break; case 64: NDT1_ATT_BWD(16) break; case 96: NDT1_ATT_BWD(24) break; default: NDT1_ATT_BWD(32) }
#undef NDT1_ATT_BWD
  NDT1_CHECK_LAUNCH();
  return 0;
}

template <typename T>
int k_attention_delta(const AttnParams& p, cudaStream_t stream) {
  const long long rows = (long long)p.B * p.L * p.nh;
  if (rows == 0) return 0;
  ndt1_launch(attn_delta_kernel<T>, ndt1_cdiv(rows, 8), 256, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  return 0;
}
template int k_attention_delta<float>(const AttnParams&, cudaStream_t);
template int k_attention_delta<bf16>(const AttnParams&, cudaStream_t);

template int k_attention_fwd<float>(const AttnParams&, cudaStream_t);
template int k_attention_fwd<bf16>(const AttnParams&, cudaStream_t);
template int k_attention_bwd<float>(const AttnParams&, cudaStream_t);
template int k_attention_bwd<bf16>(const AttnParams&, cudaStream_t);
