// Unmasked multi-head self-attention over LONG token sequences on the tcgen05 GEMM (gemm_tc.cu), forward and backward.
//
// The fused tensor-core kernels of attention_tc.cu hold one (trial, head) in a single 256-key tile; the iTransformer of
// models/itransformer.py:157-173 attends over 670 neuron tokens with heads of 96 (nn.MultiheadAttention inside
// nn.TransformerEncoderLayer: no mask, dropout on the probabilities).  Here the operator is taken apart into batched GEMMs, one
// batch per (trial, head), with the L x L probability matrices materialised in HBM in bf16 -- 115 MB per layer at the
// configs[3] size, a few tens of microseconds of HBM time per pass, against 12 ms per layer for the CUDA-core kernels:
//
//   forward    S  = Q K^T / sqrt(hd)          NT, per-batch B operand (b_sel)
//              P  = softmax(S), P~ = P * keep  one warp per row (fp32 in, bf16 out), Philox stream of the CUDA-core kernels
//              O  = P~ V                       NT against V^T (packed transposed once)
//   backward   dP~ = dO V^T                    NT
//              dS  = P * (dP~ * keep - delta) / sqrt(hd),  delta_i = sum_j P_ij dP~_ij keep_ij      one warp per row
//              dQ  = dS K                      NT against K^T
//              dK  = dS^T Q,  dV = P~^T dO     TN, one split per batch routed to its own output (epi.sel)
//
// Operands are packed head-major in bf16 by one kernel (q | k | v come as columns of the in-projection's (B L, 3H) fp32 output);
// results are unpacked the same way.  Reference semantics: F.multi_head_attention_forward (torch) as called by
// nn.TransformerEncoderLayer._sa_block; the CUDA-core kernels of attention.cu are the fp32 twin (same dropout masks).
#include "kernels.cuh"
#include "gemm_common.cuh"

namespace {

constexpr int kAlign = 256;
__host__ __device__ inline long long round_up(long long v, long long a) { return (v + a - 1) / a * a; }

// qkv (B*L, 3H) fp32  ->  Qh, Kh, Vh (B*nh, L, hd) and Kt, Vt (B*nh, hd, Lp), all bf16.  One CTA per token row (grid-stride),
// one thread per column: the head / dimension split of a column is computed once per thread, not per element.
__global__ void pack_heads_kernel(const float* __restrict__ qkv, bf16* __restrict__ Qh, bf16* __restrict__ Kh, bf16* __restrict__ Vh,
                                  bf16* __restrict__ Kt, bf16* __restrict__ Vt, int B, int L, int H, int nh, int hd, int Lp) { pdl_grid_sync();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const int h = c / hd, d = c % hd;
    for (long long bl = blockIdx.x; bl < (long long)B * L; bl += gridDim.x) {
      const int l = (int)(bl % L), b = (int)(bl / L);
      const float* src = qkv + bl * 3 * H + c;
      const long long bh = (long long)b * nh + h;
      const long long hm = (bh * L + l) * hd + d, tm = (bh * hd + d) * Lp + l;
      const bf16 q = __float2bfloat16_rn(src[0]), k = __float2bfloat16_rn(src[H]), v = __float2bfloat16_rn(src[2 * H]);
      Qh[hm] = q;
      if (Kh) Kh[hm] = k;
      if (Vh) Vh[hm] = v;
      if (Kt) Kt[tm] = k;
      if (Vt) Vt[tm] = v;
    }
  }
}

// x (B*L, H) fp32 -> Xh (B*nh, L, hd) bf16
__global__ void pack_one_kernel(const float* __restrict__ x, bf16* __restrict__ Xh, int B, int L, int H, int nh, int hd) { pdl_grid_sync();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const int h = c / hd, d = c % hd;
    for (long long bl = blockIdx.x; bl < (long long)B * L; bl += gridDim.x) {
      const int l = (int)(bl % L), b = (int)(bl / L);
      Xh[(((long long)b * nh + h) * L + l) * hd + d] = __float2bfloat16_rn(x[bl * H + c]);
    }
  }
}

// Xh (B*nh, L, hd) fp32 -> out (B*L, ld) fp32 at columns col0 + h*hd + d; up to three sources in one launch (dq | dk | dv)
__global__ void unpack_heads_kernel(const float* __restrict__ X0, const float* __restrict__ X1, const float* __restrict__ X2, float* __restrict__ out,
                                    int B, int L, int H, int nh, int hd, int ld) { pdl_grid_sync();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const int h = c / hd, d = c % hd;
    for (long long bl = blockIdx.x; bl < (long long)B * L; bl += gridDim.x) {
      const int l = (int)(bl % L), b = (int)(bl / L);
      const long long src = (((long long)b * nh + h) * L + l) * hd + d;
      float* o = out + bl * ld + c;
      o[0] = X0[src];
      if (X1) o[H] = X1[src];
      if (X2) o[2 * H] = X2[src];
    }
  }
}

__global__ void iota_kernel(long long* p, int n) { pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// keep-scale of element e of the probability-dropout stream (one Philox block per 8 consecutive elements: cached across calls)
struct KeepCache {
  unsigned long long blk; Philox4 r;
  __device__ KeepCache() : blk(~0ull) {}
  __device__ __forceinline__ float get(unsigned long long seed, unsigned long long site, unsigned long long e, uint32_t thr, float ik) {
    if ((e >> 3) != blk) { blk = e >> 3; r = philox4x32(seed, blk, site); }
    return philox_u16(r, (int)(e & 7)) >= thr ? ik : 0.f;
  }
};

// One warp per row of S (rows = B*nh*L, L columns, row stride Lp, a multiple of 8): P = softmax(S) and P~ = P * keep, both bf16;
// the pad columns of a row are written as zero.  Lane l owns the 8-column groups l, l + 32, ... -- one 32-byte read, one 16-byte
// write per group, the whole row in registers between them (kMaxGroups groups per lane: rows of up to 256 * kMaxGroups columns;
// longer rows take the strided loop).  A Philox block covers 8 consecutive elements of the FLATTENED (row, column) index, which
// is a group only when the row starts on a multiple of 8: hence the two-block cache.
constexpr int kMaxGroups = 4;

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o; o.x = *(uint32_t*)&a; o.y = *(uint32_t*)&b; o.z = *(uint32_t*)&c; o.w = *(uint32_t*)&d;
  return o;
}
__device__ __forceinline__ void unpack8_bf16(const uint4 u, float* v) {
  const __nv_bfloat162* h = (const __nv_bfloat162*)&u;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ S, bf16* __restrict__ P, bf16* __restrict__ Pd, long long rows,
                                                           int L, int Lp, float p_drop, SeedRef seed, unsigned long long site) { pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + row * Lp;
  const uint32_t thr = drop_threshold(p_drop);
  const float ik = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.f;
  const unsigned long long sd = p_drop > 0.f ? seed.get() : 0ull;
  bf16* pr = P + row * Lp;
  bf16* pd = Pd ? Pd + row * Lp : nullptr;
  KeepCache kc;
  if (Lp <= 256 * kMaxGroups) {
    float v[kMaxGroups][8];
    float m = -INFINITY;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      const int j0 = (g * 32 + lane) * 8;
      if (j0 < Lp) {
        const float4 a = *(const float4*)(s + j0), b = *(const float4*)(s + j0 + 4);
        v[g][0] = a.x; v[g][1] = a.y; v[g][2] = a.z; v[g][3] = a.w; v[g][4] = b.x; v[g][5] = b.y; v[g][6] = b.z; v[g][7] = b.w;
#pragma unroll
        for (int u = 0; u < 8; ++u) { if (j0 + u >= L) v[g][u] = -INFINITY; m = fmaxf(m, v[g][u]); }
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g)
      if ((g * 32 + lane) * 8 < Lp)
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[g][u] = __expf(v[g][u] - m); sum += v[g][u]; }      // exp(-inf) = 0 for the pad columns
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      const int j0 = (g * 32 + lane) * 8;
      if (j0 < Lp) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[g][u] *= inv;
        *(uint4*)(pr + j0) = pack8_bf16(v[g]);
        if (pd) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (p_drop > 0.f && j0 + u < L) v[g][u] *= kc.get(sd, site, (unsigned long long)row * L + j0 + u, thr, ik);
          *(uint4*)(pd + j0) = pack8_bf16(v[g]);
        }
      }
    }
    return;
  }
  float m = -INFINITY;
  for (int j = lane; j < L; j += 32) m = fmaxf(m, s[j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) sum += __expf(s[j] - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j0 = lane * 8; j0 < Lp; j0 += 256) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u;
      const float pv = j < L ? __expf(s[j] - m) * inv : 0.f;
      pr[j] = __float2bfloat16_rn(pv);
      if (pd) {
        const float kp = (j < L && p_drop > 0.f) ? kc.get(sd, site, (unsigned long long)row * L + j, thr, ik) : 1.f;
        pd[j] = __float2bfloat16_rn(pv * kp);
      }
    }
  }
}

// One warp per row: dS = P * (dP~ * keep - delta) * scale with delta = sum_j P_ij dP~_ij keep_ij, bf16 out (pad columns zero).
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ dPd, const bf16* __restrict__ P, bf16* __restrict__ dS,
                                                           long long rows, int L, int Lp, float scale, float p_drop, SeedRef seed,
                                                           unsigned long long site) { pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* g = dPd + row * Lp;
  const bf16* pr = P + row * Lp;
  const uint32_t thr = drop_threshold(p_drop);
  const float ik = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.f;
  const unsigned long long sd = p_drop > 0.f ? seed.get() : 0ull;
  bf16* o = dS + row * Lp;
  if (Lp <= 256 * kMaxGroups) {
    float gk[kMaxGroups][8], pv[kMaxGroups][8];      // dP~ * keep, P
    float delta = 0.f;
    KeepCache kc;
#pragma unroll
    for (int q = 0; q < kMaxGroups; ++q) {
      const int j0 = (q * 32 + lane) * 8;
      if (j0 < Lp) {
        const float4 a = *(const float4*)(g + j0), b = *(const float4*)(g + j0 + 4);
        gk[q][0] = a.x; gk[q][1] = a.y; gk[q][2] = a.z; gk[q][3] = a.w; gk[q][4] = b.x; gk[q][5] = b.y; gk[q][6] = b.z; gk[q][7] = b.w;
        unpack8_bf16(*(const uint4*)(pr + j0), pv[q]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (j0 + u >= L) { gk[q][u] = 0.f; pv[q][u] = 0.f; }
          else if (p_drop > 0.f) gk[q][u] *= kc.get(sd, site, (unsigned long long)row * L + j0 + u, thr, ik);
          delta = fmaf(pv[q][u], gk[q][u], delta);
        }
      }
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int q = 0; q < kMaxGroups; ++q) {
      const int j0 = (q * 32 + lane) * 8;
      if (j0 < Lp) {
#pragma unroll
        for (int u = 0; u < 8; ++u) gk[q][u] = (j0 + u < L) ? pv[q][u] * (gk[q][u] - delta) * scale : 0.f;
        *(uint4*)(o + j0) = pack8_bf16(gk[q]);
      }
    }
    return;
  }
  float delta = 0.f;
  {
    KeepCache kc;
    for (int j0 = lane * 8; j0 < L; j0 += 256)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j >= L) break;
        const float kp = p_drop > 0.f ? kc.get(sd, site, (unsigned long long)row * L + j, thr, ik) : 1.f;
        delta = fmaf(__bfloat162float(pr[j]), g[j] * kp, delta);
      }
  }
  delta = warp_sum(delta);
  KeepCache kc;
  for (int j0 = lane * 8; j0 < Lp; j0 += 256)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u;
      float v = 0.f;
      if (j < L) {
        const float kp = p_drop > 0.f ? kc.get(sd, site, (unsigned long long)row * L + j, thr, ik) : 1.f;
        v = __bfloat162float(pr[j]) * (g[j] * kp - delta) * scale;
      }
      o[j] = __float2bfloat16_rn(v);
    }
}

struct Shape { int B, L, H, nh, hd, Lp; long long BH; };
Shape shape_of(int B, int L, int H, int nh) {
  Shape s; s.B = B; s.L = L; s.H = H; s.nh = nh; s.hd = H / nh; s.Lp = (int)round_up(L, 8); s.BH = (long long)B * nh;
  return s;
}
struct Carver {
  char* cur;
  explicit Carver(void* p) : cur((char*)p) {}
  template <typename T> T* take(long long n) { cur = (char*)round_up((long long)(uintptr_t)cur, kAlign); T* r = (T*)cur; cur += n * (long long)sizeof(T); return r; }
};
// buffers kept from the forward for the backward
struct Saved { bf16 *Qh, *Vh, *Kt, *P, *Pd; long long* iota; };
Saved carve_saved(void* p, const Shape& s, bool drop) {
  Carver c(p); Saved v;
  v.iota = c.take<long long>(s.BH);
  v.Qh = c.take<bf16>(s.BH * s.L * s.hd); v.Vh = c.take<bf16>(s.BH * s.L * s.hd); v.Kt = c.take<bf16>(s.BH * s.hd * s.Lp);
  v.P = c.take<bf16>(s.BH * s.L * s.Lp);
  v.Pd = drop ? c.take<bf16>(s.BH * s.L * s.Lp) : v.P;
  return v;
}
long long saved_bytes(const Shape& s, bool drop) {
  return kAlign * 8 + s.BH * 8 + (2 * s.BH * s.L * s.hd + s.BH * s.hd * s.Lp) * 2 + (drop ? 2 : 1) * s.BH * s.L * s.Lp * 2;
}
// scratch of one call: the fp32 score matrix, the bf16 dS, head-major operands / results
long long work_bytes(const Shape& s) {
  return kAlign * 10 + s.BH * s.L * s.Lp * 4 + s.BH * s.L * s.Lp * 2 + (s.BH * s.L * s.hd + s.BH * s.hd * s.Lp) * 2 + 3 * s.BH * s.L * s.hd * 4;
}

GemmProblem problem(int mode, int M, int N, int K, int nb) {
  GemmProblem p;
  p.b_sel = nullptr;
  p.mode = mode; p.M = M; p.N = N; p.nb_out = nb; p.nchunk = 1; p.chunk_k = K;
  p.a_row_shift = p.a_col_shift = p.b_row_shift = p.b_col_shift = 0; p.b_chunk_n = N; p.split_k = 1;
  p.epi = gemm_epilogue_default();
  return p;
}
GemmOperand operand(const void* ptr, long long bs, int nb, int rows, int cols, int ld) { GemmOperand o{ptr, bs, nb, rows, cols, ld}; return o; }

int rows_grid(long long rows) { return (int)(rows < 148 * 8 ? rows : 148 * 8); }
int row_threads(int H) { return H >= 1024 ? 1024 : (H + 31) / 32 * 32; }

// C (nb, M, N) fp32 = alpha * A (nb, M, K) . B (nb, N, K)^T, every batch with its own B
int gemm_nt(const bf16* A, int lda, long long sa, const bf16* Bm, int ldb, long long sb, float* C, int ldc, int M, int N, int K, int nb,
            float alpha, const long long* iota, cudaStream_t s) {
  GemmProblem p = problem(GEMM_NT, M, N, K, nb);
  p.b_sel = iota;
  p.A = operand(A, sa, nb, M, K, lda); p.B = operand(Bm, sb, nb, N, K, ldb);
  p.epi.out = C; p.epi.ldc = ldc; p.epi.c_batch_stride = (long long)M * ldc; p.epi.alpha = alpha;
  return gemm_tc_launch(p, s);
}
// C[b] (M, N) fp32 += A[b] (K, M)^T . B[b] (K, N) for every batch b (C zeroed by the caller): one split of the reduction per batch,
// routed to its own output matrix
int gemm_tn_routed(const bf16* A, int lda, long long sa, const bf16* Bm, int ldb, long long sb, float* C, int M, int N, int K, int nb,
                   const long long* iota, cudaStream_t s) {
  GemmProblem p = problem(GEMM_TN, M, N, K, 1);
  p.nchunk = nb; p.split_k = nb; p.b_chunk_n = N;
  p.A = operand(A, sa, nb, K, M, lda); p.B = operand(Bm, sb, nb, K, N, ldb);
  p.epi.out = C; p.epi.ldc = N; p.epi.accumulate = 1;
  p.epi.sel = iota; p.epi.sel_n = nb; p.epi.c_sel_stride = (long long)M * N;
  return gemm_tc_launch(p, s);
}

int check_shape(const Shape& s) {
  NDT1_REQUIRE(s.nh > 0 && s.hd * s.nh == s.H, "attention_mm: hidden %d != heads %d x head size", s.H, s.nh);
  NDT1_REQUIRE(s.hd % 8 == 0, "attention_mm: head size %d must be a multiple of 8", s.hd);
  NDT1_REQUIRE(s.L >= 1 && s.B >= 1, "attention_mm: empty problem");
  return 0;
}

}  // namespace

size_t k_attention_mm_saved_bytes(int B, int L, int H, int nh, float p_attn) { return (size_t)saved_bytes(shape_of(B, L, H, nh), p_attn > 0.f); }
size_t k_attention_mm_workspace_bytes(int B, int L, int H, int nh) { return (size_t)work_bytes(shape_of(B, L, H, nh)); }

int k_attention_mm_fwd(const float* qkv, float* out, void* saved, void* workspace, int B, int L, int H, int nh, float p_attn, SeedRef seed,
                       unsigned long long site, cudaStream_t s) {
  const Shape sh = shape_of(B, L, H, nh);
  NDT1_TRY(check_shape(sh));
  NDT1_TRY(gemm_tc_init());
  const Saved sv = carve_saved(saved, sh, p_attn > 0.f);
  Carver w(workspace);
  float* S = w.take<float>(sh.BH * L * sh.Lp);
  (void)w.take<bf16>(sh.BH * L * sh.Lp);
  bf16* Kh = w.take<bf16>(sh.BH * L * sh.hd);
  bf16* Vt = w.take<bf16>(sh.BH * sh.hd * sh.Lp);
  float* Oh = w.take<float>(sh.BH * L * sh.hd);
  const float alpha = 1.0f / sqrtf((float)sh.hd);
  ndt1_launch(iota_kernel, (int)((sh.BH + 255) / 256), 256, 0, s, sv.iota, (int)sh.BH);
  NDT1_CHECK_LAUNCH();
  NDT1_CUDA_CHECK(cudaMemsetAsync(sv.Kt, 0, (size_t)sh.BH * sh.hd * sh.Lp * 2, s));      // (pad columns L..Lp of the transposed operands)
  NDT1_CUDA_CHECK(cudaMemsetAsync(Vt, 0, (size_t)sh.BH * sh.hd * sh.Lp * 2, s));
  ndt1_launch(pack_heads_kernel, rows_grid((long long)B * L), row_threads(H), 0, s, qkv, sv.Qh, Kh, sv.Vh, sv.Kt, Vt, B, L, H, nh, sh.hd, sh.Lp);
  NDT1_CHECK_LAUNCH();
  NDT1_TRY(gemm_nt(sv.Qh, sh.hd, (long long)L * sh.hd, Kh, sh.hd, (long long)L * sh.hd, S, sh.Lp, L, L, sh.hd, (int)sh.BH, alpha, sv.iota, s));
  const long long rows = sh.BH * L;
  ndt1_launch(softmax_fwd_kernel, (int)((rows + 7) / 8), 256, 0, s, (const float*)S, sv.P, p_attn > 0.f ? sv.Pd : (bf16*)nullptr, rows, L, sh.Lp,
              p_attn, seed, site);
  NDT1_CHECK_LAUNCH();
  NDT1_TRY(gemm_nt(sv.Pd, sh.Lp, (long long)L * sh.Lp, Vt, sh.Lp, (long long)sh.hd * sh.Lp, Oh, sh.hd, L, sh.hd, L, (int)sh.BH, 1.0f, sv.iota, s));
  ndt1_launch(unpack_heads_kernel, rows_grid((long long)B * L), row_threads(H), 0, s, (const float*)Oh, (const float*)nullptr, (const float*)nullptr, out, B, L, H, nh, sh.hd, H);
  NDT1_CHECK_LAUNCH();
  return 0;
}

int k_attention_mm_bwd(const float* dout, void* saved, void* workspace, float* dqkv, int B, int L, int H, int nh, float p_attn, SeedRef seed,
                       unsigned long long site, cudaStream_t s) {
  const Shape sh = shape_of(B, L, H, nh);
  NDT1_TRY(check_shape(sh));
  const Saved sv = carve_saved(saved, sh, p_attn > 0.f);
  Carver w(workspace);
  float* dPd = w.take<float>(sh.BH * L * sh.Lp);
  bf16* dS = w.take<bf16>(sh.BH * L * sh.Lp);
  bf16* dOh = w.take<bf16>(sh.BH * L * sh.hd);
  (void)w.take<bf16>(sh.BH * sh.hd * sh.Lp);
  float* dQh = w.take<float>(sh.BH * L * sh.hd);
  float* dKh = w.take<float>(sh.BH * L * sh.hd);
  float* dVh = w.take<float>(sh.BH * L * sh.hd);
  const float alpha = 1.0f / sqrtf((float)sh.hd);
  const long long hm = (long long)L * sh.hd, pm = (long long)L * sh.Lp;
  ndt1_launch(pack_one_kernel, rows_grid((long long)B * L), row_threads(H), 0, s, dout, dOh, B, L, H, nh, sh.hd);
  NDT1_CHECK_LAUNCH();
  NDT1_CUDA_CHECK(cudaMemsetAsync(dKh, 0, (size_t)sh.BH * hm * 4, s));            // (the routed TN GEMMs accumulate)
  NDT1_CUDA_CHECK(cudaMemsetAsync(dVh, 0, (size_t)sh.BH * hm * 4, s));
  // dP~ = dO V^T
  NDT1_TRY(gemm_nt(dOh, sh.hd, hm, sv.Vh, sh.hd, hm, dPd, sh.Lp, L, L, sh.hd, (int)sh.BH, 1.0f, sv.iota, s));
  // dV = P~^T dO
  NDT1_TRY(gemm_tn_routed(sv.Pd, sh.Lp, pm, dOh, sh.hd, hm, dVh, L, sh.hd, L, (int)sh.BH, sv.iota, s));
  const long long rows = sh.BH * L;
  ndt1_launch(softmax_bwd_kernel, (int)((rows + 7) / 8), 256, 0, s, (const float*)dPd, (const bf16*)sv.P, dS, rows, L, sh.Lp, alpha, p_attn, seed, site);
  NDT1_CHECK_LAUNCH();
  // dQ = dS K,  dK = dS^T Q
  NDT1_TRY(gemm_nt(dS, sh.Lp, pm, sv.Kt, sh.Lp, (long long)sh.hd * sh.Lp, dQh, sh.hd, L, sh.hd, L, (int)sh.BH, 1.0f, sv.iota, s));
  NDT1_TRY(gemm_tn_routed(dS, sh.Lp, pm, sv.Qh, sh.hd, hm, dKh, L, sh.hd, L, (int)sh.BH, sv.iota, s));
  ndt1_launch(unpack_heads_kernel, rows_grid((long long)B * L), row_threads(H), 0, s, (const float*)dQh, (const float*)dKh, (const float*)dVh, dqkv,
              B, L, H, nh, sh.hd, 3 * H);
  NDT1_CHECK_LAUNCH();
  return 0;
}
