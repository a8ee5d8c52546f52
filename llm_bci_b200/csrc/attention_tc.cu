// Tensor-core attention for sm_100a (bf16 mode): the masked self-attention of
// NeuralAttention.forward (models/ndt1.py:266-292) and its backward with the
// contractions on tcgen05 (accumulators in TMEM) and operands staged by TMA.
// Specialised for head size 128 and sequences of at most 256 tokens (the
// production shape is 243 tokens x 8 heads x 128): one CTA owns a 128-row tile
// and sees ALL keys at once, so the softmax is a plain two-pass row softmax
// (no online rescaling) done by 128 threads that each own one TMEM lane.
//
//   fwd    (q-tile):  S = Q K^T -> P = softmax(mask(S)) (dropout) -> smem (bf16) -> O = P V
//   bwd_q  (q-tile):  S, dP = dO V^T -> dS = P*(dP*dm - delta)*scale -> smem -> dQ = dS K
//   bwd_kv (k-tile):  for each q-tile: S^T = K Q^T, dP^T = V dO^T -> P~^T, dS^T -> smem
//                     -> dV += P~^T dO, dK += dS^T Q   (accumulated in TMEM)
//
// Every operand tile is a [rows][64 x bf16] SWIZZLE_128B tile; the same tile
// serves as a K-major operand (reduction along its columns) and as an MN-major
// operand (reduction along its rows), so nothing is ever transposed in memory.
// The mask is the predicate of attention.cu; dropout uses the same
// element-indexed Philox streams, so both implementations draw identical masks.
// Other shapes use the CUDA-core kernels of attention.cu.
#include "tc_common.cuh"
#include "kernels.cuh"

using namespace tc;

int tc_make_map(const GemmOperand& o, int box_cols, int box_rows, CUtensorMap* out);

namespace {

constexpr int HD = 128, LMAX = 256, TQ = 128, NTHREADS = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct TcAttn {
  AttnParams p;
  int cf, cb;
};

__device__ __forceinline__ bool allowed(int cf, int cb, const unsigned char* kv, int i, int j) {
  return (i == j) || ((j <= i + cf) && (j >= i - cb) && kv[j] != 0);
}

// store 8 consecutive bf16 of row `row`, columns [col, col+8) of a [128][.] operand kept as 64-column SW128 sub-tiles of 16 KB
__device__ __forceinline__ void st_row8(uint8_t* tile, int row, int col, const float* v) {
  __align__(16) bf16 o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = __float2bfloat16_rn(v[i]);
  *(uint4*)(tile + (col >> 6) * 16384 + sw128_offset(row, col & 63)) = *(const uint4*)o;
}

struct PhiloxRow {   // cached Philox block for consecutive element indices
  unsigned long long seed, stream, ctr; Philox4 r; uint32_t thr; float ik;
  __device__ __forceinline__ void init(unsigned long long s, unsigned long long st, float p) {
    seed = s; stream = st; ctr = ~0ull; thr = drop_threshold(p); ik = 1.0f / (1.0f - p);
  }
  __device__ __forceinline__ float scale(unsigned long long elem) {
    const unsigned long long c = elem >> 3;
    if (c != ctr) { ctr = c; r = philox4x32_10(seed, c, stream); }
    return philox_u16(r, (int)(elem & 7)) >= thr ? ik : 0.f;
  }
};

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map256, const TcAttn a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                 // 2 x [128 x 64]   32 KB
  uint8_t* sK = sQ + 32768;           // 2 x [256 x 64]   64 KB   (later: P as 4 x [128 x 64])
  uint8_t* sV = sK + 65536;           // 2 x [256 x 64]   64 KB
  uint8_t* sP = sK;
  unsigned char* s_kv = sV + 65536;   // [256]
  uint64_t* bars = (uint64_t*)(s_kv + 256);   // qk, v, s, o
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int j = tid; j < LMAX; j += NTHREADS) s_kv[j] = (j < L && p.key_valid[(long long)b * L + j] != 0) ? 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nks = (L + 15) / 16;            // key steps of 16 actually holding keys

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 32768 + 65536);
    for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 16384, &map128, &bars[0], h * HD + 64 * c, q0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sK + c * 32768, &map256, &bars[0], H + h * HD + 64 * c, 0, b);
    mbar_expect_tx(&bars[1], 65536);
    for (int c = 0; c < 2; ++c) tma_load_3d(sV + c * 32768, &map256, &bars[1], 2 * H + h * HD + 64 * c, 0, b);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, false, false);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint64_t ad = make_sdesc(smem_u32(sQ) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
      const uint64_t bd = make_sdesc(smem_u32(sK) + (ks >> 2) * 32768 + (ks & 3) * 32, 16, 1024);
      tc_mma_bf16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    tc_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  tc_fence_after();

  const int qi = q0 + tid;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const float sl2 = p.scale * kLog2e;
  float m = -INFINITY;
  for (int c0 = 0; c0 < LMAX; c0 += 32) {
    if (c0 >= L) break;
    uint32_t raw[32];
    tmem_ld32(trow + c0, raw);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int kj = c0 + j;
      if (qi < L && kj < L && allowed(a.cf, a.cb, s_kv, qi, kj)) m = fmaxf(m, __uint_as_float(raw[j]));
    }
  }
  const float m_s = (m == -INFINITY) ? 0.f : m * sl2;
  float l = 0.f;
  PhiloxRow ph;
  ph.init(p.seed, p.stream_attn, p.p_attn);
  const unsigned long long ebase = (((unsigned long long)b * p.nh + h) * L + qi) * (unsigned long long)L;
  for (int c0 = 0; c0 < LMAX; c0 += 32) {
    float pv[32];
    if (c0 < L) {
      uint32_t raw[32];
      tmem_ld32(trow + c0, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int kj = c0 + j;
        float v = 0.f;
        if (qi < L && kj < L && allowed(a.cf, a.cb, s_kv, qi, kj)) {
          v = exp2f(__uint_as_float(raw[j]) * sl2 - m_s);
          l += v;
          if (p.p_attn > 0.f) v *= ph.scale(ebase + kj);
        }
        pv[j] = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) pv[j] = 0.f;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) st_row8(sP, tid, c0 + g * 8, pv + g * 8);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
    for (int ks = 0; ks < nks; ++ks) {
      const uint64_t ad = make_sdesc(smem_u32(sP) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
      const uint64_t bd = make_sdesc(smem_u32(sV) + ks * 2048, 32768, 1024);
      tc_mma_bf16(tmem + 256, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    tc_commit(&bars[3]);
  }
  __syncwarp();
  mbar_wait(&bars[3], 0);
  tc_fence_after();
  const float inv = (l > 0.f) ? 1.f / l : 0.f;
  const uint32_t thr_o = drop_threshold(p.p_out);
  const float iko = p.p_out > 0.f ? 1.0f / (1.0f - p.p_out) : 1.f;
  for (int c0 = 0; c0 < HD; c0 += 32) {
    uint32_t raw[32];
    tmem_ld32(trow + 256 + c0, raw);
    tmem_ld_wait();
    if (qi < L) {
      const long long o = ((long long)b * L + qi) * H + h * HD + c0;
      __align__(16) bf16 ob[32], od[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = __uint_as_float(raw[j + i]) * inv; ob[j + i] = __float2bfloat16_rn(v[i]); }
        if (p.p_out > 0.f) {
          float ds[4];
          drop_scale_4(p.seed, p.stream_out, (unsigned long long)(o + j), thr_o, iko, ds);
#pragma unroll
          for (int i = 0; i < 4; ++i) od[j + i] = __float2bfloat16_rn(v[i] * ds[i]);
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) *(uint4*)((bf16*)p.out + o + g * 8) = *(const uint4*)(ob + g * 8);
      if (p.p_out > 0.f) {
#pragma unroll
        for (int g = 0; g < 4; ++g) *(uint4*)((bf16*)p.out_drop + o + g * 8) = *(const uint4*)(od + g * 8);
      } else if (p.out_drop != p.out) {
#pragma unroll
        for (int g = 0; g < 4; ++g) *(uint4*)((bf16*)p.out_drop + o + g * 8) = *(const uint4*)(ob + g * 8);
      }
    }
  }
  if (qi < L) p.lse[((long long)b * p.nh + h) * L + qi] = m * p.scale + logf(l);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------
// backward, query side: dQ
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_bwd_q_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map256,
                     const __grid_constant__ CUtensorMap mapdo, const TcAttn a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                 // 32 KB
  uint8_t* sdO = sQ + 32768;          // 32 KB   (sQ+sdO later: dS as 4 x [128 x 64])
  uint8_t* sK = sdO + 32768;          // 64 KB
  uint8_t* sV = sK + 65536;           // 64 KB
  uint8_t* sdS = sQ;
  unsigned char* s_kv = sV + 65536;
  uint64_t* bars = (uint64_t*)(s_kv + 256);   // loads, s, o
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int j = tid; j < LMAX; j += NTHREADS) s_kv[j] = (j < L && p.key_valid[(long long)b * L + j] != 0) ? 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nks = (L + 15) / 16;

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 32768 * 2 + 65536 * 2);
    for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 16384, &map128, &bars[0], h * HD + 64 * c, q0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sdO + c * 16384, &mapdo, &bars[0], h * HD + 64 * c, q0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sK + c * 32768, &map256, &bars[0], H + h * HD + 64 * c, 0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sV + c * 32768, &map256, &bars[0], 2 * H + h * HD + 64 * c, 0, b);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, false, false);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint32_t ko = (ks >> 2) * 32768 + (ks & 3) * 32, qo = (ks >> 2) * 16384 + (ks & 3) * 32;
      tc_mma_bf16(tmem, make_sdesc(smem_u32(sQ) + qo, 16, 1024), make_sdesc(smem_u32(sK) + ko, 16, 1024), idesc, ks > 0 ? 1u : 0u);
    }
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint32_t ko = (ks >> 2) * 32768 + (ks & 3) * 32, qo = (ks >> 2) * 16384 + (ks & 3) * 32;
      tc_mma_bf16(tmem + 256, make_sdesc(smem_u32(sdO) + qo, 16, 1024), make_sdesc(smem_u32(sV) + ko, 16, 1024), idesc, ks > 0 ? 1u : 0u);
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  const int qi = q0 + tid;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const float sl2 = p.scale * kLog2e;
  const float lse2 = qi < L ? p.lse[((long long)b * p.nh + h) * L + qi] * kLog2e : 0.f;
  const float del = qi < L ? p.delta[((long long)b * p.nh + h) * L + qi] : 0.f;
  PhiloxRow ph;
  ph.init(p.seed, p.stream_attn, p.p_attn);
  const unsigned long long ebase = (((unsigned long long)b * p.nh + h) * L + qi) * (unsigned long long)L;
  for (int c0 = 0; c0 < LMAX; c0 += 32) {
    float ds[32];
    if (c0 < L) {
      uint32_t rs[32], rp[32];
      tmem_ld32(trow + c0, rs);
      tmem_ld32(trow + 256 + c0, rp);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int kj = c0 + j;
        float v = 0.f;
        if (qi < L && kj < L && allowed(a.cf, a.cb, s_kv, qi, kj)) {
          const float pr = exp2f(__uint_as_float(rs[j]) * sl2 - lse2);
          const float dm = p.p_attn > 0.f ? ph.scale(ebase + kj) : 1.f;
          v = pr * (__uint_as_float(rp[j]) * dm - del) * p.scale;
        }
        ds[j] = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) ds[j] = 0.f;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) st_row8(sdS, tid, c0 + g * 8, ds + g * 8);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
    for (int ks = 0; ks < nks; ++ks)
      tc_mma_bf16(tmem, make_sdesc(smem_u32(sdS) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                  make_sdesc(smem_u32(sK) + ks * 2048, 32768, 1024), idesc, ks > 0 ? 1u : 0u);
    tc_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  for (int c0 = 0; c0 < HD; c0 += 32) {
    uint32_t raw[32];
    tmem_ld32(trow + c0, raw);
    tmem_ld_wait();
    if (qi < L) {
      __align__(16) bf16 ob[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ob[j] = __float2bfloat16_rn(__uint_as_float(raw[j]));
      bf16* o = (bf16*)p.dqkv + ((long long)b * L + qi) * 3 * H + h * HD + c0;
#pragma unroll
      for (int g = 0; g < 4; ++g) *(uint4*)(o + g * 8) = *(const uint4*)(ob + g * 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------
// backward, key side: dK, dV (accumulated over the query tiles in TMEM)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_bwd_kv_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap mapdo, const TcAttn a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;                 // 2 x [128 keys x 64 d]  32 KB
  uint8_t* sV = sK + 32768;
  uint8_t* sQ = sV + 32768;           // 2 x [128 q x 64 d]
  uint8_t* sdO = sQ + 32768;
  uint8_t* sPt = sdO + 32768;         // 2 x [128 keys x 64 q]
  uint8_t* sdSt = sPt + 32768;
  float* s_lse = (float*)(sdSt + 32768);   // [128]
  float* s_del = s_lse + 128;              // [128]
  unsigned char* s_kv = (unsigned char*)(s_del + 128);
  uint64_t* bars = (uint64_t*)(s_kv + 256);   // kv, q, s, acc
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int j = tid; j < LMAX; j += NTHREADS) s_kv[j] = (j < L && p.key_valid[(long long)b * L + j] != 0) ? 1 : 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nq = (L + TQ - 1) / TQ;
  const int kj = k0 + tid;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const float sl2 = p.scale * kLog2e;
  const uint32_t thr = drop_threshold(p.p_attn);
  const float ik = p.p_attn > 0.f ? 1.0f / (1.0f - p.p_attn) : 1.f;
  const float* lse_g = p.lse + ((long long)b * p.nh + h) * L;
  const float* del_g = p.delta + ((long long)b * p.nh + h) * L;

  for (int it = 0; it < nq; ++it) {
    const int q0 = it * TQ;
    const int nqs = (min(L - q0, TQ) + 15) / 16;   // query steps of 16 holding real queries
    if (it > 0) { mbar_wait(&bars[3], (it - 1) & 1); tc_fence_after(); }   // previous dV/dK MMAs done: sQ/sdO/sPt/sdSt reusable
    __syncthreads();
    if (tid < TQ) {
      const int qq = q0 + tid;
      s_lse[tid] = qq < L ? lse_g[qq] * kLog2e : 0.f;
      s_del[tid] = qq < L ? del_g[qq] : 0.f;
    }
    if (tid == 0) {
      if (it == 0) {
        mbar_expect_tx(&bars[0], 65536);
        for (int c = 0; c < 2; ++c) tma_load_3d(sK + c * 16384, &map128, &bars[0], H + h * HD + 64 * c, k0, b);
        for (int c = 0; c < 2; ++c) tma_load_3d(sV + c * 16384, &map128, &bars[0], 2 * H + h * HD + 64 * c, k0, b);
      }
      mbar_expect_tx(&bars[1], 65536);
      for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 16384, &map128, &bars[1], h * HD + 64 * c, q0, b);
      for (int c = 0; c < 2; ++c) tma_load_3d(sdO + c * 16384, &mapdo, &bars[1], h * HD + 64 * c, q0, b);
      if (it == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], it & 1);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, false);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
        tc_mma_bf16(tmem, make_sdesc(smem_u32(sK) + o, 16, 1024), make_sdesc(smem_u32(sQ) + o, 16, 1024), idesc, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
        tc_mma_bf16(tmem + 128, make_sdesc(smem_u32(sV) + o, 16, 1024), make_sdesc(smem_u32(sdO) + o, 16, 1024), idesc, ks > 0 ? 1u : 0u);
      }
      tc_commit(&bars[2]);
    }
    __syncthreads();                 // s_lse / s_del visible
    mbar_wait(&bars[2], it & 1);
    tc_fence_after();
    for (int c0 = 0; c0 < TQ; c0 += 32) {
      float pt[32], dst[32];
      if (q0 + c0 < L) {
        uint32_t rs[32], rp[32];
        tmem_ld32(trow + c0, rs);
        tmem_ld32(trow + 128 + c0, rp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qi = q0 + c0 + j;
          float pd = 0.f, ds = 0.f;
          if (qi < L && kj < L && allowed(a.cf, a.cb, s_kv, qi, kj)) {
            const float pr = exp2f(__uint_as_float(rs[j]) * sl2 - s_lse[c0 + j]);
            float dm = 1.f;
            if (p.p_attn > 0.f) {
              const unsigned long long e = (((unsigned long long)b * p.nh + h) * L + qi) * (unsigned long long)L + kj;
              dm = drop_scale_1(p.seed, p.stream_attn, e, thr, ik);
            }
            pd = pr * dm;
            ds = pr * (__uint_as_float(rp[j]) * dm - s_del[c0 + j]) * p.scale;
          }
          pt[j] = pd; dst[j] = ds;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) { pt[j] = 0.f; dst[j] = 0.f; }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) { st_row8(sPt, tid, c0 + g * 8, pt + g * 8); st_row8(sdSt, tid, c0 + g * 8, dst + g * 8); }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
      for (int ks = 0; ks < nqs; ++ks) {
        const uint32_t ao = (ks >> 2) * 16384 + (ks & 3) * 32;
        tc_mma_bf16(tmem + 256, make_sdesc(smem_u32(sPt) + ao, 16, 1024), make_sdesc(smem_u32(sdO) + ks * 2048, 16384, 1024), idesc,
                    (it > 0 || ks > 0) ? 1u : 0u);
      }
      for (int ks = 0; ks < nqs; ++ks) {
        const uint32_t ao = (ks >> 2) * 16384 + (ks & 3) * 32;
        tc_mma_bf16(tmem + 384, make_sdesc(smem_u32(sdSt) + ao, 16, 1024), make_sdesc(smem_u32(sQ) + ks * 2048, 16384, 1024), idesc,
                    (it > 0 || ks > 0) ? 1u : 0u);
      }
      tc_commit(&bars[3]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[3], (nq - 1) & 1);
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {       // 0: dV (cols 256..), 1: dK (cols 384..)
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t raw[32];
      tmem_ld32(trow + 256 + half * 128 + c0, raw);
      tmem_ld_wait();
      if (kj < L) {
        __align__(16) bf16 ob[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) ob[j] = __float2bfloat16_rn(__uint_as_float(raw[j]));
        bf16* o = (bf16*)p.dqkv + ((long long)b * L + kj) * 3 * H + (half == 0 ? 2 * H : H) + h * HD + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) *(uint4*)(o + g * 8) = *(const uint4*)(ob + g * 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int make_maps(const AttnParams& p, CUtensorMap* m128, CUtensorMap* m256, CUtensorMap* mdo) {
  GemmOperand q; q.ptr = p.qkv; q.batch_stride = (long long)p.L * 3 * p.H; q.nbatch = p.B; q.rows = p.L; q.cols = 3 * p.H; q.ld = 3 * p.H;
  NDT1_TRY(tc_make_map(q, 64, 128, m128));
  if (m256) NDT1_TRY(tc_make_map(q, 64, 256, m256));
  if (mdo) {
    GemmOperand d; d.ptr = p.dout; d.batch_stride = (long long)p.L * p.H; d.nbatch = p.B; d.rows = p.L; d.cols = p.H; d.ld = p.H;
    NDT1_TRY(tc_make_map(d, 64, 128, mdo));
  }
  return 0;
}

constexpr int SMEM_FWD = 32768 + 65536 + 65536 + 256 + 64 + 1024;
constexpr int SMEM_BQ = 32768 * 2 + 65536 * 2 + 256 + 64 + 1024;
constexpr int SMEM_BKV = 32768 * 6 + 1024 + 256 + 64 + 1024;

}  // namespace

bool k_attention_tc_supported(const AttnParams& p) { return p.hd == HD && p.L <= LMAX && p.L >= 1 && p.H % 8 == 0; }

int k_attention_tc_fwd(const AttnParams& p, cudaStream_t stream) {
  NDT1_REQUIRE(k_attention_tc_supported(p), "attention_tc: unsupported shape (head size %d, %d tokens)", p.hd, p.L);
  NDT1_TRY(gemm_tc_init());
  CUtensorMap m128, m256;
  NDT1_TRY(make_maps(p, &m128, &m256, nullptr));
  static bool attr = false;
  if (!attr) { NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD)); attr = true; }
  TcAttn a; a.p = p; a.cf = p.ctx_fwd; a.cb = p.ctx_bwd;
  dim3 grid(ndt1_cdiv(p.L, TQ), p.nh, p.B);
  attn_tc_fwd_kernel<<<grid, NTHREADS, SMEM_FWD, stream>>>(m128, m256, a);
  NDT1_CHECK_LAUNCH();
  return 0;
}

int k_attention_tc_bwd(const AttnParams& p, cudaStream_t stream) {
  NDT1_REQUIRE(k_attention_tc_supported(p), "attention_tc: unsupported shape (head size %d, %d tokens)", p.hd, p.L);
  NDT1_TRY(gemm_tc_init());
  CUtensorMap m128, m256, mdo;
  NDT1_TRY(make_maps(p, &m128, &m256, &mdo));
  static bool attr = false;
  if (!attr) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BQ));
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BKV));
    attr = true;
  }
  NDT1_TRY(k_attention_delta<bf16>(p, stream));
  TcAttn a; a.p = p; a.cf = p.ctx_fwd; a.cb = p.ctx_bwd;
  dim3 grid(ndt1_cdiv(p.L, TQ), p.nh, p.B);
  attn_tc_bwd_kv_kernel<<<grid, NTHREADS, SMEM_BKV, stream>>>(m128, mdo, a);
  NDT1_CHECK_LAUNCH();
  attn_tc_bwd_q_kernel<<<grid, NTHREADS, SMEM_BQ, stream>>>(m128, m256, mdo, a);
  NDT1_CHECK_LAUNCH();
  return 0;
}
