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
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"
#include "kernels.cuh"

using namespace tc;

int tc_make_map(const GemmOperand& o, int box_cols, int box_rows, CUtensorMap* out);

namespace {

constexpr int HD = 128, LMAX = 256, TQ = 128, NTHREADS = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct TcAttn {
  AttnParams p;
  int cf, cb;
  unsigned long long* dbg;      // optional phase timeline (tools/attn_timeline.py): 32 slots per CTA, globaltimer ns
};

__device__ __forceinline__ void dbg_mark(const TcAttn& a, int slot) {
  if (a.dbg) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    a.dbg[cta * 32 + slot] = t;
  }
}
__device__ __forceinline__ void dbg_mark64(const TcAttn& a, int slot) {       // persistent kernels: 64 slots per CTA (1-D grid)
  if (a.dbg && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[65536 + (long long)blockIdx.x * 64 + slot] = t;        // (behind the 3-D-grid kernels' 32 slots x up to 2048 CTAs)
  }
}
__device__ __forceinline__ void dbg_smid(const TcAttn& a, int slot) {
  if (a.dbg) {
    unsigned int sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    a.dbg[cta * 32 + slot] = sm;
  }
}

// ---------------------------------------------------------------------------
// Small device helpers.  The elementwise part of every kernel is what bounds it (the
// contractions are ~2K tensor-core cycles per CTA), so the mask is evaluated as 32-bit
// words (one bit per key) and the dropout mask of the probabilities is drawn ONCE in the
// forward (Philox, 16 bits per element) and kept as a bit matrix (B, heads, L, 8 x u32)
// that both backward kernels read.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// bits j of [lo, hi] intersected with [0, 31]
__device__ __forceinline__ uint32_t range_bits(int lo, int hi) {
  lo = max(lo, 0); hi = min(hi, 31);
  return (lo > hi) ? 0u : ((0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo));
}
// allowed keys [c0, c0+32) of query i:  (i == j) || (j <= i + cf && j >= i - cb && key_valid[j])
__device__ __forceinline__ uint32_t query_mask_word(int i, int c0, int cf, int cb, uint32_t kv_word, int L) {
  if (i >= L) return 0u;
  uint32_t w = kv_word & range_bits(i - cb - c0, i + cf - c0);
  const int d = i - c0;
  if (d >= 0 && d < 32) w |= 1u << d;
  return w;
}
// allowed queries [c0, c0+32) of key j (the same predicate, transposed)
__device__ __forceinline__ uint32_t key_mask_word(int j, int c0, int cf, int cb, bool kv_j, int L) {
  if (j >= L) return 0u;
  uint32_t w = kv_j ? range_bits(j - cf - c0, min(j + cb, L - 1) - c0) : 0u;
  const int d = j - c0;
  if (d >= 0 && d < 32) w |= 1u << d;
  return w;
}
__device__ __forceinline__ uint32_t keep_bits8(const Philox4& r, uint32_t thr) {
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= thr) ? 1u : 0u;  m |= ((r.x >> 16) >= thr) ? 2u : 0u;
  m |= ((r.y & 0xFFFFu) >= thr) ? 4u : 0u;  m |= ((r.y >> 16) >= thr) ? 8u : 0u;
  m |= ((r.z & 0xFFFFu) >= thr) ? 16u : 0u; m |= ((r.z >> 16) >= thr) ? 32u : 0u;
  m |= ((r.w & 0xFFFFu) >= thr) ? 64u : 0u; m |= ((r.w >> 16) >= thr) ? 128u : 0u;
  return m;
}
// keep bits of the 32 consecutive elements [e0, e0+32) of a dropout site (any alignment); same draws as drop_scale_1
__device__ __forceinline__ uint32_t keep_word32(unsigned long long seed, unsigned long long stream, unsigned long long e0, uint32_t thr,
                                                bool aligned) {
  const unsigned long long c = e0 >> 3;
  const int sh = (int)(e0 & 7);
  unsigned long long bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) bits |= (unsigned long long)keep_bits8(philox4x32(seed, c + i, stream), thr) << (8 * i);
  if (!aligned) bits |= (unsigned long long)keep_bits8(philox4x32(seed, c + 4, stream), thr) << 32;
  return (uint32_t)(bits >> sh);
}
// 16-byte store of 8 bf16 at (row, col % 8 == 0) of a [128][.] operand kept as 64-column SW128 sub-tiles of 16 KB
__device__ __forceinline__ void st_row8(uint8_t* tile, int row, int col, const float* v) {
  const uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  *(uint4*)(tile + (col >> 6) * 16384 + row * 128 + (((((col & 63) >> 3) ^ row) & 7) << 4)) = o;
}
__device__ __forceinline__ void st_global8(bf16* dst, const float* v) {
  *(uint4*)dst = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ---------------------------------------------------------------------------
// Dropout keep bits, all layers of a forward in one launch (one thread per 32-bit word).
//   bits_p[l] (B, nh, L, 8): bit j%32 of word j/32 of row (b, h, i) = probability (i, j) kept   (element index (row * L + j))
//   bits_o[l] (B*L, H/32)  : bit c%32 of word c/32 of row r = output element (r, c) kept        (element index r * H + c)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_dropbits_kernel(const AttnBitsJob job, SeedRef seed_ref, long long rows_p, int L, long long rows_o, int H,
                                                            uint32_t thr_p, uint32_t thr_o) { pdl_grid_sync();
  const unsigned long long seed = seed_ref.get();
  const int l = blockIdx.y;
  const long long words_p = job.bits_p[l] ? rows_p * 8 : 0, words_o = job.bits_o[l] ? rows_o * (H / 32) : 0;
  const long long words_m = job.bits_m[l] ? rows_o * (H / 32) : 0;
  const bool aligned = (L & 7) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words_p + words_o + words_m; i += (long long)gridDim.x * blockDim.x) {
    if (i < words_p) {
      const long long row = i >> 3; const int c0 = (int)(i & 7) * 32;
      uint32_t w = 0u;
      if (c0 < L) {
        w = keep_word32(seed, job.stream_p[l], (unsigned long long)row * (unsigned long long)L + c0, thr_p, aligned);
        if (c0 + 32 > L) w &= 0xFFFFFFFFu >> (c0 + 32 - L);
      }
      job.bits_p[l][i] = w;
    } else if (i < words_p + words_o) {
      const long long j = i - words_p;
      job.bits_o[l][j] = keep_word32(seed, job.stream_o[l], (unsigned long long)j * 32ull, thr_o, true);
    } else {
      const long long j = i - words_p - words_o;
      job.bits_m[l][j] = keep_word32(seed, job.stream_m[l], (unsigned long long)j * 32ull, thr_o, true);      // (same p as the output dropout)
    }
  }
}

// ---------------------------------------------------------------------------
// forward.  256 threads: warp w owns TMEM lanes 32*(w%4).. (one query row per lane) and the
// key columns [128*(w/4), +128) of that row; the two halves of a row meet through smem.
//
// Two CTAs share an SM (the kernel is a dependent chain load -> S -> softmax -> PV -> store, so a
// second resident CTA is what hides it): shared memory holds only Q and ONE 64 KB K/V buffer (V
// is fetched into it once S = Q K^T has retired), and tensor memory holds 256 columns:
//   S  fp32  columns [0, 256)
//   P  bf16  packed two keys per column, written IN PLACE behind each thread's read pointer.  The thread of keys
//            [0,128) walks its columns upwards and packs into [0,64); the thread of keys [128,256) walks DOWNWARDS and
//            packs into [192,256): what stays free is ONE contiguous block
//   O  fp32  columns [64, 192): a single 128 x 128 accumulator, so P V is 16 MMAs of N = 128 (an MMA costs ~60 ns to
//            issue whatever its N: two N = 64 halves would take twice as long)
// P never touches shared memory: the P V product takes its A operand from tensor memory.  Nothing here draws a random
// number: the keep bits of both dropout sites come from attn_dropbits_kernel (1 / (1 - p) is folded into the final scale).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map256,
                   const __grid_constant__ CUtensorMap mapo, const __grid_constant__ CUtensorMap mapod, const TcAttn a) { pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                 // 2 x [128 x 64]   32 KB
  uint8_t* sKV = sQ + 32768;          // 2 x [256 x 64]   64 KB   K, then V
  float* s_m = (float*)(sKV + 65536); // [2][128] row maxima of the two column halves
  float* s_l = s_m + 256;             // [2][128] row sums
  uint32_t* s_kvw = (uint32_t*)(s_l + 256);   // [8] key_valid bits
  uint64_t* bars = (uint64_t*)(s_kvw + 8);    // qk, v, s, o
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  if (tid == 0) {
    dbg_mark(a, 0); dbg_smid(a, 31);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bars[0], 32768 + 65536);               // the loads fly while the CTA sets itself up
    for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 16384, &map128, &bars[0], h * HD + 64 * c, q0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sKV + c * 32768, &map256, &bars[0], H + h * HD + 64 * c, 0, b);
  }
  {
    const int j = warp * 32 + lane;
    const uint32_t w = __ballot_sync(0xffffffffu, j < L && p.key_valid[(long long)b * L + j] != 0);
    if (lane == 0) s_kvw[warp] = w;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nks = (L + 15) / 16;            // key steps of 16 actually holding keys
  // (Measured, round 2: starting the second co-resident CTA of every SM ~4 us late, so that one CTA's load / MMA / store phases
  //  fall under the other's softmax, changes nothing -- 28.0 vs 29.6 us kernel span, no change in a step: the two CTAs do not
  //  contend, each is bound by its own chain of dependent latencies.)
  if (tid == 0) {
    dbg_mark(a, 1);                         // set-up done
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    dbg_mark(a, 2);                         // Q, K landed
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, false, false);
    const uint64_t qd = make_sdesc(smem_u32(sQ), 16, 1024), kd = make_sdesc(smem_u32(sKV), 16, 1024);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      tc_mma_bf16(tmem, sdesc_advance(qd, (ks >> 2) * 16384 + (ks & 3) * 32), sdesc_advance(kd, (ks >> 2) * 32768 + (ks & 3) * 32), idesc,
                  ks > 0 ? 1u : 0u);
    tc_commit(&bars[2]);
  }
  __syncwarp();

  // everything that does not need S is fetched while the tensor core forms it
  const int row = quarter * 32 + lane;
  const int qi = q0 + row;
  const int cbase = half * 128;
  const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
  const float sl2 = p.scale * kLog2e;
  const bool drop = p.p_attn > 0.f;
  const long long bh_row = ((long long)b * p.nh + h) * L + qi;
  uint32_t mw[4], kwv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    mw[c] = query_mask_word(qi, cbase + 32 * c, a.cf, a.cb, s_kvw[half * 4 + c], L);
    kwv[c] = (drop && qi < L) ? __ldg(p.drop_bits + bh_row * 8 + half * 4 + c) : 0xFFFFFFFFu;
  }
  uint32_t ow[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};          // keep bits of the output dropout: head-dim columns [64 half + 32 cc, +32)
  if (p.p_out > 0.f && qi < L) {
    const unsigned int* wo = p.drop_bits_o + ((long long)b * L + qi) * (H / 32) + (h * HD + half * 64) / 32;
    ow[0] = __ldg(wo); ow[1] = __ldg(wo + 1);
  }

  mbar_wait(&bars[2], 0);
  tc_fence_after();
  if (tid == 0) {                           // K is consumed: V takes its place while the softmax runs
    dbg_mark(a, 3);                         // S ready
    mbar_expect_tx(&bars[1], 65536);
    for (int c = 0; c < 2; ++c) tma_load_3d(sKV + c * 32768, &map256, &bars[1], 2 * H + h * HD + 64 * c, 0, b);
  }

  float m = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int c0 = cbase + 32 * c;
    if (c0 < L) {                                     // warp-uniform
      uint32_t raw[32];
      tmem_ld32(trow + c0, raw);
      tmem_ld_wait();
      const uint32_t w = mw[c];
      if (w == 0xFFFFFFFFu) {                         // (the common case: nothing masked in this word)
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(raw[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, (w >> j) & 1u ? __uint_as_float(raw[j]) : -INFINITY);
      }
    }
  }
  s_m[half * 128 + row] = m;
  __syncthreads();
  if (tid == 0) dbg_mark(a, 4);             // row maxima done
  m = fmaxf(s_m[row], s_m[128 + row]);
  const float m_s = (m == -INFINITY) ? 0.f : m * sl2;
  float l = 0.f;
#pragma unroll 1
  for (int ci = 0; ci < 4; ++ci) {
    const int c = half ? 3 - ci : ci;                 // keys [128,256): downwards, so that the packed P trails the read pointer from above
    const int c0 = cbase + 32 * c;
    uint32_t pk[16];
    if (c0 < L) {
      uint32_t raw[32];
      tmem_ld32(trow + c0, raw);
      tmem_ld_wait();
      const uint32_t w = mw[c], kw = kwv[c];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float s0 = fmaf(__uint_as_float(raw[j]), sl2, -m_s), s1 = fmaf(__uint_as_float(raw[j + 1]), sl2, -m_s);
        if (w != 0xFFFFFFFFu) { s0 = (w >> j) & 1u ? s0 : -INFINITY; s1 = (w >> (j + 1)) & 1u ? s1 : -INFINITY; }
        const float e0 = ex2f(s0), e1 = ex2f(s1);
        l += e0 + e1;
        uint32_t pr = pack_bf16x2(e0, e1);
        if (drop) {       // dropped probabilities become +0 (all-ones / all-zeros half-word masks from the two keep bits)
          const uint32_t lo = (uint32_t)((int32_t)(kw << (31 - j)) >> 31), hi = (uint32_t)((int32_t)(kw << (30 - j)) >> 31);
          pr &= (lo & 0x0000FFFFu) | (hi & 0xFFFF0000u);
        }
        pk[j >> 1] = pr;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = 0u;
    }
    // P in place: 16 packed columns behind this thread's read pointer (see the layout above)
    if (c0 < nks * 16) tmem_st16(trow + (half ? 192 : 0) + 16 * c, pk);
  }
  tmem_st_wait();
  s_l[half * 128 + row] = l;
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    dbg_mark(a, 5);                         // P written
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    dbg_mark(a, 6);                         // V landed
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
    const uint64_t vd = make_sdesc(smem_u32(sKV), 32768, 1024);       // V as the MN-major operand: two 64-column chunks 32 KB apart
    for (int ks = 0; ks < nks; ++ks) {
      const uint32_t pa = tmem + (ks < 8 ? ks * 8 : 192 + (ks - 8) * 8);
      tc_mma_bf16_ts(tmem + 64, pa, sdesc_advance(vd, ks * 2048), idesc, ks > 0 ? 1u : 0u);
    }
    tc_commit(&bars[3]);
  }
  __syncwarp();
  l = s_l[row] + s_l[128 + row];
  mbar_wait(&bars[3], 0);
  tc_fence_after();
  if (tid == 0) dbg_mark(a, 7);             // O ready
  const float inv = (l > 0.f) ? 1.f / l : 0.f;
  const float inv_p = drop ? inv / (1.0f - p.p_attn) : inv;          // the 1 / (1 - p) of the probability dropout
  const float iko = p.p_out > 0.f ? 1.0f / (1.0f - p.p_out) : 1.f;
  // V is dead: its buffer stages the output tile (and its dropped copy) as SWIZZLE_128B tiles for bulk tensor stores
  // (coalesced, asynchronous; rows past the sequence end are clipped by the tensor map)
  uint8_t* sOut = sKV; uint8_t* sOutD = sKV + 32768;
  const bool two = p.out_drop != p.out;
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    const int c0 = half * 64 + cc * 32;               // head-dim column
    uint32_t raw[32];
    tmem_ld32(trow + 64 + c0, raw);
    tmem_ld_wait();
    const uint32_t wo = ow[cc];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[g * 8 + i]) * inv_p;
      st_row8(sOut, row, c0 + g * 8, v);
      if (two) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (wo >> (g * 8 + i)) & 1u ? v[i] * iko : 0.f;
        st_row8(sOutD, row, c0 + g * 8, v);
      }
    }
  }
  fence_async_smem();
  if (half == 0 && qi < L) p.lse[bh_row] = m * p.scale + logf(l);
  tc_fence_before();
  __syncthreads();
  if (tid == 32) {
    dbg_mark(a, 8);                         // output tiles staged
    for (int c = 0; c < 2; ++c) {
      tma_store_3d(&mapo, sOut + c * 16384, h * HD + 64 * c, q0, b);
      if (two) tma_store_3d(&mapod, sOutD + c * 16384, h * HD + 64 * c, q0, b);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    dbg_mark(a, 9);                         // stores have read shared memory
  }
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 256); if (lane == 0) dbg_mark(a, 16); }
}

// ---------------------------------------------------------------------------
// backward, query side: dQ of one 128-query tile against all (<= 256) keys.
// 512 threads: warp w owns TMEM lanes 32*(w%4).. (one query row per lane) and the 64 key columns [64*(w/4), +64).
//   S = Q K^T -> columns [0,256), dP = dO V^T -> columns [256,512)   (two 128 x 256 x 16 MMAs per k-step: the largest
//   shape per instruction -- a tcgen05.mma costs ~60 ns to issue whatever its N);
//   dS = P (dP dm - delta) scale goes back IN PLACE as bf16 (two keys per column) into the first 32 columns of each
//   thread's own 64-column range of S; dQ = dS K takes it from tensor memory and accumulates over [256,384) (dP is dead);
//   the tile leaves through shared memory (Q's buffer) and one bulk tensor store.
// ---------------------------------------------------------------------------
constexpr int BQ_THREADS = 512;
__global__ void __launch_bounds__(BQ_THREADS, 1)
attn_tc_bwd_q_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map256,
                     const __grid_constant__ CUtensorMap mapdo, const __grid_constant__ CUtensorMap mapo,
                     const __grid_constant__ CUtensorMap mapdqkv, const TcAttn a) { pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                 // 32 KB   (later: the dQ tile for the bulk store)
  uint8_t* sdO = sQ + 32768;          // 32 KB
  uint8_t* sK = sdO + 32768;          // 64 KB
  uint8_t* sV = sK + 65536;           // 64 KB
  uint8_t* sO = sV + 65536;           // 32 KB   forward output of the tile: delta = rowsum(dO * O) is formed here, not by a kernel of its own
  float* s_part = (float*)sQ;                 // [4][128] partial row sums (Q is dead by the time they are formed)
  uint32_t* s_kvw = (uint32_t*)(sO + 32768);  // [8]
  uint64_t* bars = (uint64_t*)(s_kvw + 8);    // qk, dov, s, o
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, grp = warp >> 2;          // grp: which 64 of the 256 key columns
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bars[0], 32768 + 65536);               // the loads fly while the CTA sets itself up
    for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 16384, &map128, &bars[0], h * HD + 64 * c, q0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sK + c * 32768, &map256, &bars[0], H + h * HD + 64 * c, 0, b);
    mbar_expect_tx(&bars[1], 32768 + 65536 + 32768);
    for (int c = 0; c < 2; ++c) tma_load_3d(sdO + c * 16384, &mapdo, &bars[1], h * HD + 64 * c, q0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sV + c * 32768, &map256, &bars[1], 2 * H + h * HD + 64 * c, 0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sO + c * 16384, &mapo, &bars[1], h * HD + 64 * c, q0, b);
  }
  if (warp < 8) {
    const int j = warp * 32 + lane;
    const uint32_t w = __ballot_sync(0xffffffffu, j < L && p.key_valid[(long long)b * L + j] != 0);
    if (lane == 0) s_kvw[warp] = w;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nks = (L + 15) / 16;

  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, false, false);
    const uint64_t qd = make_sdesc(smem_u32(sQ), 16, 1024), kd = make_sdesc(smem_u32(sK), 16, 1024);
    const uint64_t od = make_sdesc(smem_u32(sdO), 16, 1024), vd = make_sdesc(smem_u32(sV), 16, 1024);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      tc_mma_bf16(tmem, sdesc_advance(qd, (ks >> 2) * 16384 + (ks & 3) * 32), sdesc_advance(kd, (ks >> 2) * 32768 + (ks & 3) * 32), idesc, ks > 0 ? 1u : 0u);
    mbar_wait(&bars[1], 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      tc_mma_bf16(tmem + 256, sdesc_advance(od, (ks >> 2) * 16384 + (ks & 3) * 32), sdesc_advance(vd, (ks >> 2) * 32768 + (ks & 3) * 32), idesc,
                  ks > 0 ? 1u : 0u);
    tc_commit(&bars[2]);
  }
  __syncwarp();

  const int row = quarter * 32 + lane;
  const int qi = q0 + row;
  const int cbase = grp * 64;
  const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
  const float sl2 = p.scale * kLog2e;
  const long long bh_row = ((long long)b * p.nh + h) * L + qi;
  const float lse2 = qi < L ? p.lse[bh_row] * kLog2e : 0.f;
  const bool drop = p.p_attn > 0.f;
  const float ik = drop ? 1.0f / (1.0f - p.p_attn) : 1.f;
  uint32_t mw[2], kw[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int c0 = cbase + 32 * c;
    mw[c] = query_mask_word(qi, c0, a.cf, a.cb, s_kvw[grp * 2 + c], L);
    kw[c] = (drop && qi < L && c0 < L) ? p.drop_bits[bh_row * 8 + (c0 >> 5)] : 0xFFFFFFFFu;
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  // delta[row] = sum_d dO[row, d] * O[row, d] from the two tiles in shared memory (after S and dP: Q's buffer, now dead, carries the partial sums):
  // the four threads of a row (grp 0..3) take 32 head-dim columns each and meet through s_part
  float del;
  {
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = grp * 32 + g * 8;
      const uint32_t off = (uint32_t)((col >> 6) * 16384 + row * 128 + (((((col & 63) >> 3) ^ row) & 7) << 4));
      const uint4 x = *(const uint4*)(sdO + off), y = *(const uint4*)(sO + off);
      const __nv_bfloat162* xa = (const __nv_bfloat162*)&x; const __nv_bfloat162* ya = (const __nv_bfloat162*)&y;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += __low2float(xa[i]) * __low2float(ya[i]) + __high2float(xa[i]) * __high2float(ya[i]);
    }
    s_part[grp * 128 + row] = acc;
    __syncthreads();
    del = s_part[row] + s_part[128 + row] + s_part[256 + row] + s_part[384 + row];
    if (grp == 0 && qi < L) p.delta[bh_row] = del;          // the key-side kernel (launched after this one) reads it
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int c0 = cbase + 32 * c;
    uint32_t ds[16];
    if (c0 < L) {                                            // warp-uniform
      uint32_t rs[32], rp[32];
      tmem_ld32(trow + c0, rs);
      tmem_ld32(trow + 256 + c0, rp);
      tmem_ld_wait();
      const uint32_t m = mw[c], k = kw[c];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float sv = fmaf(__uint_as_float(rs[j + u]), sl2, -lse2);
          const float pr = ex2f((m >> (j + u)) & 1u ? sv : -INFINITY) * p.scale;
          const float t = (k >> (j + u)) & 1u ? __uint_as_float(rp[j + u]) * ik : 0.f;
          v[u] = pr * (t - del);
        }
        ds[j >> 1] = pack_bf16x2(v[0], v[1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) ds[j] = 0u;
    }
    if (c0 < nks * 16) tmem_st16(trow + cbase + 16 * c, ds);   // keys [c0, c0+32) -> 16 packed columns inside this thread's own range
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, true);
    const uint64_t km = make_sdesc(smem_u32(sK), 32768, 1024);          // K as the MN-major operand (rows = keys)
    for (int ks = 0; ks < nks; ++ks)                                    // 16 keys = 8 packed columns: group ks / 4, offset 8 (ks % 4)
      tc_mma_bf16_ts(tmem + 256, tmem + (ks >> 2) * 64 + (ks & 3) * 8, sdesc_advance(km, ks * 2048), idesc, ks > 0 ? 1u : 0u);
    tc_commit(&bars[3]);
  }
  __syncwarp();
  mbar_wait(&bars[3], 0);
  tc_fence_after();
  {
    uint32_t raw[32];                                        // this thread: row `row`, head-dim columns [32 grp, +32)
    tmem_ld32(trow + 256 + grp * 32, raw);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) st_row8(sQ, row, grp * 32 + g * 8, (const float*)raw + g * 8);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) tma_store_3d(&mapdqkv, sQ + c * 16384, h * HD + 64 * c, q0, b);   // rows past the sequence end are clipped
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// (Measured dead end, round 2 -- git history has the kernel: a PERSISTENT query-side backward, one CTA per SM over the (trial, head)
// items with K / V resident for both query tiles, Q / dO prefetched behind the S / dP products, delta formed from S and dP in tensor
// memory instead of from an O tile, the dQ store issued by the producer warp.  Alone it takes 33 us per layer against 40 us for the
// kernel above, but inside the training step it is 45 us per step SLOWER: a CTA that holds an SM's whole shared memory for the
// kernel's duration keeps the weight-gradient GEMMs of the second stream off that SM, while the short-lived CTAs of the per-tile
// kernel interleave with them.  The same experiment on the key side, which has four times more work per CTA, does pay: kv3 below.)

// ---------------------------------------------------------------------------
// backward, key side: dK, dV (accumulated over the query tiles in TMEM).
// Lanes are keys here; warp w owns keys 32*(w%4).. of the tile and the query columns [64*(w/4), +64).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_bwd_kv_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap mapdo, const TcAttn a) { pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;                 // 2 x [128 keys x 64 d]  32 KB
  uint8_t* sV = sK + 32768;
  uint8_t* sQ = sV + 32768;           // 2 x [128 q x 64 d]
  uint8_t* sdO = sQ + 32768;
  uint8_t* sPt = sdO + 32768;         // 2 x [128 keys x 64 q]
  uint8_t* sdSt = sPt + 32768;
  float2* s_ld = (float2*)(sdSt + 32768);          // [128] (lse * log2e, delta) of the query tile
  uint32_t* s_bits = (uint32_t*)(s_ld + 128);      // [128][4] keep bits (query row, 32-key word of this key tile)
  uint64_t* bars = (uint64_t*)(s_bits + 512);      // kv, q, s, acc
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
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
  const int row = quarter * 32 + lane;
  const int kj = k0 + row;
  const bool kv_j = kj < L && p.key_valid[(long long)b * L + kj] != 0;
  const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
  const float sl2 = p.scale * kLog2e;
  const bool drop = p.p_attn > 0.f;
  const float ik = drop ? 1.0f / (1.0f - p.p_attn) : 1.f;
  const long long bh0 = ((long long)b * p.nh + h) * L;

  for (int it = 0; it < nq; ++it) {
    const int q0 = it * TQ;
    const int nqs = (min(L - q0, TQ) + 15) / 16;   // query steps of 16 holding real queries
    if (it > 0) { mbar_wait(&bars[3], (it - 1) & 1); tc_fence_after(); }   // previous dV/dK MMAs done: sQ/sdO/sPt/sdSt reusable
    __syncthreads();
    if (tid < TQ) {
      const int qq = q0 + tid;
      s_ld[tid] = qq < L ? make_float2(p.lse[bh0 + qq] * kLog2e, p.delta[bh0 + qq]) : make_float2(0.f, 0.f);
      if (drop) {
        const uint4 w = qq < L ? *(const uint4*)(p.drop_bits + (bh0 + qq) * 8 + (k0 >> 5)) : make_uint4(0u, 0u, 0u, 0u);
        *(uint4*)(s_bits + tid * 4) = w;
      }
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
    __syncthreads();                 // s_ld / s_bits visible
    mbar_wait(&bars[2], it & 1);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;          // query column of the tile
      if (q0 + c0 < L) {
        const uint32_t mw = key_mask_word(kj, q0 + c0, a.cf, a.cb, kv_j, L);
        uint32_t rs[32], rp[32];
        tmem_ld32(trow + c0, rs);
        tmem_ld32(trow + 128 + c0, rp);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pt[8], dst[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = g * 8 + i;
            const float2 ld = s_ld[c0 + j];
            const float s = fmaf(__uint_as_float(rs[j]), sl2, -ld.x);
            const float pr = ex2f((mw >> j) & 1u ? s : -INFINITY);
            float dm = 1.f;
            if (drop) dm = (s_bits[(c0 + j) * 4 + quarter] >> lane) & 1u ? ik : 0.f;
            pt[i] = pr * dm;
            dst[i] = pr * p.scale * (__uint_as_float(rp[j]) * dm - ld.y);
          }
          st_row8(sPt, row, c0 + g * 8, pt);
          st_row8(sdSt, row, c0 + g * 8, dst);
        }
      }
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
#pragma unroll
  for (int which = 0; which < 2; ++which) {       // 0: dV (cols 256..), 1: dK (cols 384..)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = half * 64 + cc * 32;
      uint32_t raw[32];
      tmem_ld32(trow + 256 + which * 128 + c0, raw);
      tmem_ld_wait();
      if (kj < L) {
        bf16* o = (bf16*)p.dqkv + ((long long)b * L + kj) * 3 * H + (which == 0 ? 2 * H : H) + h * HD + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) st_global8(o + g * 8, (const float*)raw + g * 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------
// backward, key side, pipelined: dK, dV of one 128-key tile accumulated over the queries in chunks of 64.
// Warp-specialised -- warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: 256 compute threads (lane = key) --
// and software-pipelined over the chunks: while the compute warps turn S^T / dP^T of chunk c into P~^T / dS^T, the
// tensor core already forms S^T / dP^T of chunk c + 1 (two accumulator buffers) and TMA fetches chunk c + 2 (three
// shared-memory stages).  P~^T and dS^T never touch shared memory: they are written back IN PLACE into tensor memory
// (bf16, two queries per column) and consumed from there as the A operand of the dV / dK products.
// Tensor memory (512 columns):  buffer b in {0,1}: S^T at [128 b, +64), dP^T at [128 b + 64, +64);
//   P~^T of the thread that owns query columns [32 h, +32) of the chunk: 16 packed columns at S^T + 32 h (its own
//   read range); dS^T likewise inside dP^T;   dV at [256, 384), dK at [384, 512).
// ---------------------------------------------------------------------------
constexpr int KV2_CH = 64, KV2_STAGES = 3, KV2_CWARPS = 16, KV2_THREADS = 64 + 32 * KV2_CWARPS;
constexpr int SMEM_BKV2 = 65536 + KV2_STAGES * 32768 + LMAX * 8 + LMAX * 16 + 128 + 1024;

__global__ void __launch_bounds__(KV2_THREADS, 1)   // 18 warps: 5 share a sub-partition -> at most 102 registers per thread
attn_tc_bwd_kv2_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                       const __grid_constant__ CUtensorMap mapdo64, const __grid_constant__ CUtensorMap mapdqkv, const TcAttn a) { pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;                          // 2 x [128 keys x 64 d]   32 KB
  uint8_t* sV = sK + 32768;
  uint8_t* sQ = sV + 32768;                    // stages x { Q chunk 2 x [64 q x 64 d] (16 KB), dO chunk (16 KB) }
  float2* s_ld = (float2*)(sQ + KV2_STAGES * 32768);       // [LMAX] (lse * log2e, delta) of every query
  uint32_t* s_bits = (uint32_t*)(s_ld + LMAX);             // [LMAX][4] keep bits (query, 32-key word of this key tile)
  uint64_t* bars = (uint64_t*)(s_bits + LMAX * 4);
  uint64_t* kv_full = bars;                    // [1]
  uint64_t* q_full = bars + 1;                 // [3]
  uint64_t* q_empty = bars + 4;                // [3]
  uint64_t* s_full = bars + 7;                 // [2]
  uint64_t* p_full = bars + 9;                 // [2]
  uint64_t* acc_done = bars + 11;              // [1]
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const AttnParams& p = a.p;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * TQ;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nc = (L + KV2_CH - 1) / KV2_CH;
  const long long bh0 = ((long long)b * p.nh + h) * L;
  const bool drop = p.p_attn > 0.f;

  if (tid == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < KV2_STAGES; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], KV2_CWARPS); }
    mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first loads fly while the CTA stages its per-query scalars and allocates tensor memory
    mbar_expect_tx(kv_full, 65536);
    for (int c = 0; c < 2; ++c) tma_load_3d(sK + c * 16384, &map128, kv_full, H + h * HD + 64 * c, k0, b);
    for (int c = 0; c < 2; ++c) tma_load_3d(sV + c * 16384, &map128, kv_full, 2 * H + h * HD + 64 * c, k0, b);
    for (int c = 0; c < nc && c < KV2_STAGES; ++c) {
      uint8_t* dq = sQ + c * 32768;
      mbar_expect_tx(&q_full[c], 32768);
      for (int d = 0; d < 2; ++d) tma_load_3d(dq + d * 8192, &map64, &q_full[c], h * HD + 64 * d, c * KV2_CH, b);
      for (int d = 0; d < 2; ++d) tma_load_3d(dq + 16384 + d * 8192, &mapdo64, &q_full[c], h * HD + 64 * d, c * KV2_CH, b);
    }
  }
  if (tid == 0) { dbg_mark(a, 0); dbg_smid(a, 31); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (tid == 32) dbg_mark(a, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) dbg_mark(a, 2);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int c = KV2_STAGES; c < nc; ++c) {        // (K, V and the first chunks were issued before the CTA-wide setup)
        const int st = c % KV2_STAGES;
        mbar_wait(&q_empty[st], ((c / KV2_STAGES) & 1) ^ 1);
        uint8_t* dq = sQ + st * 32768;
        mbar_expect_tx(&q_full[st], 32768);
        for (int d = 0; d < 2; ++d) tma_load_3d(dq + d * 8192, &map64, &q_full[st], h * HD + 64 * d, c * KV2_CH, b);
        for (int d = 0; d < 2; ++d) tma_load_3d(dq + 16384 + d * 8192, &mapdo64, &q_full[st], h * HD + 64 * d, c * KV2_CH, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KV2_CH, false, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, false, true);
      // descriptors are built once; inside the loops an MMA costs one 64-bit add per operand
      const uint64_t kd = make_sdesc(smem_u32(sK), 16, 1024), vd = make_sdesc(smem_u32(sV), 16, 1024);
      auto issue_s = [&](int c) {
        const int st = c % KV2_STAGES, bf = c & 1;
        mbar_wait(&q_full[st], (c / KV2_STAGES) & 1);
        tc_fence_after();
        const uint64_t qd = make_sdesc(smem_u32(sQ + st * 32768), 16, 1024), dod = sdesc_advance(qd, 16384);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          tc_mma_bf16(tmem + bf * 128, sdesc_advance(kd, (ks >> 2) * 16384 + (ks & 3) * 32), sdesc_advance(qd, (ks >> 2) * 8192 + (ks & 3) * 32),
                      idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          tc_mma_bf16(tmem + bf * 128 + 64, sdesc_advance(vd, (ks >> 2) * 16384 + (ks & 3) * 32),
                      sdesc_advance(dod, (ks >> 2) * 8192 + (ks & 3) * 32), idesc_s, ks > 0 ? 1u : 0u);
        tc_commit(&s_full[bf]);
      };
      mbar_wait(kv_full, 0);
      dbg_mark(a, 3);
      issue_s(0);
      dbg_mark(a, 4);
      for (int c = 0; c < nc; ++c) {
        if (c + 1 < nc) issue_s(c + 1);         // (its accumulator buffer was released by the dV/dK products of chunk c - 1, issued before)
        dbg_mark(a, 17 + c);
        const int st = c % KV2_STAGES, bf = c & 1;
        mbar_wait(&p_full[bf], (c >> 1) & 1);
        tc_fence_after();
        dbg_mark(a, 21 + c);
        const uint64_t qm = make_sdesc(smem_u32(sQ + st * 32768), 8192, 1024), dom = sdesc_advance(qm, 16384);   // MN-major views
#pragma unroll
        for (int ks = 0; ks < KV2_CH / 16; ++ks)
          tc_mma_bf16_ts(tmem + 256, tmem + bf * 128 + 16 * ks, sdesc_advance(dom, ks * 2048), idesc_acc, (c > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KV2_CH / 16; ++ks)
          tc_mma_bf16_ts(tmem + 384, tmem + bf * 128 + 64 + 16 * ks, sdesc_advance(qm, ks * 2048), idesc_acc, (c > 0 || ks > 0) ? 1u : 0u);
        tc_commit(&q_empty[st]);                // the chunk's shared-memory stage is free once these retire
        dbg_mark(a, 12 + c);
      }
      tc_commit(acc_done);
    }
  } else {
    // ===================== compute warps: lane = key, 16 query columns of the chunk per thread =====================
    const int cw = warp - 2;
    const int quarter = warp & 3, grp = cw >> 2;             // TMEM lane quarter; which 16 of the chunk's 64 queries
    const int row = quarter * 32 + lane;
    const int kj = k0 + row;
    const bool kv_j = kj < L && p.key_valid[(long long)b * L + kj] != 0;
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    const float ik = drop ? 1.0f / (1.0f - p.p_attn) : 1.f;
    // Idle until the first S^T arrives: stage the per-query scalars of the WHOLE sequence, and turn the forward's
    // keep bits (one word per query x 32 keys) into key-major words (this key x 32 queries) with warp ballots.
    for (int q = tid - 64; q < LMAX; q += 32 * KV2_CWARPS)
      s_ld[q] = q < L ? make_float2(p.lse[bh0 + q] * kLog2e, p.delta[bh0 + q]) : make_float2(0.f, 0.f);
    for (int q = tid - 64; q < LMAX; q += 32 * KV2_CWARPS) {   // keep bits (query q, the 128 keys of this tile): 4 words per query
      uint4 w = make_uint4(~0u, ~0u, ~0u, ~0u);
      if (drop && q < L) w = *(const uint4*)(p.drop_bits + (bh0 + q) * 8 + (k0 >> 5));
      *(uint4*)(s_bits + q * 4) = w;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * KV2_CWARPS) : "memory");
    for (int c = 0; c < nc; ++c) {
      const int bf = c & 1;
      const int qb = c * KV2_CH + grp * 16;                  // first query of this thread's columns
      mbar_wait(&s_full[bf], (c >> 1) & 1);
      tc_fence_after();
      if (tid == 64) dbg_mark(a, 5 + c);
      uint32_t pk[8], dk[8];
      if (qb < L) {                                          // warp-uniform
        const uint32_t mw = key_mask_word(kj, qb & ~31, a.cf, a.cb, kv_j, L) >> (qb & 31);
        uint32_t rs[16], rp[16];
        tmem_ld16(trow + bf * 128 + grp * 16, rs);
        tmem_ld16(trow + bf * 128 + 64 + grp * 16, rp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float pt[2], dst[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float2 ld = s_ld[qb + j + u];
            const float s = fmaf(__uint_as_float(rs[j + u]), sl2, -ld.x);
            const float pr = ex2f((mw >> (j + u)) & 1u ? s : -INFINITY);
            const float dm = (s_bits[(qb + j + u) * 4 + quarter] >> lane) & 1u ? ik : 0.f;
            pt[u] = pr * dm;
            dst[u] = pr * p.scale * (__uint_as_float(rp[j + u]) * dm - ld.y);
          }
          pk[j >> 1] = pack_bf16x2(pt[0], pt[1]);
          dk[j >> 1] = pack_bf16x2(dst[0], dst[1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { pk[j] = 0u; dk[j] = 0u; }
      }
      tmem_st8(trow + bf * 128 + grp * 16, pk);               // in place: 8 packed columns inside this thread's own 16
      tmem_st8(trow + bf * 128 + 64 + grp * 16, dk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[bf]);
    }
    // ---- epilogue: dV (cols 256..), dK (cols 384..) -> global
    if (tid == 64) dbg_mark(a, 9);
    mbar_wait(acc_done, 0);
    tc_fence_after();
    if (tid == 64) dbg_mark(a, 10);
    // K and V are dead: their buffers stage dK / dV as SWIZZLE_128B tiles for two bulk tensor stores (coalesced, asynchronous,
    // rows past the sequence end clipped by the tensor map)
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const int c0 = grp * 32;
      uint32_t r0[32];
      tmem_ld32(trow + 256 + which * 128 + c0, r0);
      tmem_ld_wait();
      uint8_t* tile = which == 0 ? sV : sK;                  // dV over V's buffer, dK over K's
#pragma unroll
      for (int g = 0; g < 4; ++g) st_row8(tile, row, c0 + g * 8, (const float*)r0 + g * 8);
    }
    fence_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(32 * KV2_CWARPS) : "memory");
    if (tid == 64) {
      for (int c = 0; c < 2; ++c) {
        tma_store_3d(&mapdqkv, sV + c * 16384, 2 * H + h * HD + 64 * c, k0, b);
        tma_store_3d(&mapdqkv, sK + c * 16384, H + h * HD + 64 * c, k0, b);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (tid == 64) dbg_mark(a, 11);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  if (tid == 32) dbg_mark(a, 16);
}

// ---------------------------------------------------------------------------
// backward, key side, PERSISTENT: the kv2 pipeline run over a list of (trial, head, key tile) items by one CTA per SM.
// A CTA of kv2 lives ~10.5 us of which ~2.4 us pass before its first MMA (barriers, tensor-memory allocation, first loads) and
// ~2 us after its last one (accumulator drain, stores, exit), and the next CTA starts ~1.4 us later: with 3.5 CTAs per SM that
// is a third of the kernel.  Here tensor memory and barriers are set up once, the producer runs ahead across item boundaries
// (the next item's first query chunks and its K / V tiles are in flight while the current item's last chunks are processed), the
// MMA warp issues the next item's first S^T / dP^T as soon as the current item's products are queued, and the accumulator drain
// of an item overlaps the next item's first products.
//   shared memory: ONE K / V buffer (64 KB; K and V are operands of S^T / dP^T only, so it is reloaded as soon as the item's last
//   S^T / dP^T have retired) + a ring of FOUR 32 KB query stages (Q chunk | dO chunk).  The two stages of an item's last two
//   chunks double as the staging tiles of its dV / dK bulk stores; they return to the ring when the stores have read them.
//   tensor memory: as kv2.
// ---------------------------------------------------------------------------
constexpr int KV3_QST = 4;
constexpr int SMEM_BKV3 = 65536 + KV3_QST * 32768 + 2 * (LMAX * 8 + LMAX * 16) + 256 + 1024;

__global__ void __launch_bounds__(KV2_THREADS, 1)
attn_tc_bwd_kv3_kernel(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                       const __grid_constant__ CUtensorMap mapdo64, const __grid_constant__ CUtensorMap mapdqkv, const TcAttn a, const int n_items) {
  pdl_grid_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;                          // 2 x [128 keys x 64 d]   32 KB
  uint8_t* sV = sK + 32768;
  uint8_t* sQ = sV + 32768;                    // ring: stages x { Q chunk 2 x [64 q x 64 d] (16 KB), dO chunk (16 KB) }
  float2* s_ld2 = (float2*)(sQ + KV3_QST * 32768);         // [2][LMAX] (lse * log2e, delta) of every query, double-buffered over items
  uint32_t* s_bits2 = (uint32_t*)(s_ld2 + 2 * LMAX);       // [2][LMAX][4] keep bits (query, 32-key word of this key tile)
  uint64_t* bars = (uint64_t*)(s_bits2 + 2 * LMAX * 4);
  uint64_t* kv_full = bars;                    // [1]  K / V of the item landed
  uint64_t* kv_free = bars + 1;                // [1]  the item's last S^T / dP^T retired: the buffer may take the next item's tiles
  uint64_t* q_full = bars + 2;                 // [4]
  uint64_t* q_empty = bars + 6;                // [4]
  uint64_t* s_full = bars + 10;                // [2]
  uint64_t* p_full = bars + 12;                // [2]
  uint64_t* acc_done = bars + 14;              // [1]  the item's dV / dK products retired
  uint64_t* acc_free = bars + 15;              // [1]  the compute warps have read the accumulators
  uint64_t* stage_ready = bars + 16;           // [1]  the item's dV / dK tiles are staged in shared memory (the producer stores them)
  uint32_t* tmem_slot = (uint32_t*)(bars + 17);

  const AttnParams& p = a.p;
  const int L = p.L, H = p.H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nc = (L + KV2_CH - 1) / KV2_CH;            // query chunks per item (2..4)
  const int nk = (L + TQ - 1) / TQ;                    // key tiles per (trial, head)
  const bool drop = p.p_attn > 0.f;
  const int my_items = blockIdx.x < n_items ? (n_items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto item_of = [&](int it, int& b, int& h, int& k0) {
    const int item = blockIdx.x + it * gridDim.x;
    const int kt = item % nk, bh = item / nk;
    k0 = kt * TQ; h = bh % p.nh; b = bh / p.nh;
  };

  if (tid == 0) {
    mbar_init(kv_full, 1); mbar_init(kv_free, 1);
    for (int i = 0; i < KV3_QST; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], KV2_CWARPS); }
    mbar_init(acc_done, 1); mbar_init(acc_free, KV2_CWARPS); mbar_init(stage_ready, KV2_CWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: one flat stream of chunks over all items; it also issues the items' bulk stores ====
    if (lane == 0) {
      const int total = my_items * nc;
      uint32_t qe_phase[KV3_QST] = {0, 0, 0, 0};   // parity of the next q_empty completion of every stage
      int stored = 0;                              // items whose dV / dK tiles have been stored
      auto store_item = [&](int it) {              // the item's two staging stages -> global; they rejoin the ring when the stores have read them
        int b, h, k0; item_of(it, b, h, k0);
        const int g_last = it * nc + nc - 1;
        uint8_t* tile_v = sQ + ((g_last - 1) % KV3_QST) * 32768;
        uint8_t* tile_k = sQ + (g_last % KV3_QST) * 32768;
        mbar_wait(stage_ready, it & 1);
        if (it < 4) dbg_mark64(a, it * 12 + 7);                // staging tiles seen by the producer
        for (int d = 0; d < 2; ++d) {
          tma_store_3d(&mapdqkv, tile_v + d * 16384, 2 * H + h * HD + 64 * d, k0, b);
          tma_store_3d(&mapdqkv, tile_k + d * 16384, H + h * HD + 64 * d, k0, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (it < 4) dbg_mark64(a, it * 12 + 8);                // stores have read shared memory
      };
      for (int g = 0; g < total; ++g) {
        const int it = g / nc, c = g - it * nc;
        int b, h, k0; item_of(it, b, h, k0);
        const int kvpos = it == 0 ? 0 : (nc > 2 ? 2 : nc - 1);     // where in the item's chunk sequence its K / V tiles are fetched
        if (c == kvpos) {
          if (it > 0) mbar_wait(kv_free, (it - 1) & 1);
          if (it < 4) dbg_mark64(a, it * 12 + 11);            // K / V of the item requested
          mbar_expect_tx(kv_full, 65536);
          for (int d = 0; d < 2; ++d) tma_load_3d(sK + d * 16384, &map128, kv_full, H + h * HD + 64 * d, k0, b);
          for (int d = 0; d < 2; ++d) tma_load_3d(sV + d * 16384, &map128, kv_full, 2 * H + h * HD + 64 * d, k0, b);
        }
        // chunks nc-2 and nc-1 reuse the stages that staged the PREVIOUS item's dV / dK: store those first (after this item's K / V
        // request above: the store waits for the previous item's accumulator drain)
        if (it > 0 && c == nc - 2 && stored < it) { store_item(it - 1); stored = it; }
        const int st = g % KV3_QST;
        // (a stage that staged a store was released by this thread itself, above: no barrier; the others come back through q_empty,
        //  which only chunks c < nc - 2 of the stage's previous user signal)
        const int gp = g - KV3_QST;                                   // the stage's previous user
        if (gp >= 0 && (gp % nc) < nc - 2) mbar_wait(&q_empty[st], qe_phase[st]), qe_phase[st] ^= 1;
        uint8_t* dq = sQ + st * 32768;
        mbar_expect_tx(&q_full[st], 32768);
        for (int d = 0; d < 2; ++d) tma_load_3d(dq + d * 8192, &map64, &q_full[st], h * HD + 64 * d, c * KV2_CH, b);
        for (int d = 0; d < 2; ++d) tma_load_3d(dq + 16384 + d * 8192, &mapdo64, &q_full[st], h * HD + 64 * d, c * KV2_CH, b);
      }
      for (int it = stored; it < my_items; ++it) store_item(it);     // the last item(s)
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && my_items > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KV2_CH, false, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, false, true);
      const uint64_t kd = make_sdesc(smem_u32(sK), 16, 1024), vd = make_sdesc(smem_u32(sV), 16, 1024);
      auto issue_s = [&](int g) {               // S^T / dP^T of global chunk g (its item's K / V must be resident)
        const int st = g % KV3_QST, bf = g & 1;
        mbar_wait(&q_full[st], (g / KV3_QST) & 1);
        tc_fence_after();
        const uint64_t qd = make_sdesc(smem_u32(sQ + st * 32768), 16, 1024), dod = sdesc_advance(qd, 16384);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          tc_mma_bf16(tmem + bf * 128, sdesc_advance(kd, (ks >> 2) * 16384 + (ks & 3) * 32), sdesc_advance(qd, (ks >> 2) * 8192 + (ks & 3) * 32),
                      idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          tc_mma_bf16(tmem + bf * 128 + 64, sdesc_advance(vd, (ks >> 2) * 16384 + (ks & 3) * 32),
                      sdesc_advance(dod, (ks >> 2) * 8192 + (ks & 3) * 32), idesc_s, ks > 0 ? 1u : 0u);
        tc_commit(&s_full[bf]);
      };
      auto acc = [&](int g, int it, int c) {   // dV += P~^T dO, dK += dS^T Q of global chunk g
        const int st = g % KV3_QST, bf = g & 1;
        mbar_wait(&p_full[bf], (g >> 1) & 1);
        if (c == 0 && it > 0) mbar_wait(acc_free, (it - 1) & 1);       // the previous item's accumulators have been read
        tc_fence_after();
        const uint64_t qm = make_sdesc(smem_u32(sQ + st * 32768), 8192, 1024), dom = sdesc_advance(qm, 16384);   // MN-major views
#pragma unroll
        for (int ks = 0; ks < KV2_CH / 16; ++ks)
          tc_mma_bf16_ts(tmem + 256, tmem + bf * 128 + 16 * ks, sdesc_advance(dom, ks * 2048), idesc_acc, (c > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KV2_CH / 16; ++ks)
          tc_mma_bf16_ts(tmem + 384, tmem + bf * 128 + 64 + 16 * ks, sdesc_advance(qm, ks * 2048), idesc_acc, (c > 0 || ks > 0) ? 1u : 0u);
        // the stage goes back to the ring when these retire -- except the item's last two, which stage its dV / dK stores first
        if (c < nc - 2) tc_commit(&q_empty[st]);
      };
      mbar_wait(kv_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int it = 0; it < my_items; ++it) {
        const int g0 = it * nc;
        for (int c = 0; c < nc; ++c) {
          if (c + 1 < nc) {
            issue_s(g0 + c + 1);
            if (c + 1 == nc - 1) tc_commit(kv_free);            // the item's last S^T / dP^T are queued: K / V may be replaced when they retire
          }
          acc(g0 + c, it, c);
        }
        tc_commit(acc_done);
        if (it < 4) dbg_mark64(a, it * 12 + 10);                 // the item's last products are queued
        if (it + 1 < my_items) {                                 // the next item's first products, while this item's accumulators drain
          mbar_wait(kv_full, (it + 1) & 1);
          tc_fence_after();
          issue_s(g0 + nc);
        }
      }
    }
  } else {
    // ===================== compute warps: lane = key, 16 query columns of the chunk per thread =====================
    const int cw = warp - 2;
    const int quarter = warp & 3, grp = cw >> 2;             // TMEM lane quarter; which 16 of the chunk's 64 queries
    const int row = quarter * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    const float ik = drop ? 1.0f / (1.0f - p.p_attn) : 1.f;
    // per-query scalars and keep bits of an item's (trial, head, key tile), into the buffer of its parity
    auto load_scalars = [&](int it) {
      int b, h, k0; item_of(it, b, h, k0);
      const long long bh0 = ((long long)b * p.nh + h) * L;
      float2* s_ld = s_ld2 + (it & 1) * LMAX;
      uint32_t* s_bits = s_bits2 + (it & 1) * LMAX * 4;
      for (int q = tid - 64; q < LMAX; q += 32 * KV2_CWARPS)
        s_ld[q] = q < L ? make_float2(p.lse[bh0 + q] * kLog2e, p.delta[bh0 + q]) : make_float2(0.f, 0.f);
      for (int q = tid - 64; q < LMAX; q += 32 * KV2_CWARPS) {
        uint4 w = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (drop && q < L) w = *(const uint4*)(p.drop_bits + (bh0 + q) * 8 + (k0 >> 5));
        *(uint4*)(s_bits + q * 4) = w;
      }
    };
    if (my_items > 0) load_scalars(0);
    for (int it = 0; it < my_items; ++it) {
      int b, h, k0; item_of(it, b, h, k0);
      const int kj = k0 + row;
      const bool kv_j = kj < L && p.key_valid[(long long)b * L + kj] != 0;
      const float2* s_ld = s_ld2 + (it & 1) * LMAX;
      const uint32_t* s_bits = s_bits2 + (it & 1) * LMAX * 4;
      if (tid == 64 && it < 4) dbg_mark64(a, it * 12 + 0);      // item start (compute warps)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * KV2_CWARPS) : "memory");      // this item's scalars are in place (all warps are past the previous item's chunks)
      for (int c = 0; c < nc; ++c) {
        const int g = it * nc + c, bf = g & 1;
        const int qb = c * KV2_CH + grp * 16;                  // first query of this thread's columns
        mbar_wait(&s_full[bf], (g >> 1) & 1);
        tc_fence_after();
        if (tid == 64 && it < 4 && c < 4) dbg_mark64(a, it * 12 + 2 + c);     // S^T / dP^T of chunk c seen
        uint32_t pk[8], dk[8];
        if (qb < L) {                                          // warp-uniform
          const uint32_t mw = key_mask_word(kj, qb & ~31, a.cf, a.cb, kv_j, L) >> (qb & 31);
          uint32_t rs[16], rp[16];
          tmem_ld16(trow + bf * 128 + grp * 16, rs);
          tmem_ld16(trow + bf * 128 + 64 + grp * 16, rp);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float pt[2], dst[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float2 ld = s_ld[qb + j + u];
              const float sv = fmaf(__uint_as_float(rs[j + u]), sl2, -ld.x);
              const float pr = ex2f((mw >> (j + u)) & 1u ? sv : -INFINITY);
              const float dm = (s_bits[(qb + j + u) * 4 + quarter] >> lane) & 1u ? ik : 0.f;
              pt[u] = pr * dm;
              dst[u] = pr * p.scale * (__uint_as_float(rp[j + u]) * dm - ld.y);
            }
            pk[j >> 1] = pack_bf16x2(pt[0], pt[1]);
            dk[j >> 1] = pack_bf16x2(dst[0], dst[1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { pk[j] = 0u; dk[j] = 0u; }
        }
        tmem_st8(trow + bf * 128 + grp * 16, pk);               // in place: 8 packed columns inside this thread's own 16
        tmem_st8(trow + bf * 128 + 64 + grp * 16, dk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[bf]);
      }
      if (tid == 64 && it < 4) dbg_mark64(a, it * 12 + 1);      // last P of the item written
      // the next item's scalars, while this item's last products retire (their buffer's readers are all behind the barrier above)
      if (it + 1 < my_items) load_scalars(it + 1);
      // ---- this item's dV (cols 256..), dK (cols 384..): tensor memory -> registers -> the two free ring stages; the producer stores them
      mbar_wait(acc_done, it & 1);
      tc_fence_after();
      if (tid == 64 && it < 4) dbg_mark64(a, it * 12 + 6);      // accumulators complete
      const int g_last = it * nc + nc - 1;
      uint8_t* tile_v = sQ + ((g_last - 1) % KV3_QST) * 32768;
      uint8_t* tile_k = sQ + (g_last % KV3_QST) * 32768;
      uint32_t r0[32], r1[32];
      tmem_ld32(trow + 256 + grp * 32, r0);
      tmem_ld32(trow + 384 + grp * 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);                    // the next item's products may overwrite the accumulators
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) st_row8(tile_v, row, grp * 32 + g4 * 8, (const float*)r0 + g4 * 8);
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) st_row8(tile_k, row, grp * 32 + g4 * 8, (const float*)r1 + g4 * 8);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(stage_ready);                 // (16 warps: the producer issues the bulk stores)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); if (lane == 0) dbg_mark64(a, 63); }
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

unsigned long long* g_attn_dbg = nullptr;
constexpr int SMEM_FWD = 32768 + 65536 + 2048 + 32 + 64 + 1024;
constexpr int SMEM_BQ = 32768 * 2 + 65536 * 2 + 32768 + 32 + 64 + 1024;
constexpr int SMEM_BKV = 32768 * 6 + 1024 + 2048 + 64 + 1024;

}  // namespace

void k_attention_tc_set_timeline(unsigned long long* buf) { g_attn_dbg = buf; }

bool k_attention_tc_supported(const AttnParams& p) { return p.hd == HD && p.L <= LMAX && p.L >= 1 && p.H % 8 == 0; }

int k_attention_tc_dropbits(const AttnBitsJob& job, const AttnParams& p, cudaStream_t stream) {
  NDT1_REQUIRE(job.n >= 1 && job.n <= 32, "attention_tc: %d layers in one keep-bit launch (max 32)", job.n);
  NDT1_REQUIRE(p.H % 32 == 0, "attention_tc: hidden size %d must be a multiple of 32", p.H);
  const long long rows_p = (long long)p.B * p.nh * p.L, rows_o = (long long)p.B * p.L;
  bool any_m = false;
  for (int l = 0; l < job.n; ++l) any_m |= job.bits_m[l] != nullptr;
  const long long words = rows_p * 8 + rows_o * (p.H / 32) * (any_m ? 2 : 1);
  if (words == 0) return 0;
  const int bx = (int)((words + 255) / 256 < 148 * 4 ? (words + 255) / 256 : 148 * 4);
  ndt1_launch(attn_dropbits_kernel, dim3(bx, job.n), 256, 0, stream, job, p.seed, rows_p, p.L, rows_o, p.H, drop_threshold(p.p_attn),
              drop_threshold(p.p_out));
  NDT1_CHECK_LAUNCH();
  return 0;
}

int k_attention_tc_fwd(const AttnParams& p, cudaStream_t stream) {
  NDT1_REQUIRE(k_attention_tc_supported(p), "attention_tc: unsupported shape (head size %d, %d tokens)", p.hd, p.L);
  NDT1_TRY(gemm_tc_init());
  NDT1_REQUIRE(p.p_attn <= 0.f || p.drop_bits, "attention_tc: probability dropout needs the drop_bits buffer");
  NDT1_REQUIRE(p.p_out <= 0.f || p.out_drop == p.out || p.drop_bits_o, "attention_tc: output dropout needs the drop_bits_o buffer");
  if (!p.bits_ready && (p.p_attn > 0.f || (p.p_out > 0.f && p.out_drop != p.out))) {      // stand-alone use: draw this layer's keep bits first
    AttnBitsJob job; memset(&job, 0, sizeof(job));
    job.n = 1; job.bits_p[0] = p.p_attn > 0.f ? p.drop_bits : nullptr; job.bits_o[0] = (p.p_out > 0.f && p.out_drop != p.out) ? p.drop_bits_o : nullptr;
    job.stream_p[0] = p.stream_attn; job.stream_o[0] = p.stream_out;
    NDT1_TRY(k_attention_tc_dropbits(job, p, stream));
  }
  CUtensorMap m128, m256;
  NDT1_TRY(make_maps(p, &m128, &m256, nullptr));
  static Ndt1PerDeviceFlag attr;
  if (!attr.here()) { NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD)); attr.here() = true; }
  TcAttn a; a.p = p; a.cf = p.ctx_fwd; a.cb = p.ctx_bwd; a.dbg = g_attn_dbg;
  dim3 grid(ndt1_cdiv(p.L, TQ), p.nh, p.B);
  GemmOperand o; o.ptr = p.out; o.batch_stride = (long long)p.L * p.H; o.nbatch = p.B; o.rows = p.L; o.cols = p.H; o.ld = p.H;
  GemmOperand od = o; od.ptr = p.out_drop;
  CUtensorMap mo, mod;
  NDT1_TRY(tc_make_map(o, 64, 128, &mo));
  NDT1_TRY(tc_make_map(od, 64, 128, &mod));
  const double mm = 2.0 * p.B * p.nh * (double)p.L * p.L * HD;       // FLOPs of one L x L x d contraction over all heads
  if (g_ndt1_prof_on) ndt1_prof_note(2 * mm, 0.0);                  // forward: Q K^T and P V
  ndt1_launch(attn_tc_fwd_kernel, grid, NTHREADS, SMEM_FWD, stream, m128, m256, mo, mod, a);
  NDT1_CHECK_LAUNCH();
  return 0;
}

int k_attention_tc_bwd(const AttnParams& p, cudaStream_t stream) {
  NDT1_REQUIRE(k_attention_tc_supported(p), "attention_tc: unsupported shape (head size %d, %d tokens)", p.hd, p.L);
  NDT1_TRY(gemm_tc_init());
  CUtensorMap m128, m256, mdo;
  NDT1_TRY(make_maps(p, &m128, &m256, &mdo));
  static Ndt1PerDeviceFlag attr;
  if (!attr.here()) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BQ));
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BKV));
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_kv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BKV2));
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_bwd_kv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BKV3));
    attr.here() = true;
  }
  TcAttn a; a.p = p; a.cf = p.ctx_fwd; a.cb = p.ctx_bwd; a.dbg = g_attn_dbg;
  dim3 grid(ndt1_cdiv(p.L, TQ), p.nh, p.B);
  GemmOperand g; g.ptr = p.dqkv; g.batch_stride = (long long)p.L * 3 * p.H; g.nbatch = p.B; g.rows = p.L; g.cols = 3 * p.H; g.ld = 3 * p.H;
  GemmOperand q; q.ptr = p.qkv; q.batch_stride = (long long)p.L * 3 * p.H; q.nbatch = p.B; q.rows = p.L; q.cols = 3 * p.H; q.ld = 3 * p.H;
  GemmOperand d; d.ptr = p.dout; d.batch_stride = (long long)p.L * p.H; d.nbatch = p.B; d.rows = p.L; d.cols = p.H; d.ld = p.H;
  GemmOperand o; o.ptr = p.out; o.batch_stride = (long long)p.L * p.H; o.nbatch = p.B; o.rows = p.L; o.cols = p.H; o.ld = p.H;
  CUtensorMap mdq, mo;
  NDT1_TRY(tc_make_map(g, 64, 128, &mdq));
  NDT1_TRY(tc_make_map(o, 64, 128, &mo));
  // query side first: it also forms delta = rowsum(dO * O), which the key side reads
  const double mm = 2.0 * p.B * p.nh * (double)p.L * p.L * HD;
  if (g_ndt1_prof_on) ndt1_prof_note(2 * mm, 0.0);                  // algorithmic: dP = dO V^T and dQ = dS K (the recomputed S is not counted)
  ndt1_launch(attn_tc_bwd_q_kernel, grid, BQ_THREADS, SMEM_BQ, stream, m128, m256, mdo, mo, mdq, a);
  NDT1_CHECK_LAUNCH();
  static const bool old_kv = getenv("NDT1_ATTN_BWD_KV1") && getenv("NDT1_ATTN_BWD_KV1")[0] == '1';
  if (old_kv) {
    ndt1_launch(attn_tc_bwd_kv_kernel, grid, NTHREADS, SMEM_BKV, stream, m128, mdo, a);
  } else {
    CUtensorMap m64, mdo64;
    NDT1_TRY(tc_make_map(q, 64, KV2_CH, &m64));
    NDT1_TRY(tc_make_map(d, 64, KV2_CH, &mdo64));
    if (g_ndt1_prof_on) ndt1_prof_note(2 * mm, 0.0);                // algorithmic: dV = P~^T dO and dK = dS^T Q
    static const bool kv2_only = getenv("NDT1_ATTN_BWD_KV2") && getenv("NDT1_ATTN_BWD_KV2")[0] == '1';
    if (!kv2_only && p.L > KV2_CH) {          // persistent pipeline (needs at least two query chunks per item for its store staging)
      const int sms = ndt1_num_sms();
      const int n_items = p.B * p.nh * ndt1_cdiv(p.L, TQ);
      ndt1_launch(attn_tc_bwd_kv3_kernel, dim3(n_items < sms ? n_items : sms), KV2_THREADS, SMEM_BKV3, stream, m128, m64, mdo64, mdq, a, n_items);
      NDT1_CHECK_LAUNCH();
      return 0;
    }
    ndt1_launch(attn_tc_bwd_kv2_kernel, grid, KV2_THREADS, SMEM_BKV2, stream, m128, m64, mdo64, mdq, a);
  }
  NDT1_CHECK_LAUNCH();
  return 0;
}
