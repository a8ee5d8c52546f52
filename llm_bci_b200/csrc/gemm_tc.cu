// Fast-mode GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) ->
// shared-memory ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) ->
// tcgen05.ld epilogue.  Persistent, warp specialised:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..9  epilogue: TMEM -> registers -> swizzled smem transpose ->
//               fused epilogue with row-contiguous (128 B per 8 lanes) global
//               reads / writes; two warps per TMEM lane quarter
// Two TMEM accumulator stages let the epilogue of tile i overlap the main
// loop of tile i+1.  Tile = 128 x BN (BN in {64,128,256}), BLOCK_K = 64.
//
// Operand layouts follow gemm_common.cuh:
//   GEMM_NT  A K-major,  B K-major      (forward linears, stack projection)
//   GEMM_NN  A K-major,  B MN-major     (data gradients, reads the forward weight)
//   GEMM_TN  A MN-major, B MN-major     (weight gradients, split over the reduction)
// so no transposed copy of a weight or an activation is ever materialised.
#include "tc_common.cuh"
#include "gemm_common.cuh"
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <algorithm>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int kEpiWarps = 8;           // two warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemPipe = 192 * 1024;  // operand ring
constexpr int kStageTile = 32 * 32 * 4;   // per-warp staging tile of the epilogue: 32 rows x 32 fp32
constexpr int kMaxRagTiles = 320;         // ragged tile table carried in the kernel parameters

struct TcParams {
  int mode;
  int M, N, nb_out;
  int nchunk, kb_per_chunk;            // k-blocks (of BK) per chunk
  int a_row_shift, a_col_shift, b_row_shift, b_col_shift, b_chunk_n;
  int split_k;
  int m_tiles, n_tiles, total_tiles;
  int chunk_k_valid;                   // reduction length per chunk in elements (for FLOP accounting)
  const long long* b_sel; int b_sel_n; // per-day B operand batch of an output trial (GEMM_NT), null = batch 0
  unsigned long long* dbg;             // optional per-CTA phase timeline (tools/gemm_timeline.py): 16 slots per CTA, globaltimer ns
  GemmEpilogue epi;
  // Ragged column tiles (CTA pairs, NT / NN, one output batch): when the regular 256-column tiling leaves the last wave partly
  // empty (7776 x 1024: 124 tiles on 74 pairs), the row strips are cut into tiles of 128 / 192 / 256 columns instead -- as many
  // tiles as a whole number of waves holds -- and dealt to the pairs by width, so that every pair carries the same number of
  // columns (+- 64) and the exposed last epilogue is a narrow one.  rag[t] = strip | first 64-column unit << 6 | (units - 1) << 12;
  // tile t goes to pair t % pairs as always.  tcgen05.mma takes N in steps of 16, the B box is loaded whole (unused rows ignored).
  int ragged;
  unsigned short rag[kMaxRagTiles];
};

using namespace tc;

__device__ __forceinline__ void dbg_stamp(unsigned long long* dbg, int slot) {
  if (dbg) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  __nv_bfloat162 x = __floats2bfloat162_rn(a, b), y = __floats2bfloat162_rn(c, d);
  uint2 o; o.x = *(uint32_t*)&x; o.y = *(uint32_t*)&y;
  return o;
}
__device__ __forceinline__ float4 unpack4_bf16(uint2 r) {
  const __nv_bfloat162 x = *(const __nv_bfloat162*)&r.x, y = *(const __nv_bfloat162*)&r.y;
  return make_float4(__low2float(x), __high2float(x), __low2float(y), __high2float(y));
}


// GELU(erf) and its derivative for the tensor-core (bf16) mode: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7,
// far below bf16 resolution) on the MUFU pipe (one ex2, one rcp) instead of erff/expf -- the epilogue of the MLP
// GEMMs is ALU-bound otherwise.  The strict fp32 mode (gemm_simt.cu) keeps erff.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void gelu_parts(float v, float& cdf, float& e) {
  const float z = fabsf(v) * 0.70710678118654752f;
  e = ex2_approx(z * z * -1.4426950408889634f);         // exp(-v^2 / 2)
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float half_erf = fmaf(-0.5f * q * t, e, 0.5f);  // erf(|v|/sqrt2) / 2
  cdf = 0.5f + copysignf(half_erf, v);
}
__device__ __forceinline__ float gelu_fast(float v) { float cdf, e; gelu_parts(v, cdf, e); return v * cdf; }
__device__ __forceinline__ float dgelu_fast(float x) { float cdf, e; gelu_parts(x, cdf, e); return fmaf(x * 0.3989422804014327f, e, cdf); }

// ---------------------------------------------------------------------------
// Epilogue.  tcgen05.ld hands every lane one ROW of the accumulator; storing that
// directly makes each warp instruction touch 32 different rows (32 L1 wavefronts
// for 512 bytes).  Instead each warp transposes its 32x32 fp32 chunk through a
// private, XOR-swizzled shared-memory tile so that 8 consecutive lanes cover 128
// contiguous bytes of one output row: the residual / saved-activation / position
// table reads and the output stores are all full-line accesses.  The side input
// of chunk i+1 is loaded into registers before chunk i is processed (and the first
// chunk of the next tile before the accumulator barrier is waited on), so its
// latency hides behind the tensor-core main loop.
//
// Lane l owns columns 4*(l%8)..+3 of rows (l/8)+4*i, i = 0..7, of the chunk.
// ---------------------------------------------------------------------------
enum { SIDE_NONE = 0, SIDE_RESID = 1, SIDE_DACT = 2, SIDE_GATHER = 3 };

// Epilogue classes: the kernel is instantiated per class with everything else compiled OUT, so that the epilogue of a
// launch is a few hundred instructions that stay in the instruction cache (the all-features epilogue does not, and
// instruction-fetch stalls then rival the arithmetic).  EF_* = features a class may use.
enum {
  EF_OUT2 = 1, EF_ACT = 2, EF_GATHER = 4, EF_DROP = 8, EF_DACT = 16, EF_RESID = 32, EF_COLSUM = 64, EF_ACCUM = 128, EF_SCALAR = 256, EF_DROPBITS = 512,
  EPI_NONE = 0,                                          // bias only: QKV, plain data gradients (no side input, no side registers)
  EPI_PLAIN = EF_RESID,                                  // bias + fp32 residual: attention out-proj
  EPI_ACT = EF_OUT2 | EF_ACT,                            // bias + activation (+ pre-activation copy): MLP up-proj, channel embedding
  EPI_DROP = EF_GATHER | EF_DROP | EF_RESID,             // bias + position rows + dropout (+ residual): stack projection
  EPI_BITSRES = EF_DROPBITS | EF_RESID,                  // bias + dropout from keep bits drawn ahead (no Philox code: no spills) + residual: MLP down-proj
  EPI_BITSBWD = EF_DROPBITS | EF_DACT | EF_COLSUM,       // backward with the keep bits: attention out-proj data gradient
  EPI_DACT = EF_DROP | EF_DACT | EF_COLSUM,              // backward: dropout mask, activation derivative, bias-gradient column sums
  EPI_DMUL = EF_DACT | EF_COLSUM,                        // backward without a dropout mask (no Philox: fewer registers): MLP down-proj data gradient
  EPI_ACCUM = EF_ACCUM,                                  // weight gradients: fp32 red.add
  EPI_ALL = 511                                          // everything, incl. the element-wise path for ragged shapes
};
__host__ __device__ constexpr bool ef_has(int cls, int f) { return (cls & f) != 0; }

struct EpiCtx {
  int side_kind;
  bool vec_ok;
  unsigned long long seed;      // the step's Philox key (read once per warp: it may live in device memory)
};

struct ChunkAt {          // where a (tile, chunk) lands in the output
  int bt, r0, n;          // trial, first row of this warp's 32-row strip, first column of this lane
  long long rowbase;      // element offset of (trial, this lane's first row, column 0): computed once per tile
  long long bias_off;     // per-day bias row (0 without routing)
  uint32_t rowmask;       // bit i: row r0 + lane/8 + 4 i of this lane is inside the output (per tile, not per chunk)
  int nch, tcol;          // 32-column chunks of this warp in the tile; this warp's first accumulator column (ragged tiles vary)
  bool live;
};

template <int EPI>
__device__ __forceinline__ void side_load(const GemmEpilogue& e, const EpiCtx& cx, const TcParams& p, const ChunkAt& at, int lane,
                                          float4* side) {
  if (!ef_has(EPI, EF_RESID | EF_DACT | EF_GATHER)) return;
  if (cx.side_kind == SIDE_NONE || !cx.vec_ok || !at.live || at.n >= p.N) return;
  const long long idx0 = at.rowbase + at.n;        // one 64-bit base per chunk; rows are 32-bit offsets from it
  const uint32_t ld4 = 4u * (uint32_t)e.ldc;
  if (ef_has(EPI, EF_RESID) && cx.side_kind == SIDE_RESID) {
    const float* b = e.resid + idx0;
#pragma unroll
    for (int i = 0; i < 8; ++i) if ((at.rowmask >> i) & 1u) side[i] = __ldg((const float4*)(b + i * ld4));
  } else if (ef_has(EPI, EF_DACT) && cx.side_kind == SIDE_DACT) {
    if (e.dact_in_bf16) {
      const bf16* b = (const bf16*)e.dact_in + idx0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if ((at.rowmask >> i) & 1u) {
          const uint2 t = __ldg((const uint2*)(b + i * ld4));
          side[i].x = __uint_as_float(t.x); side[i].y = __uint_as_float(t.y);
        }
    } else {
      const float* b = (const float*)e.dact_in + idx0;
#pragma unroll
      for (int i = 0; i < 8; ++i) if ((at.rowmask >> i) & 1u) side[i] = __ldg((const float4*)(b + i * ld4));
    }
  } else if (ef_has(EPI, EF_GATHER)) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((at.rowmask >> i) & 1u) {
        const int r = at.r0 + (lane >> 3) + 4 * i;
        const long long g = __ldg(e.gather_idx + (long long)at.bt * e.gather_idx_stride + r);
        side[i] = __ldg((const float4*)(e.gather_tab + g * e.gather_ld + at.n));
      }
  }
}

// the 4 bias values of this lane's columns: issued with the chunk's TMEM load (a dependent global load at the head of every
// chunk's arithmetic costs an L2 round trip per chunk: 1.5 us of a 4.5 us tile epilogue)
__device__ __forceinline__ float4 bias_load(const GemmEpilogue& e, const EpiCtx& cx, const TcParams& p, const ChunkAt& at) {
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!e.bias || !cx.vec_ok || !at.live || at.n >= p.N) return b;
  if (at.n + 4 <= p.N) return __ldg((const float4*)(e.bias + at.bias_off + at.n));
  b.x = __ldg(e.bias + at.bias_off + at.n);
  if (at.n + 1 < p.N) b.y = __ldg(e.bias + at.bias_off + at.n + 1);
  if (at.n + 2 < p.N) b.z = __ldg(e.bias + at.bias_off + at.n + 2);
  return b;
}

// the keep-bit words of this lane's 8 rows (one word per row, shared by the 8 lanes of a row group): issued with the chunk's
// TMEM load like the bias, for the same reason
struct KeepWords { uint32_t w[8]; };
template <int EPI>
__device__ __forceinline__ KeepWords bits_load(const GemmEpilogue& e, const TcParams& p, const ChunkAt& at, int lane) {
  KeepWords k;
#pragma unroll
  for (int i = 0; i < 8; ++i) k.w[i] = 0u;
  if (!ef_has(EPI, EF_DROPBITS) || !e.drop_bits || e.drop_p <= 0.f || !at.live || at.n >= p.N) return k;
  const uint32_t wpr = (uint32_t)p.N >> 5;
  const unsigned int* bw = e.drop_bits + ((long long)at.bt * p.M + at.r0 + (lane >> 3)) * wpr + (at.n >> 5);
#pragma unroll
  for (int i = 0; i < 8; ++i) if ((at.rowmask >> i) & 1u) k.w[i] = __ldg(bw + (unsigned)(4 * i) * wpr);
  return k;
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& e, const EpiCtx& cx, const TcParams& p, const ChunkAt& at, int lane,
                                               const float4* acc, const float4* side, const uint8_t* stg, const float4 bias4, const KeepWords& kw) {
  if (!at.live) return;
  if (ef_has(EPI, EF_SCALAR) && !cx.vec_ok) {        // ragged shapes (the 41-column head): element-wise, still row-contiguous across lanes
#pragma unroll 1           // ONE copy of the scalar epilogue (it carries every feature): code size, not speed, matters here
    for (int ij = 0; ij < 32; ++ij) {
      const int i = ij >> 2, j = ij & 3;
      const int r = at.r0 + (lane >> 3) + 4 * i;
      const int rr = (lane >> 3) + 4 * i;              // re-read from the staging tile: no dynamically indexed register array
      const float a = ((const float*)(stg + rr * 128 + ((((lane & 7) ^ rr) & 7) << 4)))[j];
      if (r < p.M && at.n + j < p.N) gemm_epilogue_store(e, p.N, p.M, a, at.bt, r, at.n + j);
    }
    return;
  }
  const bool col_ok = at.n < p.N;      // a last, partial group of 4 columns is stored whole: the row stride is padded (ldc % 4 == 0)
  float4 v[8];                         // and its accumulators beyond N are zero (out-of-range operand rows read as zero)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i].x = fmaf(acc[i].x, e.alpha, bias4.x); v[i].y = fmaf(acc[i].y, e.alpha, bias4.y);
    v[i].z = fmaf(acc[i].z, e.alpha, bias4.z); v[i].w = fmaf(acc[i].w, e.alpha, bias4.w);
  }
  // element offset of row i: idx0 + i * (4 * ldc); row-valid bits
  const long long idx0 = at.rowbase + at.n;
  const uint32_t ld4 = 4u * (uint32_t)e.ldc;       // (32-bit: a tile spans at most 32 rows of the output)
  const uint32_t okm = col_ok ? at.rowmask : 0u;
#define idx(i) (idx0 + (unsigned)((i) * ld4))
#define ok(i) ((okm >> (i)) & 1u)
  const bool deriv2 = ef_has(EPI, EF_OUT2) && ef_has(EPI, EF_ACT) && e.out2 && e.out2_deriv && e.act == ACT_GELU;
  if (ef_has(EPI, EF_OUT2) && e.out2 && !deriv2) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (ok(i)) {
        if (e.out2_bf16) *(uint2*)((bf16*)e.out2 + idx(i)) = pack4_bf16(v[i].x, v[i].y, v[i].z, v[i].w);
        else *(float4*)((float*)e.out2 + idx(i)) = v[i];
      }
  }
  if (!ef_has(EPI, EF_ACT)) {
  } else if (deriv2) {                     // GELU and its derivative from one erf / exp evaluation; out2 = GELU'(v)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 d; float cdf, ex;
      gelu_parts(v[i].x, cdf, ex); d.x = fmaf(v[i].x * 0.3989422804014327f, ex, cdf); v[i].x *= cdf;
      gelu_parts(v[i].y, cdf, ex); d.y = fmaf(v[i].y * 0.3989422804014327f, ex, cdf); v[i].y *= cdf;
      gelu_parts(v[i].z, cdf, ex); d.z = fmaf(v[i].z * 0.3989422804014327f, ex, cdf); v[i].z *= cdf;
      gelu_parts(v[i].w, cdf, ex); d.w = fmaf(v[i].w * 0.3989422804014327f, ex, cdf); v[i].w *= cdf;
      if (ok(i)) {
        if (e.out2_bf16) *(uint2*)((bf16*)e.out2 + idx(i)) = pack4_bf16(d.x, d.y, d.z, d.w);
        else *(float4*)((float*)e.out2 + idx(i)) = d;
      }
    }
  } else if (e.act == ACT_GELU) {          // (the activation switch is hoisted out of the element loops: smaller, branch-free code)
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i].x = gelu_fast(v[i].x); v[i].y = gelu_fast(v[i].y); v[i].z = gelu_fast(v[i].z); v[i].w = gelu_fast(v[i].w); }
  } else if (e.act == ACT_SOFTSIGN) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i].x *= rcp_approx(1.0f + fabsf(v[i].x)); v[i].y *= rcp_approx(1.0f + fabsf(v[i].y));
      v[i].z *= rcp_approx(1.0f + fabsf(v[i].z)); v[i].w *= rcp_approx(1.0f + fabsf(v[i].w));
    }
  } else if (e.act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f); }
  }
  if (ef_has(EPI, EF_GATHER) && e.gather_tab) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cx.side_kind == SIDE_GATHER) t = side[i];
      else if (ok(i)) {
        const int r = at.r0 + (lane >> 3) + 4 * i;
        const long long g = __ldg(e.gather_idx + (long long)at.bt * e.gather_idx_stride + r);
        t = __ldg((const float4*)(e.gather_tab + g * e.gather_ld + at.n));
      }
      if (ok(i)) { v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w; }
    }
  }
  if (ef_has(EPI, EF_DROPBITS) && e.drop_p > 0.f && e.drop_bits) {
    // keep bits drawn ahead of time on the side stream (they depend on the step's key only): 4 bits of one word per row,
    // the 8 lanes of a row group read the same word
    const float ik = 1.0f / (1.0f - e.drop_p);
    const int sh = at.n & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t nib = kw.w[i] >> sh;
      v[i].x *= (nib & 1u) ? ik : 0.f; v[i].y *= (nib & 2u) ? ik : 0.f; v[i].z *= (nib & 4u) ? ik : 0.f; v[i].w *= (nib & 8u) ? ik : 0.f;
    }
  }
  if (ef_has(EPI, EF_DROP) && e.drop_p > 0.f) {
    // One Philox block covers 8 consecutive elements = the columns of a lane PAIR.  The even lane draws the block of
    // row 2k, the odd lane the block of row 2k+1, and they swap halves: one Philox per 8 elements, as in the forward
    // of every other kernel that shares these streams.
    const uint32_t thr = drop_threshold(e.drop_p);
    const float ik = 1.0f / (1.0f - e.drop_p);
    const bool odd = lane & 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = at.r0 + (lane >> 3) + 4 * (2 * k + (odd ? 1 : 0));
      const unsigned long long elem = ((unsigned long long)at.bt * p.M + r) * (unsigned long long)p.N + (unsigned)(at.n & ~7);
      const Philox4 ph = philox4x32(cx.seed, elem >> 3, e.drop_stream);
      const uint32_t s0 = odd ? ph.x : ph.z, s1 = odd ? ph.y : ph.w;
      const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      const uint32_t a0 = odd ? r0 : ph.x, a1 = odd ? r1 : ph.y;     // row 2k
      const uint32_t b0 = odd ? ph.z : r0, b1 = odd ? ph.w : r1;     // row 2k+1
      float4& va = v[2 * k]; float4& vb = v[2 * k + 1];
      va.x *= (a0 & 0xFFFFu) >= thr ? ik : 0.f; va.y *= (a0 >> 16) >= thr ? ik : 0.f;
      va.z *= (a1 & 0xFFFFu) >= thr ? ik : 0.f; va.w *= (a1 >> 16) >= thr ? ik : 0.f;
      vb.x *= (b0 & 0xFFFFu) >= thr ? ik : 0.f; vb.y *= (b0 >> 16) >= thr ? ik : 0.f;
      vb.z *= (b1 & 0xFFFFu) >= thr ? ik : 0.f; vb.w *= (b1 >> 16) >= thr ? ik : 0.f;
    }
  }
  if (ef_has(EPI, EF_DACT) && e.dact != DACT_NONE) {
    float4 sv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cx.side_kind == SIDE_DACT) {
        s = e.dact_in_bf16 ? unpack4_bf16(make_uint2(__float_as_uint(side[i].x), __float_as_uint(side[i].y))) : side[i];
      } else if (ok(i)) {
        s = e.dact_in_bf16 ? unpack4_bf16(__ldg((const uint2*)((const bf16*)e.dact_in + idx(i))))
                           : __ldg((const float4*)((const float*)e.dact_in + idx(i)));
      }
      sv[i] = s;
    }
    if (e.dact == DACT_GELU_FROM_IN) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x *= dgelu_fast(sv[i].x); v[i].y *= dgelu_fast(sv[i].y); v[i].z *= dgelu_fast(sv[i].z); v[i].w *= dgelu_fast(sv[i].w);
      }
    } else if (e.dact == DACT_SAVED) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i].x *= sv[i].x; v[i].y *= sv[i].y; v[i].z *= sv[i].z; v[i].w *= sv[i].w; }
    } else if (e.dact == DACT_SOFTSIGN_FROM_OUT) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t;
        t = 1.0f - fabsf(sv[i].x); v[i].x *= t * t; t = 1.0f - fabsf(sv[i].y); v[i].y *= t * t;
        t = 1.0f - fabsf(sv[i].z); v[i].z *= t * t; t = 1.0f - fabsf(sv[i].w); v[i].w *= t * t;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x = sv[i].x > 0.f ? v[i].x : 0.f; v[i].y = sv[i].y > 0.f ? v[i].y : 0.f;
        v[i].z = sv[i].z > 0.f ? v[i].z : 0.f; v[i].w = sv[i].w > 0.f ? v[i].w : 0.f;
      }
    }
  }
  if (ef_has(EPI, EF_RESID) && e.resid) {           // always the prefetched side input when present
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (ok(i)) { v[i].x += side[i].x; v[i].y += side[i].y; v[i].z += side[i].z; v[i].w += side[i].w; }
  }
  if (ef_has(EPI, EF_COLSUM) && e.colsum) {          // bias gradient of the producing layer: column sums of this chunk, one red per column per warp
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (ok(i)) { cs.x += v[i].x; cs.y += v[i].y; cs.z += v[i].z; cs.w += v[i].w; }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (lane < 8 && col_ok) red_add_v4(e.colsum + at.n, cs.x, cs.y, cs.z, cs.w);
  }
  if (ef_has(EPI, EF_ACCUM) && e.accumulate) {
    float* o = (float*)e.out + idx0;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (ok(i)) red_add_v4(o + i * ld4, v[i].x, v[i].y, v[i].z, v[i].w);
  } else if (e.out_bf16) {
    bf16* o = (bf16*)e.out + idx0;
    if (okm == 0xFFu) {                      // interior tile (all but the last row tile): no per-row predicates
#pragma unroll
      for (int i = 0; i < 8; ++i) *(uint2*)(o + i * ld4) = pack4_bf16(v[i].x, v[i].y, v[i].z, v[i].w);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (ok(i)) *(uint2*)(o + i * ld4) = pack4_bf16(v[i].x, v[i].y, v[i].z, v[i].w);
    }
  } else {
    float* o = (float*)e.out + idx0;
    if (okm == 0xFFu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) *(float4*)(o + i * ld4) = v[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (ok(i)) *(float4*)(o + i * ld4) = v[i];
    }
  }
#undef idx
#undef ok
}


// ---------------------------------------------------------------------------
// CTA-pair (cta_group::2) helpers.  Two CTAs of a cluster (the two SMs of a TPC) issue ONE 256 x BN MMA: each holds
// its 128 rows of A and its half of B's columns in its own shared memory and its 128 accumulator rows in its own
// tensor memory.  Only the leader (cluster rank 0) issues tcgen05.mma; barriers that the leader waits on live in the
// leader's shared memory and are signalled remotely (TMA complete_tx of the peer's loads, the peer's epilogue arrive).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s counterpart in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope): a cluster-scope release compiles to MEMBAR.ALL.GPU, i.e. every epilogue warp would
  // wait for all of its outstanding global stores and side loads once per tile.  What the leader's MMA thread needs ordered
  // before it reuses the accumulator stage is the tcgen05.ld of this warp, which tcgen05.wait::ld + fence::before_thread_sync cover.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void tma_load_3d_g(void* dst, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1, int c2) {
  if (CTAS == 1) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void tc_mma_bf16_g(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CTAS == 1) {
    tc_mma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// MMA-completion arrive on the barrier at the same offset in every CTA of the pair
template <int CTAS>
__device__ __forceinline__ void tc_commit_g(uint64_t* bar) {
  if (CTAS == 1) {
    tc_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  }
}

// ---------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------

template <int BN, int MODE, int CTAS, int EPI>
__global__ void __launch_bounds__(kThreads, 1)   // 10 warps: 3 share one SM sub-partition -> 168 registers per thread
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcParams p) {
  constexpr int BNL = BN / CTAS;                       // columns of B held by this CTA
  constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
  constexpr int B_BYTES = BNL * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = kSmemPipe / STAGE_BYTES;
  constexpr bool A_MN = (MODE == GEMM_TN);
  constexpr bool B_MN = (MODE != GEMM_NT);
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CTAS) >> 4) << 24);
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr int NCH = BN / 64;                         // 32-column chunks per epilogue warp
  constexpr bool RAG = (CTAS == 2 && BN == 256 && MODE != GEMM_TN);     // instantiations that may be launched with ragged tiles

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_tiles = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = (uint64_t*)(stage_tiles + kEpiWarps * kStageTile);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;     // 0 = leader (issues the MMAs)
  const int unit = blockIdx.x / CTAS;                          // persistent work unit: a CTA (CTAS = 1) or a CTA pair
  const int n_units = gridDim.x / CTAS;
  if (threadIdx.x == 0) {
    dbg_stamp(p.dbg, 0);
    if (p.dbg) { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); p.dbg[(size_t)blockIdx.x * 16 + 15] = smid; }
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kEpiWarps * CTAS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();            // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 1);
  pdl_grid_sync();
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 2);       // everything above touched only shared / tensor memory and the kernel parameters

  const int total_kb = p.nchunk * p.kb_per_chunk;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = unit; tile < p.total_tiles; tile += n_units) {
        int mt, bz, n0;
        if (RAG && p.ragged) {
          const uint32_t r = p.rag[tile];
          mt = r & 63; bz = 0;
          n0 = ((r >> 6) & 63) * 64 + rank * ((int)((r >> 12) + 1) * 32);         // this CTA's half of the tile's columns
        } else {
          const int nt = tile % p.n_tiles;
          const int rest = tile / p.n_tiles;
          mt = rest % p.m_tiles;
          bz = rest / p.m_tiles;                  // trial (NT/NN) or split index (TN)
          n0 = nt * BN + rank * BNL;
        }
        const int m0 = mt * (BM * CTAS) + rank * BM;                             // this CTA's rows of A (n0: its columns of B)
        int bsel = 0;
        if (MODE == GEMM_NT && p.b_sel) { const long long d = __ldg(p.b_sel + bz); bsel = (int)(d < 0 ? 0 : (d >= p.b_sel_n ? p.b_sel_n - 1 : d)); }
        int kb0 = 0, kb1 = total_kb;
        if (MODE == GEMM_TN && p.split_k > 1) {
          const int per = (total_kb + p.split_k - 1) / p.split_k;
          kb0 = bz * per; kb1 = min(total_kb, kb0 + per);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const int j = kb / p.kb_per_chunk;
          const int kk = (kb - j * p.kb_per_chunk) * BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          // the pair's loads all complete on the LEADER's barrier (the leader's MMA thread is the only consumer)
          const uint32_t fb = CTAS == 2 ? mapa_u32(&full_bar[stage], 0) : smem_u32(&full_bar[stage]);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], STAGE_BYTES * CTAS);
          if (MODE == GEMM_TN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_3d_g<CTAS>(sa + i * (64 * BK * 2), &map_a, fb, m0 + 64 * i, kk + p.a_row_shift, j);
            const int cn = n0 / p.b_chunk_n;
            const int nc0 = n0 - cn * p.b_chunk_n;
#pragma unroll
            for (int i = 0; i < BNL / 64; ++i)
              tma_load_3d_g<CTAS>(sb + i * (64 * BK * 2), &map_b, fb, nc0 + 64 * i, kk + cn * p.b_row_shift, j);
          } else {
            tma_load_3d_g<CTAS>(sa, &map_a, fb, j * p.a_col_shift + kk, m0 + j * p.a_row_shift, bz);
            if (MODE == GEMM_NT) {
              tma_load_3d_g<CTAS>(sb, &map_b, fb, j * p.b_col_shift + kk, n0 + j * p.b_row_shift, bsel);
            } else {
#pragma unroll
              for (int i = 0; i < BNL / 64; ++i)
                tma_load_3d_g<CTAS>(sb + i * (64 * BK * 2), &map_b, fb, n0 + 64 * i + j * p.b_col_shift,
                                    kk + j * p.b_row_shift, 0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (tile == unit && kb == kb0) dbg_stamp(p.dbg, 3);
        }
      }
      dbg_stamp(p.dbg, 4);          // all loads issued
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc_stage = 0; uint32_t acc_phase = 0;
      for (int tile = unit; tile < p.total_tiles; tile += n_units) {
        int kb0 = 0, kb1 = total_kb;
        if (MODE == GEMM_TN && p.split_k > 1) {
          const int bz = (tile / p.n_tiles) / p.m_tiles;
          const int per = (total_kb + p.split_k - 1) / p.split_k;
          kb0 = bz * per; kb1 = min(total_kb, kb0 + per);
        }
        uint32_t idesc = IDESC;
        if (RAG && p.ragged) idesc = (IDESC & ~(0x3Fu << 17)) | ((((uint32_t)(p.rag[tile] >> 12) + 1) * 8u) << 17);   // N = 64 units: N >> 3 = 8 units
        mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc_stage * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (tile == unit && kb == kb0) dbg_stamp(p.dbg, 5);      // first operands landed
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          // one descriptor per operand and k-block; the four k-steps advance it with a 64-bit add
          const uint64_t ad0 = A_MN ? make_sdesc(sa, 64 * BK * 2, 1024) : make_sdesc(sa, 16, 1024);
          const uint64_t bd0 = B_MN ? make_sdesc(sb, 64 * BK * 2, 1024) : make_sdesc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_bf16_g<CTAS>(tmem_d, sdesc_advance(ad0, A_MN ? k * (16 * 128) : k * 32), sdesc_advance(bd0, B_MN ? k * (16 * 128) : k * 32), idesc,
                                (kb > kb0 || k > 0) ? 1u : 0u);
          tc_commit_g<CTAS>(&empty_bar[stage]);     // frees the smem slot (in both CTAs) when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_g<CTAS>(&tfull_bar[acc_stage]);   // accumulator ready for the epilogue warps (of both CTAs)
        dbg_stamp(p.dbg, 6);                        // (last write wins: issue of the last tile's final MMA)
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = ew >> 2;                     // which half of the tile's columns
    uint8_t* stg = stage_tiles + ew * kStageTile;
    const uint32_t stg_u32 = smem_u32(stg);
    const GemmEpilogue& e = p.epi;
    EpiCtx cx;
    const bool plain = !e.out2 && !e.gather_tab && e.drop_p <= 0.f && e.dact == DACT_NONE && !e.resid && !e.accumulate && !e.colsum;
    cx.vec_ok = (p.N % 4 == 0 || (plain && e.ldc >= p.N + (4 - p.N % 4))) && (e.ldc % 4 == 0) && (e.c_batch_stride % 4 == 0) &&
                (e.gather_tab == nullptr || e.gather_ld % 4 == 0) && (e.drop_p <= 0.f || p.N % 8 == 0);
    cx.seed = (ef_has(EPI, EF_DROP) && e.drop_p > 0.f) ? e.drop_seed.get() : 0ull;
    cx.side_kind = e.resid ? SIDE_RESID : (e.dact != DACT_NONE ? SIDE_DACT : (e.gather_tab ? SIDE_GATHER : SIDE_NONE));
    int acc_stage = 0; uint32_t acc_phase = 0;

    // tile coordinates (two integer divisions, one 64-bit multiply-add) once per tile, not per chunk
    auto tile_at = [&](int tile) {
      ChunkAt at;
      at.live = tile < p.total_tiles;
      int mt, bz, col0;
      at.nch = NCH; at.tcol = half * (BN / 2);
      if (RAG && p.ragged) {
        const uint32_t r = at.live ? p.rag[tile] : 0u;
        mt = r & 63; bz = 0;
        at.nch = (int)(r >> 12) + 1;                               // units of 64 columns = chunks of 32 per warp
        at.tcol = half * (at.nch * 32);
        col0 = ((r >> 6) & 63) * 64 + at.tcol;
      } else {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        mt = rest % p.m_tiles;
        bz = rest / p.m_tiles;
        col0 = nt * BN + half * (BN / 2);
      }
      at.bt = (MODE == GEMM_TN && !e.sel) ? 0 : bz;                // (routed GEMM_TN: the split index is the trial)
      at.r0 = mt * (BM * CTAS) + rank * BM + q * 32;
      at.n = col0 + (lane & 7) * 4;                                // chunk 0
      at.rowbase = (long long)at.bt * e.c_batch_stride + (long long)(at.r0 + (lane >> 3)) * e.ldc;
      at.rowmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) at.rowmask |= (at.r0 + (lane >> 3) + 4 * i) < p.M ? (1u << i) : 0u;
      at.bias_off = 0;
      if (e.sel && at.live) {
        long long d = __ldg(e.sel + bz);
        d = d < 0 ? 0 : (d >= e.sel_n ? e.sel_n - 1 : d);
        at.rowbase += d * e.c_sel_stride; at.bias_off = d * e.bias_sel_stride;
      }
      if (MODE == GEMM_TN && p.split_k > 1) {
        const int per = (total_kb + p.split_k - 1) / p.split_k;
        if (bz * per >= total_kb) at.live = false;       // empty split: nothing to add
      }
      return at;
    };
    auto chunk_of = [](ChunkAt t, int c) { t.n += c * 32; return t; };

    float4 side_cur[8], side_nxt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) side_cur[i] = side_nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ChunkAt t_cur = tile_at(unit);
    side_load<EPI>(e, cx, p, t_cur, lane, side_cur);
    const uint32_t tempty_addr0 = CTAS == 2 ? mapa_u32(&tempty_bar[0], 0) : 0u;   // the leader's accumulator-free barriers

    for (int tile = unit; tile < p.total_tiles; tile += n_units) {
      const ChunkAt t_nxt = tile_at(tile + n_units);
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tc_fence_after();
      if (ew == 0 && lane == 0) { if (tile == unit) dbg_stamp(p.dbg, 7); dbg_stamp(p.dbg, 8); }   // first / last accumulator ready
      const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_stage * BN + t_cur.tcol);
      const int nch = RAG ? t_cur.nch : NCH;
      // (Measured, round 2: issuing the tensor-memory read of chunk c + 1 before chunk c's arithmetic -- 32 more live registers --
      // changes nothing for the plain classes and costs the residual class 4.7 us per launch: the read is not what a chunk waits for.)
#pragma unroll 1                                   // one copy of the epilogue code: it must stay inside the instruction cache
      for (int c = 0; c < nch; ++c) {
        const ChunkAt at = chunk_of(t_cur, c);
        uint32_t raw[32];
        tmem_ld32(taddr_row + c * 32, raw);
        const float4 bias_cur = bias_load(e, cx, p, at);     // flies under the TMEM load and the transpose (not at the head of the arithmetic)
        const KeepWords kw_cur = bits_load<EPI>(e, p, at, lane);
        // side input of the next chunk (or of the next tile's first chunk) while the TMEM load is in flight
        const ChunkAt nx = (c + 1 < nch) ? chunk_of(t_cur, c + 1) : t_nxt;
        side_load<EPI>(e, cx, p, nx, lane, side_nxt);
        tmem_ld_wait();
        if (c == nch - 1) {                         // accumulator fully read: hand the TMEM stage back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 1) mbar_arrive(&tempty_bar[acc_stage]);
            else mbar_arrive_cluster(tempty_addr0 + acc_stage * 8);
          }
        }
        // transpose through the swizzled staging tile: lane = row -> lane = (row group, 4-column group)
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const uint32_t a = stg_u32 + lane * 128 + (((c4 ^ lane) & 7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(raw[4 * c4]), "r"(raw[4 * c4 + 1]), "r"(raw[4 * c4 + 2]),
                       "r"(raw[4 * c4 + 3]) : "memory");
        }
        __syncwarp();
        float4 acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = (lane >> 3) + 4 * i;
          const uint32_t a = stg_u32 + rr * 128 + ((((lane & 7) ^ rr) & 7) << 4);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc[i].x), "=f"(acc[i].y), "=f"(acc[i].z), "=f"(acc[i].w) : "r"(a) : "memory");
        }
        __syncwarp();
        epilogue_chunk<EPI>(e, cx, p, at, lane, acc, side_cur, stg, bias_cur, kw_cur);
#pragma unroll
        for (int i = 0; i < 8; ++i) side_cur[i] = side_nxt[i];
      }
      t_cur = t_nxt;
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
    if (lane == 0) dbg_stamp(p.dbg, 9 + (ew & 3));     // epilogue warps done (4 of the 8 recorded)
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) {                              // nobody leaves (or frees tensor memory) while the peer may still signal it;
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");    // pure execution barrier: a releasing arrive would be a
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");              // MEMBAR.ALL.GPU, i.e. wait for every store of the epilogue
  }
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    if (lane == 0) dbg_stamp(p.dbg, 14);
  }
}

// ---------------------------------------------------------------------------
// Host side: tensor maps and dispatch
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;

struct MapKey {
  const void* ptr; long long bs; int nb, rows, cols, ld, box_c, box_r, dev;     // (dev: the same address on two devices is two tensors)
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && bs == o.bs && nb == o.nb && rows == o.rows && cols == o.cols && ld == o.ld && box_c == o.box_c && box_r == o.box_r &&
           dev == o.dev;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    auto mix = [&](long long v) { h ^= (size_t)v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.bs); mix(k.nb); mix(k.rows); mix(k.cols); mix(k.ld); mix(k.box_c); mix(k.box_r); mix(k.dev);
    return h;
  }
};
// Encoded tensor maps, two generations: a hit in the old generation is promoted; when the young one is full it becomes the old one
// (so the maps of the buffers in use survive, those of freed buffers age out -- instead of dropping everything at a size limit).
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps, g_maps_old;
constexpr size_t kMapGeneration = 2048;

unsigned long long* g_gemm_dbg = nullptr;

int make_map(const GemmOperand& o, int box_cols, int box_rows, CUtensorMap* out) {
  NDT1_REQUIRE(((uintptr_t)o.ptr & 15) == 0, "gemm_tc: operand pointer not 16-byte aligned");
  NDT1_REQUIRE(o.ld % 8 == 0, "gemm_tc: operand row stride %d not a multiple of 8 elements", o.ld);
  NDT1_REQUIRE(o.nbatch <= 1 || o.batch_stride % 8 == 0, "gemm_tc: batch stride not a multiple of 8 elements");
  MapKey key{o.ptr, o.batch_stride, o.nbatch, o.rows, o.cols, o.ld, box_cols, box_rows, ndt1_current_device()};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  auto io = g_maps_old.find(key);
  if (io != g_maps_old.end()) {
    *out = io->second;
    if (g_maps.size() >= kMapGeneration) { g_maps_old.swap(g_maps); g_maps.clear(); }
    g_maps.emplace(key, *out);
    return 0;
  }
  cuuint64_t dims[3] = {(cuuint64_t)o.cols, (cuuint64_t)o.rows, (cuuint64_t)(o.nbatch > 0 ? o.nbatch : 1)};
  long long bs = o.batch_stride > 0 ? o.batch_stride : (long long)o.rows * o.ld;
  cuuint64_t strides[2] = {(cuuint64_t)o.ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)o.ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NDT1_REQUIRE(rc == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed rc=%d (cols=%d rows=%d nb=%d ld=%d bs=%lld box=%dx%d)",
               (int)rc, o.cols, o.rows, o.nbatch, o.ld, bs, box_cols, box_rows);
  if (g_maps.size() >= kMapGeneration) { g_maps_old.swap(g_maps); g_maps.clear(); }
  g_maps.emplace(key, *out);
  return 0;
}

template <int BN, int MODE, int CTAS, int EPI>
int launch_inst(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& tp, cudaStream_t stream) {
  constexpr int STAGE_BYTES = BM * BK * 2 + (BN / CTAS) * BK * 2;
  constexpr int STAGES = kSmemPipe / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + kEpiWarps * kStageTile + 1024 /*align*/ + (2 * STAGES + 4) * 8 + 16;
  static_assert(SMEM <= 227 * 1024, "gemm_tc: shared memory budget");
  static Ndt1PerDeviceFlag attr_set;          // (function attributes are per device)
  if (!attr_set.here()) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, CTAS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set.here() = true;
  }
  const int units = ndt1_num_sms() / CTAS;
  const int grid = CTAS * (tp.total_tiles < units ? tp.total_tiles : units);
  if (g_ndt1_prof_on)      // algorithmic FLOPs of this launch (bench.py roofline): 2 M N K over all trials / chunks
    ndt1_prof_note(2.0 * tp.M * (double)tp.N * (double)tp.nchunk * tp.chunk_k_valid * (tp.mode == GEMM_TN ? 1 : tp.nb_out), 0.0);
  NDT1_CUDA_CHECK(ndt1_launch_cluster(gemm_tc_kernel<BN, MODE, CTAS, EPI>, CTAS, dim3(grid), dim3(kThreads), SMEM, stream, ma, mb, tp));
  NDT1_CHECK_LAUNCH();
  return 0;
}

// the smallest class that covers what this launch asks for (everything else: the all-features instantiation)
int epilogue_class(const TcParams& tp, int bn) {
  const GemmEpilogue& e = tp.epi;
  int need = 0;
  if (e.out2) need |= EF_OUT2;
  if (e.act != ACT_NONE) need |= EF_ACT;
  if (e.gather_tab) need |= EF_GATHER;
  if (e.drop_p > 0.f) need |= (e.drop_bits && tp.N % 32 == 0) ? EF_DROPBITS : EF_DROP;     // (no class with the bits: falls to Philox, same mask)
  if (e.dact != DACT_NONE) need |= EF_DACT;
  if (e.resid) need |= EF_RESID;
  if (e.colsum) need |= EF_COLSUM;
  if (e.accumulate) need |= EF_ACCUM;
  const bool plain = need == 0;
  const bool vec = (tp.N % 4 == 0 || (plain && e.ldc >= tp.N + (4 - tp.N % 4))) && (e.ldc % 4 == 0) && (e.c_batch_stride % 4 == 0) &&
                   (e.gather_tab == nullptr || e.gather_ld % 4 == 0) && (e.drop_p <= 0.f || tp.N % 8 == 0);
  if (!vec || bn != 256) return EPI_ALL;
  // (only the classes launch_256 instantiates for this operand mode)
  const int nt[5] = {EPI_NONE, EPI_PLAIN, EPI_ACT, EPI_BITSRES, EPI_DROP}, nn[4] = {EPI_NONE, EPI_DMUL, EPI_BITSBWD, EPI_DACT}, tn[1] = {EPI_ACCUM};
  const int* classes = tp.mode == GEMM_NT ? nt : (tp.mode == GEMM_NN ? nn : tn);
  const int n = tp.mode == GEMM_NT ? 5 : (tp.mode == GEMM_NN ? 4 : 1);
  for (int c = 0; c < n; ++c)
    if ((need & ~classes[c]) == 0) return classes[c];
  if (need & EF_DROPBITS) {                 // no keep-bit class for this combination: the Philox classes give the same mask
    need = (need & ~EF_DROPBITS) | EF_DROP;
    for (int c = 0; c < n; ++c)
      if ((need & ~classes[c]) == 0) return classes[c];
  }
  return EPI_ALL;
}

template <int MODE, int CTAS>
int launch_256(int cls, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& tp, cudaStream_t stream) {
  // classes a mode can meet: forward (NT) plain / activation / dropout; data gradient (NN) plain / derivative; weight gradient (TN) accumulate
  if (MODE == GEMM_NT) {
    if (cls == EPI_NONE) return launch_inst<256, MODE, CTAS, EPI_NONE>(ma, mb, tp, stream);
    if (cls == EPI_PLAIN) return launch_inst<256, MODE, CTAS, EPI_PLAIN>(ma, mb, tp, stream);
    if (cls == EPI_ACT) return launch_inst<256, MODE, CTAS, EPI_ACT>(ma, mb, tp, stream);
    if (cls == EPI_BITSRES) return launch_inst<256, MODE, CTAS, EPI_BITSRES>(ma, mb, tp, stream);
    if (cls == EPI_DROP) return launch_inst<256, MODE, CTAS, EPI_DROP>(ma, mb, tp, stream);
  } else if (MODE == GEMM_NN) {
    if (cls == EPI_NONE) return launch_inst<256, MODE, CTAS, EPI_NONE>(ma, mb, tp, stream);
    if (cls == EPI_DMUL) return launch_inst<256, MODE, CTAS, EPI_DMUL>(ma, mb, tp, stream);
    if (cls == EPI_BITSBWD) return launch_inst<256, MODE, CTAS, EPI_BITSBWD>(ma, mb, tp, stream);
    if (cls == EPI_DACT) return launch_inst<256, MODE, CTAS, EPI_DACT>(ma, mb, tp, stream);
  } else {
    if (cls == EPI_ACCUM) return launch_inst<256, MODE, CTAS, EPI_ACCUM>(ma, mb, tp, stream);
  }
  return launch_inst<256, MODE, CTAS, EPI_ALL>(ma, mb, tp, stream);
}

template <int MODE>
int launch_mode(int bn, int ctas, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& tp, cudaStream_t stream) {
  if (bn == 256) {
    const int cls = epilogue_class(tp, bn);
    return ctas == 2 ? launch_256<MODE, 2>(cls, ma, mb, tp, stream) : launch_256<MODE, 1>(cls, ma, mb, tp, stream);
  }
  if (bn == 128) return ctas == 2 ? launch_inst<128, MODE, 2, EPI_ALL>(ma, mb, tp, stream) : launch_inst<128, MODE, 1, EPI_ALL>(ma, mb, tp, stream);
  return launch_inst<64, MODE, 1, EPI_ALL>(ma, mb, tp, stream);
}

}  // namespace

int tc_make_map(const GemmOperand& o, int box_cols, int box_rows, CUtensorMap* out) { return make_map(o, box_cols, box_rows, out); }

int gemm_tc_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  NDT1_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  NDT1_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
  int dev = 0, major = 0, minor = 0;
  NDT1_CUDA_CHECK(cudaGetDevice(&dev));
  NDT1_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  NDT1_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  NDT1_REQUIRE(major == 10, "gemm_tc: this library is built for sm_100a only (device is sm_%d%d)", major, minor);
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

namespace {
// Ragged column tiles (see TcParams::rag).  The regular tiling gives every strip ceil(U / 4) tiles of 4 units (64 columns each);
// with `waves` = the number of tiles the busiest pair gets, the plan keeps that number but raises the tile count to waves x pairs
// by cutting strips into more, narrower tiles (each strip's widths as even as possible, never below 2 units), then deals the
// tiles to the pairs in order of width, boustrophedon, so that the loads differ by at most one unit.  Adopted only if the busiest
// pair then carries fewer columns than before.
struct RagPlan { int S, U, pairs, total; unsigned short rag[kMaxRagTiles]; };
std::vector<RagPlan> g_rag_plans;      // one entry per (strips, units, pairs) seen: a handful of shapes per model (guarded by g_mu)

void plan_ragged_build(TcParams& tp, int pairs);
void plan_ragged(TcParams& tp, int pairs) {
  const int S = tp.m_tiles, U = tp.N / 64;
  std::lock_guard<std::mutex> lk(g_mu);
  for (const RagPlan& r : g_rag_plans)
    if (r.S == S && r.U == U && r.pairs == pairs) {
      if (r.total > 0) { memcpy(tp.rag, r.rag, sizeof(unsigned short) * r.total); tp.total_tiles = r.total; tp.ragged = 1; }
      return;
    }
  plan_ragged_build(tp, pairs);
  RagPlan r; r.S = S; r.U = U; r.pairs = pairs; r.total = tp.ragged ? tp.total_tiles : 0;
  if (tp.ragged) memcpy(r.rag, tp.rag, sizeof(unsigned short) * r.total);
  g_rag_plans.push_back(r);
}
void plan_ragged_build(TcParams& tp, int pairs) {
  const int S = tp.m_tiles, U = tp.N / 64;
  const int tmin = (U + 3) / 4, tmax = U / 2;
  const int regular = S * tmin;
  const int waves = (regular + pairs - 1) / pairs;
  int target = waves * pairs;
  if (target > S * tmax) target = S * tmax;
  if (target <= regular || target > kMaxRagTiles || U < 3) return;
  // tiles per strip: tmin everywhere, the extra ones spread over the strips (first strips get one more)
  std::vector<int> tcount(S, tmin);
  for (int extra = target - regular, i = 0; extra > 0; --extra, i = (i + 1) % S) {
    if (tcount[i] >= tmax) { bool any = false; for (int j = 0; j < S; ++j) any |= tcount[j] < tmax; if (!any) break; ++extra; continue; }
    ++tcount[i];
  }
  struct T { int strip, unit0, w; };
  std::vector<T> tiles;
  for (int sidx = 0; sidx < S; ++sidx) {
    const int t = tcount[sidx], base = U / t, rem = U % t;
    int u0 = 0;
    for (int j = 0; j < t; ++j) { const int w = base + (j < rem ? 1 : 0); tiles.push_back({sidx, u0, w}); u0 += w; }
  }
  const int n = (int)tiles.size();
  if (n > kMaxRagTiles) return;
  std::stable_sort(tiles.begin(), tiles.end(), [](const T& a, const T& b) { return a.w > b.w; });   // widest first, strip order kept inside a width
  // slot k * pairs + p belongs to pair p; odd rounds run backwards so that a pair that got a wide tile gets a narrow one next
  std::vector<int> load(pairs, 0);
  std::vector<T> slot((size_t)waves * pairs, T{-1, 0, 0});
  for (int i = 0; i < n; ++i) {
    const int k = i / pairs, j = i % pairs;
    const int pr = (k & 1) ? pairs - 1 - j : j;
    slot[(size_t)k * pairs + pr] = tiles[i];
    load[pr] += tiles[i].w;
  }
  int worst = 0;
  for (int pr = 0; pr < pairs; ++pr) worst = load[pr] > worst ? load[pr] : worst;
  if (worst >= waves * 4) return;                      // no better than the regular tiling
  // the kernel walks tile = pair, pair + pairs, ... < total_tiles: the table must have no holes before its end
  int total = 0;
  for (size_t i = 0; i < slot.size(); ++i) if (slot[i].strip >= 0) total = (int)i + 1;
  for (int i = 0; i < total; ++i) if (slot[i].strip < 0) return;
  for (int i = 0; i < total; ++i) tp.rag[i] = (unsigned short)(slot[i].strip | (slot[i].unit0 << 6) | ((slot[i].w - 1) << 12));
  tp.total_tiles = total;
  tp.ragged = 1;
}
}  // namespace

int gemm_tc_launch(const GemmProblem& p, cudaStream_t stream) {
  NDT1_TRY(gemm_tc_init());
  NDT1_REQUIRE(p.M > 0 && p.N > 0 && p.nb_out > 0 && p.nchunk > 0 && p.chunk_k > 0, "gemm_tc: empty problem");
  NDT1_REQUIRE(p.nchunk == 1 || p.mode == GEMM_TN || p.chunk_k % BK == 0, "gemm_tc: chunk_k=%d must be a multiple of %d", p.chunk_k, BK);
  NDT1_REQUIRE(p.split_k <= 1 || (p.mode == GEMM_TN && p.epi.accumulate), "gemm_tc: split_k only for accumulating GEMM_TN (0 = automatic)");
  if (p.mode == GEMM_TN) {
    NDT1_REQUIRE(p.chunk_k % BK == 0 || p.A.rows <= p.chunk_k + p.a_row_shift,
                 "gemm_tc: GEMM_TN reduction rows must be bounded by the A operand (rows=%d chunk_k=%d)", p.A.rows, p.chunk_k);
    NDT1_REQUIRE(p.b_chunk_n > 0, "gemm_tc: b_chunk_n must be set for GEMM_TN");
  }
  NDT1_REQUIRE(!p.b_sel || p.mode == GEMM_NT, "gemm_tc: b_sel (per-trial B operand) only for GEMM_NT");
  NDT1_REQUIRE(!p.epi.sel || p.mode != GEMM_TN || p.split_k == p.nchunk, "gemm_tc: a routed GEMM_TN needs split_k == nchunk (one split per trial)");
  NDT1_REQUIRE(!p.epi.drop_bits || p.N % 32 == 0, "gemm_tc: keep bits need N %% 32 == 0 (N=%d)", p.N);
  NDT1_REQUIRE(!p.epi.colsum || (p.N % 8 == 0 && p.epi.ldc % 4 == 0 && p.epi.c_batch_stride % 4 == 0),
               "gemm_tc: the fused column sum needs a vectorisable output (N=%d)", p.N);
  int bn = p.N > 128 ? 256 : (p.N > 64 ? 128 : 64);
  static const int force_bn = getenv("NDT1_GEMM_BN") ? atoi(getenv("NDT1_GEMM_BN")) : 0;   // tile-shape experiments only
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) bn = force_bn < bn ? force_bn : bn;
  if (p.mode == GEMM_TN && p.b_chunk_n % bn != 0) bn = (p.b_chunk_n % 128 == 0) ? 128 : 64;
  NDT1_REQUIRE(p.mode != GEMM_TN || p.b_chunk_n % bn == 0 || p.b_chunk_n >= p.N, "gemm_tc: b_chunk_n=%d incompatible with tile", p.b_chunk_n);

  // CTA pairs (256 x 256 tiles, cta_group::2) halve the shared-memory and L2 traffic per SM (each SM stages 128 of the
  // 256 columns of B); they are used when the tile is the full 256 columns and pairing does not cost an extra wave.
  static const int force_ctas = getenv("NDT1_GEMM_CTAS") ? atoi(getenv("NDT1_GEMM_CTAS")) : 0;
  int ctas = 1;
  if (bn >= 128 && p.M > BM) {
    const long long per = (long long)ndt1_cdiv(p.N, bn) * (p.mode == GEMM_TN ? 1 : p.nb_out);
    const long long t1 = per * ndt1_cdiv(p.M, BM), t2 = per * ndt1_cdiv(p.M, 2 * BM);
    const int num_sms = ndt1_num_sms();
    const long long w1 = (t1 + num_sms - 1) / num_sms, w2 = (t2 + num_sms / 2 - 1) / (num_sms / 2);
    if (p.mode == GEMM_TN || w2 <= w1) ctas = 2;     // (weight gradients size their split to one wave either way)
  }
  if (force_ctas == 1) ctas = 1;
  if (force_ctas == 2 && bn >= 128 && p.M > BM) ctas = 2;
  const int bm = BM * ctas;
  const int units = ndt1_num_sms() / ctas;

  TcParams tp;
  tp.mode = p.mode; tp.M = p.M; tp.N = p.N; tp.nb_out = p.nb_out;
  tp.nchunk = p.nchunk; tp.kb_per_chunk = ndt1_cdiv(p.chunk_k, BK); tp.chunk_k_valid = p.chunk_k;
  tp.a_row_shift = p.a_row_shift; tp.a_col_shift = p.a_col_shift;
  tp.b_row_shift = p.b_row_shift; tp.b_col_shift = p.b_col_shift;
  tp.b_chunk_n = p.b_chunk_n > 0 ? p.b_chunk_n : p.N;
  tp.b_sel = p.b_sel; tp.b_sel_n = p.B.nbatch;
  tp.m_tiles = ndt1_cdiv(p.M, bm); tp.n_tiles = ndt1_cdiv(p.N, bn);
  tp.split_k = p.split_k > 1 ? p.split_k : 1;
  if (p.mode == GEMM_TN && p.split_k == 0 && p.epi.accumulate) {
    // automatic split of the reduction over the work units (CTAs or CTA pairs): fewest waves x (main loop + epilogue);
    // every split adds a full fp32 red.add pass over the output, which is what the constant term charges
    const int tiles = tp.m_tiles * tp.n_tiles, kblocks = tp.nchunk * tp.kb_per_chunk;
    int best = 1; long long best_cost = -1;
    static const int max_split = getenv("NDT1_WGRAD_MAX_SPLIT") ? atoi(getenv("NDT1_WGRAD_MAX_SPLIT")) : 64;   // experiments only
    for (int sp = 1; sp <= max_split && sp <= 64 && (sp == 1 || sp * 4 <= kblocks); ++sp) {
      const long long waves = ((long long)tiles * sp + units - 1) / units;
      const long long cost = waves * (ndt1_cdiv(kblocks, sp) + 10);     // k-blocks per work item + ~10 k-blocks worth of fill / red.add epilogue
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
    }
    tp.split_k = best;
  }
  tp.total_tiles = tp.m_tiles * tp.n_tiles * (p.mode == GEMM_TN ? tp.split_k : p.nb_out);
  tp.epi = p.epi;
  tp.dbg = g_gemm_dbg;
  tp.ragged = 0;
  static const bool no_ragged = getenv("NDT1_GEMM_RAGGED") && getenv("NDT1_GEMM_RAGGED")[0] == '0';
  if (!no_ragged && p.mode != GEMM_TN && bn == 256 && ctas == 2 && p.nb_out == 1 && !p.b_sel && !p.epi.sel && p.N % 64 == 0 &&
      tp.m_tiles <= 64 && p.N / 64 <= 64)
    plan_ragged(tp, units);

  CUtensorMap ma, mb;
  if (p.mode == GEMM_TN) {
    NDT1_TRY(make_map(p.A, 64, BK, &ma));
    NDT1_TRY(make_map(p.B, 64, BK, &mb));
    return launch_mode<GEMM_TN>(bn, ctas, ma, mb, tp, stream);
  } else if (p.mode == GEMM_NN) {
    NDT1_TRY(make_map(p.A, BK, BM, &ma));
    NDT1_TRY(make_map(p.B, 64, BK, &mb));
    return launch_mode<GEMM_NN>(bn, ctas, ma, mb, tp, stream);
  }
  NDT1_TRY(make_map(p.A, BK, BM, &ma));
  NDT1_TRY(make_map(p.B, BK, bn / ctas, &mb));
  return launch_mode<GEMM_NT>(bn, ctas, ma, mb, tp, stream);
}

// ---- debugging hook (C ABI wrapper in api.cu) ----
void gemm_tc_set_timeline(unsigned long long* buf) { g_gemm_dbg = buf; }
