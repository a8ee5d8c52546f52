// Fast-mode GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) ->
// shared-memory ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) ->
// tcgen05.ld epilogue.  Persistent, warp specialised:
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  epilogue: TMEM -> registers -> fused epilogue -> global
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
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int kThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

struct TcParams {
  int mode;
  int M, N, nb_out;
  int nchunk, kb_per_chunk;            // k-blocks (of BK) per chunk
  int a_row_shift, a_col_shift, b_row_shift, b_col_shift, b_chunk_n;
  int split_k;
  int m_tiles, n_tiles, total_tiles;
  int chunk_k_valid;                   // reduction length per chunk in elements (for FLOP accounting)
  GemmEpilogue epi;
};

using namespace tc;

// ---------------------------------------------------------------------------
// Epilogue over 8 consecutive columns of one row (vector path).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_vec8(const GemmEpilogue& e, int n_total, int rows_c, const float* acc,
                                              int b, int r, int n) {
  const long long idx = (long long)b * e.c_batch_stride + (long long)r * e.ldc + n;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = acc[i] * e.alpha;
  if (e.bias) {
    const float4 b0 = *(const float4*)(e.bias + n), b1 = *(const float4*)(e.bias + n + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (e.out2) {
    if (e.out2_bf16) {
      __align__(16) bf16 o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __float2bfloat16_rn(v[i]);
      *(uint4*)((bf16*)e.out2 + idx) = *(const uint4*)o;
    } else {
      *(float4*)((float*)e.out2 + idx) = make_float4(v[0], v[1], v[2], v[3]);
      *(float4*)((float*)e.out2 + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  if (e.act != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_apply(e.act, v[i]);
  }
  if (e.gather_tab) {
    const long long g = e.gather_idx[(long long)b * e.gather_idx_stride + r];
    const float* t = e.gather_tab + g * e.gather_ld + n;
    const float4 t0 = *(const float4*)t, t1 = *(const float4*)(t + 4);
    v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
    v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
  }
  if (e.drop_p > 0.f) {
    const unsigned long long elem = ((unsigned long long)b * rows_c + r) * (unsigned long long)n_total + n;
    float ds[8];
    drop_scale_8(e.drop_seed, e.drop_stream, elem, drop_threshold(e.drop_p), 1.0f / (1.0f - e.drop_p), ds);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= ds[i];
  }
  if (e.dact != DACT_NONE) {
    float s[8];
    if (e.dact_in_bf16) {
      __align__(16) bf16 t[8];
      *(uint4*)t = *(const uint4*)((const bf16*)e.dact_in + idx);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = __bfloat162float(t[i]);
    } else {
      const float4 s0 = *(const float4*)((const float*)e.dact_in + idx);
      const float4 s1 = *(const float4*)((const float*)e.dact_in + idx + 4);
      s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= dact_apply(e.dact, s[i]);
  }
  if (e.resid) {
    const float4 r0 = *(const float4*)(e.resid + idx), r1 = *(const float4*)(e.resid + idx + 4);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (e.accumulate) {
    red_add_v4((float*)e.out + idx, v[0], v[1], v[2], v[3]);
    red_add_v4((float*)e.out + idx + 4, v[4], v[5], v[6], v[7]);
  } else if (e.out_bf16) {
    __align__(16) bf16 o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __float2bfloat16_rn(v[i]);
    *(uint4*)((bf16*)e.out + idx) = *(const uint4*)o;
  } else {
    *(float4*)((float*)e.out + idx) = make_float4(v[0], v[1], v[2], v[3]);
    *(float4*)((float*)e.out + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------
template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  constexpr int A_BYTES = BM * BK * 2;                 // 16 KB
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = kSmemBudget / STAGE_BYTES;
  constexpr bool A_MN = (MODE == GEMM_TN);
  constexpr bool B_MN = (MODE != GEMM_NT);
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_kb = p.nchunk * p.kb_per_chunk;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mt = rest % p.m_tiles;
        const int bz = rest / p.m_tiles;          // trial (NT/NN) or split index (TN)
        const int m0 = mt * BM, n0 = nt * BN;
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
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (MODE == GEMM_TN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_3d(sa + i * (64 * BK * 2), &map_a, &full_bar[stage], m0 + 64 * i, kk + p.a_row_shift, j);
            const int cn = n0 / p.b_chunk_n;
            const int nc0 = n0 - cn * p.b_chunk_n;
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_3d(sb + i * (64 * BK * 2), &map_b, &full_bar[stage], nc0 + 64 * i, kk + cn * p.b_row_shift, j);
          } else {
            tma_load_3d(sa, &map_a, &full_bar[stage], j * p.a_col_shift + kk, m0 + j * p.a_row_shift, bz);
            if (MODE == GEMM_NT) {
              tma_load_3d(sb, &map_b, &full_bar[stage], j * p.b_col_shift + kk, n0 + j * p.b_row_shift, 0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_3d(sb + i * (64 * BK * 2), &map_b, &full_bar[stage], n0 + 64 * i + j * p.b_col_shift,
                            kk + j * p.b_row_shift, 0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc_stage = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int kb0 = 0, kb1 = total_kb;
        if (MODE == GEMM_TN && p.split_k > 1) {
          const int bz = (tile / p.n_tiles) / p.m_tiles;
          const int per = (total_kb + p.split_k - 1) / p.split_k;
          kb0 = bz * per; kb1 = min(total_kb, kb0 + per);
        }
        mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc_stage * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? make_sdesc(sa + k * (16 * 128), 64 * BK * 2, 1024)
                                        : make_sdesc(sa + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_sdesc(sb + k * (16 * 128), 64 * BK * 2, 1024)
                                        : make_sdesc(sb + k * 32, 16, 1024);
            tc_mma_bf16(tmem_d, adesc, bdesc, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);           // frees the smem slot when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc_stage]);         // accumulator ready for the epilogue
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    int acc_stage = 0; uint32_t acc_phase = 0;
    const bool vec_ok = (p.N % 8 == 0) && (p.epi.ldc % 8 == 0) && (p.epi.c_batch_stride % 8 == 0) &&
                        (p.epi.gather_tab == nullptr || p.epi.gather_ld % 4 == 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = rest % p.m_tiles;
      const int bz = rest / p.m_tiles;
      const int bt = (MODE == GEMM_TN) ? 0 : bz;
      const int m0 = mt * BM, n0 = nt * BN;
      bool empty_split = false;
      if (MODE == GEMM_TN && p.split_k > 1) {
        const int per = (total_kb + p.split_k - 1) / p.split_k;
        empty_split = (bz * per >= total_kb);
      }
      mbar_wait(&tfull_bar[acc_stage], acc_phase);
      tc_fence_after();
      const int r = m0 + q * 32 + lane;
      const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_stage * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;                // warp-uniform
        uint32_t raw[32];
        tmem_ld32(taddr_row + c0, raw);
        tmem_ld_wait();
        if (r < p.M && !empty_split) {
          const float* acc = (const float*)raw;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int n = n0 + c0 + g * 8;
            if (vec_ok && n + 8 <= p.N) {
              epilogue_vec8(p.epi, p.N, p.M, acc + g * 8, bt, r, n);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (n + i < p.N) gemm_epilogue_store(p.epi, p.N, p.M, acc[g * 8 + i], bt, r, n + i);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Host side: tensor maps and dispatch
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::mutex g_mu;

struct MapKey {
  const void* ptr; long long bs; int nb, rows, cols, ld, box_c, box_r;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && bs == o.bs && nb == o.nb && rows == o.rows && cols == o.cols && ld == o.ld && box_c == o.box_c && box_r == o.box_r;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    auto mix = [&](long long v) { h ^= (size_t)v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.bs); mix(k.nb); mix(k.rows); mix(k.cols); mix(k.ld); mix(k.box_c); mix(k.box_r);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// optional per-launch timing of the tensor-core GEMM (bench.py roofline): CUDA events on the launch stream
struct ProfRec { cudaEvent_t e0, e1; double flops; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
size_t g_prof_used = 0;

int make_map(const GemmOperand& o, int box_cols, int box_rows, CUtensorMap* out) {
  NDT1_REQUIRE(((uintptr_t)o.ptr & 15) == 0, "gemm_tc: operand pointer not 16-byte aligned");
  NDT1_REQUIRE(o.ld % 8 == 0, "gemm_tc: operand row stride %d not a multiple of 8 elements", o.ld);
  NDT1_REQUIRE(o.nbatch <= 1 || o.batch_stride % 8 == 0, "gemm_tc: batch stride not a multiple of 8 elements");
  MapKey key{o.ptr, o.batch_stride, o.nbatch, o.rows, o.cols, o.ld, box_cols, box_rows};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
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
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, *out);
  return 0;
}

template <int BN, int MODE>
int launch_inst(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& tp, cudaStream_t stream) {
  constexpr int STAGE_BYTES = BM * BK * 2 + BN * BK * 2;
  constexpr int STAGES = kSmemBudget / STAGE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + (2 * STAGES + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  const int grid = tp.total_tiles < g_num_sms ? tp.total_tiles : g_num_sms;
  ProfRec* rec = nullptr;
  if (g_prof_on) {
    if (g_prof_used == g_prof.size()) {
      ProfRec r; r.flops = 0;
      NDT1_CUDA_CHECK(cudaEventCreate(&r.e0)); NDT1_CUDA_CHECK(cudaEventCreate(&r.e1));
      g_prof.push_back(r);
    }
    rec = &g_prof[g_prof_used++];
    rec->flops = 2.0 * tp.M * (double)tp.N * (double)tp.nchunk * tp.chunk_k_valid * (tp.mode == GEMM_TN ? 1 : tp.nb_out);
    NDT1_CUDA_CHECK(cudaEventRecord(rec->e0, stream));
  }
  gemm_tc_kernel<BN, MODE><<<grid, kThreads, SMEM, stream>>>(ma, mb, tp);
  NDT1_CHECK_LAUNCH();
  if (rec) NDT1_CUDA_CHECK(cudaEventRecord(rec->e1, stream));
  return 0;
}

template <int MODE>
int launch_mode(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& tp, cudaStream_t stream) {
  if (bn == 256) return launch_inst<256, MODE>(ma, mb, tp, stream);
  if (bn == 128) return launch_inst<128, MODE>(ma, mb, tp, stream);
  return launch_inst<64, MODE>(ma, mb, tp, stream);
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
  int dev = 0;
  NDT1_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  NDT1_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  NDT1_REQUIRE(prop.major == 10, "gemm_tc: this library is built for sm_100a only (device is sm_%d%d)", prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

int gemm_tc_launch(const GemmProblem& p, cudaStream_t stream) {
  NDT1_TRY(gemm_tc_init());
  NDT1_REQUIRE(p.M > 0 && p.N > 0 && p.nb_out > 0 && p.nchunk > 0 && p.chunk_k > 0, "gemm_tc: empty problem");
  NDT1_REQUIRE(p.nchunk == 1 || p.mode == GEMM_TN || p.chunk_k % BK == 0, "gemm_tc: chunk_k=%d must be a multiple of %d", p.chunk_k, BK);
  NDT1_REQUIRE(p.split_k <= 1 || (p.mode == GEMM_TN && p.epi.accumulate), "gemm_tc: split_k only for accumulating GEMM_TN");
  if (p.mode == GEMM_TN) {
    NDT1_REQUIRE(p.chunk_k % BK == 0 || p.A.rows <= p.chunk_k + p.a_row_shift,
                 "gemm_tc: GEMM_TN reduction rows must be bounded by the A operand (rows=%d chunk_k=%d)", p.A.rows, p.chunk_k);
    NDT1_REQUIRE(p.b_chunk_n > 0, "gemm_tc: b_chunk_n must be set for GEMM_TN");
  }
  int bn = p.N > 128 ? 256 : (p.N > 64 ? 128 : 64);
  if (p.mode == GEMM_TN && p.b_chunk_n % bn != 0) bn = (p.b_chunk_n % 128 == 0) ? 128 : 64;
  NDT1_REQUIRE(p.mode != GEMM_TN || p.b_chunk_n % bn == 0 || p.b_chunk_n >= p.N, "gemm_tc: b_chunk_n=%d incompatible with tile", p.b_chunk_n);

  TcParams tp;
  tp.mode = p.mode; tp.M = p.M; tp.N = p.N; tp.nb_out = p.nb_out;
  tp.nchunk = p.nchunk; tp.kb_per_chunk = ndt1_cdiv(p.chunk_k, BK); tp.chunk_k_valid = p.chunk_k;
  tp.a_row_shift = p.a_row_shift; tp.a_col_shift = p.a_col_shift;
  tp.b_row_shift = p.b_row_shift; tp.b_col_shift = p.b_col_shift;
  tp.b_chunk_n = p.b_chunk_n > 0 ? p.b_chunk_n : p.N;
  tp.split_k = p.split_k > 1 ? p.split_k : 1;
  tp.m_tiles = ndt1_cdiv(p.M, BM); tp.n_tiles = ndt1_cdiv(p.N, bn);
  tp.total_tiles = tp.m_tiles * tp.n_tiles * (p.mode == GEMM_TN ? tp.split_k : p.nb_out);
  tp.epi = p.epi;

  CUtensorMap ma, mb;
  if (p.mode == GEMM_TN) {
    NDT1_TRY(make_map(p.A, 64, BK, &ma));
    NDT1_TRY(make_map(p.B, 64, BK, &mb));
    return launch_mode<GEMM_TN>(bn, ma, mb, tp, stream);
  } else if (p.mode == GEMM_NN) {
    NDT1_TRY(make_map(p.A, BK, BM, &ma));
    NDT1_TRY(make_map(p.B, 64, BK, &mb));
    return launch_mode<GEMM_NN>(bn, ma, mb, tp, stream);
  }
  NDT1_TRY(make_map(p.A, BK, BM, &ma));
  NDT1_TRY(make_map(p.B, BK, bn, &mb));
  return launch_mode<GEMM_NT>(bn, ma, mb, tp, stream);
}

// ---- profiling hooks (C ABI wrappers in api.cu) ----
int gemm_tc_profile_begin() {
  g_prof_used = 0; g_prof_on = true;
  return 0;
}
int gemm_tc_profile_end(double* flops, double* ms, long long* launches) {
  g_prof_on = false;
  double f = 0, t = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    NDT1_CUDA_CHECK(cudaEventSynchronize(g_prof[i].e1));
    float m = 0;
    NDT1_CUDA_CHECK(cudaEventElapsedTime(&m, g_prof[i].e0, g_prof[i].e1));
    f += g_prof[i].flops; t += m;
  }
  *flops = f; *ms = t; *launches = (long long)g_prof_used;
  return 0;
}
