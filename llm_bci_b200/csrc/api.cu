// C-ABI wrappers of the stand-alone operators declared in include/ndt1_b200.h.
#include "../../include/ndt1_b200.h"
#include "kernels.cuh"
#include <string.h>

namespace {
// g = dy * act'(saved), cast to T with a padded row stride (the operand of the data- and weight-gradient GEMMs)
template <typename T>
__global__ void dact_prep_kernel(const float* __restrict__ dy, const float* __restrict__ saved, T* __restrict__ out, long long rows, int cols,
                                 int ld_out, int dact, float drop_p, unsigned long long seed, unsigned long long site) { pdl_grid_sync();
  const long long total = rows * ld_out;
  const uint32_t thr = drop_threshold(drop_p);
  const float ik = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_out; const int c = (int)(i % ld_out);
    float v = 0.f;
    if (c < cols) {
      const long long e = r * cols + c;
      v = dy[e];
      if (drop_p > 0.f) v *= drop_scale_1(seed, site, (unsigned long long)e, thr, ik);     // the forward's keep mask (output dropout)
      if (dact != DACT_NONE) v *= dact_apply(dact, saved[e]);
    }
    out[i] = from_f32<T>(v);
  }
}
// nn.CrossEntropyLoss(reduction="none").sum() over rows of (B, V) logits with int64 class labels: one warp per row;
// loss += sum_b (logsumexp(z_b) - z_b[label_b]);  dlogits = (softmax(z) - onehot(label)) * dloss
__global__ void xent_kernel(const float* __restrict__ z, const long long* __restrict__ labels, float* __restrict__ dz, float* __restrict__ loss,
                            int B, int V, const float* __restrict__ dloss) { pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* zr = z + (long long)row * V;
  float m = -INFINITY;
  for (int j = lane; j < V; j += 32) m = fmaxf(m, zr[j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < V; j += 32) sum += expf(zr[j] - m);
  sum = warp_sum(sum);
  const float lse = m + logf(sum);
  long long lab = labels[row];
  lab = lab < 0 ? 0 : (lab >= V ? V - 1 : lab);
  const float gs = dloss ? *dloss : 1.f;
  if (dz)
    for (int j = lane; j < V; j += 32) dz[(long long)row * V + j] = (expf(zr[j] - lse) - (j == (int)lab ? 1.f : 0.f)) * gs;
  if (lane == 0 && loss) atomicAdd(loss, lse - zr[lab]);
}

// the same for cols % 8 == 0 and an unpadded bf16 operand (ld_out == cols): 8 consecutive elements per thread -- two 16-byte reads
// per input, one Philox block (the 8 elements are exactly one block of the flat index), one 16-byte write
__global__ void dact_prep8_kernel(const float* __restrict__ dy, const float* __restrict__ saved, bf16* __restrict__ out, long long total8, int dact,
                                  float drop_p, unsigned long long seed, unsigned long long site) { pdl_grid_sync();
  const uint32_t thr = drop_threshold(drop_p);
  const float ik = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    const float4 a = *(const float4*)(dy + e), b = *(const float4*)(dy + e + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (drop_p > 0.f) {
      float k[8];
      drop_scale_8(seed, site, (unsigned long long)e, thr, ik, k);
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] *= k[u];
    }
    if (dact != DACT_NONE) {
      const float4 c = *(const float4*)(saved + e), d = *(const float4*)(saved + e + 4);
      const float sv[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] *= dact_apply(dact, sv[u]);
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 o; o.x = *(uint32_t*)&p0; o.y = *(uint32_t*)&p1; o.z = *(uint32_t*)&p2; o.w = *(uint32_t*)&p3;
    *(uint4*)(out + e) = o;
  }
}
}  // namespace

namespace {
// out[b, j, :] = j < d ? a[b, j, :] : (j < d + Ls ? ins[b, j - d, :] : a[b, j - Ls, :]),  d = split[b] clamped to [0, La]
// (models/bci.py:143-166: the spike features spliced between the two halves of the prompt; the same rule for the attention
// mask and, with a constant row instead of `ins`, for the targets).  `unsplice` runs it backwards: the gradient of `out`
// scattered back to a (accumulating nothing: every row has exactly one source).
template <typename T>
__global__ void splice_rows_kernel(const T* __restrict__ a, const T* __restrict__ ins, const long long* __restrict__ split, T* __restrict__ out,
                                   int B, int La, int Ls, int W, int use_fill, T fill) { pdl_grid_sync();
  const long long total = (long long)B * (La + Ls) * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W);
    const long long row = i / W;
    const int j = (int)(row % (La + Ls)), b = (int)(row / (La + Ls));
    long long d = split[b]; d = d < 0 ? 0 : (d > La ? La : d);
    T v;
    if (j < d) v = a[((long long)b * La + j) * W + c];
    else if (j < d + Ls) v = use_fill ? fill : ins[((long long)b * Ls + (j - d)) * W + c];
    else v = a[((long long)b * La + (j - Ls)) * W + c];
    out[i] = v;
  }
}
template <typename T>
__global__ void unsplice_rows_kernel(const T* __restrict__ dout, const long long* __restrict__ split, T* __restrict__ da, T* __restrict__ dins,
                                     int B, int La, int Ls, int W) { pdl_grid_sync();
  const long long total = (long long)B * (La + Ls) * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W);
    const long long row = i / W;
    const int j = (int)(row % (La + Ls)), b = (int)(row / (La + Ls));
    long long d = split[b]; d = d < 0 ? 0 : (d > La ? La : d);
    const T v = dout[i];
    if (j < d) { if (da) da[((long long)b * La + j) * W + c] = v; }
    else if (j < d + Ls) { if (dins) dins[((long long)b * Ls + (j - d)) * W + c] = v; }
    else if (da) da[((long long)b * La + (j - Ls)) * W + c] = v;
  }
}
// features (B, T, H) with a validity mask (B, T): rows padded with zeros to a multiple of `stacking`, then
// mask_out[b, r] = all `stacking` rows of group r valid (models/bci.py:127-141)
__global__ void stack_valid_kernel(const long long* __restrict__ mask, long long* __restrict__ out, int B, int T, int stacking, int Ts) { pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ts) return;
  const int b = i / Ts, r = i % Ts;
  long long ok = 1;
  for (int k = 0; k < stacking; ++k) {
    const int t = r * stacking + k;
    ok &= (t < T && mask[(long long)b * T + t] != 0) ? 1 : 0;
  }
  out[i] = ok;
}
}  // namespace

extern "C" {

int ndt1_smooth_noise(const float* x, float* out, int B, int T, int N, const float* taps_host, int K, float white_sd,
                      float offset_sd, const float* white, const float* offset, int use_philox, uint64_t seed, const uint64_t* seed_ptr,
                      void* out_bf16, int ld_bf16, void* stream) {
  const SeedRef sr = seed_ptr ? SeedRef::at((const unsigned long long*)seed_ptr) : SeedRef((unsigned long long)seed);
  return k_smooth_noise(x, out, B, T, N, taps_host, K, white_sd, offset_sd, white, offset, use_philox, sr, (cudaStream_t)stream, (bf16*)out_bf16,
                        ld_bf16);
}

int ndt1_masker_apply(float* spikes, int B, int T, int N, int mode, int timespan, const uint8_t* mask_draw, const uint8_t* zero_draw,
                      const uint8_t* random_draw, const float* rand, int64_t* mask_out, int64_t* targets_mask, void* scratch,
                      void* stream) {
  NDT1_REQUIRE(spikes && mask_draw && zero_draw && random_draw && rand && scratch, "masker_apply: null argument");
  return k_masker_apply(spikes, B, T, N, mode, timespan, mask_draw, zero_draw, random_draw, rand, (long long*)mask_out,
                        (long long*)targets_mask, (unsigned int*)scratch, (cudaStream_t)stream);
}
int ndt1_bernoulli_u8(uint8_t* out, int64_t n, float prob, uint64_t seed, uint64_t stream_id, void* stream) {
  return k_bernoulli_u8(out, n, prob, seed, stream_id, (cudaStream_t)stream);
}
int ndt1_uniform_f32(float* out, int64_t n, uint64_t seed, uint64_t stream_id, void* stream) {
  return k_uniform_f32(out, n, seed, stream_id, (cudaStream_t)stream);
}

int ndt1_pad_pack(const void* src, const int64_t* offsets, void* dst, int B, int P, int inner, int elem_size, int side_left, int full,
                  double value, void* stream) {
  return k_pad_pack(src, (const long long*)offsets, dst, B, P, inner, elem_size, side_left, full, value, (cudaStream_t)stream);
}

size_t ndt1_ctc_workspace_bytes(int B, int L, int S) { return k_ctc_workspace_floats(B, L, S) * sizeof(float); }

int ndt1_ctc_loss(const float* logits, float* logp, const int64_t* targets, const int64_t* input_lengths, const int64_t* target_lengths,
                  int B, int L, int V, int S, int blank, int zero_infinity, void* workspace, float* nll, float* loss, float* dlogits,
                  const float* dloss, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  NDT1_TRY(k_log_softmax(logits, logp, (long long)B * L, V, s));
  return k_ctc_fwd_bwd(logp, (const long long*)targets, (const long long*)input_lengths, (const long long*)target_lengths, B, L, V, S,
                       blank, zero_infinity, (float*)workspace, nll, loss, dlogits, dloss, s);
}
int ndt1_edit_distance(const int64_t* pred_ids, const int64_t* pred_len, int Lp, const int64_t* target_ids, const int64_t* target_len, int Lt,
                       int B, int64_t* errors, void* stream) {
  return k_edit_distance((const long long*)pred_ids, (const long long*)pred_len, Lp, (const long long*)target_ids, (const long long*)target_len, Lt,
                         B, (long long*)errors, (cudaStream_t)stream);
}
int ndt1_ctc_greedy_decode(const float* logp, int B, int L, int V, int blank, int64_t* out_ids, int64_t* out_len, void* stream) {
  return k_ctc_greedy_decode(logp, B, L, V, blank, (long long*)out_ids, (long long*)out_len, (cudaStream_t)stream);
}

int ndt1_recon_loss(const float* pred, const float* target, float* dpred, const int64_t* targets_mask, const int64_t* pad_mask, int B,
                    int T, int N, int loss_kind, int shift_by_one, int relu_out, float* loss, int64_t* count, const float* dloss,
                    void* stream) {
  return k_recon_loss(pred, target, dpred, (const long long*)targets_mask, (const long long*)pad_mask, B, T, N, loss_kind, shift_by_one,
                      relu_out, loss, (long long*)count, dloss, (cudaStream_t)stream);
}

int ndt1_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int64_t rows, int H,
                       float eps, void* stream) {
  return k_layernorm_fwd<float>(x, gamma, beta, y, mean, rstd, rows, H, eps, (cudaStream_t)stream);
}

static int act_code_of(int act) {
  return act == NDT1_ACT_SOFTSIGN ? ACT_SOFTSIGN : act == NDT1_ACT_GELU ? ACT_GELU : act == NDT1_ACT_RELU ? ACT_RELU : ACT_NONE;
}
static GemmProblem plain_problem(int mode, int M, int N, int K) {
  GemmProblem p;
  p.b_sel = nullptr;
  p.mode = mode; p.M = M; p.N = N; p.nb_out = 1; p.nchunk = 1; p.chunk_k = K;
  p.a_row_shift = p.a_col_shift = p.b_row_shift = p.b_col_shift = 0; p.b_chunk_n = N; p.split_k = 1;
  p.epi = gemm_epilogue_default();
  return p;
}
static bf16* ws_take(char*& cur, size_t elems) {
  cur = (char*)(((uintptr_t)cur + 255) & ~(uintptr_t)255);
  bf16* r = (bf16*)cur;
  cur += elems * sizeof(bf16);
  return r;
}

size_t ndt1_linear_workspace_bytes(int M, int N, int K) {
  const size_t ldk = (K + 7) / 8 * 8, ldn = (N + 7) / 8 * 8;
  return ((size_t)M * ldk + (size_t)N * ldk + (size_t)M * ldn) * sizeof(bf16) + 4 * 256;
}

int ndt1_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* pre, int M, int N, int K, int act, int precision,
                    void* workspace, size_t workspace_bytes, void* stream) {
  return ndt1_linear_drop_fwd(x, w, bias, y, pre, M, N, K, act, precision, workspace, workspace_bytes, 0.f, 0, 0, stream);
}

int ndt1_linear_drop_fwd(const float* x, const float* w, const float* bias, float* y, float* pre, int M, int N, int K, int act, int precision,
                         void* workspace, size_t workspace_bytes, float drop_p, uint64_t seed, uint64_t site, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (M == 0 || N == 0) return 0;
  NDT1_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "linear_fwd: dropout p must be in [0,1)");
  GemmProblem p = plain_problem(GEMM_NT, M, N, K);
  p.epi.out = y; p.epi.ldc = N; p.epi.bias = bias; p.epi.act = act_code_of(act);
  if (drop_p > 0.f) { p.epi.drop_p = drop_p; p.epi.drop_seed = SeedRef(seed); p.epi.drop_stream = site; }
  if (pre) { p.epi.out2 = pre; p.epi.out2_bf16 = 0; }
  if (precision == NDT1_PRECISION_FP32) {
    p.A = {x, 0, 1, M, K, K}; p.B = {w, 0, 1, N, K, K};
    return gemm_simt_launch(p, 0, s);
  }
  const int ldk = (K + 7) / 8 * 8;
  NDT1_REQUIRE(workspace && workspace_bytes >= ndt1_linear_workspace_bytes(M, N, K), "linear_fwd: bf16 mode needs %zu workspace bytes",
               ndt1_linear_workspace_bytes(M, N, K));
  char* cur = (char*)workspace;
  bf16* xa = ws_take(cur, (size_t)M * ldk);
  bf16* wa = ws_take(cur, (size_t)N * ldk);
  NDT1_TRY(k_cast_f32_bf16(x, xa, M, K, K, ldk, s));
  NDT1_TRY(k_cast_f32_bf16(w, wa, N, K, K, ldk, s));
  p.A = {xa, 0, 1, M, K, ldk}; p.B = {wa, 0, 1, N, K, ldk};
  return gemm_tc_launch(p, s);
}


int ndt1_linear_bwd(const float* dy, const float* x, const float* w, const float* saved, float* dx, float* dw, float* db, int M, int N,
                    int K, int act, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  return ndt1_linear_drop_bwd(dy, x, w, saved, dx, dw, db, M, N, K, act, precision, workspace, workspace_bytes, 0.f, 0, 0, stream);
}

int ndt1_linear_drop_bwd(const float* dy, const float* x, const float* w, const float* saved, float* dx, float* dw, float* db, int M, int N,
                         int K, int act, int precision, void* workspace, size_t workspace_bytes, float drop_p, uint64_t seed, uint64_t site,
                         void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (M == 0 || N == 0) return 0;
  NDT1_REQUIRE(dy && x && w, "linear_bwd: null argument");
  const int dact = act == NDT1_ACT_SOFTSIGN ? DACT_SOFTSIGN_FROM_OUT : act == NDT1_ACT_GELU ? DACT_GELU_FROM_IN
                 : act == NDT1_ACT_RELU ? DACT_RELU_FROM_OUT : DACT_NONE;
  NDT1_REQUIRE(dact == DACT_NONE || saved, "linear_bwd: the activation derivative needs the saved output (pre-activation for gelu)");
  const int ldk = (K + 7) / 8 * 8, ldn = (N + 7) / 8 * 8;
  NDT1_REQUIRE(workspace && workspace_bytes >= ndt1_linear_workspace_bytes(M, N, K), "linear_bwd: needs %zu workspace bytes",
               ndt1_linear_workspace_bytes(M, N, K));
  char* cur = (char*)workspace;
  const long long tot = (long long)M * ldn;
  const int blocks = (int)((tot + 255) / 256 < 148 * 8 ? (tot + 255) / 256 : 148 * 8);
  if (precision == NDT1_PRECISION_FP32) {
    // fp32: g (M, N) in the workspace (2 bf16 slots per float: the byte budget of the bf16 layout covers M * N floats only when
    // ldk >= N; keep it simple and exact instead: g needs M * N floats)
    NDT1_REQUIRE(workspace_bytes >= (size_t)M * N * 4 + 256, "linear_bwd: fp32 mode needs %zu workspace bytes", (size_t)M * N * 4 + 256);
    float* g = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    ndt1_launch(dact_prep_kernel<float>, blocks, 256, 0, s, dy, saved, g, (long long)M, N, N, dact, drop_p, (unsigned long long)seed, (unsigned long long)site);
    NDT1_CHECK_LAUNCH();
    if (db) NDT1_TRY(k_colsum<float>(g, db, M, N, N, s));
    if (dw) {
      GemmProblem p = plain_problem(GEMM_TN, N, K, M);
      p.b_chunk_n = K;
      p.A = {g, 0, 1, M, N, N}; p.B = {x, 0, 1, M, K, K};
      p.epi.out = dw; p.epi.ldc = K; p.epi.accumulate = 1;
      NDT1_TRY(gemm_simt_launch(p, 0, s));
    }
    if (dx) {
      GemmProblem p = plain_problem(GEMM_NN, M, K, N);
      p.A = {g, 0, 1, M, N, N}; p.B = {w, 0, 1, N, K, K};
      p.epi.out = dx; p.epi.ldc = K;
      NDT1_TRY(gemm_simt_launch(p, 0, s));
    }
    return 0;
  }
  bf16* xa = ws_take(cur, (size_t)M * ldk);
  bf16* wa = ws_take(cur, (size_t)N * ldk);
  bf16* ga = ws_take(cur, (size_t)M * ldn);
  if (N % 8 == 0 && (((uintptr_t)dy | (uintptr_t)saved) & 15) == 0) {
    const long long t8 = (long long)M * N / 8;
    ndt1_launch(dact_prep8_kernel, (int)((t8 + 255) / 256 < 148 * 8 ? (t8 + 255) / 256 : 148 * 8), 256, 0, s, dy, saved, ga, t8, dact, drop_p,
                (unsigned long long)seed, (unsigned long long)site);
  } else {
    ndt1_launch(dact_prep_kernel<bf16>, blocks, 256, 0, s, dy, saved, ga, (long long)M, N, ldn, dact, drop_p, (unsigned long long)seed, (unsigned long long)site);
  }
  NDT1_CHECK_LAUNCH();
  if (db) NDT1_TRY(k_colsum<bf16>(ga, db, M, N, ldn, s));
  if (dw) {
    NDT1_TRY(k_cast_f32_bf16(x, xa, M, K, K, ldk, s));
    GemmProblem p = plain_problem(GEMM_TN, N, K, M);
    p.b_chunk_n = K;
    p.A = {ga, 0, 1, M, N, ldn}; p.B = {xa, 0, 1, M, K, ldk};
    p.epi.out = dw; p.epi.ldc = K; p.epi.accumulate = 1; p.split_k = 0;
    NDT1_TRY(gemm_tc_launch(p, s));
  }
  if (dx) {
    NDT1_TRY(k_cast_f32_bf16(w, wa, N, K, K, ldk, s));
    GemmProblem p = plain_problem(GEMM_NN, M, K, N);
    p.A = {ga, 0, 1, M, N, ldn}; p.B = {wa, 0, 1, N, K, ldk};
    p.epi.out = dx; p.epi.ldc = K;
    NDT1_TRY(gemm_tc_launch(p, s));
  }
  return 0;
}


int ndt1_splice_rows(const void* a, const void* ins, const int64_t* split, void* out, int B, int La, int Ls, int W, int elem_size,
                     int use_fill, double fill, void* stream) {
  NDT1_REQUIRE(a && split && out && (ins || use_fill), "splice_rows: null argument");
  NDT1_REQUIRE(elem_size == 4 || elem_size == 8, "splice_rows: element size %d (4 = float32, 8 = int64)", elem_size);
  const long long total = (long long)B * (La + Ls) * W;
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (elem_size == 4)
    ndt1_launch(splice_rows_kernel<float>, blocks, 256, 0, (cudaStream_t)stream, (const float*)a, (const float*)ins, (const long long*)split, (float*)out,
                B, La, Ls, W, use_fill, (float)fill);
  else
    ndt1_launch(splice_rows_kernel<long long>, blocks, 256, 0, (cudaStream_t)stream, (const long long*)a, (const long long*)ins, (const long long*)split,
                (long long*)out, B, La, Ls, W, use_fill, (long long)fill);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int ndt1_unsplice_rows(const float* dout, const int64_t* split, float* da, float* dins, int B, int La, int Ls, int W, void* stream) {
  NDT1_REQUIRE(dout && split, "unsplice_rows: null argument");
  const long long total = (long long)B * (La + Ls) * W;
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  ndt1_launch(unsplice_rows_kernel<float>, blocks, 256, 0, (cudaStream_t)stream, dout, (const long long*)split, da, dins, B, La, Ls, W);
  NDT1_CHECK_LAUNCH();
  return 0;
}
int ndt1_stack_valid(const int64_t* mask, int64_t* out, int B, int T, int stacking, void* stream) {
  NDT1_REQUIRE(mask && out && stacking >= 1, "stack_valid: bad argument");
  const int Ts = (T + stacking - 1) / stacking;
  if (B * Ts == 0) return 0;
  ndt1_launch(stack_valid_kernel, ndt1_cdiv((long long)B * Ts, 256), 256, 0, (cudaStream_t)stream, (const long long*)mask, (long long*)out, B, T, stacking, Ts);
  NDT1_CHECK_LAUNCH();
  return 0;
}

size_t ndt1_attention_workspace_bytes(int B, int L, int n_heads) {
  // delta (B, heads, L) floats | probability keep bits (B, heads, L, 8) u32 | output keep bits (B * L, 4 * heads) u32
  const size_t nrow = ((size_t)B * n_heads * L + 3) / 4 * 4;
  return nrow * 4 + (size_t)B * n_heads * L * 8 * 4 + (size_t)B * L * n_heads * 4 * 4;
}

int ndt1_attention_bf16(const void* qkv, void* out, void* out_drop, float* lse, const int64_t* key_valid, int B, int L, int H, int n_heads,
                        int context_forward, int context_backward, float p_attn, float p_out, uint64_t seed, uint64_t site_attn,
                        uint64_t site_out, const void* dout, void* dqkv, float* delta_ws, int use_tensor_cores, void* stream) {
  AttnParams ap;
  ap.qkv = qkv; ap.out = out; ap.out_drop = out_drop ? out_drop : out; ap.lse = lse; ap.key_valid = (const long long*)key_valid;
  ap.B = B; ap.L = L; ap.H = H; ap.nh = n_heads; ap.hd = H / n_heads;
  const int unb = 1 << 29;
  if (context_forward == -2 && context_backward == -2) { ap.ctx_fwd = unb; ap.ctx_bwd = unb; }
  else { ap.ctx_fwd = context_forward >= -1 ? context_forward : unb; ap.ctx_bwd = context_backward >= -1 ? context_backward : unb; }
  ap.scale = 1.0f / sqrtf((float)ap.hd); ap.p_attn = p_attn; ap.p_out = p_out; ap.seed = seed; ap.stream_attn = site_attn; ap.stream_out = site_out;
  ap.dout = dout; ap.dqkv = dqkv; ap.delta = delta_ws;
  // workspace layout: delta (B, heads, L) floats, then the keep bits (B, heads, L, 8) u32 on a 16-byte boundary
  const long long nrow = ((long long)B * n_heads * L + 3) / 4 * 4;
  ap.drop_bits = delta_ws ? (unsigned int*)(delta_ws + nrow) : nullptr;
  ap.drop_bits_o = delta_ws ? ap.drop_bits + (long long)B * n_heads * L * 8 : nullptr;
  ap.bits_ready = 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool tcp = use_tensor_cores && k_attention_tc_supported(ap);
  NDT1_REQUIRE(!use_tensor_cores || tcp, "attention_bf16: tensor-core path needs head size 128 and at most 256 tokens");
  NDT1_REQUIRE(!(tcp && (p_attn > 0.f || p_out > 0.f)) || delta_ws, "attention_bf16: the tensor-core path with dropout needs the workspace");
  if (tcp) NDT1_TRY(k_attention_tc_fwd(ap, s)); else NDT1_TRY(k_attention_fwd<bf16>(ap, s));
  if (dout) {
    NDT1_REQUIRE(dqkv && delta_ws, "attention_bf16: backward needs dqkv and delta_ws");
    if (tcp) NDT1_TRY(k_attention_tc_bwd(ap, s)); else NDT1_TRY(k_attention_bwd<bf16>(ap, s));
  }
  return 0;
}

int ndt1_attention_f32(const float* qkv, float* out, float* out_drop, float* lse, const int64_t* key_valid, int B, int L, int H, int n_heads,
                       int context_forward, int context_backward, float p_attn, float p_out, uint64_t seed, uint64_t site_attn,
                       uint64_t site_out, const float* dout, float* dqkv, float* delta_ws, void* stream) {
  AttnParams ap;
  ap.qkv = qkv; ap.out = out; ap.out_drop = out_drop ? out_drop : out; ap.lse = lse; ap.key_valid = (const long long*)key_valid;
  ap.B = B; ap.L = L; ap.H = H; ap.nh = n_heads; ap.hd = H / n_heads;
  const int unb = 1 << 29;
  if (context_forward == -2 && context_backward == -2) { ap.ctx_fwd = unb; ap.ctx_bwd = unb; }
  else { ap.ctx_fwd = context_forward >= -1 ? context_forward : unb; ap.ctx_bwd = context_backward >= -1 ? context_backward : unb; }
  ap.scale = 1.0f / sqrtf((float)ap.hd); ap.p_attn = p_attn; ap.p_out = p_out; ap.seed = seed; ap.stream_attn = site_attn; ap.stream_out = site_out;
  ap.dout = dout; ap.dqkv = dqkv; ap.delta = delta_ws;
  ap.drop_bits = nullptr; ap.drop_bits_o = nullptr; ap.bits_ready = 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (!dout) return k_attention_fwd<float>(ap, s);
  NDT1_REQUIRE(dqkv && delta_ws, "attention_f32: backward needs dqkv and delta_ws");
  return k_attention_bwd<float>(ap, s);        // (the backward alone: out / lse are the forward's, nothing is recomputed into them)
}

size_t ndt1_attention_mm_saved_bytes(int B, int L, int H, int n_heads, float p_attn) { return k_attention_mm_saved_bytes(B, L, H, n_heads, p_attn); }
size_t ndt1_attention_mm_workspace_bytes(int B, int L, int H, int n_heads) { return k_attention_mm_workspace_bytes(B, L, H, n_heads); }
int ndt1_attention_mm_fwd(const float* qkv, float* out, void* saved, void* workspace, int B, int L, int H, int n_heads, float p_attn,
                          uint64_t seed, uint64_t site_attn, void* stream) {
  NDT1_REQUIRE(qkv && out && saved && workspace, "attention_mm_fwd: null argument");
  return k_attention_mm_fwd(qkv, out, saved, workspace, B, L, H, n_heads, p_attn, SeedRef(seed), site_attn, (cudaStream_t)stream);
}
int ndt1_attention_mm_bwd(const float* dout, void* saved, void* workspace, float* dqkv, int B, int L, int H, int n_heads, float p_attn,
                          uint64_t seed, uint64_t site_attn, void* stream) {
  NDT1_REQUIRE(dout && dqkv && saved && workspace, "attention_mm_bwd: null argument");
  return k_attention_mm_bwd(dout, saved, workspace, dqkv, B, L, H, n_heads, p_attn, SeedRef(seed), site_attn, (cudaStream_t)stream);
}

int ndt1_xent_loss(const float* logits, const int64_t* labels, float* dlogits, float* loss, int B, int V, const float* dloss, void* stream) {
  NDT1_REQUIRE(logits && labels && V >= 1, "xent_loss: null argument");
  if (B == 0) return 0;
  ndt1_launch(xent_kernel, (B + 7) / 8, 256, 0, (cudaStream_t)stream, logits, (const long long*)labels, dlogits, loss, B, V, dloss);
  NDT1_CHECK_LAUNCH();
  return 0;
}

int ndt1_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx, float* dgamma,
                       float* dbeta, int64_t rows, int H, void* stream) {
  NDT1_REQUIRE(dy && x && gamma && mean && rstd && dx, "layernorm_bwd: null argument");
  return k_layernorm_bwd<float>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, nullptr, 0.f, SeedRef(), 0ull, rows, H, nullptr, (cudaStream_t)stream);
}

int ndt1_dropout_inplace(float* x, int64_t n, float p, uint64_t seed, uint64_t site, void* stream) {
  NDT1_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0,1)");
  if (n == 0 || p == 0.f) return 0;
  return k_dropout_inplace<float>(x, n, p, SeedRef(seed), site, (cudaStream_t)stream);
}

int ndt1_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, void* stream) {
  return k_adamw(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream);
}

int ndt1_adamw_step_fused(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, float grad_scale, void* shadow_bf16, int zero_grad, void* stream) {
  return k_adamw_fused(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (bf16*)shadow_bf16,
                       zero_grad, (cudaStream_t)stream);
}

namespace {
__global__ void dropout_scales_kernel(float* out, long long n, float p, unsigned long long seed, unsigned long long site) { pdl_grid_sync();
  const uint32_t thr = drop_threshold(p);
  const float ik = 1.0f / (1.0f - p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = drop_scale_1(seed, site, (unsigned long long)i, thr, ik);
}
}  // namespace

int ndt1_profile_gemm_begin(void) { return ndt1_profile_begin(); }
int ndt1_profile_gemm_end(double* flops, double* ms, int64_t* launches) {   // the tensor-core GEMM rows of the generic profile
  NDT1_REQUIRE(flops && ms && launches, "profile_gemm_end: null argument");
  static ndt1_profile_entry ent[256];
  int n = 0;
  NDT1_TRY(ndt1_profile_end(ent, 256, &n));
  *flops = 0; *ms = 0; *launches = 0;
  for (int i = 0; i < n; ++i)
    if (strncmp(ent[i].name, "gemm_tc_kernel", 14) == 0 || strstr(ent[i].name, "::gemm_tc_kernel")) {
      *flops += ent[i].flops; *ms += ent[i].ms; *launches += ent[i].launches;
    }
  return 0;
}
int64_t ndt1_launch_counter(void) { return g_ndt1_launches; }

// events that cross the boundary of a captured step (see include/ndt1_b200.h)
int ndt1_event_create(void** event) {
  NDT1_REQUIRE(event, "event_create: null argument");
  cudaEvent_t e;
  NDT1_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event = (void*)e;
  return 0;
}
int ndt1_event_destroy(void* event) {
  if (event) NDT1_CUDA_CHECK(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
int ndt1_event_record(void* event, void* stream) {
  NDT1_REQUIRE(event, "event_record: null event");
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  NDT1_CUDA_CHECK(cudaStreamIsCapturing((cudaStream_t)stream, &st));
  if (st == cudaStreamCaptureStatusActive) NDT1_CUDA_CHECK(cudaEventRecordWithFlags((cudaEvent_t)event, (cudaStream_t)stream, cudaEventRecordExternal));
  else NDT1_CUDA_CHECK(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return 0;
}
int ndt1_stream_wait_event(void* stream, void* event) {
  NDT1_REQUIRE(event, "stream_wait_event: null event");
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  NDT1_CUDA_CHECK(cudaStreamIsCapturing((cudaStream_t)stream, &st));
  NDT1_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, st == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : 0));
  return 0;
}
int ndt1_debug_gemm_timeline(uint64_t* buf) { gemm_tc_set_timeline((unsigned long long*)buf); return 0; }
int ndt1_debug_attention_timeline(uint64_t* buf) { k_attention_tc_set_timeline((unsigned long long*)buf); return 0; }

int ndt1_dropout_scales(float* out, int64_t n, float p, uint64_t seed, uint64_t site, void* stream) {
  if (n == 0) return 0;
  NDT1_REQUIRE(p >= 0.f && p < 1.f, "dropout_scales: p must be in [0,1)");
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  ndt1_launch(dropout_scales_kernel, blocks, 256, 0, (cudaStream_t)stream, out, n, p, seed, site);
  NDT1_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
