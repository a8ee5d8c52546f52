// C-ABI wrappers of the stand-alone operators declared in include/ndt1_b200.h.
#include "../../include/ndt1_b200.h"
#include "kernels.cuh"
#include <string.h>

extern "C" {

int ndt1_smooth_noise(const float* x, float* out, int B, int T, int N, const float* taps_host, int K, float white_sd,
                      float offset_sd, const float* white, const float* offset, int use_philox, uint64_t seed, void* stream) {
  return k_smooth_noise(x, out, B, T, N, taps_host, K, white_sd, offset_sd, white, offset, use_philox, seed, (cudaStream_t)stream);
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

int ndt1_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int N, int K, int act, int precision,
                    void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GemmProblem p;
  p.b_sel = nullptr;
  p.mode = GEMM_NT; p.M = M; p.N = N; p.nb_out = 1; p.nchunk = 1; p.chunk_k = K;
  p.a_row_shift = p.a_col_shift = p.b_row_shift = p.b_col_shift = 0; p.b_chunk_n = N; p.split_k = 1;
  p.epi = gemm_epilogue_default();
  p.epi.out = y; p.epi.ldc = N; p.epi.bias = bias;
  p.epi.act = act == NDT1_ACT_SOFTSIGN ? ACT_SOFTSIGN : act == NDT1_ACT_GELU ? ACT_GELU : act == NDT1_ACT_RELU ? ACT_RELU : ACT_NONE;
  if (precision == NDT1_PRECISION_FP32) {
    p.A = {x, 0, 1, M, K, K}; p.B = {w, 0, 1, N, K, K};
    return gemm_simt_launch(p, 0, s);
  }
  const int ldk = (K + 7) / 8 * 8;
  const size_t need = ((size_t)M * ldk + (size_t)N * ldk) * sizeof(bf16) + 512;
  NDT1_REQUIRE(workspace && workspace_bytes >= need, "linear_fwd: bf16 mode needs %zu workspace bytes", need);
  bf16* xa = (bf16*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  bf16* wa = (bf16*)(((uintptr_t)(xa + (size_t)M * ldk) + 255) & ~(uintptr_t)255);
  NDT1_TRY(k_cast_f32_bf16(x, xa, M, K, K, ldk, s));
  NDT1_TRY(k_cast_f32_bf16(w, wa, N, K, K, ldk, s));
  p.A = {xa, 0, 1, M, K, ldk}; p.B = {wa, 0, 1, N, K, ldk};
  return gemm_tc_launch(p, s);
}

size_t ndt1_attention_workspace_bytes(int B, int L, int n_heads) {
  const size_t nrow = ((size_t)B * n_heads * L + 3) / 4 * 4;
  return nrow * 4 + (size_t)B * n_heads * L * 8 * 4;
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
  cudaStream_t s = (cudaStream_t)stream;
  const bool tcp = use_tensor_cores && k_attention_tc_supported(ap);
  NDT1_REQUIRE(!use_tensor_cores || tcp, "attention_bf16: tensor-core path needs head size 128 and at most 256 tokens");
  NDT1_REQUIRE(!(tcp && p_attn > 0.f) || delta_ws, "attention_bf16: the tensor-core path with dropout needs the workspace");
  if (tcp) NDT1_TRY(k_attention_tc_fwd(ap, s)); else NDT1_TRY(k_attention_fwd<bf16>(ap, s));
  if (dout) {
    NDT1_REQUIRE(dqkv && delta_ws, "attention_bf16: backward needs dqkv and delta_ws");
    if (tcp) NDT1_TRY(k_attention_tc_bwd(ap, s)); else NDT1_TRY(k_attention_bwd<bf16>(ap, s));
  }
  return 0;
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
