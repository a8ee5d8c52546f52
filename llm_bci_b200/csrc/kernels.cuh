// Internal launcher declarations (one per kernel family).  Every launcher
// returns 0 on success and records a message via ndt1_set_error otherwise.
#pragma once
#include "common.cuh"
#include "gemm_common.cuh"

// elementwise.cu
int k_smooth_noise(const float* x, float* out, int B, int T, int N, const float* taps, int K, float white_sd, float offset_sd,
                   const float* white, const float* offset, int use_philox, SeedRef seed, cudaStream_t stream, bf16* out_bf16 = nullptr,
                   int ld_bf16 = 0);
int k_masker_apply(float* spikes, int B, int T, int N, int mode, int timespan, const unsigned char* mask_draw,
                   const unsigned char* zero_draw, const unsigned char* random_draw, const float* rand, long long* mask_out,
                   long long* targets_mask, unsigned int* scratch, cudaStream_t stream);
int k_bernoulli_u8(unsigned char* out, long long n, float prob, unsigned long long seed, unsigned long long stream_id, cudaStream_t stream);
int k_uniform_f32(float* out, long long n, unsigned long long seed, unsigned long long stream_id, cudaStream_t stream);
int k_pad_pack(const void* src, const long long* offsets, void* dst, int B, int P, int inner, int elem_size, int side_left, int full,
               double value, cudaStream_t stream);
int k_cast_f32_bf16(const float* in, bf16* out, long long rows, int cols, long long ld_in, long long ld_out, cudaStream_t stream);
constexpr int CAST_MAX_SEGS = 64;
struct CastSegs {            // fp32 -> bf16 copies done by one launch (k_cast_multi fills first_chunk / total_chunks)
  const float* src[CAST_MAX_SEGS]; bf16* dst[CAST_MAX_SEGS]; long long count[CAST_MAX_SEGS]; long long first_chunk[CAST_MAX_SEGS];
  long long total_chunks; int n;
};
int k_cast_multi(CastSegs& segs, cudaStream_t stream);
template <typename T> int k_colsum(const T* in, float* out, long long rows, int cols, long long ld, cudaStream_t stream);
template <typename T>
int k_grad_prep(const float* g, T* out, long long rows, int cols, float drop_p, SeedRef seed, unsigned long long stream_id,
                float* dtab, const long long* idx, int tab_ld, int rows_per_b, long long idx_stride, int prefix, cudaStream_t stream);
int k_recon_loss(const float* pred, const float* target, float* dpred, const long long* tmask, const long long* pmask, int B, int T,
                 int N, int kind, int shift, int relu_out, float* loss, long long* count, const float* dloss, cudaStream_t stream);
int k_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd, int step,
            float gscale, cudaStream_t stream);
int k_adamw_fused(float* p, float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd, int step,
                  float gscale, bf16* shadow, int zero_grad, cudaStream_t stream);
int k_stack_mask(const long long* mask, long long* out, int B, int T, int Tp, int size, int stride, int n_prefix, cudaStream_t stream);
int k_token_rows(const float* table, const long long* idx, float* x, int B, int L, int H, int slot, cudaStream_t stream);
template <typename T> int k_token_rows_grad(float* dtable, const long long* idx, const T* dx, int B, int L, int H, int slot, cudaStream_t stream);
template <typename T> int k_scale_cast_pad(const float* in, T* out, long long rows, int cols, int ld_out, const float* scale, cudaStream_t stream);
int k_stacked_lens(const long long* lens, long long* out, int B, int stack, int size, int stride, cudaStream_t stream);
int k_set_i64(long long* p, long long v, cudaStream_t stream);
int k_relu_inplace(float* x, long long n, cudaStream_t stream);
// options off in the shipped yaml: RoPE (in place on q|k of the packed qkv rows), dropout in front of the factors projection,
// per-day embedding (bias-gradient column sums routed by day, gradient scatter)
template <typename T>
int k_rope(T* qkv, const long long* ts, long long ts_stride, const float* cs, const float* sn, long long rows, int L, int H, int nh,
           int max_F, int inverse, cudaStream_t stream);
template <typename T>
int k_dropout_inplace(T* x, long long n, float p, SeedRef seed, unsigned long long stream_id, cudaStream_t stream);
template <typename T>
int k_colsum_sel(const T* in, float* out, const long long* sel, int n_sel, long long out_stride, int B, int rows_per_b, int cols,
                 cudaStream_t stream);
int k_add_inplace(float* dst, const float* src, long long n, cudaStream_t stream);
template <typename T> int k_dact_inplace(T* g, const T* saved, long long n, int dact, cudaStream_t stream);
template <typename T> int k_cast_to_f32(const T* in, float* out, long long n, cudaStream_t stream);
int k_and_mask(const long long* tmask, const long long* pmask, long long* out, int B, int T, int N, cudaStream_t stream);

// layernorm.cu
template <typename T>
int k_layernorm_fwd(const float* x, const float* gamma, const float* beta, T* y, float* mean, float* rstd, long long rows, int H,
                    float eps, cudaStream_t stream);
// dres (fp32, in/out) += LN'(dy);  optional out_lp = T(dres_new * dropscale)
template <typename T>
int k_layernorm_bwd(const T* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dres, float* dgamma,
                    float* dbeta, T* out_lp, float drop_p, SeedRef seed, unsigned long long stream_id, long long rows, int H,
                    float* partials, cudaStream_t stream, float* colsum_out = nullptr);   // colsum_out[c] += sum_r out_lp[r, c]
size_t k_layernorm_bwd_partials_bytes(int H);

// attention.cu (CUDA-core path, both precisions)
struct AttnParams {
  const void* qkv;          // (B*L, 3H): q | k | v, head h at columns h*hd
  void* out;                // (B*L, H) attention output (pre output-dropout)
  void* out_drop;           // (B*L, H) after output dropout (== out when p_out == 0)
  float* lse;               // (B, nh, L)
  const long long* key_valid;  // (B, L)
  int B, L, H, nh, hd;
  int ctx_fwd, ctx_bwd;     // effective band half-widths (INT_MAX/2 = unbounded); self may be excluded by -1
  float scale;
  float p_attn, p_out;
  SeedRef seed; unsigned long long stream_attn, stream_out;
  // backward
  const void* dout;         // (B*L, H) gradient w.r.t. out_drop input of out_proj (already through output dropout)
  void* dqkv;               // (B*L, 3H)
  float* delta;             // (B, nh, L) scratch
  // tensor-core path, p_attn > 0: keep bits of the probability dropout, written by the forward, read by the backward
  unsigned int* drop_bits;  // (B, nh, L, 8) u32: bit j%32 of word j/32 = key j kept
  unsigned int* drop_bits_o; // (B*L, H/32) u32: keep bits of the OUTPUT dropout (element (row, col) -> bit col%32 of word col/32), or null
  int bits_ready;           // 1: drop_bits / drop_bits_o were already drawn for this forward (k_attention_tc_dropbits on a side
                            //    stream, off the critical path); 0: the forward launcher draws them first
};
template <typename T> int k_attention_fwd(const AttnParams& p, cudaStream_t stream);
template <typename T> int k_attention_bwd(const AttnParams& p, cudaStream_t stream);
template <typename T> int k_attention_delta(const AttnParams& p, cudaStream_t stream);
// attention_mm.cu: unmasked attention over long sequences as batched tcgen05 GEMMs (probabilities materialised in bf16)
size_t k_attention_mm_saved_bytes(int B, int L, int H, int nh, float p_attn);
size_t k_attention_mm_workspace_bytes(int B, int L, int H, int nh);
int k_attention_mm_fwd(const float* qkv, float* out, void* saved, void* workspace, int B, int L, int H, int nh, float p_attn, SeedRef seed,
                       unsigned long long site, cudaStream_t stream);
int k_attention_mm_bwd(const float* dout, void* saved, void* workspace, float* dqkv, int B, int L, int H, int nh, float p_attn, SeedRef seed,
                       unsigned long long site, cudaStream_t stream);
// attention_tc.cu (tcgen05 path: bf16, head size 128, at most 256 tokens)
bool k_attention_tc_supported(const AttnParams& p);
int k_attention_tc_fwd(const AttnParams& p, cudaStream_t stream);
int k_attention_tc_bwd(const AttnParams& p, cudaStream_t stream);
// keep bits of the probability and output dropout of up to NDT1_MAX_LAYERS attention layers in ONE launch (same Philox draws as
// drop_scale_1 of the CUDA-core kernels): they depend on the step's key only, so the engine draws them on its second stream
// while the embedding GEMMs run, and no attention kernel spends an instruction on the generator
struct AttnBitsJob { unsigned int* bits_p[32]; unsigned int* bits_o[32]; unsigned int* bits_m[32]; unsigned long long stream_p[32], stream_o[32], stream_m[32]; int n; };
// (bits_m: a third site per layer in the layout of bits_o -- the MLP dropout, read by the down-projection GEMM's epilogue)
int k_attention_tc_dropbits(const AttnBitsJob& job, const AttnParams& shape, cudaStream_t stream);
void k_attention_tc_set_timeline(unsigned long long* buf);   // debugging: per-CTA phase timestamps (32 u64 per CTA), null = off

// ctc.cu
int k_log_softmax(const float* logits, float* logp, long long rows, int V, cudaStream_t stream, int ld_in = 0);   // ld_in: row stride of logits (0 = V)
// log-probs (B, L, V); per-trial NLL summed into *loss; dlogits = (softmax - posterior) * (*dloss) for t < len
int k_ctc_fwd_bwd(const float* logp, const long long* targets, const long long* in_len, const long long* tgt_len, int B, int L, int V,
                  int S, int blank, int zero_infinity, float* alpha_ws, float* nll, float* loss, float* dlogits, const float* dloss,
                  cudaStream_t stream, int fast_math = 0);
size_t k_ctc_workspace_floats(int B, int L, int S);
int k_edit_distance(const long long* pred, const long long* pred_len, int Lp, const long long* tgt, const long long* tgt_len, int Lt, int B,
                    long long* out, cudaStream_t stream);
int k_ctc_greedy_decode(const float* logp, int B, int L, int V, int blank, long long* out_ids, long long* out_len, cudaStream_t stream);
