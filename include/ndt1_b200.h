/* ndt1_b200.h -- C ABI of the B200-native NDT1 hot path.
 *
 * Drop-in boundary for the data-parallel hot path of colehurwitz/llm_bci:
 * the NDT1 neural-encoder forward and backward over binned spike trains.
 * The reference implements this path as PyTorch module calls; each entry
 * point below cites the reference code it replaces (paths relative to the
 * reference checkout).  The reference-side binding is a ctypes stub, see
 * INTEGRATION.md; llm_bci_b200/_C.py is that stub in this repository.
 *
 * Conventions
 *   - every function returns 0 on success; otherwise ndt1_last_error() holds
 *     a message for the calling thread (the Python shim raises RuntimeError);
 *   - all pointers are DEVICE pointers unless stated otherwise; tensors are
 *     contiguous row-major with the shapes given; integer tensors are int64
 *     exactly as the reference's collate produces them;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the
 *     device; no global state besides the engine object;
 *   - built for sm_100a only; there is no CPU fallback.
 */
#ifndef NDT1_B200_H
#define NDT1_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDT1_ABI_VERSION 3

/* activations (transformers ACT2FN names used by configs/ndt1.yaml) */
enum { NDT1_ACT_IDENTITY = 0, NDT1_ACT_SOFTSIGN = 1, NDT1_ACT_GELU = 2, NDT1_ACT_RELU = 3 };
/* training method, models/ndt1.py:480-491 */
enum { NDT1_METHOD_CTC = 0, NDT1_METHOD_MLM = 1, NDT1_METHOD_AUTOREGRESSIVE = 2 };
/* loss, models/ndt1.py:507-517 */
enum { NDT1_LOSS_CTC = 0, NDT1_LOSS_POISSON_LOG = 1, NDT1_LOSS_POISSON_RATE = 2, NDT1_LOSS_MSE = 3 };
/* arithmetic */
enum { NDT1_PRECISION_FP32 = 0, NDT1_PRECISION_BF16 = 1 };
/* masker modes, models/masker.py:54-83 */
enum { NDT1_MASK_TEMPORAL = 0, NDT1_MASK_NEURON = 1, NDT1_MASK_RANDOM = 2, NDT1_MASK_COSMOOTH = 3 };

const char* ndt1_last_error(void);
int ndt1_abi_version(void);

/* ------------------------------------------------------------------------
 * Stand-alone operators (each is one or two kernel launches)
 * ---------------------------------------------------------------------- */

/* SmoothAndNoise.forward, models/ndt1.py:92-107.
 * out[b,t,n] = sum_i taps[i]*x[b,t+i-(K-1)/2,n] + white_sd*W[b,t,n] + offset_sd*O[b,n]
 * taps: HOST pointer, K odd (0 = no smoothing).  white/offset: injected N(0,1)
 * draws (device) or NULL; with NULL and use_philox != 0 the kernel draws its own (key = *seed_ptr when seed_ptr is
 * non-NULL, read on the device, else seed).  out (fp32) and / or out_bf16 (row stride ld_bf16 elements): the bf16 copy is
 * the operand of the channel-embedding GEMM, handed to the engine as ndt1_batch.spikes_bf16 so that no separate cast runs. */
int ndt1_smooth_noise(const float* x, float* out, int B, int T, int N, const float* taps_host, int K, float white_sd,
                      float offset_sd, const float* white, const float* offset, int use_philox, uint64_t seed,
                      const uint64_t* seed_ptr, void* out_bf16, int ld_bf16, void* stream);

/* Masker.forward given its random draws, models/masker.py:44-104, in place on
 * spikes (B,T,N).  mask_draw has the mode's own shape ((B,T) temporal, (B,N)
 * neuron/region, (B,T,N) random, (N) co-smooth); zero_draw/random_draw (B,T,N)
 * are the Bernoulli(zero_ratio)/(random_ratio) draws, all as uint8 0/1; rand
 * (B,T,N) uniform [0,1).  mask_out (B,T,N) int64 receives the mask; targets_mask
 * (B,T,N) int64, if given, is OR-ed (models/ndt1.py:424-427).  scratch: 4 bytes. */
int ndt1_masker_apply(float* spikes, int B, int T, int N, int mode, int timespan, const uint8_t* mask_draw,
                      const uint8_t* zero_draw, const uint8_t* random_draw, const float* rand, int64_t* mask_out,
                      int64_t* targets_mask, void* scratch, void* stream);
/* device-side draws for the masker's fast path (own Philox streams) */
int ndt1_bernoulli_u8(uint8_t* out, int64_t n, float prob, uint64_t seed, uint64_t stream_id, void* stream);
int ndt1_uniform_f32(float* out, int64_t n, uint64_t seed, uint64_t stream_id, void* stream);

/* padded_array, data_utils/datasets.py:191-221, on the device: ragged rows
 * stored back to back (row b = src[offsets[b]..offsets[b+1]) in units of
 * `inner` elements) -> dst (B, P, inner).  Rows shorter than `pad_to` are padded
 * with `value` on the left or right up to `pad_to`; the first P entries are kept.
 * The reference uses pad_to = P = min(truncate, max(max_len, min_length)).
 * elem_size 4 (float32) or 8 (int64). */
int ndt1_pad_pack(const void* src, const int64_t* offsets, void* dst, int B, int P, int inner, int elem_size, int side_left,
                  int pad_to, double value, void* stream);

/* nn.LogSoftmax + nn.CTCLoss(reduction="none").sum(), models/ndt1.py:499,517,581.
 * logits (B,L,V) -> logp (B,L,V); per-trial nll (B); *loss += sum; optional
 * dlogits (B,L,V) = d loss / d logits scaled by *dloss (NULL = 1).
 * workspace: ndt1_ctc_workspace_bytes(B,L,S) bytes. */
size_t ndt1_ctc_workspace_bytes(int B, int L, int S);
int ndt1_ctc_loss(const float* logits, float* logp, const int64_t* targets, const int64_t* input_lengths,
                  const int64_t* target_lengths, int B, int L, int V, int S, int blank, int zero_infinity, void* workspace,
                  float* nll, float* loss, float* dlogits, const float* dloss, void* stream);
/* argmax + format_ctc, main.py:69 and utils/eval_bci.py:41-48.  out_ids (B,L) padded with -1, out_len (B). */
int ndt1_ctc_greedy_decode(const float* logp, int B, int L, int V, int blank, int64_t* out_ids, int64_t* out_len, void* stream);

/* editdistance.eval(pred, target) per trial (word_edit_distance, utils/eval_bci.py:11-14; the CER metric of main.py:67-73 is
 * sum(errors) / sum(max(target_len, 1))): Levenshtein distance between the first pred_len[b] ids of pred_ids (B,Lp) -- the
 * output of ndt1_ctc_greedy_decode -- and the first target_len[b] ids of target_ids (B,Lt).  errors (B). */
int ndt1_edit_distance(const int64_t* pred_ids, const int64_t* pred_len, int Lp, const int64_t* target_ids, const int64_t* target_len, int Lt,
                       int B, int64_t* errors, void* stream);

/* masked Poisson-NLL / MSE, models/ndt1.py:508-515,548-578.  *loss += sum, *count += weights. */
int ndt1_recon_loss(const float* pred, const float* target, float* dpred, const int64_t* targets_mask, const int64_t* pad_mask,
                    int B, int T, int N, int loss_kind, int shift_by_one, int relu_out, float* loss, int64_t* count,
                    const float* dloss, void* stream);

/* nn.CrossEntropyLoss(reduction="none")(logits, labels).sum(), models/itransformer.py:300-301,375 (the `stat_behaviour` method):
 * logits (B,V), labels (B) int64;  *loss += sum;  optional dlogits (B,V) = (softmax - onehot) * *dloss (NULL = 1). */
int ndt1_xent_loss(const float* logits, const int64_t* labels, float* dlogits, float* loss, int B, int V, const float* dloss, void* stream);

/* nn.LayerNorm (eps 1e-5), models/ndt1.py:309-311,402.  fp32 in/out. */
int ndt1_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int64_t rows,
                       int H, float eps, void* stream);

/* The unmasked attention of nn.TransformerEncoderLayer (models/itransformer.py:157-173) for sequences beyond the fused kernels'
 * 256 tokens, on the tcgen05 GEMM: batched S = Q K^T, row softmax (+ dropout on the probabilities, the same Philox stream as
 * ndt1_attention_f32), O = P V, and the five products of the backward; the L x L probabilities are kept in bf16.
 * qkv (B*L, 3H) fp32 packed q | k | v -> out (B*L, H) fp32;  dout -> dqkv (B*L, 3H).  Head size a multiple of 8.
 * `saved` (ndt1_attention_mm_saved_bytes) carries the forward's operands and probabilities to the backward;
 * `workspace` (ndt1_attention_mm_workspace_bytes) is scratch of one call.  Both 256-byte aligned. */
size_t ndt1_attention_mm_saved_bytes(int B, int L, int H, int n_heads, float p_attn);
size_t ndt1_attention_mm_workspace_bytes(int B, int L, int H, int n_heads);
int ndt1_attention_mm_fwd(const float* qkv, float* out, void* saved, void* workspace, int B, int L, int H, int n_heads, float p_attn,
                          uint64_t seed, uint64_t site_attn, void* stream);
int ndt1_attention_mm_bwd(const float* dout, void* saved, void* workspace, float* dqkv, int B, int L, int H, int n_heads, float p_attn,
                          uint64_t seed, uint64_t site_attn, void* stream);

/* Backward of ndt1_layernorm_fwd (nn.LayerNorm under autograd; the post-LN layers of models/itransformer.py:157-173 and the
 * LayerNorms of its embedders :108-150):  dx (rows,H) += dLN/dx(dy),  dgamma[H] += sum_r dy * xhat,  dbeta[H] += sum_r dy.
 * All three accumulate: zero-fill them for plain gradients, or pre-load dx with the gradient of a residual branch. */
int ndt1_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx, float* dgamma,
                       float* dbeta, int64_t rows, int H, void* stream);

/* nn.Dropout in place with the library's Philox streams: x[i] *= keep(seed, site, i) / (1 - p).  The same call on the gradient
 * is its backward. */
int ndt1_dropout_inplace(float* x, int64_t n, float p, uint64_t seed, uint64_t site, void* stream);

/* y = act(x W^T + b) in fp32 (CUDA cores) or bf16 tensor cores (tcgen05); x (M,K), W (N,K).
 * Replaces the nn.Linear calls of models/ndt1.py:130,140,219-221,247-257,494. */
int ndt1_linear_fwd(const float* x, const float* w, const float* bias, float* y, float* pre /* optional: x W^T + b before act */,
                    int M, int N, int K, int act, int precision, void* workspace, size_t workspace_bytes, void* stream);
/* Backward of ndt1_linear_fwd (the projector MLP of the BCI coupler, models/bci.py:88-96,137-141, trains through it):
 * g = dy * act'(saved) (saved = y for relu / softsign, the pre-activation for gelu);  db[n] += sum_m g;  dw (N,K) += g^T x;
 * dx (M,K) = g W.  Any of dx / dw / db may be NULL.  Workspace: ndt1_linear_workspace_bytes (bf16 mode), M*N*4 + 256 (fp32). */
int ndt1_linear_bwd(const float* dy, const float* x, const float* w, const float* saved, float* dx, float* dw, float* db, int M, int N,
                    int K, int act, int precision, void* workspace, size_t workspace_bytes, void* stream);
size_t ndt1_linear_workspace_bytes(int M, int N, int K);
/* The same pair with nn.Dropout(p) applied to the OUTPUT inside the GEMM epilogue (y = dropout(act(x W^T + b)), the library's
 * Philox stream (seed, site) over the flat (row, column) index -- the mask ndt1_dropout_inplace(y, ...) would apply), and, in
 * the backward, to dy before the activation derivative.  The Linear -> activation -> Dropout runs of
 * nn.TransformerEncoderLayer and torchvision's MLP (models/itransformer.py:108-116, 157-165) in one pass each way.
 * `saved` for relu / softsign is the forward's (dropped) output. */
int ndt1_linear_drop_fwd(const float* x, const float* w, const float* bias, float* y, float* pre, int M, int N, int K, int act, int precision,
                         void* workspace, size_t workspace_bytes, float drop_p, uint64_t seed, uint64_t site, void* stream);
int ndt1_linear_drop_bwd(const float* dy, const float* x, const float* w, const float* saved, float* dx, float* dw, float* db, int M, int N,
                         int K, int act, int precision, void* workspace, size_t workspace_bytes, float drop_p, uint64_t seed, uint64_t site,
                         void* stream);

/* BCI.prepare_embeds, models/bci.py:143-166: out (B, La+Ls, W) = a[b, :split[b]] | ins[b] | a[b, split[b]:] per trial (or a
 * constant row `fill` instead of ins: the -100 targets over the spike positions).  elem_size 4 (float32) or 8 (int64).
 * ndt1_unsplice_rows is its backward for float32: the gradient of `out` routed back to a and ins.
 * ndt1_stack_valid, models/bci.py:127-141: out (B, ceil(T / stacking)) = 1 where all `stacking` rows of the group are valid. */
int ndt1_splice_rows(const void* a, const void* ins, const int64_t* split, void* out, int B, int La, int Ls, int W, int elem_size,
                     int use_fill, double fill, void* stream);
int ndt1_unsplice_rows(const float* dout, const int64_t* split, float* da, float* dins, int B, int La, int Ls, int W, void* stream);
int ndt1_stack_valid(const int64_t* mask, int64_t* out, int B, int T, int stacking, void* stream);

/* F.scaled_dot_product_attention with the reference's mask (models/ndt1.py:276-290, 30-41, 435-437),
 * bf16 in/out: qkv (B*L, 3H) packed q|k|v, out/out_drop (B*L, H) before/after the output dropout,
 * lse (B, heads, L).  With dout (gradient w.r.t. out) the backward runs too: dqkv (B*L, 3H).
 * delta_ws: scratch of ndt1_attention_workspace_bytes(B, L, heads) bytes, 16-byte aligned (row sums for
 * the backward + the keep bits of the probability dropout; required for the backward and for the
 * tensor-core forward with p_attn > 0).  use_tensor_cores selects the tcgen05 kernels (head size 128,
 * L <= 256) or the CUDA-core kernels. */
size_t ndt1_attention_workspace_bytes(int B, int L, int n_heads);
int ndt1_attention_bf16(const void* qkv, void* out, void* out_drop, float* lse, const int64_t* key_valid, int B, int L, int H, int n_heads,
                        int context_forward, int context_backward, float p_attn, float p_out, uint64_t seed, uint64_t site_attn,
                        uint64_t site_out, const void* dout, void* dqkv, float* delta_ws, int use_tensor_cores, void* stream);
/* The same operator in fp32 on CUDA cores (head sizes 16 / 32 / 64 / 96 / 128, any length; context -2 / -2 = no band, i.e. the
 * unmasked attention of nn.TransformerEncoderLayer, models/itransformer.py:157-173).  Without dout: the forward (out, out_drop,
 * lse).  With dout: the backward ALONE (dqkv from the forward's qkv / out / lse; delta_ws: B * n_heads * L floats). */
int ndt1_attention_f32(const float* qkv, float* out, float* out_drop, float* lse, const int64_t* key_valid, int B, int L, int H, int n_heads,
                       int context_forward, int context_backward, float p_attn, float p_out, uint64_t seed, uint64_t site_attn,
                       uint64_t site_out, const float* dout, float* dqkv, float* delta_ws, void* stream);


/* torch.optim.AdamW step on one flat buffer, models/trainer.py:229,340. */
int ndt1_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* The same update in one pass that also (optionally) writes the bf16 shadow of the parameters (see
 * ndt1_engine_set_weight_shadow) and clears the gradient for the next step (optimizer.zero_grad, trainer.py:343). */
int ndt1_adamw_step_fused(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int step, float grad_scale, void* shadow_bf16, int zero_grad,
                          void* stream);

/* ------------------------------------------------------------------------
 * The encoder + head engine: everything of NDT1.forward after the masker
 * (models/ndt1.py:429-450, 542-589) and its backward, as one call each.
 * ---------------------------------------------------------------------- */

typedef struct ndt1_engine ndt1_engine;

typedef struct {
  int32_t abi_version;       /* NDT1_ABI_VERSION */
  int32_t precision;         /* NDT1_PRECISION_* */
  /* embedder, configs/ndt1.yaml:34-53 */
  int32_t n_channels, input_dim, max_F;
  int32_t embed_bias, embed_act, pos;
  int32_t stack_active, stack_size, stack_stride;
  int32_t block_token, day_token, n_blocks, n_days, adapt;
  /* transformer, configs/ndt1.yaml:56-71 */
  int32_t n_layers, hidden, n_heads, inter, attention_bias, mlp_bias, mlp_act;
  int32_t use_rope;
  float rope_theta;
  /* context, configs/ndt1.yaml:21-24 */
  int32_t context_forward, context_backward;
  /* factors, configs/ndt1.yaml:74-81 */
  int32_t factors_active, factors_size, factors_act, factors_bias;
  /* head + loss */
  int32_t method, loss_kind, n_outputs, blank_id, zero_infinity, decoder_relu;
  /* dropout probabilities (0 disables a site) */
  float p_embed, p_transformer, p_factors;
  /* capacity of the activation arena */
  int32_t max_batch, max_T, max_targets;
} ndt1_config;

/* Parameter / gradient pointer tables, in the order of the reference's
 * state_dict; absent tensors are NULL. */
#define NDT1_MAX_LAYERS 32
#define NDT1_MAX_DAYS 64
typedef struct {
  float* embed_w; float* embed_b;      /* embed_spikes (D, N), (D); unused (NULL) with adapt */
  float* proj_w;  float* proj_b;       /* stack_projection (H, S*D) or projection (H, D) */
  float* pos_w;                        /* embed_pos (max_F, H) */
  float* block_emb; float* day_emb;    /* (n_blocks, H), (n_days, H) */
  /* adapt (models/ndt1.py:118-127,170-171): embed_spikes.{d}.weight (D, N) / .bias (D) of day d < n_days <= NDT1_MAX_DAYS;
   * trial b uses day_idx[b].  Anywhere in memory: the engine packs them per forward and scatters the gradients back. */
  float* embed_w_day[NDT1_MAX_DAYS]; float* embed_b_day[NDT1_MAX_DAYS];
  struct {
    float *ln1_w, *ln1_b, *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b, *ln2_w, *ln2_b, *up_w, *up_b, *down_w, *down_b;
  } layer[NDT1_MAX_LAYERS];
  float* out_norm_w; float* out_norm_b;
  float* factors_w; float* factors_b;  /* out_proj.proj.0 */
  float* dec_w; float* dec_b;          /* decoder.0 (n_outputs, H or factors_size) */
} ndt1_tensors;

typedef struct {
  const float* spikes;            /* (B,T,N) after smoothing/noise/masking */
  const int64_t* spikes_mask;     /* (B,T) 1 = valid */
  const int64_t* spikes_timestamp;/* (B,T) */
  const int64_t* spikes_lengths;  /* (B) */
  const int64_t* block_idx;       /* (B) or NULL */
  const int64_t* day_idx;         /* (B) or NULL */
  const int64_t* targets;         /* ctc: (B,S) */
  const int64_t* targets_lengths; /* ctc: (B) */
  const float* recon_targets;     /* mlm/autoregressive: (B,T,N) original spikes */
  const int64_t* targets_mask;    /* mlm: (B,T,N) from the masker */
  int32_t B, T, S;
  int32_t training;               /* dropout on */
  int32_t need_backward;          /* keep activations and loss gradients for ndt1_engine_backward */
  int32_t encoder_only;           /* stop after the encoder (NeuralEncoder.forward, models/ndt1.py:408-450): no head, no loss */
  uint64_t seed;                  /* Philox key of this step's dropout */
  const void* spikes_bf16;        /* optional (bf16 mode): the input already as bf16 (B, T, n_channels), n_channels % 8 == 0, e.g. written by
                                   * ndt1_smooth_noise; the engine then skips its own cast of `spikes`.  Must stay valid until the backward. */
  const uint64_t* seed_ptr;       /* optional: the key in DEVICE memory instead (read when the step executes, so a captured CUDA graph of
                                   * the step is replayed with a new key by updating that word); NULL = use `seed` */
} ndt1_batch;

typedef struct {
  float* loss;                    /* (1) sum loss */
  int64_t* n_examples;            /* (1) */
  float* preds;                   /* ctc: (B,L',V) log-probs; ssl: (B,T,N) (log-)rates */
  int64_t* out_mask;              /* (B,L) stacked padding mask incl. prefix tokens, or NULL */
  int64_t* loss_mask;             /* mlm: (B,T,N) targets_mask & padding, or NULL */
  int64_t* out_lengths;           /* (B) get_stacked_lens(spikes_lengths), or NULL */
  float* features;                /* (B,L',H_out) encoder output, or NULL */
} ndt1_outputs;

int ndt1_engine_create(const ndt1_config* cfg, ndt1_engine** out);
void ndt1_engine_destroy(ndt1_engine* e);
size_t ndt1_engine_arena_bytes(const ndt1_engine* e);
/* output sequence length for T input bins: 1 + (T - size) / stride when stacking (models/ndt1.py:138-140) */
int ndt1_engine_out_len(const ndt1_engine* e, int T);
/* forward; activations stay in the engine for the matching backward */
int ndt1_engine_forward(ndt1_engine* e, const ndt1_tensors* params, const ndt1_batch* batch, const ndt1_outputs* out, void* stream);
/* backward of the last forward: grads->X += dloss * dLoss/dX (buffers are caller-zeroed) */
int ndt1_engine_backward(ndt1_engine* e, const ndt1_tensors* params, const ndt1_tensors* grads, const float* dloss, void* stream);
/* backward of an encoder-only forward (batch.encoder_only = 1, need_backward = 1): dfeatures (B,L',H_out) is the caller's
 * gradient w.r.t. outputs.features -- what autograd hands back when the encoder feeds another model
 * (NeuralEncoder.forward used as a sub-module, models/bci.py:125).  Same accumulation rule as ndt1_engine_backward. */
int ndt1_engine_backward_features(ndt1_engine* e, const ndt1_tensors* params, const ndt1_tensors* grads, const float* dfeatures, void* stream);
/* The backward runs its weight-gradient GEMMs on an engine-owned second stream, concurrently with the data-gradient
 * chain (joined before ndt1_engine_backward's work on `stream` ends).  on = 0 serialises everything on `stream`
 * (used to time single kernels); default 1, or 0 when NDT1_OVERLAP=0 is set at engine creation. */
int ndt1_engine_set_overlap(ndt1_engine* e, int on);
/* bf16 mode: the caller keeps a bf16 copy (`shadow_bf16`, n elements, same offsets) of the flat fp32 arena its
 * parameters live in (`params_fp32`) -- ndt1_adamw_step_fused writes it -- so the forward reads the weights from there
 * instead of casting them every step.  Weights outside the arena, or a NULL shadow, use the per-forward cast.  The caller
 * guarantees the shadow matches the parameters whenever ndt1_engine_forward runs. */
int ndt1_engine_set_weight_shadow(ndt1_engine* e, const float* params_fp32, const void* shadow_bf16, int64_t n);
/* Gradient stages of the last backward in completion order: 0 = decoder + out_norm,
 * 1..n_layers = layers n_layers-1..0, n_layers+1 = embedder.  ndt1_engine_wait_stage makes
 * `stream` wait (cudaStreamWaitEvent) until that stage's gradients are final, so a bucketed
 * all-reduce on a side stream overlaps the rest of the backward (DDP semantics,
 * models/trainer.py:258-262,339). */
int ndt1_engine_stage_count(const ndt1_engine* e);
int ndt1_engine_wait_stage(ndt1_engine* e, int stage, void* stream);
/* use_rope (models/ndt1.py:44-71, 262-266): the engine builds the (max_F, head_size) cos / sin tables itself at creation
 * (get_cos_sin in double precision, rounded to fp32).  A caller that wants the exact table values of another
 * implementation (the Python host passes the ones torch computes, like the reference) overrides them here; device
 * pointers, `rows` >= max_F rows of head_size floats, copied. */
int ndt1_engine_set_rope_tables(ndt1_engine* e, const float* cos_table, const float* sin_table, int rows);
/* number of kernels launched by the last forward+backward (bench.py gpu_launches) */
int64_t ndt1_engine_launch_count(const ndt1_engine* e);
/* Measurement hooks (bench.py): between begin and end every tensor-core GEMM launch is bracketed by
 * CUDA events on its own stream; end returns the summed algorithmic FLOPs, device milliseconds and
 * the number of launches.  ndt1_launch_counter: kernels launched by this library on the calling thread. */
int ndt1_profile_gemm_begin(void);
int ndt1_profile_gemm_end(double* flops, double* ms, int64_t* launches);
int64_t ndt1_launch_counter(void);
/* The same for EVERY kernel of the library, grouped by kernel (template instantiations separately): between begin and end
 * each launch is bracketed by CUDA events on the stream it is launched on.  `flops` / `bytes` are the ALGORITHMIC figures the
 * launchers attach (0 where a kernel has none): GEMMs 2 M N K; attention 2 (forward) + 2 + 2 (backward) contractions of
 * 2 L^2 d per head (recomputed scores are not counted); LayerNorm, AdamW, smoothing, CTC, reductions their one pass over
 * the operands.  ndt1_profile_end writes at most `capacity` entries and their number to *n_out. */
typedef struct ndt1_profile_entry {
  char name[160];
  int64_t launches;
  double ms, flops, bytes;
} ndt1_profile_entry;
int ndt1_profile_begin(void);
int ndt1_profile_end(ndt1_profile_entry* out, int capacity, int* n_out);
/* Events that cross the boundary of a CAPTURED step.  A training step (ndt1_smooth_noise + ndt1_engine_forward +
 * ndt1_engine_backward) only enqueues work on the caller's stream and allocates nothing, so the caller may capture it into a
 * CUDA graph (the per-step Philox key is then passed by device pointer: ndt1_batch.seed_ptr).  Work that stays OUTSIDE the graph
 * -- the gradient exchange and the optimizer on their own stream -- is ordered against it with these events: when `stream` is
 * being captured, ndt1_event_record / ndt1_stream_wait_event add EXTERNAL event nodes (cudaEventRecordExternal /
 * cudaEventWaitExternal), which record / wait on the real event at every replay; otherwise they are plain cudaEventRecord /
 * cudaStreamWaitEvent.  The engine records its gradient-stage events (ndt1_engine_wait_stage) the same way. */
int ndt1_event_create(void** event);
int ndt1_event_destroy(void* event);
int ndt1_event_record(void* event, void* stream);
int ndt1_stream_wait_event(void* stream, void* event);
/* Debugging aid (tools/attn_timeline.py): the tensor-core attention kernels write per-CTA phase timestamps
 * (32 uint64 per CTA, %globaltimer ns) into buf; NULL switches it off. */
int ndt1_debug_attention_timeline(uint64_t* buf);
/* The same for the tensor-core GEMM (tools/gemm_timeline.py): 16 uint64 per CTA of every launch while set (each launch
 * overwrites the previous one's slots). */
int ndt1_debug_gemm_timeline(uint64_t* buf);
/* keep-scale (0 or 1/(1-p)) of a dropout site, for tests: site 0 embed, 1+4*l attn-probs, 2+4*l attn-out, 3+4*l mlp */
int ndt1_dropout_scales(float* out, int64_t n, float p, uint64_t seed, uint64_t site, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NDT1_B200_H */
