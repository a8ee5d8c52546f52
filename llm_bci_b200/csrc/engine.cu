// The NDT1 encoder + head engine: NeuralEncoder.forward after the masker,
// the decoder and the loss (models/ndt1.py:429-450, 542-589) and the whole
// backward, as two C-ABI calls that only enqueue kernels on the caller's
// stream.  Activations needed by the backward live in one arena sized at
// creation (bump allocated, 256-byte aligned): nothing is allocated, freed or
// synchronised per step, so a step can be captured in a CUDA graph.
//
// Data layout in HBM (T = float in NDT1_PRECISION_FP32, bf16 in _BF16):
//   residual stream   x[2*layers+1]  fp32 (B*L, H)      LayerNorm inputs, kept for backward
//   LN outputs        h1[l], h2[l], hn  T (B*L, H)      GEMM A operands / wgrad B operands
//   qkv[l]            T (B*L, 3H)  q|k|v packed so one GEMM (N = 3H) produces it
//   att[l], attd[l]   T (B*L, H)   attention output before / after output dropout
//   u[l], g[l]        T (B*L, I)   MLP activation derivative GELU'(pre-activation) (the pre-activation itself for other activations) / activation
//   emb               T (B*T, D)   softsign(embed_spikes(x)); the stack projection
//                                  reads it as (B, T/stride, stride*D) shifted windows
//   bf16 mode keeps bf16 copies of the weights (QKV concatenated) refreshed per forward.
#include "../../include/ndt1_b200.h"
#include "kernels.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

constexpr int kUnbounded = 1 << 29;

struct Arena {
  char* base = nullptr; size_t cap = 0, off = 0;
  template <typename U> U* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    U* p = (U*)(base ? base + off : nullptr);
    off += n * sizeof(U);
    return p;
  }
};

template <typename T> struct IsBf16 { static constexpr int v = 0; };
template <> struct IsBf16<bf16> { static constexpr int v = 1; };

}  // namespace

struct ndt1_engine {
  ndt1_config c;
  long long launches = 0;
  cudaEvent_t stage_ev[NDT1_MAX_LAYERS + 2] = {};   // gradient stages of the last backward, in completion order
  int overlap = 1;              // weight gradients on a second stream, concurrent with the data-gradient chain
  const float* shadow_src = nullptr; const void* shadow_bf16 = nullptr; long long shadow_n = 0;   // caller-maintained bf16 copy of its parameter arena
  int n_stages() const { return c.n_layers + 2; }
  virtual ~ndt1_engine() {}
  virtual int forward(const ndt1_tensors* P, const ndt1_batch* b, const ndt1_outputs* o, cudaStream_t s) = 0;
  virtual int backward(const ndt1_tensors* P, const ndt1_tensors* G, const float* dloss, cudaStream_t s, const float* dfeatures = nullptr) = 0;
  virtual size_t arena_bytes() const = 0;
  virtual int set_rope_tables(const float* cs, const float* sn, int rows) = 0;
  int out_len(int T) const { return c.stack_active ? (T - c.stack_size) / c.stack_stride + 1 : T; }
};

namespace {

template <typename T>
struct Engine : ndt1_engine {
  static constexpr int kBf16 = IsBf16<T>::v;
  Arena ar;
  bool force_simt = false;      // NDT1_FORCE_SIMT=1: every GEMM and the attention on the CUDA cores (debug)
  bool simt_attention = false;  // NDT1_SIMT_ATTENTION=1: only the attention on the CUDA cores
  int n_prefix = 0;
  // arena views ------------------------------------------------------------
  T* xin = nullptr; int ldN = 0;
  const T* xin_ext = nullptr;                  // caller-provided bf16 input of the last forward (ndt1_batch.spikes_bf16)
  T* emb = nullptr; T* emb_pre = nullptr;      // emb_pre: pre-activation, kept only for a GELU embedder (its derivative needs it)
  float* rope_cos = nullptr; float* rope_sin = nullptr;   // (max_F, head size), models/ndt1.py:44-53
  // embedder.adapt: per-day weights / biases packed back to back (T-typed copy refreshed per forward) and their packed gradients
  T* wd_pack = nullptr; float* bd_pack = nullptr; float* dwd_pack = nullptr; float* dbd_pack = nullptr; int ldW = 0;
  std::vector<float*> xs;           // 2L+1 residual snapshots
  std::vector<float*> mean, rstd;   // 2L+1
  std::vector<T*> h1, h2, qkv, att, attd, u, g;
  std::vector<float*> lse;
  std::vector<unsigned int*> dropbits;   // keep bits of the attention-probability dropout (tensor-core path)
  std::vector<unsigned int*> dropbits_o; // keep bits of the attention-output dropout (tensor-core path)
  std::vector<unsigned int*> dropbits_m; // keep bits of the MLP dropout (read by the down-projection's epilogue instead of running Philox)
  bool bits_drawn = false;               // this forward drew the keep bits (the backward's epilogues may read them)
  cudaEvent_t bits_fork_ev = nullptr, bits_done_ev = nullptr;
  T* hn = nullptr; T* fac = nullptr; T* fpre = nullptr;
  float* logits = nullptr; float* logp = nullptr; float* dlogits = nullptr; float* nll = nullptr; float* ctc_ws = nullptr;
  long long* key_valid = nullptr; long long* out_lens = nullptr;
  unsigned long long* seed_dev = nullptr;      // this step's Philox key in device memory: every stochastic kernel reads it through a SeedRef,
                                               // so a captured step (CUDA graph) is replayed with a new key by updating 8 bytes
  float* feat32 = nullptr;
  // backward scratch
  // operands of the weight-gradient GEMMs are kept PER LAYER so that those GEMMs may trail the data-gradient chain
  float* dX = nullptr; T* dYe = nullptr; T* dH = nullptr; T* dA = nullptr; T* dlog = nullptr;
  std::vector<T*> dYm, dYa, dUl, dqkvl;
  cudaStream_t wstream = nullptr;            // weight-gradient stream
  std::vector<cudaEvent_t> fork_ev; size_t fork_used = 0;
  cudaEvent_t join_ev = nullptr;
  T* dEmb = nullptr; T* dhn = nullptr; T* dfac = nullptr;
  float* delta = nullptr; float* ln_part = nullptr;
  int ldV = 0, ldL = 0;
  // bf16 weight copies
  bf16* w_emb = nullptr; bf16* w_proj = nullptr; bf16* w_fac = nullptr; bf16* w_dec = nullptr;
  std::vector<bf16*> w_qkv, w_o, w_up, w_down;
  // the bf16 weights the GEMMs actually read: the engine's own copies (cast per forward) or, when the caller keeps a bf16
  // shadow of its flat fp32 parameter arena up to date (ndt1_engine_set_weight_shadow), pointers into that shadow
  const bf16* u_emb = nullptr; const bf16* u_proj = nullptr; const bf16* u_fac = nullptr; const bf16* u_dec = nullptr;
  std::vector<const bf16*> u_qkv, u_o, u_up, u_down;
  std::vector<float*> b_qkv;
  // state of the last forward
  int B = 0, Tn = 0, Tp = 0, L = 0, S = 0, training = 0; bool have_fwd = false, fwd_encoder_only = false;
  const float* spikes_ptr = nullptr; const long long* ts_ptr = nullptr; const long long* block_ptr = nullptr; const long long* day_ptr = nullptr;

  size_t arena_bytes() const override { return ar.cap; }

  void carve() {
    const ndt1_config& k = c;
    const int Bm = k.max_batch, Tm = k.max_T;
    const int Lm = n_prefix + out_len(Tm);
    const long long Mm = (long long)Bm * Lm, MT = (long long)Bm * Tm;
    const int H = k.hidden, I = k.inter, D = k.input_dim, NL = k.n_layers;
    const int Hout = k.factors_active ? k.factors_size : H;
    ldN = (k.n_channels + 7) / 8 * 8;
    ldV = (k.n_outputs + 7) / 8 * 8;
    const long long Mout = (k.method == NDT1_METHOD_CTC) ? (long long)Bm * out_len(Tm) : Mm;
    if (kBf16) xin = ar.take<T>(MT * ldN);
    emb = ar.take<T>(MT * D);
    if (k.embed_act == NDT1_ACT_GELU) emb_pre = ar.take<T>(MT * D);
    if (k.use_rope) { rope_cos = ar.take<float>((long long)k.max_F * (H / k.n_heads)); rope_sin = ar.take<float>((long long)k.max_F * (H / k.n_heads)); }
    ldW = kBf16 ? ldN : k.n_channels;
    if (k.adapt) {
      wd_pack = ar.take<T>((long long)k.n_days * D * ldW); bd_pack = ar.take<float>((long long)k.n_days * D);
      dwd_pack = ar.take<float>((long long)k.n_days * D * k.n_channels); dbd_pack = ar.take<float>((long long)k.n_days * D);
    }
    xs.resize(2 * NL + 1); mean.resize(2 * NL + 1); rstd.resize(2 * NL + 1);
    for (int i = 0; i < 2 * NL + 1; ++i) { xs[i] = ar.take<float>(Mm * H); mean[i] = ar.take<float>(Mm); rstd[i] = ar.take<float>(Mm); }
    h1.resize(NL); h2.resize(NL); qkv.resize(NL); att.resize(NL); attd.resize(NL); u.resize(NL); g.resize(NL); lse.resize(NL); dropbits.resize(NL); dropbits_o.resize(NL); dropbits_m.resize(NL);
    for (int l = 0; l < NL; ++l) {
      h1[l] = ar.take<T>(Mm * H); h2[l] = ar.take<T>(Mm * H); qkv[l] = ar.take<T>(Mm * 3 * H);
      att[l] = ar.take<T>(Mm * H); attd[l] = (k.p_transformer > 0.f) ? ar.take<T>(Mm * H) : att[l];
      u[l] = ar.take<T>(Mm * I); g[l] = ar.take<T>(Mm * I); lse[l] = ar.take<float>((long long)Bm * k.n_heads * Lm);
      dropbits[l] = (kBf16 && k.p_transformer > 0.f) ? ar.take<unsigned int>((long long)Bm * k.n_heads * Lm * 8) : nullptr;
      dropbits_o[l] = (kBf16 && k.p_transformer > 0.f && H % 32 == 0) ? ar.take<unsigned int>(Mm * (H / 32)) : nullptr;
      dropbits_m[l] = (kBf16 && k.p_transformer > 0.f && H % 32 == 0) ? ar.take<unsigned int>(Mm * (H / 32)) : nullptr;
    }
    hn = ar.take<T>(Mm * H);
    if (k.factors_active) { fac = ar.take<T>(Mm * Hout); fpre = ar.take<T>(Mm * Hout); dfac = ar.take<T>(Mm * Hout); }
    ldL = (k.method == NDT1_METHOD_CTC) ? (k.n_outputs + 3) / 4 * 4 : k.n_outputs;   // phoneme logits: rows padded to 16 bytes (vector epilogue)
    logits = ar.take<float>(Mout * ldL); logp = ar.take<float>(Mout * k.n_outputs); dlogits = ar.take<float>(Mout * k.n_outputs);
    nll = ar.take<float>(Bm);
    if (k.method == NDT1_METHOD_CTC) ctc_ws = ar.take<float>(k_ctc_workspace_floats(Bm, out_len(Tm), k.max_targets));
    key_valid = ar.take<long long>(Mm); out_lens = ar.take<long long>(Bm);
    seed_dev = ar.take<unsigned long long>(2);
    feat32 = ar.take<float>(Mm * Hout);
    dX = ar.take<float>(Mm * H); dYe = ar.take<T>(Mm * H); dH = ar.take<T>(Mm * H); dA = ar.take<T>(Mm * H);
    dYm.resize(NL); dYa.resize(NL); dUl.resize(NL); dqkvl.resize(NL);
    for (int l = 0; l < NL; ++l) {
      dYm[l] = ar.take<T>(Mm * H); dYa[l] = ar.take<T>(Mm * H); dUl[l] = ar.take<T>(Mm * I); dqkvl[l] = ar.take<T>(Mm * 3 * H);
    }
    dlog = ar.take<T>(Mout * ldV); dEmb = ar.take<T>(MT * D); dhn = ar.take<T>(Mm * H);
    delta = ar.take<float>((long long)Bm * k.n_heads * Lm);
    ln_part = (float*)ar.take<char>(k_layernorm_bwd_partials_bytes(H));
    if (kBf16) {
      const int KP = k.stack_active ? k.stack_size * D : D;
      w_emb = ar.take<bf16>((long long)D * ldN); w_proj = ar.take<bf16>((long long)H * KP);
      if (k.factors_active) w_fac = ar.take<bf16>((long long)Hout * H);
      w_dec = ar.take<bf16>((long long)k.n_outputs * Hout);
      w_qkv.resize(NL); w_o.resize(NL); w_up.resize(NL); w_down.resize(NL); b_qkv.resize(NL);
      u_qkv.resize(NL); u_o.resize(NL); u_up.resize(NL); u_down.resize(NL);
      for (int l = 0; l < NL; ++l) {
        w_qkv[l] = ar.take<bf16>(3LL * H * H); w_o[l] = ar.take<bf16>((long long)H * H);
        w_up[l] = ar.take<bf16>((long long)I * H); w_down[l] = ar.take<bf16>((long long)H * I);
        b_qkv[l] = ar.take<float>(3 * H);
      }
    }
  }

  int init() {
    n_prefix = (c.block_token ? 1 : 0) + (c.day_token ? 1 : 0);
    const char* fs = getenv("NDT1_FORCE_SIMT");
    force_simt = fs && fs[0] == '1';
    const char* sa = getenv("NDT1_SIMT_ATTENTION");
    simt_attention = sa && sa[0] == '1';
    carve();                       // dry run: sizes only
    ar.cap = ar.off + 256; ar.off = 0;
    NDT1_CUDA_CHECK(cudaMalloc((void**)&ar.base, ar.cap));
    NDT1_CUDA_CHECK(cudaMemset(ar.base, 0, ar.cap));
    carve();
    for (int i = 0; i < n_stages(); ++i) NDT1_CUDA_CHECK(cudaEventCreateWithFlags(&stage_ev[i], cudaEventDisableTiming));
    const char* ov = getenv("NDT1_OVERLAP");
    overlap = !(ov && ov[0] == '0');
    NDT1_CUDA_CHECK(cudaStreamCreateWithFlags(&wstream, cudaStreamNonBlocking));
    fork_ev.resize(6 * c.n_layers + 12);
    for (auto& e : fork_ev) NDT1_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    NDT1_CUDA_CHECK(cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming));
    NDT1_CUDA_CHECK(cudaEventCreateWithFlags(&bits_fork_ev, cudaEventDisableTiming));
    NDT1_CUDA_CHECK(cudaEventCreateWithFlags(&bits_done_ev, cudaEventDisableTiming));
    if (kBf16 && !force_simt) NDT1_TRY(gemm_tc_init());
    if (c.use_rope) {
      // get_cos_sin (models/ndt1.py:44-53): inv_freq_i = base^(-2i/dim), angle = t * inv_freq_i, table = cat(angles, angles)
      const int hd = c.hidden / c.n_heads, half = hd / 2;
      std::vector<float> hc((size_t)c.max_F * hd), hs((size_t)c.max_F * hd);
      for (int i = 0; i < half; ++i) {
        const float inv = (float)(1.0 / pow((double)c.rope_theta, (double)(2 * i) / (double)hd));
        for (int t = 0; t < c.max_F; ++t) {
          const float ang = (float)t * inv;
          hc[(size_t)t * hd + i] = hc[(size_t)t * hd + half + i] = (float)cos((double)ang);
          hs[(size_t)t * hd + i] = hs[(size_t)t * hd + half + i] = (float)sin((double)ang);
        }
      }
      NDT1_CUDA_CHECK(cudaMemcpy(rope_cos, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
      NDT1_CUDA_CHECK(cudaMemcpy(rope_sin, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
    }
    return 0;
  }
  int set_rope_tables(const float* cs, const float* sn, int rows) override {
    NDT1_REQUIRE(c.use_rope, "engine_set_rope_tables: the engine was created without use_rope");
    NDT1_REQUIRE(cs && sn && rows >= c.max_F, "engine_set_rope_tables: need cos and sin tables of at least max_F = %d rows", c.max_F);
    const size_t bytes = (size_t)c.max_F * (c.hidden / c.n_heads) * 4;
    NDT1_CUDA_CHECK(cudaMemcpy(rope_cos, cs, bytes, cudaMemcpyDeviceToDevice));
    NDT1_CUDA_CHECK(cudaMemcpy(rope_sin, sn, bytes, cudaMemcpyDeviceToDevice));
    return 0;
  }
  ~Engine() override {
    for (int i = 0; i < NDT1_MAX_LAYERS + 2; ++i) if (stage_ev[i]) cudaEventDestroy(stage_ev[i]);
    for (auto& e : fork_ev) if (e) cudaEventDestroy(e);
    if (join_ev) cudaEventDestroy(join_ev);
    if (bits_fork_ev) cudaEventDestroy(bits_fork_ev);
    if (bits_done_ev) cudaEventDestroy(bits_done_ev);
    if (wstream) cudaStreamDestroy(wstream);
    if (ar.base) cudaFree(ar.base);
  }

  // ---- GEMM helpers -------------------------------------------------------
  int run(GemmProblem& p, cudaStream_t s) {
    if (kBf16 && !force_simt) return gemm_tc_launch(p, s);
    return gemm_simt_launch(p, kBf16, s);
  }
  static GemmOperand op(const void* ptr, long long bs, int nb, int rows, int cols, int ld) {
    GemmOperand o; o.ptr = ptr; o.batch_stride = bs; o.nbatch = nb; o.rows = rows; o.cols = cols; o.ld = ld; return o;
  }
  static GemmProblem prob(int mode, int M, int N, int K) {
    GemmProblem p;
    p.b_sel = nullptr;
    p.mode = mode; p.M = M; p.N = N; p.nb_out = 1; p.nchunk = 1; p.chunk_k = K;
    p.a_row_shift = p.a_col_shift = p.b_row_shift = p.b_col_shift = 0; p.b_chunk_n = N; p.split_k = 1;
    p.epi = gemm_epilogue_default();
    return p;
  }
  // y[M,N] = epi(a[M,K] w[N,K]^T)
  int linear_fwd(const T* a, int lda, const void* w, int ldw, int M, int N, int K, GemmEpilogue e, cudaStream_t s) {
    GemmProblem p = prob(GEMM_NT, M, N, K);
    p.A = op(a, 0, 1, M, K, lda); p.B = op(w, 0, 1, N, K, ldw); p.epi = e;
    return run(p, s);
  }
  // dx[M,K] = epi(dy[M,N] w[N,K])
  int linear_dgrad(const T* dy, int lddy, const void* w, int ldw, int M, int N, int K, GemmEpilogue e, cudaStream_t s) {
    GemmProblem p = prob(GEMM_NN, M, K, N);
    p.A = op(dy, 0, 1, M, N, lddy); p.B = op(w, 0, 1, N, K, ldw); p.epi = e;
    return run(p, s);
  }
  // dw[N,K] += dy[M,N]^T a[M,K]
  int linear_wgrad(const T* dy, int lddy, const T* a, int lda, float* dw, int lddw, int M, int N, int K, cudaStream_t s) {
    if (!dw) return 0;
    GemmProblem p = prob(GEMM_TN, N, K, M);
    p.A = op(dy, 0, 1, M, N, lddy); p.B = op(a, 0, 1, M, K, lda);
    p.epi.out = dw; p.epi.ldc = lddw; p.epi.accumulate = 1;
    // tensor-core kernel: 0 = let the launcher size the split for one wave of its (CTA or CTA-pair) work units
    p.split_k = (kBf16 && !force_simt) ? 0 : pick_split(ndt1_cdiv(N, 128) * ndt1_cdiv(K, K > 128 ? 256 : (K > 64 ? 128 : 64)), ndt1_cdiv(M, 64));
    return run(p, s);
  }
  // split the reduction so that tiles*split fills whole waves of 148 SMs
  int pick_split(int tiles, int kblocks) const {
    if (!(kBf16 && !force_simt)) { int sp = 148 / (tiles > 0 ? tiles : 1); return sp > 8 ? 8 : (sp < 1 ? 1 : sp); }
    // every split adds a full fp32 red.add pass over the output (the L2 atomics are what bounds these kernels),
    // so take the largest split that still fits ONE wave of CTAs
    int best = 148 / (tiles > 0 ? tiles : 1);
    while (best > 1 && best * 4 > kblocks) --best;
    return best < 1 ? 1 : best;
  }
  const void* W(const float* master, const bf16* copy) const { return kBf16 ? (const void*)copy : (const void*)master; }

  unsigned long long site_attn_p(int l) const { return 1 + 4ull * l; }
  unsigned long long site_attn_o(int l) const { return 2 + 4ull * l; }
  unsigned long long site_mlp(int l) const { return 3 + 4ull * l; }
  unsigned long long site_factors() const { return 1 + 4ull * c.n_layers; }     // dropout in front of the factors projection (models/ndt1.py:355,372)

  int act_code(int a) const { return a == NDT1_ACT_SOFTSIGN ? ACT_SOFTSIGN : a == NDT1_ACT_GELU ? ACT_GELU : a == NDT1_ACT_RELU ? ACT_RELU : ACT_NONE; }
  // derivative selector given what the forward kept (out = activation value, in = pre-activation)
  int dact_from_out(int a) const { return a == NDT1_ACT_SOFTSIGN ? DACT_SOFTSIGN_FROM_OUT : a == NDT1_ACT_RELU ? DACT_RELU_FROM_OUT : DACT_NONE; }

  void ctx(int& f, int& b) const {
    if (c.context_forward == -2 && c.context_backward == -2) { f = kUnbounded; b = kUnbounded; return; }
    f = c.context_forward >= -1 ? c.context_forward : kUnbounded;
    b = c.context_backward >= -1 ? c.context_backward : kUnbounded;
  }

  // ---- forward ------------------------------------------------------------
  int forward(const ndt1_tensors* P, const ndt1_batch* bt, const ndt1_outputs* o, cudaStream_t s) override {
    const ndt1_config& k = c;
    const long long launches0 = g_ndt1_launches;
    have_fwd = false;
    NDT1_REQUIRE(bt->B >= 0 && bt->B <= k.max_batch, "engine: batch %d exceeds max_batch %d", bt->B, k.max_batch);
    NDT1_REQUIRE(bt->T <= k.max_T, "engine: %d bins exceed max_T %d", bt->T, k.max_T);
    NDT1_REQUIRE(!k.stack_active || bt->T >= k.stack_size, "engine: %d bins are fewer than the stack size %d", bt->T, k.stack_size);
    NDT1_REQUIRE(!k.pos || out_len(bt->T) <= k.max_F, "engine: %d positions exceed max_F %d", out_len(bt->T), k.max_F);
    NDT1_REQUIRE(k.method != NDT1_METHOD_CTC || bt->S <= k.max_targets, "engine: %d targets exceed max_targets %d", bt->S, k.max_targets);
    B = bt->B; Tn = bt->T; Tp = out_len(Tn); L = n_prefix + Tp; S = bt->S; training = bt->training;
    spikes_ptr = bt->spikes; ts_ptr = (const long long*)bt->spikes_timestamp; block_ptr = (const long long*)bt->block_idx; day_ptr = (const long long*)bt->day_idx;
    if (B == 0) {
      // an empty shard (global batch < world size): loss 0, no examples; the other outputs have no elements
      NDT1_CUDA_CHECK(cudaMemsetAsync(o->loss, 0, sizeof(float), s));
      if (o->n_examples) NDT1_CUDA_CHECK(cudaMemsetAsync(o->n_examples, 0, 8, s));
      fwd_encoder_only = bt->encoder_only != 0;
      have_fwd = bt->need_backward != 0;
      return 0;
    }
    const int H = k.hidden, I = k.inter, D = k.input_dim, N = k.n_channels, NL = k.n_layers, V = k.n_outputs;
    const int Hout = k.factors_active ? k.factors_size : H;
    const long long M = (long long)B * L, MT = (long long)B * Tn;
    const float pe = training ? k.p_embed : 0.f, ptr_ = training ? k.p_transformer : 0.f;
    // the step's Philox key goes to device memory: copied from the caller's device word (graph replays re-read it), or set from the value
    if (bt->seed_ptr) NDT1_CUDA_CHECK(cudaMemcpyAsync(seed_dev, bt->seed_ptr, 8, cudaMemcpyDeviceToDevice, s));
    else NDT1_TRY(k_set_i64((long long*)seed_dev, (long long)bt->seed, s));
    const SeedRef seed = SeedRef::at(seed_dev);
    // The keep bits of the attention dropouts (probabilities and output, every layer) depend on the key only: one launch on the
    // second stream, concurrent with the embedding GEMMs; the first attention layer waits for it.
    const bool tc_attention = kBf16 && !force_simt && !simt_attention && (H / k.n_heads) == 128 && L <= 256 && L >= 1 && H % 32 == 0;
    const bool draw_bits = tc_attention && ptr_ > 0.f;
    // (the GEMM epilogues of the MLP dropout and of the attention-output dropout's backward read the same kind of bits, in
    //  classes without any Philox code; NDT1_GEMM_DROPBITS=0 keeps them on Philox: identical masks either way)
    static const bool no_gemm_bits = getenv("NDT1_GEMM_DROPBITS") && getenv("NDT1_GEMM_DROPBITS")[0] == '0';
    const bool use_gemm_bits = draw_bits && !no_gemm_bits;
    bits_drawn = use_gemm_bits;
    if (draw_bits) {
      AttnBitsJob job; memset(&job, 0, sizeof(job));
      job.n = NL;
      for (int l = 0; l < NL; ++l) {
        job.bits_p[l] = dropbits[l]; job.bits_o[l] = dropbits_o[l]; job.bits_m[l] = use_gemm_bits ? dropbits_m[l] : nullptr;
        job.stream_p[l] = site_attn_p(l); job.stream_o[l] = site_attn_o(l); job.stream_m[l] = site_mlp(l);
      }
      AttnParams shp; memset(&shp, 0, sizeof(shp));
      shp.B = B; shp.L = L; shp.H = H; shp.nh = k.n_heads; shp.hd = H / k.n_heads; shp.p_attn = ptr_; shp.p_out = ptr_; shp.seed = seed;
      cudaStream_t bs = overlap ? wstream : s;
      if (overlap) { NDT1_CUDA_CHECK(cudaEventRecord(bits_fork_ev, s)); NDT1_CUDA_CHECK(cudaStreamWaitEvent(bs, bits_fork_ev, 0)); }
      NDT1_TRY(k_attention_tc_dropbits(job, shp, bs));
      if (overlap) NDT1_CUDA_CHECK(cudaEventRecord(bits_done_ev, bs));
    }

    // 0. precision staging: bf16 copies of the weights and of the input
    xin_ext = nullptr;
    if (kBf16) {
      if (bt->spikes_bf16 && ldN == N) xin_ext = (const T*)bt->spikes_bf16;       // the prologue wrote the bf16 operand itself
      else NDT1_TRY(k_cast_f32_bf16(bt->spikes, (bf16*)xin, MT, N, N, ldN, s));
      CastSegs cs; cs.n = 0;
      // a weight either comes from the caller's bf16 shadow (no work here) or is cast into the engine's copy:
      // contiguous matrices through ONE multi-segment launch, anything ragged through the strided kernel
      int rc_add = 0;
      auto use = [&](const float* src, bf16* dst, long long rows, int cols, long long ld_out) -> const bf16* {
        const long long cnt = rows * cols;
        if (shadow_bf16 && ld_out == cols && src >= shadow_src && src + cnt <= shadow_src + shadow_n) {
          const bf16* sp = (const bf16*)shadow_bf16 + (src - shadow_src);
          if (((uintptr_t)sp & 15) == 0) return sp;
        }
        if (ld_out == cols && cnt % 8 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 && cs.n < CAST_MAX_SEGS) {
          cs.src[cs.n] = src; cs.dst[cs.n] = dst; cs.count[cs.n] = cnt; ++cs.n;
        } else if (!rc_add) {
          rc_add = k_cast_f32_bf16(src, dst, rows, cols, cols, ld_out, s);
        }
        return dst;
      };
      if (!k.adapt) u_emb = use(P->embed_w, w_emb, D, N, ldN);
      const int KP = k.stack_active ? k.stack_size * D : D;
      u_proj = use(P->proj_w, w_proj, H, KP, KP);
      for (int l = 0; l < NL; ++l) {
        const auto& q = P->layer[l];
        const bf16* uq = use(q.q_w, w_qkv[l], H, H, H);
        const bf16* uk = use(q.k_w, w_qkv[l] + (long long)H * H, H, H, H);
        const bf16* uv = use(q.v_w, w_qkv[l] + 2LL * H * H, H, H, H);
        if (uk != uq + (long long)H * H || uv != uk + (long long)H * H) {      // shadow without adjacent q|k|v: fall back to the engine's packed copy
          NDT1_TRY(k_cast_f32_bf16(q.q_w, w_qkv[l], H, H, H, H, s));
          NDT1_TRY(k_cast_f32_bf16(q.k_w, w_qkv[l] + (long long)H * H, H, H, H, H, s));
          NDT1_TRY(k_cast_f32_bf16(q.v_w, w_qkv[l] + 2LL * H * H, H, H, H, H, s));
          uq = w_qkv[l];
        }
        u_qkv[l] = uq;
        u_o[l] = use(q.o_w, w_o[l], H, H, H);
        u_up[l] = use(q.up_w, w_up[l], I, H, H);
        u_down[l] = use(q.down_w, w_down[l], H, I, I);
        if (k.attention_bias && !(q.k_b == q.q_b + H && q.v_b == q.k_b + H)) {   // scattered biases: gather them once
          NDT1_CUDA_CHECK(cudaMemcpyAsync(b_qkv[l], q.q_b, H * 4, cudaMemcpyDeviceToDevice, s));
          NDT1_CUDA_CHECK(cudaMemcpyAsync(b_qkv[l] + H, q.k_b, H * 4, cudaMemcpyDeviceToDevice, s));
          NDT1_CUDA_CHECK(cudaMemcpyAsync(b_qkv[l] + 2 * H, q.v_b, H * 4, cudaMemcpyDeviceToDevice, s));
        }
      }
      if (k.factors_active) u_fac = use(P->factors_w, w_fac, Hout, H, H);
      u_dec = use(P->dec_w, w_dec, V, Hout, Hout);
      NDT1_TRY(rc_add);
      NDT1_TRY(k_cast_multi(cs, s));
    }
    if (k.adapt) {
      // per-day embedding: pack this step's weights / biases back to back (they may live anywhere)
      NDT1_REQUIRE(k.n_days >= 1 && k.n_days <= NDT1_MAX_DAYS, "engine: n_days %d outside [1,%d]", k.n_days, NDT1_MAX_DAYS);
      NDT1_REQUIRE(day_ptr, "engine: embedder.adapt needs day_idx");
      for (int d = 0; d < k.n_days; ++d) {
        NDT1_REQUIRE(P->embed_w_day[d], "engine: embed_w_day[%d] is null", d);
        if (kBf16) NDT1_TRY(k_cast_f32_bf16(P->embed_w_day[d], (bf16*)wd_pack + (long long)d * D * ldW, D, N, N, ldW, s));
        else NDT1_CUDA_CHECK(cudaMemcpyAsync(wd_pack + (long long)d * D * ldW, P->embed_w_day[d], (size_t)D * N * 4, cudaMemcpyDeviceToDevice, s));
        if (k.embed_bias) {
          NDT1_REQUIRE(P->embed_b_day[d], "engine: embed_b_day[%d] is null", d);
          NDT1_CUDA_CHECK(cudaMemcpyAsync(bd_pack + (long long)d * D, P->embed_b_day[d], (size_t)D * 4, cudaMemcpyDeviceToDevice, s));
        }
      }
    }
    const T* x_in = kBf16 ? (xin_ext ? xin_ext : xin) : (const T*)bt->spikes;
    const int ldx = kBf16 ? ldN : N;

    // 1. channel embedding + activation   (models/ndt1.py:170-176)
    {
      GemmEpilogue e = gemm_epilogue_default();
      e.out = emb; e.out_bf16 = kBf16; e.ldc = D; e.bias = k.embed_bias ? P->embed_b : nullptr; e.act = act_code(k.embed_act);
      if (emb_pre) { e.out2 = emb_pre; e.out2_bf16 = kBf16; }
      if (!k.adapt) {
        NDT1_TRY(linear_fwd(x_in, ldx, W(P->embed_w, u_emb), kBf16 ? ldN : N, (int)MT, D, N, e, s));
      } else {
        // one GEMM over all trials; trial b reads day_idx[b]'s weight matrix (third TMA coordinate) and bias row
        GemmProblem p = prob(GEMM_NT, Tn, D, N);
        p.nb_out = B; p.b_sel = day_ptr;
        p.A = op(x_in, (long long)Tn * ldx, B, Tn, N, ldx);
        p.B = op(wd_pack, (long long)D * ldW, k.n_days, D, N, ldW);
        e.c_batch_stride = (long long)Tn * D;
        e.bias = k.embed_bias ? bd_pack : nullptr; e.sel = day_ptr; e.sel_n = k.n_days; e.bias_sel_stride = D;
        p.epi = e;
        NDT1_TRY(run(p, s));
      }
    }
    // 2. stack projection / projection + position table + embedding dropout  (models/ndt1.py:179-203)
    float* x0 = xs[0];
    {
      GemmEpilogue e = gemm_epilogue_default();
      e.out = x0 + (long long)n_prefix * H; e.out_bf16 = 0; e.ldc = H; e.c_batch_stride = (long long)L * H;
      e.bias = P->proj_b;
      if (k.pos) { e.gather_tab = P->pos_w; e.gather_idx = ts_ptr; e.gather_idx_stride = Tn; e.gather_ld = H; }
      if (pe > 0.f && n_prefix == 0) { e.drop_p = pe; e.drop_seed = seed; e.drop_stream = 0; }
      GemmProblem p;
      if (k.stack_active) {
        NDT1_REQUIRE(k.stack_size % k.stack_stride == 0, "engine: stack size %d must be a multiple of the stride %d", k.stack_size, k.stack_stride);
        const int K4 = k.stack_stride * D, nch = k.stack_size / k.stack_stride;
        p = prob(GEMM_NT, Tp, H, K4);
        p.nb_out = B; p.nchunk = nch; p.a_row_shift = 1; p.b_col_shift = K4;
        p.A = op(emb, (long long)Tn * D, B, Tn / k.stack_stride, K4, K4);
        p.B = op(W(P->proj_w, u_proj), 0, 1, H, nch * K4, nch * K4);
      } else {
        p = prob(GEMM_NT, Tp, H, D);
        p.nb_out = B;
        p.A = op(emb, (long long)Tn * D, B, Tn, D, D);
        p.B = op(W(P->proj_w, u_proj), 0, 1, H, D, D);
      }
      p.epi = e;
      NDT1_TRY(run(p, s));
      int slot = 0;
      if (k.day_token) { NDT1_TRY(k_token_rows(P->day_emb, day_ptr, x0, B, L, H, slot, s)); ++slot; }
      if (k.block_token) { NDT1_TRY(k_token_rows(P->block_emb, block_ptr, x0, B, L, H, slot, s)); ++slot; }
      if (pe > 0.f && n_prefix > 0) NDT1_TRY(k_grad_prep<float>(x0, x0, M, H, pe, seed, 0, nullptr, nullptr, 0, 1, 0, 0, s));
    }
    NDT1_TRY(k_stack_mask((const long long*)bt->spikes_mask, key_valid, B, Tn, Tp, k.stack_active ? k.stack_size : 0, k.stack_active ? k.stack_stride : 1,
                          n_prefix, s));
    if (o->out_mask) NDT1_CUDA_CHECK(cudaMemcpyAsync(o->out_mask, key_valid, M * 8, cudaMemcpyDeviceToDevice, s));

    // 3. transformer layers  (models/ndt1.py:317-330)
    NDT1_REQUIRE(!k.use_rope || (n_prefix == 0 && ts_ptr), "engine: use_rope needs timestamps and no block / day token (the reference indexes cos/sin by the stacked timestamps)");
    int cf, cb; ctx(cf, cb);
    for (int l = 0; l < NL; ++l) {
      const auto& q = P->layer[l];
      float* xa = xs[2 * l]; float* xm = xs[2 * l + 1]; float* xo = xs[2 * l + 2];
      NDT1_TRY(k_layernorm_fwd<T>(xa, q.ln1_w, q.ln1_b, h1[l], mean[2 * l], rstd[2 * l], M, H, 1e-5f, s));
      if (kBf16) {
        GemmEpilogue e = gemm_epilogue_default();
        const bool packed_bias = (q.k_b == q.q_b + H && q.v_b == q.k_b + H);   // flat parameter arena: q|k|v biases adjacent
        e.out = qkv[l]; e.out_bf16 = 1; e.ldc = 3 * H; e.bias = k.attention_bias ? (packed_bias ? q.q_b : b_qkv[l]) : nullptr;
        NDT1_TRY(linear_fwd(h1[l], H, u_qkv[l], H, (int)M, 3 * H, H, e, s));
      } else {
        const float* ws[3] = {q.q_w, q.k_w, q.v_w}; const float* bs[3] = {q.q_b, q.k_b, q.v_b};
        for (int j = 0; j < 3; ++j) {
          GemmEpilogue e = gemm_epilogue_default();
          e.out = qkv[l] + (long long)j * H; e.out_bf16 = 0; e.ldc = 3 * H; e.bias = k.attention_bias ? bs[j] : nullptr;
          NDT1_TRY(linear_fwd(h1[l], H, ws[j], H, (int)M, H, H, e, s));
        }
      }
      if (k.use_rope) NDT1_TRY(k_rope<T>(qkv[l], ts_ptr, Tn, rope_cos, rope_sin, M, L, H, k.n_heads, k.max_F, 0, s));
      AttnParams ap;
      ap.qkv = qkv[l]; ap.out = att[l]; ap.out_drop = (ptr_ > 0.f) ? attd[l] : att[l]; ap.lse = lse[l]; ap.key_valid = key_valid;
      ap.B = B; ap.L = L; ap.H = H; ap.nh = k.n_heads; ap.hd = H / k.n_heads; ap.ctx_fwd = cf; ap.ctx_bwd = cb;
      ap.scale = 1.0f / sqrtf((float)(H / k.n_heads)); ap.p_attn = ptr_; ap.p_out = ptr_;
      ap.seed = seed; ap.stream_attn = site_attn_p(l); ap.stream_out = site_attn_o(l);
      ap.dout = nullptr; ap.dqkv = nullptr; ap.delta = nullptr; ap.drop_bits = dropbits[l]; ap.drop_bits_o = dropbits_o[l]; ap.bits_ready = draw_bits;
      if (l == 0 && draw_bits && overlap) NDT1_CUDA_CHECK(cudaStreamWaitEvent(s, bits_done_ev, 0));
      if (kBf16 && !force_simt && !simt_attention && k_attention_tc_supported(ap)) NDT1_TRY(k_attention_tc_fwd(ap, s));
      else NDT1_TRY(k_attention_fwd<T>(ap, s));
      {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = xm; e.ldc = H; e.bias = k.attention_bias ? q.o_b : nullptr; e.resid = xa;
        NDT1_TRY(linear_fwd((const T*)ap.out_drop, H, W(q.o_w, kBf16 ? u_o[l] : nullptr), H, (int)M, H, H, e, s));
      }
      NDT1_TRY(k_layernorm_fwd<T>(xm, q.ln2_w, q.ln2_b, h2[l], mean[2 * l + 1], rstd[2 * l + 1], M, H, 1e-5f, s));
      {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = g[l]; e.out_bf16 = kBf16; e.ldc = I; e.bias = k.mlp_bias ? q.up_b : nullptr; e.act = act_code(k.mlp_act);
        e.out2 = u[l]; e.out2_bf16 = kBf16; e.out2_deriv = k.mlp_act == NDT1_ACT_GELU;     // u[l] = GELU'(pre-activation) then
        NDT1_TRY(linear_fwd(h2[l], H, W(q.up_w, kBf16 ? u_up[l] : nullptr), H, (int)M, I, H, e, s));
      }
      {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = xo; e.ldc = H; e.bias = k.mlp_bias ? q.down_b : nullptr; e.resid = xm;
        if (ptr_ > 0.f) { e.drop_p = ptr_; e.drop_seed = seed; e.drop_stream = site_mlp(l); if (bits_drawn) e.drop_bits = dropbits_m[l]; }
        NDT1_TRY(linear_fwd(g[l], I, W(q.down_w, kBf16 ? u_down[l] : nullptr), I, (int)M, H, I, e, s));
      }
    }
    // 4. output norm, factors, head   (models/ndt1.py:442-450, 545)
    NDT1_TRY(k_layernorm_fwd<T>(xs[2 * NL], P->out_norm_w, P->out_norm_b, hn, mean[2 * NL], rstd[2 * NL], M, H, 1e-5f, s));
    const float pf = training ? k.p_factors : 0.f;       // NeuralFactorsProjection.forward drops its input whether or not the projection is active
    NDT1_TRY(k_dropout_inplace<T>(hn, M * H, pf, seed, site_factors(), s));
    const T* head_in = hn; int head_ld = H;
    if (k.factors_active) {
      GemmEpilogue e = gemm_epilogue_default();
      e.out = fac; e.out_bf16 = kBf16; e.ldc = Hout; e.bias = k.factors_bias ? P->factors_b : nullptr; e.act = act_code(k.factors_act);
      e.out2 = fpre; e.out2_bf16 = kBf16;
      NDT1_TRY(linear_fwd(hn, H, W(P->factors_w, u_fac), H, (int)M, Hout, H, e, s));
      head_in = fac; head_ld = Hout;
    }
    if (o->features) {
      // fp32 copy of the encoder output without the prefix tokens
      if (k.factors_active) {
        NDT1_TRY((k_scale_cast_features(head_in, feat32, M, Hout, s)));
      } else {
        NDT1_TRY(k_layernorm_fwd<float>(xs[2 * NL], P->out_norm_w, P->out_norm_b, feat32, mean[2 * NL], rstd[2 * NL], M, H, 1e-5f, s));
        NDT1_TRY(k_dropout_inplace<float>(feat32, M * H, pf, seed, site_factors(), s));
      }
      NDT1_CUDA_CHECK(cudaMemcpy2DAsync(o->features, (size_t)Tp * Hout * 4, feat32 + (long long)n_prefix * Hout, (size_t)L * Hout * 4,
                                        (size_t)Tp * Hout * 4, B, cudaMemcpyDeviceToDevice, s));
    }
    fwd_encoder_only = bt->encoder_only != 0;
    if (bt->encoder_only) {
      have_fwd = bt->need_backward != 0;     // ndt1_engine_backward_features continues from d(features)
      NDT1_CUDA_CHECK(cudaMemsetAsync(o->loss, 0, sizeof(float), s));
      launches = g_ndt1_launches - launches0;
      return 0;
    }
    {
      GemmProblem p = prob(GEMM_NT, Tp, V, Hout);
      p.nb_out = B;
      p.A = op(head_in + (long long)n_prefix * head_ld, (long long)L * head_ld, B, Tp, Hout, head_ld);
      p.B = op(W(P->dec_w, u_dec), 0, 1, V, Hout, Hout);
      p.epi.out = logits; p.epi.ldc = ldL; p.epi.c_batch_stride = (long long)Tp * ldL; p.epi.bias = P->dec_b;
      NDT1_TRY(run(p, s));
    }
    // 5. loss   (models/ndt1.py:548-589)
    NDT1_CUDA_CHECK(cudaMemsetAsync(o->loss, 0, sizeof(float), s));
    NDT1_TRY(k_stacked_lens((const long long*)bt->spikes_lengths, out_lens, B, k.stack_active, k.stack_size, k.stack_stride, s));
    if (o->out_lengths) NDT1_CUDA_CHECK(cudaMemcpyAsync(o->out_lengths, out_lens, B * 8, cudaMemcpyDeviceToDevice, s));
    const long long Mo = (long long)B * Tp;
    if (k.method == NDT1_METHOD_CTC) {
      NDT1_REQUIRE(bt->targets && bt->targets_lengths, "engine: ctc needs targets and targets_lengths");
      NDT1_TRY(k_log_softmax(logits, logp, Mo, V, s, ldL));
      NDT1_TRY(k_ctc_fwd_bwd(logp, (const long long*)bt->targets, out_lens, (const long long*)bt->targets_lengths, B, Tp, V, S, k.blank_id, k.zero_infinity, ctc_ws, nll,
                             o->loss, bt->need_backward ? dlogits : nullptr, nullptr, s, kBf16));
      if (o->preds) NDT1_CUDA_CHECK(cudaMemcpyAsync(o->preds, logp, Mo * V * 4, cudaMemcpyDeviceToDevice, s));
      if (o->n_examples) NDT1_TRY(k_set_i64((long long*)o->n_examples, B, s));
    } else {
      NDT1_REQUIRE(bt->recon_targets, "engine: mlm/autoregressive need the original spikes as targets");
      NDT1_REQUIRE(k.method != NDT1_METHOD_MLM || bt->targets_mask, "engine: mlm needs the masker's targets_mask");
      NDT1_REQUIRE(n_prefix == 0 && !k.stack_active, "engine: ssl methods need unstacked inputs without prefix tokens");
      if (k.decoder_relu) NDT1_TRY(k_relu_inplace(logits, Mo * V, s));
      if (o->n_examples) NDT1_CUDA_CHECK(cudaMemsetAsync(o->n_examples, 0, 8, s));
      NDT1_TRY(k_recon_loss(logits, bt->recon_targets, bt->need_backward ? dlogits : nullptr, (const long long*)bt->targets_mask, key_valid, B,
                            Tn, V, k.loss_kind, k.method == NDT1_METHOD_AUTOREGRESSIVE, k.decoder_relu, o->loss, (long long*)o->n_examples,
                            nullptr, s));
      if (o->preds) NDT1_CUDA_CHECK(cudaMemcpyAsync(o->preds, logits, Mo * V * 4, cudaMemcpyDeviceToDevice, s));
      if (o->loss_mask && k.method == NDT1_METHOD_MLM)
        NDT1_TRY(k_and_mask((const long long*)bt->targets_mask, key_valid, (long long*)o->loss_mask, B, Tn, V, s));
    }
    have_fwd = bt->need_backward != 0;
    launches = g_ndt1_launches - launches0;
    return 0;
  }

  int k_scale_cast_features(const T* in, float* out, long long rows, int cols, cudaStream_t s) {
    // T -> fp32 copy (the optional `features` output with an active factors projection)
    return k_cast_to_f32<T>(in, out, rows * cols, s);
  }

  // ---- backward -----------------------------------------------------------
  int backward(const ndt1_tensors* P, const ndt1_tensors* G, const float* dloss, cudaStream_t s, const float* dfeatures) override {
    const ndt1_config& k = c;
    NDT1_REQUIRE(have_fwd, "engine: backward without a matching forward (need_backward = 1)");
    NDT1_REQUIRE(fwd_encoder_only == (dfeatures != nullptr), "engine: an encoder-only forward is continued by ndt1_engine_backward_features, a full one by ndt1_engine_backward");
    const long long launches0 = g_ndt1_launches;
    // Inside a stream capture (the step as a CUDA graph) the stage events become EXTERNAL event-record nodes: every replay records
    // the real event, so a stream outside the graph (the trainer's all-reduce / optimizer stream) can wait on it.
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    NDT1_CUDA_CHECK(cudaStreamIsCapturing(s, &cap_status));
    const bool capturing = cap_status == cudaStreamCaptureStatusActive;
    auto record_stage = [&](int i, cudaStream_t st) -> cudaError_t {
      return capturing ? cudaEventRecordWithFlags(stage_ev[i], st, cudaEventRecordExternal) : cudaEventRecord(stage_ev[i], st);
    };
    if (B == 0) {       // empty shard: no gradient contribution, but every stage is "complete" for ndt1_engine_wait_stage
      for (int i = 0; i < n_stages(); ++i) NDT1_CUDA_CHECK(record_stage(i, s));
      return 0;
    }
    const int H = k.hidden, I = k.inter, D = k.input_dim, N = k.n_channels, NL = k.n_layers, V = k.n_outputs;
    const int Hout = k.factors_active ? k.factors_size : H;
    const long long M = (long long)B * L, MT = (long long)B * Tn, Mo = (long long)B * Tp;
    const float pe = training ? k.p_embed : 0.f, ptr_ = training ? k.p_transformer : 0.f;
    const SeedRef seed = SeedRef::at(seed_dev);          // (still this forward's key)

    // Weight gradients (and the bias reductions that are not fused elsewhere) go to `ws`: they only need the operand the
    // data-gradient chain has just produced, so they run concurrently with the rest of that chain and fill the SMs its
    // persistent kernels leave idle in their last wave.  fork() orders `ws` after everything enqueued on `s` so far.
    cudaStream_t ws = overlap ? wstream : s;
    fork_used = 0;
    auto fork = [&]() -> int {
      if (!overlap) return 0;
      NDT1_REQUIRE(fork_used < fork_ev.size(), "engine: out of fork events");
      cudaEvent_t e = fork_ev[fork_used++];
      NDT1_CUDA_CHECK(cudaEventRecord(e, s));
      NDT1_CUDA_CHECK(cudaStreamWaitEvent(ws, e, 0));
      return 0;
    };
    const T* head_in = k.factors_active ? fac : hn; const int head_ld = k.factors_active ? Hout : H;
    T* d_head_in = k.factors_active ? dfac : dhn;
    if (n_prefix > 0) NDT1_CUDA_CHECK(cudaMemsetAsync(d_head_in, 0, M * head_ld * sizeof(T), s));
    if (dfeatures) {
      // encoder-only (NeuralEncoder.forward as used by models/bci.py:125): the caller's gradient w.r.t. the (B, T', H_out) features
      for (int b = 0; b < (n_prefix > 0 ? B : 1); ++b) {
        const long long rows = n_prefix > 0 ? Tp : Mo;
        NDT1_TRY(k_scale_cast_pad<T>(dfeatures + (long long)b * Tp * Hout, d_head_in + ((long long)b * L + n_prefix) * head_ld, rows, Hout, head_ld, nullptr, s));
      }
      if (k.factors_active) {
        if (k.factors_act == NDT1_ACT_GELU) NDT1_TRY(k_dact_inplace<T>(dfac, fpre, M * Hout, DACT_GELU_FROM_IN, s));
        else NDT1_TRY(k_dact_inplace<T>(dfac, fac, M * Hout, dact_from_out(k.factors_act), s));
      }
    } else {
    // head
    NDT1_TRY(k_scale_cast_pad<T>(dlogits, dlog, Mo, V, ldV, dloss, s));
    NDT1_TRY(fork());
    if (G->dec_b) NDT1_TRY(k_colsum<T>(dlog, G->dec_b, Mo, V, ldV, ws));
    if (G->dec_w) {
      GemmProblem p = prob(GEMM_TN, V, Hout, Tp);
      p.nchunk = B;
      p.A = op(dlog, (long long)Tp * ldV, B, Tp, V, ldV);
      p.B = op(head_in + (long long)n_prefix * head_ld, (long long)L * head_ld, B, Tp, Hout, head_ld);
      p.epi.out = G->dec_w; p.epi.ldc = Hout; p.epi.accumulate = 1;
      if (n_prefix == 0) {  // rows are contiguous across trials: one long reduction
        p.nchunk = 1; p.chunk_k = (int)Mo;
        p.A = op(dlog, 0, 1, (int)Mo, V, ldV); p.B = op(head_in, 0, 1, (int)Mo, Hout, head_ld);
      }
      const int kb = p.nchunk * ndt1_cdiv(p.chunk_k, 64);
      p.split_k = kb >= 32 ? 16 : 1;
      if (!(kBf16 && !force_simt) && p.split_k > 8) p.split_k = 8;
      if (kBf16 && !force_simt) p.split_k = 0;
      NDT1_TRY(run(p, ws));
    }
    {
      GemmProblem p = prob(GEMM_NN, Tp, Hout, V);
      p.nb_out = B;
      p.A = op(dlog, (long long)Tp * ldV, B, Tp, V, ldV);
      p.B = op(W(P->dec_w, u_dec), 0, 1, V, Hout, Hout);
      p.epi.out = d_head_in + (long long)n_prefix * head_ld; p.epi.out_bf16 = kBf16; p.epi.ldc = head_ld; p.epi.c_batch_stride = (long long)L * head_ld;
      if (k.factors_active) {
        // through the factors activation: needs the pre-activation for gelu, the output otherwise
        if (k.factors_act == NDT1_ACT_GELU) { p.epi.dact = DACT_GELU_FROM_IN; p.epi.dact_in = fpre; }
        else { p.epi.dact = dact_from_out(k.factors_act); p.epi.dact_in = fac; }
        p.epi.dact_in_bf16 = kBf16;
      }
      NDT1_TRY(run(p, s));
    }
    }   // (full backward)
    if (k.factors_active) {
      NDT1_TRY(fork());
      if (G->factors_b && k.factors_bias) NDT1_TRY(k_colsum<T>(dfac, G->factors_b, M, Hout, Hout, ws));
      NDT1_TRY(linear_wgrad(dfac, Hout, hn, H, G->factors_w, H, (int)M, Hout, H, ws));
      GemmEpilogue e = gemm_epilogue_default();
      e.out = dhn; e.out_bf16 = kBf16; e.ldc = H;
      NDT1_TRY(linear_dgrad(dfac, Hout, W(P->factors_w, u_fac), H, (int)M, Hout, H, e, s));
    }
    NDT1_TRY(k_dropout_inplace<T>(dhn, M * H, training ? k.p_factors : 0.f, seed, site_factors(), s));
    // out_norm
    NDT1_CUDA_CHECK(cudaMemsetAsync(dX, 0, M * H * sizeof(float), s));
    // every LayerNorm backward also emits the column sums of the operand it hands to the next GEMM pair = that layer's bias gradient
    NDT1_TRY(k_layernorm_bwd<T>(dhn, xs[2 * NL], P->out_norm_w, mean[2 * NL], rstd[2 * NL], dX, G->out_norm_w, G->out_norm_b, dYm[NL - 1], ptr_, seed,
                                site_mlp(NL - 1), M, H, ln_part, s, k.mlp_bias ? G->layer[NL - 1].down_b : nullptr));
    NDT1_TRY(fork());
    NDT1_CUDA_CHECK(record_stage(0, ws));   // decoder + out_norm gradients complete
    int cf, cb; ctx(cf, cb);
    for (int l = NL - 1; l >= 0; --l) {
      const auto& q = P->layer[l]; const auto& gq = G->layer[l];
      // MLP: x_out = x_mid + drop(down(act(up(h2))))      dY = T(dX * mlp mask)
      T* dY = dYm[l]; T* dU = dUl[l]; T* dqkv = dqkvl[l];
      NDT1_TRY(linear_wgrad(dY, H, g[l], I, gq.down_w, I, (int)M, H, I, ws));      // (ws is already ordered after the LayerNorm backward that wrote dY)
      {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = dU; e.out_bf16 = kBf16; e.ldc = I;
        if (k.mlp_act == NDT1_ACT_GELU) { e.dact = DACT_SAVED; e.dact_in = u[l]; }
        else { e.dact = dact_from_out(k.mlp_act); e.dact_in = g[l]; }
        e.dact_in_bf16 = kBf16;
        // the up-projection's bias gradient = column sums of dU: fused into this GEMM's epilogue (a separate reduction on the second
        // stream costs 12 us of SM time per layer that the overlapped weight-gradient GEMMs need: measured ~20 us per step)
        static const bool separate_cs = getenv("NDT1_FUSE_COLSUM") && getenv("NDT1_FUSE_COLSUM")[0] == '0';
        const bool fuse_cs = kBf16 && !force_simt && (!overlap || !separate_cs) && gq.up_b && k.mlp_bias && I % 8 == 0;
        if (fuse_cs) e.colsum = gq.up_b;
        NDT1_TRY(linear_dgrad(dY, H, W(q.down_w, kBf16 ? u_down[l] : nullptr), I, (int)M, H, I, e, s));
        NDT1_TRY(fork());
        if (!fuse_cs && gq.up_b && k.mlp_bias) NDT1_TRY(k_colsum<T>(dU, gq.up_b, M, I, I, ws));
      }
      NDT1_TRY(linear_wgrad(dU, I, h2[l], H, gq.up_w, H, (int)M, I, H, ws));
      {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = dH; e.out_bf16 = kBf16; e.ldc = H;
        NDT1_TRY(linear_dgrad(dU, I, W(q.up_w, kBf16 ? u_up[l] : nullptr), H, (int)M, I, H, e, s));
      }
      NDT1_TRY(k_layernorm_bwd<T>(dH, xs[2 * l + 1], q.ln2_w, mean[2 * l + 1], rstd[2 * l + 1], dX, gq.ln2_w, gq.ln2_b, dYa[l], 0.f, seed, 0, M, H,
                                  ln_part, s, k.attention_bias ? gq.o_b : nullptr));
      dY = dYa[l];
      NDT1_TRY(fork());
      // attention block: x_mid = x_in + out_proj(drop(att))      dY = T(dX)
      const T* ad = (ptr_ > 0.f) ? attd[l] : att[l];
      NDT1_TRY(linear_wgrad(dY, H, ad, H, gq.o_w, H, (int)M, H, H, ws));
      {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = dA; e.out_bf16 = kBf16; e.ldc = H;
        if (ptr_ > 0.f) { e.drop_p = ptr_; e.drop_bwd = 1; e.drop_seed = seed; e.drop_stream = site_attn_o(l); if (bits_drawn) e.drop_bits = dropbits_o[l]; }
        NDT1_TRY(linear_dgrad(dY, H, W(q.o_w, kBf16 ? u_o[l] : nullptr), H, (int)M, H, H, e, s));
      }
      AttnParams ap;
      ap.qkv = qkv[l]; ap.out = att[l]; ap.out_drop = (void*)ad; ap.lse = lse[l]; ap.key_valid = key_valid;
      ap.B = B; ap.L = L; ap.H = H; ap.nh = k.n_heads; ap.hd = H / k.n_heads; ap.ctx_fwd = cf; ap.ctx_bwd = cb;
      ap.scale = 1.0f / sqrtf((float)(H / k.n_heads)); ap.p_attn = ptr_; ap.p_out = ptr_;
      ap.seed = seed; ap.stream_attn = site_attn_p(l); ap.stream_out = site_attn_o(l);
      ap.dout = dA; ap.dqkv = dqkv; ap.delta = delta; ap.drop_bits = dropbits[l]; ap.drop_bits_o = dropbits_o[l]; ap.bits_ready = 1;
      if (kBf16 && !force_simt && !simt_attention && k_attention_tc_supported(ap)) NDT1_TRY(k_attention_tc_bwd(ap, s));
      else NDT1_TRY(k_attention_bwd<T>(ap, s));
      if (k.use_rope) NDT1_TRY(k_rope<T>(dqkv, ts_ptr, Tn, rope_cos, rope_sin, M, L, H, k.n_heads, k.max_F, 1, s));   // transpose of the rotation
      NDT1_TRY(fork());
      float* gw[3] = {gq.q_w, gq.k_w, gq.v_w}; float* gb[3] = {gq.q_b, gq.k_b, gq.v_b};
      // flat gradient arena: q|k|v weights (and biases) adjacent -> one (3H x H) weight gradient, one bias reduction
      if (gb[0] && gb[1] == gb[0] + H && gb[2] == gb[1] + H && k.attention_bias) {
        NDT1_TRY(k_colsum<T>(dqkv, gb[0], M, 3 * H, 3 * H, ws));
      } else {
        for (int j = 0; j < 3; ++j)
          if (gb[j] && k.attention_bias) NDT1_TRY(k_colsum<T>(dqkv + (long long)j * H, gb[j], M, H, 3 * H, ws));
      }
      if (gw[0] && gw[1] == gw[0] + (long long)H * H && gw[2] == gw[1] + (long long)H * H) {
        NDT1_TRY(linear_wgrad(dqkv, 3 * H, h1[l], H, gw[0], H, (int)M, 3 * H, H, ws));
      } else {
        for (int j = 0; j < 3; ++j) NDT1_TRY(linear_wgrad(dqkv + (long long)j * H, 3 * H, h1[l], H, gw[j], H, (int)M, H, H, ws));
      }
      if (kBf16) {
        GemmEpilogue e = gemm_epilogue_default();
        e.out = dH; e.out_bf16 = 1; e.ldc = H;
        NDT1_TRY(linear_dgrad(dqkv, 3 * H, u_qkv[l], H, (int)M, 3 * H, H, e, s));
      } else {
        const float* wq[3] = {q.q_w, q.k_w, q.v_w};
        for (int j = 0; j < 3; ++j) {
          GemmEpilogue e = gemm_epilogue_default();
          e.out = dH; e.out_bf16 = 0; e.ldc = H; e.accumulate = j > 0;
          NDT1_TRY(linear_dgrad(dqkv + (long long)j * H, 3 * H, wq[j], H, (int)M, H, H, e, s));
        }
      }
      const bool first = (l == 0);
      NDT1_TRY(k_layernorm_bwd<T>(dH, xs[2 * l], q.ln1_w, mean[2 * l], rstd[2 * l], dX, gq.ln1_w, gq.ln1_b, first ? (T*)nullptr : dYm[first ? 0 : l - 1],
                                  first ? 0.f : ptr_, seed, first ? 0 : site_mlp(l - 1), M, H, ln_part, s,
                                  (!first && k.mlp_bias) ? G->layer[l - 1].down_b : nullptr));
      NDT1_TRY(fork());                                          // also orders ws after this LayerNorm (its affine gradients, next dY)
      NDT1_CUDA_CHECK(record_stage(NL - l, ws));   // layer l gradients complete
    }
    // embedding: dX is the gradient w.r.t. the (dropped) embedding output.
    // One pass: apply the embedding dropout mask, cast for the GEMMs, scatter into the position table.
    T* dY = dYe;
    NDT1_TRY(k_grad_prep<T>(dX, dY, M, H, pe, seed, 0, (k.pos && G->pos_w) ? G->pos_w : nullptr, ts_ptr, H, L, Tn, n_prefix, s));
    if (n_prefix > 0) {
      int slot = 0;
      if (k.day_token) { if (G->day_emb) NDT1_TRY(k_token_rows_grad<T>(G->day_emb, day_ptr, dY, B, L, H, slot, s)); ++slot; }
      if (k.block_token) { if (G->block_emb) NDT1_TRY(k_token_rows_grad<T>(G->block_emb, block_ptr, dY, B, L, H, slot, s)); ++slot; }
      NDT1_CUDA_CHECK(cudaMemset2DAsync(dY, (size_t)L * H * sizeof(T), 0, (size_t)n_prefix * H * sizeof(T), B, s));
    }
    NDT1_TRY(fork());
    if (G->proj_b) NDT1_TRY(k_colsum<T>(dY, G->proj_b, M, H, H, ws));
    const T* dE = dY + (long long)n_prefix * H;
    const int eact = k.embed_act;
    // (not in the stacked layout: there one output row of the overlap-add GEMM holds `stride` bins)
    const bool fuse_embed_cs = kBf16 && !force_simt && G->embed_b && k.embed_bias && D % 8 == 0 && !k.stack_active && !k.adapt;
    // derivative of the embedder activation: from the kept pre-activation for GELU, from the output otherwise
    const int e_dact = eact == NDT1_ACT_GELU ? DACT_GELU_FROM_IN : dact_from_out(eact);
    const T* e_dact_in = eact == NDT1_ACT_GELU ? emb_pre : emb;
    if (k.stack_active) {
      const int K4 = k.stack_stride * D, nch = k.stack_size / k.stack_stride, R4 = Tn / k.stack_stride;
      if (G->proj_w) {
        GemmProblem p = prob(GEMM_TN, H, nch * K4, Tp);
        p.nchunk = B; p.b_chunk_n = K4; p.b_row_shift = 1;
        p.A = op(dE, (long long)L * H, B, Tp, H, H);
        p.B = op(emb, (long long)Tn * D, B, R4, K4, K4);
        p.epi.out = G->proj_w; p.epi.ldc = nch * K4; p.epi.accumulate = 1;
        p.split_k = (kBf16 && !force_simt) ? 0 : 1;
        NDT1_TRY(run(p, ws));
      }
      if (Tn % k.stack_stride != 0) NDT1_CUDA_CHECK(cudaMemsetAsync(dEmb, 0, MT * D * sizeof(T), s));
      GemmProblem p = prob(GEMM_NN, R4, K4, H);
      p.nb_out = B; p.nchunk = nch; p.a_row_shift = -1; p.b_col_shift = K4;
      p.A = op(dE, (long long)L * H, B, Tp, H, H);
      p.B = op(W(P->proj_w, u_proj), 0, 1, H, nch * K4, nch * K4);
      p.epi.out = dEmb; p.epi.out_bf16 = kBf16; p.epi.ldc = K4; p.epi.c_batch_stride = (long long)Tn * D;
      p.epi.dact = e_dact; p.epi.dact_in = e_dact_in; p.epi.dact_in_bf16 = kBf16;
      if (fuse_embed_cs) p.epi.colsum = G->embed_b;
      NDT1_TRY(run(p, s));
    } else {
      if (G->proj_w) {
        GemmProblem p = prob(GEMM_TN, H, D, Tp);
        p.nchunk = B;
        p.A = op(dE, (long long)L * H, B, Tp, H, H);
        p.B = op(emb, (long long)Tn * D, B, Tn, D, D);
        p.epi.out = G->proj_w; p.epi.ldc = D; p.epi.accumulate = 1;
        NDT1_TRY(run(p, ws));
      }
      GemmProblem p = prob(GEMM_NN, Tp, D, H);
      p.nb_out = B;
      p.A = op(dE, (long long)L * H, B, Tp, H, H);
      p.B = op(W(P->proj_w, u_proj), 0, 1, H, D, D);
      p.epi.out = dEmb; p.epi.out_bf16 = kBf16; p.epi.ldc = D; p.epi.c_batch_stride = (long long)Tn * D;
      p.epi.dact = e_dact; p.epi.dact_in = e_dact_in; p.epi.dact_in_bf16 = kBf16;
      if (fuse_embed_cs) p.epi.colsum = G->embed_b;
      NDT1_TRY(run(p, s));
    }
    NDT1_TRY(fork());
    if (k.adapt) {
      // per-day gradients: trial b adds into day_idx[b]'s packed matrix / bias row, then each day's sum goes to its own tensor
      const T* x_in = kBf16 ? (xin_ext ? xin_ext : xin) : (const T*)spikes_ptr;
      const int ldx = kBf16 ? ldN : N;
      NDT1_CUDA_CHECK(cudaMemsetAsync(dwd_pack, 0, (size_t)k.n_days * D * N * 4, ws));
      NDT1_CUDA_CHECK(cudaMemsetAsync(dbd_pack, 0, (size_t)k.n_days * D * 4, ws));
      if (k.embed_bias) NDT1_TRY(k_colsum_sel<T>(dEmb, dbd_pack, day_ptr, k.n_days, D, B, Tn, D, ws));
      GemmProblem p = prob(GEMM_TN, D, N, Tn);
      p.nchunk = B; p.split_k = B;
      p.A = op(dEmb, (long long)Tn * D, B, Tn, D, D); p.B = op(x_in, (long long)Tn * ldx, B, Tn, N, ldx);
      p.epi.out = dwd_pack; p.epi.ldc = N; p.epi.accumulate = 1;
      p.epi.sel = day_ptr; p.epi.sel_n = k.n_days; p.epi.c_sel_stride = (long long)D * N;
      NDT1_TRY(run(p, ws));
      for (int d = 0; d < k.n_days; ++d) {
        if (G->embed_w_day[d]) NDT1_TRY(k_add_inplace(G->embed_w_day[d], dwd_pack + (long long)d * D * N, (long long)D * N, ws));
        if (G->embed_b_day[d] && k.embed_bias) NDT1_TRY(k_add_inplace(G->embed_b_day[d], dbd_pack + (long long)d * D, D, ws));
      }
    }
    if (!fuse_embed_cs && G->embed_b && k.embed_bias && !k.adapt) NDT1_TRY(k_colsum<T>(dEmb, G->embed_b, MT, D, D, ws));
    if (G->embed_w && !k.adapt) {
      const T* x_in = kBf16 ? (xin_ext ? xin_ext : xin) : (const T*)spikes_ptr;
      const int ldx = kBf16 ? ldN : N;
      GemmProblem p = prob(GEMM_TN, D, N, (int)MT);
      p.A = op(dEmb, 0, 1, (int)MT, D, D); p.B = op(x_in, 0, 1, (int)MT, N, ldx);
      p.epi.out = G->embed_w; p.epi.ldc = N; p.epi.accumulate = 1;
      const int kb = ndt1_cdiv(MT, 64);
      int split = kb / 8; if (split > 64) split = 64; if (split < 1) split = 1;
      if (!(kBf16 && !force_simt) && split > 8) split = 8;
      p.split_k = (kBf16 && !force_simt) ? 0 : split;
      NDT1_TRY(run(p, ws));
    }
    NDT1_CUDA_CHECK(record_stage(NL + 1, ws));     // embedding gradients complete
    if (overlap) {                                               // join: the caller's stream sees the whole backward
      NDT1_CUDA_CHECK(cudaEventRecord(join_ev, ws));
      NDT1_CUDA_CHECK(cudaStreamWaitEvent(s, join_ev, 0));
    }
    launches += g_ndt1_launches - launches0;
    return 0;
  }
};

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int ndt1_abi_version(void) { return NDT1_ABI_VERSION; }

int ndt1_engine_create(const ndt1_config* cfg, ndt1_engine** out) {
  NDT1_REQUIRE(cfg && out, "engine_create: null argument");
  NDT1_REQUIRE(cfg->abi_version == NDT1_ABI_VERSION, "engine_create: ABI version %d, library is %d", cfg->abi_version, NDT1_ABI_VERSION);
  NDT1_REQUIRE(cfg->n_layers >= 1 && cfg->n_layers <= NDT1_MAX_LAYERS, "engine_create: n_layers %d outside [1,%d]", cfg->n_layers, NDT1_MAX_LAYERS);
  NDT1_REQUIRE(cfg->hidden % cfg->n_heads == 0, "engine_create: Hidden dim is not multiple of head size");
  NDT1_REQUIRE(!cfg->adapt || (cfg->n_days >= 1 && cfg->n_days <= NDT1_MAX_DAYS), "engine_create: adapt needs 1 <= n_days <= %d (got %d)", NDT1_MAX_DAYS, cfg->n_days);
  NDT1_REQUIRE(!cfg->use_rope || ((cfg->hidden / cfg->n_heads) % 2 == 0 && cfg->max_F > 0), "engine_create: use_rope needs an even head size and max_F > 0");
  NDT1_REQUIRE(cfg->max_batch > 0 && cfg->max_T > 0, "engine_create: max_batch / max_T must be positive");
  NDT1_REQUIRE(cfg->hidden % 8 == 0 && cfg->inter % 8 == 0 && cfg->input_dim % 8 == 0, "engine_create: hidden, inter and input_dim must be multiples of 8");
  ndt1_engine* e = nullptr;
  int rc;
  if (cfg->precision == NDT1_PRECISION_BF16) { auto* t = new Engine<bf16>(); t->c = *cfg; rc = t->init(); e = t; }
  else if (cfg->precision == NDT1_PRECISION_FP32) { auto* t = new Engine<float>(); t->c = *cfg; rc = t->init(); e = t; }
  else { NDT1_REQUIRE(false, "engine_create: unknown precision %d", cfg->precision); }
  if (rc) { delete e; return rc; }
  *out = e;
  return 0;
}
void ndt1_engine_destroy(ndt1_engine* e) { delete e; }
size_t ndt1_engine_arena_bytes(const ndt1_engine* e) { return e->arena_bytes(); }
int ndt1_engine_out_len(const ndt1_engine* e, int T) { return e->out_len(T); }
int ndt1_engine_forward(ndt1_engine* e, const ndt1_tensors* params, const ndt1_batch* batch, const ndt1_outputs* out, void* stream) {
  NDT1_REQUIRE(e && params && batch && out && out->loss, "engine_forward: null argument");
  return e->forward(params, batch, out, (cudaStream_t)stream);
}
int ndt1_engine_backward(ndt1_engine* e, const ndt1_tensors* params, const ndt1_tensors* grads, const float* dloss, void* stream) {
  NDT1_REQUIRE(e && params && grads, "engine_backward: null argument");
  return e->backward(params, grads, dloss, (cudaStream_t)stream);
}
int ndt1_engine_backward_features(ndt1_engine* e, const ndt1_tensors* params, const ndt1_tensors* grads, const float* dfeatures, void* stream) {
  NDT1_REQUIRE(e && params && grads && dfeatures, "engine_backward_features: null argument");
  return e->backward(params, grads, nullptr, (cudaStream_t)stream, dfeatures);
}
int64_t ndt1_engine_launch_count(const ndt1_engine* e) { return e->launches; }
int ndt1_engine_set_weight_shadow(ndt1_engine* e, const float* params_fp32, const void* shadow_bf16, int64_t n) {
  NDT1_REQUIRE(e, "engine_set_weight_shadow: null engine");
  NDT1_REQUIRE(!shadow_bf16 || (params_fp32 && n > 0), "engine_set_weight_shadow: a shadow needs the arena it mirrors");
  e->shadow_src = shadow_bf16 ? params_fp32 : nullptr; e->shadow_bf16 = shadow_bf16; e->shadow_n = shadow_bf16 ? n : 0;
  return 0;
}
int ndt1_engine_set_overlap(ndt1_engine* e, int on) {
  NDT1_REQUIRE(e, "engine_set_overlap: null engine");
  e->overlap = on != 0;
  return 0;
}
int ndt1_engine_set_rope_tables(ndt1_engine* e, const float* cos_table, const float* sin_table, int rows) {
  NDT1_REQUIRE(e, "engine_set_rope_tables: null engine");
  return e->set_rope_tables(cos_table, sin_table, rows);
}
int ndt1_engine_stage_count(const ndt1_engine* e) { return e->n_stages(); }
int ndt1_engine_wait_stage(ndt1_engine* e, int stage, void* stream) {
  NDT1_REQUIRE(e && stage >= 0 && stage < e->n_stages(), "engine_wait_stage: stage %d out of range", stage);
  NDT1_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, e->stage_ev[stage], 0));
  return 0;
}

}  // extern "C"
