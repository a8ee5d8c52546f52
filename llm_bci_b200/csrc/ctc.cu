// Phoneme head loss: LogSoftmax + CTC forward/backward, and greedy decode.
// Reference: nn.LogSoftmax + nn.CTCLoss(reduction="none", blank, zero_infinity)
// at models/ndt1.py:493-500,517,581 (conventions: SURVEY.md A.6) and the
// greedy collapse of utils/eval_bci.py:41-48.
//
// One CTA per trial; one thread per state of the extended label sequence
// (blank, l1, blank, l2, ..., blank) in each of two groups: alpha is swept
// forward and beta backward AT THE SAME TIME (L dependent steps, not 2 L), then
// the gradient w.r.t. the LOGITS,
//     dlogits[t,c] = (softmax[t,c] - posterior[t,c]) * dloss      (t <  len)
//                  = 0                                            (t >= len)
// is emitted by a parallel pass, so no separate log-softmax backward exists.
// The kernel is latency-, not bandwidth-bound.
#include "kernels.cuh"

namespace {

__global__ void log_softmax_kernel(const float* __restrict__ logits, float* __restrict__ logp, long long rows, int V) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = logits + r * V;
  float m = -INFINITY;
  for (int c = lane; c < V; c += 32) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(x[c] - m);
  s = warp_sum(s);
  const float lz = m + logf(s);
  for (int c = lane; c < V; c += 32) logp[r * V + c] = x[c] - lz;
}

struct CtcParams {
  const float* logp; const long long* targets; const long long* in_len; const long long* tgt_len;
  int B, L, V, S, blank, zero_infinity, lp_in_smem, LXP;
  float* alpha; float* beta; float* nll; float* dlogits; const float* dloss;
};

constexpr int kRenorm = 8;   // re-centre the alpha/beta rows every kRenorm frames

// FAST: MUFU ex2/lg2 (bf16 engine mode); otherwise libm-accurate expf/logf (strict fp32 mode, stand-alone operator)
template <bool FAST> __device__ __forceinline__ float exp_t(float x) { return FAST ? __expf(x) : expf(x); }
template <bool FAST> __device__ __forceinline__ float log_t(float x) { return FAST ? __logf(x) : logf(x); }
template <bool FAST>
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + log_t<FAST>(exp_t<FAST>(a - m) + exp_t<FAST>(b - m) + exp_t<FAST>(c - m));
}

// One CTA per trial.  The alpha sweep (threads [0, LXP), one per state of the extended label sequence) and the
// beta sweep (threads [LXP, 2 LXP)) run CONCURRENTLY, one frame per barrier, so the sequential depth is L frames
// instead of 2 L; both keep their rows in shared memory and stream the re-centred rows to a workspace.  A final,
// fully parallel pass (one warp per frame) turns alpha + beta into label posteriors and writes the gradient
// w.r.t. the logits.  Rows are kept re-centred (alpha_hat = alpha - C_t, beta_hat = beta - D_t, offsets in
// double) so the fp32 log-space values stay O(10) instead of O(loss): posteriors keep ~1e-6 relative accuracy
// where plain fp32 log-space CTC (torch's kernel included) loses ~3e-5 on a 250-frame utterance.
template <bool FAST>
__global__ void ctc_kernel(const CtcParams p) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, nwarps = blockDim.x >> 5, warp = tid >> 5, lane = tid & 31;
  const int LX = 2 * p.S + 1, LXP = p.LXP;
  double* Cs = (double*)sm_raw;                     // [L] forward offsets
  double* Ds = Cs + p.L;                            // [L] backward offsets
  float* rowA = (float*)(Ds + p.L);                 // [2][LX + 4]  (2 -inf pads on the left)
  float* rowB = rowA + 2 * (LX + 4);                // [2][LX + 4]  (2 -inf pads on the right)
  float* red = rowB + 2 * (LX + 4);                 // [32]
  float* post = red + 32;                           // [nwarps][V]
  int* lab = (int*)(post + nwarps * p.V);           // [LX]
  float* lp_s = (float*)(lab + LX);                 // [L*V] when it fits
  __shared__ double s_ll;

  int S = (int)p.tgt_len[b];
  S = S < 0 ? 0 : (S > p.S ? p.S : S);
  const int Lx = 2 * S + 1;
  const long long tn_ll = p.in_len[b];
  const int Tn = tn_ll > p.L ? p.L : (tn_ll < 0 ? 0 : (int)tn_ll);
  const float* lp_g = p.logp + (long long)b * p.L * p.V;
  const float* lp = p.lp_in_smem ? lp_s : lp_g;
  float* alpha = p.alpha + (long long)b * p.L * LX;
  float* beta = p.beta + (long long)b * p.L * LX;
  float* dl = p.dlogits ? p.dlogits + (long long)b * p.L * p.V : nullptr;

  if (p.lp_in_smem) for (int i = tid; i < p.L * p.V; i += blockDim.x) lp_s[i] = lp_g[i];
  for (int i = tid; i < 4 * (LX + 4); i += blockDim.x) rowA[i] = -INFINITY;
  for (int i = tid; i < LX; i += blockDim.x) lab[i] = (i < Lx && (i & 1)) ? (int)p.targets[(long long)b * p.S + (i >> 1)] : p.blank;
  __syncthreads();

  const bool fwd = tid < LXP;                       // alpha group / beta group
  const int s = fwd ? tid : tid - LXP;
  const bool live = s < Lx;
  const int my = live ? lab[s] : p.blank;
  const bool skip_b = live && s >= 2 && my != p.blank && my != lab[s - 2];                   // alpha: s-2 -> s allowed
  const bool skip_f = live && (s + 2 < Lx) && lab[s + 2] != p.blank && lab[s + 2] != my;     // beta:  s -> s+2 allowed
  const int gw0 = fwd ? 0 : LXP / 32, gw1 = fwd ? LXP / 32 : nwarps;                         // this group's warps

  double ll = -INFINITY;
  if (Tn > 0) {
    float* prev = fwd ? rowA + 2 : rowB;
    float* cur = prev + (LX + 4);
    double off = 0.0;
    for (int k = 0; k < Tn; ++k) {
      const int t = fwd ? k : Tn - 1 - k;
      float v = -INFINITY;
      if (live) {
        const float e = lp[t * p.V + my];
        if (k == 0) {
          v = fwd ? (s < 2 ? e : -INFINITY) : (s >= Lx - 2 ? e : -INFINITY);
        } else {
          const float x0 = prev[s];
          const float x1 = fwd ? prev[s - 1] : prev[s + 1];
          const float x2 = fwd ? (skip_b ? prev[s - 2] : -INFINITY) : (skip_f ? prev[s + 2] : -INFINITY);
          v = lse3<FAST>(x0, x1, x2);
          v = (v == -INFINITY) ? -INFINITY : v + e;
        }
      }
      if ((k % kRenorm) == kRenorm - 1) {
        const float wm = warp_max(v);
        if (lane == 0) red[warp] = wm;
        __syncthreads();
        float mx = -INFINITY;
        for (int w = gw0; w < gw1; ++w) mx = fmaxf(mx, red[w]);
        if (mx != -INFINITY) { v = (v == -INFINITY) ? v : v - mx; off += (double)mx; }
      }
      if (live) { cur[s] = v; (fwd ? alpha : beta)[(long long)t * LX + s] = v; }
      if (s == 0) (fwd ? Cs : Ds)[t] = off;
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
    if (tid == 0) {
      float v = prev[Lx - 1];
      if (Lx > 1) v = lse3<false>(v, prev[Lx - 2], -INFINITY);
      s_ll = (v == -INFINITY) ? -INFINITY : off + (double)v;
    }
    __syncthreads();
    ll = s_ll;
  } else if (S == 0) {
    ll = 0.0;
  }
  const bool feasible = (ll != -INFINITY) && (ll == ll);
  if (tid == 0) p.nll[b] = feasible ? (float)(-ll) : (p.zero_infinity ? 0.f : INFINITY);
  if (!dl) return;
  const float gs = p.dloss ? *p.dloss : 1.f;
  const int t_zero_from = (feasible && Tn > 0) ? Tn : 0;
  for (long long e = (long long)t_zero_from * p.V + tid; e < (long long)p.L * p.V; e += blockDim.x) dl[e] = 0.f;
  if (!feasible || Tn <= 0) return;

  // posteriors: one warp per frame.  dlogits[t,c] = (softmax[t,c] - sum_{s: l'(s)=c} alpha beta / (p_t(c) P)) * dloss
  float* pw = post + warp * p.V;
  for (int t = warp; t < Tn; t += nwarps) {
    for (int c = lane; c < p.V; c += 32) pw[c] = 0.f;
    __syncwarp();
    const float kt = (float)(Cs[t] + Ds[t] - ll);
    for (int s2 = lane; s2 < Lx; s2 += 32) {
      const float al = alpha[(long long)t * LX + s2], be = beta[(long long)t * LX + s2];
      if (al != -INFINITY && be != -INFINITY) {
        const int c = lab[s2];
        atomicAdd(&pw[c], exp_t<FAST>(al + be - lp[(long long)t * p.V + c] + kt));
      }
    }
    __syncwarp();
    for (int c = lane; c < p.V; c += 32) dl[(long long)t * p.V + c] = (exp_t<FAST>(lp[(long long)t * p.V + c]) - pw[c]) * gs;
    __syncwarp();
  }
}

__global__ void sum_nll_kernel(const float* nll, int B, float* loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += nll[b];
    *loss += s;
  }
}

// argmax over V then the reference's collapse: emit when id != last EMITTED id and id != blank.
__global__ void greedy_kernel(const float* logp, int B, int L, int V, int blank, long long* out_ids, long long* out_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long last = -1; int n = 0;
  for (int t = 0; t < L; ++t) {
    const float* x = logp + ((long long)b * L + t) * V;
    int best = 0; float bv = x[0];
    for (int c = 1; c < V; ++c) if (x[c] > bv) { bv = x[c]; best = c; }
    if (best != last && best != blank) { out_ids[(long long)b * L + n] = best; ++n; last = best; }
  }
  for (int t = n; t < L; ++t) out_ids[(long long)b * L + t] = -1;
  out_len[b] = n;
}

}  // namespace

int k_log_softmax(const float* logits, float* logp, long long rows, int V, cudaStream_t stream) {
  if (rows == 0) return 0;
  log_softmax_kernel<<<ndt1_cdiv(rows, 8), 256, 0, stream>>>(logits, logp, rows, V);
  NDT1_CHECK_LAUNCH();
  return 0;
}

size_t k_ctc_workspace_floats(int B, int L, int S) { return 2 * (size_t)B * L * (2 * S + 1) + B; }

int k_ctc_fwd_bwd(const float* logp, const long long* targets, const long long* in_len, const long long* tgt_len, int B, int L, int V,
                  int S, int blank, int zero_infinity, float* alpha_ws, float* nll, float* loss, float* dlogits, const float* dloss,
                  cudaStream_t stream, int fast_math) {
  if (B == 0) return 0;
  NDT1_REQUIRE(blank >= 0 && blank < V, "ctc: blank id %d outside the vocabulary (%d)", blank, V);
  const int LX = 2 * S + 1;
  const int LXP = ((LX + 31) / 32) * 32;
  const int threads = 2 * LXP;
  NDT1_REQUIRE(threads <= 1024, "ctc: target length %d too long for one CTA (max 255 labels)", S);
  const int nwarps = threads / 32;
  const size_t base = (size_t)2 * L * sizeof(double) + (size_t)(4 * (LX + 4) + 32 + nwarps * V + LX) * sizeof(float);
  const size_t with_lp = base + (size_t)L * V * sizeof(float);
  const int lp_in_smem = with_lp <= 200 * 1024;
  const size_t smem = lp_in_smem ? with_lp : base;
  NDT1_REQUIRE(smem <= 200 * 1024, "ctc: %d frames x %d labels do not fit one CTA", L, S);
  float* beta_ws = alpha_ws + (size_t)B * L * LX;
  CtcParams p{logp, targets, in_len, tgt_len, B, L, V, S, blank, zero_infinity, lp_in_smem, LXP, alpha_ws, beta_ws, nll, dlogits, dloss};
  static size_t attr = 0;
  if (smem > attr) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(ctc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(ctc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  if (fast_math) ctc_kernel<true><<<B, threads, smem, stream>>>(p);
  else ctc_kernel<false><<<B, threads, smem, stream>>>(p);
  NDT1_CHECK_LAUNCH();
  if (loss) {
    sum_nll_kernel<<<1, 32, 0, stream>>>(nll, B, loss);
    NDT1_CHECK_LAUNCH();
  }
  return 0;
}

int k_ctc_greedy_decode(const float* logp, int B, int L, int V, int blank, long long* out_ids, long long* out_len, cudaStream_t stream) {
  if (B == 0) return 0;
  greedy_kernel<<<ndt1_cdiv(B, 64), 64, 0, stream>>>(logp, B, L, V, blank, out_ids, out_len);
  NDT1_CHECK_LAUNCH();
  return 0;
}
