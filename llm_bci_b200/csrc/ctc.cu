// Phoneme head loss: LogSoftmax + CTC forward/backward, and greedy decode.
// Reference: nn.LogSoftmax + nn.CTCLoss(reduction="none", blank, zero_infinity)
// at models/ndt1.py:493-500,517,581 (conventions: SURVEY.md A.6) and the
// greedy collapse of utils/eval_bci.py:41-48.
//
// One CTA per trial; thread s owns state s of the extended label sequence
// (blank, l1, blank, l2, ..., blank).  alpha is swept forward and kept in a
// workspace; beta is swept backward and the gradient w.r.t. the LOGITS,
//     dlogits[t,c] = (softmax[t,c] - posterior[t,c]) * dloss      (t <  len)
//                  = 0                                            (t >= len)
// is emitted directly, so no separate log-softmax backward pass exists.
// The next step's emission is prefetched before each barrier: the kernel is
// latency-, not bandwidth-bound (2*L dependent steps).
#include "kernels.cuh"

namespace {

__device__ __forceinline__ float lse2(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

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
  int B, L, V, S, blank, zero_infinity, lp_in_smem;
  float* alpha; float* nll; float* dlogits; const float* dloss;
};

constexpr int kRenorm = 8;   // re-centre the alpha/beta rows every kRenorm frames

__device__ __forceinline__ float block_max(float v, float* red, int nwarps) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = -INFINITY;
  for (int w = 0; w < nwarps; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  return m;
}

// One CTA per trial, one thread per state of the extended label sequence.
// Rows are kept re-centred (alpha_hat = alpha - C_t, beta_hat = beta - D_t with the offsets in double),
// so the fp32 log-space values stay O(10) instead of O(loss): posteriors keep ~1e-6 relative accuracy
// where plain fp32 log-space CTC (torch's kernel included) loses ~3e-5 on a 250-frame utterance.
__global__ void ctc_kernel(const CtcParams p) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, nwarps = blockDim.x >> 5;
  const int LX = 2 * p.S + 1;
  double* Cs = (double*)sm_raw;                     // [L] forward offsets
  float* row0 = (float*)(Cs + p.L);                 // [LX + 4] (2 pads each side)
  float* row1 = row0 + LX + 4;
  float* post = row1 + LX + 4;                      // [2][V]
  float* red = post + 2 * p.V;                      // [32]
  float* lp_s = red + 32;                           // [L*V] when it fits
  __shared__ double s_ll;

  int S = (int)p.tgt_len[b];
  S = S < 0 ? 0 : (S > p.S ? p.S : S);
  const int Lx = 2 * S + 1;
  const long long tn_ll = p.in_len[b];
  const int Tn = tn_ll > p.L ? p.L : (int)tn_ll;
  const float* lp_g = p.logp + (long long)b * p.L * p.V;
  const float* lp = p.lp_in_smem ? lp_s : lp_g;
  float* alpha = p.alpha + (long long)b * p.L * LX;
  float* dl = p.dlogits ? p.dlogits + (long long)b * p.L * p.V : nullptr;

  if (p.lp_in_smem) for (int i = tid; i < p.L * p.V; i += blockDim.x) lp_s[i] = lp_g[i];
  for (int i = tid; i < 2 * (LX + 4); i += blockDim.x) row0[i] = -INFINITY;
  for (int i = tid; i < 2 * p.V; i += blockDim.x) post[i] = 0.f;
  const int s = tid;
  const bool live = s < Lx;
  const int my = live ? ((s & 1) ? (int)p.targets[(long long)b * p.S + (s >> 1)] : p.blank) : p.blank;
  const int my_m2 = (live && s >= 2) ? ((s & 1) ? (int)p.targets[(long long)b * p.S + ((s - 2) >> 1)] : p.blank) : -1;
  const int my_p2 = (s + 2 < Lx) ? ((s & 1) ? (int)p.targets[(long long)b * p.S + ((s + 2) >> 1)] : p.blank) : -1;
  const bool skip_b = live && s >= 2 && my != p.blank && my != my_m2;       // alpha: s-2 -> s allowed
  const bool skip_f = (s + 2 < Lx) && my_p2 != p.blank && my_p2 != my;       // beta:  s -> s+2 allowed
  __syncthreads();

  double ll = -INFINITY;
  if (Tn > 0) {
    float* prev = row0 + 2; float* cur = row1 + 2;
    double Coff = 0.0;
    float a = (live && s < 2) ? lp[my] : -INFINITY;
    if (live) { prev[s] = a; alpha[s] = a; }
    if (tid == 0) Cs[0] = 0.0;
    __syncthreads();
    for (int t = 1; t < Tn; ++t) {
      a = -INFINITY;
      if (live) {
        const float e = lp[(long long)t * p.V + my];
        a = lse2(prev[s], prev[s - 1]);
        if (skip_b) a = lse2(a, prev[s - 2]);
        a = (a == -INFINITY) ? -INFINITY : a + e;
      }
      if ((t % kRenorm) == kRenorm - 1) {
        const float mx = block_max(a, red, nwarps);
        if (mx != -INFINITY) { a = (a == -INFINITY) ? a : a - mx; Coff += (double)mx; }
      }
      if (live) { cur[s] = a; alpha[(long long)t * LX + s] = a; }
      if (tid == 0) Cs[t] = Coff;
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
    if (tid == 0) {
      float v = prev[Lx - 1];
      if (Lx > 1) v = lse2(v, prev[Lx - 2]);
      s_ll = (v == -INFINITY) ? -INFINITY : Coff + (double)v;
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

  // backward sweep: rows indexed from 0 with two trailing -inf pads
  __syncthreads();
  for (int i = tid; i < 2 * (LX + 4); i += blockDim.x) row0[i] = -INFINITY;
  __syncthreads();
  float* prev = row0; float* cur = row1;
  double Doff = 0.0;
  for (int t = Tn - 1; t >= 0; --t) {
    const float al = live ? alpha[(long long)t * LX + s] : -INFINITY;      // issued early, consumed after the recursion
    float bt = -INFINITY, e = 0.f;
    if (live) {
      e = lp[(long long)t * p.V + my];
      if (t == Tn - 1) {
        bt = (s >= Lx - 2) ? e : -INFINITY;
      } else {
        float v = lse2(prev[s], prev[s + 1]);
        if (skip_f) v = lse2(v, prev[s + 2]);
        bt = (v == -INFINITY) ? -INFINITY : v + e;
      }
    }
    if (((Tn - 1 - t) % kRenorm) == kRenorm - 1) {
      const float mx = block_max(bt, red, nwarps);
      if (mx != -INFINITY) { bt = (bt == -INFINITY) ? bt : bt - mx; Doff += (double)mx; }
    }
    if (live) {
      cur[s] = bt;
      if (al != -INFINITY && bt != -INFINITY) {
        const float k = (float)(Cs[t] + Doff - ll);
        atomicAdd(&post[(t & 1) * p.V + my], expf(al + bt - e + k));
      }
    }
    __syncthreads();
    if (tid < p.V) {
      float* pp = &post[(t & 1) * p.V + tid];
      dl[(long long)t * p.V + tid] = (expf(lp[(long long)t * p.V + tid]) - *pp) * gs;
      *pp = 0.f;
    }
    float* tmp = prev; prev = cur; cur = tmp;
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

size_t k_ctc_workspace_floats(int B, int L, int S) { return (size_t)B * L * (2 * S + 1) + B; }

int k_ctc_fwd_bwd(const float* logp, const long long* targets, const long long* in_len, const long long* tgt_len, int B, int L, int V,
                  int S, int blank, int zero_infinity, float* alpha_ws, float* nll, float* loss, float* dlogits, const float* dloss,
                  cudaStream_t stream) {
  if (B == 0) return 0;
  NDT1_REQUIRE(blank >= 0 && blank < V, "ctc: blank id %d outside the vocabulary (%d)", blank, V);
  const int LX = 2 * S + 1;
  int threads = ((LX + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  NDT1_REQUIRE(threads <= 1024, "ctc: target length %d too long for one CTA (max 511 labels)", S);
  NDT1_REQUIRE(V <= threads, "ctc: vocabulary %d larger than the CTA (%d threads)", V, threads);
  const size_t base = (size_t)L * sizeof(double) + (size_t)(2 * (LX + 4) + 2 * V + 32) * sizeof(float);
  const size_t with_lp = base + (size_t)L * V * sizeof(float);
  const int lp_in_smem = with_lp <= 200 * 1024;
  const size_t smem = lp_in_smem ? with_lp : base;
  NDT1_REQUIRE(smem <= 200 * 1024, "ctc: %d frames x %d labels do not fit one CTA", L, S);
  CtcParams p{logp, targets, in_len, tgt_len, B, L, V, S, blank, zero_infinity, lp_in_smem, alpha_ws, nll, dlogits, dloss};
  static size_t attr = 0;
  if (smem > attr) { NDT1_CUDA_CHECK(cudaFuncSetAttribute(ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  ctc_kernel<<<B, threads, smem, stream>>>(p);
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
