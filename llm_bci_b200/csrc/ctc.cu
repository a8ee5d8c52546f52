// Phoneme head loss: LogSoftmax + CTC forward/backward, and greedy decode.
// Reference: nn.LogSoftmax + nn.CTCLoss(reduction="none", blank, zero_infinity)
// at models/ndt1.py:493-500,517,581 (conventions: SURVEY.md A.6) and the
// greedy collapse of utils/eval_bci.py:41-48.
//
// One CTA per trial.  The states of the extended label sequence (blank, l1,
// blank, l2, ..., blank) live in the registers of ONE warp per sweep; alpha is
// swept forward and beta backward AT THE SAME TIME by two warps (L dependent
// steps, not 2 L, no block barrier inside), then the gradient w.r.t. the LOGITS,
//     dlogits[t,c] = (softmax[t,c] - posterior[t,c]) * dloss      (t <  len)
//                  = 0                                            (t >= len)
// is emitted by a parallel pass, so no separate log-softmax backward exists.
// The kernel is latency-, not bandwidth-bound.
#include "kernels.cuh"
#include <stdlib.h>

namespace {

__global__ void log_softmax_kernel(const float* __restrict__ logits, float* __restrict__ logp, long long rows, int V, int ld_in) { pdl_grid_sync();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = logits + r * ld_in;
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
  float* alpha; float* beta; float* nll; float* dlogits; const float* dloss;
  double* offs;              // [B][2][L] forward / backward row offsets + [B] log-likelihoods (in the recursion's log unit), for the posterior kernel
};

constexpr int kRenorm = 8;    // re-centre the alpha/beta rows every kRenorm frames
constexpr int kCtcWarps = 8;

// FAST (bf16 engine mode): the whole recursion runs in BASE 2 -- log-probabilities are scaled by log2(e) once, so an
// exponential / logarithm is a single MUFU ex2 / lg2 with no multiply around it.  Otherwise (strict fp32 mode, the
// stand-alone operator): natural logarithms with libm-accurate expf / logf.
template <bool FAST> __device__ __forceinline__ float exp_t(float x) {
  if (FAST) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  return expf(x);
}
template <bool FAST> __device__ __forceinline__ float log_t(float x) {
  if (FAST) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  return logf(x);
}
template <bool FAST> __device__ __forceinline__ constexpr float log_unit() { return FAST ? 1.4426950408889634f : 1.0f; }
template <bool FAST>
__device__ __forceinline__ float lse3(float a, float b, float c) {
  // branch-free (the SPT independent recurrences of a lane interleave): all -inf -> exp(-inf) = 0 -> log(0) = -inf
  const float m = fmaxf(a, fmaxf(b, c));
  const float ms = (m == -INFINITY) ? 0.f : m;
  return ms + log_t<FAST>(exp_t<FAST>(a - ms) + exp_t<FAST>(b - ms) + exp_t<FAST>(c - ms));
}

// One sweep (alpha: FWD, beta: !FWD) by ONE warp: lane j keeps the SPT consecutive states [j*SPT, (j+1)*SPT) of the
// extended label sequence in registers, neighbours across lanes come by shuffle -- no shared memory, no block
// barrier on the L-step dependent chain.  Rows are re-centred every kRenorm frames (offsets in double) and
// streamed to the workspace for the posterior pass.
template <bool FAST, int SPT, bool FWD>
__device__ __forceinline__ void ctc_sweep(const CtcParams& p, const float* __restrict__ lp, float lps, const int* __restrict__ lab, int Lx, int LX,
                                          int Tn, float* __restrict__ rows, double* __restrict__ offs, float* fin) {
  // lp: log-probabilities, lps: factor that brings them to the recursion's log unit (1 when lp is already scaled)
  const int lane = threadIdx.x & 31;
  const int s0 = lane * SPT;
  int my[SPT]; bool live[SPT], skip[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = s0 + i;
    live[i] = s < Lx;
    my[i] = live[i] ? lab[s] : p.blank;
    if (FWD) skip[i] = live[i] && s >= 2 && my[i] != p.blank && my[i] != lab[s - 2];          // s-2 -> s allowed
    else skip[i] = live[i] && (s + 2 < Lx) && lab[s + 2] != p.blank && lab[s + 2] != my[i];   // s -> s+2 allowed
  }
  float a[SPT], e[SPT];
  double off = 0.0;
  const int t0 = FWD ? 0 : Tn - 1;
#pragma unroll
  for (int i = 0; i < SPT; ++i) e[i] = lp[t0 * p.V + my[i]] * lps;
  for (int k = 0; k < Tn; ++k) {
    const int t = FWD ? k : Tn - 1 - k;
    float v[SPT];
    if (k == 0) {
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const int s = s0 + i;
        const bool init = FWD ? (s < 2) : (s >= Lx - 2);
        v[i] = (live[i] && init) ? e[i] : -INFINITY;
      }
    } else {
      // the two states beyond this lane's block, from the neighbouring lane(s)
      float n1, n2;
      if (FWD) {
        n1 = __shfl_up_sync(0xffffffffu, a[SPT - 1], 1);
        n2 = SPT >= 2 ? __shfl_up_sync(0xffffffffu, a[SPT >= 2 ? SPT - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, a[0], 2);
        if (lane < 1) n1 = -INFINITY;
        if (lane < (SPT >= 2 ? 1 : 2)) n2 = -INFINITY;
      } else {
        n1 = __shfl_down_sync(0xffffffffu, a[0], 1);
        n2 = SPT >= 2 ? __shfl_down_sync(0xffffffffu, a[SPT >= 2 ? 1 : 0], 1) : __shfl_down_sync(0xffffffffu, a[0], 2);
        if (lane > 30) n1 = -INFINITY;
        if (lane > (SPT >= 2 ? 30 : 29)) n2 = -INFINITY;
      }
      // the SPT recurrences of a lane are independent: written stage by stage so that their MUFU latencies overlap
      float x0[SPT], x1[SPT], x2[SPT], ms[SPT], sum[SPT];
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        x0[i] = a[i];
        if (FWD) {
          x1[i] = i >= 1 ? a[i >= 1 ? i - 1 : 0] : n1;
          x2[i] = i >= 2 ? a[i >= 2 ? i - 2 : 0] : (i == 1 ? n1 : n2);
        } else {
          x1[i] = i + 1 < SPT ? a[i + 1 < SPT ? i + 1 : 0] : n1;
          x2[i] = i + 2 < SPT ? a[i + 2 < SPT ? i + 2 : 0] : (i + 2 == SPT ? n1 : n2);
        }
        x2[i] = skip[i] ? x2[i] : -INFINITY;
      }
      // log(e^a + e^b + e^c) = m + log(1 + e^(r - m) + e^(q - m)) with {m, r, q} the three values sorted so that m is the largest:
      // the largest term is exactly 1, two exponentials per state instead of three (the MUFU pipe is what a frame step waits for)
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const float hi = fmaxf(x0[i], x1[i]), lo = fminf(x0[i], x1[i]);
        const float m = fmaxf(hi, x2[i]);
        x1[i] = fminf(hi, x2[i]); x0[i] = lo;
        ms[i] = m;
      }
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const float mz = (ms[i] == -INFINITY) ? 0.f : ms[i];   // all -inf: no NaN from (-inf) - (-inf); the result is forced below
        x0[i] = exp_t<FAST>(x0[i] - mz); x1[i] = exp_t<FAST>(x1[i] - mz);
      }
#pragma unroll
      for (int i = 0; i < SPT; ++i) sum[i] = 1.0f + x0[i] + x1[i];
#pragma unroll
      for (int i = 0; i < SPT; ++i) sum[i] = log_t<FAST>(sum[i]);
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const float r = ms[i] + sum[i];                        // ms = -inf stays -inf
        v[i] = (live[i] && r != -INFINITY) ? r + e[i] : -INFINITY;
      }
    }
    if (k + 1 < Tn) {                 // next frame's emissions: independent of the recurrence, issued early
      const int tn = FWD ? t + 1 : t - 1;
#pragma unroll
      for (int i = 0; i < SPT; ++i) e[i] = lp[tn * p.V + my[i]] * lps;
    }
    if ((k % kRenorm) == kRenorm - 1) {
      float m = v[0];
#pragma unroll
      for (int i = 1; i < SPT; ++i) m = fmaxf(m, v[i]);
      m = warp_max(m);
      if (m != -INFINITY) {
#pragma unroll
        for (int i = 0; i < SPT; ++i) v[i] = (v[i] == -INFINITY) ? v[i] : v[i] - m;
        off += (double)m;
      }
    }
    float* row = rows + (long long)t * LX + s0;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      if (s0 + i < LX) row[i] = v[i];
      a[i] = v[i];
    }
    if (lane == 0) offs[t] = off;
  }
  if (FWD) {        // log-likelihood = lse(alpha[Tn-1][Lx-1], alpha[Tn-1][Lx-2]) + offset
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      if (s0 + i == Lx - 1) fin[0] = a[i];
      if (s0 + i == Lx - 2) fin[1] = a[i];
    }
  }
}

// (Measured dead end, round 2: the same sweep in the PROBABILITY domain -- two adds and a multiply per state and frame instead of
// max / 3 ex2 / add / lg2, rows rescaled to a maximum of 1 every second frame -- cut the sweep from 92 to ~60 us, but is wrong as
// soon as the network is confident: the row maximum sits on states that can no longer reach the end (all-blank prefixes), the
// states of the correct alignment fall more than 2^126 below it and flush to zero, and the posteriors (alpha * beta) lose them:
// gradient errors of order 1 at logit scales >= 12.  Per-frame scaling cannot fix a per-STATE dynamic range; the log domain stays.)
// One CTA (8 warps) per trial.  Warp 0 sweeps alpha forward while warp 1 sweeps beta backward (L dependent steps,
// not 2 L); then all warps turn alpha + beta into label posteriors, one frame per warp, and write the gradient
// w.r.t. the logits.  The posterior of label c sums the states that carry c; the states are grouped by label once
// per trial (a CSR index built in shared memory), so no atomics are involved and the sums are deterministic.
// Rows are kept re-centred (alpha_hat = alpha - C_t, beta_hat = beta - D_t) so the fp32 log-space values stay O(10)
// instead of O(loss): posteriors keep ~1e-6 relative accuracy where plain fp32 log-space CTC (torch's kernel
// included) loses ~3e-5 on a 250-frame utterance.
template <bool FAST, int SPT>
__global__ void __launch_bounds__(kCtcWarps * 32) ctc_kernel(const CtcParams p) { pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int LX = 2 * p.S + 1, LXA = SPT * 32;
  double* Cs = (double*)sm_raw;                     // [L] forward offsets
  double* Ds = Cs + p.L;                            // [L] backward offsets
  float* wbuf = (float*)(Ds + p.L);                 // [warps][LXA] posterior weights of one frame
  int* lab = (int*)(wbuf + kCtcWarps * LXA);        // [LXA + 2] extended labels
  int* cstart = lab + LXA + 2;                      // [V + 1] states grouped by label: start offsets ...
  int* cpos = cstart + p.V + 1;                     // [S]     ... and target positions
  float* fin = (float*)(cpos + (p.S > 0 ? p.S : 1));   // [2]
  float* lp_s = fin + 2;                            // [L*V] when it fits
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
  constexpr float kUnit = log_unit<FAST>();
  if (p.lp_in_smem) {              // staged once, already in the recursion's log unit; 8 independent loads in flight per thread
    const int n = p.L * p.V;
    for (int i0 = tid; i0 < n; i0 += blockDim.x * 8) {
      float t8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; t8[u] = i < n ? __ldg(lp_g + i) : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; if (i < n) lp_s[i] = t8[u] * kUnit; }
    }
  }
  const float lps = p.lp_in_smem ? 1.0f : kUnit;
  for (int i = tid; i < LXA + 2; i += blockDim.x) lab[i] = (i < Lx && (i & 1)) ? (int)p.targets[(long long)b * p.S + (i >> 1)] : p.blank;
  if (tid < 2) fin[tid] = -INFINITY;
  __syncthreads();
  double ll = -INFINITY;
  if (Tn > 0) {
    if (warp == 0) ctc_sweep<FAST, SPT, true>(p, lp, lps, lab, Lx, LX, Tn, alpha, Cs, fin);
    else if (warp == 1) ctc_sweep<FAST, SPT, false>(p, lp, lps, lab, Lx, LX, Tn, beta, Ds, fin);
    __syncthreads();
    if (tid == 0) {
      const float v = lse3<FAST>(fin[0], fin[1], -INFINITY);     // (in the recursion's log unit)
      s_ll = (v == -INFINITY) ? -INFINITY : Cs[Tn - 1] + (double)v;
    }
    __syncthreads();
    ll = s_ll;
  } else if (S == 0) {
    ll = 0.0;
  }
  const bool feasible = (ll != -INFINITY) && (ll == ll);
  if (tid == 0) p.nll[b] = feasible ? (float)(-ll / (double)kUnit) : (p.zero_infinity ? 0.f : INFINITY);
  if (!p.dlogits) return;
  // hand the row offsets and the log-likelihood to the posterior kernel (one CTA per trial cannot fill the GPU; that pass can)
  double* og = p.offs + (long long)b * 2 * p.L;
  for (int t = tid; t < Tn; t += blockDim.x) { og[t] = Cs[t]; og[p.L + t] = Ds[t]; }
  if (tid == 0) p.offs[(long long)p.B * 2 * p.L + b] = feasible ? ll : -INFINITY;
}

// Gradient pass: dlogits[t,c] = (softmax[t,c] - sum_{s: l'(s)=c} alpha beta / (p_t(c) P)) * dloss for t < len, 0 beyond (and
// everywhere for an infeasible trial).  One warp per frame, kPostFrames frames per CTA, grid (frame chunks, trials): the
// label posterior sums the states that carry the label; states are grouped by label once per CTA (a CSR index in shared
// memory), so no atomics are involved and the sums are deterministic.
constexpr int kPostFrames = 32;
template <bool FAST, int SPT>
__global__ void __launch_bounds__(kCtcWarps * 32) ctc_posterior_kernel(const CtcParams p) { pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int LX = 2 * p.S + 1, LXA = SPT * 32;
  float* wbuf = (float*)sm_raw;                     // [warps][LXA] posterior weights of one frame
  int* lab = (int*)(wbuf + kCtcWarps * LXA);        // [LXA + 2] extended labels
  int* cstart = lab + LXA + 2;                      // [V + 1] states grouped by label: start offsets ...
  int* cpos = cstart + p.V + 1;                     // [S]     ... and target positions
  int S = (int)p.tgt_len[b];
  S = S < 0 ? 0 : (S > p.S ? p.S : S);
  const int Lx = 2 * S + 1;
  const long long tn_ll = p.in_len[b];
  const int Tn = tn_ll > p.L ? p.L : (tn_ll < 0 ? 0 : (int)tn_ll);
  const float* lp = p.logp + (long long)b * p.L * p.V;
  const float* alpha = p.alpha + (long long)b * p.L * LX;
  const float* beta = p.beta + (long long)b * p.L * LX;
  float* dl = p.dlogits + (long long)b * p.L * p.V;
  const double* og = p.offs + (long long)b * 2 * p.L;
  const double ll = p.offs[(long long)p.B * 2 * p.L + b];
  const bool feasible = (ll != -INFINITY) && (ll == ll) && Tn > 0;
  constexpr float kUnit = log_unit<FAST>();
  const float gs = p.dloss ? *p.dloss : 1.f;
  const int t_lo = blockIdx.x * kPostFrames, t_hi = min(p.L, t_lo + kPostFrames);
  const int t_live = feasible ? Tn : 0;                 // frames [t_live, L) get a zero gradient
  for (int e = max(t_lo, t_live) * p.V + tid; e < t_hi * p.V; e += blockDim.x) dl[e] = 0.f;
  if (t_lo >= t_live) return;

  for (int i = tid; i < LXA + 2; i += blockDim.x) lab[i] = (i < Lx && (i & 1)) ? (int)p.targets[(long long)b * p.S + (i >> 1)] : p.blank;
  __syncthreads();
  if (tid < p.V) {
    int n = 0;
    for (int j = 0; j < S; ++j) n += (lab[2 * j + 1] == tid);
    cstart[tid + 1] = n;
  }
  if (tid == 0) cstart[0] = 0;
  __syncthreads();
  if (tid == 0) for (int c = 0; c < p.V; ++c) cstart[c + 1] += cstart[c];
  __syncthreads();
  if (tid < p.V) {
    int n = cstart[tid];
    for (int j = 0; j < S; ++j) if (lab[2 * j + 1] == tid) cpos[n++] = j;
  }
  __syncthreads();

  float* wb = wbuf + warp * LXA;
  for (int t = t_lo + warp; t < min(t_hi, t_live); t += kCtcWarps) {
    const float kt = (float)(og[t] + og[p.L + t] - ll);
    float blank_acc = 0.f;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s2 = lane + 32 * i;
      float w = 0.f;
      if (s2 < Lx) {
        const float al = alpha[(long long)t * LX + s2], be = beta[(long long)t * LX + s2];
        if (al != -INFINITY && be != -INFINITY) w = exp_t<FAST>(al + be - lp[t * p.V + lab[s2]] * kUnit + kt);
      }
      wb[s2] = w;
      if (!(s2 & 1)) blank_acc += w;
    }
    blank_acc = warp_sum(blank_acc);
    __syncwarp();
    for (int c = lane; c < p.V; c += 32) {
      float acc = (c == p.blank) ? blank_acc : 0.f;
      for (int k = cstart[c]; k < cstart[c + 1]; ++k) acc += wb[2 * cpos[k] + 1];
      dl[(long long)t * p.V + c] = (exp_t<FAST>(lp[t * p.V + c] * kUnit) - acc) * gs;
    }
    __syncwarp();
  }
}

__global__ void sum_nll_kernel(const float* nll, int B, float* loss) { pdl_grid_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += nll[b];
    *loss += s;
  }
}

// argmax over V then the reference's collapse: emit when id != last EMITTED id and id != blank.
__global__ void greedy_kernel(const float* logp, int B, int L, int V, int blank, long long* out_ids, long long* out_len) { pdl_grid_sync();
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

// Levenshtein distance between a decoded id sequence and its target (editdistance.eval in utils/eval_bci.py:11-14), one warp
// per pair, one DP row per step.  new[j] = min(t[j], new[j-1] + 1) with t[j] = min(prev[j] + 1, prev[j-1] + cost) unrolls to
// new[j] = j + min_{k<=j}(t[k] - k): the row is a prefix minimum (warp scan, 32 columns at a time, carry across blocks).
constexpr int kEdWarps = 4;
__global__ void __launch_bounds__(kEdWarps * 32) edit_distance_kernel(const long long* __restrict__ pred, const long long* __restrict__ pred_len, int Lp,
                                                                     const long long* __restrict__ tgt, const long long* __restrict__ tgt_len, int Lt,
                                                                     int B, long long* __restrict__ out) { pdl_grid_sync();
  extern __shared__ int ed_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kEdWarps + warp;
  if (b >= B) return;
  long long ml = pred_len[b], nl = tgt_len[b];
  const int m = (int)(ml < 0 ? 0 : (ml > Lp ? Lp : ml)), n = (int)(nl < 0 ? 0 : (nl > Lt ? Lt : nl));
  int* prev = ed_sm + warp * 2 * (Lt + 1);
  int* cur = prev + (Lt + 1);
  const long long* pa = pred + (long long)b * Lp;
  const long long* tb = tgt + (long long)b * Lt;
  for (int j = lane; j <= n; j += 32) prev[j] = j;
  __syncwarp();
  for (int i = 1; i <= m; ++i) {
    const long long a = pa[i - 1];
    int carry = i;                                  // new[0] - 0
    for (int j0 = 1; j0 <= n; j0 += 32) {
      const int j = j0 + lane;
      int v = 1 << 29;
      if (j <= n) v = min(prev[j] + 1, prev[j - 1] + (a != tb[j - 1] ? 1 : 0)) - j;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = min(v, u);
      }
      v = min(v, carry);
      if (j <= n) cur[j] = v + j;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) cur[0] = i;
    __syncwarp();
    int* t = prev; prev = cur; cur = t;
  }
  if (lane == 0) out[b] = prev[n];
}

}  // namespace

int k_log_softmax(const float* logits, float* logp, long long rows, int V, cudaStream_t stream, int ld_in) {
  if (rows == 0) return 0;
  ndt1_launch(log_softmax_kernel, ndt1_cdiv(rows, 8), 256, 0, stream, logits, logp, rows, V, ld_in > 0 ? ld_in : V);
  NDT1_CHECK_LAUNCH();
  return 0;
}

// alpha and beta rows + (8-byte aligned) the row offsets and log-likelihoods the sweep kernel hands to the posterior kernel
size_t k_ctc_workspace_floats(int B, int L, int S) { return 2 * (size_t)B * L * (2 * S + 1) + 2 + 2 * ((size_t)B * 2 * L + B); }

template <bool FAST, int SPT>
static int ctc_launch(const CtcParams& p, size_t smem, cudaStream_t stream) {
  static Ndt1PerDeviceSize attr;
  if (smem > attr.here()) {
    NDT1_CUDA_CHECK(cudaFuncSetAttribute(ctc_kernel<FAST, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.here() = smem;
  }
  // log-probs in, alpha + beta rows out (latency-bound: 2 L dependent steps per trial)
  if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)p.B * p.L * (4.0 * p.V + 8.0 * (2 * p.S + 1)));
  ndt1_launch(ctc_kernel<FAST, SPT>, p.B, kCtcWarps * 32, smem, stream, p);
  NDT1_CHECK_LAUNCH();
  if (p.dlogits) {
    const int LXA = SPT * 32;
    const size_t smem2 = (size_t)(kCtcWarps * LXA + (LXA + 2) + (p.V + 1) + (p.S > 0 ? p.S : 1)) * sizeof(float);
    if (g_ndt1_prof_on) ndt1_prof_note(0.0, (double)p.B * p.L * (8.0 * p.V + 8.0 * (2 * p.S + 1)));   // alpha, beta, log-probs in; d(logits) out
    ndt1_launch(ctc_posterior_kernel<FAST, SPT>, dim3(ndt1_cdiv(p.L, kPostFrames), p.B), kCtcWarps * 32, smem2, stream, p);
    NDT1_CHECK_LAUNCH();
  }
  return 0;
}
template <bool FAST>
static int ctc_dispatch(const CtcParams& p, int spt, size_t smem, cudaStream_t stream) {
  switch (spt) {
    case 1: return ctc_launch<FAST, 1>(p, smem, stream);
    case 2: return ctc_launch<FAST, 2>(p, smem, stream);
    case 4: return ctc_launch<FAST, 4>(p, smem, stream);
    case 8: return ctc_launch<FAST, 8>(p, smem, stream);
    default: return ctc_launch<FAST, 16>(p, smem, stream);
  }
}

int k_ctc_fwd_bwd(const float* logp, const long long* targets, const long long* in_len, const long long* tgt_len, int B, int L, int V,
                  int S, int blank, int zero_infinity, float* alpha_ws, float* nll, float* loss, float* dlogits, const float* dloss,
                  cudaStream_t stream, int fast_math) {
  if (B == 0) return 0;
  NDT1_REQUIRE(blank >= 0 && blank < V, "ctc: blank id %d outside the vocabulary (%d)", blank, V);
  NDT1_REQUIRE(V <= kCtcWarps * 32, "ctc: vocabulary %d larger than the CTA (%d threads)", V, kCtcWarps * 32);
  const int LX = 2 * S + 1;
  int spt = 1;
  while (spt * 32 < LX) spt *= 2;
  NDT1_REQUIRE(spt <= 16, "ctc: target length %d too long for one warp per sweep (max 255 labels)", S);
  const int LXA = spt * 32;
  const size_t base = (size_t)2 * L * sizeof(double) + (size_t)(kCtcWarps * LXA + (LXA + 2) + (V + 1) + (S > 0 ? S : 1) + 2) * sizeof(float);
  const size_t with_lp = base + (size_t)L * V * sizeof(float);
  const int lp_in_smem = with_lp <= 200 * 1024;
  const size_t smem = lp_in_smem ? with_lp : base;
  NDT1_REQUIRE(smem <= 200 * 1024, "ctc: %d frames x %d labels do not fit one CTA", L, S);
  float* beta_ws = alpha_ws + (size_t)B * L * LX;
  double* offs = (double*)(((uintptr_t)(beta_ws + (size_t)B * L * LX) + 7) & ~(uintptr_t)7);
  CtcParams p{logp, targets, in_len, tgt_len, B, L, V, S, blank, zero_infinity, lp_in_smem, alpha_ws, beta_ws, nll, dlogits, dloss, offs};
  static const int force_fast = getenv("NDT1_CTC_FAST") && getenv("NDT1_CTC_FAST")[0] == '1';      // debugging: the stand-alone operator in FAST mode
  if (fast_math || force_fast) NDT1_TRY(ctc_dispatch<true>(p, spt, smem, stream));
  else NDT1_TRY(ctc_dispatch<false>(p, spt, smem, stream));
  if (loss) {
    ndt1_launch(sum_nll_kernel, 1, 32, 0, stream, nll, B, loss);
    NDT1_CHECK_LAUNCH();
  }
  return 0;
}

int k_ctc_greedy_decode(const float* logp, int B, int L, int V, int blank, long long* out_ids, long long* out_len, cudaStream_t stream) {
  if (B == 0) return 0;
  ndt1_launch(greedy_kernel, ndt1_cdiv(B, 64), 64, 0, stream, logp, B, L, V, blank, out_ids, out_len);
  NDT1_CHECK_LAUNCH();
  return 0;
}

int k_edit_distance(const long long* pred, const long long* pred_len, int Lp, const long long* tgt, const long long* tgt_len, int Lt, int B,
                    long long* out, cudaStream_t stream) {
  if (B == 0) return 0;
  const size_t smem = (size_t)kEdWarps * 2 * (Lt + 1) * sizeof(int);
  NDT1_REQUIRE(smem <= 48 * 1024, "edit_distance: targets of %d ids are too long", Lt);
  ndt1_launch(edit_distance_kernel, ndt1_cdiv(B, kEdWarps), kEdWarps * 32, smem, stream, pred, pred_len, Lp, tgt, tgt_len, Lt, B, out);
  NDT1_CHECK_LAUNCH();
  return 0;
}
