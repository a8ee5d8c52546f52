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
  int B, L, V, S, blank, zero_infinity;
  float* alpha; float* nll; float* dlogits; const float* dloss;
};

__global__ void ctc_kernel(const CtcParams p) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int LX = 2 * p.S + 1;                  // allocated states per trial
  int* ext = (int*)sm;                         // [LX]
  float* row0 = sm + LX;                       // [LX + 2] (two leading -inf pads)
  float* row1 = row0 + LX + 2;
  float* post = row1 + LX + 2;                 // [V]
  __shared__ float s_ll;

  int S = (int)p.tgt_len[b];
  if (S < 0) S = 0;
  if (S > p.S) S = p.S;
  const int Lx = 2 * S + 1;
  long long tn_ll = p.in_len[b];
  int Tn = tn_ll > p.L ? p.L : (int)tn_ll;
  const float* lp = p.logp + (long long)b * p.L * p.V;
  float* alpha = p.alpha + (long long)b * p.L * LX;
  float* dl = p.dlogits ? p.dlogits + (long long)b * p.L * p.V : nullptr;

  for (int s = threadIdx.x; s < Lx; s += blockDim.x) ext[s] = (s & 1) ? (int)p.targets[(long long)b * p.S + (s >> 1)] : p.blank;
  if (threadIdx.x < 2) { row0[threadIdx.x] = -INFINITY; row1[threadIdx.x] = -INFINITY; }
  __syncthreads();

  float ll = -INFINITY;
  if (Tn > 0) {
    float* prev = row0 + 2; float* cur = row1 + 2;
    // t = 0
    for (int s = threadIdx.x; s < Lx; s += blockDim.x) {
      const float a = (s < 2) ? lp[ext[s]] : -INFINITY;
      prev[s] = a; alpha[s] = a;
    }
    __syncthreads();
    for (int t = 1; t < Tn; ++t) {
      for (int s = threadIdx.x; s < Lx; s += blockDim.x) {
        const float e = lp[(long long)t * p.V + ext[s]];
        float a = lse2(prev[s], prev[s - 1]);
        if (s >= 2 && ext[s] != p.blank && ext[s] != ext[s - 2]) a = lse2(a, prev[s - 2]);
        a = (a == -INFINITY) ? -INFINITY : a + e;
        cur[s] = a; alpha[(long long)t * LX + s] = a;
      }
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
    if (threadIdx.x == 0) {
      float v = prev[Lx - 1];
      if (Lx > 1) v = lse2(v, prev[Lx - 2]);
      s_ll = v;
    }
    __syncthreads();
    ll = s_ll;
  } else if (S == 0) {
    ll = 0.f;
  }
  const bool feasible = ll != -INFINITY && ll == ll;
  if (threadIdx.x == 0) p.nll[b] = feasible ? -ll : (p.zero_infinity ? 0.f : INFINITY);
  if (!dl) return;
  const float gs = p.dloss ? *p.dloss : 1.f;
  // rows beyond the input length (and infeasible trials) receive zero gradient
  const int t_zero_from = (feasible && Tn > 0) ? Tn : 0;
  for (long long e = (long long)t_zero_from * p.V + threadIdx.x; e < (long long)p.L * p.V; e += blockDim.x) dl[e] = 0.f;
  if (!feasible || Tn <= 0) return;

  // backward sweep; beta rows padded with two trailing -inf
  float* prev = row0; float* cur = row1;       // use [0, Lx) + 2 trailing pads
  __syncthreads();
  for (int s = threadIdx.x; s < Lx + 2; s += blockDim.x) { prev[s] = -INFINITY; cur[s] = -INFINITY; }
  __syncthreads();
  for (int t = Tn - 1; t >= 0; --t) {
    for (int c = threadIdx.x; c < p.V; c += blockDim.x) post[c] = 0.f;
    __syncthreads();
    for (int s = threadIdx.x; s < Lx; s += blockDim.x) {
      const float e = lp[(long long)t * p.V + ext[s]];
      float bt;
      if (t == Tn - 1) {
        bt = (s >= Lx - 2) ? e : -INFINITY;
      } else {
        float a = lse2(prev[s], prev[s + 1]);
        if (s + 2 < Lx && ext[s + 2] != p.blank && ext[s + 2] != ext[s]) a = lse2(a, prev[s + 2]);
        bt = (a == -INFINITY) ? -INFINITY : a + e;
      }
      cur[s] = bt;
      const float ab = alpha[(long long)t * LX + s] + bt;
      if (ab != -INFINITY) atomicAdd(&post[ext[s]], expf(ab - e - ll));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.V; c += blockDim.x) dl[(long long)t * p.V + c] = (expf(lp[(long long)t * p.V + c]) - post[c]) * gs;
    float* tmp = prev; prev = cur; cur = tmp;
    __syncthreads();
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
  CtcParams p{logp, targets, in_len, tgt_len, B, L, V, S, blank, zero_infinity, alpha_ws, nll, dlogits, dloss};
  const int LX = 2 * S + 1;
  int threads = ((LX + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > 512) threads = 512;
  const size_t smem = (size_t)(LX + 2 * (LX + 2) + V) * sizeof(float);
  NDT1_REQUIRE(smem <= 48 * 1024, "ctc: target length %d too long for one CTA", S);
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
