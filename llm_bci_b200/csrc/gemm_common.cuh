// GEMM problem description shared by the SIMT fp32 kernel (strict mode) and
// the tcgen05/TMA bf16 kernel (fast mode).  One description covers every dense
// contraction of the NDT1 path (SURVEY.md K4, K5, K12, K15, K16, K18 and their
// dgrad / wgrad):
//
//   GEMM_NT  C[b,r,n]  = sum_{j,kc} A[b, r + j*a_row_shift, j*a_col_shift + kc]
//                                 * B[n + j*b_row_shift,    j*b_col_shift + kc]
//            forward linears; the stride-4 / 32-bin stack projection
//            (reference models/ndt1.py:138-140,180) is the case
//            A = activations viewed as (B, T/4, 4*D), a_row_shift = 1,
//            b_col_shift = 4*D, 8 chunks: no im2col buffer is ever built.
//   GEMM_NN  C[b,r,n]  = sum_{j,kc} A[b, r + j*a_row_shift, j*a_col_shift + kc]
//                                 * B[kc + j*b_row_shift,   n + j*b_col_shift]
//            data gradients (B = the forward weight, read "MN-major"); the
//            overlap-add (col2im) backward of the stack is a_row_shift = -1.
//   GEMM_TN  C[n1,n2]  = sum_{j<nchunk} sum_{r<chunk_k}
//                          A[j, r + a_row_shift, n1]
//                        * B[j, r + (n2 / b_chunk_n)*b_row_shift, n2 % b_chunk_n]
//            weight gradients (reduction over trials and rows).
//
// Out-of-range rows / columns of an operand read as zero (TMA OOB fill in the
// tensor-core kernel, explicit predicate in the SIMT kernel).
#pragma once
#include "common.cuh"

enum { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2 };

struct GemmOperand {
  const void* ptr;          // float* (strict) or bf16* (fast)
  long long batch_stride;   // elements between trials
  int nbatch;
  int rows;                 // valid rows per trial
  int cols;                 // valid columns
  int ld;                   // elements between rows
};

struct GemmEpilogue {
  void* out;                // [nb_out][M][ldc]
  int out_bf16;
  long long ldc;
  long long c_batch_stride;
  void* out2;               // optional copy of the pre-activation value ...
  int out2_bf16;
  int out2_deriv;           // ... or (1, GELU only) of act'(pre-activation): the backward epilogue then multiplies (DACT_SAVED) instead of
                            // evaluating erf and exp again; exp(-v^2/2) is already at hand in the forward
  const float* bias;        // [N] or null
  float alpha;
  int act;                  // ACT_*
  const float* gather_tab;  // optional: += gather_tab[gather_idx[b*gather_idx_stride + r]][n]
  const long long* gather_idx;
  long long gather_idx_stride;
  int gather_ld;
  float drop_p;             // dropout after activation/gather (forward) ...
  int drop_bwd;             // ... or the same mask applied to a gradient (backward)
  SeedRef drop_seed; unsigned long long drop_stream;
  const unsigned int* drop_bits;   // optional (tensor-core kernel, N % 32 == 0): the keep bits of this site already drawn -- bit n % 32 of word
                                   // (b * M + r) * (N / 32) + n / 32, the layout of attn_dropbits_kernel -- read instead of running Philox
  const float* resid;       // fp32 residual, indexed like out
  int dact;                 // DACT_*: multiply by act'(saved)
  const void* dact_in;      // indexed like out
  int dact_in_bf16;
  int accumulate;           // 1: out (fp32) += value (atomic when split)
  float* colsum;            // optional (tensor-core kernel only): colsum[n] += sum over rows and trials of the stored value
  // per-day routing (embedder.adapt, models/ndt1.py:118-127,170-171): sel[b] (device, clamped to [0, sel_n)) picks the bias row
  // of output trial b (NT / NN) or, in GEMM_TN run with split_k = trials, the output matrix the trial's product is added to
  const long long* sel;
  int sel_n;
  long long bias_sel_stride, c_sel_stride;
};

struct GemmProblem {
  int mode;
  int M, N;                 // per-trial output rows / output columns
  int nb_out;               // trials in the output (1 for GEMM_TN)
  int nchunk, chunk_k;      // reduction = nchunk * chunk_k
  int a_row_shift, a_col_shift;
  int b_row_shift, b_col_shift;
  int b_chunk_n;            // GEMM_TN only (N if unused)
  int split_k;              // GEMM_TN: split the reduction over this many CTAs (needs accumulate)
  const long long* b_sel;   // GEMM_NT: B operand batch (B.nbatch matrices, B.batch_stride apart) of output trial b = b_sel[b]; null = batch 0
  GemmOperand A, B;
  GemmEpilogue epi;
};

static inline GemmEpilogue gemm_epilogue_default() {
  GemmEpilogue e;
  e.out = nullptr; e.out_bf16 = 0; e.ldc = 0; e.c_batch_stride = 0;
  e.out2 = nullptr; e.out2_bf16 = 0; e.out2_deriv = 0; e.bias = nullptr; e.alpha = 1.f; e.act = ACT_NONE;
  e.gather_tab = nullptr; e.gather_idx = nullptr; e.gather_idx_stride = 0; e.gather_ld = 0;
  e.drop_p = 0.f; e.drop_bwd = 0; e.drop_seed = SeedRef(); e.drop_stream = 0; e.drop_bits = nullptr;
  e.resid = nullptr; e.dact = DACT_NONE; e.dact_in = nullptr; e.dact_in_bf16 = 0;
  e.accumulate = 0; e.colsum = nullptr;
  e.sel = nullptr; e.sel_n = 0; e.bias_sel_stride = 0; e.c_sel_stride = 0;
  return e;
}

// The epilogue of ONE output element.  `acc` is the fp32 accumulator.
// Order (forward):  v = alpha*acc + bias ; out2 = v ; v = act(v) ; v += gather ;
//                   v = dropout(v) ; v += resid
// Order (backward): v = alpha*acc ; v *= dropmask ; v *= act'(saved)
__device__ __forceinline__ void gemm_epilogue_store(const GemmEpilogue& e, int n_total, int rows_c,
                                                    float acc, int b, int r, int n) {
  long long idx = (long long)b * e.c_batch_stride + (long long)r * e.ldc + n;
  long long bias_off = 0;
  if (e.sel) {
    long long d = e.sel[b];
    d = d < 0 ? 0 : (d >= e.sel_n ? e.sel_n - 1 : d);
    idx += d * e.c_sel_stride; bias_off = d * e.bias_sel_stride;
  }
  float v = acc * e.alpha;
  if (e.bias) v += e.bias[bias_off + n];
  if (e.out2) store_from_f32(e.out2, idx, e.out2_bf16, e.out2_deriv ? dact_apply(DACT_GELU_FROM_IN, v) : v);
  v = act_apply(e.act, v);
  if (e.gather_tab) {
    const long long g = e.gather_idx[(long long)b * e.gather_idx_stride + r];
    v += e.gather_tab[g * e.gather_ld + n];
  }
  if (e.drop_p > 0.f) {
    const unsigned long long elem = ((unsigned long long)b * rows_c + r) * (unsigned long long)n_total + n;
    v *= drop_scale_1(e.drop_seed.get(), e.drop_stream, elem, drop_threshold(e.drop_p), 1.0f / (1.0f - e.drop_p));
  }
  if (e.dact != DACT_NONE) v *= dact_apply(e.dact, load_as_f32(e.dact_in, idx, e.dact_in_bf16));
  if (e.resid) v += e.resid[idx];
  if (e.accumulate) atomicAdd((float*)e.out + idx, v);
  else store_from_f32(e.out, idx, e.out_bf16, v);
}

// launchers (gemm_simt.cu / gemm_tc.cu)
int gemm_simt_launch(const GemmProblem& p, int in_bf16, cudaStream_t stream);
int gemm_tc_launch(const GemmProblem& p, cudaStream_t stream);
int gemm_tc_init();
void gemm_tc_set_timeline(unsigned long long* buf);   // debugging: per-CTA phase timestamps (16 u64 per CTA), null = off
