// Stand-alone check of the tcgen05 GEMM (gemm_tc.cu) against the CUDA-core
// GEMM (gemm_simt.cu) on the NDT1 shapes, plus a CUDA-event timing of each
// shape.  Not part of the library; run on a B200:
//   ./gemm_selftest [quick]
#include "../gemm_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <math.h>

extern "C" const char* ndt1_last_error(void);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static uint32_t g_rng = 12345u;
static float frand() { g_rng = g_rng * 1664525u + 1013904223u; return ((g_rng >> 8) & 0xFFFF) / 65536.0f - 0.5f; }

static bf16* dev_bf16(size_t n, bool zero = false) {
  std::vector<bf16> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(zero ? 0.f : frand());
  bf16* d; CK(cudaMalloc(&d, n * sizeof(bf16)));
  CK(cudaMemcpy(d, h.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
  return d;
}
static float* dev_f32(size_t n, bool rnd) {
  std::vector<float> h(n, 0.f);
  if (rnd) for (size_t i = 0; i < n; ++i) h[i] = frand();
  float* d; CK(cudaMalloc(&d, n * sizeof(float)));
  CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

static int g_fail = 0;

static void run_case(const char* name, GemmProblem p, size_t out_elems, double flops, bool out_bf16 = false) {
  float* ref = nullptr; void* out = nullptr;
  CK(cudaMalloc(&ref, out_elems * sizeof(float)));
  CK(cudaMalloc(&out, out_elems * (out_bf16 ? 2 : 4)));
  CK(cudaMemset(ref, 0, out_elems * sizeof(float)));
  CK(cudaMemset(out, 0, out_elems * (out_bf16 ? 2 : 4)));
  GemmProblem pr = p; pr.epi.out = ref; pr.epi.out_bf16 = 0;
  if (gemm_simt_launch(pr, 1, 0)) { printf("%-28s SIMT launch failed: %s\n", name, ndt1_last_error()); g_fail++; return; }
  CK(cudaDeviceSynchronize());
  GemmProblem pt = p; pt.epi.out = out; pt.epi.out_bf16 = out_bf16;
  if (gemm_tc_launch(pt, 0)) { printf("%-28s TC launch failed: %s\n", name, ndt1_last_error()); g_fail++; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s TC kernel failed: %s\n", name, cudaGetErrorString(e)); exit(2); }
  std::vector<float> hr(out_elems), ho(out_elems);
  CK(cudaMemcpy(hr.data(), ref, out_elems * 4, cudaMemcpyDeviceToHost));
  if (out_bf16) {
    std::vector<bf16> hb(out_elems);
    CK(cudaMemcpy(hb.data(), out, out_elems * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < out_elems; ++i) ho[i] = __bfloat162float(hb[i]);
  } else {
    CK(cudaMemcpy(ho.data(), out, out_elems * 4, cudaMemcpyDeviceToHost));
  }
  double maxref = 0, maxerr = 0; size_t bad = 0;
  for (size_t i = 0; i < out_elems; ++i) {
    maxref = fmax(maxref, fabs(hr[i]));
    double d = fabs((double)hr[i] - ho[i]);
    if (d > maxerr) { maxerr = d; bad = i; }
  }
  const double tol = (out_bf16 ? 1e-2 : 2e-4) * fmax(maxref, 1e-6);
  // timing (accumulating epilogues keep adding; harmless)
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) gemm_tc_launch(pt, 0);
  CK(cudaEventRecord(e0));
  const int iters = 20;
  for (int i = 0; i < iters; ++i) gemm_tc_launch(pt, 0);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
  const bool ok = maxerr <= tol && maxref > 0;
  printf("%-28s %s maxref=%.4g maxerr=%.3g (at %zu ref=%.5g got=%.5g)  %.1f us  %.1f TFLOP/s\n", name, ok ? "OK  " : "FAIL", maxref,
         maxerr, bad, hr[bad], ho[bad], ms * 1e3, flops / (ms * 1e-3) / 1e12);
  if (!ok) g_fail++;
  if (getenv("GEMM_TL") && strstr(name, getenv("GEMM_TL"))) {      // per-CTA phase timeline of this case (see tools/gemm_timeline.py)
    unsigned long long* tl = nullptr;
    CK(cudaMalloc(&tl, 148 * 16 * 8)); CK(cudaMemset(tl, 0, 148 * 16 * 8));
    gemm_tc_set_timeline(tl);
    gemm_tc_launch(pt, 0);
    CK(cudaDeviceSynchronize());
    gemm_tc_set_timeline(nullptr);
    std::vector<unsigned long long> h(148 * 16);
    CK(cudaMemcpy(h.data(), tl, h.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < 148; ++c) if (h[c * 16] && h[c * 16] < t0) t0 = h[c * 16];
    const char* names[16] = {"CTA start", "barriers + TMEM ready", "griddep wait passed", "first TMA issued", "all loads issued", "first operands landed",
                             "last MMA issued", "first accumulator ready", "last accumulator ready", "epilogue warp 0/4 done", "epilogue warp 1/5 done",
                             "epilogue warp 2/6 done", "epilogue warp 3/7 done", "", "TMEM freed (exit)", ""};
    for (int k : {0, 1, 3, 5, 7, 4, 6, 8, 9, 12, 14}) {
      double mn = 1e30, mx = 0, sum = 0; int n = 0;
      for (int c = 0; c < 148; ++c) {
        const unsigned long long v = h[c * 16 + k];
        if (!v || !h[c * 16]) continue;
        const double us = (double)(v - t0) / 1e3;
        mn = fmin(mn, us); mx = fmax(mx, us); sum += us; ++n;
      }
      if (n) printf("    %-28s mean %6.2f  min %6.2f  max %6.2f  (n=%d)\n", names[k], sum / n, mn, mx, n);
    }
    CK(cudaFree(tl));
  }
  CK(cudaFree(ref)); CK(cudaFree(out));
}

static GemmProblem base_problem(int mode, int M, int N, int K) {
  GemmProblem p;
  p.b_sel = nullptr;
  p.mode = mode; p.M = M; p.N = N; p.nb_out = 1; p.nchunk = 1; p.chunk_k = K;
  p.a_row_shift = p.a_col_shift = p.b_row_shift = p.b_col_shift = 0; p.b_chunk_n = N; p.split_k = 1;
  p.epi = gemm_epilogue_default();
  return p;
}

int main(int argc, char** argv) {
  // usage: gemm_selftest [trials]   (default 8; 32 = the benchmark batch)
  const bool quick = argc > 1 && atoi(argv[1]) <= 0;
  const int trials = (argc > 1 && atoi(argv[1]) > 0) ? atoi(argv[1]) : 8;
  if (gemm_tc_init()) { printf("init failed: %s\n", ndt1_last_error()); return 1; }
  const int Bt = quick ? 2 : trials, T = 1000, D = 256, H = 1024, Tp = (T - 32) / 4 + 1;  // 243
  const int M = Bt * Tp;

  {  // plain NT: y = x W^T + b, fp32 out
    bf16* x = dev_bf16((size_t)M * H); bf16* w = dev_bf16((size_t)H * H); float* bias = dev_f32(H, true);
    GemmProblem p = base_problem(GEMM_NT, M, H, H);
    p.A = {x, 0, 1, M, H, H}; p.B = {w, 0, 1, H, H, H};
    p.epi.ldc = H; p.epi.bias = bias;
    run_case("NT 1024x1024 bias", p, (size_t)M * H, 2.0 * M * H * H);
    p.epi.act = ACT_GELU;
    run_case("NT gelu bf16-out", p, (size_t)M * H, 2.0 * M * H * H, true);
    cudaFree(x); cudaFree(w); cudaFree(bias);
  }
  {  // NT, N=3072 (QKV)
    bf16* x = dev_bf16((size_t)M * H); bf16* w = dev_bf16((size_t)3 * H * H);
    GemmProblem p = base_problem(GEMM_NT, M, 3 * H, H);
    p.A = {x, 0, 1, M, H, H}; p.B = {w, 0, 1, 3 * H, H, 3 * H > 0 ? H : H};
    p.epi.ldc = 3 * H;
    run_case("NT qkv 3072", p, (size_t)M * 3 * H, 2.0 * M * 3 * H * H);
    cudaFree(x); cudaFree(w);
  }
  {  // NT head: N=41, K=1024, fp32 out ld 41
    bf16* x = dev_bf16((size_t)M * H); bf16* w = dev_bf16((size_t)41 * H);
    GemmProblem p = base_problem(GEMM_NT, M, 41, H);
    p.A = {x, 0, 1, M, H, H}; p.B = {w, 0, 1, 41, H, H};
    p.epi.ldc = 41;
    run_case("NT head N=41", p, (size_t)M * 41, 2.0 * M * 41 * H);
    cudaFree(x); cudaFree(w);
  }
  {  // NT embed: (B*T,256)x(256,256), softsign, bf16 out
    const int Me = Bt * T;
    bf16* x = dev_bf16((size_t)Me * D); bf16* w = dev_bf16((size_t)D * D); float* bias = dev_f32(D, true);
    GemmProblem p = base_problem(GEMM_NT, Me, D, D);
    p.A = {x, 0, 1, Me, D, D}; p.B = {w, 0, 1, D, D, D};
    p.epi.ldc = D; p.epi.bias = bias; p.epi.act = ACT_SOFTSIGN;
    run_case("NT embed softsign", p, (size_t)Me * D, 2.0 * Me * D * D);
    cudaFree(x); cudaFree(w); cudaFree(bias);
  }
  {  // stack projection forward: A = (Bt, T/4, 1024) shifted rows, 8 chunks
    const int R4 = T / 4, K4 = 4 * D;
    bf16* x = dev_bf16((size_t)Bt * T * D); bf16* w = dev_bf16((size_t)H * 8 * K4);
    GemmProblem p = base_problem(GEMM_NT, Tp, H, K4);
    p.nb_out = Bt; p.nchunk = 8; p.a_row_shift = 1; p.b_col_shift = K4;
    p.A = {x, (long long)T * D, Bt, R4, K4, K4}; p.B = {w, 0, 1, H, 8 * K4, 8 * K4};
    p.epi.ldc = H; p.epi.c_batch_stride = (long long)Tp * H;
    run_case("NT stack fwd", p, (size_t)M * H, 2.0 * M * H * 8 * K4);
    cudaFree(x); cudaFree(w);
  }
  {  // NN dgrad: dX = dY W
    bf16* dy = dev_bf16((size_t)M * H); bf16* w = dev_bf16((size_t)H * H);
    GemmProblem p = base_problem(GEMM_NN, M, H, H);
    p.A = {dy, 0, 1, M, H, H}; p.B = {w, 0, 1, H, H, H};
    p.epi.ldc = H;
    run_case("NN dgrad 1024", p, (size_t)M * H, 2.0 * M * H * H);
    cudaFree(dy); cudaFree(w);
  }
  {  // NN dgrad qkv: K=3072 -> N=1024
    bf16* dy = dev_bf16((size_t)M * 3 * H); bf16* w = dev_bf16((size_t)3 * H * H);
    GemmProblem p = base_problem(GEMM_NN, M, H, 3 * H);
    p.A = {dy, 0, 1, M, 3 * H, 3 * H}; p.B = {w, 0, 1, 3 * H, H, H};
    p.epi.ldc = H;
    run_case("NN dgrad qkv", p, (size_t)M * H, 2.0 * M * H * 3 * H);
    cudaFree(dy); cudaFree(w);
  }
  {  // NN stack dgrad (overlap-add): dX4[b,q,c] = sum_j dY[b,q-j,:] W[:, j*1024+c]
    const int R4 = T / 4, K4 = 4 * D;
    bf16* dy = dev_bf16((size_t)M * H); bf16* w = dev_bf16((size_t)H * 8 * K4);
    GemmProblem p = base_problem(GEMM_NN, R4, K4, H);
    p.nb_out = Bt; p.nchunk = 8; p.a_row_shift = -1; p.b_col_shift = K4;
    p.A = {dy, (long long)Tp * H, Bt, Tp, H, H}; p.B = {w, 0, 1, H, 8 * K4, 8 * K4};
    p.epi.ldc = K4; p.epi.c_batch_stride = (long long)T * D;
    run_case("NN stack dgrad", p, (size_t)Bt * T * D, 2.0 * Bt * R4 * K4 * 8 * H);
    cudaFree(dy); cudaFree(w);
  }
  {  // TN wgrad: dW[n,k] = sum_m dY[m,n] X[m,k], split-K accumulate
    bf16* dy = dev_bf16((size_t)M * H); bf16* x = dev_bf16((size_t)M * H);
    GemmProblem p = base_problem(GEMM_TN, H, H, M);
    p.A = {dy, 0, 1, M, H, H}; p.B = {x, 0, 1, M, H, H};
    p.epi.ldc = H; p.epi.accumulate = 1; p.split_k = 4;
    run_case("TN wgrad split4", p, (size_t)H * H, 2.0 * M * H * H);
    cudaFree(dy); cudaFree(x);
  }
  {  // TN stack wgrad: dW[h, j*1024+c] = sum_b sum_r dY[b,r,h] X4[b,r+j,c]
    const int R4 = T / 4, K4 = 4 * D;
    bf16* dy = dev_bf16((size_t)M * H); bf16* x = dev_bf16((size_t)Bt * T * D);
    GemmProblem p = base_problem(GEMM_TN, H, 8 * K4, Tp);
    p.nchunk = Bt; p.b_chunk_n = K4; p.b_row_shift = 1;
    p.A = {dy, (long long)Tp * H, Bt, Tp, H, H}; p.B = {x, (long long)T * D, Bt, R4, K4, K4};
    p.epi.ldc = 8 * K4; p.epi.accumulate = 1; p.split_k = 1;
    run_case("TN stack wgrad", p, (size_t)H * 8 * K4, 2.0 * M * H * 8 * K4);
    cudaFree(dy); cudaFree(x);
  }
  {  // TN head wgrad: dW[41,1024] = dlogits(M x 64 padded)^T hn
    bf16* dl = dev_bf16((size_t)M * 64); bf16* x = dev_bf16((size_t)M * H);
    GemmProblem p = base_problem(GEMM_TN, 41, H, M);
    p.A = {dl, 0, 1, M, 64, 64}; p.B = {x, 0, 1, M, H, H};
    p.epi.ldc = H; p.epi.accumulate = 1; p.split_k = 8;
    run_case("TN head wgrad", p, (size_t)41 * H, 2.0 * M * 41 * H);
    cudaFree(dl); cudaFree(x);
  }
  {  // epilogue features: dropout + residual + gather (fp32 out), then dact/drop backward
    bf16* x = dev_bf16((size_t)M * H); bf16* w = dev_bf16((size_t)H * H);
    float* bias = dev_f32(H, true); float* resid = dev_f32((size_t)M * H, true); float* tab = dev_f32((size_t)Tp * H, true);
    std::vector<long long> hidx(M); for (int i = 0; i < M; ++i) hidx[i] = i % Tp;
    long long* idx; CK(cudaMalloc(&idx, M * 8)); CK(cudaMemcpy(idx, hidx.data(), M * 8, cudaMemcpyHostToDevice));
    GemmProblem p = base_problem(GEMM_NT, M, H, H);
    p.A = {x, 0, 1, M, H, H}; p.B = {w, 0, 1, H, H, H};
    p.epi.ldc = H; p.epi.bias = bias; p.epi.resid = resid; p.epi.gather_tab = tab; p.epi.gather_idx = idx; p.epi.gather_ld = H;
    p.epi.drop_p = 0.4f; p.epi.drop_seed = 77; p.epi.drop_stream = 3;
    run_case("NT drop+resid+gather", p, (size_t)M * H, 2.0 * M * H * H);
    GemmProblem pb = base_problem(GEMM_NN, M, H, H);
    pb.A = {x, 0, 1, M, H, H}; pb.B = {w, 0, 1, H, H, H};
    pb.epi.ldc = H; pb.epi.dact = DACT_GELU_FROM_IN; pb.epi.dact_in = resid; pb.epi.drop_p = 0.4f; pb.epi.drop_bwd = 1; pb.epi.drop_seed = 5; pb.epi.drop_stream = 9;
    run_case("NN dgelu+dropmask bf16", pb, (size_t)M * H, 2.0 * M * H * H, true);
    cudaFree(x); cudaFree(w); cudaFree(bias); cudaFree(resid); cudaFree(tab); cudaFree(idx);
  }
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
