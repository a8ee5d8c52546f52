// Strict-mode GEMM: fp32 FMA on the CUDA cores, fp32 (or bf16) operands,
// fp32 accumulate.  Used (a) as the arithmetic of the fp32 parity mode
// (BASELINE.json north_star: loss/grads within 1e-4 of the fp32 reference,
// which single-pass TF32/bf16 tensor-core products cannot reach)
// and (b) as the on-device cross-check of the tcgen05 kernel in gemm_tc.cu.
// Same GemmProblem semantics as the tensor-core kernel (gemm_common.cuh).
#include "gemm_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T>
__device__ __forceinline__ float ld_elem(const GemmOperand& o, int b, int row, int col) {
  if (row < 0 || row >= o.rows || col < 0 || col >= o.cols) return 0.f;
  return to_f32(((const T*)o.ptr)[(long long)b * o.batch_stride + (long long)row * o.ld + col]);
}

template <typename T>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmProblem p) { pdl_grid_sync();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, 4x4 outputs each
  const int m_tiles = (p.M + BM - 1) / BM;
  const int bt = blockIdx.y / m_tiles;     // output trial
  const int m0 = (blockIdx.y % m_tiles) * BM;
  const int n0 = blockIdx.x * BN;
  int bsel = 0;                            // B operand batch (per-day weights)
  if (p.b_sel) { const long long d = p.b_sel[bt]; bsel = (int)(d < 0 ? 0 : (d >= p.B.nbatch ? p.B.nbatch - 1 : d)); }

  // reduction range of this CTA (split only in GEMM_TN)
  const int kblocks_per_chunk = (p.chunk_k + BK - 1) / BK;
  const int total_kb = p.nchunk * kblocks_per_chunk;
  int kb_begin = 0, kb_end = total_kb;
  if (p.split_k > 1) {
    const int per = (total_kb + p.split_k - 1) / p.split_k;
    kb_begin = blockIdx.z * per;
    kb_end = min(total_kb, kb_begin + per);
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kb = kb_begin; kb < kb_end; ++kb) {
    const int j = kb / kblocks_per_chunk;
    const int kc0 = (kb % kblocks_per_chunk) * BK;
    if (p.mode == GEMM_TN) {
      // tiles are [BK reduce rows][64 contiguous columns]
      const int kr = tid / 16, c4 = (tid % 16) * 4;
      const int kc = kc0 + kr;
      const bool kin = kc < p.chunk_k;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n1 = m0 + c4 + q;
        As[kr][c4 + q] = (kin && n1 < p.M) ? ld_elem<T>(p.A, j, kc + p.a_row_shift, n1) : 0.f;
        const int n2 = n0 + c4 + q;
        float bv = 0.f;
        if (kin && n2 < p.N) {
          const int cn = n2 / p.b_chunk_n;
          bv = ld_elem<T>(p.B, j, kc + cn * p.b_row_shift, n2 - cn * p.b_chunk_n);
        }
        Bs[kr][c4 + q] = bv;
      }
    } else {
      {  // A tile: 64 rows x 16 k, k contiguous in memory
        const int row = tid / 4, k4 = (tid % 4) * 4;
        const int r = m0 + row;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kc = kc0 + k4 + q;
          As[k4 + q][row] = (kc < p.chunk_k && r < p.M)
                                ? ld_elem<T>(p.A, bt, r + j * p.a_row_shift, kc + j * p.a_col_shift)
                                : 0.f;
        }
      }
      if (p.mode == GEMM_NT) {
        const int row = tid / 4, k4 = (tid % 4) * 4;
        const int n = n0 + row;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kc = kc0 + k4 + q;
          Bs[k4 + q][row] = (kc < p.chunk_k && n < p.N)
                                ? ld_elem<T>(p.B, bsel, n + j * p.b_row_shift, kc + j * p.b_col_shift)
                                : 0.f;
        }
      } else {  // GEMM_NN: B is [reduce rows][n contiguous]
        const int kr = tid / 16, c4 = (tid % 16) * 4;
        const int kc = kc0 + kr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n0 + c4 + q;
          Bs[kr][c4 + q] = (kc < p.chunk_k && n < p.N)
                               ? ld_elem<T>(p.B, 0, kc + j * p.b_row_shift, n + j * p.b_col_shift)
                               : 0.f;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = Bs[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(a[i], b[jn], acc[i][jn]);
    }
    __syncthreads();
  }

  if (kb_begin >= kb_end && p.split_k > 1) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= p.M) continue;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int n = n0 + tx * 4 + jn;
      if (n >= p.N) continue;
      gemm_epilogue_store(p.epi, p.N, p.M, acc[i][jn], (p.mode == GEMM_TN && p.epi.sel) ? (int)blockIdx.z : bt, r, n);
    }
  }
}

}  // namespace

int gemm_simt_launch(const GemmProblem& p, int in_bf16, cudaStream_t stream) {
  NDT1_REQUIRE(p.M > 0 && p.N > 0 && p.nb_out > 0, "gemm_simt: empty problem M=%d N=%d nb=%d", p.M, p.N, p.nb_out);
  NDT1_REQUIRE(p.split_k <= 1 || p.epi.accumulate, "gemm_simt: split_k needs an accumulating epilogue");
  dim3 grid(ndt1_cdiv(p.N, BN), ndt1_cdiv(p.M, BM) * p.nb_out, p.split_k > 1 ? p.split_k : 1);
  if (in_bf16) ndt1_launch(gemm_simt_kernel<bf16>, grid, NT, 0, stream, p);
  else ndt1_launch(gemm_simt_kernel<float>, grid, NT, 0, stream, p);
  NDT1_CHECK_LAUNCH();
  return 0;
}
