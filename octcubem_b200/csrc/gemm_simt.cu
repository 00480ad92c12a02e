// fp32 CUDA-core GEMM: the 1e-4 parity mode.  The reference's flash model cannot run fp32 at all
// (flash_attn/modules/mha.py:100 asserts half), so fp32 parity is defined against the CPU oracle; this kernel
// gives exact-fp32 products/accumulation (FFMA) for that mode.  The throughput path is gemm_tc.cu (tcgen05).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;  // 256 threads, 4x4 outputs each

struct SimtGemmArgs {
  const float* A; const float* B; float* D;
  int64_t M, N, K;
  int64_t a_sm, a_sk;  // A(m,k) = A[m*a_sm + k*a_sk]
  int64_t b_sn, b_sk;  // B(n,k) = B[n*b_sn + k*b_sk]
  int64_t ldd;
  int epilogue; const float* bias; float* aux; int beta;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtGemmArgs p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // loader mapping: choose the fastest-varying thread index along the contiguous global dimension
  const bool a_k_contig = (p.a_sk == 1), b_k_contig = (p.b_sk == 1);
  for (int64_t k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      const int e = tid + i * 256;
      const int kk = a_k_contig ? (e % BK) : (e / BM);
      const int mm = a_k_contig ? (e / BK) : (e % BM);
      const int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < p.K) ? p.A[m * p.a_sm + k * p.a_sk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      const int e = tid + i * 256;
      const int kk = b_k_contig ? (e % BK) : (e / BN);
      const int nn = b_k_contig ? (e / BK) : (e % BN);
      const int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < p.K) ? p.B[n * p.b_sn + k * p.b_sk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      float* d = p.D + m * p.ldd + n;
      switch (p.epilogue) {
        case OCT_EPI_BIAS: v += p.bias[n]; break;
        case OCT_EPI_BIAS_GELU: v += p.bias[n]; p.aux[m * p.ldd + n] = v; v = gelu_erf(v); break;
        case OCT_EPI_DGELU: v *= gelu_erf_grad(p.aux[m * p.ldd + n]); break;
        default: if (p.beta) v += *d; break;
      }
      *d = v;
    }
  }
}

}  // namespace

int oct_gemm_simt_f32(int layout, const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t K, int64_t lda,
                      int64_t ldb, int64_t ldd, int epilogue, const float* bias, float* aux, int beta,
                      cudaStream_t stream) {
  SimtGemmArgs p;
  p.A = A; p.B = B; p.D = D; p.M = M; p.N = N; p.K = K; p.ldd = ldd;
  p.epilogue = epilogue; p.bias = bias; p.aux = aux; p.beta = beta;
  switch (layout) {
    case OCT_GEMM_NT: p.a_sm = lda; p.a_sk = 1; p.b_sn = ldb; p.b_sk = 1; break;   // A[M,K], B[N,K]
    case OCT_GEMM_NN: p.a_sm = lda; p.a_sk = 1; p.b_sn = 1; p.b_sk = ldb; break;   // A[M,K], B[K,N]
    case OCT_GEMM_TN: p.a_sm = 1; p.a_sk = lda; p.b_sn = 1; p.b_sk = ldb; break;   // A[K,M], B[K,N]
    default: oct_set_error("oct_gemm: bad layout %d", layout); return OCT_ERR_INVALID;
  }
  if (M == 0 || N == 0) return OCT_OK;
  dim3 grid((unsigned)ceil_div64(N, BN), (unsigned)ceil_div64(M, BM));
  OCT_REQUIRE(grid.y <= 65535, "oct_gemm(f32): M too large");
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
  return oct_check_launch("oct_gemm(f32)");
}
