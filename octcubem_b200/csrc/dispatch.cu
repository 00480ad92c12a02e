// C-ABI dispatchers that pick the compute path (fp32 CUDA-core parity kernels vs tcgen05 tensor-core kernels).
#include "common.cuh"

int oct_gemm_simt_f32(int layout, const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t K, int64_t lda,
                      int64_t ldb, int64_t ldd, int epilogue, const float* bias, float* aux, int beta, cudaStream_t stream);
int oct_gemm_tc_bf16(int layout, const void* A, const void* B, void* D, int d_dtype, int64_t M, int64_t N, int64_t K,
                     int64_t lda, int64_t ldb, int64_t ldd, int epilogue, const float* bias, void* aux, int beta,
                     float* colsum, cudaStream_t st);
int oct_attn_fwd_simt(int io_dtype, const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, int64_t d,
                      float scale, cudaStream_t st);
int oct_attn_bwd_simt(int io_dtype, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      float* delta, int64_t B, int64_t S, int64_t H, int64_t d, float scale, cudaStream_t st);
int oct_attn_fwd_tc(const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, int64_t d, float scale,
                    cudaStream_t st);
int oct_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* ws,
                    size_t ws_bytes, int64_t B, int64_t S, int64_t H, int64_t d, float scale, cudaStream_t st);
size_t oct_attn_bwd_tc_ws_bytes(int64_t B, int64_t S, int64_t H, int64_t d);

extern "C" int oct_gemm(int compute, int layout, const void* A, const void* B, void* D, int d_dtype, int64_t M, int64_t N,
                        int64_t K, int64_t lda, int64_t ldb, int64_t ldd, int epilogue, const float* bias, void* aux,
                        int beta, oct_stream_t stream) {
  OCT_REQUIRE(A && B && D, "oct_gemm: null pointer");
  OCT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "oct_gemm: negative dimension");
  OCT_REQUIRE(epilogue >= OCT_EPI_NONE && epilogue <= OCT_EPI_DGELU, "oct_gemm: bad epilogue %d", epilogue);
  OCT_REQUIRE((epilogue != OCT_EPI_BIAS && epilogue != OCT_EPI_BIAS_GELU) || bias, "oct_gemm: epilogue needs bias");
  OCT_REQUIRE((epilogue != OCT_EPI_BIAS_GELU && epilogue != OCT_EPI_DGELU) || aux, "oct_gemm: epilogue needs aux");
  OCT_REQUIRE(beta == 0 || (beta == 1 && d_dtype == OCT_F32 && epilogue == OCT_EPI_NONE),
              "oct_gemm: beta=1 needs fp32 D and EPI_NONE");
  if (compute == OCT_F32) {
    OCT_REQUIRE(d_dtype == OCT_F32, "oct_gemm(f32): D must be fp32");
    return oct_gemm_simt_f32(layout, (const float*)A, (const float*)B, (float*)D, M, N, K, lda, ldb, ldd, epilogue, bias,
                             (float*)aux, beta, (cudaStream_t)stream);
  }
  if (compute == OCT_BF16) {
    OCT_REQUIRE(d_dtype == OCT_F32 || d_dtype == OCT_BF16, "oct_gemm(bf16): bad D dtype");
    return oct_gemm_tc_bf16(layout, A, B, D, d_dtype, M, N, K, lda, ldb, ldd, epilogue, bias, aux, beta, nullptr,
                            (cudaStream_t)stream);
  }
  OCT_REQUIRE(false, "oct_gemm: bad compute mode %d", compute);
}

extern "C" int oct_gemm_wgrad_bias(int compute, const void* dy, const void* x, float* dw, float* db, int64_t n_out,
                                   int64_t k_in, int64_t tokens, int64_t ld_dy, int64_t ld_x, int64_t ld_dw, int beta,
                                   oct_stream_t stream) {
  OCT_REQUIRE(dy && x && dw && db, "oct_gemm_wgrad_bias: null pointer");
  OCT_REQUIRE(n_out >= 0 && k_in >= 0 && tokens > 0, "oct_gemm_wgrad_bias: bad dimension");
  OCT_REQUIRE(beta == 0 || beta == 1, "oct_gemm_wgrad_bias: beta must be 0 or 1");
  OCT_REQUIRE(compute == OCT_BF16, "oct_gemm_wgrad_bias: only the bf16 tensor-core path fuses the bias gradient "
                                   "(fp32 mode: oct_gemm + oct_colsum)");
  return oct_gemm_tc_bf16(OCT_GEMM_TN, dy, x, dw, OCT_F32, n_out, k_in, tokens, ld_dy, ld_x, ld_dw, OCT_EPI_NONE, nullptr,
                          nullptr, beta, db, (cudaStream_t)stream);
}

extern "C" int oct_attn_fwd(int compute, const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H,
                            int64_t d, float scale, oct_stream_t stream) {
  OCT_REQUIRE(qkv && out && lse, "oct_attn_fwd: null pointer");
  OCT_REQUIRE(B >= 0 && S > 0 && H > 0 && H <= 65535 && B <= 65535, "oct_attn_fwd: bad sizes");
  if (B == 0) return OCT_OK;
  if (compute == OCT_F32) return oct_attn_fwd_simt(OCT_F32, qkv, out, lse, B, S, H, d, scale, (cudaStream_t)stream);
  if (compute == OCT_SIMT_BF16) return oct_attn_fwd_simt(OCT_BF16, qkv, out, lse, B, S, H, d, scale, (cudaStream_t)stream);
  if (compute == OCT_BF16) return oct_attn_fwd_tc(qkv, out, lse, B, S, H, d, scale, (cudaStream_t)stream);
  OCT_REQUIRE(false, "oct_attn_fwd: bad compute mode %d", compute);
}

extern "C" size_t oct_attn_bwd_ws_bytes(int compute, int64_t B, int64_t S, int64_t H, int64_t d) {
  if (compute == OCT_BF16) return oct_attn_bwd_tc_ws_bytes(B, S, H, d);
  return (size_t)B * H * S * sizeof(float);  // delta
}

extern "C" int oct_attn_bwd(int compute, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                            void* ws, size_t ws_bytes, int64_t B, int64_t S, int64_t H, int64_t d, float scale,
                            oct_stream_t stream) {
  OCT_REQUIRE(qkv && out && dout && lse && dqkv, "oct_attn_bwd: null pointer");
  OCT_REQUIRE(B >= 0 && S > 0 && H > 0 && H <= 65535 && B <= 65535, "oct_attn_bwd: bad sizes");
  if (!ws || ws_bytes < oct_attn_bwd_ws_bytes(compute, B, S, H, d)) {
    oct_set_error("oct_attn_bwd: workspace too small");
    return OCT_ERR_WORKSPACE;
  }
  if (B == 0) return OCT_OK;
  if (compute == OCT_F32)
    return oct_attn_bwd_simt(OCT_F32, qkv, out, dout, lse, dqkv, (float*)ws, B, S, H, d, scale, (cudaStream_t)stream);
  if (compute == OCT_SIMT_BF16)
    return oct_attn_bwd_simt(OCT_BF16, qkv, out, dout, lse, dqkv, (float*)ws, B, S, H, d, scale, (cudaStream_t)stream);
  if (compute == OCT_BF16) return oct_attn_bwd_tc(qkv, out, dout, lse, dqkv, ws, ws_bytes, B, S, H, d, scale, (cudaStream_t)stream);
  OCT_REQUIRE(false, "oct_attn_bwd: bad compute mode %d", compute);
}
