// Token pooling for the encoder-only ViT (OCTCube/models_vit_st_flash_attn.py:247-251): the mean over a row range
// [row0, row1) of every sample — `x[:, 1:, :].mean(dim=1)` (global pool, cls skipped) or `x[:, 0]` (range [0, 1)) — and
// its backward.  HBM-bound: the rows are read once.
#include "common.cuh"

constexpr int kPoolThreads = 256;
constexpr int kPoolRowsPerCta = 64;  // S = 5121, C = 1024: 81 chunks x (C / 1024) x B CTAs keep every SM busy

// stage 1: ws[b, chunk, c] = sum of the chunk's rows (fp32, fixed order)
template <typename T>
__global__ void __launch_bounds__(kPoolThreads) mean_pool_partial_kernel(const T* __restrict__ x, float* __restrict__ ws,
                                                                        int S, int C, int row0, int row1, int n_chunks) {
  const int chunk = blockIdx.x, b = blockIdx.z;
  const int c = (blockIdx.y * kPoolThreads + threadIdx.x) * 4;
  if (c >= C) return;
  const int s0 = row0 + chunk * kPoolRowsPerCta;
  const int s1 = min(s0 + kPoolRowsPerCta, row1);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const T* p = x + ((size_t)b * S + s0) * C + c;
#pragma unroll 4
  for (int s = s0; s < s1; ++s, p += C) {
    const float4 v = Vec4<T>::ld(p);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4*>(ws + ((size_t)b * n_chunks + chunk) * C + c) = a;
}

// stage 2: out[b, c] = (sum over chunks) / n
template <typename T>
__global__ void __launch_bounds__(kPoolThreads) mean_pool_finish_kernel(const float* __restrict__ ws, T* __restrict__ out, int C,
                                                                       int n_chunks, float inv_n) {
  const int b = blockIdx.y;
  const int c = (blockIdx.x * kPoolThreads + threadIdx.x) * 4;
  if (c >= C) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = ws + (size_t)b * n_chunks * C + c;
  for (int k = 0; k < n_chunks; ++k, p += C) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  a.x *= inv_n; a.y *= inv_n; a.z *= inv_n; a.w *= inv_n;
  Vec4<T>::st(out + (size_t)b * C + c, a);
}

// backward: dx[b, s, :] = dout[b, :] / n for s in [row0, row1), 0 for the other rows; one CTA per row
template <typename TG, typename T>
__global__ void mean_pool_bwd_kernel(const TG* __restrict__ dout, T* __restrict__ dx, int S, int C, int row0, int row1,
                                     float inv_n) {
  const int s = blockIdx.x, b = blockIdx.y;
  T* o = dx + ((size_t)b * S + s) * C;
  const TG* g = dout + (size_t)b * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= row0 && s < row1) {
      v = Vec4<TG>::ld(g + c);
      v.x *= inv_n; v.y *= inv_n; v.z *= inv_n; v.w *= inv_n;
    }
    Vec4<T>::st(o + c, v);
  }
}

static int64_t pool_chunks(int64_t row0, int64_t row1) { return ceil_div64(row1 - row0, kPoolRowsPerCta); }

extern "C" size_t oct_mean_pool_ws_bytes(int64_t B, int64_t C, int64_t row0, int64_t row1) {
  if (B <= 0 || row1 <= row0 || C <= 0) return 0;
  return (size_t)B * pool_chunks(row0, row1) * C * sizeof(float);
}

extern "C" int oct_mean_pool_fwd(const void* x, int x_dtype, void* out, int out_dtype, int64_t B, int64_t S, int64_t C,
                                 int64_t row0, int64_t row1, void* ws, size_t ws_bytes, oct_stream_t stream) {
  OCT_REQUIRE(x && out, "oct_mean_pool_fwd: null pointer");
  OCT_REQUIRE(C % 4 == 0 && row0 >= 0 && row1 > row0 && row1 <= S, "oct_mean_pool_fwd: need C%%4==0 and 0 <= row0 < row1 <= S");
  OCT_REQUIRE(B <= 65535, "oct_mean_pool_fwd: B too large");
  OCT_REQUIRE((x_dtype == OCT_F32 || x_dtype == OCT_BF16) && (out_dtype == OCT_F32 || out_dtype == OCT_BF16),
              "oct_mean_pool_fwd: bad dtype");
  if (!ws || ws_bytes < oct_mean_pool_ws_bytes(B, C, row0, row1)) {
    oct_set_error("oct_mean_pool_fwd: workspace too small");
    return OCT_ERR_WORKSPACE;
  }
  if (B == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_chunks = (int)pool_chunks(row0, row1);
  const unsigned cb = (unsigned)ceil_div64(C / 4, kPoolThreads);
  const float inv_n = 1.f / (float)(row1 - row0);
  dim3 g1((unsigned)n_chunks, cb, (unsigned)B), g2(cb, (unsigned)B);
  if (x_dtype == OCT_F32)
    mean_pool_partial_kernel<float><<<g1, kPoolThreads, 0, st>>>((const float*)x, (float*)ws, (int)S, (int)C, (int)row0,
                                                                 (int)row1, n_chunks);
  else
    mean_pool_partial_kernel<__nv_bfloat16><<<g1, kPoolThreads, 0, st>>>((const __nv_bfloat16*)x, (float*)ws, (int)S, (int)C,
                                                                         (int)row0, (int)row1, n_chunks);
  int rc = oct_check_launch("oct_mean_pool_fwd(partial)");
  if (rc) return rc;
  if (out_dtype == OCT_F32)
    mean_pool_finish_kernel<float><<<g2, kPoolThreads, 0, st>>>((const float*)ws, (float*)out, (int)C, n_chunks, inv_n);
  else
    mean_pool_finish_kernel<__nv_bfloat16><<<g2, kPoolThreads, 0, st>>>((const float*)ws, (__nv_bfloat16*)out, (int)C,
                                                                        n_chunks, inv_n);
  return oct_check_launch("oct_mean_pool_fwd(finish)");
}

extern "C" int oct_mean_pool_bwd(const void* dout, int dout_dtype, void* dx, int dx_dtype, int64_t B, int64_t S, int64_t C,
                                 int64_t row0, int64_t row1, oct_stream_t stream) {
  OCT_REQUIRE(dout && dx, "oct_mean_pool_bwd: null pointer");
  OCT_REQUIRE(C % 4 == 0 && row0 >= 0 && row1 > row0 && row1 <= S, "oct_mean_pool_bwd: need C%%4==0 and 0 <= row0 < row1 <= S");
  OCT_REQUIRE(B <= 65535, "oct_mean_pool_bwd: B too large");
  if (B == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = (int)((C / 4 < 256) ? ((C / 4 + 31) / 32 * 32) : 256);
  const float inv_n = 1.f / (float)(row1 - row0);
  dim3 grid((unsigned)S, (unsigned)B);
#define LAUNCH(TG, T)                                                                                                    \
  mean_pool_bwd_kernel<TG, T><<<grid, threads, 0, st>>>((const TG*)dout, (T*)dx, (int)S, (int)C, (int)row0, (int)row1, inv_n)
  if (dout_dtype == OCT_F32 && dx_dtype == OCT_F32) LAUNCH(float, float);
  else if (dout_dtype == OCT_F32 && dx_dtype == OCT_BF16) LAUNCH(float, __nv_bfloat16);
  else if (dout_dtype == OCT_BF16 && dx_dtype == OCT_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (dout_dtype == OCT_BF16 && dx_dtype == OCT_F32) LAUNCH(__nv_bfloat16, float);
  else OCT_REQUIRE(false, "oct_mean_pool_bwd: bad dtype");
#undef LAUNCH
  return oct_check_launch("oct_mean_pool_bwd");
}

// ---------------------------------------------------------------------------------------------------------------
// Sparse (ELL) matrix x dense table: y[r, :] = sum_k w[r, k] * x[idx[r, k], :]
// The bicubic 32x32 -> 16x16 resampling of the learnable spatial pos tables (models:419-421, 537-539) is a FIXED linear map
// with 16 taps per output cell (and its transpose, used by the backward, has at most 9 per input cell): as a dense fp32
// CUDA-core GEMM it cost 52 us per call, four calls per step (profiles/r2_step_launches.md).  fp32, fixed summation order.
// ---------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) ell_spmm_kernel(const int* __restrict__ idx, const float* __restrict__ w,
                                                       const float* __restrict__ x, float* __restrict__ y, int K, int C) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const float wk = __ldg(w + (size_t)r * K + k);
      if (wk != 0.f) {
        const float4 v = *reinterpret_cast<const float4*>(x + (size_t)__ldg(idx + (size_t)r * K + k) * C + c);
        acc.x = fmaf(wk, v.x, acc.x); acc.y = fmaf(wk, v.y, acc.y); acc.z = fmaf(wk, v.z, acc.z); acc.w = fmaf(wk, v.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(y + (size_t)r * C + c) = acc;
  }
}
}  // namespace

extern "C" int oct_ell_spmm(const int* idx, const float* w, const float* x, float* y, int64_t R, int64_t K, int64_t C,
                            oct_stream_t stream) {
  OCT_REQUIRE(idx && w && x && y, "oct_ell_spmm: null pointer");
  OCT_REQUIRE(R >= 0 && K >= 1 && C >= 4 && C % 4 == 0 && R < (1 << 30), "oct_ell_spmm: need K >= 1, C %% 4 == 0");
  OCT_REQUIRE(aligned16(x) && aligned16(y), "oct_ell_spmm: x / y must be 16-byte aligned");
  if (R == 0) return OCT_OK;
  const int threads = (int)((C / 4 < 256) ? ((C / 4 + 31) / 32 * 32) : 256);
  ell_spmm_kernel<<<(unsigned)R, threads, 0, (cudaStream_t)stream>>>(idx, w, x, y, (int)K, (int)C);
  return oct_check_launch("oct_ell_spmm");
}
