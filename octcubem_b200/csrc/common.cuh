// Shared device/host helpers for the octcube_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/octcube_b200.h"

// ------------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message, SURVEY §8b "Errors")
// ------------------------------------------------------------------------------------------------
void oct_set_error(const char* fmt, ...);
int oct_check_launch(const char* what);  // cudaGetLastError() -> code, message

#define OCT_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      oct_set_error(__VA_ARGS__);              \
      return OCT_ERR_INVALID;                  \
    }                                          \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int oct_num_sms();  // cached SM count of the current device (148 on B200)

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); result broadcast to all threads
template <int kMaxWarps = 32>
__device__ __forceinline__ float block_sum(float v, float* smem /* >= kMaxWarps floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? smem[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// exact (erf) GELU and its derivative — nn.GELU() default, flash_attn/modules/mlp.py:49
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast GELU for the tensor-core epilogues.  erf via Abramowitz & Stegun 7.1.26 (|abs err| < 1.5e-7, two orders of
// magnitude below bf16 rounding) with approximate SFU reciprocal / exp: 2 SFU ops + ~12 FMA-pipe ops instead of erff's
// ~40 instructions.  The derivative shares the single exponential: exp(-(x/sqrt2)^2) == exp(-x^2/2).
// The fp32 parity path keeps erff (gelu_erf / gelu_erf_grad).
__device__ __forceinline__ void gelu_fast_parts(float x, float& cdf, float& e) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));  // exp(-x^2/2)
  const float half_tail = 0.5f * poly * e;            // 0.5 * (1 - erf(|x|/sqrt2))
  cdf = (x >= 0.f) ? (1.f - half_tail) : half_tail;   // Phi(x)
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, e;
  gelu_fast_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
  float cdf, e;
  gelu_fast_parts(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// generic typed load/store to float
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4-wide vector access: float4 for fp32, 8 bytes for bf16
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

#endif  // __CUDACC__
