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

// Programmatic dependent launch: a kernel launched through oct_launch() may become resident while the previous kernel of
// the stream (or captured graph) is still draining; it runs its prologue (mbarrier init, TMEM allocation, tensor-map
// prefetch) and blocks in pdl_wait() until that kernel has completed and its writes are visible.  Every kernel signals
// pdl_launch_dependents() first thing, so the overlap window is the tail of the grid.  OFF by default (OCT_PDL=1 enables
// the launch attribute; without it the device-side instructions are no-ops): measured on the training step it LOSES 3 %
// (36.2 vs 35.0 ms) because the early-resident CTAs of the next kernel in the chain take the SMs that the side-stream
// weight-gradient GEMMs (ops.wgrad_bias_async) would otherwise fill.
bool oct_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t oct_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (oct_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// Fast GELU for the tensor-core epilogues.  erfc via Abramowitz & Stegun 7.1.26 (|abs err| < 1.5e-7, two orders of
// magnitude below bf16 rounding) with approximate SFU reciprocal / exp.  Written for instruction count: the epilogue warps
// of the GELU GEMMs are issue- and XU-pipe-bound (profiles/r1_gemm_ncu.md), so the 1/2 is folded into the coefficients,
// the sign select is replaced by max(x,0) - |x|*tail, and the derivative shares the single exponential
// (exp(-(x/sqrt2)^2) == exp(-x^2/2)).  The fp32 parity path keeps erff (gelu_erf / gelu_erf_grad).
//   tail = Phi(-|x|) = 0.5*erfc(|x|/sqrt2),   e = exp(-x^2/2)
__device__ __forceinline__ void gelu_fast_parts(float x, float& tail, float& e) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (x * -0.72134752044448170368f)));
  tail = poly * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float tail, e;
  gelu_fast_parts(x, tail, e);
  return fmaf(-fabsf(x), tail, fmaxf(x, 0.f));  // x >= 0: x - x*tail = x*Phi(x);  x < 0: x*tail = x*Phi(x)
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
  float tail, e;
  gelu_fast_parts(x, tail, e);
  const float cdf = (x >= 0.f) ? (1.f - tail) : tail;
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
