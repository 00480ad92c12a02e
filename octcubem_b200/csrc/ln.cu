// Residual add + LayerNorm forward/backward, GELU, column sums, casts: the HBM-bound stages between the GEMMs.
//   oct_add_ln_fwd/bwd : flash_attn/modules/block.py:126-130,163-167 (dropout p=0 + add + LN, fp32 residual) and the
//                        final norms models_mae_joint_res_flash_attn.py:489,592
// One warp owns one row; the row lives in registers between the statistics passes (read once, written once).
#include "common.cuh"
#include <cstdlib>

constexpr int kLnWarps = 4;  // warps per CTA

template <typename T> struct LpLoad4 {  // low-precision or fp32 row chunk -> float4
  static __device__ __forceinline__ float4 ld(const void* base, size_t idx) { return Vec4<T>::ld((const T*)base + idx); }
  static __device__ __forceinline__ void st(void* base, size_t idx, float4 v) { Vec4<T>::st((T*)base + idx, v); }
};

template <int NV, typename TH, typename TY>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_fwd_kernel(const TH* __restrict__ h,
                                                                   const float* __restrict__ res_in,
                                                                   float* __restrict__ res_out,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, TY* __restrict__ y,
                                                                   float* __restrict__ mean_out,
                                                                   float* __restrict__ rstd_out, int64_t M, int C,
                                                                   float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * kLnWarps;
  const int C4 = C >> 2;
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5); row < M; row += warps_total) {
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < C4) {
        float4 a = Vec4<TH>::ld(h + row * C + c4 * 4);
        if (res_in) {
          float4 r = *reinterpret_cast<const float4*>(res_in + row * C + c4 * 4);
          a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
        }
        if (res_out) *reinterpret_cast<float4*>(res_out + row * C + c4 * 4) = a;
        v[k] = a;
        sum += (a.x + a.y) + (a.z + a.w);
      } else {
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < C4) {
        const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
        sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float var = warp_sum(sq) / (float)C;
    const float rstd = rsqrtf(var + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < C4) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + c4 * 4);
        const float4 b = *reinterpret_cast<const float4*>(beta + c4 * 4);
        float4 o;
        o.x = (v[k].x - mean) * rstd * g.x + b.x;
        o.y = (v[k].y - mean) * rstd * g.y + b.y;
        o.z = (v[k].z - mean) * rstd * g.z + b.z;
        o.w = (v[k].w - mean) * rstd * g.w + b.w;
        Vec4<TY>::st(y + row * C + c4 * 4, o);
      }
    }
  }
}

template <int NV, typename TH, typename TY>
static int launch_ln_fwd(const void* h, const float* res_in, float* res_out, const float* gamma, const float* beta,
                         void* y, float* mean, float* rstd, int64_t M, int C, float eps, cudaStream_t st) {
  int64_t blocks = ceil_div64(M, kLnWarps);
  const int64_t cap = (int64_t)oct_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  oct_launch(add_ln_fwd_kernel<NV, TH, TY>, dim3((unsigned)blocks), dim3(kLnWarps * 32), 0, st, 1, (const TH*)h, res_in, res_out,
             gamma, beta, (TY*)y, mean, rstd, M, C, eps);
  return oct_check_launch("oct_add_ln_fwd");
}

template <typename TH, typename TY>
static int dispatch_ln_fwd_nv(int nv, const void* h, const float* res_in, float* res_out, const float* gamma,
                              const float* beta, void* y, float* mean, float* rstd, int64_t M, int C, float eps,
                              cudaStream_t st) {
  switch (nv) {
    case 1: return launch_ln_fwd<1, TH, TY>(h, res_in, res_out, gamma, beta, y, mean, rstd, M, C, eps, st);
    case 2: return launch_ln_fwd<2, TH, TY>(h, res_in, res_out, gamma, beta, y, mean, rstd, M, C, eps, st);
    case 4: return launch_ln_fwd<4, TH, TY>(h, res_in, res_out, gamma, beta, y, mean, rstd, M, C, eps, st);
    case 8: return launch_ln_fwd<8, TH, TY>(h, res_in, res_out, gamma, beta, y, mean, rstd, M, C, eps, st);
    default: return launch_ln_fwd<16, TH, TY>(h, res_in, res_out, gamma, beta, y, mean, rstd, M, C, eps, st);
  }
}

static int nv_for(int64_t C) {
  const int64_t need = ceil_div64(C / 4, 32);
  int nv = 1;
  while (nv < need) nv <<= 1;
  return nv;
}

extern "C" int oct_add_ln_fwd(const void* h, int h_dtype, const float* res_in, float* res_out, const float* gamma,
                              const float* beta, void* y, int y_dtype, float* mean, float* rstd, int64_t M, int64_t C,
                              float eps, oct_stream_t stream) {
  OCT_REQUIRE(h && gamma && beta && y && mean && rstd, "oct_add_ln_fwd: null pointer");
  OCT_REQUIRE(C > 0 && C % 4 == 0 && C <= 2048, "oct_add_ln_fwd: need C%%4==0 and C<=2048 (got %lld)", (long long)C);
  if (M == 0) return OCT_OK;
  const int nv = nv_for(C);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_ARGS h, res_in, res_out, gamma, beta, y, mean, rstd, M, (int)C, eps, st
  if (h_dtype == OCT_F32 && y_dtype == OCT_F32) return dispatch_ln_fwd_nv<float, float>(nv, LN_ARGS);
  if (h_dtype == OCT_F32 && y_dtype == OCT_BF16) return dispatch_ln_fwd_nv<float, __nv_bfloat16>(nv, LN_ARGS);
  if (h_dtype == OCT_BF16 && y_dtype == OCT_F32) return dispatch_ln_fwd_nv<__nv_bfloat16, float>(nv, LN_ARGS);
  if (h_dtype == OCT_BF16 && y_dtype == OCT_BF16) return dispatch_ln_fwd_nv<__nv_bfloat16, __nv_bfloat16>(nv, LN_ARGS);
#undef LN_ARGS
  OCT_REQUIRE(false, "oct_add_ln_fwd: bad dtype");
}

// ------------------------------------------------------------------------------------------------
// backward.  dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma ;  dgamma = sum dy*xhat ; dbeta = sum dy.
// Per-warp register partials of dgamma/dbeta over a grid-stride set of rows -> smem reduce per CTA -> ws[cta][2][C]
// -> second kernel sums the CTAs in order (deterministic).
// ------------------------------------------------------------------------------------------------
template <int NV, typename TDY, typename TX, typename TLP>
__global__ void __launch_bounds__(kLnWarps * 32) add_ln_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ dres_in,
                                                                   float* __restrict__ dx_f32, TLP* __restrict__ dx_lp,
                                                                   float* __restrict__ ws, int64_t M, int C) {
  extern __shared__ float sred[];  // [kLnWarps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * kLnWarps;
  const int C4 = C >> 2;
  float4 dg[NV], db[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < M; row += warps_total) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < C4) {
        const float4 d = Vec4<TDY>::ld(dy + row * C + c4 * 4);
        const float4 xv = Vec4<TX>::ld(x + row * C + c4 * 4);
        const float4 gm = *reinterpret_cast<const float4*>(gamma + c4 * 4);
        xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[k] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        s1 += (g[k].x + g[k].y) + (g[k].z + g[k].w);
        s2 += (g[k].x * xh[k].x + g[k].y * xh[k].y) + (g[k].z * xh[k].z + g[k].w * xh[k].w);
        dg[k].x += d.x * xh[k].x; dg[k].y += d.y * xh[k].y; dg[k].z += d.z * xh[k].z; dg[k].w += d.w * xh[k].w;
        db[k].x += d.x; db[k].y += d.y; db[k].z += d.z; db[k].w += d.w;
      }
    }
    const float c1 = warp_sum(s1) / (float)C, c2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (c4 < C4) {
        float4 o;
        o.x = rs * (g[k].x - c1 - xh[k].x * c2);
        o.y = rs * (g[k].y - c1 - xh[k].y * c2);
        o.z = rs * (g[k].z - c1 - xh[k].z * c2);
        o.w = rs * (g[k].w - c1 - xh[k].w * c2);
        if (dres_in) {
          const float4 r = *reinterpret_cast<const float4*>(dres_in + row * C + c4 * 4);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + row * C + c4 * 4) = o;
        if (dx_lp) Vec4<TLP>::st(dx_lp + row * C + c4 * 4, o);
      }
    }
  }
  // CTA-level reduction of the parameter-gradient partials
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    if (c4 < C4) {
      *reinterpret_cast<float4*>(sred + ((size_t)warp * 2 + 0) * C + c4 * 4) = dg[k];
      *reinterpret_cast<float4*>(sred + ((size_t)warp * 2 + 1) * C + c4 * 4) = db[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) a += sred[(size_t)w * 2 * C + i];
    ws[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// ---- pipelined variant for C = 128 * NV * WPR (the model's 512 / 1024 and the toy widths) -------------------------
// The kernel above keeps ~16 rows per SM in flight but every warp alternates "load, reduce, load dres, store", so its
// loads are outstanding only about half of the time: 3.4 TB/s at the decoder shape (profiles/r1_step_launches.md).
// Here a row is owned by WPR warps (NV float4 per lane: registers stay bounded for wide rows), ALL of a row's inputs
// (dy, x, dres_in, mean, rstd) are requested one row ahead into raw registers, and the statistics pass / output pass
// recompute xhat and g from those raw values instead of keeping them.
template <typename T> struct Raw4;
template <> struct Raw4<float> {
  typedef float4 type;
  static __device__ __forceinline__ type ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ float4 f(type v) { return v; }
};
template <> struct Raw4<__nv_bfloat16> {
  typedef uint2 type;
  static __device__ __forceinline__ type ld(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  static __device__ __forceinline__ float4 f(type v) {
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

template <int NV, int WPR, typename TDY, typename TX>
__global__ void __launch_bounds__(kLnWarps * 32, 3) add_ln_bwd_pipe_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                                        const float* __restrict__ mean,
                                                                        const float* __restrict__ rstd,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ dres_in,
                                                                        float* __restrict__ dx_f32,
                                                                        __nv_bfloat16* __restrict__ dx_lp,
                                                                        float* __restrict__ ws, int64_t M) {
  constexpr int C = 128 * NV * WPR, kTeams = kLnWarps / WPR, kPart = 128 * NV;  // columns per warp
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sred[kLnWarps][2][kPart];
  __shared__ float2 sx[2][kLnWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int team = warp / WPR, part = warp % WPR;
  const int col0 = part * kPart + lane * 4;  // + 128 k
  typedef typename Raw4<TDY>::type RD;
  typedef typename Raw4<TX>::type RX;
  float4 gm[NV], dg[NV], db[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    gm[k] = *reinterpret_cast<const float4*>(gamma + col0 + 128 * k);
    dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t stride = (int64_t)gridDim.x * kTeams;
  int64_t row = (int64_t)blockIdx.x * kTeams + team;
  RD d_n[NV];
  RX x_n[NV];
  float4 r_n[NV];
  float mu_n = 0.f, rs_n = 0.f;
  auto fetch = [&](int64_t r) {
    const size_t base = (size_t)r * C + col0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      d_n[k] = Raw4<TDY>::ld(dy + base + 128 * k);
      x_n[k] = Raw4<TX>::ld(x + base + 128 * k);
      if (dres_in) r_n[k] = *reinterpret_cast<const float4*>(dres_in + base + 128 * k);
    }
    mu_n = mean[r];
    rs_n = rstd[r];
  };
  if (row < M) fetch(row);
  int it = 0;
  for (; row < M; row += stride, ++it) {
    RD d_c[NV];
    RX x_c[NV];
    float4 r_c[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) { d_c[k] = d_n[k]; x_c[k] = x_n[k]; r_c[k] = r_n[k]; }
    const float mu = mu_n, rs = rs_n;
    if (row + stride < M) fetch(row + stride);  // next row's requests are in flight while this one is reduced and stored
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 d = Raw4<TDY>::f(d_c[k]), xv = Raw4<TX>::f(x_c[k]);
      const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      const float4 g = make_float4(d.x * gm[k].x, d.y * gm[k].y, d.z * gm[k].z, d.w * gm[k].w);
      s1 += (g.x + g.y) + (g.z + g.w);
      s2 += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
      dg[k].x += d.x * xh.x; dg[k].y += d.y * xh.y; dg[k].z += d.z * xh.z; dg[k].w += d.w * xh.w;
      db[k].x += d.x; db[k].y += d.y; db[k].z += d.z; db[k].w += d.w;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (WPR > 1) {  // the row's other warps: two alternating slots, one named barrier per row and team
      if (lane == 0) sx[it & 1][warp] = make_float2(s1, s2);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(WPR * 32) : "memory");
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int w = 0; w < WPR; ++w) { const float2 v = sx[it & 1][team * WPR + w]; s1 += v.x; s2 += v.y; }
    }
    const float c1 = s1 * (1.f / (float)C), c2 = s2 * (1.f / (float)C);
    const size_t base = (size_t)row * C + col0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 d = Raw4<TDY>::f(d_c[k]), xv = Raw4<TX>::f(x_c[k]);
      float4 o;
      o.x = rs * (d.x * gm[k].x - c1 - (xv.x - mu) * rs * c2);
      o.y = rs * (d.y * gm[k].y - c1 - (xv.y - mu) * rs * c2);
      o.z = rs * (d.z * gm[k].z - c1 - (xv.z - mu) * rs * c2);
      o.w = rs * (d.w * gm[k].w - c1 - (xv.w - mu) * rs * c2);
      if (dres_in) { o.x += r_c[k].x; o.y += r_c[k].y; o.z += r_c[k].z; o.w += r_c[k].w; }
      if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + base + 128 * k) = o;
      if (dx_lp) Vec4<__nv_bfloat16>::st(dx_lp + base + 128 * k, o);
    }
  }
  // CTA-level reduction of the parameter-gradient partials: warps with the same `part` hold the same columns
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    *reinterpret_cast<float4*>(&sred[warp][0][lane * 4 + 128 * k]) = dg[k];
    *reinterpret_cast<float4*>(&sred[warp][1][lane * 4 + 128 * k]) = db[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i % C, pp = c / kPart, cc = c % kPart;
    float a = 0.f;
#pragma unroll
    for (int t = 0; t < kTeams; ++t) a += sred[t * WPR + pp][which][cc];
    ws[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// 32 columns x 32 row-slices per CTA; fixed summation order (slice-strided partials, then slices 0..31)
__global__ void __launch_bounds__(1024) ln_bwd_finish_kernel(const float* __restrict__ ws, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int nblocks, int C) {
  __shared__ float sm[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (i < 2 * C)
    for (int b = threadIdx.y; b < nblocks; b += 32) a += ws[(size_t)b * 2 * C + i];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && i < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    if (i < C) dgamma[i] = t; else dbeta[i - C] = t;
  }
}

static int64_t ln_bwd_blocks(int64_t M) {
  int64_t blocks = ceil_div64(M, (int64_t)kLnWarps);  // one row per warp when M is small: the kernel is latency-bound there
  const int64_t cap = (int64_t)oct_num_sms() * 8;
  return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}

extern "C" size_t oct_add_ln_bwd_ws_bytes(int64_t M, int64_t C) {
  return (size_t)ln_bwd_blocks(M) * 2 * C * sizeof(float);
}

template <int NV, typename TDY, typename TX, typename TLP>
static int launch_ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                         const float* dres_in, float* dx_f32, void* dx_lp, float* dgamma, float* dbeta, float* ws,
                         int64_t M, int C, cudaStream_t st, int* nblocks_out = nullptr) {
  const int64_t blocks = ln_bwd_blocks(M);
  const size_t smem = (size_t)kLnWarps * 2 * C * sizeof(float);
  auto kern = add_ln_bwd_kernel<NV, TDY, TX, TLP>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { oct_set_error("oct_add_ln_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  }
  kern<<<(unsigned)blocks, kLnWarps * 32, smem, st>>>((const TDY*)dy, (const TX*)x, mean, rstd, gamma, dres_in, dx_f32,
                                                     (TLP*)dx_lp, ws, M, C);
  int rc = oct_check_launch("oct_add_ln_bwd");
  if (rc) return rc;
  if (nblocks_out) { *nblocks_out = (int)blocks; return OCT_OK; }  // the caller runs the finish (another stream)
  ln_bwd_finish_kernel<<<(unsigned)ceil_div64(2 * C, 32), dim3(32, 32), 0, st>>>(ws, dgamma, dbeta, (int)blocks, C);
  return oct_check_launch("oct_add_ln_bwd(finish)");
}

template <int NV, int WPR>
static int launch_ln_bwd_pipe_t(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                                const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp, float* dgamma,
                                float* dbeta, float* ws, int64_t M, cudaStream_t st, int* nblocks_out) {
  constexpr int C = 128 * NV * WPR, kTeams = kLnWarps / WPR;
  int64_t blocks = ceil_div64(M, kTeams);
  const int64_t cap = (int64_t)oct_num_sms() * 3;  // ~150 registers per thread: three CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks > ln_bwd_blocks(M)) blocks = ln_bwd_blocks(M);  // the workspace is sized by oct_add_ln_bwd_ws_bytes
#define LNP(TDY, TX)                                                                                                   \
  oct_launch(add_ln_bwd_pipe_kernel<NV, WPR, TDY, TX>, dim3((unsigned)blocks), dim3(kLnWarps * 32), 0, st, 1,         \
             (const TDY*)dy, (const TX*)x, mean, rstd, gamma, dres_in, dx_f32, (__nv_bfloat16*)dx_lp, ws, M)
  if (dy_dtype == OCT_BF16 && x_dtype == OCT_F32) LNP(__nv_bfloat16, float);
  else if (dy_dtype == OCT_F32 && x_dtype == OCT_F32) LNP(float, float);
  else if (dy_dtype == OCT_BF16 && x_dtype == OCT_BF16) LNP(__nv_bfloat16, __nv_bfloat16);
  else LNP(float, __nv_bfloat16);
#undef LNP
  int rc = oct_check_launch("oct_add_ln_bwd(pipe)");
  if (rc) return rc;
  if (nblocks_out) { *nblocks_out = (int)blocks; return OCT_OK; }
  oct_launch(ln_bwd_finish_kernel, dim3((unsigned)ceil_div64(2 * C, 32)), dim3(32, 32), 0, st, 1, (const float*)ws, dgamma, dbeta,
             (int)blocks, C);
  return oct_check_launch("oct_add_ln_bwd(finish)");
}

static int launch_ln_bwd_pipe(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                              const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp, float* dgamma,
                              float* dbeta, float* ws, int64_t M, int C, cudaStream_t st, int* nblocks_out) {
#define A dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres_in, dx_f32, dx_lp, dgamma, dbeta, ws, M, st, nblocks_out
  switch (C) {
    case 128: return launch_ln_bwd_pipe_t<1, 1>(A);
    case 256: return launch_ln_bwd_pipe_t<2, 1>(A);
    case 512: return launch_ln_bwd_pipe_t<4, 1>(A);
    case 1024: return launch_ln_bwd_pipe_t<4, 2>(A);
    default: return OCT_ERR_UNSUPPORTED;
  }
#undef A
}

template <typename TDY, typename TX, typename TLP>
static int dispatch_ln_bwd_nv(int nv, const void* dy, const void* x, const float* mean, const float* rstd,
                              const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp, float* dgamma,
                              float* dbeta, float* ws, int64_t M, int C, cudaStream_t st, int* nblocks_out) {
#define A dy, x, mean, rstd, gamma, dres_in, dx_f32, dx_lp, dgamma, dbeta, ws, M, C, st, nblocks_out
  switch (nv) {
    case 1: return launch_ln_bwd<1, TDY, TX, TLP>(A);
    case 2: return launch_ln_bwd<2, TDY, TX, TLP>(A);
    case 4: return launch_ln_bwd<4, TDY, TX, TLP>(A);
    case 8: return launch_ln_bwd<8, TDY, TX, TLP>(A);
    default: return launch_ln_bwd<16, TDY, TX, TLP>(A);
  }
#undef A
}

static int add_ln_bwd_impl(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean,
                              const float* rstd, const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp,
                              int dx_lp_dtype, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int64_t M,
                              int64_t C, oct_stream_t stream, int* nblocks_out) {
  OCT_REQUIRE(dy && x && mean && rstd && gamma && (nblocks_out || (dgamma && dbeta)), "oct_add_ln_bwd: null pointer");
  OCT_REQUIRE(C > 0 && C % 4 == 0 && C <= 2048, "oct_add_ln_bwd: need C%%4==0 and C<=2048 (got %lld)", (long long)C);
  if (!ws || ws_bytes < oct_add_ln_bwd_ws_bytes(M, C)) {
    oct_set_error("oct_add_ln_bwd: workspace too small (%zu < %zu)", ws_bytes, oct_add_ln_bwd_ws_bytes(M, C));
    return OCT_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) {
    if (nblocks_out) { *nblocks_out = 0; return OCT_OK; }
    cudaMemsetAsync(dgamma, 0, C * sizeof(float), st);
    cudaMemsetAsync(dbeta, 0, C * sizeof(float), st);
    return OCT_OK;
  }
  const int nv = nv_for(C);
  // the low-precision copy is only ever bf16 (or absent); dy and x each f32|bf16
  OCT_REQUIRE(!dx_lp || dx_lp_dtype == OCT_BF16, "oct_add_ln_bwd: dx_lp must be bf16");
  if (getenv("OCT_LN_NO_PIPE") == nullptr) {
    const int rc = launch_ln_bwd_pipe(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres_in, dx_f32, dx_lp, dgamma, dbeta,
                                      (float*)ws, M, (int)C, st, nblocks_out);
    if (rc != OCT_ERR_UNSUPPORTED) return rc;  // widths the pipelined kernel does not cover fall through
  }
#define A nv, dy, x, mean, rstd, gamma, dres_in, dx_f32, dx_lp, dgamma, dbeta, (float*)ws, M, (int)C, st, nblocks_out
  if (dy_dtype == OCT_F32 && x_dtype == OCT_F32) return dispatch_ln_bwd_nv<float, float, __nv_bfloat16>(A);
  if (dy_dtype == OCT_BF16 && x_dtype == OCT_F32) return dispatch_ln_bwd_nv<__nv_bfloat16, float, __nv_bfloat16>(A);
  if (dy_dtype == OCT_F32 && x_dtype == OCT_BF16) return dispatch_ln_bwd_nv<float, __nv_bfloat16, __nv_bfloat16>(A);
  if (dy_dtype == OCT_BF16 && x_dtype == OCT_BF16)
    return dispatch_ln_bwd_nv<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(A);
#undef A
  OCT_REQUIRE(false, "oct_add_ln_bwd: bad dtype");
}

extern "C" int oct_add_ln_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean,
                              const float* rstd, const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp,
                              int dx_lp_dtype, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int64_t M,
                              int64_t C, oct_stream_t stream) {
  return add_ln_bwd_impl(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres_in, dx_f32, dx_lp, dx_lp_dtype, dgamma, dbeta, ws,
                         ws_bytes, M, C, stream, nullptr);
}

// The same backward split in two launches for callers that keep the parameter-gradient reduction OFF the dgrad chain: `main`
// writes dx (+ per-CTA partial sums of dgamma / dbeta into ws) and returns the number of partial rows; `finish` reduces them in a
// fixed order on whatever stream the caller likes (the weight-gradient side stream in ops.AddLNFn).
extern "C" int oct_add_ln_bwd_main(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean,
                                   const float* rstd, const float* gamma, const float* dres_in, float* dx_f32, void* dx_lp,
                                   int dx_lp_dtype, void* ws, size_t ws_bytes, int64_t M, int64_t C, int* nblocks,
                                   oct_stream_t stream) {
  OCT_REQUIRE(nblocks, "oct_add_ln_bwd_main: null nblocks");
  return add_ln_bwd_impl(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres_in, dx_f32, dx_lp, dx_lp_dtype, nullptr, nullptr, ws,
                         ws_bytes, M, C, stream, nblocks);
}

extern "C" int oct_add_ln_bwd_finish(const void* ws, int nblocks, int64_t C, float* dgamma, float* dbeta, oct_stream_t stream) {
  OCT_REQUIRE(dgamma && dbeta && (ws || nblocks == 0) && C > 0 && nblocks >= 0, "oct_add_ln_bwd_finish: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (nblocks == 0) {
    cudaMemsetAsync(dgamma, 0, C * sizeof(float), st);
    cudaMemsetAsync(dbeta, 0, C * sizeof(float), st);
    return OCT_OK;
  }
  ln_bwd_finish_kernel<<<(unsigned)ceil_div64(2 * C, 32), dim3(32, 32), 0, st>>>((const float*)ws, dgamma, dbeta, nblocks, (int)C);
  return oct_check_launch("oct_add_ln_bwd_finish");
}

// ------------------------------------------------------------------------------------------------
// GELU / cast / colsum
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = Vec4<T>::ld(x + i * 4);
    v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
    Vec4<T>::st(y + i * 4, v);
  }
}
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = Vec4<T>::ld(x + i * 4), d = Vec4<T>::ld(dy + i * 4);
    d.x *= gelu_erf_grad(v.x); d.y *= gelu_erf_grad(v.y); d.z *= gelu_erf_grad(v.z); d.w *= gelu_erf_grad(v.w);
    Vec4<T>::st(dx + i * 4, d);
  }
}

static unsigned ew_blocks(int64_t n4) {
  int64_t b = ceil_div64(n4, 256);
  const int64_t cap = (int64_t)oct_num_sms() * 8;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

extern "C" int oct_gelu_fwd(const void* x, void* y, int dtype, int64_t n, oct_stream_t stream) {
  OCT_REQUIRE(x && y && n % 4 == 0, "oct_gelu_fwd: null pointer or n%%4 != 0");
  if (n == 0) return OCT_OK;
  if (dtype == OCT_F32) gelu_fwd_kernel<float><<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, n / 4);
  else if (dtype == OCT_BF16) gelu_fwd_kernel<__nv_bfloat16><<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 4);
  else OCT_REQUIRE(false, "oct_gelu_fwd: bad dtype");
  return oct_check_launch("oct_gelu_fwd");
}
extern "C" int oct_gelu_bwd(const void* dy, const void* x, void* dx, int dtype, int64_t n, oct_stream_t stream) {
  OCT_REQUIRE(dy && x && dx && n % 4 == 0, "oct_gelu_bwd: null pointer or n%%4 != 0");
  if (n == 0) return OCT_OK;
  if (dtype == OCT_F32) gelu_bwd_kernel<float><<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>((const float*)dy, (const float*)x, (float*)dx, n / 4);
  else if (dtype == OCT_BF16) gelu_bwd_kernel<__nv_bfloat16><<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (__nv_bfloat16*)dx, n / 4);
  else OCT_REQUIRE(false, "oct_gelu_bwd: bad dtype");
  return oct_check_launch("oct_gelu_bwd");
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    Vec4<__nv_bfloat16>::st(dst + i * 4, *reinterpret_cast<const float4*>(src + i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16_rn(src[(n4 << 2) + threadIdx.x]);
}
extern "C" int oct_cast_f32_to_bf16(const float* src, void* dst, int64_t n, oct_stream_t stream) {
  OCT_REQUIRE(src && dst, "oct_cast_f32_to_bf16: null pointer");
  OCT_REQUIRE(aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "oct_cast_f32_to_bf16: misaligned");
  if (n == 0) return OCT_OK;
  oct_launch(cast_f32_bf16_kernel, dim3(ew_blocks(n / 4 + 1)), dim3(256), 0, (cudaStream_t)stream, 1, src, (__nv_bfloat16*)dst,
             (int64_t)n);
  return oct_check_launch("oct_cast_f32_to_bf16");
}

// All weight shadows of a model in ONE launch: table rows [src pointer, dst pointer, element count] per chunk (<= 16384 elements,
// never crossing a tensor; chunk starts 4-element aligned).  The per-parameter form above is ~200 launches per step for the MAE
// (0.25 ms of launch gaps on top of the 0.34 ms the 2 GB of traffic take).
__global__ void __launch_bounds__(256) cast_f32_bf16_multi_kernel(const int64_t* __restrict__ table, int n_chunks) {
  pdl_launch_dependents();
  pdl_wait();
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const float* src = reinterpret_cast<const float*>(table[3 * (int64_t)c]);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(table[3 * (int64_t)c + 1]);
    const int n = (int)table[3 * (int64_t)c + 2], n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) Vec4<__nv_bfloat16>::st(dst + i * 4, *reinterpret_cast<const float4*>(src + i * 4));
    if ((int)threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16_rn(src[(n4 << 2) + threadIdx.x]);
  }
}
extern "C" int oct_cast_f32_to_bf16_multi(const int64_t* table, int64_t n_chunks, oct_stream_t stream) {
  OCT_REQUIRE(table || n_chunks == 0, "oct_cast_f32_to_bf16_multi: null table");
  OCT_REQUIRE(n_chunks >= 0 && n_chunks < (1 << 30), "oct_cast_f32_to_bf16_multi: bad chunk count");
  if (n_chunks == 0) return OCT_OK;
  const int64_t cap = (int64_t)oct_num_sms() * 8;
  oct_launch(cast_f32_bf16_multi_kernel, dim3((unsigned)(n_chunks < cap ? n_chunks : cap)), dim3(256), 0, (cudaStream_t)stream, 1, table,
             (int)n_chunks);
  return oct_check_launch("oct_cast_f32_to_bf16_multi");
}

// colsum: block = 32 column-quads x 8 row lanes; grid.y row slices -> ws[slice][N] -> finish
constexpr int kCsRows = 8;
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, float* __restrict__ ws, int64_t M, int N, int64_t ldx,
                                      int64_t rows_per_slice) {
  __shared__ float4 sm[kCsRows][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  int64_t r1 = r0 + rows_per_slice;
  if (r1 > M) r1 = M;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < N) {
    // 4 independent row streams per thread keep 4 loads in flight (a single dependent stream ran at ~20 % of HBM rate)
    float4 b = a, d = a, e = a;
    int64_t r = r0 + threadIdx.y;
    for (; r + 3 * kCsRows < r1; r += 4 * kCsRows) {
      const float4 v0 = Vec4<T>::ld(x + r * ldx + c);
      const float4 v1 = Vec4<T>::ld(x + (r + kCsRows) * ldx + c);
      const float4 v2 = Vec4<T>::ld(x + (r + 2 * kCsRows) * ldx + c);
      const float4 v3 = Vec4<T>::ld(x + (r + 3 * kCsRows) * ldx + c);
      a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
      b.x += v1.x; b.y += v1.y; b.z += v1.z; b.w += v1.w;
      d.x += v2.x; d.y += v2.y; d.z += v2.z; d.w += v2.w;
      e.x += v3.x; e.y += v3.y; e.z += v3.z; e.w += v3.w;
    }
    for (; r < r1; r += kCsRows) {
      const float4 v = Vec4<T>::ld(x + r * ldx + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x += (b.x + d.x) + e.x; a.y += (b.y + d.y) + e.y; a.z += (b.z + d.z) + e.z; a.w += (b.w + d.w) + e.w;
  }
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
#pragma unroll
    for (int k = 1; k < kCsRows; ++k) {
      float4 v = sm[k][threadIdx.x];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(ws + (size_t)blockIdx.y * N + c) = a;
  }
}
__global__ void colsum_finish_kernel(const float* __restrict__ ws, float* __restrict__ out, int slices, int N, int beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float a = 0.f;
  for (int s = 0; s < slices; ++s) a += ws[(size_t)s * N + c];
  out[c] = beta ? out[c] + a : a;
}
// any N / ldx (classifier heads with a handful of classes): one thread per column, rows in order
template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t M, int N, int64_t ldx, int beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float a = 0.f;
  for (int64_t r = 0; r < M; ++r) a += ldf<T>(x + r * ldx + c);
  out[c] = beta ? out[c] + a : a;
}
static int colsum_slices(int64_t M) {
  int64_t s = ceil_div64(M, 128);   // >= 128 rows (16 per row lane) per slice
  if (s > 296) s = 296;             // 2 slices per SM at most
  return (int)(s < 1 ? 1 : s);
}
extern "C" size_t oct_colsum_ws_bytes(int64_t M, int64_t N) { return (size_t)colsum_slices(M) * N * sizeof(float); }
extern "C" int oct_colsum(const void* x, int x_dtype, float* out, int64_t M, int64_t N, int64_t ldx, int beta, void* ws,
                          size_t ws_bytes, oct_stream_t stream) {
  OCT_REQUIRE(x && out, "oct_colsum: null pointer");
  OCT_REQUIRE(x_dtype == OCT_F32 || x_dtype == OCT_BF16, "oct_colsum: bad dtype");
  if (!ws || ws_bytes < oct_colsum_ws_bytes(M, N)) { oct_set_error("oct_colsum: workspace too small"); return OCT_ERR_WORKSPACE; }
  if (N == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (N % 4 != 0 || ldx % 4 != 0 || (x_dtype == OCT_F32 ? !aligned16(x) : (reinterpret_cast<uintptr_t>(x) & 7) != 0)) {
    const unsigned g = (unsigned)ceil_div64(N, 128);
    if (x_dtype == OCT_F32) colsum_scalar_kernel<float><<<g, 128, 0, st>>>((const float*)x, out, M, (int)N, ldx, beta);
    else colsum_scalar_kernel<__nv_bfloat16><<<g, 128, 0, st>>>((const __nv_bfloat16*)x, out, M, (int)N, ldx, beta);
    return oct_check_launch("oct_colsum(scalar)");
  }
  const int slices = colsum_slices(M);
  const int64_t rps = ceil_div64(M > 0 ? M : 1, slices);
  dim3 grid((unsigned)ceil_div64(N, 128), (unsigned)slices), block(32, kCsRows);
  if (x_dtype == OCT_F32) colsum_partial_kernel<float><<<grid, block, 0, st>>>((const float*)x, (float*)ws, M, (int)N, ldx, rps);
  else if (x_dtype == OCT_BF16) colsum_partial_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, (float*)ws, M, (int)N, ldx, rps);
  else OCT_REQUIRE(false, "oct_colsum: bad dtype");
  int rc = oct_check_launch("oct_colsum");
  if (rc) return rc;
  colsum_finish_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>((const float*)ws, out, slices, (int)N, beta);
  return oct_check_launch("oct_colsum(finish)");
}
