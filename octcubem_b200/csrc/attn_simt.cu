// fp32-math CUDA-core self-attention forward/backward on packed qkv [B,S,3,H,d] — the 1e-4 parity mode of
// flash_attn_qkvpacked_func (flash_attn/modules/mha.py:122-130; maths as SelfAttention, mha.py:247-277):
// non-causal, no dropout, softmax(scale * q k^T) v.  Flash-style (never materialises S x S), exact expf.
// The throughput path is attn_tc.cu (tcgen05).
#include "common.cuh"

namespace {

constexpr int kQThreads = 128;  // queries per CTA in fwd / dQ kernels (one thread = one query row)
constexpr int kKT = 32;         // keys per shared-memory tile

template <int HD, typename T>
__global__ void __launch_bounds__(kQThreads) attn_fwd_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                                   float* __restrict__ lse, int S, int H, float scale) {
  __shared__ float Ks[kKT][HD];
  __shared__ float Vs[kKT][HD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * kQThreads + threadIdx.x;
  const bool active = i < S;
  const size_t tok_stride = (size_t)3 * H * HD;
  const T* base = qkv + (size_t)b * S * tok_stride + (size_t)h * HD;
  float q[HD], acc[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) {
    q[c] = active ? ldf(base + (size_t)i * tok_stride + c) * scale : 0.f;
    acc[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < S; k0 += kKT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKT * HD; e += kQThreads) {
      const int j = e / HD, c = e % HD;
      const bool ok = (k0 + j) < S;
      const T* kp = base + (size_t)(k0 + j) * tok_stride + (size_t)H * HD + c;
      Ks[j][c] = ok ? ldf(kp) : 0.f;
      Vs[j][c] = ok ? ldf(kp + (size_t)H * HD) : 0.f;
    }
    __syncthreads();
    const int nk = min(kKT, S - k0);
    float s[kKT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < kKT; ++j) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) d = fmaf(q[c], Ks[j][c], d);
      s[j] = (j < nk) ? d : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);  // exp(-inf) = 0 on the first tile
    l *= corr;
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] *= corr;
#pragma unroll
    for (int j = 0; j < kKT; ++j) {
      const float p = expf(s[j] - m_new);  // masked keys: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] = fmaf(p, Vs[j][c], acc[c]);
    }
    m = m_new;
  }
  if (active) {
    const float inv = 1.f / l;
    T* o = out + ((size_t)b * S + i) * H * HD + (size_t)h * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) stf(o + c, acc[c] * inv);
    lse[((size_t)b * H + h) * S + i] = m + logf(l);
  }
}

// dQ (thread per query) + delta_i = sum_c dO_i O_i (stored for the dK/dV kernel)
template <int HD, typename T>
__global__ void __launch_bounds__(kQThreads) attn_bwd_dq_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ out,
                                                                      const T* __restrict__ dout,
                                                                      const float* __restrict__ lse, T* __restrict__ dqkv,
                                                                      float* __restrict__ delta, int S, int H,
                                                                      float scale) {
  __shared__ float Ks[kKT][HD];
  __shared__ float Vs[kKT][HD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * kQThreads + threadIdx.x;
  const bool active = i < S;
  const size_t tok_stride = (size_t)3 * H * HD;
  const T* base = qkv + (size_t)b * S * tok_stride + (size_t)h * HD;
  float q[HD], dq[HD], dO[HD];
  float dlt = 0.f;
#pragma unroll
  for (int c = 0; c < HD; ++c) {
    q[c] = active ? ldf(base + (size_t)i * tok_stride + c) * scale : 0.f;
    const size_t oi = ((size_t)b * S + i) * H * HD + (size_t)h * HD + c;
    dO[c] = active ? ldf(dout + oi) : 0.f;
    dlt += active ? dO[c] * ldf(out + oi) : 0.f;
    dq[c] = 0.f;
  }
  const float my_lse = active ? lse[((size_t)b * H + h) * S + i] : 0.f;
  if (active) delta[((size_t)b * H + h) * S + i] = dlt;
  for (int k0 = 0; k0 < S; k0 += kKT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKT * HD; e += kQThreads) {
      const int j = e / HD, c = e % HD;
      const bool ok = (k0 + j) < S;
      const T* kp = base + (size_t)(k0 + j) * tok_stride + (size_t)H * HD + c;
      Ks[j][c] = ok ? ldf(kp) : 0.f;
      Vs[j][c] = ok ? ldf(kp + (size_t)H * HD) : 0.f;
    }
    __syncthreads();
    const int nk = min(kKT, S - k0);
    for (int j = 0; j < nk; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        s = fmaf(q[c], Ks[j][c], s);
        dp = fmaf(dO[c], Vs[j][c], dp);
      }
      const float p = expf(s - my_lse);
      const float ds = p * (dp - dlt);
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = fmaf(ds, Ks[j][c], dq[c]);
    }
  }
  if (active) {
    T* o = dqkv + ((size_t)b * S + i) * tok_stride + (size_t)h * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) stf(o + c, dq[c] * scale);
  }
}

// dK, dV (thread per key); own K/V rows and the Q/dO tile live in shared memory, dk/dv accumulators in registers
constexpr int kKThreads = 64;
constexpr int kQT = 16;

template <int HD, typename T>
__global__ void __launch_bounds__(kKThreads) attn_bwd_dkv_simt_kernel(const T* __restrict__ qkv,
                                                                       const T* __restrict__ dout,
                                                                       const float* __restrict__ lse,
                                                                       const float* __restrict__ delta,
                                                                       T* __restrict__ dqkv, int S, int H, float scale) {
  extern __shared__ float smem[];
  float* Ks = smem;                          // [kKThreads][HD+1]
  float* Vs = Ks + kKThreads * (HD + 1);     // [kKThreads][HD+1]
  float* Qs = Vs + kKThreads * (HD + 1);     // [kQT][HD]
  float* Os = Qs + kQT * HD;                 // [kQT][HD]
  float* Ls = Os + kQT * HD;                 // [kQT] lse
  float* Ds = Ls + kQT;                      // [kQT] delta
  const int b = blockIdx.z, h = blockIdx.y;
  const int j0 = blockIdx.x * kKThreads;
  const int j = j0 + threadIdx.x;
  const bool active = j < S;
  const size_t tok_stride = (size_t)3 * H * HD;
  const T* base = qkv + (size_t)b * S * tok_stride + (size_t)h * HD;
  for (int e = threadIdx.x; e < kKThreads * HD; e += kKThreads) {
    const int r = e / HD, c = e % HD;
    const bool ok = (j0 + r) < S;
    const T* kp = base + (size_t)(j0 + r) * tok_stride + (size_t)H * HD + c;
    Ks[r * (HD + 1) + c] = ok ? ldf(kp) : 0.f;
    Vs[r * (HD + 1) + c] = ok ? ldf(kp + (size_t)H * HD) : 0.f;
  }
  float dk[HD], dv[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) dk[c] = dv[c] = 0.f;
  const float* myK = Ks + threadIdx.x * (HD + 1);
  const float* myV = Vs + threadIdx.x * (HD + 1);
  for (int i0 = 0; i0 < S; i0 += kQT) {
    __syncthreads();
    for (int e = threadIdx.x; e < kQT * HD; e += kKThreads) {
      const int r = e / HD, c = e % HD;
      const bool ok = (i0 + r) < S;
      Qs[e] = ok ? ldf(base + (size_t)(i0 + r) * tok_stride + c) : 0.f;
      Os[e] = ok ? ldf(dout + ((size_t)b * S + i0 + r) * H * HD + (size_t)h * HD + c) : 0.f;
    }
    if (threadIdx.x < kQT) {
      const bool ok = (i0 + threadIdx.x) < S;
      Ls[threadIdx.x] = ok ? lse[((size_t)b * H + h) * S + i0 + threadIdx.x] : 0.f;
      Ds[threadIdx.x] = ok ? delta[((size_t)b * H + h) * S + i0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    const int nq = min(kQT, S - i0);
    for (int r = 0; r < nq; ++r) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        s = fmaf(Qs[r * HD + c], myK[c], s);
        dp = fmaf(Os[r * HD + c], myV[c], dp);
      }
      const float p = expf(s * scale - Ls[r]);
      const float ds = p * (dp - Ds[r]) * scale;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        dv[c] = fmaf(p, Os[r * HD + c], dv[c]);
        dk[c] = fmaf(ds, Qs[r * HD + c], dk[c]);
      }
    }
  }
  if (active) {
    T* o = dqkv + ((size_t)b * S + j) * tok_stride + (size_t)H * HD + (size_t)h * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      stf(o + c, dk[c]);
      stf(o + (size_t)H * HD + c, dv[c]);
    }
  }
}

template <int HD, typename T>
int launch_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, float scale, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(S, kQThreads), (unsigned)H, (unsigned)B);
  attn_fwd_simt_kernel<HD, T><<<grid, kQThreads, 0, st>>>((const T*)qkv, (T*)out, lse, (int)S, (int)H, scale);
  return oct_check_launch("oct_attn_fwd(f32)");
}

template <int HD, typename T>
int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int64_t B,
               int64_t S, int64_t H, float scale, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(S, kQThreads), (unsigned)H, (unsigned)B);
  attn_bwd_dq_simt_kernel<HD, T><<<grid, kQThreads, 0, st>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, (T*)dqkv,
                                                            delta, (int)S, (int)H, scale);
  int rc = oct_check_launch("oct_attn_bwd(f32,dq)");
  if (rc) return rc;
  const size_t smem = (size_t)(2 * kKThreads * (HD + 1) + 2 * kQT * HD + 2 * kQT) * sizeof(float);
  auto kern = attn_bwd_dkv_simt_kernel<HD, T>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { oct_set_error("oct_attn_bwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  }
  dim3 grid2((unsigned)ceil_div64(S, kKThreads), (unsigned)H, (unsigned)B);
  kern<<<grid2, kKThreads, smem, st>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, (int)S, (int)H, scale);
  return oct_check_launch("oct_attn_bwd(f32,dkv)");
}

}  // namespace

int oct_attn_fwd_simt(int io_dtype, const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, int64_t d,
                      float scale, cudaStream_t st) {
#define CASE(HD)                                                                                  \
  case HD:                                                                                        \
    return io_dtype == OCT_F32 ? launch_fwd<HD, float>(qkv, out, lse, B, S, H, scale, st)         \
                               : launch_fwd<HD, __nv_bfloat16>(qkv, out, lse, B, S, H, scale, st);
  switch (d) {
    CASE(16) CASE(32) CASE(64) CASE(128)
    default: oct_set_error("oct_attn_fwd(f32): head dim %lld unsupported (16/32/64/128)", (long long)d); return OCT_ERR_UNSUPPORTED;
  }
#undef CASE
}

int oct_attn_bwd_simt(int io_dtype, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      float* delta, int64_t B, int64_t S, int64_t H, int64_t d, float scale, cudaStream_t st) {
#define CASE(HD)                                                                                                      \
  case HD:                                                                                                            \
    return io_dtype == OCT_F32 ? launch_bwd<HD, float>(qkv, out, dout, lse, dqkv, delta, B, S, H, scale, st)          \
                               : launch_bwd<HD, __nv_bfloat16>(qkv, out, dout, lse, dqkv, delta, B, S, H, scale, st);
  switch (d) {
    CASE(16) CASE(32) CASE(64) CASE(128)
    default: oct_set_error("oct_attn_bwd(f32): head dim %lld unsupported (16/32/64/128)", (long long)d); return OCT_ERR_UNSUPPORTED;
  }
#undef CASE
}
