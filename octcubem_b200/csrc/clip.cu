// Contrastive step of OCTCube-IR: F.normalize + ClipLoss in the recipe's configuration
// (--local-loss --gather-with-grad; retinal-COEM/src/open_clip/model.py:661-683, loss.py:21-63,148-229).
//
// The reference all-gathers both feature matrices with an autograd-aware collective (loss.py:51-52), multiplies
// `logit_scale * local @ all^T` (loss.py:188-189), takes two cross-entropies (:223-227) and lets autograd run the
// reduce-scatter that the gather implies.  Here the exchange IS the kernel:
//   * every rank owns one exchange buffer in peer-mapped memory (NVLink 5 / NVSwitch); `clip_publish_kernel` copies the
//     rank's two feature matrices into it and raises a per-rank epoch flag on every peer with a release store;
//   * `clip_loss_fwd_kernel` computes its logits tile by tile straight from the PEERS' buffers (P2P loads) as their flags
//     arrive — peer (rank + 1) first, so the ranks do not hammer one GPU — with an online softmax, so neither the gathered
//     feature matrix nor the [B, B*W] logits ever exist in memory;
//   * `clip_loss_bwd_kernel` needs no reduce-scatter at all: with G_s = X_r Y_s^T (the same Gram tile the forward used),
//         dX_r = scale/(2B) * sum_s [ exp(scale G_s - lse_r[i]) + exp(scale G_s - lse'_s[j]) - 2 [s==r][i==j] ] Y_s
//     where lse'_s is the OTHER direction's log-sum-exp of rank s, a B-vector read from the peer's buffer
//     (closed form of oracle/clip_loss_oracle.py: the term a rank would receive from the reduce-scatter is recomputed locally
//     from the transposed tile, 4 MFLOP at B = 32, W = 8).
// Buffers are double-buffered by epoch parity; the epoch lives on the device so the three launches replay from a CUDA graph.
// The work is ~4 MFLOP per rank: CUDA-core fp32 FMA, latency-bound by design — tensor cores have nothing to do here.
#include "common.cuh"
#include <cstring>

namespace {

constexpr int kMaxWorld = 64;
constexpr int kRows = 8, kCols = 32, kThreads = 256;

// exchange buffer layout (32-bit words): [0, 64) feat_ready[s], [64, 128) lse_ready[s], then feat[2][2][B][D], lse[2][2][B]
__host__ __device__ inline size_t xchg_feat_off(int parity, int modality, int64_t B, int64_t D) {
  return 128 + ((size_t)parity * 2 + modality) * (size_t)B * D;
}
__host__ __device__ inline size_t xchg_lse_off(int parity, int dir, int64_t B, int64_t D) {
  return 128 + 4 * (size_t)B * D + ((size_t)parity * 2 + dir) * (size_t)B;
}
inline size_t xchg_words(int64_t B, int64_t D) { return 128 + 4 * (size_t)B * D + 4 * (size_t)B; }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded spin (~10 s): a peer that never publishes (a rank that died, ranks calling in different orders) must not wedge the
// GPU; the kernel then raises state[3] and carries on with whatever is in the buffer — the host side checks the flag.
__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned epoch, unsigned* timeout_flag) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
      __nanosleep(256);
      if (++spins > (1u << 25)) { atomicExch(timeout_flag, 1u); break; }
    }
  }
  __syncthreads();
}

struct PeerTable { float* buf[kMaxWorld]; };

// ---------------------------------------------------------------------------------------------------------------
// F.normalize(x, dim=-1): y = x / max(||x||, eps)      (model.py:663,667)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void l2norm_fwd_kernel(const T* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm, int B, int D, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const T* xr = x + (size_t)row * D;
  float ss = 0.f;
  for (int k = lane; k < D; k += 32) { const float v = ldf(xr + k); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int k = lane; k < D; k += 32) y[(size_t)row * D + k] = ldf(xr + k) * inv;
  if (lane == 0) inv_norm[row] = (sqrtf(ss) > eps) ? inv : -inv;  // sign bit marks the clamped (constant-scale) branch
}

template <typename T>
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ inv_norm,
                                  T* __restrict__ dx, int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* dyr = dy + (size_t)row * D;
  const float* yr = y + (size_t)row * D;
  const float inv = inv_norm[row];
  float dot = 0.f;
  if (inv > 0.f) {
    for (int k = lane; k < D; k += 32) dot = fmaf(dyr[k], yr[k], dot);
    dot = warp_sum(dot);
  }
  const float a = fabsf(inv);
  for (int k = lane; k < D; k += 32) stf(dx + (size_t)row * D + k, a * (dyr[k] - yr[k] * dot));
}

// ---------------------------------------------------------------------------------------------------------------
// exchange
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) clip_publish_kernel(const float* __restrict__ image, const float* __restrict__ enface,
                                                            PeerTable peers, const unsigned* __restrict__ epoch, int rank, int W,
                                                            int B, int D) {
  const unsigned e = *epoch + 1;
  float* mine = peers.buf[rank];
  const int n4 = (B * D) >> 2;
  float4* d0 = reinterpret_cast<float4*>(mine + xchg_feat_off(e & 1, 0, B, D));
  float4* d1 = reinterpret_cast<float4*>(mine + xchg_feat_off(e & 1, 1, B, D));
  const float4* s0 = reinterpret_cast<const float4*>(image);
  const float4* s1 = reinterpret_cast<const float4*>(enface);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) { d0[i] = s0[i]; d1[i] = s1[i]; }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < W) st_release_sys(reinterpret_cast<unsigned*>(peers.buf[threadIdx.x]) + rank, e);
}

// shared-memory tiles: As [kRows][D+1], Ys [kCols][D+1], Cs [kRows][kCols]
__device__ __forceinline__ void load_rows(float* dst, const float* src, int rows, int row0, int B, int D, bool remote) {
  const int ld = D + 1, d4 = D >> 2;
  for (int i = threadIdx.x; i < rows * d4; i += kThreads) {
    const int r = i / d4, c4 = i - r * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < B) {
      const float4* p = reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * D) + c4;
      v = remote ? __ldcg(p) : *p;   // peer memory: never through this SM's L1 (the same addresses are rewritten every 2nd step)
    }
    float* d = dst + r * ld + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

__device__ __forceinline__ float tile_dot(const float* As, const float* Ys, int i, int j, int D) {
  const float* a = As + i * (D + 1);
  const float* y = Ys + j * (D + 1);
  float acc0 = 0.f, acc1 = 0.f;
  for (int k = 0; k < D; k += 2) { acc0 = fmaf(a[k], y[k], acc0); acc1 = fmaf(a[k + 1], y[k + 1], acc1); }
  return acc0 + acc1;
}

// grid (ceil(B / kRows), 2 directions).  dir 0: rows = image, columns = all enface (logits_per_image);  dir 1: the converse.
__global__ void __launch_bounds__(kThreads) clip_loss_fwd_kernel(const float* __restrict__ image, const float* __restrict__ enface,
                                                                 const float* __restrict__ logit_scale, PeerTable peers,
                                                                 unsigned* __restrict__ epoch, unsigned* __restrict__ counter,
                                                                 float* __restrict__ row_loss, float* __restrict__ lse_out,
                                                                 float* __restrict__ loss, int rank, int W, int B, int D) {
  extern __shared__ float smem[];
  float* As = smem;
  float* Ys = smem + kRows * (D + 1);
  const int dir = blockIdx.y, row0 = blockIdx.x * kRows;
  const int i = threadIdx.x >> 5, j = threadIdx.x & 31;  // warp i owns local row row0 + i
  const unsigned e = *epoch + 1;
  const float scale = *logit_scale;
  const float* rows_src = dir == 0 ? image : enface;
  const float* cols_local = dir == 0 ? enface : image;
  load_rows(As, rows_src, kRows, row0, B, D, false);
  float m = -INFINITY, l = 0.f, diag = 0.f;
  for (int p = 0; p < W; ++p) {
    const int s = (rank + p) % W;
    const float* src = cols_local;
    if (s != rank) {
      wait_flag(reinterpret_cast<const unsigned*>(peers.buf[rank]) + s, e, epoch + 3);
      src = peers.buf[s] + xchg_feat_off(e & 1, dir == 0 ? 1 : 0, B, D);
    }
    for (int c0 = 0; c0 < B; c0 += kCols) {
      __syncthreads();
      load_rows(Ys, src, kCols, c0, B, D, s != rank);
      __syncthreads();
      const bool valid = (c0 + j < B) && (row0 + i < B);
      const float z = valid ? scale * tile_dot(As, Ys, i, j, D) : -INFINITY;
      if (valid && s == rank && c0 + j == row0 + i) diag = z;
      const float tm = warp_max(z);
      if (tm > m) { l *= __expf(m - tm); m = tm; }          // (m = -inf, l = 0 initially: exp(-inf) = 0)
      l += warp_sum(valid ? __expf(z - m) : 0.f);
    }
  }
  diag = warp_sum(diag);                                     // exactly one lane holds it
  if (j == 0 && row0 + i < B) {
    const float lse = m + __logf(l);
    lse_out[dir * B + row0 + i] = lse;
    peers.buf[rank][xchg_lse_off(e & 1, dir, B, D) + row0 + i] = lse;
    row_loss[dir * B + row0 + i] = lse - diag;               // cross-entropy of this row (labels = arange(B) + B * rank)
  }
  // last CTA: the loss in a fixed order, publish the lse vectors, advance the epoch
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float acc = 0.f;
    for (int k = threadIdx.x; k < 2 * B; k += 32) acc += __ldcg(row_loss + k);
    acc = warp_sum(acc);
    if (threadIdx.x == 0) {
      loss[0] = acc / (2.f * (float)B);
      *counter = 0;
      *epoch = e;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < W) st_release_sys(reinterpret_cast<unsigned*>(peers.buf[threadIdx.x]) + 64 + rank, e);
}

// grid (ceil(B / kRows), 2).  dir 0 -> d image, dir 1 -> d enface.  D <= 256 * kMaxK.
constexpr int kMaxK = 4;
__global__ void __launch_bounds__(kThreads) clip_loss_bwd_kernel(const float* __restrict__ image, const float* __restrict__ enface,
                                                                 const float* __restrict__ logit_scale, const float* __restrict__ dloss,
                                                                 PeerTable peers, const unsigned* __restrict__ epoch,
                                                                 unsigned* __restrict__ counter, const float* __restrict__ lse_local,
                                                                 float* __restrict__ d_image, float* __restrict__ d_enface,
                                                                 float* __restrict__ dscale_part, float* __restrict__ d_scale, int rank,
                                                                 int W, int B, int D) {
  extern __shared__ float smem[];
  float* As = smem;
  float* Ys = smem + kRows * (D + 1);
  float* Cs = Ys + kCols * (D + 1);
  __shared__ float red[kThreads / 32];
  const int dir = blockIdx.y, row0 = blockIdx.x * kRows;
  const int i = threadIdx.x >> 5, j = threadIdx.x & 31;
  const unsigned e = *epoch;
  const float scale = *logit_scale;
  const float* rows_src = dir == 0 ? image : enface;
  const float* cols_local = dir == 0 ? enface : image;
  load_rows(As, rows_src, kRows, row0, B, D, false);
  const float lse_i = (row0 + i < B) ? lse_local[dir * B + row0 + i] : 0.f;
  float acc[kRows][kMaxK];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int kk = 0; kk < kMaxK; ++kk) acc[r][kk] = 0.f;
  float ds = 0.f;
  for (int p = 0; p < W; ++p) {
    const int s = (rank + p) % W;
    const float* src = cols_local;
    const float* lse_peer = lse_local + (1 - dir) * B;      // the other direction's lse of the rank that owns the columns
    if (s != rank) {
      wait_flag(reinterpret_cast<const unsigned*>(peers.buf[rank]) + 64 + s, e, const_cast<unsigned*>(epoch) + 3);
      src = peers.buf[s] + xchg_feat_off(e & 1, dir == 0 ? 1 : 0, B, D);
      lse_peer = peers.buf[s] + xchg_lse_off(e & 1, 1 - dir, B, D);
    }
    for (int c0 = 0; c0 < B; c0 += kCols) {
      __syncthreads();
      load_rows(Ys, src, kCols, c0, B, D, s != rank);
      __syncthreads();
      const bool valid = (c0 + j < B) && (row0 + i < B);
      float c = 0.f;
      if (valid) {
        const float g = tile_dot(As, Ys, i, j, D), z = scale * g;
        const float lp = (s != rank) ? __ldcg(lse_peer + c0 + j) : lse_peer[c0 + j];
        const float hit = (s == rank && c0 + j == row0 + i) ? 1.f : 0.f;
        const float p_own = __expf(z - lse_i) - hit;           // (P_r - Y_r)[i, s*B + c0 + j]
        c = p_own + __expf(z - lp) - hit;                      // + (Q_s - Y_s)[c0 + j, r*B + i]  (the reduce-scatter term)
        ds = fmaf(p_own, g, ds);
      }
      Cs[i * kCols + j] = c;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kMaxK; ++kk) {
        const int k = threadIdx.x + kk * kThreads;
        if (k < D) {
#pragma unroll 4
          for (int jj = 0; jj < kCols; ++jj) {
            const float y = Ys[jj * (D + 1) + k];
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r][kk] = fmaf(Cs[r * kCols + jj], y, acc[r][kk]);
          }
        }
      }
    }
  }
  const float coef = dloss[0] * scale / (2.f * (float)B);
  float* dst = dir == 0 ? d_image : d_enface;
#pragma unroll
  for (int kk = 0; kk < kMaxK; ++kk) {
    const int k = threadIdx.x + kk * kThreads;
    if (k < D) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (row0 + r < B) dst[(size_t)(row0 + r) * D + k] = coef * acc[r][kk];
    }
  }
  // d logit_scale = dloss / (2B) * sum over both directions of <P - Y, G>: per-CTA partials, summed in a fixed order
  ds = warp_sum(ds);
  if (j == 0) red[i] = ds;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    dscale_part[blockIdx.y * gridDim.x + blockIdx.x] = t;
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1) {
      __threadfence();
      float tot = 0.f;
      for (unsigned k = 0; k < gridDim.x * gridDim.y; ++k) tot += __ldcg(dscale_part + k);
      d_scale[0] = dloss[0] * tot / (2.f * (float)B);
      *counter = 0;
    }
  }
}

int check_common(const void* const* peer_bufs, int rank, int W, int64_t B, int64_t D, const char* who) {
  OCT_REQUIRE(W >= 1 && W <= kMaxWorld && rank >= 0 && rank < W, "%s: bad rank / world size (world <= %d)", who, kMaxWorld);
  OCT_REQUIRE(B >= 1 && B <= 65535 && D >= 4 && D % 4 == 0 && D <= 256 * kMaxK, "%s: need 1 <= B, D %% 4 == 0, D <= %d", who, 256 * kMaxK);
  OCT_REQUIRE(peer_bufs, "%s: null peer table", who);
  for (int s = 0; s < W; ++s) OCT_REQUIRE(peer_bufs[s] && aligned16(peer_bufs[s]), "%s: peer buffer %d null or unaligned", who, s);
  return OCT_OK;
}

PeerTable make_table(const void* const* peer_bufs, int W) {
  PeerTable t;
  for (int s = 0; s < kMaxWorld; ++s) t.buf[s] = s < W ? (float*)peer_bufs[s] : nullptr;
  return t;
}

}  // namespace

extern "C" size_t oct_clip_xchg_bytes(int64_t B, int64_t D) { return xchg_words(B, D) * 4; }
extern "C" size_t oct_clip_state_bytes(int64_t B) { return (size_t)(8 + 4 * B + 2 * ceil_div64(B, kRows)) * 4; }

extern "C" int oct_l2norm_fwd(const void* x, int x_dtype, float* y, float* inv_norm, int64_t B, int64_t D, float eps, oct_stream_t stream) {
  OCT_REQUIRE(x && y && inv_norm && B >= 0 && D >= 1, "oct_l2norm_fwd: bad arguments");
  if (B == 0) return OCT_OK;
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div64(B, wpb));
  if (x_dtype == OCT_F32) l2norm_fwd_kernel<float><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>((const float*)x, y, inv_norm, (int)B, (int)D, eps);
  else if (x_dtype == OCT_BF16) l2norm_fwd_kernel<__nv_bfloat16><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, inv_norm, (int)B, (int)D, eps);
  else OCT_REQUIRE(false, "oct_l2norm_fwd: bad dtype");
  return oct_check_launch("oct_l2norm_fwd");
}

extern "C" int oct_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, void* dx, int dx_dtype, int64_t B, int64_t D,
                              oct_stream_t stream) {
  OCT_REQUIRE(dy && y && inv_norm && dx && B >= 0 && D >= 1, "oct_l2norm_bwd: bad arguments");
  if (B == 0) return OCT_OK;
  const int wpb = 8;
  dim3 grid((unsigned)ceil_div64(B, wpb));
  if (dx_dtype == OCT_F32) l2norm_bwd_kernel<float><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, (float*)dx, (int)B, (int)D);
  else if (dx_dtype == OCT_BF16) l2norm_bwd_kernel<__nv_bfloat16><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(dy, y, inv_norm, (__nv_bfloat16*)dx, (int)B, (int)D);
  else OCT_REQUIRE(false, "oct_l2norm_bwd: bad dtype");
  return oct_check_launch("oct_l2norm_bwd");
}

// state (device, zero-initialised by the caller, oct_clip_state_bytes): [0] epoch, [1] fwd counter, [2] bwd counter, [3] set to 1
// when a wait for a peer timed out, [8 ...) lse
// [2][B], row_loss [2][B], dscale partials [2 * ceil(B / 8)].
extern "C" int oct_clip_loss_fwd(const float* image, const float* enface, const float* logit_scale, const void* const* peer_bufs,
                                 void* state, float* loss, int rank, int world, int64_t B, int64_t D, oct_stream_t stream) {
  OCT_REQUIRE(image && enface && logit_scale && state && loss, "oct_clip_loss_fwd: null pointer");
  int rc = check_common(peer_bufs, rank, world, B, D, "oct_clip_loss_fwd");
  if (rc) return rc;
  OCT_REQUIRE(aligned16(image) && aligned16(enface), "oct_clip_loss_fwd: features must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  PeerTable t = make_table(peer_bufs, world);
  unsigned* u = (unsigned*)state;
  float* f = (float*)state;
  clip_publish_kernel<<<1, 1024, 0, st>>>(image, enface, t, u, rank, world, (int)B, (int)D);
  rc = oct_check_launch("oct_clip_loss_fwd(publish)");
  if (rc) return rc;
  const size_t smem = (size_t)(kRows + kCols) * (D + 1) * 4;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(clip_loss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(clip_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div64(B, kRows), 2);
  clip_loss_fwd_kernel<<<grid, kThreads, smem, st>>>(image, enface, logit_scale, t, u, u + 1, f + 8 + 2 * B, f + 8, loss, rank, world,
                                                     (int)B, (int)D);
  return oct_check_launch("oct_clip_loss_fwd");
}

extern "C" int oct_clip_loss_bwd(const float* image, const float* enface, const float* logit_scale, const float* dloss,
                                 const void* const* peer_bufs, void* state, float* d_image, float* d_enface, float* d_scale, int rank,
                                 int world, int64_t B, int64_t D, oct_stream_t stream) {
  OCT_REQUIRE(image && enface && logit_scale && dloss && state && d_image && d_enface && d_scale, "oct_clip_loss_bwd: null pointer");
  int rc = check_common(peer_bufs, rank, world, B, D, "oct_clip_loss_bwd");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PeerTable t = make_table(peer_bufs, world);
  unsigned* u = (unsigned*)state;
  float* f = (float*)state;
  const size_t smem = ((size_t)(kRows + kCols) * (D + 1) + kRows * kCols) * 4;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(clip_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div64(B, kRows), 2);
  clip_loss_bwd_kernel<<<grid, kThreads, smem, st>>>(image, enface, logit_scale, dloss, t, u, u + 2, f + 8, d_image, d_enface,
                                                     f + 8 + 4 * B, d_scale, rank, world, (int)B, (int)D);
  return oct_check_launch("oct_clip_loss_bwd");
}

// ---------------------------------------------------------------------------------------------------------------
// peer-mapped memory for the exchange buffers (one process per GPU: CUDA IPC over NVLink).  The only entry points of the
// library that allocate; they are called once at set-up, never inside a step.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int oct_peer_alloc(void** ptr, int64_t bytes) {
  OCT_REQUIRE(ptr && bytes > 0, "oct_peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) { oct_set_error("oct_peer_alloc: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { oct_set_error("oct_peer_alloc(memset): %s", cudaGetErrorString(e)); return (int)e; }
  return OCT_OK;
}
extern "C" int oct_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { oct_set_error("oct_peer_free: %s", cudaGetErrorString(e)); return (int)e; }
  return OCT_OK;
}
extern "C" int oct_peer_export(void* ptr, void* handle64) {
  OCT_REQUIRE(ptr && handle64, "oct_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, ptr);
  if (e != cudaSuccess) { oct_set_error("oct_peer_export: %s", cudaGetErrorString(e)); return (int)e; }
  return OCT_OK;
}
extern "C" int oct_peer_open(const void* handle64, void** ptr) {
  OCT_REQUIRE(ptr && handle64, "oct_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { oct_set_error("oct_peer_open: %s", cudaGetErrorString(e)); return (int)e; }
  return OCT_OK;
}
extern "C" int oct_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { oct_set_error("oct_peer_close: %s", cudaGetErrorString(e)); return (int)e; }
  return OCT_OK;
}
