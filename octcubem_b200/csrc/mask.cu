// random_masking and the token-routing kernels around it (HBM-bound integer / copy work).
//   oct_mask_sort          models_mae_joint_res_flash_attn.py:349-369
//   oct_patchify           :289-314 (+ kept-token variant used by the patch-embed wgrad)
//   oct_gather_tokens_*    :406-478
//   oct_unshuffle_*        :515-573
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// mask_sort: rank-by-counting.  rank(i) = #{j : (key_j, j) < (key_i, i)} is exactly the position of i
// in a STABLE ascending sort, i.e. ids_restore.  No data movement, no inter-thread dependency, the whole
// chip works on B rows at once (a per-row block sort would occupy only B SMs).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t orderable(float f) {
  if (f == 0.f) f = 0.f;  // -0.0 == +0.0 for comparison sorts
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kSortThreads = 128;

__global__ void __launch_bounds__(kSortThreads) mask_sort_kernel(const float* __restrict__ noise, int L, int keep,
                                                                  int64_t* __restrict__ ids_restore,
                                                                  int64_t* __restrict__ ids_keep,
                                                                  float* __restrict__ mask) {
  extern __shared__ uint32_t skeys[];  // L (+ padding to a multiple of 4)
  const int b = blockIdx.y;
  const float* row = noise + (size_t)b * L;
  const int L4 = (L + 3) & ~3;
  for (int j = threadIdx.x; j < L4; j += kSortThreads) skeys[j] = (j < L) ? orderable(row[j]) : 0xffffffffu;
  __syncthreads();
  const int i = blockIdx.x * kSortThreads + threadIdx.x;
  if (i >= L) return;
  const uint32_t ki = skeys[i];
  // j < i : count key_j <= key_i ; j > i : count key_j < key_i   (stable tie-break by index)
  int rank = 0;
  const uint4* s4 = reinterpret_cast<const uint4*>(skeys);
  const int i4 = i >> 2;
  int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 4
  for (int q = 0; q < i4; ++q) {
    uint4 v = s4[q];
    r0 += (v.x <= ki); r1 += (v.y <= ki); r2 += (v.z <= ki); r3 += (v.w <= ki);
  }
  {  // the quad containing i
    const int base = i4 << 2;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = base + t;
      const uint32_t kj = skeys[j];
      rank += (j < i) ? (kj <= ki) : ((j > i) ? (kj < ki) : 0);
    }
  }
#pragma unroll 4
  for (int q = i4 + 1; q < (L4 >> 2); ++q) {
    uint4 v = s4[q];
    r0 += (v.x < ki); r1 += (v.y < ki); r2 += (v.z < ki); r3 += (v.w < ki);
  }
  rank += r0 + r1 + r2 + r3;  // padding keys are 0xffffffff: never < ki, and sit after i
  ids_restore[(size_t)b * L + i] = rank;
  mask[(size_t)b * L + i] = (rank >= keep) ? 1.f : 0.f;
  if (rank < keep) ids_keep[(size_t)b * keep + rank] = i;
}

extern "C" int oct_mask_sort(const float* noise, int64_t B, int64_t L, int64_t keep, int64_t* ids_restore,
                             int64_t* ids_keep, float* mask, oct_stream_t stream) {
  OCT_REQUIRE(B >= 0 && L >= 0 && keep >= 0 && keep <= L, "oct_mask_sort: bad sizes B=%lld L=%lld keep=%lld",
              (long long)B, (long long)L, (long long)keep);
  OCT_REQUIRE(B == 0 || L == 0 || (noise && ids_restore && mask), "oct_mask_sort: null pointer");
  OCT_REQUIRE(L <= 16384, "oct_mask_sort: L=%lld > 16384 unsupported", (long long)L);
  OCT_REQUIRE(B == 0 || keep == 0 || ids_keep, "oct_mask_sort: ids_keep is null");
  OCT_REQUIRE(B <= 65535, "oct_mask_sort: B too large");
  if (B == 0 || L == 0) return OCT_OK;
  dim3 grid((unsigned)ceil_div64(L, kSortThreads), (unsigned)B);
  size_t smem = (size_t)((L + 3) & ~3) * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mask_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { oct_set_error("oct_mask_sort: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  }
  mask_sort_kernel<<<grid, kSortThreads, smem, (cudaStream_t)stream>>>(noise, (int)L, (int)keep, ids_restore, ids_keep,
                                                                        mask);
  return oct_check_launch("oct_mask_sort");
}

// ------------------------------------------------------------------------------------------------
// patchify: out row r <- patch of token (b, tok); one CTA per row, float4 per thread
// ------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void patchify_kernel(const float* __restrict__ imgs, TO* __restrict__ out,
                                const int64_t* __restrict__ ids_keep, const int64_t* __restrict__ frame_idx, int T,
                                int H, int W, int p, int u, int rows_per_b, int L) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int hp = H / p, wp = W / p, G = hp * wp;
  const int tok = ids_keep ? (int)ids_keep[(size_t)b * rows_per_b + r] : r;
  const int t = tok / G, s = tok - t * G, h = s / wp, w = s - h * wp;
  const int P = u * p * p, p4 = p >> 2;
  TO* o = out + ((size_t)b * rows_per_b + r) * P;
  for (int e4 = threadIdx.x; e4 < (P >> 2); e4 += blockDim.x) {
    const int kw4 = e4 % p4, kh = (e4 / p4) % p, kt = e4 / (p4 * p);
    int f = t * u + kt;
    if (frame_idx) f = (int)frame_idx[f];
    const float* src = imgs + (((size_t)b * T + f) * H + (h * p + kh)) * W + w * p + kw4 * 4;
    Vec4<TO>::st(o + e4 * 4, *reinterpret_cast<const float4*>(src));
  }
}

extern "C" int oct_patchify(const float* imgs, void* out, int out_dtype, const int64_t* ids_keep,
                            const int64_t* frame_idx, int64_t B, int64_t T, int64_t H, int64_t W, int64_t p, int64_t u,
                            int64_t T_sel, int64_t keep, oct_stream_t stream) {
  OCT_REQUIRE(imgs && out, "oct_patchify: null pointer");
  OCT_REQUIRE(p > 0 && u > 0 && p % 4 == 0 && H % p == 0 && W % p == 0 && T_sel % u == 0,
              "oct_patchify: need p%%4==0, H%%p==0, W%%p==0, T%%u==0 (models:300)");
  OCT_REQUIRE(frame_idx || T_sel == T, "oct_patchify: T_sel != T needs frame_idx");
  const int64_t L = (T_sel / u) * (H / p) * (W / p);
  const int64_t rows = ids_keep ? keep : L;
  if (B == 0 || rows == 0) return OCT_OK;
  OCT_REQUIRE(B <= 65535, "oct_patchify: B too large");
  dim3 grid((unsigned)rows, (unsigned)B);
  const int threads = 192;
  if (out_dtype == OCT_F32)
    patchify_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>(imgs, (float*)out, ids_keep, frame_idx, (int)T,
                                                                       (int)H, (int)W, (int)p, (int)u, (int)rows, (int)L);
  else if (out_dtype == OCT_BF16)
    patchify_kernel<__nv_bfloat16><<<grid, threads, 0, (cudaStream_t)stream>>>(
        imgs, (__nv_bfloat16*)out, ids_keep, frame_idx, (int)T, (int)H, (int)W, (int)p, (int)u, (int)rows, (int)L);
  else
    OCT_REQUIRE(false, "oct_patchify: bad dtype");
  return oct_check_launch("oct_patchify");
}

// ------------------------------------------------------------------------------------------------
// gather_tokens forward: one CTA per output row
// ------------------------------------------------------------------------------------------------
// kGathered: x holds the kept rows only ([B, keep, C], e.g. the gather-first patch embedding) — ids_keep then selects the
// positional rows, not the x row.
template <typename TX, bool kGathered = false>
__global__ void gather_tokens_fwd_kernel(const TX* __restrict__ x, const int64_t* __restrict__ ids_keep,
                                         const float* __restrict__ pos_sp, const float* __restrict__ pos_tmp,
                                         const float* __restrict__ cls_row, float* __restrict__ out, int L, int keep,
                                         int G, int C, int has_cls) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int rows = keep + has_cls;
  float* o = out + ((size_t)b * rows + r) * C;
  if (has_cls && r == 0) {
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4)
      *reinterpret_cast<float4*>(o + c) = *reinterpret_cast<const float4*>(cls_row + c);
    return;
  }
  const int tok = (int)ids_keep[(size_t)b * keep + (r - has_cls)];
  const int t = tok / G, s = tok - t * G;
  const TX* xr = kGathered ? x + ((size_t)b * keep + (r - has_cls)) * C : x + ((size_t)b * L + tok) * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 v = Vec4<TX>::ld(xr + c);
    if (pos_sp) {
      float4 a = *reinterpret_cast<const float4*>(pos_sp + (size_t)s * C + c);
      if (pos_tmp) {
        float4 tt = *reinterpret_cast<const float4*>(pos_tmp + (size_t)t * C + c);
        // reference order: (spatial.repeat + temporal.repeat_interleave) first, then x + pos  (models:429-436,478)
        a.x += tt.x; a.y += tt.y; a.z += tt.z; a.w += tt.w;
      }
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(o + c) = v;
  }
}

extern "C" int oct_gather_tokens_fwd(const void* x, int x_dtype, const int64_t* ids_keep, const float* pos_sp,
                                     const float* pos_tmp, const float* cls_row, float* out, int64_t B, int64_t L,
                                     int64_t keep, int64_t G, int64_t C, oct_stream_t stream) {
  OCT_REQUIRE(x && ids_keep && out, "oct_gather_tokens_fwd: null pointer");
  if (!pos_sp) { OCT_REQUIRE(!pos_tmp, "oct_gather_tokens_fwd: pos_tmp without pos_sp"); G = L > 0 ? L : 1; }
  OCT_REQUIRE(C % 4 == 0 && G > 0 && L % G == 0, "oct_gather_tokens_fwd: need C%%4==0 and L%%G==0");
  OCT_REQUIRE(!pos_sp || pos_tmp || L == G, "oct_gather_tokens_fwd: pos_tmp may be NULL only when T'==1");
  const int has_cls = cls_row ? 1 : 0;
  if (B == 0 || keep + has_cls == 0) return OCT_OK;
  dim3 grid((unsigned)(keep + has_cls), (unsigned)B);
  const int threads = (int)((C / 4 < 256) ? ((C / 4 + 31) / 32 * 32) : 256);
  if (x_dtype == OCT_F32)
    gather_tokens_fwd_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>(
        (const float*)x, ids_keep, pos_sp, pos_tmp, cls_row, out, (int)L, (int)keep, (int)G, (int)C, has_cls);
  else if (x_dtype == OCT_BF16)
    gather_tokens_fwd_kernel<__nv_bfloat16><<<grid, threads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, ids_keep, pos_sp, pos_tmp, cls_row, out, (int)L, (int)keep, (int)G, (int)C, has_cls);
  else
    OCT_REQUIRE(false, "oct_gather_tokens_fwd: bad dtype");
  return oct_check_launch("oct_gather_tokens_fwd");
}

// x_keep [B, keep, C] (rows already gathered) + cls row + positional rows selected by ids_keep -> out [B, keep (+1), C] f32
extern "C" int oct_posadd_tokens_fwd(const void* x_keep, int x_dtype, const int64_t* ids_keep, const float* pos_sp,
                                     const float* pos_tmp, const float* cls_row, float* out, int64_t B, int64_t L,
                                     int64_t keep, int64_t G, int64_t C, oct_stream_t stream) {
  OCT_REQUIRE(x_keep && ids_keep && out && pos_sp, "oct_posadd_tokens_fwd: null pointer");
  OCT_REQUIRE(C % 4 == 0 && G > 0 && L % G == 0, "oct_posadd_tokens_fwd: need C%%4==0 and L%%G==0");
  OCT_REQUIRE(pos_tmp || L == G, "oct_posadd_tokens_fwd: pos_tmp may be NULL only when T'==1");
  const int has_cls = cls_row ? 1 : 0;
  if (B == 0 || keep + has_cls == 0) return OCT_OK;
  dim3 grid((unsigned)(keep + has_cls), (unsigned)B);
  const int threads = (int)((C / 4 < 256) ? ((C / 4 + 31) / 32 * 32) : 256);
  if (x_dtype == OCT_F32)
    gather_tokens_fwd_kernel<float, true><<<grid, threads, 0, (cudaStream_t)stream>>>(
        (const float*)x_keep, ids_keep, pos_sp, pos_tmp, cls_row, out, (int)L, (int)keep, (int)G, (int)C, has_cls);
  else if (x_dtype == OCT_BF16)
    gather_tokens_fwd_kernel<__nv_bfloat16, true><<<grid, threads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x_keep, ids_keep, pos_sp, pos_tmp, cls_row, out, (int)L, (int)keep, (int)G, (int)C, has_cls);
  else
    OCT_REQUIRE(false, "oct_posadd_tokens_fwd: bad dtype");
  return oct_check_launch("oct_posadd_tokens_fwd");
}

// backward part 1: dx_keep rows (plain copy/cast of dout rows 1..)
template <typename TD>
__global__ void gather_tokens_bwd_copy_kernel(const float* __restrict__ dout, TD* __restrict__ dx, int keep, int C,
                                              int has_cls) {
  const int r = blockIdx.x, b = blockIdx.y;
  const float* src = dout + ((size_t)b * (keep + has_cls) + r + has_cls) * C;
  TD* dst = dx + ((size_t)b * keep + r) * C;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4)
    Vec4<TD>::st(dst + c, *reinterpret_cast<const float4*>(src + c));
}

// backward part 2: deterministic segmented sums.  CTA x in [0,Gs): spatial slot x ; [Gs, Gs+Tp): temporal slot ; last: cls.
// Each CTA first collects the kept tokens that fall into its slot (parallel scan of the B*keep ids into a shared list),
// orders that short list by flat index, then accumulates the matching rows in that fixed order (reproducible sums).
constexpr int kPosMaxList = 2048;
__global__ void gather_tokens_bwd_pos_kernel(const float* __restrict__ dout, const int64_t* __restrict__ ids_keep,
                                             float* __restrict__ d_pos_sp, float* __restrict__ d_pos_tmp,
                                             float* __restrict__ d_cls_row, int B, int keep, int G, int Gs, int Tp,
                                             int C, int has_cls) {
  __shared__ int s_list[kPosMaxList];
  __shared__ int s_sorted[kPosMaxList];
  __shared__ int s_count;
  const int slot = blockIdx.x;
  const int rows = keep + has_cls;
  if (slot == Gs + Tp) {  // cls
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int b = 0; b < B; ++b) {
        float4 v = *reinterpret_cast<const float4*>(dout + (size_t)b * rows * C + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(d_cls_row + c) = acc;
    }
    return;
  }
  const bool spatial = slot < Gs;
  const int want = spatial ? slot : slot - Gs;
  // each thread owns up to 4 float4 column chunks (C <= 4 * 4 * blockDim.x, checked by the host)
  float4 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int total = B * keep;
  // flat-index segments of at most kPosMaxList ids: a segment's matches always fit the shared list, and processing
  // segments in increasing order keeps the global summation order = increasing flat index
  for (int seg = 0; seg < total; seg += kPosMaxList) {
    __syncthreads();
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int seg_end = min(total, seg + kPosMaxList);
    for (int f = seg + threadIdx.x; f < seg_end; f += blockDim.x) {
      const int tok = (int)ids_keep[f];
      const int key = spatial ? (tok % G) : (tok / G);
      if (key == want) s_list[atomicAdd(&s_count, 1)] = f;
    }
    __syncthreads();
    const int n = s_count;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {  // rank sort (flat indices are unique)
      const int v = s_list[i];
      int r = 0;
      for (int j = 0; j < n; ++j) r += (s_list[j] < v);
      s_sorted[r] = v;
    }
    __syncthreads();
    // rows are ADDED in list order (reproducible), but fetched eight at a time: a temporal slot collects ~B*keep/Tp rows
    // (204 at the bench shape), and one dependent L2 round trip per row made the 16 temporal CTAs the whole kernel (92 us)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = (threadIdx.x + k * blockDim.x) * 4;
      if (c >= C) continue;
      for (int i = 0; i < n; i += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (i + j < n) {
            const int f = s_sorted[i + j];
            const int b = f / keep, r = f - b * keep;
            v[j] = *reinterpret_cast<const float4*>(dout + ((size_t)b * rows + r + has_cls) * C + c);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (i + j < n) { acc[k].x += v[j].x; acc[k].y += v[j].y; acc[k].z += v[j].z; acc[k].w += v[j].w; }
        }
      }
    }
  }
  float* dst = spatial ? (d_pos_sp + (size_t)want * C) : (d_pos_tmp + (size_t)want * C);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = (threadIdx.x + k * blockDim.x) * 4;
    if (c < C) *reinterpret_cast<float4*>(dst + c) = acc[k];
  }
}

extern "C" int oct_gather_tokens_bwd(const float* dout, const int64_t* ids_keep, void* dx_keep, int dx_dtype,
                                     float* d_pos_sp, float* d_pos_tmp, float* d_cls_row, int64_t B, int64_t L,
                                     int64_t keep, int64_t G, int64_t C, oct_stream_t stream) {
  OCT_REQUIRE(dout && ids_keep, "oct_gather_tokens_bwd: null pointer");
  if (!d_pos_sp) { OCT_REQUIRE(!d_pos_tmp, "oct_gather_tokens_bwd: d_pos_tmp without d_pos_sp"); G = L > 0 ? L : 1; }
  OCT_REQUIRE(C % 4 == 0 && G > 0 && L % G == 0, "oct_gather_tokens_bwd: need C%%4==0 and L%%G==0");
  const int has_cls = d_cls_row ? 1 : 0;
  const int Tp = d_pos_tmp ? (int)(L / G) : 0;
  if (B == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = (int)((C / 4 < 256) ? ((C / 4 + 31) / 32 * 32) : 256);
  if (dx_keep && keep > 0) {
    dim3 grid((unsigned)keep, (unsigned)B);
    if (dx_dtype == OCT_F32)
      gather_tokens_bwd_copy_kernel<float><<<grid, threads, 0, st>>>(dout, (float*)dx_keep, (int)keep, (int)C, has_cls);
    else if (dx_dtype == OCT_BF16)
      gather_tokens_bwd_copy_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(dout, (__nv_bfloat16*)dx_keep, (int)keep,
                                                                            (int)C, has_cls);
    else
      OCT_REQUIRE(false, "oct_gather_tokens_bwd: bad dtype");
    int rc = oct_check_launch("oct_gather_tokens_bwd(copy)");
    if (rc) return rc;
  }
  const int Gs = d_pos_sp ? (int)G : 0;  // no spatial slots when the caller has no pos tables (plain random_masking)
  if (Gs + Tp + has_cls == 0) return OCT_OK;
  OCT_REQUIRE(C <= 16 * threads, "oct_gather_tokens_bwd: C too large");
  dim3 grid2((unsigned)(Gs + Tp + has_cls), 1);
  // slot numbering inside the kernel: [0,Gs) spatial, [Gs,Gs+Tp) temporal, Gs+Tp cls
  gather_tokens_bwd_pos_kernel<<<grid2, threads, 0, st>>>(dout, ids_keep, d_pos_sp, d_pos_tmp, d_cls_row, (int)B,
                                                          (int)keep, (int)G, Gs, Tp, (int)C, has_cls);
  return oct_check_launch("oct_gather_tokens_bwd(pos)");
}

// ------------------------------------------------------------------------------------------------
// decoder unshuffle forward: one CTA per output row
// ------------------------------------------------------------------------------------------------
template <typename TY>
__global__ void unshuffle_fwd_kernel(const TY* __restrict__ y, const int64_t* __restrict__ ids_restore,
                                     const float* __restrict__ mask_token, const float* __restrict__ pos_sp,
                                     const float* __restrict__ pos_tmp, const float* __restrict__ cls_row,
                                     float* __restrict__ out, int L, int keep, int G, int D, int has_cls,
                                     int y_row0) {
  const int r = blockIdx.x, b = blockIdx.y;
  float* o = out + ((size_t)b * (L + has_cls) + r) * D;
  const TY* yb = y + (size_t)b * (keep + y_row0) * D;  // this sample's rows: [cls row (y_row0 == 1)] + keep token rows
  if (has_cls && r == 0) {
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
      float4 v = *reinterpret_cast<const float4*>(cls_row + c);
      if (y_row0) {  // per-sample cls token that went through decoder_embed (OCTCube/models_mae_flash_attn.py:301-309)
        const float4 a = Vec4<TY>::ld(yb + c);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      *reinterpret_cast<float4*>(o + c) = v;
    }
    return;
  }
  const int j = r - has_cls;
  const int src = (int)ids_restore[(size_t)b * L + j];
  const int t = j / G, s = j - t * G;
  const TY* yr = yb + ((size_t)y_row0 + src) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 v = (src < keep) ? Vec4<TY>::ld(yr + c) : *reinterpret_cast<const float4*>(mask_token + c);
    float4 a = *reinterpret_cast<const float4*>(pos_sp + (size_t)s * D + c);
    if (pos_tmp) {
      float4 tt = *reinterpret_cast<const float4*>(pos_tmp + (size_t)t * D + c);
      a.x += tt.x; a.y += tt.y; a.z += tt.z; a.w += tt.w;
    }
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    *reinterpret_cast<float4*>(o + c) = v;
  }
}

extern "C" int oct_unshuffle_fwd(const void* y, int y_dtype, const int64_t* ids_restore, const float* mask_token,
                                 const float* pos_sp, const float* pos_tmp, const float* cls_row, float* out, int64_t B,
                                 int64_t L, int64_t keep, int64_t G, int64_t D, int64_t y_row0, oct_stream_t stream) {
  OCT_REQUIRE(ids_restore && mask_token && pos_sp && out && (y || keep + y_row0 == 0), "oct_unshuffle_fwd: null pointer");
  OCT_REQUIRE(y_row0 == 0 || (y_row0 == 1 && cls_row), "oct_unshuffle_fwd: y_row0 must be 0, or 1 together with cls_row");
  OCT_REQUIRE(D % 4 == 0 && G > 0 && L % G == 0, "oct_unshuffle_fwd: need D%%4==0 and L%%G==0");
  OCT_REQUIRE(pos_tmp || L == G, "oct_unshuffle_fwd: pos_tmp may be NULL only when T'==1");
  const int has_cls = cls_row ? 1 : 0;
  if (B == 0 || L + has_cls == 0) return OCT_OK;
  dim3 grid((unsigned)(L + has_cls), (unsigned)B);
  const int threads = (int)((D / 4 < 256) ? ((D / 4 + 31) / 32 * 32) : 256);
  if (y_dtype == OCT_F32)
    unshuffle_fwd_kernel<float><<<grid, threads, 0, (cudaStream_t)stream>>>(
        (const float*)y, ids_restore, mask_token, pos_sp, pos_tmp, cls_row, out, (int)L, (int)keep, (int)G, (int)D, has_cls,
        (int)y_row0);
  else if (y_dtype == OCT_BF16)
    unshuffle_fwd_kernel<__nv_bfloat16><<<grid, threads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)y, ids_restore, mask_token, pos_sp, pos_tmp, cls_row, out, (int)L, (int)keep, (int)G, (int)D,
        has_cls, (int)y_row0);
  else
    OCT_REQUIRE(false, "oct_unshuffle_fwd: bad dtype");
  return oct_check_launch("oct_unshuffle_fwd");
}

// backward 1: dy[b, y_row0 + r] = dout[b, 1 + j] where r = ids_restore[b, j] < keep (a permutation: every r written once);
// with y_row0 == 1 the extra CTA j == L copies the cls row: dy[b, 0] = dout[b, 0]
template <typename TD>
__global__ void unshuffle_bwd_scatter_kernel(const float* __restrict__ dout, const int64_t* __restrict__ ids_restore,
                                             TD* __restrict__ dy, int L, int keep, int D, int has_cls, int y_row0) {
  const int j = blockIdx.x, b = blockIdx.y;
  TD* dyb = dy + (size_t)b * (keep + y_row0) * D;
  if (j == L) {
    const float* src = dout + (size_t)b * (L + has_cls) * D;
    for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
      Vec4<TD>::st(dyb + c, *reinterpret_cast<const float4*>(src + c));
    return;
  }
  const int r = (int)ids_restore[(size_t)b * L + j];
  if (r >= keep) return;
  const float* src = dout + ((size_t)b * (L + has_cls) + j + has_cls) * D;
  TD* dst = dyb + ((size_t)y_row0 + r) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
    Vec4<TD>::st(dst + c, *reinterpret_cast<const float4*>(src + c));
}

// backward 2a: per (t, b) partial sums over the G spatial slots: all rows -> ws_tmp[b,t,:], masked rows -> ws_mt[b,t,:]
__global__ void unshuffle_bwd_rowsum_kernel(const float* __restrict__ dout, const int64_t* __restrict__ ids_restore,
                                            float* __restrict__ ws_tmp, float* __restrict__ ws_mt, int L, int keep,
                                            int G, int D, int has_cls) {
  const int t = blockIdx.x, b = blockIdx.y, Tp = gridDim.x;
  const int c = (blockIdx.z * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), m = a;
  const float* base = dout + ((size_t)b * (L + has_cls) + has_cls + (size_t)t * G) * D + c;
  const int64_t* ids = ids_restore + (size_t)b * L + (size_t)t * G;
  // rows are added in order, fetched eight at a time (128 CTAs x 4 warps: the loop is bound by load latency, not bandwidth)
  for (int s0 = 0; s0 < G; s0 += 8) {
    float4 v[8];
    bool msk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (s0 + j < G) {
        v[j] = *reinterpret_cast<const float4*>(base + (size_t)(s0 + j) * D);
        msk[j] = ids[s0 + j] >= keep;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (s0 + j < G) {
        a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w;
        if (msk[j]) { m.x += v[j].x; m.y += v[j].y; m.z += v[j].z; m.w += v[j].w; }
      }
    }
  }
  *reinterpret_cast<float4*>(ws_tmp + ((size_t)b * Tp + t) * D + c) = a;
  *reinterpret_cast<float4*>(ws_mt + ((size_t)b * Tp + t) * D + c) = m;
}

// backward 2b: finish temporal / mask-token / cls sums (fixed order over b, t)
__global__ void unshuffle_bwd_finish_kernel(const float* __restrict__ dout, const float* __restrict__ ws_tmp,
                                            const float* __restrict__ ws_mt, float* __restrict__ d_pos_tmp,
                                            float* __restrict__ d_mask_token, float* __restrict__ d_cls_row, int B,
                                            int Tp, int L, int D, int has_cls) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float mt = 0.f;
  for (int t = 0; t < Tp; ++t) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) {
      a += ws_tmp[((size_t)b * Tp + t) * D + c];
      mt += ws_mt[((size_t)b * Tp + t) * D + c];
    }
    if (d_pos_tmp) d_pos_tmp[(size_t)t * D + c] = a;
  }
  d_mask_token[c] = mt;
  if (d_cls_row) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += dout[(size_t)b * (L + has_cls) * D + c];
    d_cls_row[c] = a;
  }
}

// backward 2c: spatial table: d_pos_sp[s] = sum_{b,t} dout[b, 1 + t*G + s]
__global__ void unshuffle_bwd_spatial_kernel(const float* __restrict__ dout, float* __restrict__ d_pos_sp, int B, int L,
                                             int G, int D, int has_cls) {
  const int s = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const int Tp = L / G;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < Tp; ++t) {
      float4 v = *reinterpret_cast<const float4*>(dout + ((size_t)b * (L + has_cls) + has_cls + (size_t)t * G + s) * D + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  *reinterpret_cast<float4*>(d_pos_sp + (size_t)s * D + c) = a;
}

extern "C" size_t oct_unshuffle_bwd_ws_bytes(int64_t B, int64_t L, int64_t G, int64_t D) {
  if (G <= 0) return 0;
  return (size_t)2 * B * (L / G) * D * sizeof(float);
}

extern "C" int oct_unshuffle_bwd(const float* dout, const int64_t* ids_restore, void* dy, int dy_dtype,
                                 float* d_mask_token, float* d_pos_sp, float* d_pos_tmp, float* d_cls_row, void* ws,
                                 size_t ws_bytes, int64_t B, int64_t L, int64_t keep, int64_t G, int64_t D, int has_cls,
                                 int64_t y_row0, oct_stream_t stream) {
  OCT_REQUIRE(dout && ids_restore && d_mask_token && d_pos_sp, "oct_unshuffle_bwd: null pointer");
  OCT_REQUIRE(y_row0 == 0 || (y_row0 == 1 && has_cls && dy), "oct_unshuffle_bwd: y_row0 must be 0, or 1 together with has_cls and dy");
  OCT_REQUIRE(D % 4 == 0 && G > 0 && L % G == 0, "oct_unshuffle_bwd: need D%%4==0 and L%%G==0");
  OCT_REQUIRE((has_cls != 0) == (d_cls_row != nullptr), "oct_unshuffle_bwd: has_cls / d_cls_row mismatch");
  if (ws_bytes < oct_unshuffle_bwd_ws_bytes(B, L, G, D) || !ws) {
    oct_set_error("oct_unshuffle_bwd: workspace too small");
    return OCT_ERR_WORKSPACE;
  }
  if (B == 0 || L == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int Tp = (int)(L / G);
  const int threads = (int)((D / 4 < 256) ? ((D / 4 + 31) / 32 * 32) : 256);
  if (dy && keep + y_row0 > 0) {
    dim3 grid((unsigned)(L + y_row0), (unsigned)B);
    if (dy_dtype == OCT_F32)
      unshuffle_bwd_scatter_kernel<float><<<grid, threads, 0, st>>>(dout, ids_restore, (float*)dy, (int)L, (int)keep,
                                                                    (int)D, has_cls, (int)y_row0);
    else if (dy_dtype == OCT_BF16)
      unshuffle_bwd_scatter_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(dout, ids_restore, (__nv_bfloat16*)dy,
                                                                           (int)L, (int)keep, (int)D, has_cls, (int)y_row0);
    else
      OCT_REQUIRE(false, "oct_unshuffle_bwd: bad dtype");
    int rc = oct_check_launch("oct_unshuffle_bwd(scatter)");
    if (rc) return rc;
  }
  float* ws_tmp = (float*)ws;
  float* ws_mt = ws_tmp + (size_t)B * Tp * D;
  dim3 g1((unsigned)Tp, (unsigned)B, (unsigned)ceil_div64(D / 4, threads));
  unshuffle_bwd_rowsum_kernel<<<g1, threads, 0, st>>>(dout, ids_restore, ws_tmp, ws_mt, (int)L, (int)keep, (int)G,
                                                      (int)D, has_cls);
  int rc = oct_check_launch("oct_unshuffle_bwd(rowsum)");
  if (rc) return rc;
  unshuffle_bwd_finish_kernel<<<(unsigned)ceil_div64(D, 128), 128, 0, st>>>(dout, ws_tmp, ws_mt, d_pos_tmp, d_mask_token,
                                                                           d_cls_row, (int)B, Tp, (int)L, (int)D, has_cls);
  rc = oct_check_launch("oct_unshuffle_bwd(finish)");
  if (rc) return rc;
  dim3 g3((unsigned)G, (unsigned)ceil_div64(D / 4, threads));
  unshuffle_bwd_spatial_kernel<<<g3, threads, 0, st>>>(dout, d_pos_sp, (int)B, (int)L, (int)G, (int)D, has_cls);
  return oct_check_launch("oct_unshuffle_bwd(spatial)");
}
