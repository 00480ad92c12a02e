// forward_loss (models_mae_joint_res_flash_attn.py:613-667): masked MSE over (optionally per-patch normalised) pixels.
// The target patch is read straight from the volume through patch indexing — no patchify copy (models:628-642) —
// and only masked tokens are touched (kept tokens have weight 0 in every sum of models:649-663).
#include "common.cuh"

constexpr int kLossThreads = 192;  // 192 threads x float4 = 768 = 3*16*16 pixels in one pass

constexpr int kLossNormPix = 1, kLossChannelLast = 2, kLossAllTokens = 4;  // OCT_LOSS_* of include/octcube_b200.h

// loads element quad e4 of the target patch of token (b,t,h,w)
__device__ __forceinline__ float4 load_target4(const float* __restrict__ imgs, const int64_t* __restrict__ frame_idx,
                                               int b, int t, int h, int w, int e4, int T, int H, int W, int p, int u,
                                               bool chan_last) {
  if (chan_last) {
    // 2D model (OCTCube/models_mae_flash_attn.py:214-226, 'nchpwq->nhwpqc'): element e = (kh * p + kw) * u + c, the u
    // channels are the "frames" of the [B,u,H,W] image; four consecutive elements straddle channels -> scalar loads
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = e4 * 4 + i;
      const int c = e % u, q = e / u;
      const int kw = q % p, kh = q / p;
      r[i] = imgs[(((size_t)b * T + (t * u + c)) * H + (h * p + kh)) * W + w * p + kw];
    }
    return make_float4(r[0], r[1], r[2], r[3]);
  }
  const int p4 = p >> 2;
  const int kw4 = e4 % p4, kh = (e4 / p4) % p, kt = e4 / (p4 * p);
  int f = t * u + kt;
  if (frame_idx) f = (int)frame_idx[f];
  return *reinterpret_cast<const float4*>(imgs + (((size_t)b * T + f) * H + (h * p + kh)) * W + w * p + kw4 * 4);
}

template <typename TP, bool kBackward, typename TD>
__global__ void __launch_bounds__(kLossThreads) mse_loss_token_kernel(
    const float* __restrict__ imgs, const int64_t* __restrict__ frame_idx, const TP* __restrict__ pred,
    const float* __restrict__ mask, float* __restrict__ loss_tok, const float* __restrict__ mask_sum,
    const float* __restrict__ dloss, TD* __restrict__ dpred, int T, int H, int W, int p, int u, int L, int pred_rows,
    int pred_row0, int flags) {
  __shared__ float sred[32];
  const bool norm_pix = flags & kLossNormPix, cl = flags & kLossChannelLast;
  const int b = blockIdx.y;
  const int P = u * p * p, P4 = P >> 2;
  int j;  // token index
  if (kBackward) {
    const int r = blockIdx.x;  // dpred row
    j = r - pred_row0;
    TD* drow = dpred + ((size_t)b * pred_rows + r) * P;
    const bool live = (j >= 0) && (j < L) && (mask[(size_t)b * L + j] != 0.f);  // rows past the L tokens: zero gradient
    if (!live) {
      for (int e4 = threadIdx.x; e4 < P4; e4 += kLossThreads) Vec4<TD>::st(drow + e4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      return;
    }
  } else {
    j = blockIdx.x;
    if (!(flags & kLossAllTokens) && mask[(size_t)b * L + j] == 0.f) {
      if (threadIdx.x == 0) loss_tok[(size_t)b * L + j] = 0.f;
      return;
    }
  }
  const int hp = H / p, wp = W / p, G = hp * wp;
  const int t = j / G, s = j - t * G, h = s / wp, w = s - h * wp;
  const TP* prow = pred + ((size_t)b * pred_rows + pred_row0 + j) * P;

  // The recipe's patch (3 x 16 x 16 = 768 pixels) is exactly one float4 per thread: it is fetched ONCE and kept in registers
  // across the mean / variance / error passes (three dependent trips to L1 per token before).
  const bool single = (P4 <= kLossThreads);
  float4 tg0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (single && (int)threadIdx.x < P4) tg0 = load_target4(imgs, frame_idx, b, t, h, w, threadIdx.x, T, H, W, p, u, cl);
  auto target = [&](int e4) { return single ? tg0 : load_target4(imgs, frame_idx, b, t, h, w, e4, T, H, W, p, u, cl); };

  float mean = 0.f, inv_std = 1.f;
  if (norm_pix) {  // models:644-647 — mean, UNBIASED variance, eps 1e-6
    float sum = 0.f;
    for (int e4 = threadIdx.x; e4 < P4; e4 += kLossThreads) {
      float4 v = target(e4);
      sum += (v.x + v.y) + (v.z + v.w);
    }
    mean = block_sum(sum, sred) / (float)P;
    float sq = 0.f;
    for (int e4 = threadIdx.x; e4 < P4; e4 += kLossThreads) {
      float4 v = target(e4);
      const float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
      sq += (a * a + c * c) + (d * d + e * e);
    }
    const float var = block_sum(sq, sred) / (float)(P - 1);
    inv_std = 1.f / sqrtf(var + 1.0e-6f);
  }

  if (!kBackward) {
    float acc = 0.f;
    for (int e4 = threadIdx.x; e4 < P4; e4 += kLossThreads) {
      float4 tg = target(e4);
      const float4 pr = Vec4<TP>::ld(prow + e4 * 4);
      const float a = pr.x - (tg.x - mean) * inv_std, c = pr.y - (tg.y - mean) * inv_std;
      const float d = pr.z - (tg.z - mean) * inv_std, e = pr.w - (tg.w - mean) * inv_std;
      acc += (a * a + c * c) + (d * d + e * e);
    }
    acc = block_sum(acc, sred);
    if (threadIdx.x == 0) loss_tok[(size_t)b * L + j] = acc / (float)P;
  } else {
    const float coef = dloss[0] * 2.f / ((float)P * mask_sum[0]);
    TD* drow = dpred + ((size_t)b * pred_rows + pred_row0 + j) * P;
    for (int e4 = threadIdx.x; e4 < P4; e4 += kLossThreads) {
      float4 tg = target(e4);
      const float4 pr = Vec4<TP>::ld(prow + e4 * 4);
      float4 o;
      o.x = coef * (pr.x - (tg.x - mean) * inv_std);
      o.y = coef * (pr.y - (tg.y - mean) * inv_std);
      o.z = coef * (pr.z - (tg.z - mean) * inv_std);
      o.w = coef * (pr.w - (tg.w - mean) * inv_std);
      Vec4<TD>::st(drow + e4 * 4, o);
    }
  }
}

// One CTA: per-(b,t') masked sums in a fixed order, then the total (models:655-663).
__global__ void __launch_bounds__(1024) mse_loss_finish_kernel(const float* loss_tok, const float* mask,
                                                                    float* loss, float* mask_sum, float* frame_losses,
                                                                    int BT, int G) {
  __shared__ float part[2 * 4096];
  __shared__ float tot[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) { tot[0] = 0.f; tot[1] = 0.f; }
  for (int base = 0; base < BT; base += 4096) {  // any B * T': 4096 (b,t') groups per pass, totals carried in a fixed order
    const int n = min(4096, BT - base);
    for (int i = warp; i < n; i += nw) {
      const int g = base + i;
      float ls = 0.f, ms = 0.f;
      for (int s = lane; s < G; s += 32) {
        const float m = mask[(size_t)g * G + s];
        ls += loss_tok[(size_t)g * G + s] * m;
        ms += m;
      }
      ls = warp_sum(ls);
      ms = warp_sum(ms);
      if (lane == 0) {
        part[i] = ls;
        part[4096 + i] = ms;
        frame_losses[g] = ls / (ms + 1e-6f);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float ls = tot[0], ms = tot[1];
      for (int i = 0; i < n; ++i) { ls += part[i]; ms += part[4096 + i]; }
      tot[0] = ls; tot[1] = ms;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss[0] = tot[0] / tot[1];
    mask_sum[0] = tot[1];
  }
}

static int check_loss_args(int64_t B, int64_t T, int64_t T_sel, const int64_t* frame_idx, int64_t H, int64_t W,
                           int64_t p, int64_t u, int64_t pred_rows, int64_t pred_row0, int64_t* L_out, const char* who) {
  OCT_REQUIRE(p > 0 && u > 0 && p % 4 == 0 && H % p == 0 && W % p == 0 && T_sel % u == 0, "%s: bad patch geometry", who);
  OCT_REQUIRE(frame_idx || T_sel == T, "%s: T_sel != T needs frame_idx", who);
  const int64_t L = (T_sel / u) * (H / p) * (W / p);
  OCT_REQUIRE(pred_row0 >= 0 && pred_rows >= pred_row0 + L, "%s: pred_rows < pred_row0 + L", who);
  OCT_REQUIRE(B <= 65535, "%s: B too large", who);
  *L_out = L;
  return OCT_OK;
}

extern "C" int oct_mse_loss_fwd(const float* imgs, const int64_t* frame_idx, const void* pred, int pred_dtype,
                                const float* mask, float* loss_tok, float* loss, float* mask_sum, float* frame_losses,
                                int64_t B, int64_t T, int64_t T_sel, int64_t H, int64_t W, int64_t p, int64_t u,
                                int64_t pred_rows, int64_t pred_row0, int flags, oct_stream_t stream) {
  OCT_REQUIRE(imgs && pred && mask && loss_tok && loss && mask_sum && frame_losses, "oct_mse_loss_fwd: null pointer");
  OCT_REQUIRE(!(flags & kLossChannelLast) || (!frame_idx && T == u), "oct_mse_loss_fwd: channel-last needs T == u, no frame_idx");
  int64_t L;
  int rc = check_loss_args(B, T, T_sel, frame_idx, H, W, p, u, pred_rows, pred_row0, &L, "oct_mse_loss_fwd");
  if (rc) return rc;
  if (B == 0 || L == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)L, (unsigned)B);
#define LAUNCH(TP)                                                                                                   \
  mse_loss_token_kernel<TP, false, float><<<grid, kLossThreads, 0, st>>>(imgs, frame_idx, (const TP*)pred, mask,     \
      loss_tok, nullptr, nullptr, nullptr, (int)T, (int)H, (int)W, (int)p, (int)u, (int)L, (int)pred_rows,            \
      (int)pred_row0, flags)
  if (pred_dtype == OCT_F32) LAUNCH(float);
  else if (pred_dtype == OCT_BF16) LAUNCH(__nv_bfloat16);
  else OCT_REQUIRE(false, "oct_mse_loss_fwd: bad dtype");
#undef LAUNCH
  rc = oct_check_launch("oct_mse_loss_fwd");
  if (rc) return rc;
  const int G = (int)((H / p) * (W / p));
  const int BT = (int)(B * (T_sel / u));
  mse_loss_finish_kernel<<<1, 1024, 0, st>>>(loss_tok, mask, loss, mask_sum, frame_losses, BT, G);
  return oct_check_launch("oct_mse_loss_fwd(finish)");
}

extern "C" int oct_mse_loss_bwd(const float* imgs, const int64_t* frame_idx, const void* pred, int pred_dtype,
                                const float* mask, const float* mask_sum, const float* dloss, void* dpred,
                                int dpred_dtype, int64_t B, int64_t T, int64_t T_sel, int64_t H, int64_t W, int64_t p,
                                int64_t u, int64_t pred_rows, int64_t pred_row0, int flags, oct_stream_t stream) {
  OCT_REQUIRE(imgs && pred && mask && mask_sum && dloss && dpred, "oct_mse_loss_bwd: null pointer");
  OCT_REQUIRE(!(flags & kLossChannelLast) || (!frame_idx && T == u), "oct_mse_loss_bwd: channel-last needs T == u, no frame_idx");
  OCT_REQUIRE(pred_dtype == dpred_dtype, "oct_mse_loss_bwd: dpred dtype must equal pred dtype");
  int64_t L;
  int rc = check_loss_args(B, T, T_sel, frame_idx, H, W, p, u, pred_rows, pred_row0, &L, "oct_mse_loss_bwd");
  if (rc) return rc;
  if (B == 0 || pred_rows == 0) return OCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)pred_rows, (unsigned)B);
#define LAUNCH(TP)                                                                                                   \
  mse_loss_token_kernel<TP, true, TP><<<grid, kLossThreads, 0, st>>>(imgs, frame_idx, (const TP*)pred, mask, nullptr, \
      mask_sum, dloss, (TP*)dpred, (int)T, (int)H, (int)W, (int)p, (int)u, (int)L, (int)pred_rows, (int)pred_row0,    \
      flags)
  if (pred_dtype == OCT_F32) LAUNCH(float);
  else if (pred_dtype == OCT_BF16) LAUNCH(__nv_bfloat16);
  else OCT_REQUIRE(false, "oct_mse_loss_bwd: bad dtype");
#undef LAUNCH
  return oct_check_launch("oct_mse_loss_bwd");
}
