// Volume ingest on the device: the reference's loader turns the uint8 B-scans of a cube into the fp32
// [B,1,T,H,W] step input on the CPU — ToTensor's /255 (PatientDataset_inhouse.py:420), centre padding with zero frames or
// centre cropping to `padding_num_frames` (:436-450), RandFlipd along the frame axis and along the width
// (create_3d_transforms, :59-62) — and ships 4 bytes per pixel over PCIe.  Here the uint8 cube crosses the bus (1 byte per
// pixel) and ONE kernel writes the step's fp32 input: scale, pad / crop and both flips are index arithmetic on the way.
// All four operations are exact, so the result is bit-identical to the CPU pipeline's.  HBM-bound: 1 B read + 4 B written.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) ingest_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                        const uint8_t* __restrict__ flip_t, const uint8_t* __restrict__ flip_w,
                                                        int T_src, int T, int H, int W, int shift, float divisor) {
  const int W4 = W >> 2;
  const int64_t per_sample = (int64_t)T * H * W4;
  const int b = blockIdx.y;
  const bool ft = flip_t && flip_t[b], fw = flip_w && flip_w[b];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic (the host checks T*H*W/4 < 2^31): three 64-bit divisions per item made the first version of
    // this kernel instruction-bound (57 us for 25 M pixels, profiles/r1_membound_ncu.md)
    const unsigned q = (unsigned)i / (unsigned)W4;
    const int w4 = (int)((unsigned)i - q * (unsigned)W4);
    const int t = (int)(q / (unsigned)H);
    const int h = (int)(q - (unsigned)t * (unsigned)H);
    const int tp = ft ? T - 1 - t : t;     // frame of the padded / cropped cube that lands at output frame t
    const int ts = tp - shift;             // its index in the source cube (shift > 0: left padding, < 0: cropping)
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ts >= 0 && ts < T_src) {
      const int ws = fw ? W - 4 - 4 * w4 : 4 * w4;
      const uchar4 v = *reinterpret_cast<const uchar4*>(src + (((int64_t)b * T_src + ts) * H + h) * W + ws);
      const float a = __fdiv_rn((float)v.x, divisor), c = __fdiv_rn((float)v.y, divisor);
      const float d = __fdiv_rn((float)v.z, divisor), e = __fdiv_rn((float)v.w, divisor);
      o = fw ? make_float4(e, d, c, a) : make_float4(a, c, d, e);
    }
    *reinterpret_cast<float4*>(dst + (((int64_t)b * T + t) * H + h) * W + 4 * w4) = o;
  }
}


// ---- CropForegroundd + Resized(trilinear) + RandFlipd of create_3d_transforms (PatientDataset_inhouse.py:56-63) ----------
// The loader pads / crops the cube to `padding_num_frames` first (:436-450), THEN crops the bounding box of the non-zero
// voxels (monai CropForegroundd: select_fn = x > 0, margin 0) and resamples the box to (num_frames, H, W) with
// F.interpolate(mode="trilinear", align_corners=False) (monai Resized), then flips.  On the device: one reduction kernel
// leaves the box in 6 ints (in PADDED coordinates), one kernel resamples straight from the uint8 cube — ToTensor's /255 is
// applied to the 8 corner voxels in registers — with ATen's source-index rule and blend order
// (aten/src/ATen/native/cuda/UpSampleTrilinear3d.cu: src = max(scale * (dst + 0.5) - 0.5, 0), scale = in / out in fp32).
__global__ void __launch_bounds__(256) fg_bbox_kernel(const uint8_t* __restrict__ src, int T_src, int H, int W, int f0, int f1,
                                                      int shift, int* __restrict__ box) {
  // box = {t0, t1, h0, h1, w0, w1} (ends exclusive), pre-set by the host launch to {INT_MAX, -1, ...}
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
  const int W4 = W >> 2;
  const int64_t n = (int64_t)(f1 - f0) * H * W4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned q = (unsigned)i / (unsigned)W4;
    const int w4 = (int)((unsigned)i - q * (unsigned)W4);
    const int t = (int)(q / (unsigned)H), h = (int)(q - (unsigned)t * (unsigned)H);
    const uchar4 v = *reinterpret_cast<const uchar4*>(src + ((int64_t)(f0 + t) * H + h) * W + 4 * w4);
    if (v.x | v.y | v.z | v.w) {
      const int tp = f0 + t + shift;
      lo[0] = min(lo[0], tp); hi[0] = max(hi[0], tp);
      lo[1] = min(lo[1], h); hi[1] = max(hi[1], h);
      const int wl = 4 * w4 + (v.x ? 0 : v.y ? 1 : v.z ? 2 : 3), wh = 4 * w4 + (v.w ? 3 : v.z ? 2 : v.y ? 1 : 0);
      lo[2] = min(lo[2], wl); hi[2] = max(hi[2], wh);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0 && hi[a] >= 0) {
      atomicMin(box + 2 * a, lo[a]);
      atomicMax(box + 2 * a + 1, hi[a] + 1);
    }
  }
}

__global__ void fg_bbox_init_kernel(int* box) {
  if (threadIdx.x < 3) { box[2 * threadIdx.x] = 0x7fffffff; box[2 * threadIdx.x + 1] = -1; }
}

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& step, float& l1) {
  const float r = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = (int)r;
  step = (i0 < in_size - 1) ? 1 : 0;
  l1 = r - (float)i0;
}

__global__ void __launch_bounds__(256) resize_trilinear_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                                  const int* __restrict__ box, int T_src, int Hs, int Ws, int shift,
                                                                  int T_pad, int T, int H, int W, int flip_t, int flip_w,
                                                                  float divisor) {
  // region of the padded cube that is resampled: the foreground box, or the whole padded cube (box == nullptr; also when the
  // cube has no foreground at all)
  int t0 = 0, t1 = T_pad, h0 = 0, h1 = Hs, w0 = 0, w1 = Ws;
  if (box && box[1] > 0) { t0 = box[0]; t1 = box[1]; h0 = box[2]; h1 = box[3]; w0 = box[4]; w1 = box[5]; }
  const int it = t1 - t0, ih = h1 - h0, iw = w1 - w0;
  const float st = (float)it / (float)T, sh = (float)ih / (float)H, sw = (float)iw / (float)W;
  const int64_t n = (int64_t)T * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned q = (unsigned)i / (unsigned)W;
    const int w = (int)((unsigned)i - q * (unsigned)W);
    const int t = (int)(q / (unsigned)H), h = (int)(q - (unsigned)t * (unsigned)H);
    int ta, tp, ha, hp, wa, wp;
    float tl, hl, wl;
    src_index(st, flip_t ? T - 1 - t : t, it, ta, tp, tl);
    src_index(sh, h, ih, ha, hp, hl);
    src_index(sw, flip_w ? W - 1 - w : w, iw, wa, wp, wl);
    float v[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ts = t0 + ta + a * tp - shift;   // frame of the SOURCE cube (padding frames are zero)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float x = 0.f;
          if (ts >= 0 && ts < T_src) x = __fdiv_rn((float)src[((int64_t)ts * Hs + (h0 + ha + b * hp)) * Ws + (w0 + wa + c * wp)], divisor);
          v[a][b][c] = x;
        }
    }
    const float t0l = 1.f - tl, h0l = 1.f - hl, w0l = 1.f - wl;
    dst[i] = t0l * (h0l * (w0l * v[0][0][0] + wl * v[0][0][1]) + hl * (w0l * v[0][1][0] + wl * v[0][1][1])) +
             tl * (h0l * (w0l * v[1][0][0] + wl * v[1][0][1]) + hl * (w0l * v[1][1][0] + wl * v[1][1][1]));
  }
}

}  // namespace

extern "C" int oct_ingest_u8(const uint8_t* src, float* dst, const uint8_t* flip_t, const uint8_t* flip_w, int64_t B,
                             int64_t T_src, int64_t T, int64_t H, int64_t W, float divisor, oct_stream_t stream) {
  OCT_REQUIRE(src && dst, "oct_ingest_u8: null pointer");
  OCT_REQUIRE(B >= 0 && T_src > 0 && T > 0 && H > 0 && W > 0 && W % 4 == 0, "oct_ingest_u8: need positive sizes and W%%4==0");
  OCT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 3) == 0 && aligned16(dst), "oct_ingest_u8: misaligned");
  OCT_REQUIRE(divisor > 0.f, "oct_ingest_u8: divisor must be positive");
  OCT_REQUIRE(B <= 65535 && T * H * (W / 4) < (1ll << 31), "oct_ingest_u8: too large");
  if (B == 0) return OCT_OK;
  // PatientDataset_inhouse.py:439-450: left_padding = (T - T_src) // 2 zero frames, or frames [left_idx, left_idx + T)
  const int shift = T_src <= T ? (int)((T - T_src) / 2) : -(int)((T_src - T) / 2);
  const int64_t per_sample = T * H * (W / 4);
  int64_t blocks = ceil_div64(per_sample, 256 * 4);  // four float4 per thread
  const int64_t cap = (int64_t)oct_num_sms() * 16;
  if (blocks * B > cap) blocks = ceil_div64(cap, B);
  if (blocks < 1) blocks = 1;
  ingest_u8_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(src, dst, flip_t, flip_w, (int)T_src,
                                                                                          (int)T, (int)H, (int)W, shift, divisor);
  return oct_check_launch("oct_ingest_u8");
}

static inline int pad_shift(int64_t T_src, int64_t T_pad) {
  return T_src <= T_pad ? (int)((T_pad - T_src) / 2) : -(int)((T_src - T_pad) / 2);
}

extern "C" int oct_fg_bbox_u8(const uint8_t* src, int* box, int64_t T_src, int64_t H, int64_t W, int64_t T_pad, oct_stream_t stream) {
  OCT_REQUIRE(src && box, "oct_fg_bbox_u8: null pointer");
  OCT_REQUIRE(T_src > 0 && T_pad > 0 && H > 0 && W > 0 && W % 4 == 0, "oct_fg_bbox_u8: need positive sizes and W%%4==0");
  OCT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 3) == 0, "oct_fg_bbox_u8: misaligned");
  OCT_REQUIRE(T_src * H * (W / 4) < (1ll << 31), "oct_fg_bbox_u8: too large");
  const int shift = pad_shift(T_src, T_pad);
  const int f0 = shift < 0 ? -shift : 0;                      // frames of the source that survive the centre crop
  const int f1 = shift < 0 ? f0 + (int)T_pad : (int)T_src;
  cudaStream_t st = (cudaStream_t)stream;
  fg_bbox_init_kernel<<<1, 32, 0, st>>>(box);
  int64_t blocks = ceil_div64((int64_t)(f1 - f0) * H * (W / 4), 256 * 8);
  const int64_t cap = (int64_t)oct_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  fg_bbox_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, (int)T_src, (int)H, (int)W, f0, f1, shift, box);
  return oct_check_launch("oct_fg_bbox_u8");
}

extern "C" int oct_resize_trilinear_u8(const uint8_t* src, float* dst, const int* box, int64_t T_src, int64_t H_src, int64_t W_src,
                                       int64_t T_pad, int64_t T, int64_t H, int64_t W, int flip_t, int flip_w, float divisor,
                                       oct_stream_t stream) {
  OCT_REQUIRE(src && dst, "oct_resize_trilinear_u8: null pointer");
  OCT_REQUIRE(T_src > 0 && H_src > 0 && W_src > 0 && T_pad > 0 && T > 0 && H > 0 && W > 0, "oct_resize_trilinear_u8: need positive sizes");
  OCT_REQUIRE(divisor > 0.f, "oct_resize_trilinear_u8: divisor must be positive");
  OCT_REQUIRE(T * H * W < (1ll << 31), "oct_resize_trilinear_u8: too large");
  int64_t blocks = ceil_div64(T * H * W, 256 * 4);
  const int64_t cap = (int64_t)oct_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  resize_trilinear_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, box, (int)T_src, (int)H_src, (int)W_src,
                                                                                 pad_shift(T_src, T_pad), (int)T_pad, (int)T, (int)H,
                                                                                 (int)W, flip_t, flip_w, divisor);
  return oct_check_launch("oct_resize_trilinear_u8");
}
