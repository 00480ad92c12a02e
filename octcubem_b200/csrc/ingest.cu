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
