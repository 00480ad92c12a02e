// Fused multi-tensor AdamW step: replaces torch.optim._multi_tensor.AdamW at
// Pre-training/main_pretrain_oph_joint_2d512_flash_attn.py:451-455 (betas 0.9 / 0.95, decoupled weight decay, no amsgrad)
// plus the GradScaler unscale that precedes it (custom_util/misc.py:326-344) and the bf16 weight cast that autocast repeats at
// the top of the next forward.  One launch per parameter group; HBM-bound: 16 B read + 12 B (+2 B shadow) written per parameter.
//
// Work is described by a device table of chunks (<= 16384 elements each, never crossing a tensor):
//   row = { param f32*, grad f32*, exp_avg f32*, exp_avg_sq f32*, shadow bf16* (or 0), count }
// Arithmetic order follows torch/optim/adamw.py (_multi_tensor): p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g g ;
// denom = sqrt(v) / sqrt(bias_correction2) + eps ; p -= (lr / bias_correction1) * m / denom.
#include "common.cuh"

namespace {

struct AdamWArgs {
  float lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2_sqrt, grad_scale;
  const float* grad_scale_dev;  // optional device scalar multiplied into grad_scale (the clip coefficient of oct_grad_norm)
};

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, const AdamWArgs& a) {
  g *= a.grad_scale;
  p *= 1.f - a.lr * a.weight_decay;
  m = m + (g - m) * (1.f - a.beta1);            // lerp(m, g, 1 - beta1), as torch._foreach_lerp_
  v = v * a.beta2 + (g * g) * (1.f - a.beta2);  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = sqrtf(v) / a.bias_corr2_sqrt + a.eps;
  p -= (a.lr / a.bias_corr1) * (m / denom);
}

// Device-resident optimizer clock (16 bytes): a captured CUDA graph replays the SAME launch parameters every step, so the
// quantities that change per step — the step count behind the bias corrections and the per-iteration cosine learning rate of
// custom_util/lr_sched.py:10-28 — live on the device and are advanced by a one-thread kernel inside the graph.
struct AdamWClock {
  int step;               // optimizer steps taken (1-based after the first advance)
  float lr;               // learning rate of this step (before a group's lr_scale)
  float bias_corr1;       // 1 - beta1^step
  float bias_corr2_sqrt;  // sqrt(1 - beta2^step)
};

__global__ void adamw_clock_advance_kernel(AdamWClock* clk, float base_lr, float min_lr, float warmup_epochs, float epochs,
                                           float epochs_per_step, float beta1, float beta2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int step = clk->step + 1;
  // lr_sched.adjust_learning_rate at the fractional epoch of the accumulation group's first iteration
  const double e = (double)(step - 1) * (double)epochs_per_step;
  double lr;
  if (e < (double)warmup_epochs) lr = (double)base_lr * e / (double)warmup_epochs;
  else lr = (double)min_lr + ((double)base_lr - (double)min_lr) * 0.5 *
                             (1.0 + cos(3.14159265358979323846 * (e - (double)warmup_epochs) / ((double)epochs - (double)warmup_epochs)));
  clk->step = step;
  clk->lr = (float)lr;
  clk->bias_corr1 = (float)(1.0 - pow((double)beta1, (double)step));
  clk->bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const int64_t* __restrict__ table, int n_chunks, AdamWArgs a,
                                                          const AdamWClock* __restrict__ clk) {
  if (clk) {  // a.lr carries the group's lr_scale
    a.lr *= clk->lr;
    a.bias_corr1 = clk->bias_corr1;
    a.bias_corr2_sqrt = clk->bias_corr2_sqrt;
  }
  if (a.grad_scale_dev) a.grad_scale *= __ldg(a.grad_scale_dev);
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int64_t* row = table + 6 * (int64_t)c;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* m = reinterpret_cast<float*>(row[2]);
    float* v = reinterpret_cast<float*>(row[3]);
    __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(row[4]);
    const int n = (int)row[5], n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 p4 = *reinterpret_cast<float4*>(p + 4 * i);
      const float4 g4 = *reinterpret_cast<const float4*>(g + 4 * i);
      float4 m4 = *reinterpret_cast<float4*>(m + 4 * i), v4 = *reinterpret_cast<float4*>(v + 4 * i);
      adamw1(p4.x, g4.x, m4.x, v4.x, a);
      adamw1(p4.y, g4.y, m4.y, v4.y, a);
      adamw1(p4.z, g4.z, m4.z, v4.z, a);
      adamw1(p4.w, g4.w, m4.w, v4.w, a);
      *reinterpret_cast<float4*>(p + 4 * i) = p4;
      *reinterpret_cast<float4*>(m + 4 * i) = m4;
      *reinterpret_cast<float4*>(v + 4 * i) = v4;
      if (sh) Vec4<__nv_bfloat16>::st(sh + 4 * i, p4);
    }
    if (threadIdx.x < (n & 3)) {  // tail of a tensor whose size is not a multiple of 4
      const int i = (n4 << 2) + threadIdx.x;
      float pp = p[i], mm = m[i], vv = v[i];
      adamw1(pp, g[i], mm, vv, a);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (sh) sh[i] = __float2bfloat16_rn(pp);
    }
  }
}

// ---- global gradient norm + clip coefficient (custom_util/misc.py:356-373 get_grad_norm_ / clip_grad_norm_) ----------
// partial[c] = sum of squares of chunk c (fixed order inside the chunk); one block then adds the partials in order:
// out[0] = grad_scale * sqrt(sum), out[1] = max_norm > 0 ? min(1, max_norm / (out[0] + 1e-6)) : 1  (torch's clip coefficient)
__global__ void __launch_bounds__(256) grad_sqnorm_partial_kernel(const int64_t* __restrict__ table, int n_chunks,
                                                                  float* __restrict__ partial) {
  __shared__ float sm[8];
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const float* g = reinterpret_cast<const float*>(table[6 * (int64_t)c + 1]);
    const int n = (int)table[6 * (int64_t)c + 5], n4 = n >> 2;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 v = *reinterpret_cast<const float4*>(g + 4 * i);
      acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    if (threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; acc += v * v; }
    const float t = block_sum<8>(acc, sm);
    if (threadIdx.x == 0) partial[c] = t;
  }
}
__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int n_chunks, float grad_scale,
                                                                float max_norm, float* __restrict__ out) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_chunks; i += 1024) acc += partial[i];
  const float t = block_sum<32>(acc, sm);
  if (threadIdx.x == 0) {
    const float norm = grad_scale * sqrtf(t);
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  }
}

}  // namespace

extern "C" int oct_grad_norm(const int64_t* table, int64_t n_chunks, float grad_scale, float max_norm, float* partial,
                             float* out, oct_stream_t stream) {
  OCT_REQUIRE((table && partial) || n_chunks == 0, "oct_grad_norm: null table / workspace");
  OCT_REQUIRE(out, "oct_grad_norm: null output");
  OCT_REQUIRE(n_chunks >= 0 && n_chunks < (1 << 30), "oct_grad_norm: bad chunk count");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_chunks > 0) {
    const int64_t cap = (int64_t)oct_num_sms() * 8;
    grad_sqnorm_partial_kernel<<<(unsigned)(n_chunks < cap ? n_chunks : cap), 256, 0, st>>>(table, (int)n_chunks, partial);
    int rc = oct_check_launch("oct_grad_norm(partial)");
    if (rc) return rc;
  }
  grad_norm_finish_kernel<<<1, 1024, 0, st>>>(partial, (int)n_chunks, grad_scale, max_norm, out);
  return oct_check_launch("oct_grad_norm(finish)");
}

extern "C" int oct_adamw_step(const int64_t* table, int64_t n_chunks, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int64_t step, float grad_scale, const float* grad_scale_dev,
                              oct_stream_t stream) {
  OCT_REQUIRE(table || n_chunks == 0, "oct_adamw_step: null table");
  OCT_REQUIRE(n_chunks >= 0 && n_chunks < (1 << 30), "oct_adamw_step: bad chunk count");
  OCT_REQUIRE(step >= 1, "oct_adamw_step: step counts from 1");
  OCT_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "oct_adamw_step: bad hyper-parameters");
  if (n_chunks == 0) return OCT_OK;
  AdamWArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_scale = grad_scale;
  a.grad_scale_dev = grad_scale_dev;
  a.bias_corr1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const int64_t cap = (int64_t)oct_num_sms() * 8;
  adamw_multi_kernel<<<(unsigned)(n_chunks < cap ? n_chunks : cap), 256, 0, (cudaStream_t)stream>>>(table, (int)n_chunks, a,
                                                                                                  nullptr);
  return oct_check_launch("oct_adamw_step");
}

extern "C" int oct_adamw_clock_advance(void* clock, float base_lr, float min_lr, float warmup_epochs, float epochs,
                                       float epochs_per_step, float beta1, float beta2, oct_stream_t stream) {
  OCT_REQUIRE(clock && aligned16(clock), "oct_adamw_clock_advance: clock must be a 16-byte aligned device buffer");
  OCT_REQUIRE(epochs > warmup_epochs && warmup_epochs >= 0.f && epochs_per_step >= 0.f, "oct_adamw_clock_advance: bad schedule");
  OCT_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "oct_adamw_clock_advance: bad betas");
  adamw_clock_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((AdamWClock*)clock, base_lr, min_lr, warmup_epochs, epochs,
                                                                 epochs_per_step, beta1, beta2);
  return oct_check_launch("oct_adamw_clock_advance");
}

extern "C" int oct_adamw_step_clocked(const int64_t* table, int64_t n_chunks, const void* clock, float lr_scale, float beta1,
                                      float beta2, float eps, float weight_decay, float grad_scale, const float* grad_scale_dev,
                                      oct_stream_t stream) {
  OCT_REQUIRE(table || n_chunks == 0, "oct_adamw_step_clocked: null table");
  OCT_REQUIRE(clock, "oct_adamw_step_clocked: null clock");
  OCT_REQUIRE(n_chunks >= 0 && n_chunks < (1 << 30), "oct_adamw_step_clocked: bad chunk count");
  OCT_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "oct_adamw_step_clocked: bad hyper-parameters");
  if (n_chunks == 0) return OCT_OK;
  AdamWArgs a;
  a.lr = lr_scale; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_scale = grad_scale;
  a.grad_scale_dev = grad_scale_dev;
  a.bias_corr1 = 1.f; a.bias_corr2_sqrt = 1.f;  // read from the clock
  const int64_t cap = (int64_t)oct_num_sms() * 8;
  adamw_multi_kernel<<<(unsigned)(n_chunks < cap ? n_chunks : cap), 256, 0, (cudaStream_t)stream>>>(table, (int)n_chunks, a,
                                                                                                  (const AdamWClock*)clock);
  return oct_check_launch("oct_adamw_step_clocked");
}
