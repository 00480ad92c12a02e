// placeholder until the tcgen05 flash kernels land
#include "tc_common.cuh"
int oct_attn_fwd_tc(const void*, void*, float*, int64_t, int64_t, int64_t, int64_t, float, cudaStream_t) {
  oct_set_error("oct_attn_fwd(bf16): tcgen05 kernel not built"); return OCT_ERR_UNSUPPORTED;
}
int oct_attn_bwd_tc(const void*, const void*, const void*, const float*, void*, void*, size_t, int64_t, int64_t, int64_t, int64_t, float, cudaStream_t) {
  oct_set_error("oct_attn_bwd(bf16): tcgen05 kernel not built"); return OCT_ERR_UNSUPPORTED;
}
size_t oct_attn_bwd_tc_ws_bytes(int64_t B, int64_t S, int64_t H, int64_t d) { return (size_t)B * H * S * sizeof(float); }
