// PatchEmbed.forward (custom_util/video_vit.py:74-83): Conv3d(1 -> E, kernel = stride = (u,16,16)) + flatten +
// 'ncts->ntsc', as an im2col-free GEMM.  The fp32 volume is never re-laid out: a 5-D TMA box (kw, kh, w, h, frame)
// drops 128 tokens x 16 k-elements (one patch row: 64 contiguous bytes, hence the 64B swizzle mode — the TMA row
// pitch in smem equals the inner box extent) into a K-major smem tile, tcgen05.mma kind::tf32 contracts it
// against the fp32 Conv3d weight viewed as [E, u*16*16] (vv:69-72 weight order == patchify order), the
// accumulator lives in TMEM, and the epilogue adds the bias and writes token-major [B, T'*h*w, E] directly.
//
// One CTA = 128 tokens (8 patch rows x 16 patch columns of one temporal slot) x 256 output channels.
// 2 CTAs/SM co-reside (96 KB smem, 256 TMEM columns each) so one CTA's epilogue overlaps the other's main loop.
#include "tc_common.cuh"

namespace {

constexpr int PE_BLOCK_M = 128, PE_BLOCK_N = 256, PE_BLOCK_K = 16 /* fp32 elements = 64 B */, PE_UMMA_K = 8;
constexpr int PE_STAGES = 4;
constexpr int PE_ROW_BYTES = PE_BLOCK_K * 4;
constexpr int PE_A_BYTES = PE_BLOCK_M * PE_ROW_BYTES, PE_B_BYTES = PE_BLOCK_N * PE_ROW_BYTES;
constexpr int PE_STAGE_BYTES = PE_A_BYTES + PE_B_BYTES;
constexpr int PE_SMEM = PE_STAGES * PE_STAGE_BYTES + 1024 + 128;
constexpr int PE_THREADS = 192;

struct PeParams {
  int T, u, hp, wp, E, Tp;   // frames per volume, temporal patch, patch grid, channels, T/u
  int h_tiles, w_tiles;      // ceil(hp/8), ceil(wp/16)
  const float* bias;
  void* out;
  int out_bf16;
};

__global__ void __launch_bounds__(PE_THREADS, 2) patch_embed_tc_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                       const __grid_constant__ CUtensorMap tmap_w,
                                                                       const PeParams p) {
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ provenance (LDS/STS instead of generic LD/ST): offset the array, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + PE_STAGES * PE_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PE_STAGES * PE_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + PE_STAGES;
  uint64_t* acc_bar = empty_bar + PE_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x -> (b, t', h_tile, w_tile) ; blockIdx.y -> n tile
  int m_tile = blockIdx.x;
  const int wt = m_tile % p.w_tiles; m_tile /= p.w_tiles;
  const int ht = m_tile % p.h_tiles; m_tile /= p.h_tiles;
  const int tp = m_tile % p.Tp;
  const int b = m_tile / p.Tp;
  const int n0 = blockIdx.y * PE_BLOCK_N;
  const int num_kb = p.u * 16;  // one (kt, kh) patch row of 16 pixels per k-block

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_x);
    tc::prefetch_tmap(&tmap_w);
    for (int s = 0; s < PE_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(acc_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<PE_BLOCK_N>(tmem_slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[stage], PE_STAGE_BYTES);
        const int kt = kb >> 4, kh0 = kb & 15;
        tc::tma_load_5d(smem_a + stage * PE_A_BYTES, &tmap_x, &full_bar[stage], 0, kh0, wt * 16, ht * 8,
                        b * p.T + tp * p.u + kt);
        tc::tma_load_2d(smem_b + stage * PE_B_BYTES, &tmap_w, &full_bar[stage], kb * PE_BLOCK_K, n0);
        if (++stage == PE_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(tc::kFmtTF32, false, false, PE_BLOCK_M, PE_BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::tcgen05_fence_after();
        const uint32_t a_addr = tc::smem_u32(smem_a + stage * PE_A_BYTES);
        const uint32_t b_addr = tc::smem_u32(smem_b + stage * PE_B_BYTES);
#pragma unroll
        for (int k = 0; k < PE_BLOCK_K / PE_UMMA_K; ++k) {
          const uint64_t da = tc::make_smem_desc(a_addr + k * 32, 16, 8 * PE_ROW_BYTES, tc::kSwz64);
          const uint64_t db = tc::make_smem_desc(b_addr + k * 32, 16, 8 * PE_ROW_BYTES, tc::kSwz64);
          tc::mma_ss_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        tc::mma_commit(&empty_bar[stage]);
        if (++stage == PE_STAGES) { stage = 0; phase ^= 1; }
      }
      tc::mma_commit(acc_bar);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;          // tile row = hl*16 + wl
    const int h = ht * 8 + (r >> 4), w = wt * 16 + (r & 15);
    const bool row_ok = (h < p.hp) && (w < p.wp);
    const size_t token = ((size_t)b * p.Tp + tp) * p.hp * p.wp + (size_t)h * p.wp + w;
    tc::mbar_wait(acc_bar, 0);
    tc::tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int ch = 0; ch < PE_BLOCK_N / 32; ++ch) {
      const int nc = n0 + ch * 32;
      if (nc >= p.E) break;
      uint32_t rr[32];
      tc::tmem_ld_x32(taddr + ch * 32, rr);
      tc::tmem_ld_wait();
      if (row_ok) {
        const int nvalid = min(32, p.E - nc);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < nvalid) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc + i));
          v[i] = __uint_as_float(rr[i]) + b4.x; v[i + 1] = __uint_as_float(rr[i + 1]) + b4.y;
          v[i + 2] = __uint_as_float(rr[i + 2]) + b4.z; v[i + 3] = __uint_as_float(rr[i + 3]) + b4.w;
        }
        if (p.out_bf16) {
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.out) + token * p.E + nc;
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            if (i < nvalid) {
              uint4 o;
              o.x = pack_bf16x2(v[i], v[i + 1]); o.y = pack_bf16x2(v[i + 2], v[i + 3]);
              o.z = pack_bf16x2(v[i + 4], v[i + 5]); o.w = pack_bf16x2(v[i + 6], v[i + 7]);
              *reinterpret_cast<uint4*>(d + i) = o;
            }
        } else {
          float* d = reinterpret_cast<float*>(p.out) + token * p.E + nc;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (i < nvalid) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<PE_BLOCK_N>(tmem_base);
  }
}

}  // namespace

extern "C" int oct_patch_embed_fwd(const float* imgs, const float* weight, const float* bias, void* out, int out_dtype,
                                   int64_t B, int64_t T, int64_t H, int64_t W, int64_t p, int64_t u, int64_t E,
                                   oct_stream_t stream) {
  OCT_REQUIRE(imgs && weight && bias && out, "oct_patch_embed_fwd: null pointer");
  OCT_REQUIRE(p == 16, "oct_patch_embed_fwd: patch size must be 16 (got %lld)", (long long)p);
  // the reference asserts the spatial size against the module's img_size (vv:76-78); divisibility is what we need here
  OCT_REQUIRE(H % 16 == 0 && W % 16 == 0 && u > 0 && T % u == 0, "oct_patch_embed_fwd: need H,W %% 16 == 0 and T %% u == 0");
  OCT_REQUIRE(E % 8 == 0, "oct_patch_embed_fwd: E must be a multiple of 8");
  OCT_REQUIRE(aligned16(imgs) && aligned16(weight) && aligned16(bias) && aligned16(out), "oct_patch_embed_fwd: misaligned");
  OCT_REQUIRE(out_dtype == OCT_F32 || out_dtype == OCT_BF16, "oct_patch_embed_fwd: bad dtype");
  if (B == 0) return OCT_OK;
  const int64_t hp = H / 16, wp = W / 16, K = u * 256;
  CUtensorMap tx, tw;
  {
    uint64_t dims[5] = {16, 16, (uint64_t)wp, (uint64_t)hp, (uint64_t)(B * T)};
    uint64_t str[4] = {(uint64_t)W * 4, 64, (uint64_t)16 * W * 4, (uint64_t)H * W * 4};
    uint32_t box[5] = {16, 1, 16, 8, 1};
    int rc = oct_make_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, imgs, dims, str, box, "oct_patch_embed_fwd(volume)",
                           CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)E};
    uint64_t str[1] = {(uint64_t)K * 4};
    uint32_t box[2] = {PE_BLOCK_K, PE_BLOCK_N};
    int rc = oct_make_tmap(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, weight, dims, str, box, "oct_patch_embed_fwd(weight)",
                           CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  PeParams prm;
  prm.T = (int)T; prm.u = (int)u; prm.hp = (int)hp; prm.wp = (int)wp; prm.E = (int)E; prm.Tp = (int)(T / u);
  prm.h_tiles = (int)ceil_div64(hp, 8); prm.w_tiles = (int)ceil_div64(wp, 16);
  prm.bias = bias; prm.out = out; prm.out_bf16 = (out_dtype == OCT_BF16);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM);
    if (e != cudaSuccess) { oct_set_error("oct_patch_embed_fwd: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  const int64_t m_tiles = B * prm.Tp * prm.h_tiles * prm.w_tiles;
  OCT_REQUIRE(m_tiles < (1ll << 31), "oct_patch_embed_fwd: too many tiles");
  dim3 grid((unsigned)m_tiles, (unsigned)ceil_div64(E, PE_BLOCK_N));
  patch_embed_tc_kernel<<<grid, PE_THREADS, PE_SMEM, (cudaStream_t)stream>>>(tx, tw, prm);
  return oct_check_launch("oct_patch_embed_fwd");
}
