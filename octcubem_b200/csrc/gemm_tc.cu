// bf16 GEMM on 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in TMEM)
// -> tcgen05.ld epilogue (bias / GELU / dGELU / fp32 accumulate) -> global.  Replaces the cuBLASLt calls behind
// nn.Linear forward (flash_attn/modules/mha.py:635,703, mlp.py:48-50, models_mae_joint_res_flash_attn.py:511,595)
// and their autograd dgrad / wgrad.
//
// Persistent, warp-specialised CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator),
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Two TMEM accumulator stages let the epilogue of tile i
// overlap the main loop of tile i+1.
//
// Operand layouts (all row-major in global memory):
//   "K-major"  operand: [rows, K] with K contiguous     -> one TMA box {64 K-elements, rows}, UMMA major = K
//   "MN-major" operand: [K, rows] with rows contiguous  -> rows/64 TMA boxes {64 rows-elements, 64 k}, UMMA major = MN
//   NT: A K-major,  B K-major      NN: A K-major, B MN-major      TN: A MN-major, B MN-major
#include "tc_common.cuh"
#include <mutex>

// ------------------------------------------------------------------------------------------------
// host: driver entry point + tensor map helper
// ------------------------------------------------------------------------------------------------
oct_encode_tiled_fn oct_get_encode_tiled() {
  static oct_encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (oct_encode_tiled_fn)p;
  });
  if (!fn) oct_set_error("cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

int oct_make_tmap(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, const char* who, CUtensorMapSwizzle swizzle) {
  oct_encode_tiled_fn enc = oct_get_encode_tiled();
  if (!enc) return OCT_ERR_UNSUPPORTED;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    oct_set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d; base %p dims %llu,%llu stride %llu box %u,%u)", who,
                  (int)r, base, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    return OCT_ERR_INVALID;
  }
  return OCT_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle line
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kGroupM = 8;    // tile rasterisation: groups of 8 m-blocks sweep all n-blocks (L2 reuse of both operands)

template <int BLOCK_N>
struct Cfg {
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages (power of two: 256 or 512)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int64_t ldd;
  void* D;
  int d_bf16;
  int epilogue;
  const float* bias;
  void* aux;
  int beta;
  int splits;        // split-K factor (fp32 D, EPI_NONE only): partial products are reduced with red.global.add
  int kb_per_split;  // k-blocks per split
};

template <bool A_MN, bool B_MN, int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                              const __grid_constant__ CUtensorMap tmap_b,
                                                              const GemmParams p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ provenance (LDS/STS instead of generic LD/ST): offset the array, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M, num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_work = num_tiles * p.splits;  // work item w: tile = w % num_tiles, split = w / num_tiles

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    for (int s = 0; s < C::kStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tmem_full[s], 1);
      tc::mbar_init(&tmem_empty[s], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<C::kTmemCols>(tmem_base_slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
    const int per_group = kGroupM * num_n;
    const int g = t / per_group, first_m = g * kGroupM;
    const int gsz = min(num_m - first_m, kGroupM);
    m_blk = first_m + (t % per_group) % gsz;
    n_blk = (t % per_group) / gsz;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(w % num_tiles, m_blk, n_blk);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
        const int kb0 = (w / num_tiles) * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          uint8_t* sa = smem_a + stage * C::kABytes;
          uint8_t* sb = smem_b + stage * C::kBBytes;
          const int k0 = kb * BLOCK_K;
          if (!A_MN) {
            tc::tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tc::tma_load_2d(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + c * 64, k0);
          }
          if (!B_MN) {
            tc::tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_N / 64; ++c)
              tc::tma_load_2d(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + c * 64, k0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = tc::make_idesc(tc::kFmtBF16, A_MN, B_MN, BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      if (lane == 0) {
        tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        const int num_kb = min(num_kb_total, (w / num_tiles + 1) * p.kb_per_split) - (w / num_tiles) * p.kb_per_split;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tcgen05_fence_after();
          const uint32_t a_addr = tc::smem_u32(smem_a + stage * C::kABytes);
          const uint32_t b_addr = tc::smem_u32(smem_b + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: advance 16 elements = 32 bytes inside the swizzled line; MN-major: advance 16 k-rows = 2048 B
            const uint64_t da = A_MN ? tc::make_smem_desc(a_addr + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : tc::make_smem_desc(a_addr + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? tc::make_smem_desc(b_addr + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                     : tc::make_smem_desc(b_addr + k * (UMMA_K * 2), 16, 1024);
            tc::mma_ss(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          tc::mma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(w % num_tiles, m_blk, n_blk);
      const int row = m_blk * BLOCK_M + quarter * 32 + lane;
      const int n0 = n_blk * BLOCK_N;
      tc::mbar_wait(&tmem_full[acc], acc_phase);
      tc::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        const int nc = n0 + ch * 32;
        if (nc >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tc::tmem_ld_x32(taddr + ch * 32, r);
        tc::tmem_ld_wait();
        if (row < p.M) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          const int nvalid = min(32, p.N - nc);  // multiple of 8 (N % 8 == 0)
          if (p.epilogue == OCT_EPI_BIAS || p.epilogue == OCT_EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i < nvalid) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc + i));
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            }
          }
          const size_t off = (size_t)row * p.ldd + nc;
          if (p.d_bf16) {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + off;
            __nv_bfloat16* ax = reinterpret_cast<__nv_bfloat16*>(p.aux) + off;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (i < nvalid) {
                if (p.epilogue == OCT_EPI_BIAS_GELU) {
                  uint4 pre;
                  pre.x = pack_bf16x2(v[i], v[i + 1]); pre.y = pack_bf16x2(v[i + 2], v[i + 3]);
                  pre.z = pack_bf16x2(v[i + 4], v[i + 5]); pre.w = pack_bf16x2(v[i + 6], v[i + 7]);
                  *reinterpret_cast<uint4*>(ax + i) = pre;
                  // GELU is evaluated on the bf16-rounded pre-activation, like nn.GELU on a bf16 tensor
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[i + j] = gelu_erf(bf16_round(v[i + j]));
                } else if (p.epilogue == OCT_EPI_DGELU) {
                  const uint4 pre = *reinterpret_cast<const uint4*>(ax + i);
                  const float2 a0 = unpack_bf16x2(pre.x), a1 = unpack_bf16x2(pre.y), a2 = unpack_bf16x2(pre.z),
                               a3 = unpack_bf16x2(pre.w);
                  v[i] *= gelu_erf_grad(a0.x); v[i + 1] *= gelu_erf_grad(a0.y);
                  v[i + 2] *= gelu_erf_grad(a1.x); v[i + 3] *= gelu_erf_grad(a1.y);
                  v[i + 4] *= gelu_erf_grad(a2.x); v[i + 5] *= gelu_erf_grad(a2.y);
                  v[i + 6] *= gelu_erf_grad(a3.x); v[i + 7] *= gelu_erf_grad(a3.y);
                }
                uint4 o;
                o.x = pack_bf16x2(v[i], v[i + 1]); o.y = pack_bf16x2(v[i + 2], v[i + 3]);
                o.z = pack_bf16x2(v[i + 4], v[i + 5]); o.w = pack_bf16x2(v[i + 6], v[i + 7]);
                *reinterpret_cast<uint4*>(d + i) = o;
              }
            }
          } else {
            float* d = reinterpret_cast<float*>(p.D) + off;
            float* ax = reinterpret_cast<float*>(p.aux) + off;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i < nvalid) {
                float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                if (p.epilogue == OCT_EPI_BIAS_GELU) {
                  *reinterpret_cast<float4*>(ax + i) = o;
                  o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
                } else if (p.epilogue == OCT_EPI_DGELU) {
                  const float4 a = *reinterpret_cast<const float4*>(ax + i);
                  o.x *= gelu_erf_grad(a.x); o.y *= gelu_erf_grad(a.y); o.z *= gelu_erf_grad(a.z); o.w *= gelu_erf_grad(a.w);
                } else if (p.splits > 1) {  // split-K partial: reduce in L2 (D was zeroed, or holds beta*D)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + i), "f"(o.x), "f"(o.y),
                               "f"(o.z), "f"(o.w) : "memory");
                  continue;
                } else if (p.beta) {
                  const float4 a = *reinterpret_cast<const float4*>(d + i);
                  o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                }
                *reinterpret_cast<float4*>(d + i) = o;
              }
            }
          }
        }
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <bool A_MN, bool B_MN, int BLOCK_N>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  auto kern = gemm_tc_kernel<A_MN, B_MN, BLOCK_N>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { oct_set_error("oct_gemm(bf16): smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  const int num_tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * ((p.N + BLOCK_N - 1) / BLOCK_N);
  const int num_work = num_tiles * p.splits;
  const int grid = num_work < oct_num_sms() ? num_work : oct_num_sms();
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, p);
  return oct_check_launch("oct_gemm(bf16)");
}

// operand map: K-major [rows, K] (ld = elements per row) or MN-major [K, rows]
int make_operand_map(CUtensorMap* map, const void* base, bool mn_major, int64_t rows, int64_t K, int64_t ld, int block_rows,
                     const char* who) {
  uint64_t dims[2], strides[1];
  uint32_t box[2];
  if (!mn_major) {
    dims[0] = (uint64_t)K; dims[1] = (uint64_t)rows; strides[0] = (uint64_t)ld * 2;
    box[0] = BLOCK_K; box[1] = (uint32_t)block_rows;
  } else {
    dims[0] = (uint64_t)rows; dims[1] = (uint64_t)K; strides[0] = (uint64_t)ld * 2;
    box[0] = 64; box[1] = BLOCK_K;
  }
  return oct_make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, who);
}

}  // namespace

int oct_gemm_tc_bf16(int layout, const void* A, const void* B, void* D, int d_dtype, int64_t M, int64_t N, int64_t K,
                     int64_t lda, int64_t ldb, int64_t ldd, int epilogue, const float* bias, void* aux, int beta,
                     cudaStream_t st) {
  OCT_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D), "oct_gemm(bf16): operands must be 16-byte aligned");
  OCT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "oct_gemm(bf16): lda/ldb must be multiples of 8 (TMA 16-byte strides)");
  OCT_REQUIRE(N % 8 == 0 && ldd % 8 == 0, "oct_gemm(bf16): N and ldd must be multiples of 8");
  OCT_REQUIRE(M < (1 << 30) && N < (1 << 30) && K < (1 << 30), "oct_gemm(bf16): dimension too large");
  OCT_REQUIRE(!aux || aligned16(aux), "oct_gemm(bf16): aux must be 16-byte aligned");
  OCT_REQUIRE(!bias || aligned16(bias), "oct_gemm(bf16): bias must be 16-byte aligned");
  if (M == 0 || N == 0) return OCT_OK;
  OCT_REQUIRE(K > 0, "oct_gemm(bf16): K must be positive");
  const bool a_mn = (layout == OCT_GEMM_TN), b_mn = (layout != OCT_GEMM_NT);
  const int block_n = (N > 128) ? 256 : 128;
  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, A, a_mn, M, K, lda, BLOCK_M, "oct_gemm(bf16) A");
  if (rc) return rc;
  rc = make_operand_map(&tb, B, b_mn, N, K, ldb, block_n, "oct_gemm(bf16) B");
  if (rc) return rc;
  GemmParams p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.ldd = ldd; p.D = D; p.d_bf16 = (d_dtype == OCT_BF16);
  p.epilogue = epilogue; p.bias = bias; p.aux = aux; p.beta = beta;
  // split-K: wgrad-shaped problems (few output tiles, long contraction) would otherwise occupy a fraction of the chip
  p.splits = 1;
  const int num_kb = (int)ceil_div64(K, BLOCK_K);
  p.kb_per_split = num_kb;
  if (d_dtype == OCT_F32 && epilogue == OCT_EPI_NONE) {
    const int64_t tiles = ceil_div64(M, BLOCK_M) * ceil_div64(N, block_n);
    const int sms = oct_num_sms();
    if (tiles * 2 <= sms && num_kb >= 16) {
      int splits = (int)(sms / tiles);
      if (splits > num_kb / 8) splits = num_kb / 8;  // keep >= 8 k-blocks (512 of K) per split
      if (splits > 1) {
        p.kb_per_split = (num_kb + splits - 1) / splits;
        p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
        if (!beta) {
          cudaError_t e = cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, st);
          if (e != cudaSuccess) { oct_set_error("oct_gemm(bf16): memset: %s", cudaGetErrorString(e)); return (int)e; }
        }
      }
    }
  }
#define GO(AMN, BMN)                                                                    \
  return block_n == 256 ? launch<AMN, BMN, 256>(ta, tb, p, st) : launch<AMN, BMN, 128>(ta, tb, p, st)
  switch (layout) {
    case OCT_GEMM_NT: GO(false, false);
    case OCT_GEMM_NN: GO(false, true);
    case OCT_GEMM_TN: GO(true, true);
    default: oct_set_error("oct_gemm: bad layout %d", layout); return OCT_ERR_INVALID;
  }
#undef GO
}
