// bf16 GEMM on 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in TMEM)
// -> tcgen05.ld epilogue (bias / GELU / dGELU / fp32 accumulate) -> global.  Replaces the cuBLASLt calls behind
// nn.Linear forward (flash_attn/modules/mha.py:635,703, mlp.py:48-50, models_mae_joint_res_flash_attn.py:511,595)
// and their autograd dgrad / wgrad.
//
// Persistent, warp-specialised CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator),
// warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter split the columns).  Two TMEM accumulator stages let the epilogue of tile i
// overlap the main loop of tile i+1.
//
// Operand layouts (all row-major in global memory):
//   "K-major"  operand: [rows, K] with K contiguous     -> one TMA box {64 K-elements, rows}, UMMA major = K
//   "MN-major" operand: [K, rows] with rows contiguous  -> rows/64 TMA boxes {64 rows-elements, 64 k}, UMMA major = MN
//   NT: A K-major,  B K-major      NN: A K-major, B MN-major      TN: A MN-major, B MN-major
#include "tc_common.cuh"
#include <mutex>
#include <cstdlib>

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
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    // A thread that has not touched the CUDA runtime yet (e.g. a fresh autograd worker): bind the context of the device
    // that owns the operand, then retry.  Kernel launches do this implicitly, the driver-level encode call does not.
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, base) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
        cudaSetDevice(at.device) == cudaSuccess && cudaFree(nullptr) == cudaSuccess)
      r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    (void)cudaGetLastError();
  }
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
constexpr int kThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kEpiThreads = 256;
constexpr int kGroupM = 8;    // tile rasterisation: groups of 8 m-blocks sweep all n-blocks (L2 reuse of both operands)

constexpr int kEpiStageBytes = 128 * 128;  // one [128 rows x 128 B] output group per column half

constexpr int kOnesBytes = 16 * 128;  // [16 rows x 64 bf16] tile of ones: B operand of the bias-gradient MMA (kColsum)

// kColsum (wgrad + bias gradient in one kernel, TN layout only): next to D = A^T B the kernel accumulates
// colsum[m] = sum_k A[k, m] = (A^T 1)[m] with one extra 128x16x16 MMA per k-step against a constant tile of ones, into 16
// TMEM columns behind a SINGLE accumulator stage (2 x 256 + 16 columns do not fit the 512-column TMEM).  It replaces a
// separate pass over dY per nn.Linear (131 column-sum launches, 5.5 % of the step, profiles/r1_step_launches.md);
// wgrad-shaped problems map one work item to each CTA, so the second accumulator stage was idle there anyway.
template <int BLOCK_N, bool kColsum, bool kPair>
struct Cfg {
  // a CTA of a pair keeps only ITS half of the B tile (the 2-CTA MMA reads the other half from the peer's smem)
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = (kPair ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  // operand ring + 64 KB of output staging (+ 2 KB ones tile) within 227 KB
  static constexpr int kStages = kPair ? (BLOCK_N == 256 ? (kColsum ? 4 : 5) : 6) : ((BLOCK_N == 256) ? 3 : (kColsum ? 4 : 5));
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kAccStages = kColsum ? 1 : 2;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages, or one + the colsum columns (power of two: 256 or 512)
  static constexpr int kColsumCol = BLOCK_N;     // TMEM column of the bias-gradient accumulator (kColsum)
  static constexpr int kBiasBytes = BLOCK_N * 4;  // the tile's bias values, staged once per tile by the epilogue warps
  static constexpr int kSmemBytes = kStages * kStageBytes + 4 * kEpiStageBytes + (kColsum ? kOnesBytes : 0) + kBiasBytes +
                                    1024 /*align slack*/ + 256 /*barriers*/;
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
  int dbg;           // OCT_GEMM_DBG timing experiments (0 in production): 1 = epilogue skips its global stores
  int splits;        // split-K factor (fp32 D, EPI_NONE only): partial products are reduced with red.global.add
  int kb_per_split;  // k-blocks per split
  float* colsum;     // kColsum: [M] fp32, (+)= column sums of the MN-major A operand (the bias gradient of a wgrad)
};

// kPair: the grid is launched in clusters of two CTAs (one TPC) that own vertically adjacent 128-row tiles of the same
// 256-column strip and run ONE tcgen05.mma.cta_group::2 (M = 256) per k-step, issued by the even ("leader") CTA.  Each CTA
// stages its own 128 rows of A and only ITS HALF of the B tile (32 KB per k-block instead of 48 KB): with single-CTA
// MMAs the operand ring needs 96 B/clk of smem writes (TMA) plus 96 B/clk of reads (MMA) against a 128 B/clk port, which
// capped the main loop at ~65 % of the tensor rate (36.6 us for the 3280x1024x4096 fc2 tile, profiles/r1_step_launches.md).
// Protocol: both CTAs' TMA loads complete on the LEADER's full barrier (which the leader arms with the bytes of both);
// the leader's tcgen05.commit multicasts onto both CTAs' empty / tmem_full barriers; the epilogue warps of both CTAs
// release an accumulator stage on the leader's tmem_empty barrier (one arrival per warp).
#ifndef OCT_GEMM_TRACE  // build with -DOCT_GEMM_TRACE=1: clock64 stamps of the phases of a CTA, printed by CTAs 0 / 1 / last
#define OCT_GEMM_TRACE 0
#endif
#if OCT_GEMM_TRACE
#define G_TRACE(id) do { gtrace[id] = clock64(); } while (0)
#define G_TRACEW(id) do { if (lane == 0 && w == cta) gtrace[16 + (warp - 2) * 8 + (id)] = clock64(); } while (0)
#else
#define G_TRACE(id) do {} while (0)
#define G_TRACEW(id) do {} while (0)
#endif

template <bool A_MN, bool B_MN, int BLOCK_N, bool kPair, bool kColsum>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                              const __grid_constant__ CUtensorMap tmap_b,
                                                              const __grid_constant__ CUtensorMap tmap_d,
                                                              const __grid_constant__ CUtensorMap tmap_aux,
                                                              const GemmParams p) {
  using C = Cfg<BLOCK_N, kColsum, kPair>;
  static_assert(!kColsum || A_MN, "the bias-gradient MMA sums the MN-major A operand over K");
#if OCT_GEMM_TRACE
  __shared__ long long gtrace[16 + 8 * 8];
  if (threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); gtrace[15] = (long long)g; G_TRACE(0); }
#endif
  pdl_launch_dependents();
  static_assert(!kPair || BLOCK_N == 256 || (BLOCK_N == 128 && !kColsum), "pair mode: 256 x 256 (or 256 x 128) output tile per CTA pair");
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ provenance (LDS/STS instead of generic LD/ST): offset the array, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint8_t* smem_epi = smem + C::kStages * C::kStageBytes;  // 2 column halves x 2 buffers x 16 KB, 1024-aligned
  uint8_t* smem_ones = smem_epi + 4 * kEpiStageBytes;      // kColsum: 2 KB of bf16 1.0, 1024-aligned
  float* smem_bias = reinterpret_cast<float*>(smem_ones + (kColsum ? kOnesBytes : 0));  // [2 column halves][BLOCK_N / 2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_bias) + C::kBiasBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;  // [column half][buffer]: dGELU pre-activation tiles landed in the staging buffers
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(aux_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = kPair ? (int)tc::cluster_ctarank() : 0;
  const int cta = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // index of this CTA (pair) in the work loop
  const int ncta = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // tile rows are handed out in units of one m-block (two vertically adjacent m-blocks for a pair)
  const int num_m = kPair ? (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) : (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_work = num_tiles * p.splits;  // work item w: tile = w % num_tiles, split = w / num_tiles

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
    tc::prefetch_tmap(&tmap_d);
    for (int s = 0; s < C::kStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tmem_full[s], 1);
      tc::mbar_init(&tmem_empty[s], kPair ? 2 * (kEpiThreads / 32) : kEpiThreads);  // pair: one arrival per epilogue warp of both CTAs
    }
    for (int s = 0; s < 4; ++s) tc::mbar_init(&aux_full[s], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tc::tmem_alloc_pair<C::kTmemCols>(tmem_base_slot); else tc::tmem_alloc<C::kTmemCols>(tmem_base_slot);
  }
  if (kColsum) {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem_ones)[i] = 0x3F803F80u;
    tc::fence_proxy_async();  // generic-proxy stores -> tcgen05.mma operand reads
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (kPair) tc::cluster_sync_all();  // the peer's barriers are initialised before anything is multicast at them
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) G_TRACE(1);
  pdl_wait();  // everything above overlapped the previous kernel's tail; operands / outputs are touched only below

  auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
    const int per_group = kGroupM * num_n;
    const int g = t / per_group, first_m = g * kGroupM;
    const int gsz = min(num_m - first_m, kGroupM);
    m_blk = first_m + (t % per_group) % gsz;
    if (kPair) m_blk = 2 * m_blk + rank;  // may point one block past M for the odd tail: TMA zero-fills, stores are masked
    n_blk = (t % per_group) / gsz;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cta; w < num_work; w += ncta) {
        int m_blk, n_blk;
        tile_coords(w % num_tiles, m_blk, n_blk);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
        const int kb0 = (w / num_tiles) * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * C::kABytes;
          uint8_t* sb = smem_b + stage * C::kBBytes;
          const int k0 = kb * BLOCK_K;
          if (!kPair) {
            tc::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
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
          } else {
            // the leader arms its full barrier with the bytes of BOTH CTAs; every load of the pair completes there
            if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            if (!A_MN) {
              tc::tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)
                tc::tma_load_2d_pair(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + c * 64, k0);
            }
            const int nh = n0 + rank * (BLOCK_N / 2);  // my half of the B tile's N rows
            if (!B_MN) {
              tc::tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], k0, nh);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 128; ++c)
                tc::tma_load_2d_pair(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[stage], nh + c * 64, k0);
            }
          }
          if (kb == kb0 && w == cta) G_TRACE(2);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
      G_TRACE(3);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr int kMmaM = kPair ? 2 * BLOCK_M : BLOCK_M;  // cta_group::2: one instruction covers both CTAs' rows
    constexpr uint32_t idesc = tc::make_idesc(tc::kFmtBF16, A_MN, B_MN, kMmaM, BLOCK_N);
    constexpr uint32_t idesc_cs = tc::make_idesc(tc::kFmtBF16, A_MN, false, kMmaM, 16);  // A^T x ones[16 x K] (K-major)
    const uint64_t d_ones = tc::make_smem_desc(tc::smem_u32(smem_ones), 16, 1024);         // every k-slice is all ones
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = cta; w < num_work; w += ncta) {
      if (lane == 0 && rank == 0) {  // pair: the peer's warp 1 only allocates / frees TMEM
        bool with_colsum = false;
        if (kColsum) {  // the n-block-0 tile of every m-block row carries the column sums of its A tiles
          int m_blk, n_blk;
          tile_coords(w % num_tiles, m_blk, n_blk);
          with_colsum = (n_blk == 0);
        }
        tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        const int num_kb = min(num_kb_total, (w / num_tiles + 1) * p.kb_per_split) - (w / num_tiles) * p.kb_per_split;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          if (kb == 0 && w == cta) G_TRACE(4);
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
            if (kPair) {
              tc::mma_ss_pair(tmem_d, da, db, idesc, (kb | k) != 0);
              if (kColsum && with_colsum) tc::mma_ss_pair(tmem_base + C::kColsumCol, da, d_ones, idesc_cs, (kb | k) != 0);
            } else {
              tc::mma_ss(tmem_d, da, db, idesc, (kb | k) != 0);
              if (kColsum && with_colsum) tc::mma_ss(tmem_base + C::kColsumCol, da, d_ones, idesc_cs, (kb | k) != 0);
            }
          }
          // frees the smem slot once these MMAs have read it (pair: in both CTAs)
          if (kPair) tc::mma_commit_pair(&empty_bar[stage], 3); else tc::mma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (pair: of both CTAs)
        if (kPair) tc::mma_commit_pair(&tmem_full[acc], 3); else tc::mma_commit(&tmem_full[acc]);
        if (w == cta) G_TRACE(5);
      }
      __syncwarp();
      if (++acc == C::kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps share a TMEM lane quarter and split the tile's columns.  Results leave the SM through a swizzled smem
    // staging tile and ONE TMA store per 128-byte-wide column group: per-lane global stores (32 different rows per warp
    // instruction) made output-heavy GEMMs store-bound — 70 us with, 38 us without them at 32776x1536x512
    // (profiles/r1_gemm_ncu.md).  fp32 accumulate / split-K partials use the reducing store (cp.reduce ... .add).
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    const int colhalf = (warp - 2) >> 2;
    const int tile_row = quarter * 32 + lane;
    uint8_t* stg_base = smem_epi + colhalf * 2 * kEpiStageBytes;  // two buffers, used alternately
    int stg_sel = 0;
    const int bar_id = 1 + colhalf;
    const bool issuer = (quarter == 0) && (lane == 0);
    // hand an accumulator stage back to the MMA issuer: every thread (single CTA), or one arrival per warp on the LEADER's
    // barrier (pair; the peer arrives remotely)
    const uint32_t tmem_empty_leader[2] = {kPair ? tc::mapa_u32(&tmem_empty[0], 0) : 0u, kPair ? tc::mapa_u32(&tmem_empty[1], 0) : 0u};
    auto release_acc = [&](int a) {
      tc::tcgen05_fence_before();
      if (!kPair) {
        tc::mbar_arrive(&tmem_empty[a]);
      } else {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(tmem_empty_leader[a]);
      }
    };
    // 128-byte row segment (32 x 32-bit) -> staging tile -> TMA store of the [128 rows x 128 B] group at (col, row0)
    auto stage_and_store = [&](const uint32_t (&o)[32], const CUtensorMap* map, int col, int row0, bool reduce_add) {
      uint8_t* stg = stg_base + stg_sel * kEpiStageBytes;
      uint8_t* stg_row = stg + tile_row * 128;
      const uint32_t stg_addr = tc::smem_u32(stg);
      stg_sel ^= 1;
      // the store issued two calls ago (same buffer) has left smem; the most recent one may still be in flight
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(stg_row + ((q ^ (tile_row & 7)) << 4)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      tc::fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (issuer && !(p.dbg & 1)) {
        if (reduce_add)
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(map)), "r"(stg_addr), "r"(col), "r"(row0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(map)), "r"(stg_addr), "r"(col), "r"(row0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int w = cta; w < num_work; w += ncta) {
      int m_blk, n_blk;
      tile_coords(w % num_tiles, m_blk, n_blk);
      const int row0 = m_blk * BLOCK_M;
      const int n0 = n_blk * BLOCK_N;
      if (p.epilogue == OCT_EPI_DGELU && issuer) {
        // dGELU: fetch this tile's pre-activations straight into the staging buffers (TMA, same swizzle as the store)
        // while the MMAs of the tile are still running; the epilogue then works in place.  Per-lane global loads of
        // 32 different rows per instruction cost as much as the per-lane stores did (profiles/r1_gemm_ncu.md).
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // earlier stores have drained both buffers
        constexpr int kGroups = BLOCK_N / 128;
#pragma unroll
        for (int gg = 0; gg < kGroups; ++gg) {
          const int nc = n0 + (colhalf * kGroups + gg) * 64;
          if (nc < p.N) {
            tc::mbar_arrive_expect_tx(&aux_full[colhalf * 2 + gg], kEpiStageBytes);
            tc::tma_load_2d(stg_base + gg * kEpiStageBytes, &tmap_aux, &aux_full[colhalf * 2 + gg], nc, row0);
          }
        }
      }
      const bool has_bias = (p.epilogue == OCT_EPI_BIAS || p.epilogue == OCT_EPI_BIAS_GELU);
      float* sbias = smem_bias + colhalf * (BLOCK_N / 2);
      if (has_bias) {
        // this column half's bias values -> smem while the tile's MMAs are still running: the per-thread global loads of
        // the same 64 values sat on the epilogue's critical path (long_scoreboard 25 %, profiles/r1_gemm_ncu.md).  The
        // previous tile's readers are past the second bar.sync of their last staging call.
        const int bc = n0 + colhalf * (BLOCK_N / 2) + tile_row;
        if (tile_row < BLOCK_N / 2) sbias[tile_row] = (bc < p.N) ? __ldg(p.bias + bc) : 0.f;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      tc::mbar_wait(&tmem_full[acc], acc_phase);
      if (w == cta && warp == 2 && lane == 0) G_TRACE(6);
      G_TRACEW(0);
      tc::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
      if (kColsum && n_blk == 0 && colhalf == 0) {
        // bias gradient of rows [row0, row0 + 128): all 16 columns of the ones-MMA hold the same sum; lane <-> row
        const float cs = __uint_as_float(tc::tmem_ld_x1(tmem_base + ((uint32_t)(quarter * 32) << 16) + C::kColsumCol));
        if (row0 + tile_row < p.M) {
          if (p.beta != 0 || p.splits > 1) atomicAdd(p.colsum + row0 + tile_row, cs);
          else p.colsum[row0 + tile_row] = cs;
        }
      }
      if (p.d_bf16) {
        constexpr int kGroups = BLOCK_N / 128;  // 64-column (128-byte) groups per column half
#pragma unroll 1
        for (int gg = 0; gg < kGroups; ++gg) {
          const int g = colhalf * kGroups + gg;
          const int nc = n0 + g * 64;
          if (nc >= p.N) break;  // uniform over the 128 threads of this column half
          uint32_t r0[32], r1[32];
          tc::tmem_ld_x32(taddr + g * 64, r0);
          tc::tmem_ld_x32(taddr + g * 64 + 32, r1);
          tc::tmem_ld_wait();
          G_TRACEW(1 + 3 * gg);
          if (gg == kGroups - 1 || nc + 64 >= p.N) release_acc(acc);  // last TMEM read of this tile
          float v[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sbias + gg * 64 + i);  // broadcast read
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          }
          uint32_t o[32];
          if (p.epilogue == OCT_EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            stage_and_store(o, &tmap_aux, nc, row0, false);  // pre-activation
            // GELU is evaluated on the bf16-rounded pre-activation, like nn.GELU on a bf16 tensor
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float2 pre = unpack_bf16x2(o[i]);
              tc::gelu_fast2(pre.x, pre.y, v[2 * i], v[2 * i + 1]);
            }
          } else if (p.epilogue == OCT_EPI_DGELU) {
            uint8_t* stg = stg_base + gg * kEpiStageBytes;
            uint8_t* stg_row = stg + tile_row * 128;
            tc::mbar_wait(&aux_full[colhalf * 2 + gg], (aux_phase >> gg) & 1);
            aux_phase ^= 1u << gg;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint4* slot = reinterpret_cast<uint4*>(stg_row + ((q ^ (tile_row & 7)) << 4));
              const uint4 pre = *slot;
              const float2 a0 = unpack_bf16x2(pre.x), a1 = unpack_bf16x2(pre.y), a2 = unpack_bf16x2(pre.z),
                           a3 = unpack_bf16x2(pre.w);
              const int i = 8 * q;
              uint4 out;
              float g0, g1;
              tc::gelu_fast_grad2(a0.x, a0.y, v[i], v[i + 1], g0, g1);
              out.x = pack_bf16x2(g0, g1);
              tc::gelu_fast_grad2(a1.x, a1.y, v[i + 2], v[i + 3], g0, g1);
              out.y = pack_bf16x2(g0, g1);
              tc::gelu_fast_grad2(a2.x, a2.y, v[i + 4], v[i + 5], g0, g1);
              out.z = pack_bf16x2(g0, g1);
              tc::gelu_fast_grad2(a3.x, a3.y, v[i + 6], v[i + 7], g0, g1);
              out.w = pack_bf16x2(g0, g1);
              *slot = out;  // in place: each thread owns its 128-byte row segment
            }
            tc::fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (issuer && !(p.dbg & 1)) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmap_d)), "r"(tc::smem_u32(stg)), "r"(nc), "r"(row0) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            continue;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          G_TRACEW(2 + 3 * gg);
          stage_and_store(o, &tmap_d, nc, row0, false);
          G_TRACEW(3 + 3 * gg);
        }
      } else {
        constexpr int kGroups = BLOCK_N / 64;  // 32-column (128-byte) fp32 groups per column half
#pragma unroll 1
        for (int gg = 0; gg < kGroups; ++gg) {
          const int g = colhalf * kGroups + gg;
          const int nc = n0 + g * 32;
          if (nc >= p.N) break;
          uint32_t o[32];
          tc::tmem_ld_x32(taddr + g * 32, o);
          tc::tmem_ld_wait();
          if (gg == kGroups - 1 || nc + 32 >= p.N) release_acc(acc);
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              {
                const float4 b4 = *reinterpret_cast<const float4*>(sbias + gg * 32 + i);
                o[i] = __float_as_uint(__uint_as_float(o[i]) + b4.x);
                o[i + 1] = __float_as_uint(__uint_as_float(o[i + 1]) + b4.y);
                o[i + 2] = __float_as_uint(__uint_as_float(o[i + 2]) + b4.z);
                o[i + 3] = __float_as_uint(__uint_as_float(o[i + 3]) + b4.w);
              }
            }
          }
          // beta = 1 and split-K partial products accumulate in L2 through the reducing TMA store
          stage_and_store(o, &tmap_d, nc, row0, p.beta != 0 || p.splits > 1);
        }
      }
      // a tile whose column half lies entirely beyond N never reached an arrive above
      if (n0 + colhalf * (BLOCK_N / 2) >= p.N) release_acc(acc);
      if (w == cta && warp == 2 && lane == 0) G_TRACE(7);
      if (++acc == C::kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    // the staging buffers have been read; global visibility of the stores is the kernel boundary's business (waiting for it
    // here, `wait_group 0` without .read, kept every CTA ~1.3 us past its last store: clock64 trace)
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (warp == 2 && lane == 0) G_TRACE(8);
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (kPair) tc::cluster_sync_all_relaxed();  // nobody leaves while the peer may still multicast into / arrive on this CTA
  if (warp == 1) {
    tc::tcgen05_fence_after();
    if (kPair) tc::tmem_dealloc_pair<C::kTmemCols>(tmem_base); else tc::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
#if OCT_GEMM_TRACE
  __syncthreads();
  if (threadIdx.x == 0 && (p.dbg & 2) && (blockIdx.x < 2 || blockIdx.x == gridDim.x / 2 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    const long long t0 = gtrace[0];
    printf("GTRACE cta %d gt0 %lld gt1 %lld | pro %lld tma1 %lld tmaend %lld full1 %lld mmaend %lld epi0 %lld epi1 %lld stored %lld exit %lld\n", (int)blockIdx.x,
           gtrace[15], (long long)g, gtrace[1] - t0, gtrace[2] - t0, gtrace[3] - t0, gtrace[4] - t0, gtrace[5] - t0, gtrace[6] - t0,
           gtrace[7] - t0, gtrace[8] - t0, (long long)clock64() - t0);
    if (blockIdx.x == 0)
      for (int ww = 0; ww < 8; ++ww)
        printf("GTRACEW warp %d wake %lld | ld %lld packed %lld staged %lld | ld %lld packed %lld staged %lld\n", ww + 2, gtrace[16 + ww * 8] - t0,
               gtrace[16 + ww * 8 + 1] - t0, gtrace[16 + ww * 8 + 2] - t0, gtrace[16 + ww * 8 + 3] - t0, gtrace[16 + ww * 8 + 4] - t0,
               gtrace[16 + ww * 8 + 5] - t0, gtrace[16 + ww * 8 + 6] - t0);
  }
#endif
}

template <bool A_MN, bool B_MN, int BLOCK_N, bool kPair, bool kColsum = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tx, const GemmParams& p,
           cudaStream_t st) {
  using C = Cfg<BLOCK_N, kColsum, kPair>;
  auto kern = gemm_tc_kernel<A_MN, B_MN, BLOCK_N, kPair, kColsum>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { oct_set_error("oct_gemm(bf16): smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  const int unit_m = kPair ? 2 * BLOCK_M : BLOCK_M;
  const int num_tiles = ((p.M + unit_m - 1) / unit_m) * ((p.N + BLOCK_N - 1) / BLOCK_N);
  const int num_work = num_tiles * p.splits;
  {
    const int units = kPair ? oct_num_sms() / 2 : oct_num_sms();
    const int grid = (num_work < units ? num_work : units) * (kPair ? 2 : 1);
    cudaError_t e = oct_launch(kern, dim3(grid), dim3(kThreads), (size_t)C::kSmemBytes, st, kPair ? 2 : 1, ta, tb, td, tx, p);
    if (e != cudaSuccess) { oct_set_error("oct_gemm(bf16): launch: %s", cudaGetErrorString(e)); return (int)e; }
  }
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
                     float* colsum, cudaStream_t st) {
  OCT_REQUIRE(!colsum || (layout == OCT_GEMM_TN && d_dtype == OCT_F32 && epilogue == OCT_EPI_NONE),
              "oct_gemm(bf16): the fused column sum needs the TN layout, fp32 D and no epilogue");
  OCT_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D), "oct_gemm(bf16): operands must be 16-byte aligned");
  OCT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "oct_gemm(bf16): lda/ldb must be multiples of 8 (TMA 16-byte strides)");
  OCT_REQUIRE(N % 8 == 0 && ldd % 8 == 0, "oct_gemm(bf16): N and ldd must be multiples of 8");
  OCT_REQUIRE(M < (1 << 30) && N < (1 << 30) && K < (1 << 30), "oct_gemm(bf16): dimension too large");
  OCT_REQUIRE(!aux || aligned16(aux), "oct_gemm(bf16): aux must be 16-byte aligned");
  OCT_REQUIRE(!bias || aligned16(bias), "oct_gemm(bf16): bias must be 16-byte aligned");
  if (M == 0 || N == 0) return OCT_OK;
  OCT_REQUIRE(K > 0, "oct_gemm(bf16): K must be positive");
  const bool a_mn = (layout == OCT_GEMM_TN), b_mn = (layout != OCT_GEMM_NT);
  int block_n = (N > 128) ? 256 : 128;
  static const int force_bn = [] { const char* e = getenv("OCT_GEMM_BN"); return e ? atoi(e) : 0; }();  // experiment knob
  if (force_bn == 128 || force_bn == 256) block_n = force_bn;
  CUtensorMap ta, tb;
  int rc = make_operand_map(&ta, A, a_mn, M, K, lda, BLOCK_M, "oct_gemm(bf16) A");
  if (rc) return rc;
  // pair mode (2-CTA clusters running cta_group::2 MMAs on 256 x 256 tiles) whenever there are at least two m-blocks.
  // At K <= 512 (8 k-blocks per tile) the tile time is the epilogue's: the GELU / dGELU epilogues are slower in pair mode
  // (93 -> 97 us at the decoder fc1), the plain and bias ones faster since the remote accumulator release stopped
  // costing a cluster-scope fence (dec Wqkv 56.1 -> 49.2 us, out_proj 29.4 -> 26.6 us; tools/time_gemm_shapes.py)
  static const int pair_min_k = [] { const char* e = getenv("OCT_GEMM_PAIR_MIN_K"); return e ? atoi(e) : 512; }();
  const bool light_epilogue = (epilogue == OCT_EPI_NONE || epilogue == OCT_EPI_BIAS) && d_dtype == OCT_BF16;
  const bool pair = (M > BLOCK_M) && (block_n == 256 || (block_n == 128 && !colsum)) && (K > pair_min_k || light_epilogue) &&
                    (getenv("OCT_GEMM_NO_PAIR") == nullptr);
  rc = make_operand_map(&tb, B, b_mn, N, K, ldb, pair ? block_n / 2 : block_n, "oct_gemm(bf16) B");
  if (rc) return rc;
  // output maps for the TMA-store epilogue: [M, N] row-major, one box = 128 rows x 128 bytes
  OCT_REQUIRE(d_dtype == OCT_BF16 || (epilogue != OCT_EPI_BIAS_GELU && epilogue != OCT_EPI_DGELU),
              "oct_gemm(bf16): GELU epilogues need a bf16 D");
  CUtensorMap td, tx;
  {
    const bool f32 = (d_dtype == OCT_F32);
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldd * (f32 ? 4 : 2)};
    uint32_t box[2] = {f32 ? 32u : 64u, (uint32_t)BLOCK_M};
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    rc = oct_make_tmap(&td, dt, 2, D, dims, strides, box, "oct_gemm(bf16) D");
    if (rc) return rc;
    tx = td;
    if (epilogue == OCT_EPI_BIAS_GELU || epilogue == OCT_EPI_DGELU) {
      rc = oct_make_tmap(&tx, dt, 2, aux, dims, strides, box, "oct_gemm(bf16) aux");
      if (rc) return rc;
    }
  }
  GemmParams p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.ldd = ldd; p.D = D; p.d_bf16 = (d_dtype == OCT_BF16);
  p.epilogue = epilogue; p.bias = bias; p.aux = aux; p.beta = beta; p.colsum = colsum;
  { const char* e = getenv("OCT_GEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  // split-K: wgrad-shaped problems (few output tiles, long contraction) would otherwise occupy a fraction of the chip
  p.splits = 1;
  const int num_kb = (int)ceil_div64(K, BLOCK_K);
  p.kb_per_split = num_kb;
  if (d_dtype == OCT_F32 && epilogue == OCT_EPI_NONE && num_kb >= 16) {
    // Work items = tiles x splits are dealt round-robin to one CTA per SM: pick the split count that minimises
    // rounds x (k-blocks per item + the item's epilogue, ~6 k-blocks' worth for a 128x256 fp32 reducing store), keeping
    // >= 8 k-blocks (512 of K) per split.  E.g. the encoder Wqkv wgrad (96 tiles, 52 k-blocks) runs as 3 x 96 items in two
    // rounds of 18 instead of one round of 52 on 96 of the 148 SMs.
    const int64_t tiles = ceil_div64(M, pair ? 2 * BLOCK_M : BLOCK_M) * ceil_div64(N, block_n) * (pair ? 2 : 1);
    const int sms = oct_num_sms();
    const int kEpiCost = 6;
    int best = 1;
    int64_t best_cost = ceil_div64(tiles, sms) * (num_kb + kEpiCost);
    for (int s = 2; s <= num_kb / 8 && s <= 64; ++s) {
      const int per = (num_kb + s - 1) / s;
      const int eff = (num_kb + per - 1) / per;
      const int64_t cost = ceil_div64(tiles * eff, sms) * (per + kEpiCost);
      if (cost < best_cost) { best_cost = cost; best = s; }
    }
    if (best > 1) {
      p.kb_per_split = (num_kb + best - 1) / best;
      p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
      if (!beta) {
        cudaError_t e = cudaMemset2DAsync(D, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, st);
        if (e != cudaSuccess) { oct_set_error("oct_gemm(bf16): memset: %s", cudaGetErrorString(e)); return (int)e; }
      }
    }
  }
  if (colsum && p.splits > 1 && !beta) {
    cudaError_t e = cudaMemsetAsync(colsum, 0, (size_t)M * sizeof(float), st);
    if (e != cudaSuccess) { oct_set_error("oct_gemm(bf16): memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  if (colsum) {
    if (pair) return launch<true, true, 256, true, true>(ta, tb, td, tx, p, st);
    return block_n == 256 ? launch<true, true, 256, false, true>(ta, tb, td, tx, p, st) : launch<true, true, 128, false, true>(ta, tb, td, tx, p, st);
  }
#define GO(AMN, BMN)                                                                                         \
  if (pair) return block_n == 256 ? launch<AMN, BMN, 256, true>(ta, tb, td, tx, p, st) : launch<AMN, BMN, 128, true>(ta, tb, td, tx, p, st); \
  return block_n == 256 ? launch<AMN, BMN, 256, false>(ta, tb, td, tx, p, st) : launch<AMN, BMN, 128, false>(ta, tb, td, tx, p, st)
  switch (layout) {
    case OCT_GEMM_NT: GO(false, false);
    case OCT_GEMM_NN: GO(false, true);
    case OCT_GEMM_TN: GO(true, true);
    default: oct_set_error("oct_gemm: bad layout %d", layout); return OCT_ERR_INVALID;
  }
#undef GO
}
