// Backward of the tcgen05 flash attention (autograd of flash_attn_qkvpacked_func, flash_attn/modules/mha.py:122-130).
//
// One CTA owns one 128-row K/V tile of one (batch, head) and sweeps the query tiles.  Everything is computed in the
// TRANSPOSED orientation so that the softmax threads own kv rows (= TMEM lanes) and P^T / dS^T can be fed back to the
// tensor core as TMEM A-operands without ever leaving TMEM:
//     S^T  = K Q^T              (SS, K-major x K-major)          P^T  = exp2(S^T c - lse)
//     dP^T = V dO^T             (SS)                             dS^T = P^T o (dP^T - delta)
//     dV  += P^T  dO            (TS, dO tile as MN-major B)      dK  += dS^T Q     (TS, Q tile as MN-major B)
//     dQ_m = dS K               (SS, dS^T staged in smem and read as an MN-major A operand, K tile as MN-major B)
// dV / dK accumulate in TMEM across the sweep; dQ_m tiles are reduced across CTAs with vector fp32 reductions into a
// workspace that a small kernel scales and converts to bf16.
//
// 320 threads: warp 0 TMA, warp 1 MMA issuer, warps 2-9 softmax-backward.  Warps w and w+4 share a TMEM lane quarter and
// split the 128 query columns of a tile in halves, so every scheduler has two warps feeding the SFU (the kernel is
// ex2-bound for head_dim 32).  For head_dim 32 the bf16 P^T / dS^T live in their own TMEM columns, which lets the
// issuer queue S^T/dP^T of the next query tile ahead of the dV/dK/dQ MMAs of the current one; for head_dim 64 TMEM is
// too small for that (512 columns) and P^T / dS^T overwrite S^T / dP^T in place.
#include "tc_common.cuh"

namespace {

constexpr int AB_T = 128, AB_THREADS = 320, AB_Q_STAGES = 2, AB_SM_THREADS = 256;
constexpr float kLog2e = 1.4426950408889634f;

template <int HD>
struct AbCfg {
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kTileBytes = 128 * kRowBytes;
  static constexpr uint32_t kSwz = (HD == 64) ? tc::kSwz128 : tc::kSwz64;
  static constexpr uint32_t kSBO = 8 * kRowBytes;
  static constexpr int kDsBytes = 2 * 128 * 128;  // dS^T staging: two 64-column chunks of [128 kv rows x 128 B]
  // smem: K, V | Q[2], dO[2] | dS | lse2[2][128], delta[2][128] | barriers
  static constexpr int kSmem = 2 * kTileBytes + 2 * AB_Q_STAGES * kTileBytes + kDsBytes + 4 * 128 * 4 + 1024 + 256;
  static constexpr bool kInPlace = (HD == 64);
  // TMEM columns
  static constexpr uint32_t kColST = 0, kColDPT = 128;
  static constexpr uint32_t kColPT = kInPlace ? kColST : 256;    // bf16 P^T  (64 columns)
  static constexpr uint32_t kColDST = kInPlace ? kColDPT : 320;  // bf16 dS^T (64 columns)
  static constexpr uint32_t kColAcc = kInPlace ? 256 : 384;
  static constexpr uint32_t kColDV = kColAcc, kColDK = kColAcc + HD, kColDQ = kColAcc + 2 * HD;
  static_assert(kColDQ + HD <= 512, "TMEM budget");
  // TMEM column offset (inside the P^T / dS^T region) of the 16-query slice k16 (= one UMMA K step, 8 columns).
  // In place, each half-warpgroup keeps its bf16 output inside the fp32 columns it has itself consumed
  // (half 0: columns [0,32), half 1: [64,96)), so the two halves never overwrite each other's unread scores.
  __host__ __device__ static constexpr uint32_t slice_off(int k16) {
    return kInPlace ? (uint32_t)((k16 >> 2) * 64 + (k16 & 3) * 8) : (uint32_t)(k16 * 8);
  }
};

struct AbParams {
  int S, H, Spad;
  float scale, scale_log2e;
  const float* lse;     // [B,H,S]
  const float* delta;   // [B,H,S]
  float* dq_acc;        // [B,H,Spad,HD] fp32, zero-initialised
  __nv_bfloat16* dqkv;  // [B,S,3,H,HD]
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                    const __grid_constant__ CUtensorMap tmap_do,
                                                                    const AbParams p) {
  using C = AbCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ provenance (LDS/STS instead of generic LD/ST): offset the array, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + C::kTileBytes;
  uint8_t* sQ = sV + C::kTileBytes;                       // [AB_Q_STAGES]
  uint8_t* sDO = sQ + AB_Q_STAGES * C::kTileBytes;        // [AB_Q_STAGES]
  uint8_t* sDS = sDO + AB_Q_STAGES * C::kTileBytes;       // 32 KB, 1024-aligned (all tiles are multiples of 8 KB)
  float* sLse = reinterpret_cast<float*>(sDS + C::kDsBytes);  // [2][128]
  float* sDelta = sLse + 2 * 128;                             // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 2 * 128);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;                 // [2]
  uint64_t* q_empty = q_full + AB_Q_STAGES;    // [2]
  uint64_t* sdp_full = q_empty + AB_Q_STAGES;
  uint64_t* p_ready = sdp_full + 1;
  uint64_t* dq_full = p_ready + 1;
  uint64_t* dq_free = dq_full + 1;
  uint64_t* acc_full = dq_free + 1;
  uint64_t* p_free = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.S + AB_T - 1) / AB_T;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_qkv);
    tc::prefetch_tmap(&tmap_do);
    tc::mbar_init(kv_full, 1);
    for (int s = 0; s < AB_Q_STAGES; ++s) { tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1); }
    tc::mbar_init(sdp_full, 1);
    tc::mbar_init(p_ready, AB_SM_THREADS);
    tc::mbar_init(dq_full, 1);
    tc::mbar_init(dq_free, AB_SM_THREADS);
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(p_free, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(kv_full, 2 * C::kTileBytes);
      tc::tma_load_4d(sK, &tmap_qkv, kv_full, 0, p.H + h, n0, b);
      tc::tma_load_4d(sV, &tmap_qkv, kv_full, 0, 2 * p.H + h, n0, b);
      int stage = 0; uint32_t phase = 0;
      for (int m = 0; m < n_q; ++m) {
        tc::mbar_wait(&q_empty[stage], phase ^ 1);
        tc::mbar_arrive_expect_tx(&q_full[stage], 2 * C::kTileBytes);
        tc::tma_load_4d(sQ + stage * C::kTileBytes, &tmap_qkv, &q_full[stage], 0, h, m * AB_T, b);
        tc::tma_load_4d(sDO + stage * C::kTileBytes, &tmap_do, &q_full[stage], 0, h, m * AB_T, b);
        if (++stage == AB_Q_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_st = tc::make_idesc(tc::kFmtBF16, false, false, 128, 128);   // K Q^T, V dO^T
      constexpr uint32_t idesc_acc = tc::make_idesc(tc::kFmtBF16, false, true, 128, HD);     // P^T dO, dS^T Q (TS)
      constexpr uint32_t idesc_dq = tc::make_idesc(tc::kFmtBF16, true, true, 128, HD);       // dS K (A MN-major)
      const uint32_t k_addr = tc::smem_u32(sK), v_addr = tc::smem_u32(sV), ds_addr = tc::smem_u32(sDS);
      auto issue_sdp = [&](int stage) {  // S^T = K Q^T ; dP^T = V dO^T
        const uint32_t q_addr = tc::smem_u32(sQ + stage * C::kTileBytes);
        const uint32_t do_addr = tc::smem_u32(sDO + stage * C::kTileBytes);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t da = tc::make_smem_desc(k_addr + k * 32, 16, C::kSBO, C::kSwz);
          const uint64_t db = tc::make_smem_desc(q_addr + k * 32, 16, C::kSBO, C::kSwz);
          tc::mma_ss(tmem_base + C::kColST, da, db, idesc_st, k != 0);
        }
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t da = tc::make_smem_desc(v_addr + k * 32, 16, C::kSBO, C::kSwz);
          const uint64_t db = tc::make_smem_desc(do_addr + k * 32, 16, C::kSBO, C::kSwz);
          tc::mma_ss(tmem_base + C::kColDPT, da, db, idesc_st, k != 0);
        }
        tc::mma_commit(sdp_full);
      };
      tc::mbar_wait(kv_full, 0);
      tc::mbar_wait(&q_full[0], 0);
      tc::tcgen05_fence_after();
      issue_sdp(0);
      int stage = 0;
      int nstage = 1 % AB_Q_STAGES; uint32_t nphase = (AB_Q_STAGES == 1) ? 1 : 0;
      for (int m = 0; m < n_q; ++m) {
        const bool more = (m + 1) < n_q;
        const uint32_t q_addr = tc::smem_u32(sQ + stage * C::kTileBytes);
        const uint32_t do_addr = tc::smem_u32(sDO + stage * C::kTileBytes);
        auto issue_dv_dk = [&]() {
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k) {  // dV += P^T dO
            const uint64_t db = tc::make_smem_desc(do_addr + k * 16 * C::kRowBytes, C::kTileBytes, C::kSBO, C::kSwz);
            tc::mma_ts(tmem_base + C::kColDV, tmem_base + C::kColPT + C::slice_off(k), db, idesc_acc, (m | k) != 0);
          }
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k) {  // dK += dS^T Q
            const uint64_t db = tc::make_smem_desc(q_addr + k * 16 * C::kRowBytes, C::kTileBytes, C::kSBO, C::kSwz);
            tc::mma_ts(tmem_base + C::kColDK, tmem_base + C::kColDST + C::slice_off(k), db, idesc_acc, (m | k) != 0);
          }
        };
        auto issue_dq = [&]() {
          if (m > 0) {
            tc::mbar_wait(dq_free, (m - 1) & 1);  // previous dQ tile drained from TMEM
            tc::tcgen05_fence_after();
          }
#pragma unroll
          for (int k = 0; k < 128 / 16; ++k) {  // dQ_m = dS K : A = dS^T smem tile read MN-major (M = q contiguous)
            const uint64_t da = tc::make_smem_desc(ds_addr + k * 16 * 128, 128 * 128, 1024, tc::kSwz128);
            const uint64_t db = tc::make_smem_desc(k_addr + k * 16 * C::kRowBytes, C::kTileBytes, C::kSBO, C::kSwz);
            tc::mma_ss(tmem_base + C::kColDQ, da, db, idesc_dq, k != 0);
          }
          tc::mma_commit(dq_full);
        };
        auto issue_next_sdp = [&]() {
          if (more) {
            tc::mbar_wait(&q_full[nstage], nphase);
            tc::tcgen05_fence_after();
            issue_sdp(nstage);
          }
        };
        tc::mbar_wait(p_ready, m & 1);  // softmax(m) done: fp32 S^T/dP^T consumed, bf16 P^T/dS^T (+ smem dS) ready
        tc::tcgen05_fence_after();
        if (!C::kInPlace) {
          // dQ first (its drain is on the threads' critical path), then the next tile's scores, then dV/dK, whose
          // completion (p_free) is only needed when the threads want to overwrite the bf16 P^T/dS^T columns again
          issue_dq();
          issue_next_sdp();
          issue_dv_dk();
        } else {
          issue_dv_dk();
          issue_dq();
          issue_next_sdp();  // overwrites P^T/dS^T in place: must follow their consumers in the in-order MMA pipe
        }
        tc::mma_commit(p_free);
        tc::mma_commit(&q_empty[stage]);
        if (more && ++nstage == AB_Q_STAGES) { nstage = 0; nphase ^= 1; }
        if (++stage == AB_Q_STAGES) stage = 0;
      }
      tc::mma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== softmax-backward threads: thread <-> (kv row, half of the query columns) ==============
    const int half = (warp - 2) >> 2;     // 0: query columns 0..63, 1: 64..127
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // kv row inside the tile == TMEM lane
    const int tid = threadIdx.x - 64;     // 0..255
    const bool kv_ok = (n0 + row) < p.S;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const size_t bh = (size_t)b * p.H + h;
    // per-query statistics of a q tile: threads 0..127 fetch lse (-> log2 domain), 128..255 fetch delta; the fetch for
    // tile m+1 is issued one iteration ahead so its latency hides behind the softmax work of tile m
    auto load_stat = [&](int m) -> float {  // raw value; transformed only when it is stored (keeps the LDG in flight)
      const int qi = min(m * AB_T + (tid & 127), p.S - 1);
      return (tid < 128) ? p.lse[bh * p.S + qi] : p.delta[bh * p.S + qi];
    };
    float stat = load_stat(0);
    for (int m = 0; m < n_q; ++m) {
      const int slot = m & 1;
      {
        const bool ok = (m * AB_T + (tid & 127)) < p.S;
        if (tid < 128) sLse[slot * 128 + tid] = ok ? stat * kLog2e : INFINITY;
        else sDelta[slot * 128 + (tid & 127)] = ok ? stat : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (m + 1 < n_q) stat = load_stat(m + 1);
      tc::mbar_wait(sdp_full, m & 1);
      tc::tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;  // 32-column chunk of the tile
        uint32_t s[32], dp[32];
        tc::tmem_ld_x32(lane_addr + C::kColST + c * 32, s);
        tc::tmem_ld_x32(lane_addr + C::kColDPT + c * 32, dp);
        tc::tmem_ld_wait();
        uint32_t pk[16], dk[16];
        const float4* l4 = reinterpret_cast<const float4*>(sLse + slot * 128 + c * 32);
        const float4* d4 = reinterpret_cast<const float4*>(sDelta + slot * 128 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 lv = l4[i], dv = d4[i];
          float p0 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * i]), p.scale_log2e, -lv.x));
          float p1 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * i + 1]), p.scale_log2e, -lv.y));
          float p2 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * i + 2]), p.scale_log2e, -lv.z));
          float p3 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * i + 3]), p.scale_log2e, -lv.w));
          if (!kv_ok) { p0 = 0.f; p1 = 0.f; p2 = 0.f; p3 = 0.f; }
          const float d0 = p0 * (__uint_as_float(dp[4 * i]) - dv.x);
          const float d1 = p1 * (__uint_as_float(dp[4 * i + 1]) - dv.y);
          const float d2 = p2 * (__uint_as_float(dp[4 * i + 2]) - dv.z);
          const float d3 = p3 * (__uint_as_float(dp[4 * i + 3]) - dv.w);
          pk[2 * i] = pack_bf16x2(p0, p1);
          pk[2 * i + 1] = pack_bf16x2(p2, p3);
          dk[2 * i] = pack_bf16x2(d0, d1);
          dk[2 * i + 1] = pack_bf16x2(d2, d3);
        }
        if (cc == 0 && m > 0) {  // dV/dK of the previous tile have finished reading the bf16 P^T / dS^T columns
          tc::mbar_wait(p_free, (m - 1) & 1);
          tc::tcgen05_fence_after();
        }
        tc::tmem_st_x16(lane_addr + C::kColPT + C::slice_off(2 * c), pk);    // P^T  (bf16 pairs), 2 K-slices
        tc::tmem_st_x16(lane_addr + C::kColDST + C::slice_off(2 * c), dk);   // dS^T (bf16 pairs)
        // dS^T row -> smem (MN-major A operand of dQ = dS K): 64-column chunk = half, 16-byte pieces cc*4 .. +3
        uint8_t* rowp = sDS + half * (128 * 128) + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int piece = cc * 4 + i;
          uint4 v = make_uint4(dk[4 * i], dk[4 * i + 1], dk[4 * i + 2], dk[4 * i + 3]);
          *reinterpret_cast<uint4*>(rowp + ((piece ^ (row & 7)) << 4)) = v;
        }
      }
      tc::tmem_st_wait();
      tc::fence_proxy_async();  // st.shared (generic proxy) -> tcgen05.mma reads (async proxy)
      tc::tcgen05_fence_before();
      tc::mbar_arrive(p_ready);
      // drain dQ_m and reduce it into the fp32 accumulator (TMEM lane = query row here; the two halves split the columns)
      tc::mbar_wait(dq_full, m & 1);
      tc::tcgen05_fence_after();
      float* dq_row = p.dq_acc + (bh * p.Spad + (size_t)m * AB_T + row) * HD + half * (HD / 2);
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[16];
        tc::tmem_ld_x16(lane_addr + C::kColDQ + half * (HD / 2) + c * 16, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(dq_row + c * 16 + i, __uint_as_float(o[i]), __uint_as_float(o[i + 1]), __uint_as_float(o[i + 2]),
                     __uint_as_float(o[i + 3]));
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(dq_free);
    }
    // epilogue: half 0 stores dV, half 1 stores dK of this kv tile
    tc::mbar_wait(acc_full, 0);
    tc::tcgen05_fence_after();
    const int kv = n0 + row;
    const uint32_t col = half ? C::kColDK : C::kColDV;
    const float sc = half ? p.scale : 1.f;
    __nv_bfloat16* dst = p.dqkv + ((((size_t)b * p.S + kv) * 3 + (half ? 1 : 2)) * p.H + h) * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[32];
      tc::tmem_ld_x32(lane_addr + col + c * 32, o);
      tc::tmem_ld_wait();
      if (kv_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[i]) * sc, __uint_as_float(o[i + 1]) * sc);
          v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * sc, __uint_as_float(o[i + 3]) * sc);
          v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * sc, __uint_as_float(o[i + 5]) * sc);
          v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * sc, __uint_as_float(o[i + 7]) * sc);
          *reinterpret_cast<uint4*>(dst + c * 32 + i) = v;
        }
      }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// delta[b,h,s] = sum_c dO[b,s,h,c] * O[b,s,h,c]   (one warp per (b,s,h))
template <int HD>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int64_t total, int S, int H) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const int64_t base = w * HD;
  float acc = 0.f;
  if (lane * 2 < HD) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(out + base + lane * 2));
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + base + lane * 2));
    acc = a.x * d.x + a.y * d.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const int hh = (int)(w % H);
    const int64_t bs = w / H;
    const int s = (int)(bs % S);
    const int64_t bb = bs / S;
    delta[(bb * H + hh) * S + s] = acc;
  }
}

// dq (bf16, into dqkv[:, :, 0]) = scale * dq_acc
template <int HD>
__global__ void attn_dq_convert_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, int S, int H,
                                       int Spad, float scale, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*S*H*HD/4
  if (i >= total4) return;
  const int c4 = (int)(i % (HD / 4));
  int64_t r = i / (HD / 4);
  const int hh = (int)(r % H); r /= H;
  const int s = (int)(r % S);
  const int64_t bb = r / S;
  const float4 v = *reinterpret_cast<const float4*>(dq_acc + ((bb * H + hh) * Spad + s) * HD + c4 * 4);
  __nv_bfloat16* dst = dqkv + (((bb * S + s) * 3 + 0) * H + hh) * HD + c4 * 4;
  Vec4<__nv_bfloat16>::st(dst, make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale));
}

int make_map4(CUtensorMap* map, const void* base, int64_t d, int64_t hdim, int64_t S, int64_t B, const char* who) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)hdim, (uint64_t)S, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)hdim * d * 2, (uint64_t)S * hdim * d * 2};
  uint32_t box[4] = {(uint32_t)d, 1, 128, 1};
  return oct_make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box, who,
                       d == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int HD>
int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* ws, int64_t B,
               int64_t S, int64_t H, float scale, cudaStream_t st) {
  using C = AbCfg<HD>;
  const int64_t Spad = ceil_div64(S, 128) * 128;
  float* dq_acc = (float*)ws;
  float* delta = dq_acc + (size_t)B * H * Spad * HD;
  CUtensorMap mq, md;
  int rc = make_map4(&mq, qkv, HD, 3 * H, S, B, "oct_attn_bwd(bf16) qkv");
  if (rc) return rc;
  rc = make_map4(&md, dout, HD, H, S, B, "oct_attn_bwd(bf16) dout");
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)B * H * Spad * HD * sizeof(float), st);
  if (e != cudaSuccess) { oct_set_error("oct_attn_bwd(bf16): memset: %s", cudaGetErrorString(e)); return (int)e; }
  const int64_t rows = B * S * H;
  attn_delta_kernel<HD><<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>((const __nv_bfloat16*)out,
                                                                              (const __nv_bfloat16*)dout, delta, rows,
                                                                              (int)S, (int)H);
  rc = oct_check_launch("oct_attn_bwd(bf16,delta)");
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    e = cudaFuncSetAttribute(attn_bwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) { oct_set_error("oct_attn_bwd(bf16): smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  AbParams p;
  p.S = (int)S; p.H = (int)H; p.Spad = (int)Spad; p.scale = scale; p.scale_log2e = scale * kLog2e;
  p.lse = lse; p.delta = delta; p.dq_acc = dq_acc; p.dqkv = (__nv_bfloat16*)dqkv;
  dim3 grid((unsigned)ceil_div64(S, AB_T), (unsigned)H, (unsigned)B);
  attn_bwd_tc_kernel<HD><<<grid, AB_THREADS, C::kSmem, st>>>(mq, md, p);
  rc = oct_check_launch("oct_attn_bwd(bf16)");
  if (rc) return rc;
  const int64_t total4 = B * S * H * (HD / 4);
  attn_dq_convert_kernel<HD><<<(unsigned)ceil_div64(total4, 256), 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dqkv, (int)S,
                                                                                (int)H, (int)Spad, scale, total4);
  return oct_check_launch("oct_attn_bwd(bf16,dq)");
}

}  // namespace

size_t oct_attn_bwd_tc_ws_bytes(int64_t B, int64_t S, int64_t H, int64_t d) {
  const int64_t Spad = ceil_div64(S, 128) * 128;
  return (size_t)B * H * Spad * d * sizeof(float) + (size_t)B * H * S * sizeof(float) + 64;
}

int oct_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* ws,
                    size_t ws_bytes, int64_t B, int64_t S, int64_t H, int64_t d, float scale, cudaStream_t st) {
  OCT_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv) && aligned16(ws),
              "oct_attn_bwd(bf16): pointers must be 16-byte aligned");
  (void)ws_bytes;
  if (d == 64) return launch_bwd<64>(qkv, out, dout, lse, dqkv, ws, B, S, H, scale, st);
  if (d == 32) return launch_bwd<32>(qkv, out, dout, lse, dqkv, ws, B, S, H, scale, st);
  oct_set_error("oct_attn_bwd(bf16): head dim %lld unsupported by the tcgen05 kernel (32 or 64)", (long long)d);
  return OCT_ERR_UNSUPPORTED;
}
