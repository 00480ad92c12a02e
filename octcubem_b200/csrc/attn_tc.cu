// Flash-style self-attention on tcgen05 / TMEM (replaces flash_attn_qkvpacked_func, flash_attn/modules/mha.py:122-130:
// non-causal, dropout 0, softmax scale d^-0.5; the pip package's sm_100 build is an mma.sync (HMMA) kernel).
// qkv packed [B,S,3,H,d] bf16 is read in place through one 4-D TMA map (d, 3H, S, B).
//
// Forward: one CTA = 256 query rows (two 128-row tiles) of one (batch, head).  320 threads:
//   warp 0     TMA producer: both Q tiles once, then K_j / V_j tiles through a ring (shared by the two Q tiles)
//   warp 1     MMA issuer:   S_t = Q_t K_j^T (SS, K-major x K-major) into one of THREE rotating TMEM score buffers
//              (score tile n = 2j + t lives in buffer n % 3); O_t += P_t V_j (TS: P read from TMEM as the A operand, V
//              taken from the same row-major smem tile through an MN-major descriptor).  Issue order
//              QK(0) QK(1) QK(2) | PV(n) QK(n+3) ...: QK(n+3) reuses the buffer PV(n) has just consumed (the tensor
//              pipe executes in issue order), so a warpgroup's next score tile is complete before it finishes the
//              current one — with one buffer per Q tile the PV -> QK -> commit turnaround (~500 clk) sat on every
//              warpgroup's critical path and the SFU idled 40 % of the time (profiles/r1_attention_ncu.md).
//   warps 2-5  softmax warpgroup of Q tile 0, warps 6-9 of Q tile 1: thread <-> query row (TMEM lane), online softmax in
//              the log2 domain with lazy rescaling (O is only rescaled when the running max grows by more than 2^8),
//              P written back over S as bf16.
// The kernel is MUFU(ex2)-bound for head_dim 32 (128 MMA flops per score element): two warps per scheduler
// keep the SFU pipe busy.  head_dim 64 uses 128B swizzle, head_dim 32 uses 64B swizzle (TMA row pitch = inner box extent).
#include "tc_common.cuh"

#ifndef OCT_AT_POLY_EVERY  // one exponential pair in OCT_AT_POLY_EVERY is evaluated on the FMA pipe (0 = none)
#define OCT_AT_POLY_EVERY 4
#endif
#ifndef AT_KNOCK  // timing experiments ONLY (results wrong): 1 no exponentials, 2 no row maximum, 4 no P store, 8 no PV MMAs, 16 no QK MMAs
#define AT_KNOCK 0
#endif
#ifndef OCT_AT_SPIN        // 1: poll the issuer <-> softmax hand-off barriers instead of suspending on them
#define OCT_AT_SPIN 0
#endif
#if OCT_AT_SPIN
#define AT_WAIT tc::mbar_wait_spin
#else
#define AT_WAIT tc::mbar_wait
#endif

namespace {

constexpr int AT_BM = 128, AT_BN = 128, AT_QT = 2, AT_THREADS = 64 + AT_QT * 128, AT_KV_STAGES = 4, AT_SBUFS = 3;
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

template <int HD>
struct AtCfg {
  static constexpr int kRowBytes = HD * 2;                       // 64 or 128
  static constexpr int kTileBytes = 128 * kRowBytes;             // one 128-row Q / K / V tile
  static constexpr uint32_t kSwz = (HD == 64) ? tc::kSwz128 : tc::kSwz64;
  static constexpr uint32_t kSBO = 8 * kRowBytes;                // 8-row swizzle atom
  static constexpr int kSmem = kTileBytes * (AT_QT + 2 * AT_KV_STAGES) + 1024 + 256;
  static constexpr uint32_t kColS = 0;                  // score tile n = 2j + t at kColS + 128 (n % AT_SBUFS)
  static constexpr uint32_t kColO = 128 * AT_SBUFS;     // O_t at kColO + HD t  (384 + 2 HD <= 512)
};

struct AtParams {
  int S, H;
  float scale_log2e;
  __nv_bfloat16* out;  // [B,S,H,HD]
  float* lse;          // [B,H,S]
};

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // 3-input max (sm_100)
  return d;
}

// One 128x128 score tile of one query row: online-softmax update (m, l in the log2 domain), P written over S as bf16.
// kMasked = last kv tile (columns >= valid do not exist); the common full tile carries no per-element predicates.
template <int HD, bool kMasked>
__device__ __forceinline__ void softmax_tile(uint32_t tmem_s, uint32_t tmem_o, uint64_t* o_done, int j, int valid,
                                             float scale_log2e, float& m, float& l) {
  // Single pass over TMEM: the whole 128-column score row is held in registers (tcgen05.ld runs at ~64 B/clk/SM, so
  // reading S twice — once for the maximum, once for the exponentials — made the kernel TMEM-load-bound).
  uint32_t sr[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c) tc::tmem_ld_x32(tmem_s + c * 32, sr[c]);
  tc::tmem_ld_wait();
  if (kMasked) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c * 32 + i >= valid) sr[c][i] = 0xff800000u;  // -inf
  }
  // eight independent 3-input max chains (8 deep) instead of two (32 deep): the row maximum sits on the critical path of
  // every tile (nothing else of this warp can issue until it is known), so its latency, not its instruction count, matters
  float mx[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) mx[q] = -INFINITY;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; i += 16) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        mx[q] = max3(mx[q], __uint_as_float(sr[c][i + 2 * q]), __uint_as_float(sr[c][i + 2 * q + 1]));
    }
#if AT_KNOCK & 2
  const float m_tile = 1.0f;
#else
  const float m_tile = fmaxf(max3(mx[0], mx[1], mx[2]), fmaxf(max3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7]))) * scale_log2e;
#endif
  const bool need = m_tile > m + kRescaleThreshold;  // always true on the first tile (m = -inf)
  if (__any_sync(0xffffffffu, need)) {
    const float m_new = need ? m_tile : m;
    if (j > 0) {
      const float factor = need ? tc::fast_exp2(m - m_new) : 1.f;
      tc::mbar_wait(o_done, (j - 1) & 1);  // O += P(j-1) V_{j-1} has landed
      tc::tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t o[16];
        tc::tmem_ld_x16(tmem_o + c * 16, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
        tc::tmem_st_x16(tmem_o + c * 16, o);
      }
      l *= factor;
    }
    m = m_new;
  }
  // P = exp2(s * c - m) as bf16 pairs written over S (masked columns hold -inf -> exactly 0).  Packed fp32x2 math
  // (FFMA2 / FADD2) halves the issue slots around the exponentials, and one pair in kPolyEvery is evaluated by
  // exp2_poly2 on the FMA pipe: the loop is bound by MUFU.EX2 (16 / clk / SM), so moving a quarter of the exponentials
  // off the SFU shortens it (same idea as FlashAttention-4's software exp2).  The masked tile keeps the SFU for all
  // columns so that -inf stays exactly 0.
  constexpr int kPolyEvery = OCT_AT_POLY_EVERY;
  const uint64_t sc2 = tc::pack2(scale_log2e, scale_log2e), nm2 = tc::pack2(-m, -m);
  uint64_t la = tc::pack2(0.f, 0.f), lb = la;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint64_t x = tc::fma2(tc::pack2(__uint_as_float(sr[c][2 * i]), __uint_as_float(sr[c][2 * i + 1])), sc2, nm2);
      float p0, p1;
      if (AT_KNOCK & 1) {
        tc::unpack2(x, p0, p1);
      } else if (!kMasked && kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
        tc::exp2_poly2(x, p0, p1);
      } else {
        float x0, x1;
        tc::unpack2(x, x0, x1);
        p0 = tc::fast_exp2(x0);
        p1 = tc::fast_exp2(x1);
      }
      if (i & 1) lb = tc::add2(lb, tc::pack2(p0, p1)); else la = tc::add2(la, tc::pack2(p0, p1));
      pk[i] = pack_bf16x2(p0, p1);
    }
    if (!(AT_KNOCK & 4) || pk[3] == 0x12345678u) tc::tmem_st_x16(tmem_s + c * 16, pk);
  }
  float l0, l1;
  tc::unpack2(tc::add2(la, lb), l0, l1);
  l += l0 + l1;
}

// The ragged last kv tile when it holds at most 16 keys (S = 128 k + 1: the cls token of the MAE decoder / ViT sequences puts
// ONE key into a 33rd tile): scores from an N = 16 MMA, 16 columns of softmax, one K-step of PV.  A full-width masked tile for
// that single key cost every CTA 1/33 of its loop (tools/attn_overhead.py: S = 4097 vs 4096).
template <int HD>
__device__ __forceinline__ void softmax_tile_narrow(uint32_t tmem_s, uint32_t tmem_o, uint64_t* o_done, int j, int valid,
                                                    float scale_log2e, float& m, float& l) {
  uint32_t sr[16];
  tc::tmem_ld_x16(tmem_s, sr);
  tc::tmem_ld_wait();
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i >= valid) sr[i] = 0xff800000u;  // -inf
    mx = fmaxf(mx, __uint_as_float(sr[i]));
  }
  const float m_tile = mx * scale_log2e;
  const bool need = m_tile > m + kRescaleThreshold;  // always true on the first tile (m = -inf)
  if (__any_sync(0xffffffffu, need)) {
    const float m_new = need ? m_tile : m;
    if (j > 0) {
      const float factor = need ? tc::fast_exp2(m - m_new) : 1.f;
      tc::mbar_wait(o_done, (j - 1) & 1);  // O += P(j-1) V_{j-1} has landed
      tc::tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t o[16];
        tc::tmem_ld_x16(tmem_o + c * 16, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
        tc::tmem_st_x16(tmem_o + c * 16, o);
      }
      l *= factor;
    }
    m = m_new;
  }
  uint32_t pk[8];
  float ls = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // masked columns hold -inf -> exactly 0
    const float p0 = tc::fast_exp2(fmaf(__uint_as_float(sr[2 * i]), scale_log2e, -m));
    const float p1 = tc::fast_exp2(fmaf(__uint_as_float(sr[2 * i + 1]), scale_log2e, -m));
    ls += p0 + p1;
    pk[i] = pack_bf16x2(p0, p1);
  }
  tc::tmem_st_x8(tmem_s, pk);
  l += ls;
}

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                    const AtParams p) {
  using C = AtCfg<HD>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ provenance (LDS/STS instead of generic LD/ST): offset the array, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // [AT_QT]
  uint8_t* sK = smem + AT_QT * C::kTileBytes;           // [AT_KV_STAGES]
  uint8_t* sV = sK + AT_KV_STAGES * C::kTileBytes;      // [AT_KV_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT_KV_STAGES * C::kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + AT_KV_STAGES;
  // s_full / p_full are per score BUFFER (tile n and n + 3): QK(n+3) is issued behind PV(n), which waits for p_full(n),
  // which the warpgroup arrives on after consuming s_full(n) — no barrier can run two phases ahead of its waiter
  uint64_t* s_full = kv_empty + AT_KV_STAGES;  // [AT_SBUFS]
  uint64_t* p_full = s_full + AT_SBUFS;        // [AT_SBUFS]
  uint64_t* o_done = p_full + AT_SBUFS;        // [AT_QT] every PV of Q tile t (waited on only when O must be rescaled)
  uint64_t* o_final = o_done + AT_QT;          // [AT_QT] the LAST PV of Q tile t (single phase: the epilogue's wait)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + AT_QT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (AT_QT * AT_BM), h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.S + AT_BN - 1) / AT_BN;
  // The last CTA of a (batch, head) may own a single live Q tile (S = 4097: one query row in the 17th CTA); then score
  // tile n is simply kv tile n of Q tile 0 and warpgroup 1 idles, instead of sweeping the keys for 128 rows that do not exist.
  const int qsh = (q0 + AT_BM < p.S) ? 1 : 0;  // log2(number of live Q tiles): tile n -> (t = n & qsh, j = n >> qsh)
  const bool narrow_last = (p.S - (n_kv - 1) * AT_BN) <= 16;  // the last kv tile holds <= 16 keys: N = 16 scores, one PV K-step

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_qkv);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < AT_KV_STAGES; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < AT_SBUFS; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 128);
    }
    for (int t = 0; t < AT_QT; ++t) { tc::mbar_init(&o_done[t], 1); tc::mbar_init(&o_final[t], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(q_full, AT_QT * C::kTileBytes);
#pragma unroll
      for (int t = 0; t < AT_QT; ++t) tc::tma_load_4d(sQ + t * C::kTileBytes, &tmap_qkv, q_full, 0, h, q0 + t * AT_BM, b);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        tc::mbar_wait(&kv_empty[stage], phase ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[stage], 2 * C::kTileBytes);
        tc::tma_load_4d(sK + stage * C::kTileBytes, &tmap_qkv, &kv_full[stage], 0, p.H + h, j * AT_BN, b);
        tc::tma_load_4d(sV + stage * C::kTileBytes, &tmap_qkv, &kv_full[stage], 0, 2 * p.H + h, j * AT_BN, b);
        if (++stage == AT_KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and the waits; the MMAs are issued under elect.sync.  Inside a `lane == 0` branch
    // ptxas wraps EVERY tcgen05 instruction in an ELECT / BRA.U.ANY serialisation loop (it cannot prove a single active
    // thread), which made this warp — not the SFU — the bottleneck of the kernel (93 % busy, profiles/r1_attention_ncu.md).
    {
      constexpr uint32_t idesc_qk = tc::make_idesc(tc::kFmtBF16, false, false, AT_BM, AT_BN);
      constexpr uint32_t idesc_qk16 = tc::make_idesc(tc::kFmtBF16, false, false, AT_BM, 16);
      constexpr uint32_t idesc_pv = tc::make_idesc(tc::kFmtBF16, false, true, AT_BM, HD);
      auto issue_qk = [&](int n) {  // score tile n: Q tile n & qsh against K_(n >> qsh)
        const int t = n & qsh, stage = (n >> qsh) % AT_KV_STAGES;
        const uint32_t q_addr = tc::smem_u32(sQ + t * C::kTileBytes);
        const uint32_t k_addr = tc::smem_u32(sK + stage * C::kTileBytes);
        const uint32_t d_addr = tmem_base + C::kColS + (n % AT_SBUFS) * 128;
        const uint32_t idesc = (narrow_last && (n >> qsh) == n_kv - 1) ? idesc_qk16 : idesc_qk;
#pragma unroll
        for (int k = 0; k < ((AT_KNOCK & 16) ? 0 : HD / 16); ++k) {
          const uint64_t da = tc::make_smem_desc(q_addr + k * 32, 16, C::kSBO, C::kSwz);
          const uint64_t db = tc::make_smem_desc(k_addr + k * 32, 16, C::kSBO, C::kSwz);
          tc::mma_ss(d_addr, da, db, idesc, k != 0);
        }
        tc::mma_commit(&s_full[n % AT_SBUFS]);
      };
      auto issue_pv = [&](int n) {
        const int t = n & qsh, j = n >> qsh, stage = j % AT_KV_STAGES;
        const uint32_t v_addr = tc::smem_u32(sV + stage * C::kTileBytes);
        const uint32_t p_addr = tmem_base + C::kColS + (n % AT_SBUFS) * 128;
        if (narrow_last && j == n_kv - 1) {  // narrow tile: P holds 16 keys, one K-step
          const uint64_t db = tc::make_smem_desc(v_addr, C::kTileBytes, C::kSBO, C::kSwz);
          tc::mma_ts(tmem_base + C::kColO + t * HD, p_addr, db, idesc_pv, j != 0);
        } else {
#pragma unroll
          for (int k = 0; k < ((AT_KNOCK & 8) ? 0 : AT_BN / 16); ++k) {
            // V as an MN-major B operand: 16 kv rows per step; one MN chunk (= HD elements) so LBO is unused
            const uint64_t db = tc::make_smem_desc(v_addr + k * 16 * C::kRowBytes, C::kTileBytes, C::kSBO, C::kSwz);
            tc::mma_ts(tmem_base + C::kColO + t * HD, p_addr + k * 8, db, idesc_pv, (j | k) != 0);
          }
        }
        tc::mma_commit(&o_done[t]);
        if (j == n_kv - 1) tc::mma_commit(&o_final[t]);
      };
      auto wait_kv = [&](int j) {
        tc::mbar_wait(&kv_full[j % AT_KV_STAGES], (j / AT_KV_STAGES) & 1);
        tc::tcgen05_fence_after();
      };
      const int n_tiles = n_kv << qsh;
      // Score tiles are issued `ahead` tiles in front of their PV.  o_done[t] is only waited on when O must be rescaled, so
      // the barrier must never get two phases ahead of its waiter (the parity wait would alias): s_full(n) implies
      // PV(n - ahead) is complete, i.e. this warpgroup's PV(j - 2) with two live Q tiles and ahead = 3; with a single live
      // Q tile every PV belongs to warpgroup 0, hence ahead = 2 there.
      const int ahead = 2 + qsh;
      tc::mbar_wait(q_full, 0);
      const int n_pro = min(ahead, n_tiles);
      for (int n = 0; n < n_pro; ++n)
        if ((n & qsh) == 0) wait_kv(n >> qsh);
      if (tc::elect_one()) {
        for (int n = 0; n < n_pro; ++n) issue_qk(n);
      }
      __syncwarp();
      for (int n = 0; n < n_tiles; ++n) {
        const int t = n & qsh, j = n >> qsh;
        AT_WAIT(&p_full[n % AT_SBUFS], (n / AT_SBUFS) & 1);  // P(n) in TMEM (and O_t rescaled if needed)
        tc::tcgen05_fence_after();
        const bool more = n + ahead < n_tiles;
        if (more && ((n + ahead) & qsh) == 0) wait_kv((n + ahead) >> qsh);
        if (tc::elect_one()) {
          issue_pv(n);
          if (t == qsh) tc::mma_commit(&kv_empty[j % AT_KV_STAGES]);  // K_j, V_j fully consumed (by the last live Q tile)
          if (more) issue_qk(n + ahead);  // into the buffer PV(n) (or PV(n - 1)) has read: ordered behind it on the tensor pipe
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (((warp - 2) >> 2) <= qsh) {
    // ===================== softmax / correction / epilogue: warpgroup t owns Q tile t =====================
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tmem_o = lane_addr + C::kColO + t * HD;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int n = (j << qsh) + t, buf = n % AT_SBUFS;
      AT_WAIT(&s_full[buf], (n / AT_SBUFS) & 1);
      tc::tcgen05_fence_after();
      const int valid = p.S - j * AT_BN;  // columns >= valid are out of range (TMA zero-filled K rows)
      const uint32_t tmem_s = lane_addr + C::kColS + buf * 128;
      if (valid >= AT_BN)
        softmax_tile<HD, false>(tmem_s, tmem_o, &o_done[t], j, valid, p.scale_log2e, m, l);
      else if (narrow_last)
        softmax_tile_narrow<HD>(tmem_s, tmem_o, &o_done[t], j, valid, p.scale_log2e, m, l);
      else
        softmax_tile<HD, true>(tmem_s, tmem_o, &o_done[t], j, valid, p.scale_log2e, m, l);
      tc::tmem_st_wait();
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&p_full[buf]);
    }
    // epilogue: O / l -> bf16, lse.  The last PV of this Q tile arrives on its own single-phase barrier: a parity wait on o_done
    // is only unambiguous while that barrier is at most one phase behind the phase waited for, and nothing guarantees that
    // PV(n_kv - 2) has completed when the last tile's softmax is done (it normally has — the softmax takes longer than a PV round
    // trip — but a 16-key ragged-tile experiment did not, and read O two PVs early).
    tc::mbar_wait(&o_final[t], 0);
    tc::tcgen05_fence_after();
    const int qi = q0 + t * AT_BM + row;
    const float inv = 1.f / l;
    __nv_bfloat16* orow = p.out + (((size_t)b * p.S + qi) * p.H + h) * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[32];
      tc::tmem_ld_x32(tmem_o + c * 32, o);
      tc::tmem_ld_wait();
      if (qi < p.S) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + i) = v;
        }
      }
    }
    if (qi < p.S) p.lse[((size_t)b * p.H + h) * p.S + qi] = (m + log2f(l)) * kLn2;
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

int make_qkv_map(CUtensorMap* map, const void* qkv, int64_t B, int64_t S, int64_t H, int64_t d, uint32_t box_rows,
                 const char* who) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)3 * H * d * 2, (uint64_t)S * 3 * H * d * 2};
  uint32_t box[4] = {(uint32_t)d, 1, box_rows, 1};
  return oct_make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, qkv, dims, str, box, who,
                       d == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int HD>
int launch_fwd(const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, float scale, cudaStream_t st) {
  using C = AtCfg<HD>;
  CUtensorMap map;
  int rc = make_qkv_map(&map, qkv, B, S, H, HD, 128, "oct_attn_fwd(bf16)");
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) { oct_set_error("oct_attn_fwd(bf16): smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  AtParams p;
  p.S = (int)S; p.H = (int)H; p.scale_log2e = scale * kLog2e; p.out = (__nv_bfloat16*)out; p.lse = lse;
  dim3 grid((unsigned)ceil_div64(S, AT_QT * AT_BM), (unsigned)H, (unsigned)B);
  {
    cudaError_t e = oct_launch(attn_fwd_tc_kernel<HD>, grid, dim3(AT_THREADS), (size_t)C::kSmem, st, 1, map, p);
    if (e != cudaSuccess) { oct_set_error("oct_attn_fwd(bf16): launch: %s", cudaGetErrorString(e)); return (int)e; }
  }
  return oct_check_launch("oct_attn_fwd(bf16)");
}

}  // namespace

int oct_attn_fwd_tc(const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H, int64_t d, float scale,
                    cudaStream_t st) {
  OCT_REQUIRE(aligned16(qkv) && aligned16(out), "oct_attn_fwd(bf16): pointers must be 16-byte aligned");
  OCT_REQUIRE(S < (1 << 24), "oct_attn_fwd(bf16): S too large");
  if (d == 64) return launch_fwd<64>(qkv, out, lse, B, S, H, scale, st);
  if (d == 32) return launch_fwd<32>(qkv, out, lse, B, S, H, scale, st);
  oct_set_error("oct_attn_fwd(bf16): head dim %lld unsupported by the tcgen05 kernel (32 or 64)", (long long)d);
  return OCT_ERR_UNSUPPORTED;
}
