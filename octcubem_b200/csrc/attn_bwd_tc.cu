// Backward of the tcgen05 flash attention (autograd of flash_attn_qkvpacked_func, flash_attn/modules/mha.py:122-130).
//
// One CTA owns one 128-row K/V tile of one (batch, head) and sweeps the query tiles in 64-row sub-tiles.  Everything is
// computed in the TRANSPOSED orientation so that the softmax threads own kv rows (= TMEM lanes) and P^T / dS^T can be
// fed back to the tensor core as TMEM A-operands without ever leaving TMEM:
//     S^T  = K Q_h^T            (SS, K-major x K-major, N = 64)  P^T  = exp2(S^T c - lse)
//     dP^T = V dO_h^T           (SS)                             dS^T = P^T o (dP^T - delta)
//     dV  += P^T  dO_h          (TS, dO rows as MN-major B)      dK  += dS^T Q_h   (TS, Q rows as MN-major B)
//     dQ_m = dS K               (SS, once per 128-row query tile: dS^T staged in smem, read as an MN-major A operand)
// dV / dK accumulate in TMEM across the sweep; dQ_m tiles go TMEM -> registers -> swizzled smem -> one asynchronous
// cp.reduce.async.bulk (.add.f32) into an fp32 workspace that a small kernel scales and converts to bf16.
//
// What measurements say bounds the sweep (profiles/r2_attention_ncu.md: knock-out timings, clock64 traces of every role, ncu
// stall samples): not the tensor pipe (removing any MMA family changes nothing; tools/microbench/mma.cu: the 16 MMAs of a sub-tile
// occupy it ~490 clk of ~1000) and not the MUFU (moving exponentials to the FMA pipe is slower), but the LATENCY of the softmax
// warps' dependent instruction stream and of the stage hand-offs.  Hence:
//   * the fp32 S^T / dP^T sub-tiles have THREE stages in TMEM at head_dim 32 (3 x 128 + 96 columns; two at head_dim 64);
//   * kAlt: the two softmax warps of a scheduler work on DIFFERENT sub-tiles (group g = warp >> 2 owns the sub-tiles of its
//     parity, all 64 query columns, in four 16-column chunks) instead of running in phase on the two column halves of one;
//   * the softmax statistics (-lse log2e, -delta, pre-masked by attn_delta_kernel) arrive by bulk copy with the query tile in its
//     ring stage: per-thread global loads + an smem staging round cost every warp ~500 clk per sub-tile;
//   * bf16 P^T / dS^T overwrite the fp32 columns their own thread has consumed (no extra columns, no cross-warp hazard);
//   * one arrival per WARP on the hand-off barriers;
//   * the MMAs are issued by two warps with their own barriers: warp 9 the gradient MMAs G(i) and, right behind them in the same
//     in-order stream, the scores S/dP(i + kNS) into the stage G(i) has just read; warp 10 the dQ MMAs (ds_ready / ds_free).  One
//     warp issuing everything executed ~200 instructions per sub-tile as one dependent stream (~600 clk with the pipe idle).
// 512 threads in four warpgroups: warps 0-7 softmax-backward, warp 8 TMA producer, warps 9 / 10 MMA issuers (11 only with
// -DAB_MERGE=0), warps 12-15 drain the dQ tiles (TMEM -> smem -> bulk reduction).  setmaxnreg moves the registers of the light
// warpgroups to the softmax warpgroups.  Per-CTA fixed cost: 4.9 us of 39 us at S = 4097 (tools/attn_overhead.py); a persistent
// variant halved it but ran the sweep 12 % slower and was dropped (profiles/r2_attention_ncu.md).
#include "tc_common.cuh"
#include <type_traits>
#include <cstdlib>

namespace {

constexpr int AB_T = 128, AB_SUB = 64, AB_THREADS = 512, AB_SM_THREADS = 256, AB_DRAIN_THREADS = 128;
// register budget (setmaxnreg, per warpgroup): 8 softmax warps x 176 + 8 control / drain warps x 80 = 64 K registers
constexpr int AB_REGS_SOFTMAX = 176, AB_REGS_OTHER = 80;
constexpr float kLog2e = 1.4426950408889634f;
#ifndef AB_NS32   // experiment knobs (tools/build_variant.sh): score stages / alternating softmax groups at head_dim 32
#define AB_NS32 3
#endif
#ifndef AB_ALT32
#define AB_ALT32 1
#endif
#ifndef AB_POLY   // pairs out of every 4 (two q-iterations) whose exponentials run on the FMA pipe: 0 .. 4
#define AB_POLY 0
#endif
#ifndef AB_MERGE  // 1: warp 9 issues G(i) and S/dP(i + kNS) back to back (warp 11 idle); 0: warp 11 issues S/dP after g_done
#define AB_MERGE 1
#endif
#ifndef AB_KNOCK  // timing experiments ONLY (results are wrong): 1 no exponentials, 2 no dS^T st.shared, 4 no dQ MMAs, 8 no G MMAs,
#define AB_KNOCK 0  // 16 no tcgen05.st of P^T / dS^T, 32 no tcgen05.ld of the scores, 64 no S/dP MMAs
#endif

template <int HD>
struct AbCfg {
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kTileBytes = 128 * kRowBytes;
  static constexpr uint32_t kSwz = (HD == 64) ? tc::kSwz128 : tc::kSwz64;
  static constexpr uint32_t kSBO = 8 * kRowBytes;
  static constexpr int kDsBytes = 2 * 128 * 128;  // one dS^T staging tile: two 64-query chunks of [128 kv rows x 128 B]
  static constexpr int kDqBytes = 128 * HD * 4;   // fp32 dQ tile staged for the bulk reduce (16-byte chunks XOR-swizzled)
  // smem: K, V | Q[2], dO[2] | dS[2] | dQ staging | lse2[2][128], delta[2][128] | barriers
  // Q / dO ring depth.  The stage of query tile m + kQStages is refilled once tile m has been consumed; with two stages that
  // is ONE sub-tile (~1400 clk) before the issuer needs the data, less than a TMA round trip: the clock64 trace showed the
  // issuer waiting ~1100 clk for q_full at every second sub-tile (38 % of the kernel).  head_dim 64 has no room for more.
  static constexpr int kQStages = (HD == 32) ? 4 : 2;
  // fp32 score stages in TMEM.  head_dim 32 has room for three (3 x 128 + 96 columns): the scores of sub-tile i + 3 are
  // queued behind the gradient MMAs of sub-tile i, so a softmax group that finishes sub-tile i finds S/dP(i + 2) complete
  // instead of waiting for its own p_ready -> issuer -> tensor pipe round trip.
  static constexpr int kNS = (HD == 32) ? AB_NS32 : 2;
  // kAlt: the two softmax warps of a scheduler work on DIFFERENT sub-tiles (group g = warp >> 2 owns the sub-tiles
  // i = g mod 2, all 64 query columns) instead of on the two column halves of the same one.  In the column-split form both
  // warps wait on the same barrier and run in phase: their exponentials contend for the MUFU (512 clk for the pair) and
  // then BOTH sit in the TMEM store / fence / barrier tail with the MUFU idle (~1230 clk per sub-tile, clock64 trace).
  static constexpr bool kAlt = (HD == 32) && (AB_ALT32 != 0);
  static constexpr int kStatFloats = kQStages * 256;  // per Q/dO ring stage: [128 x -lse*log2e | 128 x -delta] of the query tile
  static constexpr int kSmem = 2 * kTileBytes + 2 * kQStages * kTileBytes + 2 * kDsBytes + kDqBytes + kStatFloats * 4 + 1024 + 512;
  // TMEM columns: stage s of the fp32 sub-tiles: S^T at 128 s, dP^T at 128 s + 64; accumulators behind them
  static constexpr uint32_t kColST = 0, kColDPT = 64, kStageCols = 128;
  static constexpr uint32_t kColDV = kNS * 128, kColDK = kNS * 128 + HD, kColDQ = kNS * 128 + 2 * HD;
  static_assert(kColDQ + HD <= 512, "TMEM budget");
  // bf16 K-slice k (16 queries, 8 columns) of a sub-tile: the warp that owns query columns [32 c, 32 c + 32) writes its
  // bf16 output over its own fp32 columns -> slices 0,1 at columns 0,8 and slices 2,3 at columns 32,40
  __host__ __device__ static constexpr uint32_t slice_off(int k) { return (uint32_t)((k >> 1) * 32 + (k & 1) * 8); }
};

struct AbParams {
  int S, H, Spad;
  float scale, scale_log2e;
  const float* nlse2;   // [B,H,Spad]  -lse * log2(e)  (-inf for s >= S), written by attn_delta_kernel
  const float* ndelta;  // [B,H,Spad]  -delta          (0 for s >= S)
  float* dq_acc;        // [B,H,Spad/128][lane quarter][HD/4 chunks][32 rows][4] fp32, zero-initialised (see the drain warps)
  __nv_bfloat16* dqkv;  // [B,S,3,H,HD]
  int dbg;              // OCT_ATTN_BWD_DBG=16: record a cycle trace of CTA (1,0,0) (diagnostics only)
};

#ifndef OCT_AB_TRACE  // build with -DOCT_AB_TRACE=1 for the clock64 trace; the probes cost ~10 % of the issued instructions
#define OCT_AB_TRACE 0
#endif
#ifndef OCT_AB_SPIN   // 1: poll the issuer <-> softmax hand-off barriers instead of suspending on them
#define OCT_AB_SPIN 0
#endif
#if OCT_AB_SPIN
#define AB_WAIT tc::mbar_wait_spin
#else
#define AB_WAIT tc::mbar_wait
#endif
__device__ long long g_ab_trace[8 * 48];
#if !OCT_AB_TRACE
#define AB_TRACE(id) do {} while (0)
#define AB_TRACEW(id) do {} while (0)
#else
#define AB_TRACE(id)                                                                                              \
  do {                                                                                                           \
    if ((p.dbg & 16) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp >= 9 && warp <= 11 || warp == 0 || (C::kAlt && warp == 4)) && \
        i >= 8 && i < 16)                                                                                        \
      g_ab_trace[(i - 8) * 48 + (id)] = clock64();                                                               \
  } while (0)
#define AB_TRACEW(id)                                                                                             \
  do {                                                                                                           \
    if ((p.dbg & 16) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && i >= 8 && i < 16)   \
      g_ab_trace[(i - 8) * 48 + 16 + 4 * warp + (id)] = clock64();                                                \
  } while (0)
#endif

template <int HD>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const AbParams p) {
  using C = AbCfg<HD>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ provenance (LDS/STS instead of generic LD/ST): offset the array, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = sK + C::kTileBytes;
  uint8_t* sQ = sV + C::kTileBytes;                       // [C::kQStages]
  uint8_t* sDO = sQ + C::kQStages * C::kTileBytes;        // [C::kQStages]
  uint8_t* sDS = sDO + C::kQStages * C::kTileBytes;       // [2] x 32 KB, 1024-aligned (all tiles are multiples of 8 KB)
  uint8_t* sDQ = sDS + 2 * C::kDsBytes;                   // [128][HD] fp32, swizzled
  float* sStat = reinterpret_cast<float*>(sDQ + C::kDqBytes);  // [kQStages][128 x -lse*log2e | 128 x -delta], filled by the producer
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + C::kStatFloats);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;                 // [2]
  uint64_t* q_empty = q_full + C::kQStages;    // [2]
  uint64_t* sdp_full = q_empty + C::kQStages;  // [kNS] per fp32 stage, one completion every kNS sub-tiles
  uint64_t* p_ready = sdp_full + C::kNS;       // [kNS] per fp32 stage (one arrival per softmax warp that worked on it)
  uint64_t* g_done = p_ready + C::kNS;         // [kNS] G(i) complete: the stage may receive the scores of sub-tile i + kNS
  uint64_t* ds_ready = g_done + C::kNS;        // [2] per dS^T buffer: both 64-query chunks written (softmax warps of two sub-tiles)
  uint64_t* ds_free = ds_ready + 2;            // [2] per dS^T buffer: dQ_m has read it
  uint64_t* dq_full = ds_free + 2;             // once per query tile
  uint64_t* dq_free = dq_full + 1;             // once per query tile (the 128 drain threads)
  uint64_t* acc_full = dq_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.S + AB_T - 1) / AB_T;
  // 64-query sub-tiles that hold at least one query: the second half of the last 128-row tile may be empty (S = 4097:
  // 65 sub-tiles, not 66).  The dQ MMA of that tile then reads a stale (finite or zeroed) second dS^T chunk, which only
  // reaches accumulator rows >= S that nobody reads.
  const int n_sub = (p.S + AB_SUB - 1) / AB_SUB;

  if (warp == 8 && lane == 0) {
    tc::prefetch_tmap(&tmap_qkv);
    tc::prefetch_tmap(&tmap_do);
    tc::mbar_init(kv_full, 1);
    for (int s = 0; s < C::kQStages; ++s) { tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < C::kNS; ++s) {
      tc::mbar_init(&sdp_full[s], 1);
      tc::mbar_init(&p_ready[s], C::kAlt ? 4 : 8);
      tc::mbar_init(&g_done[s], 1);
    }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&ds_ready[s], C::kAlt ? 8 : 16); tc::mbar_init(&ds_free[s], 1); }
    tc::mbar_init(dq_full, 1);
    tc::mbar_init(dq_free, AB_DRAIN_THREADS);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 9) tc::tmem_alloc<512>(tmem_slot);
  if ((n_sub & 1) && warp < 8) {  // never-written half of the last tile's dS^T buffer (read by its dQ MMA, see above)
    uint4* z = reinterpret_cast<uint4*>(sDS + ((n_q - 1) & 1) * C::kDsBytes + 128 * 128);
    for (int i = threadIdx.x; i < 128 * 128 / 16; i += AB_SM_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // setmaxnreg sits at the top of each role's branch: ptxas budgets the registers of a region by the setmaxnreg that
  // dominates it (after a common if / else it applies the smaller value to everything that follows).
  if (warp >= 8 && warp < 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_OTHER));
  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(kv_full, 2 * C::kTileBytes);
      tc::tma_load_4d(sK, &tmap_qkv, kv_full, 0, p.H + h, n0, b);
      tc::tma_load_4d(sV, &tmap_qkv, kv_full, 0, 2 * p.H + h, n0, b);
      int stage = 0; uint32_t phase = 0;
      for (int m = 0; m < n_q; ++m) {
        tc::mbar_wait(&q_empty[stage], phase ^ 1);
        tc::mbar_arrive_expect_tx(&q_full[stage], 2 * C::kTileBytes + 2 * AB_T * 4);
        tc::tma_load_4d(sQ + stage * C::kTileBytes, &tmap_qkv, &q_full[stage], 0, h, m * AB_T, b);
        tc::tma_load_4d(sDO + stage * C::kTileBytes, &tmap_do, &q_full[stage], 0, h, m * AB_T, b);
        // the tile's softmax statistics ride in the same ring stage: per-thread global loads + an smem staging round per
        // sub-tile cost the softmax warps ~500 clk of their ~1800 clk period (clock64 trace; -10 % kernel time without the loads)
        const size_t srow = ((size_t)b * p.H + h) * p.Spad + (size_t)m * AB_T;
        tc::bulk_load_1d(sStat + stage * 256, p.nlse2 + srow, AB_T * 4, &q_full[stage]);
        tc::bulk_load_1d(sStat + stage * 256 + 128, p.ndelta + srow, AB_T * 4, &q_full[stage]);
        if (++stage == C::kQStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9 || warp == 10 || warp == 11) {
    // ===================== three MMA issuers =====================
    // One warp issuing everything executed ~200 instructions per sub-tile (three barrier waits, descriptor arithmetic in
    // uniform registers, 16-24 MMAs, commits) in ONE dependent stream: ~600 clk per sub-tile on top of the ~620 clk its
    // MMAs keep the tensor pipe busy, with the pipe idle during the former (clock64 trace, profiles/r2_attention_ncu.md).
    // The three MMA families have different producers and consumers, so each gets its own warp and its own barriers:
    //   warp 9   G(i):    dV += P^T dO_h, dK += dS^T Q_h    after p_ready[st]          -> g_done[st], q_empty, acc_full
    //   warp 10  dQ(m):   dQ_m = dS K                       after ds_ready[m&1], dq_free -> dq_full, ds_free[m&1]
    //   warp 11  S/dP(j): S^T = K Q_h^T, dP^T = V dO_h^T    after g_done[st] (G(j - kNS) has read the stage's bf16 contents),
    //                     q_full, ds_free (dQ of query tile (j>>1) - 2 has read the dS^T buffer the softmax warps of
    //                     sub-tile j are about to overwrite)                            -> sdp_full[st]
    // tcgen05.commit only covers the MMAs of the committing thread; every cross-warp ordering goes through one of the
    // barriers above.  Each warp runs its loop and waits with all lanes; MMAs / commits are issued under elect.sync
    // (inside a `lane == 0` branch ptxas wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop, attn_tc.cu).
    constexpr uint32_t kStageStep = C::kTileBytes >> 4, kHalfStep = (AB_SUB * C::kRowBytes) >> 4;
    constexpr uint32_t kKStepK = 32 >> 4, kKStepMN = (16 * C::kRowBytes) >> 4, kKStepDS = (16 * 128) >> 4;
    constexpr uint32_t kDsBufStep = C::kDsBytes >> 4;
    // Descriptors are built once; inside the loops only their 14-bit start-address field (units of 16 B) is advanced.
    if (warp == 9) {
      constexpr uint32_t idesc_acc = tc::make_idesc(tc::kFmtBF16, false, true, 128, HD);       // P^T dO_h, dS^T Q_h (TS)
      const uint64_t dQ0_mn = tc::make_smem_desc(tc::smem_u32(sQ), C::kTileBytes, C::kSBO, C::kSwz);
      const uint64_t dDO0_mn = tc::make_smem_desc(tc::smem_u32(sDO), C::kTileBytes, C::kSBO, C::kSwz);
#if AB_MERGE
      // AB_MERGE: this warp also queues S/dP(i + kNS) right behind G(i) (same stage; the in-order MMA stream of one thread
      // needs no completion wait in between: one commit -> barrier -> wake hop (~500 clk) less per stage round trip)
      constexpr uint32_t idesc_st = tc::make_idesc(tc::kFmtBF16, false, false, 128, AB_SUB);
      const uint64_t dK_kmaj = tc::make_smem_desc(tc::smem_u32(sK), 16, C::kSBO, C::kSwz);
      const uint64_t dV_kmaj = tc::make_smem_desc(tc::smem_u32(sV), 16, C::kSBO, C::kSwz);
      const uint64_t dQ0_kmaj = tc::make_smem_desc(tc::smem_u32(sQ), 16, C::kSBO, C::kSwz);
      const uint64_t dDO0_kmaj = tc::make_smem_desc(tc::smem_u32(sDO), 16, C::kSBO, C::kSwz);
      auto issue_sdp = [&](int j) {  // elected lane
        const uint32_t joff = ((j >> 1) % C::kQStages) * kStageStep + (j & 1) * kHalfStep;
        const uint32_t jcol = tmem_base + (j % C::kNS) * C::kStageCols;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::mma_ss(jcol + C::kColST, dK_kmaj + k * kKStepK, dQ0_kmaj + joff + k * kKStepK, idesc_st, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::mma_ss(jcol + C::kColDPT, dV_kmaj + k * kKStepK, dDO0_kmaj + joff + k * kKStepK, idesc_st, k != 0);
        tc::mma_commit(&sdp_full[j % C::kNS]);
      };
      auto wait_for_sdp = [&](int j) {  // all lanes: operands of S/dP(j) present, dS^T buffer of its query tile released
        if ((j & 1) == 0) {
          const int t = j >> 1;
          tc::mbar_wait(&q_full[t % C::kQStages], (t / C::kQStages) & 1);
          if (t >= 2) tc::mbar_wait(&ds_free[t & 1], ((t - 2) >> 1) & 1);
        }
      };
      tc::mbar_wait(kv_full, 0);
      for (int j = 0; j < C::kNS && j < n_sub; ++j) {
        wait_for_sdp(j);
        tc::tcgen05_fence_after();
        if (tc::elect_one()) issue_sdp(j);
        __syncwarp();
      }
#endif
      for (int i = 0; i < n_sub; ++i) {
        const int m = i >> 1, hh = i & 1, st = i % C::kNS;
        const uint32_t off = (m % C::kQStages) * kStageStep + hh * kHalfStep;
        const uint32_t tcol = tmem_base + st * C::kStageCols;
        AB_TRACE(0);
#if AB_MERGE
        if (i + C::kNS < n_sub) wait_for_sdp(i + C::kNS);  // long satisfied; off the p_ready -> G critical path
#endif
        AB_WAIT(&p_ready[st], (i / C::kNS) & 1);  // bf16 P^T / dS^T of sub-tile i in TMEM
        tc::tcgen05_fence_after();
        AB_TRACE(1);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < ((AB_KNOCK & 8) ? 0 : AB_SUB / 16); ++k)  // dV += P^T dO_h
            tc::mma_ts(tmem_base + C::kColDV, tcol + C::kColST + C::slice_off(k), dDO0_mn + off + k * kKStepMN, idesc_acc,
                       (i | k) != 0);
#pragma unroll
          for (int k = 0; k < ((AB_KNOCK & 8) ? 0 : AB_SUB / 16); ++k)  // dK += dS^T Q_h
            tc::mma_ts(tmem_base + C::kColDK, tcol + C::kColDPT + C::slice_off(k), dQ0_mn + off + k * kKStepMN, idesc_acc,
                       (i | k) != 0);
#if !AB_MERGE
          tc::mma_commit(&g_done[st]);
#endif
          // Q_m / dO_m fully consumed: this G is causally after S/dP of both halves (p_ready <- softmax <- sdp_full)
          if (hh == 1 || i == n_sub - 1) tc::mma_commit(&q_empty[m % C::kQStages]);
#if AB_MERGE
          if (i + C::kNS < n_sub) issue_sdp(i + C::kNS);
#endif
        }
        __syncwarp();
        AB_TRACE(3);
      }
      if (tc::elect_one()) tc::mma_commit(acc_full);
      __syncwarp();
    } else if (warp == 10) {
      constexpr uint32_t idesc_dq = tc::make_idesc(tc::kFmtBF16, true, true, 128, HD);         // dS K (A MN-major)
      const uint64_t dK_mn = tc::make_smem_desc(tc::smem_u32(sK), C::kTileBytes, C::kSBO, C::kSwz);          // K as MN-major B
      const uint64_t dDS_mn = tc::make_smem_desc(tc::smem_u32(sDS), 128 * 128, 1024, tc::kSwz128);           // dS^T as MN-major A
      tc::mbar_wait(kv_full, 0);
      for (int m = 0; m < n_q; ++m) {
        [[maybe_unused]] const int i = 2 * m + 1;                 // (trace slot)
        AB_WAIT(&ds_ready[m & 1], (m >> 1) & 1);            // both 64-query chunks of dS^T(m) are in smem
        if (m > 0) tc::mbar_wait(dq_free, (m - 1) & 1);           // previous dQ tile drained from TMEM
        tc::tcgen05_fence_after();
        AB_TRACE(12);
        if (tc::elect_one()) {
          const uint32_t dsoff = (m & 1) * kDsBufStep;
#pragma unroll
          for (int k = 0; k < ((AB_KNOCK & 4) ? 0 : 128 / 16); ++k)  // dQ_m = dS K : A = dS^T smem tile read MN-major (M = q contiguous)
            tc::mma_ss(tmem_base + C::kColDQ, dDS_mn + dsoff + k * kKStepDS, dK_mn + k * kKStepMN, idesc_dq, k != 0);
          tc::mma_commit(dq_full);
          tc::mma_commit(&ds_free[m & 1]);
        }
        __syncwarp();
        AB_TRACE(13);
      }
    } else if (!AB_MERGE) {
      constexpr uint32_t idesc_st = tc::make_idesc(tc::kFmtBF16, false, false, 128, AB_SUB);  // K Q_h^T, V dO_h^T
      const uint64_t dK_kmaj = tc::make_smem_desc(tc::smem_u32(sK), 16, C::kSBO, C::kSwz);       // K as K-major A
      const uint64_t dV_kmaj = tc::make_smem_desc(tc::smem_u32(sV), 16, C::kSBO, C::kSwz);       // V as K-major A
      const uint64_t dQ0_kmaj = tc::make_smem_desc(tc::smem_u32(sQ), 16, C::kSBO, C::kSwz);      // stage 0, half 0
      const uint64_t dDO0_kmaj = tc::make_smem_desc(tc::smem_u32(sDO), 16, C::kSBO, C::kSwz);
      tc::mbar_wait(kv_full, 0);
      // sub-tile j -> (q tile j>>1, half j&1); its Q/dO ring stage is (j>>1) % C::kQStages, its fp32 stage is j % kNS
      for (int j = 0; j < n_sub; ++j) {
        const int st = j % C::kNS, t = j >> 1;
        [[maybe_unused]] const int i = j;
        AB_TRACE(10);
        if (j >= C::kNS) AB_WAIT(&g_done[st], ((j - C::kNS) / C::kNS) & 1);
        if ((j & 1) == 0) {
          tc::mbar_wait(&q_full[t % C::kQStages], (t / C::kQStages) & 1);   // first use of Q/dO tile t
          if (t >= 2) tc::mbar_wait(&ds_free[t & 1], ((t - 2) >> 1) & 1);   // dQ(t - 2) has read the dS^T buffer of tile t
        }
        tc::tcgen05_fence_after();
        AB_TRACE(11);
        if (tc::elect_one()) {
          const uint32_t off = (t % C::kQStages) * kStageStep + (j & 1) * kHalfStep;
          const uint32_t tcol = tmem_base + st * C::kStageCols;
#pragma unroll
          for (int k = 0; k < ((AB_KNOCK & 64) ? 0 : HD / 16); ++k)
            tc::mma_ss(tcol + C::kColST, dK_kmaj + k * kKStepK, dQ0_kmaj + off + k * kKStepK, idesc_st, k != 0);
#pragma unroll
          for (int k = 0; k < ((AB_KNOCK & 64) ? 0 : HD / 16); ++k)
            tc::mma_ss(tcol + C::kColDPT, dV_kmaj + k * kKStepK, dDO0_kmaj + off + k * kKStepK, idesc_st, k != 0);
          tc::mma_commit(&sdp_full[st]);
        }
        __syncwarp();
        AB_TRACE(2);
      }
    }
  }
  } else if (warp >= 12) {
    // ===================== dQ drain: warp 12 + q owns TMEM lane quarter q =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AB_REGS_OTHER));
    // TMEM -> registers -> per-warp smem block -> ONE asynchronous bulk reduction of that block into the fp32 accumulator.
    // Accumulator layout per 128-query tile: [lane quarter][16-byte chunk][32 rows][4 floats] (attn_dq_convert_kernel
    // undoes it): contiguous per warp, and the st.shared.v4 of a warp hit 32 different bank groups.
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const size_t bh = (size_t)b * p.H + h;
    constexpr int kWarpDqBytes = 32 * HD * 4;  // 32 rows x HD columns fp32
    uint8_t* sdq_w = sDQ + quarter * kWarpDqBytes;
    for (int m = 0; m < n_q; ++m) {
      tc::mbar_wait(dq_full, m & 1);
      tc::tcgen05_fence_after();
      uint32_t o[HD];
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t(&oc)[32] = reinterpret_cast<uint32_t(&)[32]>(o[c * 32]);
        tc::tmem_ld_x32(lane_addr + C::kColDQ + c * 32, oc);
      }
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();
      tc::mbar_arrive(dq_free);  // dQ is in registers: the issuer may overwrite the TMEM tile
      // this warp's bulk reduction of the previous query tile must have finished reading its staging block
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int q = 0; q < HD / 4; ++q)
        *reinterpret_cast<uint4*>(sdq_w + (q * 32 + lane) * 16) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !(AB_KNOCK & 128)) {
        float* dst = p.dq_acc + (bh * p.Spad + (size_t)m * AB_T) * HD + quarter * (kWarpDqBytes / 4);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                     "r"(tc::smem_u32(sdq_w)), "r"((uint32_t)kWarpDqBytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem read; visibility = kernel boundary
  } else {
    // ===================== softmax-backward threads =====================
    // column-split form: (kv row, 32 query columns) of EVERY sub-tile; kAlt: (kv row, 64 query columns) of every OTHER one
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AB_REGS_SOFTMAX));
    const int quarter = warp & 3, colhalf = warp >> 2;  // kAlt: colhalf is the GROUP (parity of the sub-tiles it owns)
    const int row = quarter * 32 + lane;  // kv row inside the tile == TMEM lane
    const bool kv_ok = (n0 + row) < p.S;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const size_t bh = (size_t)b * p.H + h;
    constexpr int kNC = C::kAlt ? 2 : 1;  // 32-column halves per warp and sub-tile
    // The eight warps never synchronise with each other (only through p_ready -> the MMA issuers).  The per-query statistics
    // (-lse log2e, -delta; masked for s >= S by attn_delta_kernel) arrive with the query tile in its ring stage.
    const int i_first = C::kAlt ? colhalf : 0, i_step = C::kAlt ? 2 : 1;
    for (int i = i_first; i < n_sub; i += i_step) {
      const int m = i >> 1, hh = i & 1, st = i % C::kNS;
      const float* stat = sStat + (m % C::kQStages) * 256 + hh * AB_SUB;
      // the stage's bulk copies completed on q_full (long ago: the S/dP issuer waited for it before the scores of this
      // sub-tile were queued); observing the phase here makes their bytes visible to this thread
      tc::mbar_wait(&q_full[m % C::kQStages], (m / C::kQStages) & 1);
      AB_TRACE(4);
      AB_TRACEW(2);
      AB_WAIT(&sdp_full[st], (i / C::kNS) & 1);
      tc::tcgen05_fence_after();
      AB_TRACE(5);
      AB_TRACEW(0);
      // Packed fp32x2 math (FFMA2 / FADD2 / FMUL2): half the issue slots around the exponentials.  No masking is
      // needed: query columns past S carry lse = +inf (P = 0), and kv rows past S have zero-filled K / V rows, so
      // their (finite) P and dS only reach dV / dK rows that are never stored and add dS * 0 to dQ.
      const uint64_t sc2 = tc::pack2(p.scale_log2e, p.scale_log2e);
      // 16-query chunks (= one K-slice of the gradient MMAs each), software-pipelined: the TMEM loads of chunk k + 1 are in
      // flight while chunk k is computed, so only the first load of a sub-tile is exposed (tcgen05.wait::ld covers every
      // outstanding load: wait, THEN issue the next pair).  kAlt: chunks 0-3; column-split form: chunks 2 colhalf, 2 colhalf + 1.
      constexpr int kChunks = C::kAlt ? 4 : 2;
      const int k0 = C::kAlt ? 0 : 2 * colhalf;
      const uint32_t tst = lane_addr + st * C::kStageCols;
      uint8_t* rowp = sDS + (m & 1) * C::kDsBytes + hh * (128 * 128) + row * 128;
      uint32_t s[2][16], dp[2][16];
#if AB_KNOCK & 32
#pragma unroll
      for (int q = 0; q < 16; ++q) { s[0][q] = s[1][q] = q * lane; dp[0][q] = dp[1][q] = q + lane; }
#else
      tc::tmem_ld_x16(tst + C::kColST + k0 * 16, s[0]);
      tc::tmem_ld_x16(tst + C::kColDPT + k0 * 16, dp[0]);
      tc::tmem_ld_wait();
#endif
      AB_TRACE(6);
#pragma unroll
      for (int kk = 0; kk < kChunks; ++kk) {
        const int k = k0 + kk;  // chunk = K-slice index inside the 64-query sub-tile
        uint32_t(&sc)[16] = s[kk & 1];
        uint32_t(&dc)[16] = dp[kk & 1];
#if !(AB_KNOCK & 32)
        if (kk + 1 < kChunks) {
          tc::tmem_ld_x16(tst + C::kColST + (k + 1) * 16, s[(kk + 1) & 1]);
          tc::tmem_ld_x16(tst + C::kColDPT + (k + 1) * 16, dp[(kk + 1) & 1]);
        }
#endif
        uint32_t pk[8], dk[8];
        const float4* l4 = reinterpret_cast<const float4*>(stat + k * 16);
        const float4* d4 = reinterpret_cast<const float4*>(stat + 128 + k * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 lv = l4[q], dv = d4[q];
          const uint64_t xa = tc::fma2(tc::pack2(__uint_as_float(sc[4 * q]), __uint_as_float(sc[4 * q + 1])), sc2,
                                       tc::pack2(lv.x, lv.y));
          const uint64_t xb = tc::fma2(tc::pack2(__uint_as_float(sc[4 * q + 2]), __uint_as_float(sc[4 * q + 3])), sc2,
                                       tc::pack2(lv.z, lv.w));
          float x0, x1, x2, x3;
          tc::unpack2(xa, x0, x1);
          tc::unpack2(xb, x2, x3);
#if AB_KNOCK & 1
          const float p0 = x0, p1 = x1, p2 = x2, p3 = x3;
#else
          // AB_POLY of every 4 pairs go through the FMA-pipe polynomial (tc::exp2_poly2, 7.6e-5 relative: 25x below the
          // bf16 rounding P receives) instead of the MUFU: a softmax warp alone on its scheduler is bound by the 8 clk per
          // MUFU.EX2 warp instruction, with the FMA pipe ~75 % idle
          float p0, p1, p2, p3;
          const bool a_poly = (AB_POLY == 4) || (AB_POLY == 3 && (q & 1));
          const bool b_poly = (AB_POLY >= 2) || (AB_POLY == 1 && (q & 1));
          if (a_poly) tc::exp2_poly2(xa, p0, p1); else { p0 = tc::fast_exp2(x0); p1 = tc::fast_exp2(x1); }
          if (b_poly) tc::exp2_poly2(xb, p2, p3); else { p2 = tc::fast_exp2(x2); p3 = tc::fast_exp2(x3); }
#endif
          const uint64_t da = tc::mul2(tc::pack2(p0, p1), tc::add2(tc::pack2(__uint_as_float(dc[4 * q]), __uint_as_float(dc[4 * q + 1])),
                                                                   tc::pack2(dv.x, dv.y)));
          const uint64_t db = tc::mul2(tc::pack2(p2, p3), tc::add2(tc::pack2(__uint_as_float(dc[4 * q + 2]), __uint_as_float(dc[4 * q + 3])),
                                                                   tc::pack2(dv.z, dv.w)));
          float d0, d1, d2, d3;
          tc::unpack2(da, d0, d1);
          tc::unpack2(db, d2, d3);
          pk[2 * q] = pack_bf16x2(p0, p1);
          pk[2 * q + 1] = pack_bf16x2(p2, p3);
          dk[2 * q] = pack_bf16x2(d0, d1);
          dk[2 * q + 1] = pack_bf16x2(d2, d3);
        }
        // in place: the bf16 K-slice k goes over fp32 columns this thread has already consumed (slice_off(k) <= 16 k)
#if AB_KNOCK & 16
        if (pk[0] == 0x12345678u && dk[3] == 0x9abcdef0u) tc::tmem_st_x8(tst + C::kColST + C::slice_off(k), pk);
#else
        tc::tmem_st_x8(tst + C::kColST + C::slice_off(k), pk);    // P^T  (bf16 pairs)
        tc::tmem_st_x8(tst + C::kColDPT + C::slice_off(k), dk);   // dS^T (bf16 pairs)
#endif
        // dS^T row -> smem (MN-major A operand of dQ = dS K): buffer m&1, 64-query chunk hh, 16-byte pieces 2 k, 2 k + 1
#pragma unroll
        for (int q = 0; q < ((AB_KNOCK & 2) ? 0 : 2); ++q) {
          const int piece = k * 2 + q;
          *reinterpret_cast<uint4*>(rowp + ((piece ^ (row & 7)) << 4)) = make_uint4(dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
        }
#if !(AB_KNOCK & 32)
        if (kk + 1 < kChunks) tc::tmem_ld_wait();
#endif
      }
      AB_TRACE(7);
      tc::tmem_st_wait();
      if (!(AB_KNOCK & 1024)) tc::fence_proxy_async();  // st.shared (generic proxy) -> tcgen05.mma reads (async proxy)
      tc::tcgen05_fence_before();
      // one arrival per warp: 256 per-thread arrivals on one mbarrier serialise in the shared-memory pipe
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&p_ready[st]);
        tc::mbar_arrive(&ds_ready[m & 1]);
        if (hh == 0 && i == n_sub - 1) tc::mbar_arrive(&ds_ready[m & 1]);  // odd tail: the tile has no second sub-tile
      }
      AB_TRACE(8);
      AB_TRACEW(1);
    }
#if OCT_AB_TRACE
    if ((p.dbg & 16) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
      for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 48; ++k) printf("TRACE i%d id%d %lld\n", i + 8, k, g_ab_trace[i * 48 + k] - g_ab_trace[0]);
    }
#endif
    // epilogue: column half 0 stores dV, half 1 stores dK of this kv tile
    tc::mbar_wait(acc_full, 0);
    tc::tcgen05_fence_after();
    const int kv = n0 + row;
    const uint32_t col = colhalf ? C::kColDK : C::kColDV;
    const float sc = colhalf ? p.scale : 1.f;
    __nv_bfloat16* dst = p.dqkv + ((((size_t)b * p.S + kv) * 3 + (colhalf ? 1 : 2)) * p.H + h) * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[32];
      tc::tmem_ld_x32(lane_addr + col + c * 32, o);
      tc::tmem_ld_wait();
      if (kv_ok) {
#pragma unroll
        for (int q = 0; q < 32; q += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[q]) * sc, __uint_as_float(o[q + 1]) * sc);
          v.y = pack_bf16x2(__uint_as_float(o[q + 2]) * sc, __uint_as_float(o[q + 3]) * sc);
          v.z = pack_bf16x2(__uint_as_float(o[q + 4]) * sc, __uint_as_float(o[q + 5]) * sc);
          v.w = pack_bf16x2(__uint_as_float(o[q + 6]) * sc, __uint_as_float(o[q + 7]) * sc);
          *reinterpret_cast<uint4*>(dst + c * 32 + q) = v;
        }
      }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// delta[b,h,s] = sum_c dO[b,s,h,c] * O[b,s,h,c]   (one thread per (b,s,h) row: HD bf16 = 64 / 128 contiguous bytes), stored
// NEGATED next to -lse * log2(e) in [B,H,Spad] arrays (the form the softmax warps of attn_bwd_tc_kernel consume, fetched per
// 128-query tile by bulk copies); the rows s in [S, Spad) get -inf / 0 so that padded query columns produce P = dS = 0.
// The same thread clears the row's slot of the fp32 dQ accumulator (the drain warps of attn_bwd_tc_kernel reduce into it): a
// separate memset was one more launch per layer (32 per step).  Rows s >= S of the padded accumulator are never read.
template <int HD>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                  const float* __restrict__ lse, float* __restrict__ nlse2, float* __restrict__ ndelta,
                                  float* __restrict__ dq_acc, int64_t total, int S, int H, int Spad) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread per (b, h, s) with s FASTEST (over the padded length): the statistics and the accumulator slots of a warp are
  // contiguous; the HD-element rows it reads are H * HD elements apart but each is a whole number of sectors.  (With h fastest
  // — the first version of this kernel — every warp touched 32 different lines per statistics access and scattered 16-byte
  // zeroing stores over 32 accumulator tiles: 40 us instead of 15 us at the decoder shape.)
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (bb * H + hh) * Spad + s
  if (w >= total) return;
  const int s = (int)(w % Spad);
  const int64_t bh = w / Spad;
  if (s >= S) {  // padded query rows: P = dS = 0
    ndelta[w] = 0.f;
    nlse2[w] = -INFINITY;
    return;
  }
  const int hh = (int)(bh % H);
  const int64_t bb = bh / H;
  const int64_t row = ((bb * S + s) * H + hh) * HD;
  const uint4* a = reinterpret_cast<const uint4*>(out + row);
  const uint4* d = reinterpret_cast<const uint4*>(dout + row);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const uint4 x = a[i], y = d[i];
    const float2 x0 = unpack_bf16x2(x.x), x1 = unpack_bf16x2(x.y), x2 = unpack_bf16x2(x.z), x3 = unpack_bf16x2(x.w);
    const float2 y0 = unpack_bf16x2(y.x), y1 = unpack_bf16x2(y.y), y2 = unpack_bf16x2(y.z), y3 = unpack_bf16x2(y.w);
    acc += (x0.x * y0.x + x0.y * y0.y) + (x1.x * y1.x + x1.y * y1.y) + (x2.x * y2.x + x2.y * y2.y) +
           (x3.x * y3.x + x3.y * y3.y);
  }
  ndelta[w] = -acc;
  nlse2[w] = -lse[bh * S + s] * kLog2e;
  // accumulator layout per 128-query tile: [row quarter][16-byte chunk][32 rows][4 floats] (see the drain warps)
  float* tile = dq_acc + (bh * Spad + (s & ~127)) * HD;
  const int rr = s & 127;
#pragma unroll
  for (int c4 = 0; c4 < HD / 4; ++c4)
    *reinterpret_cast<float4*>(tile + (((rr >> 5) * (HD / 4) + c4) * 32 + (rr & 31)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// dq (bf16, into dqkv[:, :, 0]) = scale * dq_acc.  One block per 128-query tile of one (batch, head): the accumulator tile
// ([row quarter][16-byte chunk][32 rows][4 floats], see the drain warps of attn_bwd_tc_kernel) is read contiguously, transposed
// through shared memory, and written as whole HD-element rows (with one thread per 16-byte piece of the OUTPUT every read
// was a half-used sector 512 bytes from the next one).
template <int HD>
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv,
                                                              int S, int H, int Spad, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kPitch = HD + 4;  // floats per staged row: float4 stores of a quarter-warp land in 32 distinct banks
  __shared__ float tile[128 * kPitch];
  const int mt = blockIdx.x;
  const int64_t bh = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(dq_acc + (bh * Spad + (int64_t)mt * 128) * HD);
  for (int i = threadIdx.x; i < 128 * HD / 4; i += 256) {
    const int r = i & 31, c4 = (i >> 5) % (HD / 4), q = i / (32 * (HD / 4));
    *reinterpret_cast<float4*>(&tile[(q * 32 + r) * kPitch + c4 * 4]) = src[i];
  }
  __syncthreads();
  const int hh = (int)(bh % H);
  const int64_t bb = bh / H;
  for (int j = threadIdx.x; j < 128 * HD / 8; j += 256) {
    const int row = j / (HD / 8), c8 = j % (HD / 8);
    const int s = mt * 128 + row;
    if (s >= S) continue;
    const float4 v0 = *reinterpret_cast<const float4*>(&tile[row * kPitch + c8 * 8]);
    const float4 v1 = *reinterpret_cast<const float4*>(&tile[row * kPitch + c8 * 8 + 4]);
    uint4 o;
    o.x = pack_bf16x2(v0.x * scale, v0.y * scale);
    o.y = pack_bf16x2(v0.z * scale, v0.w * scale);
    o.z = pack_bf16x2(v1.x * scale, v1.y * scale);
    o.w = pack_bf16x2(v1.z * scale, v1.w * scale);
    *reinterpret_cast<uint4*>(dqkv + (((bb * S + s) * 3 + 0) * H + hh) * HD + c8 * 8) = o;
  }
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
  float* nlse2 = dq_acc + (size_t)B * H * Spad * HD;
  float* ndelta = nlse2 + (size_t)B * H * Spad;
  CUtensorMap mq, md;
  int rc = make_map4(&mq, qkv, HD, 3 * H, S, B, "oct_attn_bwd(bf16) qkv");
  if (rc) return rc;
  rc = make_map4(&md, dout, HD, H, S, B, "oct_attn_bwd(bf16) dout");
  if (rc) return rc;
  cudaError_t e;
  const int64_t rows = B * H * Spad;
  oct_launch(attn_delta_kernel<HD>, dim3((unsigned)ceil_div64(rows, 256)), dim3(256), 0, st, 1, (const __nv_bfloat16*)out,
             (const __nv_bfloat16*)dout, lse, nlse2, ndelta, dq_acc, rows, (int)S, (int)H, (int)Spad);
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
  p.nlse2 = nlse2; p.ndelta = ndelta; p.dq_acc = dq_acc; p.dqkv = (__nv_bfloat16*)dqkv;
  { const char* e = getenv("OCT_ATTN_BWD_DBG"); p.dbg = e ? atoi(e) : 0; }
  dim3 grid((unsigned)ceil_div64(S, AB_T), (unsigned)H, (unsigned)B);
  oct_launch(attn_bwd_tc_kernel<HD>, grid, dim3(AB_THREADS), (size_t)C::kSmem, st, 1, mq, md, p);
  rc = oct_check_launch("oct_attn_bwd(bf16)");
  if (rc) return rc;
  oct_launch(attn_dq_convert_kernel<HD>, dim3((unsigned)(Spad / 128), (unsigned)(B * H)), dim3(256), 0, st, 1, (const float*)dq_acc,
             (__nv_bfloat16*)dqkv, (int)S, (int)H, (int)Spad, scale);
  return oct_check_launch("oct_attn_bwd(bf16,dq)");
}

}  // namespace

size_t oct_attn_bwd_tc_ws_bytes(int64_t B, int64_t S, int64_t H, int64_t d) {
  const int64_t Spad = ceil_div64(S, 128) * 128;
  return (size_t)B * H * Spad * d * sizeof(float) + 2 * (size_t)B * H * Spad * sizeof(float) + 64;
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
