// Gradient all-reduce over NVLink 5 / NVSwitch without a collective library (replaces the NCCL all-reduce behind
// DistributedDataParallel, main_pretrain...:435-439): the gradient buckets live in SYMMETRIC memory (one allocation per rank,
// every rank's copy mapped into every process, plus — when the fabric supports it — ONE multicast address that maps all of them).
//
//   multicast path (NVLS):  rank r owns 1/W of the range.  For each 16-byte vector of its shard
//        v = multimem.ld_reduce.add.f32 [mc + off]        the SWITCH reads all W copies and adds them
//        multimem.st [mc + off], v * scale                and broadcasts the sum into all W copies
//      — per GPU (W-1)/W of the bytes cross the link once in and once out, and no SM ever touches a peer's data twice;
//   peer path (no multicast): the same two-shot schedule with plain peer loads (W reads per vector) and W peer stores.
//
// Why not NCCL: its kernels take whole SMs (channels x 512 threads with tens of KB of shared memory each) away from the
// persistent one-CTA-per-SM GEMM / attention kernels of the backward pass they overlap — 2.8 ms of a 34.9 ms step at N = 8
// (DESIGN.md §7).  This kernel uses no shared memory and 32 registers per thread, so its CTAs CO-RESIDE with the 200 KB GEMM
// CTAs: it costs issue slots and memory bandwidth, not SMs.
//
// Ranks synchronise through epoch flags in the symmetric allocation (release / acquire at system scope): "my gradients of this
// range are final" before anybody reduces, "my shard is reduced and broadcast" before anybody reads.  The epoch is device
// resident, so the launch replays from a CUDA graph.  Every rank must issue the same sequence of calls.
#include "common.cuh"

namespace {

constexpr int kArMaxWorld = 16;
constexpr int kArThreads = 128;   // 128 threads x <= 48 registers: the CTA fits NEXT TO a 320-thread x 165-register GEMM CTA on the same SM

struct ArPeers {
  float* buf[kArMaxWorld];        // rank s's buffer (peer path: data; both paths: flag block at flag_off)
};

__device__ __forceinline__ unsigned ar_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ar_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// flags (32-bit words at `flags` of every rank's buffer): [slot][0..15] ready epochs, [slot][16..31] done epochs, one slot
// per concurrently usable channel (bucket index % kSlots).  state (local): [slot] = {epoch, cta counter}.
constexpr int kSlots = 8;

template <bool kMulticast>
__global__ void __launch_bounds__(kArThreads) allreduce_kernel(float* __restrict__ mc, ArPeers peers, size_t flag_off_words,
                                                               unsigned* __restrict__ state, size_t off_elems, size_t n_elems,
                                                               int rank, int W, int slot, float scale) {
  unsigned* st = state + slot * 4;
  const unsigned e = st[0] + 1;
  unsigned* my_flags = reinterpret_cast<unsigned*>(peers.buf[rank]) + flag_off_words + slot * 32;
  // ---- 1. every rank's gradients of this range are final (their producing kernels precede this one in stream order)
  if (blockIdx.x == 0 && threadIdx.x < W)
    ar_st_release(reinterpret_cast<unsigned*>(peers.buf[threadIdx.x]) + flag_off_words + slot * 32 + rank, e);
  if (threadIdx.x < W) {
    unsigned spins = 0;
    while ((int)(ar_ld_acquire(my_flags + threadIdx.x) - e) < 0) {
      __nanosleep(128);
      if (++spins > (1u << 26)) { atomicExch(st + 2, 1u); break; }   // ~10 s: a peer never arrived; flagged, not hung
    }
  }
  __syncthreads();
  // ---- 2. my shard: reduce over the ranks, scale, broadcast
  const size_t n4 = n_elems >> 2;
  const size_t per = (n4 + W - 1) / W;
  const size_t lo = (size_t)rank * per, hi = min(n4, lo + per);
  const size_t stride = (size_t)gridDim.x * kArThreads;
  if (kMulticast) {
    float* base = mc + off_elems;
    size_t i = lo + (size_t)blockIdx.x * kArThreads + threadIdx.x;
    for (; i + 3 * stride < hi; i += 4 * stride) {   // four independent vectors in flight per thread
      float4 v0 = mc_ld_reduce(base + 4 * i), v1 = mc_ld_reduce(base + 4 * (i + stride));
      float4 v2 = mc_ld_reduce(base + 4 * (i + 2 * stride)), v3 = mc_ld_reduce(base + 4 * (i + 3 * stride));
      v0.x *= scale; v0.y *= scale; v0.z *= scale; v0.w *= scale;
      v1.x *= scale; v1.y *= scale; v1.z *= scale; v1.w *= scale;
      v2.x *= scale; v2.y *= scale; v2.z *= scale; v2.w *= scale;
      v3.x *= scale; v3.y *= scale; v3.z *= scale; v3.w *= scale;
      mc_st(base + 4 * i, v0); mc_st(base + 4 * (i + stride), v1);
      mc_st(base + 4 * (i + 2 * stride), v2); mc_st(base + 4 * (i + 3 * stride), v3);
    }
    for (; i < hi; i += stride) {
      float4 v = mc_ld_reduce(base + 4 * i);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      mc_st(base + 4 * i, v);
    }
  } else {
    for (size_t i = lo + (size_t)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < W; ++s) {   // fixed order: every rank computes a shard exactly once, all copies end up identical
        const float4 v = __ldcg(reinterpret_cast<const float4*>(peers.buf[s] + off_elems) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
      for (int s = 0; s < W; ++s) __stcg(reinterpret_cast<float4*>(peers.buf[s] + off_elems) + i, acc);
    }
  }
  // ---- 3. my shard is complete in every copy -> tell everybody; wait until everybody's shard is complete in MY copy
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(st + 1, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < W)
      ar_st_release(reinterpret_cast<unsigned*>(peers.buf[threadIdx.x]) + flag_off_words + slot * 32 + 16 + rank, e);
  }
  if (threadIdx.x < W) {
    unsigned spins = 0;
    while ((int)(ar_ld_acquire(my_flags + 16 + threadIdx.x) - e) < 0) {
      __nanosleep(128);
      if (++spins > (1u << 26)) { atomicExch(st + 2, 1u); break; }
    }
  }
  __syncthreads();
  // the epoch advances once every CTA has passed the last wait (a second counter pass)
  if (threadIdx.x == 0) {
    if (atomicAdd(st + 3, 1u) == gridDim.x - 1) {
      st[1] = 0; st[3] = 0;
      __threadfence();
      st[0] = e;
    }
  }
}

}  // namespace

extern "C" size_t oct_allreduce_flag_bytes(void) { return (size_t)kSlots * 32 * 4; }
extern "C" size_t oct_allreduce_state_bytes(void) { return (size_t)kSlots * 4 * 4; }

// In-place SUM all-reduce (times `scale`) of `n_elems` fp32 values at element offset `off_elems` of a symmetric buffer.
//   mc_ptr     multicast address of the allocation (NULL: peer path)
//   peer_bufs  HOST array [world] of device pointers: rank s's copy of the allocation as this process addresses it
//   flag_off_bytes  offset of the flag block (oct_allreduce_flag_bytes(), zero-initialised) inside every rank's allocation
//   state      local device memory, oct_allreduce_state_bytes(), zero-initialised; word 2 of slot s is raised on a time-out
//   slot       flag / epoch channel (0..7): calls that may be in flight at the same time must use different slots
extern "C" int oct_allreduce_sym(void* mc_ptr, const void* const* peer_bufs, int64_t flag_off_bytes, void* state,
                                 int64_t off_elems, int64_t n_elems, int rank, int world, int slot, float scale, int ctas,
                                 oct_stream_t stream) {
  OCT_REQUIRE(peer_bufs && state, "oct_allreduce_sym: null pointer");
  OCT_REQUIRE(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world, "oct_allreduce_sym: bad rank / world (<= %d)", kArMaxWorld);
  OCT_REQUIRE(slot >= 0 && slot < kSlots, "oct_allreduce_sym: slot out of range");
  OCT_REQUIRE(off_elems >= 0 && n_elems >= 0 && off_elems % 4 == 0 && n_elems % 4 == 0 && flag_off_bytes % 16 == 0,
              "oct_allreduce_sym: offsets and counts must be multiples of 4 elements");
  ArPeers peers;
  for (int s = 0; s < kArMaxWorld; ++s) {
    peers.buf[s] = s < world ? (float*)peer_bufs[s] : nullptr;
    OCT_REQUIRE(s >= world || (peers.buf[s] && aligned16(peers.buf[s])), "oct_allreduce_sym: peer buffer %d null or unaligned", s);
  }
  if (n_elems == 0) return OCT_OK;
  if (ctas <= 0) ctas = 96;
  const int64_t need = ceil_div64(ceil_div64(n_elems / 4, world), kArThreads);
  if (ctas > need) ctas = (int)(need < 1 ? 1 : need);
  if (ctas > 296) ctas = 296;   // every CTA must become resident while the others spin on the flags
  cudaStream_t st = (cudaStream_t)stream;
  if (mc_ptr)
    allreduce_kernel<true><<<ctas, kArThreads, 0, st>>>((float*)mc_ptr, peers, (size_t)flag_off_bytes / 4, (unsigned*)state,
                                                        (size_t)off_elems, (size_t)n_elems, rank, world, slot, scale);
  else
    allreduce_kernel<false><<<ctas, kArThreads, 0, st>>>(nullptr, peers, (size_t)flag_off_bytes / 4, (unsigned*)state,
                                                         (size_t)off_elems, (size_t)n_elems, rank, world, slot, scale);
  return oct_check_launch("oct_allreduce_sym");
}
