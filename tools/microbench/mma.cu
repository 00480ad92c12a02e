// tcgen05.mma issue / execution time per instruction as a function of N, operand source and layout (sm_100a).
// One CTA per SM; one elected thread issues kCount MMAs of one kind back to back into one TMEM accumulator, commits, and the
// CTA waits for the commit barrier: (t_done - t_start) / kCount = sustained cycles per MMA of the in-order tensor pipe;
// (t_issued - t_start) / kCount = cycles the ISSUING THREAD spends per MMA.  Variants: SS K-major A (128x16 from smem),
// SS MN-major A, TS (A from TMEM); N in {32, 64, 128, 256}; optionally two warps issuing alternately into two accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../octcubem_b200/csrc mma.cu -o mma
#include <cstdio>
#include "tc_common.cuh"

constexpr int kCount = 512;

template <int N, int MODE, int ISSUERS>  // MODE 0: SS K-major A/B, 1: SS MN-major A + MN-major B, 2: TS (A in TMEM), B MN-major
__global__ void __launch_bounds__(128) k(unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&slot);
  tc::fence_proxy_async();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t a_addr = tc::smem_u32(smem), b_addr = tc::smem_u32(smem + 32 * 1024);
  constexpr uint32_t idesc = tc::make_idesc(tc::kFmtBF16, MODE == 1, MODE != 0, 128, N);
  unsigned long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < ISSUERS) {
    t0 = clock64();
    if (tc::elect_one()) {
#pragma unroll 8
      for (int i = 0; i < kCount / ISSUERS; ++i) {
        const uint32_t koff = (i & 3);
        const uint64_t da = MODE == 1 ? tc::make_smem_desc(a_addr + koff * 2048, 64 * 128, 1024) : tc::make_smem_desc(a_addr + koff * 32, 16, 1024);
        const uint64_t db = MODE != 0 ? tc::make_smem_desc(b_addr + koff * 2048, 64 * 128, 1024) : tc::make_smem_desc(b_addr + koff * 32, 16, 1024);
        if (MODE == 2) tc::mma_ts(tmem + warp * 256, tmem + 448 + koff * 8, db, idesc, 1);
        else tc::mma_ss(tmem + warp * 256, da, db, idesc, 1);
      }
      tc::mma_commit(&bar[warp]);
    }
    __syncwarp();
    t1 = clock64();
    tc::mbar_wait(&bar[warp], 0);
    t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tcgen05_fence_after(); tc::tmem_dealloc<512>(tmem); }
}

template <int N, int MODE, int ISSUERS>
void run(const char* name) {
  unsigned long long* out;
  cudaMalloc(&out, 64);
  cudaMemset(out, 0, 64);
  cudaFuncSetAttribute(k<N, MODE, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  k<N, MODE, ISSUERS><<<148, 128, 66 * 1024>>>(out);
  k<N, MODE, ISSUERS><<<148, 128, 66 * 1024>>>(out);
  cudaDeviceSynchronize();
  unsigned long long h[4];
  cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  const double per = (double)(ISSUERS == 2 ? (h[1] > h[3] ? h[1] : h[3]) : h[1]) / kCount;
  printf("%-34s N=%3d issuers=%d: %6.1f clk/MMA sustained (%5.1f %% of the 8192 flop/clk/SM rate), issuing thread %6.1f clk/MMA  (%s)\n", name, N,
         ISSUERS, per, 100.0 * (2.0 * 128 * N * 16 / per) / 8192.0, (double)h[0] / (kCount / ISSUERS), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<32, 0, 1>("SS  A K-major  (128x16 smem)");  run<64, 0, 1>("SS  A K-major  (128x16 smem)");
  run<128, 0, 1>("SS  A K-major  (128x16 smem)"); run<256, 0, 1>("SS  A K-major  (128x16 smem)");
  run<32, 1, 1>("SS  A MN-major, B MN-major");    run<64, 1, 1>("SS  A MN-major, B MN-major");
  run<128, 1, 1>("SS  A MN-major, B MN-major");
  run<32, 2, 1>("TS  A in TMEM, B MN-major");     run<64, 2, 1>("TS  A in TMEM, B MN-major");
  run<128, 2, 1>("TS  A in TMEM, B MN-major");
  run<32, 0, 2>("SS  A K-major, two issuing warps"); run<32, 2, 2>("TS, two issuing warps"); run<64, 0, 2>("SS  A K-major, two issuing warps");
  return 0;
}
