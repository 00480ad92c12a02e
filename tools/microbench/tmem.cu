// tcgen05.ld / tcgen05.st throughput per SM (sm_100a): one CTA per SM, W warps (4 or 8; warp w owns TMEM lane quarter w % 4),
// each issuing 32x32b.x32 loads (4 KB per warp instruction) / stores back to back over its 512 allocated columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../octcubem_b200/csrc tmem.cu -o tmem
#include <cstdio>
#include "tc_common.cuh"

constexpr int kIters = 2048;

template <int MODE>  // 0: ld x32, 1: st x32, 2: ld x32 + st x16 (the softmax pattern: read fp32 S, write bf16 P)
__global__ void __launch_bounds__(256) k(unsigned long long* clk, float* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&slot);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32], acc = 0;
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
    const uint32_t col = (it * 32) & 511 & ~31;
    if (MODE == 0 || MODE == 2) {
      tc::tmem_ld_x32(base + ((warp >> 2) * 0) + col, r);
      if ((it & 3) == 3) tc::tmem_ld_wait();
      acc += r[0] ^ r[31];
    }
    if (MODE == 1) {
      tc::tmem_st_x32(base + col, r);
      if ((it & 3) == 3) tc::tmem_st_wait();
    }
    if (MODE == 2) {
      uint32_t h[16];
      for (int i = 0; i < 16; ++i) h[i] = r[2 * i];
      tc::tmem_st_x16(base + (col >> 1), h);
      if ((it & 3) == 3) tc::tmem_st_wait();
    }
  }
  tc::tmem_ld_wait();
  tc::tmem_st_wait();
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tcgen05_fence_after(); tc::tmem_dealloc<512>(slot); }
}

template <int MODE>
void run(const char* name, int warps, double bytes_per_iter) {
  const int sms = 148;
  unsigned long long* clk; float* out;
  cudaMalloc(&clk, sms * 8); cudaMalloc(&out, sms * 256 * 4);
  k<MODE><<<sms, warps * 32>>>(clk, out);
  k<MODE><<<sms, warps * 32>>>(clk, out);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  printf("%-34s %d warps: %8.1f B/clk/SM   (%.0f clk per warp instruction;  %s)\n", name, warps, warps * kIters * bytes_per_iter / avg,
         avg / kIters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(clk); cudaFree(out);
}

int main() {
  for (int w : {1, 4, 8}) {
    run<0>("tcgen05.ld 32x32b.x32 (4 KB)", w, 4096);
    run<1>("tcgen05.st 32x32b.x32 (4 KB)", w, 4096);
    run<2>("ld x32 + st x16 (4 KB + 2 KB)", w, 6144);
  }
  return 0;
}
