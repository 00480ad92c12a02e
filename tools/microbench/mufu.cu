// Throughput of the exponential instruction variants on one SM-resident persistent grid (sm_100a):
//   ex2.approx.ftz.f32 (1 result / instr), ex2.approx.f16x2, ex2.approx.ftz.bf16x2 (2 results / instr), and the FMA-pipe
//   polynomial of tc_common.cuh for reference.  Prints results per clock per SM.   nvcc -arch=sm_100a -O3 mufu.cu -o mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

constexpr int kIters = 4096, kIlp = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, unsigned long long* clk) {
  float x[kIlp];
  unsigned u[kIlp];
  for (int i = 0; i < kIlp; ++i) { x[i] = -0.001f * (threadIdx.x + i + 1); u[i] = 0xB800B800u + threadIdx.x + i; }
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kIlp; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < kIlp; ++i) s += x[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  int sms = 148;
  float* out; unsigned long long* clk;
  cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&clk, sms * 8);
  k<MODE><<<sms, 1024>>>(out, clk);
  k<MODE><<<sms, 1024>>>(out, clk);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  const double instr = 1024.0 * kIters * kIlp;
  printf("%-28s %7.2f instr/clk/SM  %7.2f results/clk/SM   (%s)\n", name, instr / avg, instr * per_instr / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<4>("tanh.approx.f16x2", 2);
  run<5>("rcp.approx.ftz.f32", 1);
  return 0;
}
