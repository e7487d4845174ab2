// Micro-benchmark (perf experiment): MUFU ex2 throughput for f32 / f16x2 / bf16x2 on sm_100a, results per clock per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 8 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (uint32_t)(t1 - t0);
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps = 4; warps <= 16; warps *= 2) {
      uint32_t cyc = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 0x3c003c00u);
        if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, 0x3c003c00u);
        if (mode == 2) k<2><<<148, warps * 32>>>(d, iters, 0x3f803f80u);
        cudaDeviceSynchronize();
        cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
      }
      double ops = (double)iters * 8 * warps * 32;   // instructions x lanes per SM
      const char* names[] = {"f32", "f16x2", "bf16x2"};
      printf("%-7s warps/SM=%2d  cycles=%u  lane-instr/clk/SM=%.2f  results/clk/SM=%.2f  err=%s\n", names[mode], warps, cyc,
             ops / cyc, ops / cyc * (mode == 0 ? 1 : 2), cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
