// Micro-benchmark: throughput of ex2.approx f32 vs packed bf16x2 / f16x2 (results per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  uint32_t r[8];
  for (int i = 0; i < 8; ++i) r[i] = 0xBC00BC00u + threadIdx.x + i;  // small negative values
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { float f = __uint_as_float(r[i]); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); r[i] = __float_as_uint(f) | 0x80000000u; }
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 4) { float f = __uint_as_float(r[i]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); r[i] = __float_as_uint(f); }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (uint32_t)(t1 - t0);
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  const char* names[5] = {"ex2.f32", "ex2.bf16x2", "ex2.f16x2", "tanh.bf16x2", "tanh.f32"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 1024>>>(d, iters);
      if (mode == 1) k<1><<<148, 1024>>>(d, iters);
      if (mode == 2) k<2><<<148, 1024>>>(d, iters);
      if (mode == 3) k<3><<<148, 1024>>>(d, iters);
      if (mode == 4) k<4><<<148, 1024>>>(d, iters);
      cudaDeviceSynchronize();
    }
    uint32_t clk; cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
    double instr = 1024.0 * iters * 8;  // thread-instructions per SM
    printf("%-12s %8u clk  %.2f thread-instr/clk/SM  (%s results/clk/SM: %.2f)\n", names[mode], clk, instr / clk,
           (mode == 0 || mode == 4) ? "1x" : "2x", instr / clk * ((mode == 0 || mode == 4) ? 1 : 2));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
