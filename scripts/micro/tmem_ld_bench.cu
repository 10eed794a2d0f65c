// Micro-benchmark: TMEM read bandwidth (tcgen05.ld.32x32b.x32) per SM as a function of the number of reading warps.
// One CTA per SM allocates all 512 columns; warp w reads the 32 lanes of quarter (w & 3), 32 columns (4 KB) per
// instruction, walking the 512 columns; `depth` loads are in flight before each tcgen05.wait::ld.
// Build on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_bench scripts/micro/tmem_ld_bench.cu
// Question it answers: the split d = 40 attention reads one 128 x 128 fp32 S tile (64 KB) per query tile and key tile out
// of TMEM; the GEMM epilogue a 128 x BN fp32 accumulator.  How many clocks is that, at best?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

template <int DEPTH>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  __shared__ long long tmax;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) tmax = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t v[DEPTH][32];
    uint32_t col = (warp >> 2) * 32;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        ld32(base + (col & 511), v[d]);
        col += 32 * ((nwarps + 3) / 4);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc ^= v[d][0] ^ v[d][31];
    }
    t1 = clock64();
    atomicMax(reinterpret_cast<unsigned long long*>(&tmax), static_cast<unsigned long long>(t1 - t0));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
  out[blockIdx.x * blockDim.x + threadIdx.x + 2] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = static_cast<uint32_t>(tmax); }
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, (148 * 512 + 2) * 4);
  const int iters = 2048;
  const int warps[4] = {1, 4, 8, 16};
  for (int depth = 1; depth <= 4; depth *= 2) {
    for (int wi = 0; wi < 4; ++wi) {
      const int nw = warps[wi];
      for (int rep = 0; rep < 2; ++rep) {
        if (depth == 1) k<1><<<148, 512>>>(d, iters, nw);
        if (depth == 2) k<2><<<148, 512>>>(d, iters, nw);
        if (depth == 4) k<4><<<148, 512>>>(d, iters, nw);
        cudaDeviceSynchronize();
      }
      uint32_t clk;
      cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
      const double bytes = static_cast<double>(nw) * iters * depth * 4096.0;
      printf("warps %2d  loads in flight %d: %9u clk  %7.1f B/clk/SM  (%.1f clk per 4 KB load per warp; a 128 x 128 fp32 tile = %.0f clk)\n", nw, depth, clk,
             bytes / clk, static_cast<double>(clk) / (iters * depth), 65536.0 / (bytes / clk));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
