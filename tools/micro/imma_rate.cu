// Microbenchmark: issue rate of mma.sync m16n8k32 u8 (IMMA.16832) and dp4a on sm_100a, per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_rate imma_rate.cu && ./imma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_imma(int* out, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  int c[8][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  int s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (int)(t1 - t0);
}
__global__ void k_dp4a(int* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u, b = a * 3;
  int c[8] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(c[j]) : "r"(a + j), "r"(b));
  }
  long long t1 = clock64();
  int s = 0;
  for (int j = 0; j < 8; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (int)(t1 - t0);
}
int main() {
  int* d; cudaMalloc(&d, 1 << 22);
  for (int warps : {1, 4, 8, 16}) {
    const int iters = 2000;
    k_imma<<<148, warps * 32>>>(d, iters); cudaDeviceSynchronize();
    k_imma<<<148, warps * 32>>>(d, iters); cudaDeviceSynchronize();
    int clk; cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
    printf("IMMA.16832: %2d warps/SM: %.2f clk per IMMA per warp, %.1f IMMA/clk/SM -> %.0f int8 MAC/clk/SM\n", warps, clk / (8.0 * iters),
           warps * 8.0 * iters / clk, warps * 8.0 * iters / clk * 4096);
    k_dp4a<<<148, warps * 32>>>(d, iters); cudaDeviceSynchronize();
    cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
    printf("dp4a      : %2d warps/SM: %.2f clk per dp4a per warp, %.2f warp-dp4a/clk/SM -> %.0f int8 MAC/clk/SM\n", warps, clk / (8.0 * iters),
           warps * 8.0 * iters / clk, warps * 8.0 * iters / clk * 128);
  }
  return 0;
}
