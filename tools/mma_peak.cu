// Microbenchmark: legacy mma.sync.m16n8k16 bf16 throughput on sm_100a (the pipe k_fb3.cu runs on).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mma_peak tools/mma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k(float *out, int iters) {
  float acc[8][4] = {};
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b0 = threadIdx.x, b1 = 5u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int n = 0; n < 8; ++n)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[n][0]), "+f"(acc[n][1]), "+f"(acc[n][2]), "+f"(acc[n][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int n = 0; n < 8; ++n) s += acc[n][0] + acc[n][1] + acc[n][2] + acc[n][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float *out; cudaMalloc(&out, 1 << 24);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000;
    k<<<p.multiProcessorCount, warps * 32>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<p.multiProcessorCount, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)p.multiProcessorCount * warps * iters * 8;
    const double clk = ms * 1e-3 * p.clockRate * 1e3;
    printf("%2d warps/SM: %.3f ms  %.2f clk per mma per SM-subpartition  %.0f TFLOP/s dense bf16\n", warps, ms,
           clk / (mmas / p.multiProcessorCount / 4), mmas * 2 * 16 * 8 * 16 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
