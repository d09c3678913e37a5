// pipe_peaks.cu -- measures the FP32-FMA and MUFU (ex2 / lg2 / rcp) issue rates of the GPU it
// runs on, in lane-operations per clock per SM.  These are the roofline denominators of the
// mixture sweeps (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_peaks pipe_peaks.cu && ./pipe_peaks
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float *out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) v[i] = fmaf(v[i], 0.999f, 0.001f);
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); v[i] = fmaf(v[i], 0.5f, -1.0f); v[i] = fmaf(v[i], v[i], 0.25f); v[i] = fmaf(v[i], 0.5f, -0.5f);}
      if (OP == 5) v[i] = fmaf(v[i], v[(i + 1) & 15], v[(i + 2) & 15]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(const char *name, int ops_per_iter, int sms, double mhz) {
  float *out;
  const int threads = 1024, blocks = sms * 2, iters = 4096;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  k<OP><<<blocks, threads>>>(out, 64, 0.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double lane_ops = (double)blocks * threads * iters * 16.0 * ops_per_iter;
  const double per_clk_sm = lane_ops / (ms * 1e-3) / (mhz * 1e6) / sms;
  printf("%-28s %8.3f ms  %7.2f lane-ops/clk/SM at %.0f MHz  (%.2f Tops/s)\n", name, ms, per_clk_sm, mhz, lane_ops / (ms * 1e-3) / 1e12);
  cudaFree(out);
  return per_clk_sm;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double mhz = p.clockRate / 1e3;
  printf("%s, %d SMs, clockRate %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  run<0>("FFMA (imm operands)", 1, p.multiProcessorCount, mhz);
  run<5>("FFMA (3 registers)", 1, p.multiProcessorCount, mhz);
  run<1>("MUFU.EX2", 1, p.multiProcessorCount, mhz);
  run<2>("MUFU.LG2", 1, p.multiProcessorCount, mhz);
  run<3>("MUFU.RCP", 1, p.multiProcessorCount, mhz);
  run<4>("EX2 + 3 FFMA (ex2 counted)", 1, p.multiProcessorCount, mhz);
  return 0;
}
