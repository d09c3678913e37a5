// pipe_peaks2.cu -- packed fp32x2 (FFMA2 / FADD2 / FMUL2) issue rates on sm_100a, alone and
// mixed with MUFU.EX2, to decide whether the mixture sweeps should use packed arithmetic.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

template <int OP>
__global__ void k(float *out, int iters, float seed) {
  unsigned long long v[8], b = pack(0.999f, 0.998f), c = pack(0.001f, 0.002f);
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = pack(seed + i * 0.001f, seed + threadIdx.x * 1e-6f); f[i] = seed + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) v[i] = fma2(v[i], b, c);
      if (OP == 1) v[i] = add2(v[i], c);
      if (OP == 2) v[i] = mul2(v[i], b);
      if (OP == 3) v[i] = fma2(v[i], v[(i + 1) & 7], v[(i + 2) & 7]);
      if (OP == 4) {  // 1 MUFU + 4 packed ops (8 fp32 lane-ops)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        v[i] = fma2(v[i], b, c); v[i] = mul2(v[i], b); v[i] = add2(v[i], c); v[i] = fma2(v[i], b, pack(f[i], f[i]));
      }
      if (OP == 5) {  // 1 MUFU + 8 scalar fp32 ops
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        float a0 = __uint_as_float((unsigned)v[i]), a1 = __uint_as_float((unsigned)(v[i] >> 32));
#pragma unroll
        for (int q = 0; q < 4; ++q) { a0 = fmaf(a0, 0.999f, f[i]); a1 = fmaf(a1, 0.998f, 0.002f); }
        v[i] = pack(a0, a1);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32)) + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, double instr_per_iter, int sms, double mhz) {
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
  const double groups = (double)blocks * threads * iters * 8.0;   // loop bodies (per lane)
  const double per_clk_sm = groups / (ms * 1e-3) / (mhz * 1e6) / sms;
  printf("%-44s %8.3f ms  %7.2f bodies/clk/SM  -> %7.2f instr-lanes/clk/SM\n", name, ms, per_clk_sm, per_clk_sm * instr_per_iter);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double mhz = p.clockRate / 1e3;
  printf("%s, %d SMs, clockRate %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  run<0>("FFMA2 (const operands)  [1 instr, 2 fma]", 1, p.multiProcessorCount, mhz);
  run<1>("FADD2                   [1 instr, 2 add]", 1, p.multiProcessorCount, mhz);
  run<2>("FMUL2                   [1 instr, 2 mul]", 1, p.multiProcessorCount, mhz);
  run<3>("FFMA2 (3 register pairs)[1 instr, 2 fma]", 1, p.multiProcessorCount, mhz);
  run<4>("MUFU + 4 packed ops     [5 instr]", 5, p.multiProcessorCount, mhz);
  run<5>("MUFU + 8 scalar FFMA    [9 instr]", 9, p.multiProcessorCount, mhz);
  return 0;
}
