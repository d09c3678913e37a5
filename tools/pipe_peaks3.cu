// pipe_peaks3.cu -- sm_100a issue rates that decide between the component-packed and the points-packed emission
// sweep: FFMA2 with scalar-broadcast operands (SASS R.F32), shared-memory loads next to MUFU (both sit behind
// the MIO queue), and the two inner bodies themselves.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fnma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\t.reg .b64 n;\n\tmov.b64 {lo, hi}, %1;\n\tneg.f32 lo, lo;\n\tneg.f32 hi, hi;\n\t"
      "mov.b64 n, {lo, hi};\n\tfma.rn.f32x2 %0, n, %2, %3;\n\t}" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 ex2_2(u64 a) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tex2.approx.ftz.f32 lo, lo;\n\tex2.approx.ftz.f32 hi, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}" : "=l"(d) : "l"(a));
  return d;
}

// OP 0: FFMA2 pair x scalar + scalar      OP 1: FFMA2 -pair x pair + scalar     OP 2: FFMA2 scalar x pair + pair
// OP 3: FFMA2 three pairs                 (OP 4 / 5: shared-memory loads alone -- not run: the compiler keeps them out of the loop)
// OP 6: 1 LDS.128 + 4 MUFU                OP 7: 4 MUFU alone
// OP 8: points-packed body per component (LDS.128, 4 FFMA2, 4 MUFU, 2 FADD2)
// OP 9: component-packed body per component pair (LDS.128 + LDS.64, 8 FFMA2, 8 MUFU, 4 FADD2)
template <int OP>
__global__ void k(float *out, int iters, float seed) {
  __shared__ float4 sm4[8 * 32 * 4 + 32];
  __shared__ float2 sm2[8 * 32 * 4 + 32];
  for (int t = threadIdx.x; t < 8 * 32 * 4 + 32; t += blockDim.x) {
    sm4[t] = make_float4(1e-3f * (t & 7), 2e-3f, -1.f - 1e-3f * (t & 3), 0.f);
    sm2[t] = make_float2(-1.f, -1.001f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 3;
  const float4 *p4 = sm4 + w * 256 + lane;
  const float2 *p2 = sm2 + w * 256 + lane;
  u64 v[8];
  float s1 = seed * 0.999f, s2 = seed * 1e-3f;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = pk(seed + i * 0.001f, seed + threadIdx.x * 1e-6f);
  u64 acc0 = 0ull, acc1 = 0ull, acc2 = 0ull, acc3 = 0ull;
  for (int it = 0; it < (OP >= 10 ? 0 : iters); ++it) {
    if (OP <= 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) v[i] = fma2(v[i], pk(s1, s1), pk(s2, s2));
        if (OP == 1) v[i] = fnma2(v[i], v[i], pk(s1, s1));
        if (OP == 2) v[i] = fma2(pk(s1, s1), v[i], v[(i + 1) & 7]);
        if (OP == 3) v[i] = fma2(v[i], v[(i + 1) & 7], v[(i + 2) & 7]);
      }
    } else if (OP == 4 || OP == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 4) { const float4 q = p4[i * 32]; acc0 = add2(acc0, pk(q.x, q.z)); }
        else { const float2 q = p2[i * 32]; acc0 = add2(acc0, pk(q.x, q.y)); }
      }
      p4 += (it & 1) ? -1 : 1; p2 += (it & 1) ? -1 : 1;     // (keeps the loads inside the loop)
    } else if (OP == 6 || OP == 7) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 6) { const float4 q = p4[i * 32]; acc0 = add2(acc0, pk(q.x, q.z)); }
        v[i] = ex2_2(v[i]);
        v[(i + 4) & 7] = ex2_2(v[(i + 4) & 7]);
      }
      p4 += (it & 1) ? -1 : 1;
    } else if (OP == 8) {
      const u64 x01 = v[0], x23 = v[1];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float4 q = p4[kk * 32];
        const u64 sb = pk(q.x, q.x), mb = pk(q.y, q.y), cb = pk(q.z, q.z);
        const u64 z01 = fma2(x01, sb, mb), z23 = fma2(x23, sb, mb);
        acc0 = add2(acc0, ex2_2(fnma2(z01, z01, cb)));
        acc1 = add2(acc1, ex2_2(fnma2(z23, z23, cb)));
      }
      v[0] = add2(v[0], pk(1e-6f, 1e-6f)); v[1] = add2(v[1], pk(1e-6f, 2e-6f));
    } else if (OP == 9) {
      float x[4];
      unpk(v[0], x[0], x[1]); unpk(v[1], x[2], x[3]);
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        const float4 q = p4[kp * 32];
        const float2 c = p2[kp * 32];
        const u64 s2_ = pk(q.x, q.y), m2 = pk(q.z, q.w), cc = pk(c.x, c.y);
        const u64 z0 = fma2(pk(x[0], x[0]), s2_, m2), z1 = fma2(pk(x[1], x[1]), s2_, m2);
        const u64 z2 = fma2(pk(x[2], x[2]), s2_, m2), z3 = fma2(pk(x[3], x[3]), s2_, m2);
        acc0 = add2(acc0, ex2_2(fnma2(z0, z0, cc)));
        acc1 = add2(acc1, ex2_2(fnma2(z1, z1, cc)));
        acc2 = add2(acc2, ex2_2(fnma2(z2, z2, cc)));
        acc3 = add2(acc3, ex2_2(fnma2(z3, z3, cc)));
      }
      v[0] = add2(v[0], pk(1e-6f, 1e-6f)); v[1] = add2(v[1], pk(1e-6f, 2e-6f));
    }
  }
  float s = 0.f, lo, hi;
  if (OP == 10 || OP == 11) {
    // statistics-sweep body, K = 8 (k_stats2 stage1 + stage2): per point 16 FFMA2 + 7 FADD2 + 4 FMUL2 + 9 MUFU + 3 scalar
    u64 s2[4], m2[4], c2[4], S0[4], S1[4], S2[4];
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      s2[kp] = pk(seed + kp, seed * 2 + kp); m2[kp] = pk(-seed, seed * 3); c2[kp] = pk(-1.f - kp, -2.f);
      S0[kp] = S1[kp] = S2[kp] = 0ull;
    }
    float x = seed, g = seed * 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int pnt = 0; pnt < 4; ++pnt) {
        const u64 xx = OP == 11 ? pk(x, x) : pk(x, x + 1e-7f * pnt);
        u64 z[4], ar[4], e[4];
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          z[kp] = fma2(xx, s2[kp], m2[kp]);
          ar[kp] = fnma2(z[kp], z[kp], c2[kp]);
          e[kp] = ex2_2(ar[kp]);
        }
        const u64 sum2 = add2(add2(e[0], e[1]), add2(e[2], e[3]));
        float sl, sh;
        unpk(sum2, sl, sh);
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaxf(sl + sh, 1e-38f)));
        inv *= g;
        const u64 inv2 = pk(inv, inv);
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) {
          const u64 rr = fma2(e[kp], inv2, 0ull);
          S0[kp] = add2(S0[kp], rr);
          S1[kp] = fma2(rr, z[kp], S1[kp]);
          S2[kp] = fma2(rr, ar[kp], S2[kp]);
        }
        x += 1e-3f;
        g = g * 0.999f + 1e-4f;
      }
    }
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) { unpk(add2(add2(S0[kp], S1[kp]), S2[kp]), lo, hi); s += lo + hi; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { unpk(v[i], lo, hi); s += lo + hi; }
  unpk(add2(add2(acc0, acc1), add2(acc2, acc3)), lo, hi);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + lo + hi;
}

template <int OP>
void run(const char *name, double units_per_iter, const char *unit, int sms, double mhz) {
  float *out = nullptr;
  const int threads = 512, blocks = sms * 2, iters = 4096;
  if (cudaMalloc(&out, sizeof(float) * threads * blocks) != cudaSuccess) { printf("cudaMalloc failed\n"); return; }
  k<OP><<<blocks, threads>>>(out, 64, 0.5f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  const double lanes = (double)blocks * threads * iters;
  const double per_clk_sm = lanes * units_per_iter / (ms * 1e-3) / (mhz * 1e6) / sms;
  printf("%-64s %8.3f ms  %8.2f %s/clk/SM\n", name, ms, per_clk_sm, unit);
  fflush(stdout);
  cudaFree(out);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double mhz = p.clockRate / 1e3;
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs, clockRate %.0f MHz\n", p.name, sms, mhz);
  run<0>("FFMA2 pair x scalar + scalar", 8, "instr-lanes", sms, mhz);
  run<1>("FFMA2 -pair x pair + scalar", 8, "instr-lanes", sms, mhz);
  run<2>("FFMA2 scalar x pair + pair", 8, "instr-lanes", sms, mhz);
  run<3>("FFMA2 three pairs", 8, "instr-lanes", sms, mhz);
  run<6>("1 LDS.128 + 4 MUFU", 32, "MUFU lane-ops", sms, mhz);
  run<7>("4 MUFU", 32, "MUFU lane-ops", sms, mhz);
  run<8>("points-packed body, K = 8 (per component: LDS.128, 4 FFMA2, 4 MUFU, 2 FADD2)", 32, "MUFU lane-ops", sms, mhz);
  run<9>("component-packed body, K = 8 (per pair: LDS.128 + LDS.64, 8 FFMA2, 8 MUFU, 4 FADD2)", 32, "MUFU lane-ops", sms, mhz);
  run<10>("statistics body, K = 8 (per point 16 FFMA2, 7 FADD2, 4 FMUL2, 9 MUFU; x as a pair)", 4 * 9, "MUFU lane-ops", sms, mhz);
  run<11>("statistics body, K = 8 (x as a scalar-broadcast operand)", 4 * 9, "MUFU lane-ops", sms, mhz);
  return 0;
}
