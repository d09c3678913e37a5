// packed_f32.cuh -- packed f32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100a) shared by the kernels.
// A pair travels as one 64-bit register pair; pk / unpk are register renames, not instructions.
// Measured on the B200 (tools/pipe_peaks2.cu): a packed instruction halves the issue slots of two scalar
// ones but occupies the dispatch port for 2 cycles (3 with three distinct register pairs).
#pragma once

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpk(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// c - a * b, packed: the negation is written on the halves so that ptxas folds it into the FFMA2
__device__ __forceinline__ u64 fnma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\t.reg .b64 n;\n\tmov.b64 {lo, hi}, %1;\n\tneg.f32 lo, lo;\n\tneg.f32 hi, hi;\n\t"
      "mov.b64 n, {lo, hi};\n\tfma.rn.f32x2 %0, n, %2, %3;\n\t}" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
