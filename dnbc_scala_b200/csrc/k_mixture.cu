// k_mixture.cu -- the two Gaussian-mixture sweeps of an EM iteration.
//
//   k_emission_px (K1): E[p][j] = ln P(o_p | state j)
//       reference arithmetic: emissionsSum of viterbiInitialize / viterbiCompute,
//       core/src/main/scala/DynamicNaiveBayesianClassifier.scala:52-56, 79-83;
//       LearnedDiscreteEdge.probability Edge.scala:69-75; LearnedContinuousEdge.probability
//       Edge.scala:114-118 -> gaussianMixturePdf GaussianUtils.scala:43-45.
//   k_stats_pk (K3): fp64 sufficient statistics from the posteriors gamma[p][j]:
//       GMM responsibilities / moments (generalises EM1D.run's E- and M-step sums,
//       core/src/main/java/Tools/ExpectationMaximization1D.java:92-124), categorical
//       histograms (generalises DiscreteEdge.learn, Edge.scala:46-51) and the gamma sums.
//
// Both sweeps evaluate every (continuous variable c, state j, component k) Gaussian at every
// point: ~8.2k exponentials per point at the scale-out shape against 80 B of observations, so
// they are bound by the MUFU pipe (measured 16 ex2/clk/SM, tools/pipe_peaks.cu) and by
// instruction issue, not by HBM.  Design rules that follow from the measurements:
//   * fp32 work is issued as PACKED f32x2 instructions (FFMA2 / FMUL2 / FADD2; the halves are two points of
//     one component in the emission sweep, two components of one point in the statistics sweep): the FP32 pipe
//     rate is unchanged but the issue slots halve, which is what lets MUFU run near its peak
//     (tools/pipe_peaks2.cu: a body of 1 MUFU + 8 fp32 ops reaches 96 % of MUFU peak packed, 79 % scalar);
//   * lanes run over hidden states (32 per warp; several points per warp when N <= 16), so
//     parameter fetches are conflict-free and E / gamma rows are read and written as whole
//     128-byte lines;
//   * emission: every warp evaluates all variables for its own 4 points per lane group; coefficients stream
//     from shared memory (one LDS.128 per component, amortised over the 4 points), observations arrive through
//     cp.async one group ahead; no block barriers, no cross-warp traffic, high occupancy;
//   * statistics: one warp per variable keeps that variable's coefficients AND its 3K moment
//     accumulators in registers (fp32 over a window of <= 4096 points, in the standardised
//     variable z and re-centred on the fp64 old mean when flushed: no x^2 - mu^2 cancellation);
//     categorical bins live in warp-private shared-memory columns (no atomics); each window
//     ends with one fp64 atomicAdd per statistic.
#include "dnbc_internal.cuh"
#include "packed_f32.cuh"

__device__ __forceinline__ u64 ex2_2(u64 a) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tex2.approx.ftz.f32 lo, lo;\n\tex2.approx.ftz.f32 hi, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}" : "=l"(d) : "l"(a));
  return d;
}
__device__ __forceinline__ float max3lohi(float m, u64 a) {
  float lo, hi;
  unpk(a, lo, hi);
  return fmaxf(m, fmaxf(lo, hi));
}
__device__ __forceinline__ float hsum(u64 a) {
  float lo, hi;
  unpk(a, lo, hi);
  return lo + hi;
}

// ---------------------------------------------------------------------------------
// K1 emission
// ---------------------------------------------------------------------------------
struct EmitArgs {
  Dims dm;
  DeviceTables tb;
  DataView dv;
  int64_t p0;
  int np;       // points of this launch (chunks are capped below 2^30 points)
  float *E;
  int LG;       // lanes per point = min(32, Npad)
  int vec;      // observation planes are 16-byte aligned at every group of 4 points
};

constexpr int EPU = 4;   // points per thread and coefficient fetch
// The log2 is the second most frequent MUFU operation of the sweep (one per state, variable and point against K
// exponentials), so two variables share one: the first variable's mixture sums are held, the second adds
// log2(held * sum) - 2 SHIFT.  The staged log2 constants carry SHIFT = EMIT_SHIFT when the kernel pairs (0 otherwise),
// so a sum that passes the underflow test is >= 1e-18 and the product of two cannot leave the fp32 range; the shift is
// taken back at once (an exact fp32 subtraction), so the accumulator keeps its small magnitude.  9,216 -> 8,704 MUFU
// operations per point at the scale-out shape: 72.3 -> 66.8 ms per 3.2e7 points.
constexpr float EMIT_SHIFT = 30.f;       // log2 scale of the staged mixture constants of a pairing kernel

// ---------------------------------------------------------------------------------
// K1 emission, packed over POINTS (k_emission_px)
// ---------------------------------------------------------------------------------
// The same sums as the component-packed kernel it replaced (k_emission_pk), but the two halves of a packed instruction are two consecutive points of ONE
// mixture component instead of two components at one point.  A component's coefficients then enter the FFMA2 as
// scalar broadcast operands (SASS: FFMA2 R, R.F32x2, R.F32, R.F32), so
//   * an odd component count costs nothing (K = 3 is three component steps, not a pair plus a scalar tail with
//     its own code path), and no per-variable count checks run inside the component loop;
//   * the mixture sums of a lane's four points are two packed registers: no horizontal add, packed log2
//     arguments, packed accumulation; one underflow test per variable covers the four points;
//   * lane groups below 32 states (N = 10: three points x 10 lanes) read 4 consecutive points per lane with one
//     16-byte load like the full-warp layout does.
// ncu of the component-packed kernel this one replaced (profiles/r02t_ncu_n512_summary.txt, N = 512, K = 3): 40 warp
// instructions per (point, variable, 32 states), issue 77 %, MUFU 63 %; at the scale-out shape (K = 8) it reached
// 90 % of the MUFU pipe, this kernel 94 %.
// KU > 0: every variable has exactly KU components (unrolled, log2 shared by pairs of variables);
// KU = 0: per-variable counts (runtime loop, one log2 per variable).
template <int KU>
struct PxScale {
  static constexpr float SHIFT = KU > 0 ? EMIT_SHIFT : 0.f;
  static constexpr float TINY = KU > 0 ? 1e-18f : 1e-30f;
};
__device__ __forceinline__ u64 lg2_2(u64 a) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tlg2.approx.ftz.f32 lo, lo;\n\tlg2.approx.ftz.f32 hi, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}" : "=l"(d) : "l"(a));
  return d;
}
// exponent-shifted evaluation of one (point, state, variable) whose mixture sum underflowed fp32: log2 of the sum
// without the staged shift, finite wherever the reference's fp64 sum is finite; kept out of line.  Not rare with
// hundreds of states: at N = 512 a quarter of the (4 points, variable, 32 states) groups have a state more than 11
// sigma from every component at some point, and elimination builds put 29 % of the sweep's time here
// (profiles/r02e_emission_elimination.jsonl).  Re-evaluating all four points of a flagged lane at once, packed, is
// slower (9.36 vs 8.77 ms: usually one of the four is flagged); what helps is fewer and shorter evaluations --
// emit_var_px's loop over each lane's next flagged point, and unrolled component loops here: 8.78 -> 8.06 ms.
// Deferring the flagged (variable, point) pairs of a lane to after its last variable, so that lanes flagged at
// different variables share their evaluations, changes nothing (7.97 vs 7.96 ms; +1.5 % at N = 10): not kept.
__device__ __noinline__ float emit_px_slow(const float4 *ps, int kc, float x, float shift) {
  float mx = -INFINITY;
  for (int k = 0; k < kc; ++k) {
    const float4 q = ps[k * 32];
    const float z = fmaf(x, q.x, q.y);
    mx = fmaxf(mx, fmaf(-z, z, q.z));
  }
  if (mx == -INFINITY) return -INFINITY;                     // no component has mass
  float s = 0.f;
  for (int k = 0; k < kc; ++k) {
    const float4 q = ps[k * 32];
    const float z = fmaf(x, q.x, q.y);
    s += ex2_approx(fmaf(-z, z, q.z) - mx);
  }
  return mx + lg2_approx(s) - shift;
}
// the same for a compile-time component count up to 4: unrolled, so the K coefficient loads and chains overlap
template <int KU>
__device__ __noinline__ float emit_px_slow_u(const float4 *ps, float x, float shift) {
  float a[KU], mx = -INFINITY, s = 0.f;
#pragma unroll
  for (int k = 0; k < KU; ++k) {
    const float4 q = ps[k * 32];
    const float z = fmaf(x, q.x, q.y);
    a[k] = fmaf(-z, z, q.z);
    mx = fmaxf(mx, a[k]);
  }
  if (mx == -INFINITY) return -INFINITY;
#pragma unroll
  for (int k = 0; k < KU; ++k) s += ex2_approx(a[k] - mx);
  return mx + lg2_approx(s) - shift;
}
// one variable, 4 points (x01, x23): log2 mixture density added to acc01 / acc23.  PAIR as in emit_var.
template <int KU, int PAIR>
__device__ __forceinline__ void emit_var_px(const float4 *ps, int kc, u64 x01, u64 x23, u64 &acc01, u64 &acc23,
                                            u64 &hold01, u64 &hold23) {
  constexpr float SH = PxScale<KU>::SHIFT;
  u64 s01 = 0ull, s23 = 0ull;
  if (KU > 0) {
#pragma unroll
    for (int k = 0; k < KU; ++k) {
      const float4 q = ps[k * 32];
      const u64 sb = pk(q.x, q.x), mb = pk(q.y, q.y), cb = pk(q.z, q.z);
      const u64 z01 = fma2(x01, sb, mb), z23 = fma2(x23, sb, mb);
      const u64 e01 = ex2_2(fnma2(z01, z01, cb)), e23 = ex2_2(fnma2(z23, z23, cb));
      s01 = k == 0 ? e01 : add2(s01, e01);
      s23 = k == 0 ? e23 : add2(s23, e23);
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < kc; ++k) {
      const float4 q = ps[k * 32];
      const u64 sb = pk(q.x, q.x), mb = pk(q.y, q.y), cb = pk(q.z, q.z);
      const u64 z01 = fma2(x01, sb, mb), z23 = fma2(x23, sb, mb);
      s01 = add2(s01, ex2_2(fnma2(z01, z01, cb)));
      s23 = add2(s23, ex2_2(fnma2(z23, z23, cb)));
    }
  }
  float v[4];
  unpk(s01, v[0], v[1]);
  unpk(s23, v[2], v[3]);
  if (fminf(fminf(v[0], v[1]), fminf(v[2], v[3])) < PxScale<KU>::TINY) {
    // some point's mixture underflowed (every component > ~13 sigma away) or has no mass: its term is completed
    // here and it joins the log2 below as the neutral factor 2^SHIFT
    // Each trip of the loop re-evaluates every lane's NEXT flagged point, whichever of the four it is: the number
    // of (out-of-line, latency-bound) evaluations a warp pays is the largest count of flagged points in one lane,
    // not the number of distinct flagged positions over its lanes.
    float x[4], ac[4];
    unpk(x01, x[0], x[1]);
    unpk(x23, x[2], x[3]);
    unpk(acc01, ac[0], ac[1]);
    unpk(acc23, ac[2], ac[3]);
    if constexpr (KU <= 4) {
      unsigned fl = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) fl |= (v[i] < PxScale<KU>::TINY ? 1u : 0u) << i;
      while (fl) {
        const int i = __ffs(fl) - 1;
        fl &= fl - 1;
        const float xi = i == 0 ? x[0] : (i == 1 ? x[1] : (i == 2 ? x[2] : x[3]));
        float L;
        if constexpr (KU > 0) L = emit_px_slow_u<KU>(ps, xi, SH);
        else L = emit_px_slow(ps, kc, xi, SH);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ac[e] += e == i ? L : 0.f;
          v[e] = e == i ? (KU > 0 ? 1073741824.f : 1.f) : v[e];
        }
      }
    } else {                                                 // (large mixtures rarely underflow as a whole: the plain form
                                                             // keeps the scale-out kernel's code as it measured fastest)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (v[i] < PxScale<KU>::TINY) {
          ac[i] += emit_px_slow(ps, KU, x[i], SH);
          v[i] = 1073741824.f;
        }
    }
    acc01 = pk(ac[0], ac[1]);
    acc23 = pk(ac[2], ac[3]);
    s01 = pk(v[0], v[1]);
    s23 = pk(v[2], v[3]);
  }
  if (PAIR == 1) {
    hold01 = s01;
    hold23 = s23;
  } else if (PAIR == 2) {
    const u64 back = pk(-2.f * SH, -2.f * SH);
    acc01 = add2(acc01, add2(lg2_2(mul2(hold01, s01)), back));
    acc23 = add2(acc23, add2(lg2_2(mul2(hold23, s23)), back));
  } else if (SH != 0.f) {
    const u64 back = pk(-SH, -SH);
    acc01 = add2(acc01, add2(lg2_2(s01), back));
    acc23 = add2(acc23, add2(lg2_2(s23), back));
  } else {
    acc01 = add2(acc01, lg2_2(s01));
    acc23 = add2(acc23, lg2_2(s23));
  }
}

// Observations reach the arithmetic through shared memory: a warp's next group of points (4 consecutive points per
// lane group, every variable) is requested with cp.async one whole group ahead and read back with one broadcast
// LDS.128 (LDS.32 for the symbols) per variable.  Register prefetch of the next variable's observations does not
// survive ptxas here: the loaded registers are copied at the merge of the aligned / ragged load paths, and the copy
// waits for the load (ncu of the first version of this kernel: 35 % of the stall samples on those copies).
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_px() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_px() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int PX_WARPS = 16;
// shared memory: coefficients of the 32-state slice [C][KS][32] x {s, m, c, -}, the categorical table
// [D][Vmax+1][32], then per warp two stages of observations [C][PG] x float4 and symbols [D][PG] x uint32
// BIGP: more than 64 staging pieces per group (many variables x many lane groups of few states): piece offsets on the fly
template <int KU, bool BIGP = false>
__global__ void __launch_bounds__(PX_WARPS * 32, 2) k_emission_px(EmitArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax, Npad = a.dm.Npad;
  const int KS = KU > 0 ? KU : Kmax;                        // component slots per variable
  const int LG = a.LG, PG = 32 / LG;
  float4 *sm_q = (float4 *)smraw;                           // [C*KS*32]
  float *sm_cat = (float *)(sm_q + (size_t)C * KS * 32);    // [D*(Vmax+1)*32]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lj = lane % LG, lp = lane / LG;
  const int xstage = C * PG, vstage = (D * PG + 3) & ~3;    // float4 / uint32 entries of one stage
  float4 *xs = (float4 *)(sm_cat + (size_t)D * (Vmax + 1) * 32) + (size_t)w * 2 * xstage;          // [2][C][PG]
  uint32_t *vs = (uint32_t *)((float4 *)(sm_cat + (size_t)D * (Vmax + 1) * 32) + (size_t)PX_WARPS * 2 * xstage) +
                 (size_t)w * 2 * vstage;                                                            // [2][D][PG]
  const int jbase = blockIdx.y * 32;
  for (int t = threadIdx.x; t < C * KS * 32; t += blockDim.x) {
    const int sl = t & 31, k = (t >> 5) % KS, c = (t >> 5) / KS;
    const int j = jbase + sl % LG;
    const bool on = j < N && k < a.tb.K[c];
    const int64_t o = ((int64_t)c * Kmax + k) * Npad + (j < Npad ? j : 0);
    // (padded lanes get one unit-mass component so they never look like an underflowed mixture)
    sm_q[t] = make_float4(on ? a.tb.gs[o] : 0.f, on ? a.tb.gm[o] : 0.f,
                          on ? a.tb.gc[o] + PxScale<KU>::SHIFT : (j < N || k > 0 ? -INFINITY : PxScale<KU>::SHIFT), 0.f);
  }
  __shared__ int sK[64];
  for (int t = threadIdx.x; t < C; t += blockDim.x) sK[t] = a.tb.K[t];
  for (int t = threadIdx.x; t < D * (Vmax + 1) * 32; t += blockDim.x) {
    const int sl = t & 31, v = (t >> 5) % (Vmax + 1), d = (t >> 5) / (Vmax + 1);
    const int j = jbase + sl % LG;
    sm_cat[t] = j < N ? a.tb.cat[((int64_t)d * (Vmax + 1) + v) * Npad + j] : -INFINITY;
  }
  __syncthreads();
  const int j = jbase + lj;
  const int group = EPU * PG;                               // points per warp iteration
  const int ngroups = (a.np + group - 1) / group;
  const int64_t P = a.dv.P;
  const int cstride = KS * 32, tstride = (Vmax + 1) * 32;
  const bool lane_on = lp < PG && j < N;                    // lanes past the last whole group and padded states idle
  const int lpc = lp < PG ? lp : 0;
  const float *cont0 = a.dv.cont + a.p0;
  const uint8_t *disc0 = a.dv.disc + a.p0;
  // request group g's observations into stage st: piece i = (variable i / PG, lane group i % PG) is 4 consecutive
  // points of one plane; lane l owns pieces l, l + 32, ... and holds the plane offsets of its first two (shapes with
  // more than 64 pieces -- many variables x many lane groups -- compute the rest on the fly).  Whole groups of aligned planes: one 16-byte cp.async per piece and one 4-byte cp.async per
  // symbol piece; the ragged last group and unaligned planes: guarded loads and ordinary stores.
  auto piece_off = [&](int i) -> int64_t {
    const int v = i / PG;
    return (int64_t)v * P + (i - v * PG) * EPU;
  };
  int64_t poff[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) poff[h] = piece_off(lane + 32 * h);
  const int npx = C * PG, npv = D * PG;
  auto request = [&](int g, int st) {
    const int pb = g * group;
    float4 *xd = xs + st * xstage;
    uint32_t *vd = vs + st * vstage;
    if (a.vec && pb + group <= a.np) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < npx) cp_async16(xd + i, cont0 + poff[h] + pb);
        if (i < npv) cp_async4(vd + i, disc0 + poff[h] + pb);
      }
      if (BIGP) {
        for (int i = lane + 64; i < (npx > npv ? npx : npv); i += 32) {
          const int64_t off = piece_off(i) + pb;
          if (i < npx) cp_async16(xd + i, cont0 + off);
          if (i < npv) cp_async4(vd + i, disc0 + off);
        }
      }
    } else {
#pragma unroll 1
      for (int i = lane; i < (BIGP ? (npx > npv ? npx : npv) : 64); i += 32) {
        const int q = pb + (i % PG) * EPU;
        const int64_t off = piece_off(i) + pb;
        if (i < npx) {
          const float *src = cont0 + off;
          float x[EPU];
#pragma unroll
          for (int e = 0; e < EPU; ++e) x[e] = q + e < a.np ? __ldg(src + e) : 0.f;
          xd[i] = make_float4(x[0], x[1], x[2], x[3]);
        }
        if (i < npv) {
          const uint8_t *src = disc0 + off;
          uint32_t w4 = 0;
#pragma unroll
          for (int e = 0; e < EPU; ++e) w4 |= (q + e < a.np ? (uint32_t)__ldg(src + e) : (uint32_t)Vmax) << (8 * e);
          vd[i] = w4;
        }
      }
    }
    cp_async_commit_px();
  };
  const int gstride = gridDim.x * PX_WARPS;
  int g = blockIdx.x * PX_WARPS + w;
  if (g < ngroups) request(g, 0);
  for (int st = 0; g < ngroups; g += gstride, st ^= 1) {
    cp_async_wait_px();
    __syncwarp();                                            // stage st is complete; stage st ^ 1 is no longer read
    if (g + gstride < ngroups) request(g + gstride, st ^ 1);
    const float4 *xq = xs + st * xstage + lpc;                // this lane group's piece of variable 0; advanced per variable
    const uint32_t *vq = vs + st * vstage + lpc;
    u64 acc01 = 0ull, acc23 = 0ull, hold01 = 0ull, hold23 = 0ull;
    const float4 *ps = sm_q + lane;
    const float *tab = sm_cat + lane;
    auto add_cat = [&]() {                                   // one symbol per byte; ids >= V mean "never seen"
      const uint32_t w4 = *vq;
      vq += PG;
      float t[EPU];
#pragma unroll
      for (int i = 0; i < EPU; ++i) {
        const int v = (int)__byte_perm(w4, 0u, 0x4440 + i);
        t[i] = tab[(v < Vmax ? v : Vmax) * 32];
      }
      acc01 = add2(acc01, pk(t[0], t[1]));
      acc23 = add2(acc23, pk(t[2], t[3]));
      tab += tstride;
    };
    int c = 0;
    if (KU > 0) {
      for (; c + 1 < C; c += 2) {                            // pairs of variables share one log2
        const float4 xa = xq[0], xb = xq[PG];
        xq += 2 * PG;
        emit_var_px<KU, 1>(ps, KU, pk(xa.x, xa.y), pk(xa.z, xa.w), acc01, acc23, hold01, hold23);
        if (c < D) add_cat();
        emit_var_px<KU, 2>(ps + cstride, KU, pk(xb.x, xb.y), pk(xb.z, xb.w), acc01, acc23, hold01, hold23);
        if (c + 1 < D) add_cat();
        ps += 2 * cstride;
      }
    }
    for (; c < C; ++c) {
      const float4 xa = xq[0];
      xq += PG;
      emit_var_px<KU, 0>(ps, KU > 0 ? KU : sK[c], pk(xa.x, xa.y), pk(xa.z, xa.w), acc01, acc23, hold01, hold23);
      if (c < D) add_cat();
      ps += cstride;
    }
    for (int d = C; d < D; ++d) add_cat();                   // more discrete than continuous variables
    if (lane_on) {
      const int q0 = g * group + lp * EPU;
      const u64 ln2 = pk((float)DNBC_LN2, (float)DNBC_LN2);
      float e[EPU];
      unpk(mul2(acc01, ln2), e[0], e[1]);
      unpk(mul2(acc23, ln2), e[2], e[3]);
      float *dst = a.E + (int64_t)q0 * N + j;
      if (q0 + EPU <= a.np) {
#pragma unroll
        for (int i = 0; i < EPU; ++i, dst += N) *dst = e[i];
      } else {
#pragma unroll
        for (int i = 0; i < EPU; ++i, dst += N)
          if (q0 + i < a.np) *dst = e[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// K3 statistics
// ---------------------------------------------------------------------------------
struct StatArgs {
  Dims dm;
  DeviceTables tb;
  DataView dv;
  int64_t p0;
  int np;
  const float *GA;
  const double *mu_old;
  double *gmm, *cat, *gsum;
  int LG, NW, window;
};

#ifndef DNBC_STATS_SPB
#define DNBC_STATS_SPB 4
#endif
constexpr int SPB = DNBC_STATS_SPB;   // points whose posteriors are fetched together, one group ahead of their use

template <int KP>
struct StatRegs {
  u64 s2[KP], m2[KP], cc2[KP], S0[KP], S1[KP], S2[KP];
};

// one point: responsibilities of this warp's variable, weighted by the posterior g.
// Branch-free: exponents are shifted by their maximum before ex2, so the mixture sum is >= 1
// whenever any component has mass (no fp32 underflow, no fallback path) and consecutive points
// interleave freely in the instruction stream.
template <int KP>
__device__ __forceinline__ void stat_point(StatRegs<KP> &q, float x, float g) {
  const u64 neg1 = pk(-1.f, -1.f);
  const u64 xx = pk(x, x);
  u64 a2[KP], z2[KP], zz2[KP];
  float mx = -INFINITY;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    z2[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
    zz2[kp] = mul2(z2[kp], z2[kp]);
    a2[kp] = fma2(zz2[kp], neg1, q.cc2[kp]);
    mx = max3lohi(mx, a2[kp]);
  }
  mx = mx == -INFINITY ? 0.f : mx;                          // no component has mass: e = 0, sum = 0
  const u64 nm = pk(-mx, -mx);
  u64 sum2 = 0ull;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    a2[kp] = ex2_2(add2(a2[kp], nm));
    sum2 = add2(sum2, a2[kp]);
  }
  const float sum = hsum(sum2);
  const float inv = sum > 0.f ? g * rcp_approx(sum) : 0.f;
  const u64 inv2 = pk(inv, inv);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 rr = mul2(a2[kp], inv2);
    q.S0[kp] = add2(q.S0[kp], rr);
    q.S1[kp] = fma2(rr, z2[kp], q.S1[kp]);
    q.S2[kp] = fma2(rr, zz2[kp], q.S2[kp]);
  }
}

// fast path, no exponent shift: 8 packed fp32 instructions + 2 MUFU per component pair
// (z, z^2, c - z^2 | sum | r, S0, S1, S2) and no branch, so consecutive points interleave.
// When the mixture sum underflows fp32 at a point that carries posterior mass (every component
// > ~13 sigma away) the point contributes nothing here and the function returns true; the
// caller then re-runs the flagged points of the block through the shifted stat_point above.
#ifndef DNBC_STATS_FAST
#define DNBC_STATS_FAST 1
#endif
// A state whose whole mixture underflows fp32 at a point almost always has a posterior far below
// fp32 resolution there as well.  Only a posterior above this threshold sends the block through the
// shifted re-evaluation; below it the point's contribution (< 1e-12 of one point's unit mass, against
// sums of 10^2..10^8 and a 1e-5 parity budget) is dropped.
#ifndef DNBC_STATS_GMIN
#define DNBC_STATS_GMIN 1e-12f
#endif
template <int KP>
__device__ __forceinline__ bool stat_point_fast(StatRegs<KP> &q, float x, float g) {
  const float TINY = 1e-30f;
  const u64 xx = pk(x, x);
  u64 e[KP], z[KP], zz[KP];
  u64 sum2;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    z[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
    zz[kp] = mul2(z[kp], z[kp]);
    e[kp] = ex2_2(sub2(q.cc2[kp], zz[kp]));
    sum2 = kp == 0 ? e[0] : add2(sum2, e[kp]);
  }
  const float sum = hsum(sum2);
  const bool under = sum < TINY;
  const float inv = under ? 0.f : g * rcp_approx(sum);
  const u64 inv2 = pk(inv, inv);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 rr = mul2(e[kp], inv2);
    q.S0[kp] = add2(q.S0[kp], rr);
    q.S1[kp] = fma2(rr, z[kp], q.S1[kp]);
    q.S2[kp] = fma2(rr, zz[kp], q.S2[kp]);
  }
  return under && g > DNBC_STATS_GMIN;
}

// The same fast path split in two stages so that stat_block can software-pipeline it: stage 1
// (z, z^2, the 2K exponentials) of point p + 1 is issued together with stage 2 (normalise,
// accumulate) of point p.  Without this the four warps of a scheduler fall into a convoy --
// all in the MUFU phase, then all in the fp32 phase -- and the two pipes never overlap
// (measured: 135 clk per warp-point = 72 MUFU + 64 fp32, profiles/r01_ncu_final.txt).
#ifndef DNBC_STATS_PIPE
#define DNBC_STATS_PIPE 1
#endif
#ifndef DNBC_STATS_UNROLL
#define DNBC_STATS_UNROLL 1
#endif
constexpr int STATS_UNROLL = DNBC_STATS_UNROLL;
// (holding z^2 of the pending point costs 8 registers and measured 7 % slower than recomputing it)
#ifndef DNBC_STATS_HOLDZZ
#define DNBC_STATS_HOLDZZ 0
#endif
#ifndef DNBC_STATS_NW
#define DNBC_STATS_NW 8
#endif
template <int KP>
struct StatPend {
  u64 z[KP], e[KP];
#if DNBC_STATS_HOLDZZ
  u64 zz[KP];
#endif
  float g;
};
template <int KP>
__device__ __forceinline__ void stat_stage1(const StatRegs<KP> &q, float x, float g, StatPend<KP> &n) {
  const u64 xx = pk(x, x);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    n.z[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
#if DNBC_STATS_HOLDZZ
    const u64 zz = mul2(n.z[kp], n.z[kp]);
    n.zz[kp] = zz;
    n.e[kp] = ex2_2(sub2(q.cc2[kp], zz));
#else
    n.e[kp] = ex2_2(fnma2(n.z[kp], n.z[kp], q.cc2[kp]));
#endif
  }
  n.g = g;
}
template <int KP>
__device__ __forceinline__ bool stat_stage2(StatRegs<KP> &q, const StatPend<KP> &p) {
  const float TINY = 1e-30f;
  u64 sum2 = p.e[0];
#pragma unroll
  for (int kp = 1; kp < KP; ++kp) sum2 = add2(sum2, p.e[kp]);
  const float sum = hsum(sum2);
  const bool under = sum < TINY;
  const float inv = under ? 0.f : p.g * rcp_approx(sum);
  const u64 inv2 = pk(inv, inv);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 rr = mul2(p.e[kp], inv2);
    q.S0[kp] = add2(q.S0[kp], rr);
#if DNBC_STATS_HOLDZZ
    q.S1[kp] = fma2(rr, p.z[kp], q.S1[kp]);
    q.S2[kp] = fma2(rr, p.zz[kp], q.S2[kp]);
#else
    const u64 rz = mul2(rr, p.z[kp]);
    q.S1[kp] = add2(q.S1[kp], rz);
    q.S2[kp] = fma2(rz, p.z[kp], q.S2[kp]);
#endif
  }
  return under && p.g > DNBC_STATS_GMIN;
}

// 32 points: the warp-uniform role flags are template parameters so the inner loop carries no
// per-point branches.  Posteriors are fetched SPB points ahead (gn) of the ones being consumed.
// FULL: all 32 points exist and the lane's state is real, so the posterior fetches carry no predicates.
// DEEP (only with FULL): posteriors are fetched 8 points ahead instead of 4.  With at most two component pairs a
// point costs ~60 clk, so four points do not cover an L2 miss (N = 512, K = 3: -7 %); with four pairs the extra
// registers cost more than the latency they hide.
template <int KP, bool HAS_C, bool HAS_D, bool GSUM, bool FULL, bool DEEP = false>
__device__ __forceinline__ void stat_block(StatRegs<KP> &q, double &gs, float *bd, const float *gp, int gstep, float xreg,
                                           int vreg, int NS, int PG, int lp, int base, int pe, bool jon) {
  // NS = steps of PG points that cover the block's 32 points (= the lane-group size when that divides 32)
  constexpr bool fast = FULL;
  constexpr int SPB = DEEP ? 2 * ::SPB : ::SPB;
  float gn[SPB];
  float gsl = 0.f;   // gamma sum of this block; folded into fp64 per block (an fp32 running sum over a
                     // 4096-point window reaches ~64 and would drop every posterior below 4e-6)
  bool bad = false;
  StatPend<KP> pend[2];
  const float *gp0 = gp;
  auto fetch = [&](int blk) {
    if (FULL) {                                              // (LG is a multiple of SPB when N >= SPB)
#pragma unroll
      for (int i = 0; i < SPB; ++i) gn[i] = __ldg(gp + i * gstep);
    } else {
#pragma unroll
      for (int i = 0; i < SPB; ++i)
        gn[i] = (blk + i < NS && (blk + i) * PG + lp < 32 && base + (blk + i) * PG + lp < pe && jon) ? __ldg(gp + i * gstep) : 0.f;
    }
    gp += SPB * gstep;
  };
  fetch(0);
  if (HAS_C && DNBC_STATS_PIPE) stat_stage1<KP>(q, __shfl_sync(0xffffffffu, xreg, lp), gn[0], pend[0]);
#pragma unroll STATS_UNROLL
  for (int blk = 0; blk < NS; blk += SPB) {                 // NS steps of PG points cover the 32 points
    float g[SPB];
#pragma unroll
    for (int i = 0; i < SPB; ++i) g[i] = gn[i];
    const bool more = blk + SPB < NS;
    if (more) fetch(blk + SPB);
#pragma unroll
    for (int i = 0; i < SPB; ++i) {
      const int qp = (blk + i) * PG + lp;                    // (steps past the block read lane qp & 31: g is 0 there)
      if (HAS_C) {
        if (DNBC_STATS_PIPE) {
          // stage 1 of the next point rides with stage 2 of this one; two buffers alternate (no copies)
          // (after the block's last point this evaluates one point too many and discards it:
          // cheaper than a branch, which costs the compiler its register assignment)
          const float xn = __shfl_sync(0xffffffffu, xreg, qp + PG);
          stat_stage1<KP>(q, xn, i + 1 < SPB ? g[i + 1] : gn[0], pend[(i + 1) & 1]);
          bad |= stat_stage2<KP>(q, pend[i & 1]);
        } else {
          const float x = __shfl_sync(0xffffffffu, xreg, qp);
          if (DNBC_STATS_FAST) bad |= stat_point_fast<KP>(q, x, g[i]);
          else stat_point<KP>(q, x, g[i]);
        }
      }
      if (HAS_D) {                                           // categorical histogram: bins[v][lane] += gamma
        const int v = __shfl_sync(0xffffffffu, vreg, qp);
        bd[v * 32] += g[i];
      }
      if (GSUM) gsl += g[i];
    }
  }
  if (GSUM) gs += (double)gsl;
  if (HAS_C && (DNBC_STATS_FAST || DNBC_STATS_PIPE) && __any_sync(0xffffffffu, bad && jon)) {
    // rare: some point of this block sits where a state's whole mixture underflows fp32.  Redo the
    // block through the shifted evaluation, each lane contributing only where its fast path gave up
    // (the test below repeats stat_point_fast's: same inputs, same instructions, same outcome).
#pragma unroll 1
    for (int st = 0; st < NS; ++st) {
      const bool on = jon && (fast || (st * PG + lp < 32 && base + st * PG + lp < pe));
      const float g = on ? __ldg(gp0 + (int64_t)st * gstep) : 0.f;
      const float x = __shfl_sync(0xffffffffu, xreg, st * PG + lp);
      StatRegs<KP> probe = q;
      probe.S0[0] = 0ull;
      const bool under = stat_point_fast<KP>(probe, x, g);
      stat_point<KP>(q, x, under ? g : 0.f);
    }
  }
}

template <int KP>
#ifndef DNBC_STATS_OCC
#define DNBC_STATS_OCC (16 / DNBC_STATS_NW > 3 ? 3 : 16 / DNBC_STATS_NW)
#endif
__global__ void __launch_bounds__(DNBC_STATS_NW * 32, DNBC_STATS_OCC) k_stats_pk(StatArgs a) {
  extern __shared__ float bins[];                           // [warps of this CTA][Vmax+1][32]
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax, Npad = a.dm.Npad;
  const int LG = a.LG, PG = 32 / LG, NS = (32 + PG - 1) / PG, NW = a.NW;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, lj = lane % LG, lp = lane / LG;
  const int j = blockIdx.y * 32 + lj;
  const int u = blockIdx.z * NW + w;                        // this warp's variable index
  // (a lane group that does not divide 32 leaves 32 - PG * LG idle lanes at the top of the warp)
  const bool has_c = u < C, has_d = u < D, jon = j < N && lp < PG;
  const int bin_d = (Vmax + 1) * 32;
  float *bd = bins + (size_t)w * bin_d + lane;
  for (int t = 0; t <= Vmax; ++t) bd[t * 32] = 0.f;

  StatRegs<KP> q;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    float v[6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * kp + h;
      const bool on = has_c && jon && k < a.tb.K[has_c ? u : 0];
      const int64_t o = ((int64_t)(has_c ? u : 0) * Kmax + (k < Kmax ? k : 0)) * Npad + (j < Npad ? j : 0);
      v[h] = on ? a.tb.gs[o] : 0.f;
      v[2 + h] = on ? a.tb.gm[o] : 0.f;
      v[4 + h] = on ? a.tb.gc[o] : -INFINITY;
    }
    q.s2[kp] = pk(v[0], v[1]); q.m2[kp] = pk(v[2], v[3]); q.cc2[kp] = pk(v[4], v[5]);
    q.S0[kp] = q.S1[kp] = q.S2[kp] = 0ull;
  }
  double gs = 0.0;
  // contiguous slab of points per CTA column, aligned to 32 so x / symbol loads stay coalesced
  const int per = ((a.np + (int)gridDim.x - 1) / (int)gridDim.x + 31) / 32 * 32;
  const int pb = (int)blockIdx.x * per;
  const int pe = pb + per < a.np ? pb + per : a.np;
  const float *xcol = a.dv.cont + (int64_t)(has_c ? u : 0) * a.dv.P + a.p0;
  const uint8_t *dcol = a.dv.disc + (int64_t)(has_d ? u : 0) * a.dv.P + a.p0;
  const float *gcol = a.GA + (jon ? j : 0);
  const int gstep = PG * N;                                 // floats between consecutive warp steps
  int inwin = 0;
  // the observation / symbol of point base + lane, fetched one 32-point block ahead
  float xnext = 0.f;
  int vnext = Vmax;
  if (pb + lane < pe) {
    if (has_c) xnext = __ldg(xcol + pb + lane);
    if (has_d) vnext = (int)__ldg(dcol + pb + lane);
  }
  for (int base = pb; base < pe; base += 32) {
    const bool full = base + 32 <= pe;
    const float xreg = xnext;
    const int vreg = vnext < Vmax ? vnext : Vmax;
    xnext = 0.f;
    vnext = Vmax;
    if (base + 32 + lane < pe) {
      if (has_c) xnext = __ldg(xcol + base + 32 + lane);
      if (has_d) vnext = (int)__ldg(dcol + base + 32 + lane);
    }
    const float *gp = gcol + (int64_t)(base + lp) * N;      // posterior of this lane's state, step 0
    // warp-uniform: padded lanes (j >= N) run the same predicate-free path on state 0's posteriors and
    // their accumulators are never flushed (a per-lane flag made the warp execute BOTH block variants)
    const bool fast = full && LG >= SPB && PG * LG == 32;
    const bool deep = KP <= 2 && LG == 32;
#define STAT_BLOCK(C_, D_, G_)                                                                          \
  do {                                                                                                 \
    if (fast && deep) stat_block<KP, C_, D_, G_, true, KP <= 2>(q, gs, bd, gp, gstep, xreg, vreg, NS, PG, lp, base, pe, jon); \
    else if (fast) stat_block<KP, C_, D_, G_, true>(q, gs, bd, gp, gstep, xreg, vreg, NS, PG, lp, base, pe, jon);   \
    else stat_block<KP, C_, D_, G_, false>(q, gs, bd, gp, gstep, xreg, vreg, NS, PG, lp, base, pe, jon);       \
  } while (0)
    if (has_c && has_d) { if (u == 0) STAT_BLOCK(true, true, true); else STAT_BLOCK(true, true, false); }
    else if (has_c) { if (u == 0) STAT_BLOCK(true, false, true); else STAT_BLOCK(true, false, false); }
    else if (has_d) { if (u == 0) STAT_BLOCK(false, true, true); else STAT_BLOCK(false, true, false); }
    else if (u == 0) STAT_BLOCK(false, false, true);
#undef STAT_BLOCK
    inwin += 32;
    if (inwin >= a.window || base + 32 >= pe) {
      inwin = 0;
      if (has_c && jon) {
        const int K = a.tb.K[u];
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          float sv[2], mv[2], a0[2], a1[2], a2v[2];
          unpk(q.s2[kp], sv[0], sv[1]); unpk(q.m2[kp], mv[0], mv[1]);
          unpk(q.S0[kp], a0[0], a0[1]); unpk(q.S1[kp], a1[0], a1[1]); unpk(q.S2[kp], a2v[0], a2v[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 * kp + h;
            if (k < K && a0[h] != 0.f) {
              // z = (x - mu') s with mu' = -m/s the fp32-rounded centre; re-centre on the fp64 old mean
              const int64_t dst = ((int64_t)u * N + j) * Kmax + k;
              const double s = (double)sv[h];
              double m0 = (double)a0[h], m1 = 0.0, m2d = 0.0;
              if (s > 0.0) {
                const double delta = -(double)mv[h] / s - a.mu_old[dst];
                const double b1 = (double)a1[h] / s, b2 = (double)a2v[h] / (s * s);
                m1 = b1 + delta * m0;
                m2d = b2 + 2.0 * delta * b1 + delta * delta * m0;
              }
              atomicAdd(a.gmm + 3 * dst + 0, m0);
              atomicAdd(a.gmm + 3 * dst + 1, m1);
              atomicAdd(a.gmm + 3 * dst + 2, m2d);
            }
          }
          q.S0[kp] = q.S1[kp] = q.S2[kp] = 0ull;
        }
      }
      if (u == 0) {
        if (jon && gs != 0.0) atomicAdd(a.gsum + j, gs);
        gs = 0.0;
      }
      if (has_d) {
        for (int v = 0; v < Vmax; ++v) {
          const float b = bd[v * 32];
          if (b != 0.f && jon) atomicAdd(a.cat + ((int64_t)u * N + j) * Vmax + v, (double)b);
          bd[v * 32] = 0.f;
        }
        bd[Vmax * 32] = 0.f;                                // unseen-symbol row is discarded
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------
// Lanes per group of states in the emission / statistics sweeps.  Normally the padded (power-of-two) state
// count, so that groups tile the warp; the exact state count when that puts more points into a warp
// (N = 10: 3 x 10 lanes instead of 2 x 16; N = 5: 6 x 5 instead of 4 x 8).
static int lane_group(const Dims &d, bool pow2_only) {
  if (d.N >= 32) return 32;
  const int pg_pad = 32 / (d.Npad < 32 ? d.Npad : 32), pg_exact = 32 / d.N;
  return (!pow2_only && pg_exact > pg_pad) ? d.N : (d.Npad < 32 ? d.Npad : 32);
}

// the statistics sweep keeps 3K coefficients + 3K accumulators in registers and handles one
// continuous and one discrete variable per warp
bool mixture_stats_fast(const dnbc_ctx *c) {
  const Dims &d = c->dm;
  const int units = d.C > d.D ? d.C : d.D;
  return d.Kmax <= 8 && units <= 64 && (size_t)8 * (d.Vmax + 1) * 32 * sizeof(float) <= 96 * 1024;
}

// points-packed kernel: component slots per variable and shared memory
static size_t emit_px_smem(const Dims &d, int KS, int PG) {
  const size_t stage = (size_t)d.C * PG * sizeof(float4) + (size_t)((d.D * PG + 3) & ~3) * sizeof(uint32_t);
  return (size_t)d.C * KS * 32 * sizeof(float4) + (size_t)d.D * (d.Vmax + 1) * 32 * sizeof(float) + PX_WARPS * 2 * stage;
}
// emission through k_emission_px: up to 8 components per mixture, coefficients + staging within one SM's shared memory
// (everything else takes the generic kernel of k_emission.cu)
static int emit_px_ku(const dnbc_ctx *c) {                   // the common component count, or 0 (per-variable counts)
  bool uniform = c->dm.C > 0;
  for (int v = 0; v < c->dm.C; ++v) uniform = uniform && c->K[v] == c->dm.Kmax;
  return uniform ? c->dm.Kmax : 0;
}
bool mixture_emission_fast(const dnbc_ctx *c) {
  const Dims &d = c->dm;
  const int KU = emit_px_ku(c);
  (void)KU;                                                  // (Kmax slots are the upper bound of both instantiations)
  return d.Kmax <= 8 && d.C <= 64 && emit_px_smem(d, d.Kmax, 32 / lane_group(d, false)) <= 200 * 1024;
}
template <int KU, bool BIGP = false>
static void launch_emit_px_t(dnbc_ctx *c, const EmitArgs &a, dim3 grid, size_t smem) {
  if (smem > 32 * 1024) cudaFuncSetAttribute(k_emission_px<KU, BIGP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // (static + dynamic shared memory count against the 48 KB default: opt in with room to spare)
  k_emission_px<KU, BIGP><<<grid, PX_WARPS * 32, smem, c->stream>>>(a);
}

void launch_mixture_emission(dnbc_ctx *c, int64_t p0, int64_t np, float *out) {
  if (np <= 0) return;
  FamTimer ft(c, FAM_EMISSION);
  const Dims &d = c->dm;
  const int vec = (c->P % 4 == 0) && (p0 % 4 == 0);
  // lane group = the state count itself when it is below 32 (N = 10: three points x 10 states fill 30 lanes;
  // the power-of-two group of 16 would fill 20).  Lanes past the last whole group repeat the first points of
  // the next group and store the same values.
  const int LG = lane_group(d, false);
  EmitArgs a{d, c->tables(), c->data(), p0, (int)np, out, LG, vec};
  // (more than 64 staging pieces per group -- e.g. 16 lane groups of 2 states x 5 variables -- take the instantiation
  // with per-variable component counts, which computes piece offsets on the fly)
  const bool bigp = d.C * (32 / LG) > 64 || d.D * (32 / LG) > 64;
  const int KU = bigp ? 0 : emit_px_ku(c);
  const size_t smem_px = emit_px_smem(d, KU > 0 ? KU : d.Kmax, 32 / LG);
  const int gy = (d.N + 31) / 32;                            // slices of 32 states that hold a real state
  int per_sm = (int)((200 * 1024) / (smem_px + 1024));
  per_sm = per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm);
  const int64_t group = (int64_t)EPU * (32 / a.LG);
  int64_t gx = (int64_t)c->sm_count * per_sm / gy;
  const int64_t maxgx = ((np + group - 1) / group + 15) / 16;
  gx = gx > maxgx ? maxgx : gx;
  gx = gx < 1 ? 1 : gx;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  switch (KU) {
    case 1: launch_emit_px_t<1>(c, a, grid, smem_px); break;
    case 2: launch_emit_px_t<2>(c, a, grid, smem_px); break;
    case 3: launch_emit_px_t<3>(c, a, grid, smem_px); break;
    case 4: launch_emit_px_t<4>(c, a, grid, smem_px); break;
    case 5: launch_emit_px_t<5>(c, a, grid, smem_px); break;
    case 6: launch_emit_px_t<6>(c, a, grid, smem_px); break;
    case 7: launch_emit_px_t<7>(c, a, grid, smem_px); break;
    case 8: launch_emit_px_t<8>(c, a, grid, smem_px); break;
    default:
      if (bigp) launch_emit_px_t<0, true>(c, a, grid, smem_px);
      else launch_emit_px_t<0>(c, a, grid, smem_px);
      break;
  }
}

// GMM moments + categorical histograms + gamma sums in one sweep over the posteriors.
// State counts that are a multiple of 32 go through k_stats2.cu (whole 32-point blocks); k_stats_pk takes every
// other shape and the ragged tail.
static void launch_stats_pk(dnbc_ctx *c, int64_t p0, int64_t np, const float *GA);
void launch_mixture_stats(dnbc_ctx *c, int64_t p0, int64_t np) {
  if (np <= 0) return;
  FamTimer ft(c, FAM_STATS_GMM);
  int64_t done = 0;
#ifndef DNBC_NO_STATS2              // (A/B builds only: tools/build_variant.sh)
  if (stats2_supported(c)) done = launch_stats2(c, p0, np);
#endif
  if (done < np) {
    if (done) c->launches++;
    launch_stats_pk(c, p0 + done, np - done, c->GA + done * c->dm.N);
  }
}

static void launch_stats_pk(dnbc_ctx *c, int64_t p0, int64_t np, const float *GA) {
  const Dims &d = c->dm;
  StatsLayout sl = stats_layout(d);
  int units = d.C > d.D ? d.C : d.D;
  if (units < 1) units = 1;
  const int NW = units > DNBC_STATS_NW ? DNBC_STATS_NW : units;
  const int gz = (units + NW - 1) / NW, gy = (d.N + 31) / 32;
  StatArgs a{d, c->tables(), c->data(), p0, (int)np, GA, c->d_mu, c->d_stats + sl.gmm, c->d_stats + sl.cat,
             c->d_stats + sl.gamma, lane_group(d, false), NW, 4096};
  const size_t smem = (size_t)NW * (d.Vmax + 1) * 32 * sizeof(float);
  // 16 warps per SM (128 registers each): CTAs of NW warps, e.g. 3 x 5 warps when 5 variables
#ifdef DNBC_STATS_PER_SM
  const int per_sm = DNBC_STATS_PER_SM;
#else
  const int per_sm = 16 / NW > 8 ? 8 : 16 / NW;
#endif
  int64_t gx = (int64_t)c->sm_count * per_sm / ((int64_t)gy * gz);
  const int64_t maxgx = (np + 255) / 256;
  if (gx > maxgx) gx = maxgx;
  if (gx < 1) gx = 1;
  const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  const int KP = (d.Kmax + 1) / 2;
#define STATS_CASE(KP_)                                                                                              \
  if (KP == KP_) {                                                                                                   \
    if (smem > 32 * 1024) cudaFuncSetAttribute(k_stats_pk<KP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    k_stats_pk<KP_><<<grid, NW * 32, smem, c->stream>>>(a);                                                          \
    return;                                                                                                          \
  }
  STATS_CASE(1) STATS_CASE(2) STATS_CASE(3) STATS_CASE(4)
#undef STATS_CASE
}
