// k_fused.cu -- the whole E-step of one sequence in ONE kernel, for small state spaces (N <= 16).
//
// BASELINE configs 1-3 (README set, toy robot, the N = 10 mixture set) have 10-12 hidden states: a
// sequence's scaled alpha is T*N*4 = 8 KB, yet the general pipeline (emission -> forward -> backward ->
// statistics, five launches) round-trips E, alpha^, gamma through HBM: 240 B of traffic per timestep
// against 20 B of observations, and at 1000 sequences every launch is latency.  Here a lane group owns a
// sequence and everything between the observations and the sufficient statistics stays in shared
// memory (north_star item 2: "alpha and beta rows in registers and shared memory"):
//
//   phase E  emission       E_t(j) = sum_d log2 B_d[j][o_td] + sum_c log2 sum_k w phi      -> BT[t][j]
//            (reference arithmetic: emissionsSum, DynamicNaiveBayesianClassifier.scala:52-56, 79-83;
//             Edge.scala:69-75, 114-118; GaussianUtils.scala:43-45)
//   phase F  forward        alpha^_t(j) = b_t(j) sum_i alpha^_{t-1}(i) a_ij / c_t         -> AL[t][j],
//            BT[t][j] <- b_t(j) / c_t,  LL += ln c_t + m_t
//   phase B  backward       beta^, gamma_t = alpha^_t beta^_t (renormalised every step)    -> AL[t][j],
//            xi-stat row i accumulated in registers, pi-stat, gamma sums
//   phase S  statistics     GMM responsibilities / moments per variable (generalising EM1D.run's sums,
//            Tools/ExpectationMaximization1D.java:92-124) and categorical histograms (Edge.scala:46-51),
//            from gamma in shared memory
//
// Mapping: a warp is a CTA; its 32 lanes form PG = 32 / LG groups of LG lanes (LG = N when that packs more
// groups, else the padded N), lane = (sequence slot, state).  Vectors the recursions exchange are rows of
// the shared-memory arrays, read back as group-wide broadcasts: no shuffles, no reductions trees -- every
// lane sums the row it has to read anyway.  The transition matrix lives in registers (column j for the
// forward pass, row i for the backward pass).  Phases E and S run variable by variable with that
// variable's mixture coefficients in registers; they are MUFU-bound like k_emission_px / k_stats2.
// fp32 accumulators (per sequence in registers, per warp in shared memory) are folded into the fp64
// statistics every ~4096 timesteps.  No reference counterpart; checked against the fp64
// oracle like the general pipeline, whose results it reproduces to fp32 rounding.
#include "dnbc_internal.cuh"
#include "packed_f32.cuh"

namespace {

constexpr float FHBIG = 1e18f;     // weight g / sum above which a point is re-evaluated with shifted exponents
constexpr float FLT_MIN_N = 1.17549435e-38f;

struct FusedArgs {
  Dims dm;
  DeviceTables tb;
  DataView dv;
  const int64_t *off;
  int64_t s0, s1, p0;
  const double *mu_old;
  double *stats;
  StatsLayout sl;
  int LG, PG;        // lanes per group, groups per warp
  int Tcap;          // longest sequence of the launch
  int rs;            // row stride of AL / BT in floats (even)
  int window;        // timesteps between fp64 flushes
  // byte offsets into dynamic shared memory
  int o_al, o_bt, o_xs, o_sy, o_cat, o_bins, o_gacc, o_zero;
  int sy_stride;     // bytes per (group, discrete variable) symbol row
  int xs_stride;     // floats per observation staging buffer (two buffers: variable c + 1 lands while c is evaluated)
};

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ u64 ex2_2(u64 a) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tex2.approx.ftz.f32 lo, lo;\n\tex2.approx.ftz.f32 hi, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}" : "=l"(d) : "l"(a));
  return d;
}
__device__ __forceinline__ float hsum(u64 a) {
  float lo, hi;
  unpk(a, lo, hi);
  return lo + hi;
}

// mixture coefficients of one (variable, state): K components as K / 2 packed pairs + an odd scalar tail
template <int K>
struct MixCoef {
  static constexpr int KP = K / 2;
  u64 s2[KP ? KP : 1], m2[KP ? KP : 1], c2[KP ? KP : 1];
  float s1, m1, c1;
};
template <int K>
__device__ __forceinline__ void load_coef(MixCoef<K> &q, const DeviceTables &tb, int c, int j, int Kmax, int Npad, bool on) {
  constexpr int KP = K / 2;
  auto at = [&](const float *tab, int k, float off_v) { return on ? tab[((int64_t)c * Kmax + k) * Npad + j] : off_v; };
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    q.s2[kp] = pk(at(tb.gs, 2 * kp, 0.f), at(tb.gs, 2 * kp + 1, 0.f));
    q.m2[kp] = pk(at(tb.gm, 2 * kp, 0.f), at(tb.gm, 2 * kp + 1, 0.f));
    // padded lanes get unit-mass components so that they never look like an underflowed mixture
    q.c2[kp] = pk(at(tb.gc, 2 * kp, 0.f), at(tb.gc, 2 * kp + 1, 0.f));
  }
  if (K & 1) {
    q.s1 = at(tb.gs, K - 1, 0.f);
    q.m1 = at(tb.gm, K - 1, 0.f);
    q.c1 = at(tb.gc, K - 1, 0.f);
  }
}
// component values e_k = 2^(c_k - z_k^2), their standardised coordinates and the mixture sum
template <int K>
struct MixEval {
  static constexpr int KP = K / 2;
  u64 z2[KP ? KP : 1], e2[KP ? KP : 1];
  float z1, e1;
};
template <int K>
__device__ __forceinline__ float mix_eval(const MixCoef<K> &q, float x, MixEval<K> &v) {
  constexpr int KP = K / 2;
  const u64 xx = pk(x, x);
  u64 sum2 = 0ull;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    v.z2[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
    v.e2[kp] = ex2_2(fnma2(v.z2[kp], v.z2[kp], q.c2[kp]));
    sum2 = kp == 0 ? v.e2[0] : add2(sum2, v.e2[kp]);
  }
  float sum = KP ? hsum(sum2) : 0.f;
  if (K & 1) {
    v.z1 = fmaf(x, q.s1, q.m1);
    v.e1 = ex2_approx(fmaf(-v.z1, v.z1, q.c1));
    sum += v.e1;
  }
  return sum;
}
// log2 of the mixture density with the exponents shifted by their maximum (every component underflowed fp32)
template <int K>
__device__ __forceinline__ float mix_log2_shifted(const MixCoef<K> &q, float x) {
  constexpr int KP = K / 2;
  const u64 xx = pk(x, x);
  float a[K > 0 ? K : 1];
  float mx = -INFINITY;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 z = fma2(xx, q.s2[kp], q.m2[kp]);
    unpk(fnma2(z, z, q.c2[kp]), a[2 * kp], a[2 * kp + 1]);
    mx = fmaxf(mx, fmaxf(a[2 * kp], a[2 * kp + 1]));
  }
  if (K & 1) {
    const float z = fmaf(x, q.s1, q.m1);
    a[K - 1] = fmaf(-z, z, q.c1);
    mx = fmaxf(mx, a[K - 1]);
  }
  if (mx == -INFINITY) return -INFINITY;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) s += ex2_approx(a[k] - mx);
  return mx + lg2_approx(s);
}

// moment accumulators of one (variable, state) over one sequence
template <int K>
struct MixAcc {
  static constexpr int KP = K / 2;
  u64 S0[KP ? KP : 1], S1[KP ? KP : 1], S2[KP ? KP : 1];
  float t0, t1, t2;
};
template <int K>
__device__ __forceinline__ void acc_add(MixAcc<K> &a, const MixEval<K> &v, float inv) {
  constexpr int KP = K / 2;
  const u64 inv2 = pk(inv, inv);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 rr = mul2(v.e2[kp], inv2);
    a.S0[kp] = add2(a.S0[kp], rr);
    const u64 rz = mul2(rr, v.z2[kp]);
    a.S1[kp] = add2(a.S1[kp], rz);
    a.S2[kp] = fma2(rz, v.z2[kp], a.S2[kp]);
  }
  if (K & 1) {
    const float r = v.e1 * inv, rz = r * v.z1;
    a.t0 += r;
    a.t1 += rz;
    a.t2 = fmaf(rz, v.z1, a.t2);
  }
}
// exponent-shifted contribution of one point (weight g; negative weights take a contribution back)
template <int K>
__device__ __forceinline__ void acc_add_shifted(MixAcc<K> &a, const MixCoef<K> &q, float x, float g) {
  constexpr int KP = K / 2;
  const u64 xx = pk(x, x);
  MixEval<K> v;
  float ar[K > 0 ? K : 1];
  float mx = -INFINITY;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    v.z2[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
    unpk(fnma2(v.z2[kp], v.z2[kp], q.c2[kp]), ar[2 * kp], ar[2 * kp + 1]);
    mx = fmaxf(mx, fmaxf(ar[2 * kp], ar[2 * kp + 1]));
  }
  if (K & 1) {
    v.z1 = fmaf(x, q.s1, q.m1);
    ar[K - 1] = fmaf(-v.z1, v.z1, q.c1);
    mx = fmaxf(mx, ar[K - 1]);
  }
  if (mx == -INFINITY) return;
  float sum = 0.f;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const float e0 = ex2_approx(ar[2 * kp] - mx), e1 = ex2_approx(ar[2 * kp + 1] - mx);
    v.e2[kp] = pk(e0, e1);
    sum += e0 + e1;
  }
  if (K & 1) {
    v.e1 = ex2_approx(ar[K - 1] - mx);
    sum += v.e1;
  }
  acc_add<K>(a, v, sum > 0.f ? g * rcp_approx(sum) : 0.f);
}

// ---- phase E for one variable: bt[t * rs] += log2 mixture density for t < T --------------------------------
template <int K>
__device__ __forceinline__ void emis_var(const FusedArgs &a, int c, int j, bool on, const float *xs, float *bt, int rs, int T) {
  MixCoef<K> q;
  load_coef<K>(q, a.tb, c, j, a.dm.Kmax, a.dm.Npad, on);
  float smin = INFINITY;
  // four timesteps per trip with every load ahead of the arithmetic: a lone warp per scheduler hides the MUFU
  // and shared-memory latencies only through independent chains (the read-modify-write of bt would otherwise
  // order the iterations: the compiler cannot tell bt from xs)
  int t = 0;
  for (; t + 4 <= T; t += 4) {
    const float4 x = *reinterpret_cast<const float4 *>(xs + t);
    float *b0 = bt + t * rs;
    const float o0 = b0[0], o1 = b0[rs], o2 = b0[2 * rs], o3 = b0[3 * rs];
    MixEval<K> v;
    const float s0 = mix_eval<K>(q, x.x, v), s1 = mix_eval<K>(q, x.y, v), s2 = mix_eval<K>(q, x.z, v), s3 = mix_eval<K>(q, x.w, v);
    smin = fminf(fminf(smin, fminf(s0, s1)), fminf(s2, s3));
    b0[0] = o0 + (s0 < 1e-30f ? 0.f : lg2_approx(s0));      // underflowed points are added below
    b0[rs] = o1 + (s1 < 1e-30f ? 0.f : lg2_approx(s1));
    b0[2 * rs] = o2 + (s2 < 1e-30f ? 0.f : lg2_approx(s2));
    b0[3 * rs] = o3 + (s3 < 1e-30f ? 0.f : lg2_approx(s3));
  }
  for (; t < T; ++t) {
    MixEval<K> v;
    const float sum = mix_eval<K>(q, xs[t], v);
    smin = fminf(smin, sum);
    bt[t * rs] += sum < 1e-30f ? 0.f : lg2_approx(sum);
  }
  if (smin < 1e-30f) {
    // rare: some point lies > ~13 sigma from every component of this state (or no component has mass): the fp32
    // sum underflowed.  Such points take the exponent-shifted evaluation, which stays finite where the reference's
    // fp64 sum is finite and gives -inf where no component has mass.
#pragma unroll 1
    for (int u = 0; u < T; ++u) {
      MixEval<K> v;
      if (mix_eval<K>(q, xs[u], v) < 1e-30f) bt[u * rs] += mix_log2_shifted<K>(q, xs[u]);
    }
  }
}
// ---- phase S for one variable: moments of one sequence, then into the warp's shared accumulators ------------
// ga / gs: this lane's gamma column and its stride (a zero word with stride 0 for padding lanes); rows t >= T of a
// slot are zero, so the sweep runs to the warp-uniform Tmax without predicates.
template <int K>
__device__ __forceinline__ void stat_var(const FusedArgs &a, int c, int j, bool on, const float *xs, const float *ga, int gs, int Tmax,
                                         float *gacc) {
  constexpr int KP = K / 2;
  MixCoef<K> q;
  load_coef<K>(q, a.tb, c, j, a.dm.Kmax, a.dm.Npad, on);
  MixAcc<K> acc;
#pragma unroll
  for (int kp = 0; kp < (KP ? KP : 1); ++kp) acc.S0[kp] = acc.S1[kp] = acc.S2[kp] = 0ull;
  acc.t0 = acc.t1 = acc.t2 = 0.f;
  float hmax = 0.f;
  int t = 0;
  for (; t + 4 <= Tmax; t += 4) {                            // (four independent chains per trip, loads first: see emis_var)
    const float4 x = *reinterpret_cast<const float4 *>(xs + t);
    const float *gp = ga + t * gs;
    const float g0 = gp[0], g1 = gp[gs], g2 = gp[2 * gs], g3 = gp[3 * gs];
    MixEval<K> v0, v1, v2, v3;
    const float s0 = mix_eval<K>(q, x.x, v0), s1 = mix_eval<K>(q, x.y, v1), s2 = mix_eval<K>(q, x.z, v2), s3 = mix_eval<K>(q, x.w, v3);
    const float i0 = g0 * rcp_approx(fmaxf(s0, FLT_MIN_N)), i1 = g1 * rcp_approx(fmaxf(s1, FLT_MIN_N));
    const float i2 = g2 * rcp_approx(fmaxf(s2, FLT_MIN_N)), i3 = g3 * rcp_approx(fmaxf(s3, FLT_MIN_N));
    hmax = fmaxf(fmaxf(hmax, fmaxf(i0, i1)), fmaxf(i2, i3));
    acc_add<K>(acc, v0, i0);
    acc_add<K>(acc, v1, i1);
    acc_add<K>(acc, v2, i2);
    acc_add<K>(acc, v3, i3);
  }
  for (; t < Tmax; ++t) {
    MixEval<K> v;
    const float sum = mix_eval<K>(q, xs[t], v);
    const float inv = ga[t * gs] * rcp_approx(fmaxf(sum, FLT_MIN_N));
    hmax = fmaxf(hmax, inv);
    acc_add<K>(acc, v, inv);
  }
  if (__any_sync(0xffffffffu, hmax > FHBIG)) {
    // rare: posterior mass at a point where fp32 has flushed (part of) the state's mixture: take the fast
    // contribution back (same instructions, weight -g) and add the exponent-shifted evaluation
#pragma unroll 1
    for (int u = 0; u < Tmax; ++u) {
      const float x = xs[u], g = ga[u * gs];
      MixEval<K> v;
      const float sum = mix_eval<K>(q, x, v);
      const float inv = g * rcp_approx(fmaxf(sum, FLT_MIN_N));
      if (inv > FHBIG) {
        acc_add<K>(acc, v, -inv);
        acc_add_shifted<K>(acc, q, x, g);
      }
    }
  }
  // this lane's column of the warp's accumulators: gacc[((c * Kmax + k) * 3 + m) * 32 + lane]
  float *dst = gacc + (size_t)c * a.dm.Kmax * 96;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    float v0[2], v1[2], v2[2];
    unpk(acc.S0[kp], v0[0], v0[1]); unpk(acc.S1[kp], v1[0], v1[1]); unpk(acc.S2[kp], v2[0], v2[1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float *d = dst + (2 * kp + h) * 96;
      d[0] += v0[h]; d[32] += v1[h]; d[64] += v2[h];
    }
  }
  if (K & 1) {
    float *d = dst + (K - 1) * 96;
    d[0] += acc.t0; d[32] += acc.t1; d[64] += acc.t2;
  }
}

#define FUSED_SWITCH_K(Kc, CALL)            \
  switch (Kc) {                             \
    case 1: { constexpr int KK = 1; CALL; } break; \
    case 2: { constexpr int KK = 2; CALL; } break; \
    case 3: { constexpr int KK = 3; CALL; } break; \
    case 4: { constexpr int KK = 4; CALL; } break; \
    case 5: { constexpr int KK = 5; CALL; } break; \
    case 6: { constexpr int KK = 6; CALL; } break; \
    case 7: { constexpr int KK = 7; CALL; } break; \
    default: { constexpr int KK = 8; CALL; } break; \
  }

// per-lane view of the warp's shared memory and of the sequence slot the lane works for
// (byte offsets into the dynamic shared memory, not pointers: the phase functions rebuild their pointers from the
// __shared__ base so that the compiler emits LDS / STS instead of generic loads)
struct Slot {
  int AL, BT, xs;            // this slot's rows / staging buffers
  int sy;
  int cat, bins, gacc;       // this lane's columns
  int zero;
  int lane, lj, j, T, Tmax;
  int64_t pt;
  bool ingrp, on;
};

// observations of variable c for the slot's sequence -> staging buffer `buf` (asynchronous; zero past the end)
__device__ __forceinline__ void stage_x(const FusedArgs &a, const Slot &s, int c, int buf) {
  extern __shared__ __align__(16) unsigned char sm[];
  if (!s.ingrp) return;
  float *dst = reinterpret_cast<float *>(sm + s.xs) + buf * a.xs_stride;
  const float *src = a.dv.cont + (int64_t)c * a.dv.P + s.pt;
  for (int t = s.lj; t < s.Tmax; t += a.LG) {
    if (t < s.T) cp_async4(dst + t, src + t);
    else dst[t] = 0.f;
  }
}

// phases E and S are the same for every state count: compiled once, called from each kernel instantiation
__device__ __noinline__ void phase_emission(const FusedArgs &a, const Slot &s) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, rs = a.rs;
  float *sBT = reinterpret_cast<float *>(sm + s.BT);
  const float *sxs = reinterpret_cast<const float *>(sm + s.xs), *scat = reinterpret_cast<const float *>(sm + s.cat);
  uint8_t *ssy = sm + s.sy;
  if (C > 0) stage_x(a, s, 0, 0);
  for (int d = 0; d < D; ++d) {                              // the sequence's symbols, every discrete variable
    const uint8_t *src = a.dv.disc + (int64_t)d * a.dv.P + s.pt;
#pragma unroll 4
    for (int t = s.lj; t < s.T; t += a.LG) {
      const int v = (int)src[t];
      ssy[d * a.sy_stride + t] = (uint8_t)(v < Vmax ? v : Vmax);
    }
  }
  __syncwarp();
  if (s.on)
    for (int t = 0; t < s.T; ++t) {
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc += scat[(d * (Vmax + 1) + ssy[d * a.sy_stride + t]) * 32];
      sBT[t * rs + s.lj] = acc;
    }
  for (int c = 0; c < C; ++c) {
    cp_async_commit_wait();
    __syncwarp();                                            // variable c has landed; the other buffer is free
    if (c + 1 < C) stage_x(a, s, c + 1, (c + 1) & 1);
    const int Kc = a.tb.K[c];
    const float *xs = sxs + (c & 1) * a.xs_stride;
    FUSED_SWITCH_K(Kc, emis_var<KK>(a, c, s.j, s.on, xs, sBT + (s.on ? s.lj : 0), rs, s.on ? s.T : 0));
  }
  __syncwarp();
}

__device__ __noinline__ void phase_statistics(const FusedArgs &a, const Slot &s) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, rs = a.rs;
  const float *sAL = reinterpret_cast<const float *>(sm + s.AL), *sxs = reinterpret_cast<const float *>(sm + s.xs);
  float *sbins = reinterpret_cast<float *>(sm + s.bins), *sgacc = reinterpret_cast<float *>(sm + s.gacc);
  const uint8_t *ssy = sm + s.sy;
  if (C > 0) stage_x(a, s, 0, 0);
  if (s.on)
    for (int d = 0; d < D; ++d)
      for (int t = 0; t < s.T; ++t) sbins[(d * (Vmax + 1) + ssy[d * a.sy_stride + t]) * 32] += sAL[t * rs + s.lj];
  const float *ga = s.on ? sAL + s.lj : reinterpret_cast<const float *>(sm + s.zero);
  const int gs = s.on ? rs : 0;
  for (int c = 0; c < C; ++c) {
    cp_async_commit_wait();
    __syncwarp();
    if (c + 1 < C) stage_x(a, s, c + 1, (c + 1) & 1);
    const int Kc = a.tb.K[c];                                // (every lane takes part: the sweep ends in a warp vote)
    const float *xs = sxs + (c & 1) * a.xs_stride;
    FUSED_SWITCH_K(Kc, stat_var<KK>(a, c, s.j, s.on, xs, ga, gs, s.Tmax, sgacc));
  }
  __syncwarp();
}

// N2 = pairs of states per row (N <= 2 N2, rows are 2 N2 floats with a zero pad entry for odd N)
template <int N2>
__global__ void __launch_bounds__(32) k_estep_fused(const __grid_constant__ FusedArgs a) {
  extern __shared__ __align__(16) unsigned char sm[];
  constexpr int rs = 2 * N2;
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax, Npad = a.dm.Npad;
  const int LG = a.LG, PG = a.PG, Tcap = a.Tcap;
  const int lane = threadIdx.x, gi = lane / LG, lj = lane % LG;
  Slot s;
  s.lane = lane; s.lj = lj;
  s.ingrp = gi < PG;                                         // lanes past the last whole group idle
  s.on = s.ingrp && lj < N;                                  // a real state of a sequence slot
  const bool on = s.on;
  const int g = s.ingrp ? gi : 0;
  s.AL = a.o_al + g * Tcap * rs * 4;
  s.BT = a.o_bt + g * Tcap * rs * 4;
  s.xs = a.o_xs + g * ((Tcap + 3) & ~3) * 4;
  s.sy = a.o_sy + g * D * a.sy_stride;
  s.cat = a.o_cat + lane * 4;                                // [D][Vmax+1][32] log2 P(v | state)
  s.bins = a.o_bins + lane * 4;                              // [D][Vmax+1][32]
  s.gacc = a.o_gacc + lane * 4;                              // [C][Kmax][3][32]
  s.zero = a.o_zero;
  s.j = on ? lj : 0;
  const int j = s.j;
  float *AL = reinterpret_cast<float *>(sm + s.AL), *BT = reinterpret_cast<float *>(sm + s.BT);
  float *scat = reinterpret_cast<float *>(sm + s.cat), *sbins = reinterpret_cast<float *>(sm + s.bins);
  float *sgacc = reinterpret_cast<float *>(sm + s.gacc);

  for (int q = 0; q < D * (Vmax + 1); ++q) {
    scat[q * 32] = on ? a.tb.cat[(int64_t)q * Npad + j] : 0.f;
    sbins[q * 32] = 0.f;
  }
  for (int q = 0; q < C * Kmax * 3; ++q) sgacc[q * 32] = 0.f;
  {
    // pad columns (odd N) of the row arrays are read by the float2 loops and must stay zero
    float *al0 = reinterpret_cast<float *>(sm + a.o_al), *bt0 = reinterpret_cast<float *>(sm + a.o_bt);
    for (int q = lane; q < PG * Tcap * rs; q += 32) al0[q] = bt0[q] = 0.f;
    if (lane < 4) reinterpret_cast<float *>(sm + a.o_zero)[lane] = 0.f;
  }

  // transition matrix: column j (forward) and row j (backward) of this lane's state; zero for padding
  float Acol[rs], Arow[rs], X[rs];
#pragma unroll
  for (int i = 0; i < rs; ++i) {
    Acol[i] = (on && i < N) ? a.tb.A[i * Npad + j] : 0.f;
    Arow[i] = (on && i < N) ? a.tb.A[j * Npad + i] : 0.f;
    X[i] = 0.f;
  }
  const float pi_j = on ? a.tb.pi[j] : 0.f;
  double ll = 0.0, piacc = 0.0, gsacc = 0.0;
  int inwin = 0;
  __syncwarp();

  auto flush = [&]() {
    if (on) {
      for (int c = 0; c < C; ++c) {
        const int K = a.tb.K[c];
        for (int k = 0; k < K; ++k) {
          float *src = sgacc + ((size_t)c * Kmax + k) * 96;
          const float a0 = src[0], a1 = src[32], a2 = src[64];
          src[0] = src[32] = src[64] = 0.f;
          if (a0 != 0.f) {
            // z = (x - mu') s with mu' = -m / s the fp32-rounded centre; re-centre on the fp64 old mean
            const int64_t o = ((int64_t)c * Kmax + k) * Npad + j;
            const int64_t dst = ((int64_t)c * N + j) * Kmax + k;
            const double sc = (double)a.tb.gs[o], m = (double)a.tb.gm[o];
            double m0 = (double)a0, m1 = 0.0, m2 = 0.0;
            if (sc > 0.0) {
              const double delta = -m / sc - a.mu_old[dst];
              const double c1 = (double)a1 / sc, c2 = (double)a2 / (sc * sc);
              m1 = c1 + delta * m0;
              m2 = c2 + 2.0 * delta * c1 + delta * delta * m0;
            }
            atomicAdd(a.stats + a.sl.gmm + 3 * dst + 0, m0);
            atomicAdd(a.stats + a.sl.gmm + 3 * dst + 1, m1);
            atomicAdd(a.stats + a.sl.gmm + 3 * dst + 2, m2);
          }
        }
      }
      for (int d = 0; d < D; ++d)
        for (int v = 0; v < Vmax; ++v) {
          float *b = sbins + (d * (Vmax + 1) + v) * 32;
          if (*b != 0.f) atomicAdd(a.stats + a.sl.cat + ((int64_t)d * N + j) * Vmax + v, (double)*b);
          *b = 0.f;
        }
#pragma unroll
      for (int i = 0; i < rs; ++i) {
        if (i < N && X[i] != 0.f) atomicAdd(a.stats + a.sl.xi + (int64_t)j * N + i, (double)X[i] * (double)Arow[i]);
        X[i] = 0.f;
      }
      if (gsacc != 0.0) atomicAdd(a.stats + a.sl.gamma + j, gsacc);
      if (piacc != 0.0) atomicAdd(a.stats + a.sl.pi + j, piacc);
      gsacc = piacc = 0.0;
    }
    for (int d = 0; d < D; ++d) sbins[(d * (Vmax + 1) + Vmax) * 32] = 0.f;     // unseen-symbol row is discarded
    inwin = 0;
  };

  const int64_t nbatch = (a.s1 - a.s0 + PG - 1) / PG;
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int64_t sq = a.s0 + batch * PG + g;
    const bool have = s.ingrp && sq < a.s1;
    const int T = have ? (int)(a.off[sq + 1] - a.off[sq]) : 0;
    s.T = T;
    s.pt = have ? a.off[sq] : 0;                             // global point index of the sequence
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    s.Tmax = Tmax;
    if (Tmax == 0) continue;

    // ---------------- phase E: log2 emissions into BT ----------------
    phase_emission(a, s);

    // ---------------- emission factors: b_t(j) = 2^(E_t(j) - max_j E_t(j)) in place, four rows per barrier ----------------
    // (independent across t: none of this sits on the recursions' dependency chains)
    float msum = 0.f;
    for (int t0 = 0; t0 < Tmax; t0 += 4) {
      float bq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = t0 + q;
        const bool step = t < T;
        const float2 *er = reinterpret_cast<const float2 *>(BT + (size_t)(step ? t : 0) * rs);
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < N2; ++i) {
          const float2 e = er[i];
          m0 = fmaxf(m0, e.x);
          if (2 * i + 1 < N) m1 = fmaxf(m1, e.y);           // (the pad entry of an odd N holds 0, not a log)
        }
        const float mx = fmaxf(m0, m1);
        const float m = mx == -INFINITY ? 0.f : mx;
        bq[q] = ex2_approx(BT[(size_t)(step ? t : 0) * rs + lj] - m);
        msum += step ? m : 0.f;
      }
      __syncwarp();                                          // every lane has read the four rows
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (on && t0 + q < T) BT[(size_t)(t0 + q) * rs + lj] = bq[q];
    }
    __syncwarp();

    // ---------------- phase F: scaled forward recursion ----------------
    // alpha~_t(j) = b_t(j) acc_t(j), acc_0 = pi, acc_t(j) = sum_i alpha~_{t-1}(i) a_ij / S_{t-1}, S_t = sum_j alpha~_t(j) = c_t.
    // Row t - 1 is read by every lane of the group (dot product and S_{t-1} in one pass); afterwards each lane
    // normalises its own entry of that row (AL <- alpha^, BT <- b / c) and writes alpha~_t.  A lone warp per
    // scheduler exposes every instruction's latency, so the loop carries no branches: `step` / `prev` only
    // predicate the stores, and the log-likelihood terms are summed in fp32 per sequence (<= 2^-17 relative).
    {
      float bprev = on && T > 0 ? BT[lj] : 0.f;
      float aprev = bprev * pi_j, lls = 0.f;
      if (on && T > 0) AL[lj] = aprev;
      __syncwarp();
      for (int t = 1; t <= Tmax; ++t) {
        const bool step = t < T, prev = t <= T;
        const float2 *ar = reinterpret_cast<const float2 *>(AL + (size_t)(prev ? t - 1 : 0) * rs);
        const float b = BT[(size_t)(step ? t : 0) * rs + lj];
        float d0 = 0.f, d1 = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < N2; ++i) {
          const float2 v = ar[i];                            // (the pad entry of an odd N is kept at zero)
          d0 = fmaf(v.x, Acol[2 * i], d0);
          d1 = fmaf(v.y, Acol[2 * i + 1], d1);
          s0 += v.x;
          s1 += v.y;
        }
        const float Sprev = s0 + s1;
        const float inv = rcp_approx(Sprev);
        const float at = b * ((d0 + d1) * inv);
        lls += prev ? lg2_approx(Sprev) : 0.f;
        __syncwarp();                                        // every lane has read row t - 1
        if (prev && on) {
          AL[(size_t)(t - 1) * rs + lj] = aprev * inv;       // alpha^_{t-1}
          BT[(size_t)(t - 1) * rs + lj] = bprev * inv;       // b_{t-1} / c_{t-1}
        }
        if (step && on) AL[(size_t)t * rs + lj] = at;
        aprev = at;
        bprev = b;
        __syncwarp();
      }
      if (have && lj == 0) ll += ((double)lls + (double)msum) * DNBC_LN2;
    }

    // ---------------- phase B: backward recursion, gamma, xi, pi ----------------
    {
      float beta = 1.f, gseq = 0.f;
      for (int k = 0; k < Tmax; ++k) {
        const int t = T - 1 - k;                             // slots are aligned at their last step
        const bool step = t >= 0, inner = step && k > 0;
        const int tc = step ? t : 0, tn = inner ? t + 1 : 0;
        const float al = on ? AL[(size_t)tc * rs + lj] : 0.f;
        if (inner && on) BT[(size_t)tn * rs + lj] *= beta;   // u_{t+1}(j) = b_{t+1}(j) beta^_{t+1}(j) / c_{t+1}
        __syncwarp();
        const float2 *ur = reinterpret_cast<const float2 *>(BT + (size_t)tn * rs);
        float d0 = 0.f, d1 = 0.f;
        const float alx = inner ? al : 0.f;                  // (no transition statistics from the last step)
#pragma unroll
        for (int i = 0; i < N2; ++i) {
          const float2 u = ur[i];
          d0 = fmaf(u.x, Arow[2 * i], d0);
          d1 = fmaf(u.y, Arow[2 * i + 1], d1);
          X[2 * i] = fmaf(alx, u.x, X[2 * i]);               // xi-stat row: alpha^_t(i) u_{t+1}(j), times a_ij at the flush
          X[2 * i + 1] = fmaf(alx, u.y, X[2 * i + 1]);
        }
        const float bt = inner ? d0 + d1 : 1.f;
        // gamma_t = alpha^_t beta^_t, renormalised: sum_i alpha^_t(i) beta^_t(i) is 1 only in exact arithmetic
        const float p = al * bt;
        if (step && on) AL[(size_t)tc * rs + lj] = p;
        __syncwarp();
        const float2 *pr = reinterpret_cast<const float2 *>(AL + (size_t)tc * rs);
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int i = 0; i < N2; ++i) { const float2 v = pr[i]; g0 += v.x; g1 += v.y; }
        const float G = g0 + g1;
        const float fix = G > 0.f ? rcp_approx(G) : 1.f;
        beta = step ? bt * fix : beta;
        const float gm = step ? p * fix : 0.f;
        __syncwarp();                                        // every lane has read the row
        if (step && on) AL[(size_t)tc * rs + lj] = gm;       // gamma_t(j) for phase S
        gseq += on ? gm : 0.f;
        if (t == 0) piacc += on ? (double)gm : 0.0;
      }
      gsacc += (double)gseq;
      // rows past this slot's end may hold an earlier, longer sequence's posteriors: phase S runs to Tmax
      if (on)
        for (int t = T; t < Tmax; ++t) AL[(size_t)t * rs + lj] = 0.f;
    }
    __syncwarp();

    // ---------------- phase S: sufficient statistics from gamma ----------------
    phase_statistics(a, s);
    inwin += Tmax;
    if (inwin >= a.window) flush();
  }
  flush();
  if (ll != 0.0) atomicAdd(a.stats, ll);
}

struct FusedPlan {
  FusedArgs a;
  size_t smem;
  int N2;
};

// shared-memory plan for the resident data; smem == 0 when the shape does not fit
FusedPlan fused_plan(const dnbc_ctx *c) {
  FusedPlan p{};
  const Dims &d = c->dm;
  if (d.Npad > 16 || d.Kmax > 8 || c->Tmax < 1) return p;
  const int pg_pad = 32 / d.Npad, pg_exact = 32 / d.N;
  const int LG = pg_exact > pg_pad ? d.N : d.Npad, PG = 32 / LG;
  p.N2 = (d.N + 1) / 2;
  const int rs = 2 * p.N2;
  const int Tcap = (int)c->Tmax;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~(size_t)15; return (int)at; };
  FusedArgs &a = p.a;
  a.LG = LG; a.PG = PG; a.rs = rs; a.Tcap = Tcap; a.window = 4096;
  a.sy_stride = (Tcap + 3) & ~3;
  a.xs_stride = PG * ((Tcap + 3) & ~3);
  a.o_al = take((size_t)PG * Tcap * rs * sizeof(float));
  a.o_bt = take((size_t)PG * Tcap * rs * sizeof(float));
  a.o_xs = take((size_t)2 * a.xs_stride * sizeof(float));
  a.o_sy = take((size_t)PG * d.D * a.sy_stride);
  a.o_cat = take((size_t)d.D * (d.Vmax + 1) * 32 * sizeof(float));
  a.o_bins = take((size_t)d.D * (d.Vmax + 1) * 32 * sizeof(float));
  a.o_gacc = take((size_t)d.C * d.Kmax * 3 * 32 * sizeof(float));
  a.o_zero = take(16);
  if (o > 200 * 1024) return p;
  p.smem = o;
  return p;
}

}  // namespace

bool fused_estep_supported(const dnbc_ctx *c) { return fused_plan(c).smem != 0; }

// One warp per scheduler exposes every latency, so the fused kernel wins where the general pipeline's five launches
// are latency-bound themselves and loses to it on throughput beyond.  The choice compares two fitted costs (B200,
// 250 .. 16,000 sequences of three shapes, profiles/r02h_fused_crossover.jsonl): a wave of the fused kernel takes
// T x (0.54 + 0.027 sum_c K_c + 0.16 D) microseconds, the pipeline 270 microseconds of launch latency plus 0.25 ns per
// point.  Measured gains where it is chosen: 1.3 - 1.5x at the config-3 shape up to 1000 sequences, 1.2 - 1.8x at the
// toy-robot shape up to 4000; the README set (5 + 5 variables, 1000 sequences) is a tie and stays on the pipeline.
bool fused_estep_wanted(const dnbc_ctx *c, int64_t n_seq) {
  if (c->opt_fused_estep == 2) return false;
  const FusedPlan p = fused_plan(c);
  if (p.smem == 0) return false;
  if (c->opt_fused_estep == 1) return true;
  int per_sm = (int)((size_t)(227 * 1024) / (p.smem + 1024));
  if (per_sm > 16) per_sm = 16;
  if (per_sm < 1) per_sm = 1;
  const int64_t nbatch = (n_seq + p.a.PG - 1) / p.a.PG, capacity = (int64_t)c->sm_count * per_sm;
  const double waves = (double)((nbatch + capacity - 1) / capacity);
  int sumK = 0;
  for (int v = 0; v < c->dm.C; ++v) sumK += c->K[v];
  const double t_fused = waves * (double)c->Tmax * (0.54 + 0.027 * sumK + 0.16 * c->dm.D);
  const double t_pipeline = 270.0 + 0.25e-3 * (double)n_seq * (double)c->Tmax;
  return t_fused < t_pipeline;
}

// E-step of sequences [s0, s1) (points from p0) into the statistics buffer: one launch
void launch_estep_fused(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0) {
  if (s1 <= s0) return;
  FusedPlan p = fused_plan(c);
  FusedArgs &a = p.a;
  a.dm = c->dm;
  a.tb = c->tables();
  a.dv = c->data();
  a.off = c->d_off;
  a.s0 = s0; a.s1 = s1; a.p0 = p0;
  a.mu_old = c->d_mu;
  a.stats = c->d_stats;
  a.sl = stats_layout(c->dm);
  FamTimer ft(c, FAM_FORWARD);
  const int64_t nbatch = (s1 - s0 + a.PG - 1) / a.PG;
  int per_sm = (int)((size_t)(227 * 1024) / (p.smem + 1024));
  if (per_sm > 16) per_sm = 16;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)c->sm_count * per_sm;
  if (grid > nbatch) grid = nbatch;
#define FUSED_CASE(N2_)                                                                                   \
  case N2_:                                                                                               \
    cudaFuncSetAttribute(k_estep_fused<N2_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);  \
    k_estep_fused<N2_><<<(unsigned)grid, 32, p.smem, c->stream>>>(a);                                    \
    break;
  switch (p.N2) {
    FUSED_CASE(1) FUSED_CASE(2) FUSED_CASE(3) FUSED_CASE(4) FUSED_CASE(5) FUSED_CASE(6) FUSED_CASE(7) FUSED_CASE(8)
  }
#undef FUSED_CASE
}
