// k_stats2.cu -- K3 statistics sweep, second generation, for state counts from 32 up that are a multiple of 4
// (the scale-out shape N = 64, the large-state shape N = 512; a last slice of fewer than 32 states idles its spare
// lanes on zero posteriors) and, in a grouped layout, for 8 <= N < 32.
//
// Same mathematics and the same fp32-window / fp64-flush scheme as k_stats_pk (k_mixture.cu): GMM
// responsibilities and moments (generalising the E- and M-step sums of EM1D.run,
// core/src/main/java/Tools/ExpectationMaximization1D.java:92-124), categorical histograms
// (DiscreteEdge.learn, core/src/main/scala/Edge.scala:46-51) and the gamma sums.
//
// Why a second kernel: ncu on k_stats_pk<4> (profiles/r01f_ncu_summary.txt) shows 73 warp instructions
// per (point, variable, 32 states) against 74 MUFU cycles: the sweep is ISSUE-bound, and 18 of the 73
// are register moves, 64-bit address arithmetic and predicate bookkeeping around the posterior
// prefetch (8 MOV, 3.7 IADD3, 3 FSETP, 1.4 LEA, 1 IMAD, 1 ISETP).  Here nothing is carried in
// registers across points except the coefficients and the accumulators:
//   * the posteriors of a 32-point x 32-state block arrive in shared memory through cp.async
//     (8 x 16 B per lane per block, double buffered, one block ahead) and are read back with one
//     conflict-free LDS per point: no prefetch registers, no rotation moves, no long-scoreboard stalls;
//   * the observation of a point is warp-uniform: the lane that loaded it stores the record
//     {x, x, 128 * symbol} once per block and every point costs one broadcast LDS.128 that yields the
//     packed pair (x, x) and the histogram row offset (no shuffle, no duplicate move);
//   * blocks are always whole (the launcher gives the ragged tail to k_stats_pk), lanes are always
//     real states: no predicates in the loop.
// Per point and warp at K = 8: 27 packed fp32 + 9 MUFU + ~10 other instructions.
#include "dnbc_internal.cuh"
#include "packed_f32.cuh"

namespace {

#ifndef S2_NW
#define S2_NW 8
#endif
#ifndef S2_WPSM
#define S2_WPSM 16               // resident warps per SM the grid is sized for
#endif
constexpr int NW_MAX = S2_NW;    // warps per CTA, one observed variable each
#ifndef S2_PTS
#define S2_PTS 0                 // points per trip of the block loop; 0 = by mixture size (A/B builds set it)
#endif

struct Stat2Args {
  Dims dm;
  DeviceTables tb;
  DataView dv;
  int64_t p0;
  int nblk;                      // whole 32-point blocks of this launch
  const float *GA;
  const double *mu_old;
  double *gmm, *cat, *gsum;
  int window;                    // points per fp32 accumulation window (multiple of 32)
  int nw;                        // warps per CTA
  int LG, PG;                    // grouped layout (N <= 32): lanes per point, points per warp step (32 / LG)
};

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
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ODD: the mixtures have 2 KP + 1 components and the last one travels alone in scalar registers (K = 3, 5, 7 with the
// same K for every variable): a padded pair would cost an exponential and seven packed instructions per point for nothing.
template <int KP, bool ODD>
struct Coef {
  u64 s2[KP], m2[KP], cc2[KP];   // z = x s + m ; exponent = cc - z^2 (log2 units), two components per register pair
  float s1, m1, c1;
};
template <int KP, bool ODD>
struct Acc {
  u64 S0[KP], S1[KP], S2[KP];    // sum r, sum r z, sum r z^2 over the current window
  float t0, t1, t2;
};
#ifndef S2_TTRICK
#define S2_TTRICK 1
#endif
template <int KP, bool ODD>
struct Pend {
  u64 z[KP], e[KP];
#if S2_TTRICK
  u64 a[KP];                     // exponent arguments c' - z^2 (see stage2)
#endif
  float z1, e1, a1;
};

// stage 1 of a point: standardised coordinate and the 2 KP exponentials
template <int KP, bool ODD>
__device__ __forceinline__ void stage1(const Coef<KP, ODD> &q, u64 xx, Pend<KP, ODD> &n) {
  if (ODD) {
    float x, xhi;
    unpk(xx, x, xhi);
    n.z1 = fmaf(x, q.s1, q.m1);
    n.a1 = fmaf(-n.z1, n.z1, q.c1);
    n.e1 = ex2_approx(n.a1);
  }
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    n.z[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
#if S2_TTRICK
    n.a[kp] = fnma2(n.z[kp], n.z[kp], q.cc2[kp]);
    n.e[kp] = ex2_2(n.a[kp]);
#else
    n.e[kp] = ex2_2(fnma2(n.z[kp], n.z[kp], q.cc2[kp]));
#endif
  }
}
// mixture sum of a pending point (the underflow decision is taken on exactly this value, here and in the redo)
template <int KP, bool ODD>
__device__ __forceinline__ float mix_sum(const Pend<KP, ODD> &p) {
  u64 sum2 = p.e[0];
#pragma unroll
  for (int kp = 1; kp < KP; ++kp) sum2 = add2(sum2, p.e[kp]);
  return ODD ? hsum(sum2) + p.e1 : hsum(sum2);
}
// stage 2: normalise by the mixture sum, weight by the posterior, accumulate the three moments.
// No predicates: the sum is clamped to the smallest normal float, so the weight g / sum is finite and a mixture
// whose every component flushed to zero contributes exactly nothing.  The weight itself is the underflow detector:
// the caller keeps its running maximum and a value above HBIG (sum < 1e-18 g: the state carries posterior mass
// at a point ~9 sigma away from all of its components, where fp32 has flushed part of the mixture) sends the
// block through the exponent-shifted re-evaluation.
constexpr float HBIG = 1e18f;
template <int KP, bool ODD>
__device__ __forceinline__ float stage2(Acc<KP, ODD> &a, const Pend<KP, ODD> &p, float g) {
  const float sum = mix_sum<KP, ODD>(p);
  const float inv = g * rcp_approx(fmaxf(sum, 1.17549435e-38f));
  const u64 inv2 = pk(inv, inv);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 rr = mul2(p.e[kp], inv2);
    a.S0[kp] = add2(a.S0[kp], rr);
#if S2_TTRICK
    // z^2 = c' - arg, so sum r z^2 = c' S0 - sum r arg: the second moment costs one FMA on the argument the
    // exponential was taken of (a.S2 holds sum r arg here; the flush converts)
    a.S1[kp] = fma2(rr, p.z[kp], a.S1[kp]);
    a.S2[kp] = fma2(rr, p.a[kp], a.S2[kp]);
#else
    const u64 rz = mul2(rr, p.z[kp]);
    a.S1[kp] = add2(a.S1[kp], rz);
    a.S2[kp] = fma2(rz, p.z[kp], a.S2[kp]);
#endif
  }
  if (ODD) {
    const float r = p.e1 * inv;
    a.t0 += r;
    a.t1 = fmaf(r, p.z1, a.t1);
#if S2_TTRICK
    a.t2 = fmaf(r, p.a1, a.t2);
#else
    a.t2 = fmaf(r * p.z1, p.z1, a.t2);
#endif
  }
  return inv;
}
// exponent-shifted evaluation for a point whose mixture sum underflows fp32 (rare; same as stat_point of k_mixture.cu)
template <int KP, bool ODD>
__device__ __forceinline__ void shifted_point(const Coef<KP, ODD> &q, Acc<KP, ODD> &a, float x, float g) {
  const u64 xx = pk(x, x);
  u64 a2[KP], z2[KP];
  float mx = -INFINITY;
  float z1 = 0.f, a1 = -INFINITY;
  if (ODD) {
    z1 = fmaf(x, q.s1, q.m1);
    a1 = fmaf(-z1, z1, q.c1);
    mx = a1;
  }
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    z2[kp] = fma2(xx, q.s2[kp], q.m2[kp]);
    a2[kp] = fnma2(z2[kp], z2[kp], q.cc2[kp]);
    float lo, hi;
    unpk(a2[kp], lo, hi);
    mx = fmaxf(mx, fmaxf(lo, hi));
  }
  mx = mx == -INFINITY ? 0.f : mx;
  const u64 nm = pk(-mx, -mx);
  u64 sum2 = 0ull, e2[KP];
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    e2[kp] = ex2_2(add2(a2[kp], nm));
    sum2 = add2(sum2, e2[kp]);
  }
  const float e1 = ODD ? ex2_approx(a1 - mx) : 0.f;
  const float sum = hsum(sum2) + e1;
  const float inv = sum > 0.f ? g * rcp_approx(sum) : 0.f;
  const u64 inv2 = pk(inv, inv);
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const u64 rr = mul2(e2[kp], inv2);
    a.S0[kp] = add2(a.S0[kp], rr);
#if S2_TTRICK
    a.S1[kp] = fma2(rr, z2[kp], a.S1[kp]);
    a.S2[kp] = fma2(rr, a2[kp], a.S2[kp]);                  // (the unshifted argument, as in stage2)
#else
    const u64 rz = mul2(rr, z2[kp]);
    a.S1[kp] = add2(a.S1[kp], rz);
    a.S2[kp] = fma2(rz, z2[kp], a.S2[kp]);
#endif
  }
  if (ODD) {
    const float r = e1 * inv;
    a.t0 += r;
    a.t1 = fmaf(r, z1, a.t1);
#if S2_TTRICK
    a.t2 = fmaf(r, a1, a.t2);
#else
    a.t2 = fmaf(r * z1, z1, a.t2);
#endif
  }
}

// shared memory of one warp: posterior tiles (2 stages), observation records (2 stages), histogram rows
struct alignas(16) Rec { float x0, x1; int voff; int pad; };

// One 32-point block.  gt = this lane's column of the posterior tile (stride 32 floats per point), rec = the block's
// observation records, bdb = this lane's histogram column as a byte address (rows are 128 bytes apart and the
// record carries 128 * symbol).  PTS points per loop trip with constant offsets from two bumped pointers.
// Nothing but the accumulators is carried from one trip to the next: the first version of this kernel software-
// pipelined stage 1 of point p + 1 with stage 2 of point p through two loop-carried buffers, and at 128 registers
// ptxas paid ~5 register moves per point for them at the back-edge.  Letting it schedule the PTS independent points of
// a trip itself measures faster at every shape (scale-out 96.0 -> 89.4 ms per 3.2e7 points with 4 points per trip;
// N = 512, K = 3: 9.93 -> 9.38 ms per 2.5e6 with 8; 16 and 32 are slower again, as are 96 or 80 registers with 20 or 24
// warps per SM: 104.6 / 143.8 ms).  tools/pipe_peaks3.cu: the bare K = 8 body -- 27 packed + 9 MUFU + 4 scalar
// instructions per point, no loads -- reaches 0.84 of the MUFU peak with 16 warps per SM; that, not 1.0, is the ceiling
// of this instruction mix.  Packing the arithmetic over two POINTS of one component instead of two components of one
// point (what k_emission_px does; no scalar tail for K = 3) was measured and dropped here: 9.68 vs 9.39 ms at N = 512,
// 1.77 vs 1.69 ms at N = 10 -- two reciprocals and two clamps per pair cost more than the scalar tail they replace.
// Packing only the odd last component over the two points of a trip: 9.42 vs 9.33 ms at N = 512, 1.67 vs 1.71 ms at
// N = 10 -- within the noise either way, not kept.
// GRP (state counts up to 32): a warp step covers PG points x LG lanes, so a block is 32 steps = 32 PG points; gstep /
// rstep are the distances between consecutive steps in the posterior tile (floats) and the record array.
template <int KP, bool ODD, bool HAS_C, bool HAS_D, bool GRP>
__device__ __forceinline__ float block32(const Coef<KP, ODD> &q, Acc<KP, ODD> &acc, const float *gt, const Rec *rec, char *bdb, float &gsl,
                                         bool gsum_on, int gstep_, int rstep_) {
  constexpr int PTS = S2_PTS > 0 ? S2_PTS : (KP >= 4 ? 4 : 8);
  const int gstep = GRP ? gstep_ : 32, rstep = GRP ? rstep_ : 1;
  float hmax = 0.f;
  const float4 *rp = reinterpret_cast<const float4 *>(rec);
#pragma unroll 1
  for (int it = 0; it < 32 / PTS; ++it, rp += PTS * rstep, gt += gstep * PTS) {
#pragma unroll
    for (int h = 0; h < PTS; h += 2) {
      const float g0 = gt[gstep * h], g1 = gt[gstep * h + gstep];
      if (HAS_C) {
        // (the observation enters the packed FMA as a scalar broadcast operand: FFMA2 R, R.F32, R.F32x2, R.F32x2)
        const float x0 = *reinterpret_cast<const float *>(rp + h * rstep);
        const float x1 = *reinterpret_cast<const float *>(rp + (h + 1) * rstep);
        Pend<KP, ODD> p0, p1;
        stage1<KP, ODD>(q, pk(x0, x0), p0);
        stage1<KP, ODD>(q, pk(x1, x1), p1);
        hmax = fmaxf(hmax, fmaxf(stage2<KP, ODD>(acc, p0, g0), stage2<KP, ODD>(acc, p1, g1)));
      }
      if (HAS_D) {
        // categorical histogram: bins[symbol][state] += gamma
        const int v0 = __float_as_int(rp[h * rstep].z), v1 = __float_as_int(rp[(h + 1) * rstep].z);
        *reinterpret_cast<float *>(bdb + v0) += g0;
        *reinterpret_cast<float *>(bdb + v1) += g1;
      }
      if (gsum_on) gsl += g0 + g1;
    }
  }
  return hmax;
}

#ifndef S2_MAXREG
#define S2_MAXREG 128
#endif
constexpr int PGMAX = 4;         // points per warp step in the grouped layout (N >= 8)

// GRP = false: N a multiple of 32; blockIdx.y selects a 32-state slice, a block is 32 points, the posterior tile is
//              the [32 points][32 states] slice of GA.
// GRP = true : N <= 32; lane = (point slot gi, state lj), a block is 32 steps of PG points, the tile is the contiguous
//              [32 PG points][N] piece of GA.  Lanes past the last whole group and padded states carry zero weight.
// PART (!GRP only): N is not a multiple of 32 and the last slice has idle lanes (a template flag: carrying the two
// predicates through the hot loop of the full-slice kernel cost it 3 % at the scale-out shape).
template <int KP, bool ODD, bool GRP, bool PART>
__global__ void __maxnreg__(S2_MAXREG) k_stats2(Stat2Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax, Npad = a.dm.Npad;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int LG = GRP ? a.LG : 32, PG = GRP ? a.PG : 1;
  const int gi = GRP ? lane / LG : 0, lj = GRP ? lane % LG : lane;
  const bool lane_on = GRP ? (gi < PG && lj < N) : (!PART || (int)blockIdx.y * 32 + lane < N);
  const int j = lane_on ? (GRP ? lj : blockIdx.y * 32 + lane) : 0;   // a real state
  const int u = blockIdx.z * a.nw + w;                      // this warp's variable index
  const bool has_c = u < C, has_d = u < D;
  if (!has_c && !has_d) return;
  const int TPB = 32 * PG;                                  // points per block
  const int tile_floats = GRP ? TPB * N : 32 * 32;
  const int nrec = TPB + PG + 1;
  const int bin_floats = (Vmax + 1) * 32;
  const size_t per_warp = (size_t)2 * tile_floats * sizeof(float) + (size_t)2 * nrec * sizeof(Rec) + (size_t)(bin_floats + 4) * sizeof(float);
  unsigned char *base_w = smem_raw + (size_t)w * per_warp;
  float *gtile = reinterpret_cast<float *>(base_w);                                   // [2][tile_floats]
  Rec *recs = reinterpret_cast<Rec *>(base_w + (size_t)2 * tile_floats * sizeof(float));   // [2][nrec]
  float *bins = reinterpret_cast<float *>(base_w + (size_t)2 * tile_floats * sizeof(float) + (size_t)2 * nrec * sizeof(Rec));
  float *bd = bins + lane;
  char *bdb = reinterpret_cast<char *>(bd);
  for (int t = 0; t <= Vmax; ++t) bd[t * 32] = 0.f;
  if (lane <= PG) {                                         // the records past a block's end: x = 0, unseen-symbol row
    recs[TPB + lane] = Rec{0.f, 0.f, Vmax * 128, 0};
    recs[nrec + TPB + lane] = Rec{0.f, 0.f, Vmax * 128, 0};
  }
  // (grouped layout: a zero word at the end of the warp's shared memory for lanes without a state)
  float *zero_word = bins + bin_floats;
  if (GRP && lane == 0) *zero_word = 0.f;

  Coef<KP, ODD> q;
  Acc<KP, ODD> acc;
#if S2_TTRICK
  // Responsibilities do not change when every component of a mixture is scaled by the same factor, so the log2
  // constants are shifted per (variable, state) to put the largest at CREF = E[z^2] of a well-fitted component:
  // the arguments c' - z^2 that the second moment is accumulated through (stage2) then average near zero for the
  // components that carry the mass, and c' S0 - sum r arg loses at most a bit to cancellation.  Components
  // without mass (padding, zero weights) get a finite constant far below the fp32 exponent range: 0 * (-inf) would be NaN.
  constexpr float CREF = 0.72134752f, CNONE = -20000.f;
  float kappa = 0.f;
  {
    float cmax = -INFINITY;
    for (int k = 0; k < Kmax; ++k)
      if (has_c && lane_on && k < a.tb.K[u]) cmax = fmaxf(cmax, a.tb.gc[((int64_t)u * Kmax + k) * Npad + j]);
    kappa = cmax == -INFINITY ? 0.f : CREF - cmax;
  }
#endif
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    float v[6];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * kp + h;
      const bool on = has_c && lane_on && k < a.tb.K[has_c ? u : 0];
      const int64_t o = ((int64_t)(has_c ? u : 0) * Kmax + (k < Kmax ? k : 0)) * Npad + j;
      v[h] = on ? a.tb.gs[o] : 0.f;
      v[2 + h] = on ? a.tb.gm[o] : 0.f;
#if S2_TTRICK
      v[4 + h] = on ? fmaxf(a.tb.gc[o] + kappa, CNONE) : CNONE;
#else
      v[4 + h] = on ? a.tb.gc[o] : -INFINITY;
#endif
    }
    q.s2[kp] = pk(v[0], v[1]); q.m2[kp] = pk(v[2], v[3]); q.cc2[kp] = pk(v[4], v[5]);
    acc.S0[kp] = acc.S1[kp] = acc.S2[kp] = 0ull;
  }
  q.s1 = q.m1 = 0.f;
  q.c1 = -20000.f;
  acc.t0 = acc.t1 = acc.t2 = 0.f;
  if (ODD && has_c && lane_on) {                            // the odd last component (every variable has 2 KP + 1)
    const int64_t o = ((int64_t)u * Kmax + 2 * KP) * Npad + j;
    q.s1 = a.tb.gs[o];
    q.m1 = a.tb.gm[o];
#if S2_TTRICK
    q.c1 = fmaxf(a.tb.gc[o] + kappa, -20000.f);
#else
    q.c1 = a.tb.gc[o];
#endif
  }
  // contiguous run of whole blocks per CTA column
  const int per = (a.nblk + (int)gridDim.x - 1) / (int)gridDim.x;
  const int b0 = (int)blockIdx.x * per;
  const int b1 = b0 + per < a.nblk ? b0 + per : a.nblk;
  if (b0 >= b1) return;
  const float *xcol = a.dv.cont + (int64_t)(has_c ? u : 0) * a.dv.P + a.p0;
  const uint8_t *dcol = a.dv.disc + (int64_t)(has_d ? u : 0) * a.dv.P + a.p0;
  const bool gsum_on = u == 0;
  double gs = 0.0;
  int inwin = 0;

  // posterior tile of a block.  !GRP: lane l copies, for i = 0..7, the 16 bytes of point 4 i + l / 8, states
  // 4 (l % 8) .. + 3 of this slice.  GRP: the tile is one contiguous piece of GA.
  // A slot of four states past N (last slice of a state count that is not a multiple of 32; N is a multiple of 4,
  // so a slot is all states or all padding) is never copied: it keeps the zeros written here, and its lanes
  // accumulate nothing (weight 0 against coefficients without mass).
  const bool slot_on = GRP || !PART || (int)blockIdx.y * 32 + (lane & 7) * 4 < N;
  const float *gsrc0 = a.GA + (int64_t)(lane >> 3) * N + blockIdx.y * 32 + (lane & 7) * 4;
  float *gdst0 = gtile + (lane >> 3) * 32 + (lane & 7) * 4;
  if (!GRP && PART && !slot_on) {
#pragma unroll
    for (int i = 0; i < 16; ++i) *reinterpret_cast<float4 *>(gdst0 + (i >> 3) * 1024 + (i & 7) * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto issue_tile = [&](int blk, int stage) {
    if (GRP) {
      const float *src = a.GA + (int64_t)blk * tile_floats;
      float *dst = gtile + stage * tile_floats;
      for (int c = lane * 4; c < tile_floats; c += 128) cp_async16(dst + c, src + c);
    } else {
      const float *src = gsrc0 + (int64_t)blk * 32 * N;
      float *dst = gdst0 + stage * 1024;
      if (slot_on) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cp_async16(dst + i * 128, src + (int64_t)i * 4 * N);
      }
    }
    cp_async_commit();
  };
  issue_tile(b0, 0);
  float xr[GRP ? PGMAX : 1];
  int vr[GRP ? PGMAX : 1];
  auto load_obs = [&](int blk) {
    if (GRP) {
#pragma unroll
      for (int i = 0; i < PGMAX; ++i) {
        if (i < PG) {
          const int64_t pnt = (int64_t)blk * TPB + i * 32 + lane;
          xr[i] = has_c ? __ldg(xcol + pnt) : 0.f;
          vr[i] = has_d ? (int)__ldg(dcol + pnt) : Vmax;
        }
      }
    } else {
      if (has_c) xr[0] = __ldg(xcol + (int64_t)blk * 32 + lane);
      if (has_d) vr[0] = (int)__ldg(dcol + (int64_t)blk * 32 + lane);
    }
  };
  xr[0] = 0.f;
  vr[0] = Vmax;
  load_obs(b0);

  for (int blk = b0; blk < b1; ++blk) {
    const int stage = (blk - b0) & 1;
    Rec *rec = recs + stage * nrec;
    if (GRP) {
#pragma unroll
      for (int i = 0; i < PGMAX; ++i)
        if (i < PG) rec[i * 32 + lane] = Rec{xr[i], xr[i], (vr[i] < Vmax ? vr[i] : Vmax) * 128, 0};
    } else {
      rec[lane] = Rec{xr[0], xr[0], (vr[0] < Vmax ? vr[0] : Vmax) * 128, 0};
    }
    cp_async_wait_all();
    __syncwarp();
    if (blk + 1 < b1) {
      issue_tile(blk + 1, stage ^ 1);
      load_obs(blk + 1);
    }
    // this lane's column of the tile and its first record; lanes without a state read a zero word with stride 0
    const float *gt = GRP ? (lane_on ? gtile + stage * tile_floats + gi * N + lj : zero_word) : gtile + stage * 1024 + lane;
    const int gstep = GRP ? (lane_on ? PG * N : 0) : 32;
    const Rec *rec0 = rec + (GRP ? (gi < PG ? gi : 0) : 0);
    float gsl = 0.f, hmax;
    if (has_c && has_d) hmax = block32<KP, ODD, true, true, GRP>(q, acc, gt, rec0, bdb, gsl, gsum_on, gstep, PG);
    else if (has_c) hmax = block32<KP, ODD, true, false, GRP>(q, acc, gt, rec0, bdb, gsl, gsum_on, gstep, PG);
    else hmax = block32<KP, ODD, false, true, GRP>(q, acc, gt, rec0, bdb, gsl, gsum_on, gstep, PG);
    gs += (double)gsl;
    if (has_c && __any_sync(0xffffffffu, hmax > HBIG)) {
      // rare: a state with posterior mass at a point where fp32 has flushed (part of) its mixture.  Each lane takes
      // back what its fast path added for such points (same instructions, weight -g) and adds the exponent-shifted
      // evaluation instead.
#pragma unroll 1
      for (int p = 0; p < 32; ++p) {
        const float x = rec0[p * PG].x0, g = gt[p * gstep];
        Pend<KP, ODD> pr;
        stage1<KP, ODD>(q, pk(x, x), pr);
        const float inv = g * rcp_approx(fmaxf(mix_sum<KP, ODD>(pr), 1.17549435e-38f));
        const bool redo = inv > HBIG;
        stage2<KP, ODD>(acc, pr, redo ? -g : 0.f);
        shifted_point<KP, ODD>(q, acc, x, redo ? g : 0.f);
      }
    }
    inwin += TPB;
    if (inwin >= a.window || blk + 1 >= b1) {
      inwin = 0;
      if (has_c && lane_on) {
        const int K = a.tb.K[u];
        // component k's window sums: z = (x - mu') s with mu' = -m/s the fp32-rounded centre; re-centre on the fp64
        // old mean; sum r z^2 = c' S0 - sum r arg (stage2)
        auto flush_one = [&](int k, float sv, float mv, float cv, double a0, double a1, double a2v) {
          if (k >= K || a0 == 0.0) return;
          const int64_t dst = ((int64_t)u * N + j) * Kmax + k;
          const double s = (double)sv;
          double m1 = 0.0, m2d = 0.0;
          if (s > 0.0) {
            const double delta = -(double)mv / s - a.mu_old[dst];
            const double zz = (double)cv * a0 - a2v;
            const double c1 = a1 / s, c2 = zz / (s * s);
            m1 = c1 + delta * a0;
            m2d = c2 + 2.0 * delta * c1 + delta * delta * a0;
          }
          atomicAdd(a.gmm + 3 * dst + 0, a0);
          atomicAdd(a.gmm + 3 * dst + 1, m1);
          atomicAdd(a.gmm + 3 * dst + 2, m2d);
        };
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) {
          float sv[2], mv[2], cv[2], a0[2], a1[2], a2v[2];
          unpk(q.s2[kp], sv[0], sv[1]); unpk(q.m2[kp], mv[0], mv[1]); unpk(q.cc2[kp], cv[0], cv[1]);
          unpk(acc.S0[kp], a0[0], a0[1]); unpk(acc.S1[kp], a1[0], a1[1]); unpk(acc.S2[kp], a2v[0], a2v[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) flush_one(2 * kp + h, sv[h], mv[h], cv[h], (double)a0[h], (double)a1[h], (double)a2v[h]);
        }
        if (ODD) flush_one(2 * KP, q.s1, q.m1, q.c1, (double)acc.t0, (double)acc.t1, (double)acc.t2);
      }
      acc.t0 = acc.t1 = acc.t2 = 0.f;
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) acc.S0[kp] = acc.S1[kp] = acc.S2[kp] = 0ull;
      if (gsum_on) {
        if (lane_on && gs != 0.0) atomicAdd(a.gsum + j, gs);
        gs = 0.0;
      }
      if (has_d) {
        for (int v = 0; v < Vmax; ++v) {
          const float b = bd[v * 32];
          if (lane_on && b != 0.f) atomicAdd(a.cat + ((int64_t)u * N + j) * Vmax + v, (double)b);
          bd[v * 32] = 0.f;
        }
        bd[Vmax * 32] = 0.f;                                // unseen-symbol row is discarded
      }
    }
  }
}

// lane groups of the grouped layout: the state count itself when that packs more points into a warp step
void group_shape(const Dims &d, int &LG, int &PG) {
  if (d.N > 16) { LG = 32; PG = 1; return; }
  const int pg_pad = 32 / d.Npad, pg_exact = 32 / d.N;
  LG = pg_exact > pg_pad ? d.N : d.Npad;
  PG = 32 / LG;
}

size_t stats2_smem(const Dims &d, int nw, bool grp) {
  int LG = 32, PG = 1;
  if (grp) group_shape(d, LG, PG);
  const size_t tile = grp ? (size_t)32 * PG * d.N : 1024;
  const size_t per_warp = 2 * tile * sizeof(float) + (size_t)2 * (32 * PG + PG + 1) * sizeof(Rec) + (size_t)((d.Vmax + 1) * 32 + 4) * sizeof(float);
  return (size_t)nw * per_warp;
}

}  // namespace

bool stats2_supported(const dnbc_ctx *c) {
  const Dims &d = c->dm;
  const int units = d.C > d.D ? d.C : d.D;
  if (d.Kmax > 8 || units < 1) return false;
  if (d.N >= 32 && d.N % 4 == 0) return stats2_smem(d, NW_MAX, false) <= 110 * 1024;   // (rows of 16-byte aligned posterior slots)
  return d.N >= 8 && d.N < 32 && stats2_smem(d, NW_MAX, true) <= 110 * 1024;
}

// whole blocks of [p0, p0 + np) (32 points, or 32 warp steps of the grouped layout); returns the number of points
// covered (the caller gives the rest to k_stats_pk)
int64_t launch_stats2(dnbc_ctx *c, int64_t p0, int64_t np) {
  const Dims &d = c->dm;
  const bool grp = d.N < 32, part = !grp && d.N % 32 != 0;
  int LG = 32, PG = 1;
  if (grp) group_shape(d, LG, PG);
  const int TPB = 32 * PG;
  const int64_t nblk = np / TPB;
  if (nblk <= 0) return 0;
  StatsLayout sl = stats_layout(d);
  const int units = d.C > d.D ? d.C : d.D;
  // 16 warps per SM at <= 128 registers: CTAs of nw warps, e.g. 3 x 5 warps when there are 5 variables
  const int nw = units > NW_MAX ? NW_MAX : units;
  const int gz = (units + nw - 1) / nw, gy = grp ? 1 : (d.N + 31) / 32;
  Stat2Args a{d, c->tables(), c->data(), p0, (int)nblk, c->GA, c->d_mu, c->d_stats + sl.gmm, c->d_stats + sl.cat,
              c->d_stats + sl.gamma, 4096, nw, LG, PG};
  const size_t smem = stats2_smem(d, nw, grp);
  const int per_sm = S2_WPSM / nw > 8 ? 8 : S2_WPSM / nw;
  int64_t gx = (int64_t)c->sm_count * per_sm / ((int64_t)gy * gz);
  const int64_t maxgx = (nblk + 7) / 8;
  if (gx > maxgx) gx = maxgx;
  if (gx < 1) gx = 1;
  const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  // every variable with the same odd K >= 3: the last component in scalar registers instead of a padded pair
  bool odd = d.C > 0 && (d.Kmax & 1) && d.Kmax >= 3;
  for (int v = 0; v < d.C; ++v) odd = odd && c->K[v] == d.Kmax;
  const int KP = odd ? d.Kmax / 2 : (d.Kmax + 1) / 2;
#define STATS2_LAUNCH(KP_, ODD_, GRP_)                                                                            \
  do {                                                                                                            \
    auto kf = (!GRP_ && part) ? k_stats2<KP_, ODD_, GRP_, !GRP_> : k_stats2<KP_, ODD_, GRP_, false>;              \
    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                             \
    kf<<<grid, nw * 32, smem, c->stream>>>(a);                                                                    \
  } while (0)
#define STATS2_CASE(KP_)                                                                                          \
  if (KP == KP_) {                                                                                                \
    if (odd && KP_ < 4) {                                                                                         \
      if (grp) STATS2_LAUNCH(KP_, true, true); else STATS2_LAUNCH(KP_, true, false);                              \
    } else {                                                                                                      \
      if (grp) STATS2_LAUNCH(KP_, false, true); else STATS2_LAUNCH(KP_, false, false);                            \
    }                                                                                                             \
  }
  STATS2_CASE(1) STATS2_CASE(2) STATS2_CASE(3) STATS2_CASE(4)
#undef STATS2_CASE
#undef STATS2_LAUNCH
  return nblk * TPB;
}
