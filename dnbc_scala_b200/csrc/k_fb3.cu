// k_fb3.cu -- K2 for 33..64 hidden states: scaled forward / backward recursions with the
// alpha_t * A product on the tensor cores, 16 sequences per warp.
//
// The reference has NO forward-backward; these kernels implement SURVEY Appendix B,
// produce exactly the arrays k_fb2.cu produces (alpha^, gamma, u, c_t, log-likelihood, pi-stat) and
// are checked against the fp64 oracle (oracle/dnbc_oracle.c fb_one).
//
// Why not tcgen05 here: at N = 64 a step is a [16 x 64] x [64 x 64] product per warp whose RESULT is
// the next step's left operand.  With mma.sync the accumulator fragment of two adjacent 8-column
// tiles IS the A fragment of the next step's 16-wide K tile (same lane, same registers), so the
// recursion never leaves the register file; a TMEM round trip per step (tcgen05.mma -> commit ->
// tcgen05.ld -> shared-memory operand) would put ~1 us of latency on a 1000-step dependency chain.
// (N >= 256 does use tcgen05, k_tc.cu: there the per-step product is large enough to amortise it.)
//
//   * operands are bf16 split pairs, x = hi + lo, product = hi*hi + hi*lo + lo*hi accumulated in
//     fp32 (relative error ~2^-17 per element; one bf16 pass would be 2^-9);
//   * the transition matrix sits in shared memory as two 64 x 64 bf16 tiles (hi, lo), rows =
//     output state, 128-byte rows with the 16-byte chunks XOR-swizzled so that ldmatrix.x4 is
//     conflict-free; forward uses [to][from], backward [from][to];
//   * emission / alpha rows are read and written straight in the accumulator layout: a lane owns
//     columns 8 nt + 2 q, + 1 of rows g and g + 8, i.e. one 8-byte access per (row, tile), whole
//     32-byte sectors per quad; a step's rows are requested at its top and first used after its
//     96 MMAs (prefetching a step ahead costs 32 registers, i.e. a CTA per SM: measured 17 % slower);
//   * the row normaliser is a 2-step quad shuffle.
// The step is bound by HBM (forward: read E, write alpha^; backward: read E, alpha^, write gamma,
// u = 4 N floats per point) instead of by the FP32 pipe as in k_fb2.cu.
#include <cuda_bf16.h>

#include "dnbc_internal.cuh"

namespace {

constexpr int NP = 64;            // padded states
constexpr int NT = NP / 8;        // 8-column accumulator tiles
constexpr int KT = NP / 16;       // 16-wide K tiles
constexpr int FB3_WARPS = 4;

struct Fb3Args {
  int N, Npad;
  const float *A;                 // [N][Npad] fp32 table (row = from-state)
  const float *pi;                // [N]
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;
  float *AL, *GA, *U, *cs;
  double *stats;
  int64_t st_pi;
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// (lo, hi) fp32 pair -> packed bf16 hi word and bf16 lo word of the split
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *(const uint32_t *)&h;
  lo = *(const uint32_t *)&l;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// transition matrix -> two swizzled bf16 tiles.  Row n = output state of the product, K index k =
// summed state: forward B[n = to][k = from] = a[from][to]; backward B[n = from][k = to] = a[from][to].
__device__ void fill_tiles(const Fb3Args &a, bool backward, unsigned char *bh, unsigned char *bl) {
  for (int idx = threadIdx.x; idx < NP * NP; idx += blockDim.x) {
    const int n = idx / NP, k = idx % NP;
    const int from = backward ? n : k, to = backward ? k : n;
    const float v = (from < a.N && to < a.N) ? a.A[from * a.Npad + to] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const int o = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *(__nv_bfloat16 *)(bh + o) = h;
    *(__nv_bfloat16 *)(bl + o) = l;
  }
}

// acc[nt] = sum_kt (xh + xl)[kt] * (Bh + Bl)[kt][nt]   (lo * lo dropped)
__device__ __forceinline__ void step_mma(float (&acc)[NT][4], const uint32_t (&xh)[KT][4], const uint32_t (&xl)[KT][4],
                                         uint32_t sbh, uint32_t sbl, const uint32_t (&loff)[KT / 2]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int kp = 0; kp < KT / 2; ++kp) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bh[4], bl[4];
      ldsm4(sbh + nt * 1024 + loff[kp], bh);
      ldsm4(sbl + nt * 1024 + loff[kp], bl);
      mma_bf16(acc[nt], xh[2 * kp], bh[0], bh[1]);
      mma_bf16(acc[nt], xh[2 * kp + 1], bh[2], bh[3]);
      mma_bf16(acc[nt], xl[2 * kp], bh[0], bh[1]);
      mma_bf16(acc[nt], xl[2 * kp + 1], bh[2], bh[3]);
      mma_bf16(acc[nt], xh[2 * kp], bl[0], bl[1]);
      mma_bf16(acc[nt], xh[2 * kp + 1], bl[2], bl[3]);
    }
  }
}

// one row (N floats) in the accumulator layout: v[nt][0..1] = row[8 nt + 2 q + {0, 1}]
__device__ __forceinline__ void load_row(const float *row, int N, int q, bool on, float fill, float (&v)[NT][2]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = 8 * nt + 2 * q;
    if (on && col < N) {                                     // N is even: col + 1 < N as well
      const float2 x = __ldg((const float2 *)(row + col));
      v[nt][0] = x.x;
      v[nt][1] = x.y;
    } else {
      v[nt][0] = v[nt][1] = fill;
    }
  }
}
__device__ __forceinline__ void store_row(float *row, int N, int q, const float (&v)[NT][2]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = 8 * nt + 2 * q;
    if (col < N) *(float2 *)(row + col) = make_float2(v[nt][0], v[nt][1]);
  }
}

struct RowInfo {
  int T[2];
  int64_t pt[2];
};
__device__ __forceinline__ RowInfo rows_of(const Fb3Args &a, int64_t sb, int g, int &Tmax) {
  RowInfo r;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t s = sb + g + 8 * h;
    r.T[h] = s < a.s1 ? (int)(a.off[s + 1] - a.off[s]) : 0;
    r.pt[h] = s < a.s1 ? a.off[s] - a.p0 : 0;
  }
  int t = r.T[0] > r.T[1] ? r.T[0] : r.T[1];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int x = __shfl_xor_sync(0xffffffffu, t, o);
    t = x > t ? x : t;
  }
  Tmax = t;
  return r;
}

// ---- forward: alpha^_t(j) = b~_t(j) sum_i alpha^_{t-1}(i) a_ij / c_t ;  LL += ln c_t + m_t
#ifndef DNBC_FB3_FWD_PREFETCH
#define DNBC_FB3_FWD_PREFETCH 0
#endif
#ifndef DNBC_FB3_FWD_OCC
#define DNBC_FB3_FWD_OCC 3
#endif
__global__ void __launch_bounds__(FB3_WARPS * 32, DNBC_FB3_FWD_OCC) k_forward3(Fb3Args a) {
  __shared__ __align__(1024) unsigned char sB[2][NP * 128];
  const int N = a.N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  fill_tiles(a, false, sB[0], sB[1]);
  __syncthreads();
  const uint32_t sbh = (uint32_t)__cvta_generic_to_shared(sB[0]), sbl = (uint32_t)__cvta_generic_to_shared(sB[1]);
  uint32_t loff[KT / 2];
#pragma unroll
  for (int kp = 0; kp < KT / 2; ++kp) loff[kp] = (lane & 7) * 128 + (((4 * kp + (lane >> 3)) ^ (lane & 7)) << 4);
  float pi[NT][2];
  load_row(a.pi, N, q, true, 0.f, pi);

  const int64_t nwarps = (int64_t)gridDim.x * FB3_WARPS;
  double ll = 0.0;
  for (int64_t sb = a.s0 + ((int64_t)blockIdx.x * FB3_WARPS + w) * 16; sb < a.s1; sb += nwarps * 16) {
    int Tmax;
    const RowInfo ri = rows_of(a, sb, g, Tmax);
    const float *Er[2] = {a.E + ri.pt[0] * N, a.E + ri.pt[1] * N};
    float *Ar[2] = {a.AL + ri.pt[0] * N, a.AL + ri.pt[1] * N};
#if DNBC_FB3_FWD_PREFETCH
    float en[2][NT][2];
#pragma unroll
    for (int h = 0; h < 2; ++h) load_row(Er[h], N, q, ri.T[h] > 0, -INFINITY, en[h]);
#endif
    uint32_t xh[KT][4], xl[KT][4];
    for (int t = 0; t < Tmax; ++t) {
      float e[2][NT][2], acc[NT][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#if DNBC_FB3_FWD_PREFETCH
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { e[h][nt][0] = en[h][nt][0]; e[h][nt][1] = en[h][nt][1]; }
        load_row(Er[h] + (int64_t)(t + 1) * N, N, q, t + 1 < ri.T[h], -INFINITY, en[h]);
#else
        // this step's emission row is requested here and first used after the step's 96 MMAs
        load_row(Er[h] + (int64_t)t * N, N, q, t < ri.T[h], -INFINITY, e[h]);
#endif
      }
      if (t == 0) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][2] = pi[nt][0]; acc[nt][1] = acc[nt][3] = pi[nt][1]; }
      } else {
        step_mma(acc, xh, xl, sbh, sbl, loff);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float m = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) m = fmaxf(m, fmaxf(e[h][nt][0], e[h][nt][1]));
        m = quad_max(m);
        if (m == -INFINITY) m = 0.f;
        const float ms = -m * (float)DNBC_LOG2E;
        float c = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            float &v = acc[nt][2 * h + x];
            v *= ex2_approx(fmaf(e[h][nt][x], (float)DNBC_LOG2E, ms));
            c += v;
          }
        }
        c = quad_sum(c);
        const float inv = rcp_approx(c);
        const bool live = t < ri.T[h];
        float al[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          al[nt][0] = acc[nt][2 * h] * inv;
          al[nt][1] = acc[nt][2 * h + 1] * inv;
        }
        if (live) {
          store_row(Ar[h] + (int64_t)t * N, N, q, al);
          if (q == 0) {
            a.cs[ri.pt[h] + t] = c;
            ll += (double)(lg2_approx(c) * (float)DNBC_LN2) + (double)m;
          }
        }
        // the accumulator tiles 2 kt, 2 kt + 1 of this row are registers h, 2 + h of K tile kt
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          split2(al[2 * kt][0], al[2 * kt][1], xh[kt][h], xl[kt][h]);
          split2(al[2 * kt + 1][0], al[2 * kt + 1][1], xh[kt][2 + h], xl[kt][2 + h]);
        }
      }
    }
  }
  if (ll != 0.0) atomicAdd(a.stats, ll);
}

// ---- backward: beta^_t(i) = sum_j a_ij u_t(j),  u_t(j) = b~_{t+1}(j) beta^_{t+1}(j) / c_{t+1};
//      gamma_t = alpha^_t beta^_t (beta^ rescaled so that the row sums to 1);  pi-stat += gamma_0
#ifndef DNBC_FB3_BWD_OCC
#define DNBC_FB3_BWD_OCC 2
#endif
__global__ void __launch_bounds__(FB3_WARPS * 32, DNBC_FB3_BWD_OCC) k_backward3(Fb3Args a) {
  __shared__ __align__(1024) unsigned char sB[2][NP * 128];
  const int N = a.N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  fill_tiles(a, true, sB[0], sB[1]);
  __syncthreads();
  const uint32_t sbh = (uint32_t)__cvta_generic_to_shared(sB[0]), sbl = (uint32_t)__cvta_generic_to_shared(sB[1]);
  uint32_t loff[KT / 2];
#pragma unroll
  for (int kp = 0; kp < KT / 2; ++kp) loff[kp] = (lane & 7) * 128 + (((4 * kp + (lane >> 3)) ^ (lane & 7)) << 4);
  float piacc[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) piacc[nt][0] = piacc[nt][1] = 0.f;

  const int64_t nwarps = (int64_t)gridDim.x * FB3_WARPS;
  for (int64_t sb = a.s0 + ((int64_t)blockIdx.x * FB3_WARPS + w) * 16; sb < a.s1; sb += nwarps * 16) {
    int Tmax;
    const RowInfo ri = rows_of(a, sb, g, Tmax);
    const int64_t r0[2] = {ri.pt[0] * N, ri.pt[1] * N};
    // rows are aligned at their LAST step: iteration k handles t = T - 1 - k
    float en[2][NT][2], csn[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) en[h][nt][0] = en[h][nt][1] = -INFINITY;
      csn[h] = 1.f;
    }
    float acc[NT][4];                                        // beta^ of the step just finished
    for (int k = 0; k < Tmax; ++k) {
      float al[2][NT][2], e[2][NT][2], cn[2];
      int t[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        t[h] = ri.T[h] - 1 - k;
        cn[h] = csn[h];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { e[h][nt][0] = en[h][nt][0]; e[h][nt][1] = en[h][nt][1]; }
        // alpha^_t is requested here and first used after this step's MMAs; E_t / c_t are step
        // t - 1's "t + 1" quantities, needed BEFORE its MMAs, so they are fetched one step ahead
        load_row(a.AL + r0[h] + (int64_t)t[h] * N, N, q, t[h] >= 0, 0.f, al[h]);
        load_row(a.E + r0[h] + (int64_t)t[h] * N, N, q, t[h] > 0, -INFINITY, en[h]);
        csn[h] = t[h] > 0 ? __ldg(a.cs + ri.pt[h] + t[h]) : 1.f;
      }
      if (k > 0) {
        uint32_t xh[KT][4], xl[KT][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float m = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) m = fmaxf(m, fmaxf(e[h][nt][0], e[h][nt][1]));
          m = quad_max(m);
          if (m == -INFINITY) m = 0.f;
          const float ms = -m * (float)DNBC_LOG2E;
          const float ic = rcp_approx(cn[h]);
          float u[NT][2];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int x = 0; x < 2; ++x)
              u[nt][x] = ex2_approx(fmaf(e[h][nt][x], (float)DNBC_LOG2E, ms)) * acc[nt][2 * h + x] * ic;
          if (t[h] >= 0) store_row(a.U + r0[h] + (int64_t)t[h] * N, N, q, u);
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) {
            split2(u[2 * kt][0], u[2 * kt][1], xh[kt][h], xl[kt][h]);
            split2(u[2 * kt + 1][0], u[2 * kt + 1][1], xh[kt][2 + h], xl[kt][2 + h]);
          }
        }
        step_mma(acc, xh, xl, sbh, sbl, loff);
      } else {
        float z[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          z[nt][0] = z[nt][1] = 0.f;
          acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 1.f;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (t[h] >= 0) store_row(a.U + r0[h] + (int64_t)t[h] * N, N, q, z);   // no transition out of the last step
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float s = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) s = fmaf(al[h][nt][0], acc[nt][2 * h], fmaf(al[h][nt][1], acc[nt][2 * h + 1], s));
        s = quad_sum(s);
        const float fix = s > 0.f ? rcp_approx(s) : 1.f;
        float ga[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          acc[nt][2 * h] *= fix;
          acc[nt][2 * h + 1] *= fix;
          ga[nt][0] = al[h][nt][0] * acc[nt][2 * h];
          ga[nt][1] = al[h][nt][1] * acc[nt][2 * h + 1];
        }
        if (t[h] >= 0) store_row(a.GA + r0[h] + (int64_t)t[h] * N, N, q, ga);
        if (t[h] == 0) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) { piacc[nt][0] += ga[nt][0]; piacc[nt][1] += ga[nt][1]; }
        }
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int j = 8 * nt + 2 * q + x;
      if (j < N && piacc[nt][x] != 0.f) atomicAdd(a.stats + a.st_pi + j, (double)piacc[nt][x]);
    }
}

}  // namespace

bool fb3_supported(const dnbc_ctx *c) {
  return c->dm.Npad == NP && c->dm.N % 2 == 0;
}

void launch_fb3(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward) {
  StatsLayout sl = stats_layout(c->dm);
  Fb3Args a{c->dm.N, c->dm.Npad, c->t_A, c->t_pi, c->d_off, s0, s1, p0, c->E, c->AL, c->GA, c->U, c->cs, c->d_stats, sl.pi};
  const int64_t groups = (s1 - s0 + 15) / 16;
  int64_t blocks = (groups + FB3_WARPS - 1) / FB3_WARPS;
  const int64_t cap = (int64_t)c->sm_count * (backward ? DNBC_FB3_BWD_OCC : DNBC_FB3_FWD_OCC);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (backward) k_backward3<<<(int)blocks, FB3_WARPS * 32, 0, c->stream>>>(a);
  else k_forward3<<<(int)blocks, FB3_WARPS * 32, 0, c->stream>>>(a);
}
