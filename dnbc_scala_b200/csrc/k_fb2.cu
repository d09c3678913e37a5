// k_fb2.cu -- K2: scaled forward / backward recursions with the transition matrix held in
// REGISTERS (state spaces up to 64).
//
// The reference has NO forward-backward: its learner is supervised counting
// (DynamicNaiveBayesianClassifier.scala:124-161).  These kernels implement SURVEY Appendix B
// and are checked against the fp64 oracle (oracle/dnbc_oracle.c fb_one) and, through the
// clamped-label bridge test, against the reference's supervised counts.
//
// Mapping: a group of LG = min(32, Npad) lanes owns SQ = 4 sequences at a time (four
// independent dependency chains per lane); lane l holds states l, l + LG (R per lane).
//   forward : lane keeps COLUMN j of A (alpha_t(j) = b_t(j) sum_i alpha_{t-1}(i) a_ij)
//   backward: lane keeps ROW i of A    (beta_t(i) = sum_j a_ij u_t(j))
// so the N^2 multiply-adds of a step read the matrix from registers; the previous vector is
// exchanged through a warp-private shared-memory line and read back as broadcast LDS.128
// (N/4 loads per step instead of the N shuffles + 2N matrix loads of the first version, which
// ncu showed to be 90 % overhead instructions).  The per-step normaliser is a shuffle
// reduction inside the group.  Groups are persistent and walk sequences with a grid stride.
#include "dnbc_internal.cuh"
#include "packed_f32.cuh"

template <int LG>
__device__ __forceinline__ unsigned gmask() {
  if (LG == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << (LG & 31)) - 1u) << ((lane / LG) * LG);
}
// A lane group is LG consecutive lanes.  Normally LG is a power of two (an xor butterfly reduces in log2 LG rounds);
// the state counts 3, 5, 6, 9, 10 get groups of exactly LG lanes instead, because that puts more sequences into a
// warp (N = 10: 3 x 10 lanes instead of 2 x 16) -- their reductions run a tree to the group's first lane plus a
// broadcast (one round more).  Lanes past the last whole group idle.
template <int LG>
struct GroupRed {
  static constexpr bool POW2 = (LG & (LG - 1)) == 0;
  static constexpr int P = LG <= 1 ? 1 : LG <= 2 ? 2 : LG <= 4 ? 4 : LG <= 8 ? 8 : LG <= 16 ? 16 : 32;   // next power of two
  // SQ sums (+) and SQ maxima reduced together over the group; every lane ends up with the results
  template <int SQ>
  static __device__ __forceinline__ void sum_max(float (&sm)[SQ], float (&mx)[SQ], unsigned mask, int lj, int base) {
    if constexpr (POW2) {
#pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          sm[q] += __shfl_xor_sync(mask, sm[q], o, LG);
          mx[q] = fmaxf(mx[q], __shfl_xor_sync(mask, mx[q], o, LG));
        }
    } else {
#pragma unroll
      for (int o = P / 2; o > 0; o >>= 1) {
        const int src = base + (lj + o < LG ? lj + o : lj);
        const bool take = lj < o && lj + o < LG;
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const float ts = __shfl_sync(mask, sm[q], src), tm = __shfl_sync(mask, mx[q], src);
          if (take) {
            sm[q] += ts;
            mx[q] = fmaxf(mx[q], tm);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < SQ; ++q) {
        sm[q] = __shfl_sync(mask, sm[q], base);
        mx[q] = __shfl_sync(mask, mx[q], base);
      }
    }
  }
  template <int SQ>
  static __device__ __forceinline__ void max_only(float (&mx)[SQ], unsigned mask, int lj, int base) {
    if constexpr (POW2) {
#pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < SQ; ++q) mx[q] = fmaxf(mx[q], __shfl_xor_sync(mask, mx[q], o, LG));
    } else {
#pragma unroll
      for (int o = P / 2; o > 0; o >>= 1) {
        const int src = base + (lj + o < LG ? lj + o : lj);
        const bool take = lj < o && lj + o < LG;
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const float tm = __shfl_sync(mask, mx[q], src);
          if (take) mx[q] = fmaxf(mx[q], tm);
        }
      }
#pragma unroll
      for (int q = 0; q < SQ; ++q) mx[q] = __shfl_sync(mask, mx[q], base);
    }
  }
};
// length of the exchanged vector: the states of a group, padded to whole float4's (pad entries stay zero)
__host__ __device__ constexpr int fb2_np(int LG, int R) { return LG * R <= 2 ? LG * R : (LG * R + 3) / 4 * 4; }

struct Fb2Args {
  Dims dm;
  DeviceTables tb;
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;   // [np][N] natural-log emissions (chunk-relative rows)
  float *AL;        // [np][N] scaled alpha
  float *GA;        // [np][N] gamma
  float *U;         // [np][N] u_t(j) = b~_{t+1}(j) beta_{t+1}(j) / c_{t+1}  (0 on the last step)
  float *cs;        // [np]    c_t
  double *stats;    // packed statistics ([0] = log-likelihood)
  int64_t st_pi;    // offset of the pi-stat
  int64_t st_xi;    // offset of the xi-stat (transition statistics fused into the backward sweep, XI variants)
};

constexpr int FB_WARPS = 4;

__device__ __forceinline__ float hsum2(u64 a, u64 b) {
  float lo, hi;
  unpk(add2(a, b), lo, hi);
  return lo + hi;
}

// dot products of the shared vector with this lane's R matrix lines.  The matrix lines are
// held as packed pairs over the reduction index, so one FFMA2 does two multiply-adds and one
// broadcast LDS.128 feeds two of them (two independent partial sums per line).
template <int NP, int R>
__device__ __forceinline__ void matvec(const float *vec, const u64 (&M)[R][(NP + 1) / 2], float (&out)[R]) {
  u64 part[R][2];
#pragma unroll
  for (int r = 0; r < R; ++r) part[r][0] = part[r][1] = 0ull;
  if (NP % 4 == 0) {
    const float4 *v4 = (const float4 *)vec;
#pragma unroll
    for (int i4 = 0; i4 < NP / 4; ++i4) {
      const float4 v = v4[i4];
      const u64 v01 = pk(v.x, v.y), v23 = pk(v.z, v.w);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        part[r][0] = fma2(v01, M[r][2 * i4], part[r][0]);
        part[r][1] = fma2(v23, M[r][2 * i4 + 1], part[r][1]);
      }
    }
  } else {                                                   // NP == 2
    const float2 v = *(const float2 *)vec;
#pragma unroll
    for (int r = 0; r < R; ++r) part[r][0] = fma2(pk(v.x, v.y), M[r][0], part[r][0]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = hsum2(part[r][0], part[r][1]);
}

// The same dot product for one matrix line, also accumulating this lane's row of the outer product
// alpha^_t(i) u_{t+1}(j) (the transition statistics): the u pairs are already in registers, so the
// statistics cost one more FFMA2 per pair instead of a separate pass over alpha^ and u (k_xi).
template <int NP>
__device__ __forceinline__ float matvec_xi(const float *vec, const u64 (&M)[(NP + 1) / 2], float al, u64 (&X)[(NP + 1) / 2]) {
  const u64 al2 = pk(al, al);
  u64 p0 = 0ull, p1 = 0ull;
  if (NP % 4 == 0) {
    const float4 *v4 = (const float4 *)vec;
#pragma unroll
    for (int i4 = 0; i4 < NP / 4; ++i4) {
      const float4 v = v4[i4];
      const u64 v01 = pk(v.x, v.y), v23 = pk(v.z, v.w);
      p0 = fma2(v01, M[2 * i4], p0);
      p1 = fma2(v23, M[2 * i4 + 1], p1);
      X[2 * i4] = fma2(v01, al2, X[2 * i4]);
      X[2 * i4 + 1] = fma2(v23, al2, X[2 * i4 + 1]);
    }
  } else {                                                   // NP == 2
    const float2 v = *(const float2 *)vec;
    p0 = fma2(pk(v.x, v.y), M[0], p0);
    X[0] = fma2(pk(v.x, v.y), al2, X[0]);
  }
  return hsum2(p0, p1);
}

// ---- forward: alpha^_t(j) = b~_t(j) sum_i alpha^_{t-1}(i) a_ij / c_t ;  LL += ln c_t + m_t
template <int LG, int R, int SQ>
__global__ void __launch_bounds__(FB_WARPS * 32, (R == 1 && LG < 32) ? 4 : 0) k_forward2(Fb2Args a) {
  constexpr int NP = fb2_np(LG, R), PG = 32 / LG;
  constexpr int PF = R == 1 ? 3 : 1;                        // emission rows in flight ahead of the step
  __shared__ __align__(16) float sv[FB_WARPS][PG][SQ][NP < 4 ? 4 : NP];
  const int N = a.dm.N, Npad = a.dm.Npad;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lj = lane % LG, gi = lane / LG;
  const unsigned mask = gmask<LG>();
  const int base = gi * LG;                                  // first lane of this group
  const bool idle = (32 % LG != 0) && gi >= PG;              // lanes past the last whole group (LG does not divide 32)
  if constexpr ((NP < 4 ? 4 : NP) > LG * R) {                // pad entries of the exchanged vectors stay zero
    if (!idle) {
#pragma unroll
      for (int q = 0; q < SQ; ++q)
        for (int i = LG * R + lj; i < (NP < 4 ? 4 : NP); i += LG) sv[w][gi][q][i] = 0.f;
    }
    __syncwarp(mask);
  }
  u64 Acol[R][(NP + 1) / 2];
  float pi[R];
  int jr[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    jr[r] = lj + LG * r;
    pi[r] = jr[r] < N ? a.tb.pi[jr[r]] : 0.f;
#pragma unroll
    for (int i = 0; i < NP; i += 2) {
      const float lo = (i < N && jr[r] < N) ? a.tb.A[i * Npad + jr[r]] : 0.f;
      const float hi = (i + 1 < N && jr[r] < N) ? a.tb.A[(i + 1) * Npad + jr[r]] : 0.f;
      Acol[r][i / 2] = pk(lo, hi);
    }
  }
  const int64_t ngroups = (int64_t)gridDim.x * FB_WARPS * PG;
  const int64_t gid = ((int64_t)blockIdx.x * FB_WARPS + w) * PG + gi;
  double ll = 0.0;
  for (int64_t sb = idle ? a.s1 : a.s0 + gid * SQ; sb < a.s1; sb += ngroups * SQ) {
    int64_t row[SQ], pt[SQ];
    int T[SQ], Tmax = 0;
    // Only the product alpha^_{t-1} A, the normaliser and the rescale sit on the step-to-step dependency chain:
    // the emission factor b~_t = exp(E_t - max_j E_t) of the NEXT step is formed while this step's chain runs
    // (its maximum rides in the same shuffle rounds as the normaliser), and the emission rows are requested PF
    // steps ahead, so neither a load nor the second reduction is ever waited for.
    float alpha[SQ][R], ering[PF][SQ][R], bcur[SQ][R], mcur[SQ];
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
      const int64_t s = sb + q;
      T[q] = s < a.s1 ? (int)(a.off[s + 1] - a.off[s]) : 0;
      pt[q] = s < a.s1 ? a.off[s] - a.p0 : 0;
      row[q] = pt[q] * N;
      Tmax = T[q] > Tmax ? T[q] : Tmax;
    }
    auto load_e = [&](int q, int r, int tt) -> float {
      return (tt < T[q] && jr[r] < N) ? a.E[row[q] + (int64_t)tt * N + jr[r]] : -INFINITY;
    };
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
      mcur[q] = -INFINITY;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        alpha[q][r] = 0.f;
        bcur[q][r] = load_e(q, r, 0);
        mcur[q] = fmaxf(mcur[q], bcur[q][r]);
#pragma unroll
        for (int i = 0; i < PF; ++i) ering[i][q][r] = load_e(q, r, 1 + i);
      }
    }
    GroupRed<LG>::template max_only<SQ>(mcur, mask, lj, base);
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
      if (mcur[q] == -INFINITY) mcur[q] = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) bcur[q][r] = ex2_approx((bcur[q][r] - mcur[q]) * (float)DNBC_LOG2E);
    }
    // The ring of requested rows is indexed statically: the step loop is unrolled PF times and step t consumes
    // and refills slot t % PF.  A ring that moves up by register copies touches every requested row one step after
    // its load was issued, and the copy waits for it (ncu, profiles/r02t_ncu_n10_summary.txt: 44 % of the forward
    // kernel's stall samples on the first copy at the loop top): the rows were in effect requested one step ahead.
    for (int t0 = 0; t0 < Tmax; t0 += PF) {
#pragma unroll
     for (int u = 0; u < PF; ++u) {
      const int t = t0 + u;
      if (t >= Tmax) break;
      // phases run over all SQ sequences so their dependency chains interleave; only the two
      // __syncwarp()s around the shared-memory exchange order the instruction stream
      float acc[SQ][R], csum[SQ], mnext[SQ];
      if (t == 0) {
#pragma unroll
        for (int q = 0; q < SQ; ++q)
#pragma unroll
          for (int r = 0; r < R; ++r) acc[q][r] = pi[r];
      } else {
#pragma unroll
        for (int q = 0; q < SQ; ++q)
#pragma unroll
          for (int r = 0; r < R; ++r) sv[w][gi][q][jr[r]] = alpha[q][r];
        __syncwarp(mask);
#pragma unroll
        for (int q = 0; q < SQ; ++q) matvec<NP, R>(sv[w][gi][q], Acol, acc[q]);
        __syncwarp(mask);
      }
#pragma unroll
      for (int q = 0; q < SQ; ++q) {
        csum[q] = 0.f;
        mnext[q] = ering[u][q][0];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[q][r] *= bcur[q][r];
          csum[q] += acc[q][r];
          mnext[q] = fmaxf(mnext[q], ering[u][q][r]);
        }
      }
      GroupRed<LG>::template sum_max<SQ>(csum, mnext, mask, lj, base);
#pragma unroll
      for (int q = 0; q < SQ; ++q) {
        const bool live = t < T[q];
        const float inv = rcp_approx(csum[q]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          alpha[q][r] = acc[q][r] * inv;
          if (live && jr[r] < N) a.AL[row[q] + (int64_t)t * N + jr[r]] = alpha[q][r];
        }
        if (live && lj == 0) {
          a.cs[pt[q] + t] = csum[q];
          ll += (double)(lg2_approx(csum[q]) * (float)DNBC_LN2) + (double)mcur[q];
        }
        // next step's emission factor; its ring slot is refilled with the row PF steps further on
        mcur[q] = mnext[q] == -INFINITY ? 0.f : mnext[q];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          bcur[q][r] = ex2_approx((ering[u][q][r] - mcur[q]) * (float)DNBC_LOG2E);
          ering[u][q][r] = load_e(q, r, t + 1 + PF);
        }
      }
     }
    }
  }
  if (lj == 0 && ll != 0.0) atomicAdd(a.stats, ll);
}

// ---- backward: beta^_t(i) = sum_j a_ij u_t(j),  u_t(j) = b~_{t+1}(j) beta^_{t+1}(j) / c_{t+1};
//      gamma_t = alpha^_t beta^_t;  pi-stat += gamma_0
//      XI (R == 1 only): xi-stat[i][j] += a_ij sum_t alpha^_t(i) u_{t+1}(j) accumulated in registers (fp32
//      over windows of ~4096 points, flushed with fp64 atomics); u is then not written to HBM at all
template <int LG, int R, int SQ, bool XI>
__global__ void __launch_bounds__(FB_WARPS * 32, (R == 1 && LG < 32) ? 4 : 0) k_backward2(Fb2Args a) {
  static_assert(!XI || R == 1, "fused transition statistics need the whole row in one register line");
  constexpr int NP = fb2_np(LG, R), PG = 32 / LG;
  // (three rings of rows here: two slots each measure faster than three -- 1.21 vs 1.28 ms per 2e7 points at N = 10 --
  // and four spill: 1.59 ms; the forward sweep has one ring and wants three, 0.80 vs 0.83 ms)
  constexpr int PF = R == 1 ? 2 : 1;                        // rows in flight ahead of the step
  __shared__ __align__(16) float sv[FB_WARPS][PG][SQ][NP < 4 ? 4 : NP];
  const int N = a.dm.N, Npad = a.dm.Npad;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lj = lane % LG, gi = lane / LG;
  const unsigned mask = gmask<LG>();
  const int base = gi * LG;                                  // first lane of this group
  const bool idle = (32 % LG != 0) && gi >= PG;              // lanes past the last whole group (LG does not divide 32)
  if constexpr ((NP < 4 ? 4 : NP) > LG * R) {                // pad entries of the exchanged vectors stay zero
    if (!idle) {
#pragma unroll
      for (int q = 0; q < SQ; ++q)
        for (int i = LG * R + lj; i < (NP < 4 ? 4 : NP); i += LG) sv[w][gi][q][i] = 0.f;
    }
    __syncwarp(mask);
  }
  u64 Arow[R][(NP + 1) / 2];
  int ir[R];
  double piacc[R];
  u64 xacc[XI ? (NP + 1) / 2 : 1];
  int xcount = 0;
#pragma unroll
  for (int jj = 0; jj < (XI ? (NP + 1) / 2 : 1); ++jj) xacc[jj] = 0ull;
  auto xi_flush = [&]() {
    if constexpr (XI) {
#pragma unroll
      for (int jj = 0; jj < (XI ? (NP + 1) / 2 : 1); ++jj) {
        float x0, x1, a0, a1;
        unpk(xacc[jj], x0, x1);
        unpk(Arow[0][jj], a0, a1);
        if (ir[0] < N && 2 * jj < N && x0 != 0.f)
          atomicAdd(a.stats + a.st_xi + (int64_t)ir[0] * N + 2 * jj, (double)x0 * (double)a0);
        if (ir[0] < N && 2 * jj + 1 < N && x1 != 0.f)
          atomicAdd(a.stats + a.st_xi + (int64_t)ir[0] * N + 2 * jj + 1, (double)x1 * (double)a1);
        xacc[jj] = 0ull;
      }
      xcount = 0;
    }
  };
#pragma unroll
  for (int r = 0; r < R; ++r) {
    ir[r] = lj + LG * r;
    piacc[r] = 0.0;
#pragma unroll
    for (int jj = 0; jj < NP; jj += 2) {
      const float lo = (jj < N && ir[r] < N) ? a.tb.A[ir[r] * Npad + jj] : 0.f;
      const float hi = (jj + 1 < N && ir[r] < N) ? a.tb.A[ir[r] * Npad + jj + 1] : 0.f;
      Arow[r][jj / 2] = pk(lo, hi);
    }
  }
  const int64_t ngroups = (int64_t)gridDim.x * FB_WARPS * PG;
  const int64_t gid = ((int64_t)blockIdx.x * FB_WARPS + w) * PG + gi;
  for (int64_t sb = idle ? a.s1 : a.s0 + gid * SQ; sb < a.s1; sb += ngroups * SQ) {
    int64_t row[SQ], pt[SQ];
    int T[SQ], Tmax = 0;
    // As in the forward sweep, only u = bu * beta, the product A u, the alpha.beta dot product and the rescale sit
    // on the dependency chain: bu = exp(E_{t+1} - max) / c_{t+1} of the NEXT iteration is formed while this one
    // runs (its maximum shares the shuffle rounds of the dot product) and alpha^ / E / c rows are requested PF
    // iterations ahead.  Iteration k works at t = T - 1 - k and needs alpha^_t, E_{t+1}, c_{t+1}.
    float beta[SQ][R], alr[PF][SQ][R], ering[PF][SQ][R], cring[PF][SQ], bu[SQ][R];
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
      const int64_t s = sb + q;
      T[q] = s < a.s1 ? (int)(a.off[s + 1] - a.off[s]) : 0;
      pt[q] = s < a.s1 ? a.off[s] - a.p0 : 0;
      row[q] = pt[q] * N;
      Tmax = T[q] > Tmax ? T[q] : Tmax;
    }
    auto load_al = [&](int q, int r, int tt) -> float {
      return (tt >= 0 && ir[r] < N) ? a.AL[row[q] + (int64_t)tt * N + ir[r]] : 0.f;
    };
    auto load_e = [&](int q, int r, int tt) -> float {       // E_tt as the "t+1" quantity of step tt-1
      return (tt >= 1 && ir[r] < N) ? a.E[row[q] + (int64_t)tt * N + ir[r]] : -INFINITY;
    };
    auto load_c = [&](int q, int tt) -> float { return tt >= 1 ? a.cs[pt[q] + tt] : 1.f; };
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
#pragma unroll
      for (int i = 0; i < PF; ++i) cring[i][q] = load_c(q, T[q] - 1 - i);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        beta[q][r] = 1.f;
        bu[q][r] = 0.f;
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          alr[i][q][r] = load_al(q, r, T[q] - 1 - i);
          ering[i][q][r] = load_e(q, r, T[q] - 1 - i);
        }
      }
    }
    // sequences of a pair are aligned at their LAST step: step index k counts back from it
    // (ring slots are indexed statically, k % PF, as in the forward sweep)
    for (int k0 = 0; k0 < Tmax; k0 += PF) {
#pragma unroll
     for (int u = 0; u < PF; ++u) {
      const int k = k0 + u;
      if (k >= Tmax) break;
      float al[SQ][R], mnext[SQ];
#pragma unroll
      for (int q = 0; q < SQ; ++q)
#pragma unroll
        for (int r = 0; r < R; ++r) al[q][r] = alr[u][q][r];
      if (k > 0) {
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const int t = T[q] - 1 - k;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float u = bu[q][r] * beta[q][r];
            if (!XI && t >= 0 && ir[r] < N) a.U[row[q] + (int64_t)t * N + ir[r]] = u;
            sv[w][gi][q][ir[r]] = (XI && t < 0) ? 0.f : u;
          }
        }
        __syncwarp(mask);
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          if constexpr (XI) beta[q][0] = matvec_xi<NP>(sv[w][gi][q], Arow[0], al[q][0], xacc);
          else matvec<NP, R>(sv[w][gi][q], Arow, beta[q]);
        }
        __syncwarp(mask);
      } else if (!XI) {
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
          const int t = T[q] - 1;
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (t >= 0 && ir[r] < N) a.U[row[q] + (int64_t)t * N + ir[r]] = 0.f;
        }
      }
      // sum_i alpha^_t(i) beta^_t(i) is 1 in exact arithmetic; in fp32 the product of a thousand
      // approximate normalisers drifts by ~3e-6, so beta^_t is rescaled to make it 1 at every step
      float gsumq[SQ];
#pragma unroll
      for (int q = 0; q < SQ; ++q) {
        gsumq[q] = 0.f;
        mnext[q] = ering[u][q][0];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          gsumq[q] = fmaf(al[q][r], beta[q][r], gsumq[q]);
          mnext[q] = fmaxf(mnext[q], ering[u][q][r]);
        }
      }
      GroupRed<LG>::template sum_max<SQ>(gsumq, mnext, mask, lj, base);
#pragma unroll
      for (int q = 0; q < SQ; ++q) {
        const int t = T[q] - 1 - k;
        const float fix = gsumq[q] > 0.f ? rcp_approx(gsumq[q]) : 1.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          beta[q][r] *= fix;
          const float g = al[q][r] * beta[q][r];
          if (t >= 0 && ir[r] < N) a.GA[row[q] + (int64_t)t * N + ir[r]] = g;
          if (t == 0) piacc[r] += (double)g;
        }
        // next iteration's factor (E_t, c_t are its "t+1" quantities); the ring slots are refilled PF steps ahead
        const float mn = mnext[q] == -INFINITY ? 0.f : mnext[q];
        const float ic = rcp_approx(cring[u][q]);
        cring[u][q] = load_c(q, t - PF);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          bu[q][r] = ex2_approx((ering[u][q][r] - mn) * (float)DNBC_LOG2E) * ic;
          ering[u][q][r] = load_e(q, r, t - PF);
          alr[u][q][r] = load_al(q, r, t - PF);
        }
      }
     }
    }
    xcount += Tmax * SQ;
    if (XI && xcount >= 4096) xi_flush();
  }
  xi_flush();
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (ir[r] < N && piacc[r] != 0.0) atomicAdd(a.stats + a.st_pi + ir[r], piacc[r]);
}

template <int LG, int R, int SQ>
static void launch_fb2_sq(dnbc_ctx *c, const Fb2Args &a, bool backward) {
  constexpr int PG = 32 / LG;
  const int64_t nseq = a.s1 - a.s0;
  const int64_t per_block = (int64_t)FB_WARPS * PG * SQ;
  int64_t blocks = (nseq + per_block - 1) / per_block;
  const int64_t cap = (int64_t)c->sm_count * (LG * R > 32 ? 2 : 4);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (!backward) k_forward2<LG, R, SQ><<<(int)blocks, FB_WARPS * 32, 0, c->stream>>>(a);
  else if (R == 1 && fb2_fuses_xi(c)) k_backward2<LG, R, SQ, R == 1><<<(int)blocks, FB_WARPS * 32, 0, c->stream>>>(a);
  else k_backward2<LG, R, SQ, false><<<(int)blocks, FB_WARPS * 32, 0, c->stream>>>(a);
}

// A lane group advances SQ = 4 sequences at a time so that their dependency chains interleave.  When the
// whole chunk fits the GPU with one sequence per group (a README-sized data set: 1000 sequences), one
// sequence per group is faster: the step time of a lone warp is its instruction count, not its chain.
template <int LG, int R>
static void launch_fb2_t(dnbc_ctx *c, const Fb2Args &a, bool backward) {
  constexpr int PG = 32 / LG;
  const int64_t one_wave = (int64_t)c->sm_count * (LG * R > 32 ? 2 : 4) * FB_WARPS * PG;
  if (R == 1 && a.s1 - a.s0 <= one_wave && !c->opt_fb2_interleave) launch_fb2_sq<LG, R, 1>(c, a, backward);
  else launch_fb2_sq<LG, R, 4>(c, a, backward);
}

bool fb2_supported(const dnbc_ctx *c) { return c->dm.Npad <= 64; }
// up to 32 states the backward sweep accumulates the transition statistics itself (launch_xi is then a no-op)
bool fb2_fuses_xi(const dnbc_ctx *c) {
  return !fb3_supported(c) && fb2_supported(c) && c->dm.Npad <= 32;
}

void launch_fb2(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward) {
  StatsLayout sl = stats_layout(c->dm);
  Fb2Args a{c->dm, c->tables(), c->d_off, s0, s1, p0, c->E, c->AL, c->GA, c->U, c->cs, c->d_stats, sl.pi, sl.xi};
  // lanes per group: the padded state count, or the exact one where that puts more sequences into a warp
  // (the rule of lane_group() in k_mixture.cu: N = 3, 5, 6, 9, 10)
  const int N = c->dm.N, Npad = c->dm.Npad;
  // (a chunk that fits one wave of power-of-two groups is latency-bound: the exact groups' extra shuffle round
  // would only lengthen its chain)
  const int lgp = Npad < 32 ? Npad : 32;
  const bool one_wave = s1 - s0 <= (int64_t)c->sm_count * 4 * FB_WARPS * (32 / lgp) && !c->opt_fb2_interleave;
  const int lg = (!one_wave && N < 32 && 32 / N > 32 / lgp) ? N : Npad;
  switch (lg) {
    case 1:
    case 2: launch_fb2_t<2, 1>(c, a, backward); break;
    case 3: launch_fb2_t<3, 1>(c, a, backward); break;
    case 4: launch_fb2_t<4, 1>(c, a, backward); break;
    case 5: launch_fb2_t<5, 1>(c, a, backward); break;
    case 6: launch_fb2_t<6, 1>(c, a, backward); break;
    case 8: launch_fb2_t<8, 1>(c, a, backward); break;
    case 9: launch_fb2_t<9, 1>(c, a, backward); break;
    case 10: launch_fb2_t<10, 1>(c, a, backward); break;
    case 16: launch_fb2_t<16, 1>(c, a, backward); break;
    case 32: launch_fb2_t<32, 1>(c, a, backward); break;
    default: launch_fb2_t<32, 2>(c, a, backward); break;   // Npad == 64
  }
}
