// k_forward_backward.cu -- K2: scaled forward / backward recursions and the xi
// (expected transition count) contraction.
//
// The reference has NO forward-backward: its learner is supervised
// counting (DynamicNaiveBayesianClassifier.scala:124-161).  These kernels implement
// SURVEY Appendix B and are checked against the fp64 oracle (oracle/dnbc_oracle.c
// fb_one) and, through the clamped-label bridge test, against the reference's
// supervised counts.
//
// Mapping: one "lane group" of LG = min(32, Npad) lanes owns one sequence; lane l
// holds states l, l+LG, ... (R per lane) of alpha / beta in registers.  alpha_{t-1}(i)
// is broadcast with a warp shuffle, A[i][*] comes from shared memory with unit stride
// across lanes, and the per-step normaliser c_t is a shuffle reduction inside the group.
// Groups are persistent: each walks sequences s, s+G, s+2G, ... of the chunk.
#include "dnbc_internal.cuh"

template <int LG>
__device__ __forceinline__ float group_max(float v, unsigned mask) {
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, LG));
  return v;
}
template <int LG>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, LG);
  return v;
}
template <int LG>
__device__ __forceinline__ unsigned group_mask() {
  if (LG == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << (LG & 31)) - 1u) << ((lane / LG) * LG);
}

struct FbArgs {
  Dims dm;
  DeviceTables tb;
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;   // [np][N] natural-log emissions (chunk-relative rows)
  float *AL;        // [np][N] scaled alpha
  float *GA;        // [np][N] gamma
  float *U;         // [np][N] u_t(j) = b~_{t+1}(j) beta_{t+1}(j) / c_{t+1}  (0 on the last step)
  float *cs;        // [np]    c_t
  double *stats;    // packed statistics
  int64_t st_pi;    // offset of the pi-stat
};

// ---- forward: alpha^_t(j) = b~_t(j) sum_i alpha^_{t-1}(i) a_ij / c_t ;  LL += ln c_t + m_t
template <int LG, int R, bool ASMEM>
__global__ void __launch_bounds__(256) k_forward(FbArgs a) {
  extern __shared__ float sm_A[];
  const int N = a.dm.N, Npad = a.dm.Npad;
  if (ASMEM) {
    for (int q = threadIdx.x; q < N * Npad; q += blockDim.x) sm_A[q] = a.tb.A[q];
    __syncthreads();
  }
  const float *__restrict__ Am = ASMEM ? sm_A : a.tb.A;
  const int lane = threadIdx.x % LG;
  const int gpb = blockDim.x / LG;
  const unsigned mask = group_mask<LG>();
  double ll = 0.0;
  for (int64_t s = a.s0 + (int64_t)blockIdx.x * gpb + threadIdx.x / LG; s < a.s1; s += (int64_t)gridDim.x * gpb) {
    const int64_t b = a.off[s] - a.p0;
    const int64_t T = a.off[s + 1] - a.off[s];
    float alpha[R], en[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + LG * r;
      en[r] = (T > 0 && j < N) ? a.E[b * N + j] : -INFINITY;
      alpha[r] = 0.f;
    }
    for (int64_t t = 0; t < T; ++t) {
      float e[R], acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + LG * r;
        e[r] = en[r];
        en[r] = (t + 1 < T && j < N) ? a.E[(b + t + 1) * N + j] : -INFINITY;
      }
      float m = e[0];
#pragma unroll
      for (int r = 1; r < R; ++r) m = fmaxf(m, e[r]);
      m = group_max<LG>(m, mask);
      if (m == -INFINITY) m = 0.f;
      if (t == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int j = lane + LG * r;
          acc[r] = j < N ? a.tb.pi[j] : 0.f;
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
        for (int rp = 0; rp < R; ++rp) {
#pragma unroll 4
          for (int lp = 0; lp < LG; ++lp) {
            const int i = lp + LG * rp;
            const float ai = __shfl_sync(mask, alpha[rp], lp, LG);
            if (i < N && ai != 0.f) {
              const float *row = Am + i * Npad + lane;
#pragma unroll
              for (int r = 0; r < R; ++r) acc[r] = fmaf(ai, row[LG * r], acc[r]);
            }
          }
        }
      }
      float csum = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r] *= ex2_approx((e[r] - m) * (float)DNBC_LOG2E);
        csum += acc[r];
      }
      csum = group_sum<LG>(csum, mask);
      const float inv = 1.0f / csum;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + LG * r;
        alpha[r] = acc[r] * inv;
        if (j < N) a.AL[(b + t) * N + j] = alpha[r];
      }
      if (lane == 0) {
        a.cs[b + t] = csum;
        ll += (double)logf(csum) + (double)m;
      }
    }
  }
  if (lane == 0 && ll != 0.0) atomicAdd(a.stats, ll);
}

// ---- backward: beta^_t(i) = sum_j a_ij u_t(j),  u_t(j) = b~_{t+1}(j) beta^_{t+1}(j) / c_{t+1};
//      gamma_t = alpha^_t beta^_t;  pi-stat += gamma_0
template <int LG, int R, bool ASMEM>
__global__ void __launch_bounds__(256) k_backward(FbArgs a) {
  extern __shared__ float sm_At[];
  const int N = a.dm.N, Npad = a.dm.Npad;
  if (ASMEM) {
    for (int q = threadIdx.x; q < N * Npad; q += blockDim.x) sm_At[q] = a.tb.At[q];
    __syncthreads();
  }
  const float *__restrict__ Atm = ASMEM ? sm_At : a.tb.At;
  const int lane = threadIdx.x % LG;
  const int gpb = blockDim.x / LG;
  const unsigned mask = group_mask<LG>();
  double piacc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) piacc[r] = 0.0;
  for (int64_t s = a.s0 + (int64_t)blockIdx.x * gpb + threadIdx.x / LG; s < a.s1; s += (int64_t)gridDim.x * gpb) {
    const int64_t b = a.off[s] - a.p0;
    const int64_t T = a.off[s + 1] - a.off[s];
    if (T <= 0) continue;
    float beta[R], aln[R], en[R];
    float csn = 1.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + LG * r;
      beta[r] = 1.f;
      aln[r] = j < N ? a.AL[(b + T - 1) * N + j] : 0.f;
      en[r] = -INFINITY;
    }
    for (int64_t t = T - 1; t >= 0; --t) {
      float al[R], e[R];
      const float c_next = csn;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + LG * r;
        al[r] = aln[r];
        e[r] = en[r];
        // prefetch for step t-1: alpha_{t-1}, and E_t / c_t (the "t+1" quantities of step t-1)
        aln[r] = (t > 0 && j < N) ? a.AL[(b + t - 1) * N + j] : 0.f;
        en[r] = (t > 0 && j < N) ? a.E[(b + t) * N + j] : -INFINITY;
      }
      csn = t > 0 ? a.cs[b + t] : 1.f;
      if (t < T - 1) {
        float m = e[0];
#pragma unroll
        for (int r = 1; r < R; ++r) m = fmaxf(m, e[r]);
        m = group_max<LG>(m, mask);
        if (m == -INFINITY) m = 0.f;
        const float ic = 1.0f / c_next;
        float u[R], acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int j = lane + LG * r;
          u[r] = ex2_approx((e[r] - m) * (float)DNBC_LOG2E) * beta[r] * ic;
          if (j < N) a.U[(b + t) * N + j] = u[r];
          acc[r] = 0.f;
        }
#pragma unroll
        for (int rp = 0; rp < R; ++rp) {
#pragma unroll 4
          for (int lp = 0; lp < LG; ++lp) {
            const int j = lp + LG * rp;
            const float uj = __shfl_sync(mask, u[rp], lp, LG);
            if (j < N && uj != 0.f) {
              const float *row = Atm + j * Npad + lane;
#pragma unroll
              for (int r = 0; r < R; ++r) acc[r] = fmaf(uj, row[LG * r], acc[r]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) beta[r] = acc[r];
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int j = lane + LG * r;
          if (j < N) a.U[(b + t) * N + j] = 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + LG * r;
        const float g = al[r] * beta[r];
        if (j < N) a.GA[(b + t) * N + j] = g;
        if (t == 0) piacc[r] += (double)g;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + LG * r;
    if (j < N && piacc[r] != 0.0) atomicAdd(a.stats + a.st_pi + j, piacc[r]);
  }
}

template <int LG, int R>
static void launch_fb_t(dnbc_ctx *c, const FbArgs &a, bool backward) {
  const int N = c->dm.N, Npad = c->dm.Npad;
  const size_t smem = (size_t)N * Npad * sizeof(float);
  const bool asmem = smem <= 96 * 1024;
  const int gpb = 256 / LG;
  const int64_t nseq = a.s1 - a.s0;
  int64_t blocks = (nseq + gpb - 1) / gpb;
  const int64_t cap = (int64_t)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (asmem) {
    auto kf = backward ? k_backward<LG, R, true> : k_forward<LG, R, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kf<<<(int)blocks, 256, smem, c->stream>>>(a);
  } else {
    auto kf = backward ? k_backward<LG, R, false> : k_forward<LG, R, false>;
    kf<<<(int)blocks, 256, 0, c->stream>>>(a);
  }
}

static void launch_fb(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward) {
  if (s1 <= s0) return;
  FamTimer ft(c, backward ? FAM_BACKWARD : FAM_FORWARD);
  if (fbs_supported(c, s1 - s0)) return launch_fbs(c, s0, s1, p0, backward);   // small even N, many sequences: lane per sequence (k_fbs.cu)
  if (fb3_supported(c)) return launch_fb3(c, s0, s1, p0, backward);   // 33..64 states: mma.sync, 16 sequences per warp (k_fb3.cu)
  if (fb2_supported(c)) return launch_fb2(c, s0, s1, p0, backward);   // register-resident A (k_fb2.cu)
  if (fb_tc_supported(c)) {                                            // tcgen05 tensor cores (k_tc.cu)
    launch_fb_tc(c, s0, s1, p0, c->h_off[s1] - c->h_off[s0], backward);
    return;
  }
  StatsLayout sl = stats_layout(c->dm);
  FbArgs a{c->dm, c->tables(), c->d_off, s0, s1, p0, c->E, c->AL, c->GA, c->U, c->cs, c->d_stats, sl.pi};
  switch (c->dm.Npad) {
    case 1:
    case 2: launch_fb_t<2, 1>(c, a, backward); break;
    case 4: launch_fb_t<4, 1>(c, a, backward); break;
    case 8: launch_fb_t<8, 1>(c, a, backward); break;
    case 16: launch_fb_t<16, 1>(c, a, backward); break;
    case 32: launch_fb_t<32, 1>(c, a, backward); break;
    case 64: launch_fb_t<32, 2>(c, a, backward); break;
    case 128: launch_fb_t<32, 4>(c, a, backward); break;
    default: launch_fb_block(c, s0, s1, p0, backward); break;  // Npad >= 256: CTA per sequence (k_large_n.cu)
  }
}

void launch_forward(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0) { launch_fb(c, s0, s1, p0, false); }
void launch_backward(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0) { launch_fb(c, s0, s1, p0, true); }

// ---------------------------------------------------------------------------------
// xi contraction: Xi[i][j] = sum_p AL[p][i] * U[p][j]  (a [N x np] x [np x N] product over
// the point axis of the chunk), then xi-stat[i][j] += a_ij * Xi[i][j].
// Register-tiled FP32 on the CUDA cores: block tile (16 TM)^2, thread tile TM^2, 32
// points per shared-memory stage; partial sums are folded into fp64 every 4096 points.
// ---------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(256) k_xi(int N, int Npad, const float *__restrict__ AL, const float *__restrict__ U,
                                            int64_t np, const float *__restrict__ Amat, double *__restrict__ xi) {
  constexpr int BT = 16 * TM, KS = 32;
  __shared__ float sA[KS][BT + 4];
  __shared__ float sB[KS][BT + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int i0 = blockIdx.y * BT, j0 = blockIdx.x * BT;
  const int64_t slab = (np + gridDim.z - 1) / gridDim.z;
  const int64_t pb = (int64_t)blockIdx.z * slab;
  const int64_t pe = pb + slab < np ? pb + slab : np;
  float acc[TM][TM];
  double acc64[TM][TM];
#pragma unroll
  for (int x = 0; x < TM; ++x)
#pragma unroll
    for (int y = 0; y < TM; ++y) { acc[x][y] = 0.f; acc64[x][y] = 0.0; }
  int stage = 0;
  for (int64_t p = pb; p < pe; p += KS) {
    for (int q = threadIdx.x; q < KS * BT; q += 256) {
      const int kk = q / BT, cc = q % BT;
      const int64_t pp = p + kk;
      sA[kk][cc] = (pp < pe && i0 + cc < N) ? AL[pp * N + i0 + cc] : 0.f;
      sB[kk][cc] = (pp < pe && j0 + cc < N) ? U[pp * N + j0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < KS; ++kk) {
      float av[TM], bv[TM];
#pragma unroll
      for (int x = 0; x < TM; ++x) { av[x] = sA[kk][ty * TM + x]; bv[x] = sB[kk][tx * TM + x]; }
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TM; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();
    if (++stage == 128) {
      stage = 0;
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TM; ++y) { acc64[x][y] += (double)acc[x][y]; acc[x][y] = 0.f; }
    }
  }
#pragma unroll
  for (int x = 0; x < TM; ++x)
#pragma unroll
    for (int y = 0; y < TM; ++y) {
      const int i = i0 + ty * TM + x, j = j0 + tx * TM + y;
      const double v = acc64[x][y] + (double)acc[x][y];
      if (i < N && j < N && v != 0.0) atomicAdd(xi + (int64_t)i * N + j, v * (double)Amat[i * Npad + j]);
    }
}

void launch_xi(dnbc_ctx *c, int64_t np) {
  if (np <= 0) return;
  if (fb2_fuses_xi(c)) return;                               // accumulated inside k_backward2 (k_fb2.cu)
  FamTimer ft(c, FAM_XI);
  if (xi_tc_supported(c)) {                                  // N >= 64: tcgen05 GEMM (k_xi_tc.cu)
    launch_xi_tc(c, np);
    return;
  }
  StatsLayout sl = stats_layout(c->dm);
  const int N = c->dm.N;
  if (N <= 16) {
    int z = (int)((np + 4095) / 4096);
    if (z > c->sm_count * 8) z = c->sm_count * 8;
    k_xi<1><<<dim3(1, 1, z), 256, 0, c->stream>>>(N, c->dm.Npad, c->AL, c->U, np, c->t_A, c->d_stats + sl.xi);
  } else {
    const int tiles = (N + 63) / 64;
    int z = (int)((np + 4095) / 4096);
    int zcap = c->sm_count * 4 / (tiles * tiles);
    if (zcap < 1) zcap = 1;
    if (z > zcap) z = zcap;
    k_xi<4><<<dim3(tiles, tiles, z), 256, 0, c->stream>>>(N, c->dm.Npad, c->AL, c->U, np, c->t_A, c->d_stats + sl.xi);
  }
}
