// k_stats.cu -- K3: E-step sufficient statistics (gamma sums, categorical histograms,
// GMM responsibilities) in fp64, and K4: the M-step normalisation.
//
// Reference arithmetic these generalise (hard labels -> posteriors):
//   DiscreteEdge.learn / learnFinalize     core/src/main/scala/Edge.scala:46-59
//   learnDiscreteEmissions                 core/src/main/scala/DynamicNaiveBayesianClassifier.scala:163-178
//   EM1D.run E-step :92-101, M-step :104-132 core/src/main/java/Tools/ExpectationMaximization1D.java
// With one-hot posteriors (clamped labels) the statistics equal the reference's counts
// and the M-step equals its normalisations (tests/test_bridge.py).
#include "dnbc_internal.cuh"

// ---------------------------------------------------------------------------------
// GMM statistics.  Thread <-> (continuous variable c, state j); its K_c components'
// coefficients and 3 K_c accumulators live in registers; the CTA streams over a slab
// of points reading gamma[p][j] (unit stride across lanes) and x[c][p] (broadcast).
// Moments are taken in the kernel's own standardised variable z = x*gs + gm and
// converted to moments about the fp64 old mean when a window is flushed, so fp32
// never sees the cancellation x^2 - mu^2.
// ---------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256) k_stats_gmm(Dims dm, DeviceTables tb, DataView dv, int64_t p0, int64_t np,
                                                   const float *__restrict__ GA, const double *__restrict__ mu_old,
                                                   double *__restrict__ gmm, int window) {
  const int N = dm.N, C = dm.C, Kmax = dm.Kmax, Npad = dm.Npad;
  const int q = blockIdx.y * blockDim.x + threadIdx.x;
  const bool valid = q < N * C;
  const int c = valid ? q / N : 0, j = valid ? q % N : 0;
  const int K = tb.K[c];
  float gs[KT], gm[KT], gc[KT], s0[KT], s1[KT], s2[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int64_t o = ((int64_t)c * Kmax + k) * Npad + j;
    const bool on = valid && k < K;
    gs[k] = on ? tb.gs[o] : 0.f;
    gm[k] = on ? tb.gm[o] : 0.f;
    gc[k] = on ? tb.gc[o] : -INFINITY;
    s0[k] = s1[k] = s2[k] = 0.f;
  }
  const int64_t slab = (np + gridDim.x - 1) / gridDim.x;
  const int64_t pb = (int64_t)blockIdx.x * slab;
  const int64_t pe = pb + slab < np ? pb + slab : np;
  const float *xcol = dv.cont + (int64_t)c * dv.P + p0;
  int inwin = 0;
  for (int64_t p = pb; p < pe; ++p) {
    const float g = valid ? __ldg(GA + p * N + j) : 0.f;
    const float x = __ldg(xcol + p);
    if (g > 0.f) {
      float z[KT], e[KT], sum = 0.f, mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        z[k] = fmaf(x, gs[k], gm[k]);
        const float a = fmaf(-z[k], z[k], gc[k]);
        mx = fmaxf(mx, a);
        e[k] = ex2_approx(a);
        sum += e[k];
      }
      if (sum < 1e-30f && mx != -INFINITY) {  // all components underflowed: redo shifted
        sum = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          e[k] = ex2_approx(fmaf(-z[k], z[k], gc[k]) - mx);
          sum += e[k];
        }
      }
      if (sum > 0.f) {
        const float inv = g / sum;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const float r = e[k] * inv;
          const float rz = r * z[k];
          s0[k] += r;
          s1[k] += rz;
          s2[k] = fmaf(rz, z[k], s2[k]);
        }
      }
    }
    if (++inwin == window || p + 1 == pe) {
      inwin = 0;
      if (valid) {
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < K && s0[k] != 0.f) {
            // z = (x - mu')*s with mu' = -gm/gs the fp32-rounded centre; re-centre on mu_old
            const int64_t dst = ((int64_t)c * N + j) * Kmax + k;
            const double s = (double)gs[k];
            double m0 = (double)s0[k], m1, m2;
            if (s > 0.0) {
              const double delta = -(double)gm[k] / s - mu_old[dst];
              const double a1 = (double)s1[k] / s, a2 = (double)s2[k] / (s * s);
              m1 = a1 + delta * m0;
              m2 = a2 + 2.0 * delta * a1 + delta * delta * m0;
            } else {
              m1 = 0.0;
              m2 = 0.0;
            }
            atomicAdd(gmm + 3 * dst + 0, m0);
            atomicAdd(gmm + 3 * dst + 1, m1);
            atomicAdd(gmm + 3 * dst + 2, m2);
          }
          s0[k] = s1[k] = s2[k] = 0.f;
        }
      }
    }
  }
}

template <int KT>
static void launch_stats_gmm_t(dnbc_ctx *c, int64_t p0, int64_t np) {
  StatsLayout sl = stats_layout(c->dm);
  const int pairs = c->dm.N * c->dm.C;
  const int tpb = pairs >= 256 ? 256 : ((pairs + 31) / 32) * 32;
  const int gy = (pairs + tpb - 1) / tpb;
  const int window = 2048;
  int64_t gx = (np + window - 1) / window;
  int64_t cap = (int64_t)c->sm_count * 8 / gy;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  k_stats_gmm<KT><<<dim3((unsigned)gx, gy), tpb, 0, c->stream>>>(c->dm, c->tables(), c->data(), p0, np, c->GA, c->d_mu,
                                                                 c->d_stats + sl.gmm, window);
}

void launch_stats_gmm(dnbc_ctx *c, int64_t p0, int64_t np) {
  if (np <= 0 || c->dm.C == 0) return;
  FamTimer ft(c, FAM_STATS_GMM);
  const int K = c->dm.Kmax;
  if (K <= 1) launch_stats_gmm_t<1>(c, p0, np);
  else if (K <= 2) launch_stats_gmm_t<2>(c, p0, np);
  else if (K <= 4) launch_stats_gmm_t<4>(c, p0, np);
  else if (K <= 8) launch_stats_gmm_t<8>(c, p0, np);
  else if (K <= 16) launch_stats_gmm_t<16>(c, p0, np);
  else launch_stats_gmm_t<32>(c, p0, np);
}

// ---------------------------------------------------------------------------------
// Categorical statistics + gamma sums.  Thread <-> (variable d, state j) with d == D
// standing for the plain gamma sum; each thread owns a private column of Vmax fp32
// bins in shared memory (layout [v][thread]: conflict-free, no atomics), flushed to
// the fp64 buffer once per window.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stats_cat(Dims dm, DataView dv, int64_t p0, int64_t np,
                                                   const float *__restrict__ GA, double *__restrict__ cat,
                                                   double *__restrict__ gsum, int window) {
  extern __shared__ float bins[];  // [Vmax][blockDim.x]
  const int N = dm.N, D = dm.D, Vmax = dm.Vmax;
  const int q = blockIdx.y * blockDim.x + threadIdx.x;
  const bool valid = q < N * (D + 1);
  const int d = valid ? q / N : 0, j = valid ? q % N : 0;
  const bool is_sum = d == D;
  for (int v = 0; v < Vmax; ++v) bins[v * blockDim.x + threadIdx.x] = 0.f;
  float tot = 0.f;
  const int64_t slab = (np + gridDim.x - 1) / gridDim.x;
  const int64_t pb = (int64_t)blockIdx.x * slab;
  const int64_t pe = pb + slab < np ? pb + slab : np;
  const uint8_t *col = is_sum ? nullptr : dv.disc + (int64_t)d * dv.P + p0;
  int inwin = 0;
  for (int64_t p = pb; p < pe; ++p) {
    if (valid) {
      const float g = __ldg(GA + p * N + j);
      if (is_sum) {
        tot += g;
      } else {
        const int v = __ldg(col + p);
        if (v < Vmax) bins[v * blockDim.x + threadIdx.x] += g;
      }
    }
    if (++inwin == window || p + 1 == pe) {
      inwin = 0;
      if (valid) {
        if (is_sum) {
          if (tot != 0.f) atomicAdd(gsum + j, (double)tot);
          tot = 0.f;
        } else {
          for (int v = 0; v < Vmax; ++v) {
            const float b = bins[v * blockDim.x + threadIdx.x];
            if (b != 0.f) {
              atomicAdd(cat + ((int64_t)d * N + j) * Vmax + v, (double)b);
              bins[v * blockDim.x + threadIdx.x] = 0.f;
            }
          }
        }
      }
    }
  }
}

void launch_stats_cat(dnbc_ctx *c, int64_t p0, int64_t np) {
  if (np <= 0) return;
  FamTimer ft(c, FAM_STATS_CAT);
  StatsLayout sl = stats_layout(c->dm);
  const int pairs = c->dm.N * (c->dm.D + 1);
  int tpb = pairs >= 256 ? 256 : ((pairs + 31) / 32) * 32;
  while ((size_t)tpb * c->dm.Vmax * sizeof(float) > 64 * 1024 && tpb > 32) tpb /= 2;
  const int gy = (pairs + tpb - 1) / tpb;
  const int window = 4096;
  int64_t gx = (np + window - 1) / window;
  int64_t cap = (int64_t)c->sm_count * 8 / gy;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  const size_t smem = (size_t)tpb * c->dm.Vmax * sizeof(float);
  if (smem > 32 * 1024) cudaFuncSetAttribute(k_stats_cat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_stats_cat<<<dim3((unsigned)gx, gy), tpb, smem, c->stream>>>(c->dm, c->data(), p0, np, c->GA, c->d_stats + sl.cat,
                                                                c->d_stats + sl.gamma, window);
}

// ---------------------------------------------------------------------------------
// K4: M-step.  One thread per row / edge, fp64 throughout (the arrays are tiny):
//   pi_j = pi-stat_j / S                               (Appendix B)
//   a_ij = xi_ij / sum_j xi_ij                         (Edge.scala:53-59 with soft counts)
//   B_d[j][v] = cat / sum_v cat                        (same)
//   w = s0 / sum_k s0, mu += s1/s0, var = s2/s0 - (s1/s0)^2   (EM1D.java:104-132)
// A row / edge whose denominator is zero keeps its previous parameters.
// The reference's EM1D has no variance floor (a collapsed component gives var = 0 and NaN
// responsibilities, EM1D.java:92-101); Baum-Welch over many iterations needs one: a new variance
// below var_floor_rel x (pooled variance of that observed variable) is raised to it and counted.
// ---------------------------------------------------------------------------------
struct MstepArgs {
  Dims dm;
  const int32_t *V, *K;
  const double *st;
  StatsLayout sl;
  double S;
  double *pi, *A, *B, *w, *mu, *var;
  double floor_rel;
  double *pooled;                 // [C] pooled variance per continuous variable (k_pooled_variance)
  unsigned long long *floored;    // [1]
};

// pooled variance of variable c over every state and component, from the moments centred on the
// old means: x = z + mu_old  =>  sum x = s1 + mu s0,  sum x^2 = s2 + 2 mu s1 + mu^2 s0
__global__ void k_pooled_variance(MstepArgs a) {
  const int N = a.dm.N, Kmax = a.dm.Kmax, c = blockIdx.x;
  __shared__ double sh[3][128];
  double n = 0.0, sx = 0.0, sxx = 0.0;
  for (int q = threadIdx.x; q < N * Kmax; q += blockDim.x) {
    if (q % Kmax >= a.K[c]) continue;
    const int64_t e = (int64_t)c * N * Kmax + q;
    const double *x = a.st + a.sl.gmm + 3 * e;
    const double mu = a.mu[e];
    n += x[0];
    sx += x[1] + mu * x[0];
    sxx += x[2] + 2.0 * mu * x[1] + mu * mu * x[0];
  }
  sh[0][threadIdx.x] = n; sh[1][threadIdx.x] = sx; sh[2][threadIdx.x] = sxx;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int q = 0; q < 3; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double tn = sh[0][0];
    double v = 0.0;
    if (tn > 0.0) {
      const double m = sh[1][0] / tn;
      v = sh[2][0] / tn - m * m;
    }
    a.pooled[c] = v > 0.0 ? v : 0.0;
  }
}

__global__ void k_mstep(MstepArgs a) {
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax;
  const int rows = N + N * D + N * C + 1;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    if (r < N) {
      const double *x = a.st + a.sl.xi + (int64_t)r * N;
      double sum = 0.0;
      for (int j = 0; j < N; ++j) sum += x[j];
      if (sum > 0.0)
        for (int j = 0; j < N; ++j) a.A[(int64_t)r * N + j] = x[j] / sum;
    } else if (r < N + N * D) {
      const int q = r - N, d = q / N, j = q % N;
      const double *x = a.st + a.sl.cat + ((int64_t)d * N + j) * Vmax;
      double sum = 0.0;
      for (int v = 0; v < a.V[d]; ++v) sum += x[v];
      if (sum > 0.0)
        for (int v = 0; v < a.V[d]; ++v) a.B[((int64_t)d * N + j) * Vmax + v] = x[v] / sum;
    } else if (r < N + N * D + N * C) {
      const int q = r - N - N * D, c = q / N, j = q % N;
      const int64_t base = ((int64_t)c * N + j) * Kmax;
      const double *x = a.st + a.sl.gmm + 3 * base;
      const double vfloor = a.floor_rel * a.pooled[c];
      double tot = 0.0;
      for (int k = 0; k < a.K[c]; ++k) tot += x[3 * k];
      if (tot > 0.0)
        for (int k = 0; k < a.K[c]; ++k) {
          const double s0 = x[3 * k], s1 = x[3 * k + 1], s2 = x[3 * k + 2];
          a.w[base + k] = s0 / tot;
          if (s0 > 0.0) {
            const double sh = s1 / s0;
            double v = s2 / s0 - sh * sh;
            if (!(v >= vfloor)) {      // also catches a NaN from cancellation on an empty component
              v = vfloor;
              atomicAdd(a.floored, 1ull);
            }
            a.mu[base + k] += sh;
            a.var[base + k] = v;
          }
        }
    } else {
      for (int j = 0; j < N; ++j) a.pi[j] = a.st[a.sl.pi + j] / a.S;
    }
  }
}

void launch_mstep(dnbc_ctx *c, int64_t total_sequences) {
  {
    FamTimer ft(c, FAM_MSTEP);
    MstepArgs a{c->dm, c->d_V, c->d_K, c->d_stats, stats_layout(c->dm), (double)total_sequences,
                c->d_pi, c->d_A, c->d_B, c->d_w, c->d_mu, c->d_var, c->var_floor_rel, c->d_pooled, c->d_floored};
    const int rows = c->dm.N * (1 + c->dm.D + c->dm.C) + 1;
    if (c->dm.C > 0) {
      k_pooled_variance<<<c->dm.C, 128, 0, c->stream>>>(a);
      c->launches++;
    }
    k_mstep<<<(rows + 127) / 128, 128, 0, c->stream>>>(a);
  }
  c->tables64_valid = false;
  launch_prepare_tables(c);
}
