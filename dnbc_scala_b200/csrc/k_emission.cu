// k_emission.cu -- K1: fused emission kernel (categorical log lookups + Gaussian-mixture
// log-densities) and the table builder that feeds it.
//
// Reference arithmetic it replaces (paths relative to the reference root):
//   emissionsSum in viterbiInitialize / viterbiCompute
//       core/src/main/scala/DynamicNaiveBayesianClassifier.scala:52-56, 79-83
//   LearnedDiscreteEdge.probability        core/src/main/scala/Edge.scala:69-75
//   LearnedContinuousEdge.probability      core/src/main/scala/Edge.scala:114-118
//   gaussianMixturePdf                     core/src/main/scala/GaussianUtils.scala:43-45
// The reference evaluates these per (state, variable, timestep) through String-keyed
// maps, boxing every value; here one launch streams the whole observation tensor.
#include "dnbc_internal.cuh"

// ---------------------------------------------------------------------------------
// table builder: fp64 master parameters -> fp32 kernel tables
// ---------------------------------------------------------------------------------
struct PrepArgs {
  Dims dm;
  const double *pi, *A, *B, *w, *mu, *var;
  const int32_t *V, *K;
  float *t_pi, *t_A, *t_At, *t_logpi, *t_logA, *t_cat, *t_gs, *t_gm, *t_gc;
};

__global__ void k_prepare_tables(PrepArgs a) {
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax, Npad = a.dm.Npad;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = tid; q < N; q += nth) {
    a.t_pi[q] = (float)a.pi[q];
    a.t_logpi[q] = (float)log(a.pi[q]);
  }
  for (int64_t q = tid; q < (int64_t)N * Npad; q += nth) {
    int i = (int)(q / Npad), j = (int)(q % Npad);
    double v = j < N ? a.A[(int64_t)i * N + j] : 0.0;
    double vt = j < N ? a.A[(int64_t)j * N + i] : 0.0;
    a.t_A[q] = (float)v;
    a.t_logA[q] = (float)log(v);
    a.t_At[q] = (float)vt;
  }
  for (int64_t q = tid; q < (int64_t)D * (Vmax + 1) * Npad; q += nth) {
    int j = (int)(q % Npad);
    int v = (int)((q / Npad) % (Vmax + 1));
    int d = (int)(q / ((int64_t)Npad * (Vmax + 1)));
    double p = (j < N && v < a.V[d]) ? a.B[((int64_t)d * N + j) * Vmax + v] : 0.0;
    a.t_cat[q] = (float)log2(p);
  }
  for (int64_t q = tid; q < (int64_t)C * Kmax * Npad; q += nth) {
    int j = (int)(q % Npad);
    int k = (int)((q / Npad) % Kmax);
    int c = (int)(q / ((int64_t)Npad * Kmax));
    float gs = 0.f, gm = 0.f, gc = -INFINITY;
    if (j < N && k < a.K[c]) {
      int64_t src = ((int64_t)c * N + j) * Kmax + k;
      double w = a.w[src], mu = a.mu[src], var = a.var[src];
      if (w > 0.0) {
        if (var > 0.0) {
          double s = sqrt(DNBC_LOG2E / (2.0 * var));
          gs = (float)s;
          gm = (float)(-mu * s);
          gc = (float)(log2(w) - 0.5 * log2(2.0 * DNBC_PI * var));
        } else {
          // degenerate covariance: spark-mllib's pseudo-inverse makes the pdf the
          // constant 1/sqrt(2 pi) (see oracle orc_mllib_pdf)
          gc = (float)(log2(w) - 0.5 * log2(2.0 * DNBC_PI));
        }
      }
    }
    a.t_gs[q] = gs;
    a.t_gm[q] = gm;
    a.t_gc[q] = gc;
  }
}

void launch_prepare_tables(dnbc_ctx *c) {
  PrepArgs a{c->dm, c->d_pi, c->d_A, c->d_B, c->d_w, c->d_mu, c->d_var, c->d_V, c->d_K,
             c->t_pi, c->t_A, c->t_At, c->t_logpi, c->t_logA, c->t_cat, c->t_gs, c->t_gm, c->t_gc};
  c->launches++;
  k_prepare_tables<<<c->sm_count, 256, 0, c->stream>>>(a);
  if (fb_tc_supported(c)) launch_tc_tables(c);   // bf16 split copies of A for the tensor-core recursion
}

// ---------------------------------------------------------------------------------
// K1 (fp32): out[n][j] = ln P(o_n | state j)
//
// Thread <-> (state j, slot g); a slot covers PPT consecutive points of the CTA's tile.
// Lanes of a warp hold consecutive states, so the parameter tables ([.][.][Npad],
// state-minor) are read with unit stride and every observation is a warp broadcast.
// Work per (point, state): D table lookups + sum_c K_c x (2 FMA + 1 MUFU.EX2 + 1 FADD
// + 1 FMNMX) + C x MUFU.LG2: the kernel is bound by the MUFU pipe, not by HBM
// (algorithmic bytes per point: D + 4C in, 4N out).
// The mixture sum is formed directly; only when it underflows fp32 (every component
// more than ~13 sigma away) is it redone shifted by the largest exponent, so results
// stay finite where the reference's fp64 sum is still finite.
// ---------------------------------------------------------------------------------
template <int PPT>
__global__ void __launch_bounds__(256) k_emission_f32(Dims dm, DeviceTables tb, DataView dv, int64_t p0,
                                                      int64_t np, float *__restrict__ out) {
  const int N = dm.N, D = dm.D, C = dm.C, Vmax = dm.Vmax, Kmax = dm.Kmax, Npad = dm.Npad;
  const int NS = Npad < (int)blockDim.x ? Npad : (int)blockDim.x;  // states spanned by the CTA at once
  const int slots = blockDim.x / NS;
  const int j0 = threadIdx.x % NS, g = threadIdx.x / NS;
  const int64_t tile = (int64_t)slots * PPT;
  const int64_t ntiles = (np + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * tile + (int64_t)g * PPT;  // first point (chunk-relative) of this slot
    for (int j = j0; j < N; j += NS) {
      float acc[PPT];
#pragma unroll
      for (int i = 0; i < PPT; ++i) acc[i] = 0.f;
      for (int d = 0; d < D; ++d) {
        const uint8_t *col = dv.disc + (int64_t)d * dv.P + p0;
        const float *tab = tb.cat + (int64_t)d * (Vmax + 1) * Npad + j;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          int64_t n = base + i;
          int v = n < np ? (int)__ldg(col + n) : 0;
          v = v < Vmax ? v : Vmax;
          acc[i] += __ldg(tab + (int64_t)v * Npad);
        }
      }
      for (int c = 0; c < C; ++c) {
        const float *xcol = dv.cont + (int64_t)c * dv.P + p0;
        const int K = __ldg(tb.K + c);
        const int64_t tofs = (int64_t)c * Kmax * Npad + j;
        float x[PPT], sum[PPT], mx[PPT];
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          int64_t n = base + i;
          x[i] = n < np ? __ldg(xcol + n) : 0.f;
          sum[i] = 0.f;
          mx[i] = -INFINITY;
        }
        for (int k = 0; k < K; ++k) {
          const float s = __ldg(tb.gs + tofs + (int64_t)k * Npad);
          const float m = __ldg(tb.gm + tofs + (int64_t)k * Npad);
          const float cc = __ldg(tb.gc + tofs + (int64_t)k * Npad);
#pragma unroll
          for (int i = 0; i < PPT; ++i) {
            float z = fmaf(x[i], s, m);
            float a = fmaf(-z, z, cc);
            mx[i] = fmaxf(mx[i], a);
            sum[i] += ex2_approx(a);
          }
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
          float L;
          if (K == 1) {
            L = mx[i];
          } else if (sum[i] >= 1e-30f) {
            L = lg2_approx(sum[i]);
          } else if (mx[i] == -INFINITY) {
            L = -INFINITY;
          } else {
            float s2 = 0.f;
            for (int k = 0; k < K; ++k) {
              const float s = __ldg(tb.gs + tofs + (int64_t)k * Npad);
              const float m = __ldg(tb.gm + tofs + (int64_t)k * Npad);
              const float cc = __ldg(tb.gc + tofs + (int64_t)k * Npad);
              float z = fmaf(x[i], s, m);
              s2 += ex2_approx(fmaf(-z, z, cc) - mx[i]);
            }
            L = mx[i] + lg2_approx(s2);
          }
          acc[i] += L;
        }
      }
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        int64_t n = base + i;
        if (n < np) out[n * N + j] = acc[i] * (float)DNBC_LN2;
      }
    }
  }
}

void launch_emission_f32(dnbc_ctx *c, int64_t p0, int64_t np, float *out) {
  if (np <= 0) return;
  if (mixture_emission_fast(c)) return launch_mixture_emission(c, p0, np, out);   // k_mixture.cu
  FamTimer ft(c, FAM_EMISSION);
  constexpr int PPT = 4;
  const int NS = c->dm.Npad < 256 ? c->dm.Npad : 256;
  const int64_t tile = (int64_t)(256 / NS) * PPT;
  int64_t ntiles = (np + tile - 1) / tile;
  int grid = (int)(ntiles < (int64_t)c->sm_count * 8 ? ntiles : (int64_t)c->sm_count * 8);
  k_emission_f32<PPT><<<grid, 256, 0, c->stream>>>(c->dm, c->tables(), c->data(), p0, np, out);
}

// ---------------------------------------------------------------------------------
// K1 (fp64 parity mode): mirrors the reference operation by operation --
// (0.0 + sum_d log P_d) + sum_c log(sum_k w_k * pdf_k(x)), pdf as spark-mllib 2.1.0
// MultivariateGaussian.pdf = exp(u + v*v*(-0.5)).  u, sqrt(1/var) and the categorical
// logs are tabulated on the HOST with glibc; explicit _rn intrinsics keep nvcc from
// contracting a*b+c into FMAs the JVM would not use.
// ---------------------------------------------------------------------------------
struct Emis64Args {
  Dims dm;
  const int32_t *K;
  const double *logB;   // [D][Vmax+1][N]
  const double *u, *rs; // [C][N][Kmax]
  const double *w, *mu; // [C][N][Kmax] master
  DataView dv;
};

__global__ void __launch_bounds__(256) k_emission_f64(Emis64Args a, int64_t p0, int64_t np, double *__restrict__ out) {
  const int N = a.dm.N, D = a.dm.D, C = a.dm.C, Vmax = a.dm.Vmax, Kmax = a.dm.Kmax;
  const int64_t total = np * N;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = q / N;
    const int j = (int)(q % N);
    double sd = 0.0, sc = 0.0;
    for (int d = 0; d < D; ++d) {
      int v = a.dv.disc[(int64_t)d * a.dv.P + p0 + n];
      v = v < Vmax ? v : Vmax;
      sd = __dadd_rn(sd, a.logB[((int64_t)d * (Vmax + 1) + v) * N + j]);
    }
    for (int c = 0; c < C; ++c) {
      const double x = a.dv.cont64 ? a.dv.cont64[(int64_t)c * a.dv.P + p0 + n]
                                   : (double)a.dv.cont[(int64_t)c * a.dv.P + p0 + n];
      const int64_t base = ((int64_t)c * N + j) * Kmax;
      double sum = 0.0;
      for (int k = 0; k < a.K[c]; ++k) {
        double v = __dmul_rn(a.rs[base + k], __dsub_rn(x, a.mu[base + k]));
        double e = __dadd_rn(a.u[base + k], __dmul_rn(__dmul_rn(v, v), -0.5));
        sum = __dadd_rn(sum, __dmul_rn(a.w[base + k], exp(e)));
      }
      sc = __dadd_rn(sc, log(sum));
    }
    double e = 0.0;
    e = __dadd_rn(e, sd);
    e = __dadd_rn(e, sc);
    out[q] = e;
  }
}

void launch_emission_f64(dnbc_ctx *c, int64_t p0, int64_t np, double *out) {
  if (np <= 0) return;
  FamTimer ft(c, FAM_EMISSION);
  Emis64Args a{c->dm, c->d_K, c->t64_logB, c->t64_u, c->t64_rs, c->d_w, c->d_mu, c->data()};
  int64_t total = np * c->dm.N;
  int grid = (int)((total + 255) / 256 < (int64_t)c->sm_count * 16 ? (total + 255) / 256 : (int64_t)c->sm_count * 16);
  k_emission_f64<<<grid, 256, 0, c->stream>>>(a, p0, np, out);
}

// E-step with clamped labels (bridge test: one-hot posteriors): E[n][j] = -inf for j != label
__global__ void k_clamp_labels(int N, const int32_t *labels, int64_t np, float *E) {
  const int64_t total = np * N;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    if (labels[q / N] != (int)(q % N)) E[q] = -INFINITY;
  }
}

void launch_clamp_labels(dnbc_ctx *c, int64_t p0, int64_t np, float *E) {
  if (np <= 0) return;
  c->launches++;
  int64_t total = np * c->dm.N;
  int grid = (int)((total + 255) / 256 < (int64_t)c->sm_count * 16 ? (total + 255) / 256 : (int64_t)c->sm_count * 16);
  k_clamp_labels<<<grid, 256, 0, c->stream>>>(c->dm.N, c->d_labels + p0, np, E);
}
