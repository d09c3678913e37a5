/*
 * dnbc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see dnbc_oracle.h).
 *
 * fp64 CPU restatement of the reference hot path.  Reference paths below are
 * relative to /root/reference/:
 *   DNBC.scala  = core/src/main/scala/DynamicNaiveBayesianClassifier.scala
 *   Edge.scala  = core/src/main/scala/Edge.scala
 *   GU.scala    = core/src/main/scala/GaussianUtils.scala
 *   Perf.scala  = core/src/main/scala/Performance.scala
 *   KMeans.java = core/src/main/java/Tools/KMeans.java
 *   EM1D.java   = core/src/main/java/Tools/ExpectationMaximization1D.java
 * Nothing here is copied: the reference is Scala/Java over String-keyed maps; this
 * is C over the dense layout described in the header.
 */
#include "dnbc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846 /* == java.lang.Math.PI */

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py's CPU arms set this explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static int64_t total_points(const orc_data *d) { return d->off[d->S]; }

/* ------------------------------------------------------------------------- */
/* Supervised learner                                                        */
/* ------------------------------------------------------------------------- */

/* One pass of the body of `originalSequences.foreach` (DNBC.scala:145-159) over
 * sequence s: categorical counts (learnDiscreteEmissions :163-178 +
 * DiscreteEdge.learn Edge.scala:46-51), transition counts (learnTransitions
 * :204-216) and the initial-state count (learnInitialEdge :218-220).
 * `known` is the key set of the `transitions` argument at sequence entry; with the
 * quirk ON the membership test at :211 looks at that immutable argument, so every
 * occurrence of a not-yet-known `from` state replaces its edge by a fresh one. */
static void learn_one_sequence(int N, int D, int Vmax, const orc_data *data, int64_t s,
                               int quirk, uint8_t *known, int64_t *cnt_pi,
                               int64_t *cnt_A, int64_t *cnt_B) {
  const int64_t P = total_points(data);
  const int64_t b = data->off[s], e = data->off[s + 1];
  if (e <= b) return;
  for (int d = 0; d < D; ++d)
    for (int64_t n = b; n < e; ++n) {
      int j = data->labels[n];
      int v = data->disc[(int64_t)d * P + n];
      if (v < Vmax) cnt_B[((int64_t)d * N + j) * Vmax + v] += 1;
    }
  for (int64_t n = b; n + 1 < e; ++n) {
    int from = data->labels[n], to = data->labels[n + 1];
    if (quirk && !known[from]) memset(cnt_A + (int64_t)from * N, 0, sizeof(int64_t) * N);
    cnt_A[(int64_t)from * N + to] += 1;
  }
  for (int64_t n = b; n + 1 < e; ++n) known[data->labels[n]] = 1;
  cnt_pi[data->labels[b]] += 1;
}

/* DNBC.scala:124-161.  ORC_DOUBLE_COUNT_FIRST reproduces `firstSequence ++ sequences`
 * (:129-133) on a re-iterable collection (the first sequence is learned twice). */
int orc_mle_counts(int N, int D, int Vmax, const orc_data *data, int flags,
                   int64_t *cnt_pi, int64_t *cnt_A, int64_t *cnt_B) {
  if (!data->labels) return -1;
  memset(cnt_pi, 0, sizeof(int64_t) * N);
  memset(cnt_A, 0, sizeof(int64_t) * N * N);
  memset(cnt_B, 0, sizeof(int64_t) * (size_t)D * N * Vmax);
  uint8_t *known = (uint8_t *)calloc(N, 1);
  const int quirk = (flags & ORC_QUIRK_TRANSITIONS) != 0;
  if ((flags & ORC_DOUBLE_COUNT_FIRST) && data->S > 0)
    learn_one_sequence(N, D, Vmax, data, 0, quirk, known, cnt_pi, cnt_A, cnt_B);
  for (int64_t s = 0; s < data->S; ++s)
    learn_one_sequence(N, D, Vmax, data, s, quirk, known, cnt_pi, cnt_A, cnt_B);
  free(known);
  return 0;
}

/* DiscreteEdge.learnFinalize, Edge.scala:53-59: P = count / sum(count).
 * Rows with no counts stay all-zero (the key is absent from the reference's map). */
void orc_normalize_counts(const int64_t *cnt, int rows, int cols, double *prob,
                          uint8_t *row_nonempty) {
  for (int r = 0; r < rows; ++r) {
    int64_t sum = 0;
    for (int c = 0; c < cols; ++c) sum += cnt[(int64_t)r * cols + c];
    for (int c = 0; c < cols; ++c)
      prob[(int64_t)r * cols + c] = sum ? (double)cnt[(int64_t)r * cols + c] / (double)sum : 0.0;
    if (row_nonempty) row_nonempty[r] = sum != 0;
  }
}

/* ContinuousEdge.learn, Edge.scala:83-85 (via learnContinuousEmissions,
 * DNBC.scala:180-202): the observations of variable c labelled `state`, in learn order. */
int64_t orc_edge_points(const orc_data *data, int c, int state, int flags, double *out,
                        int64_t cap) {
  const int64_t P = total_points(data);
  const double *x = data->cont + (int64_t)c * P;
  int64_t m = 0;
  int first_twice = (flags & ORC_DOUBLE_COUNT_FIRST) && data->S > 0;
  for (int pass = first_twice ? 0 : 1; pass < 2; ++pass) {
    int64_t b = data->off[0], e = pass == 0 ? data->off[1] : P;
    for (int64_t n = b; n < e; ++n)
      if (data->labels[n] == state) {
        if (out && m < cap) out[m] = x[n];
        ++m;
      }
  }
  return m;
}

/* ------------------------------------------------------------------------- */
/* 1-D k-means (KMeans.java)                                                 */
/* ------------------------------------------------------------------------- */

/* KMeans.run :20-37 with the random `initialize` (:46-72, unseeded java.util.Random)
 * replaced by caller-supplied centroids.  repartitionStep :83-106 uses
 * sqrt((x-c)^2) (PVector.Minus/norm2) and strict `<`, so the lowest index wins ties
 * and a NaN centroid is never selected; centroidStep :116-123 multiplies the
 * in-order sum by 1.0/size (NaN for an empty cluster).  Returns iterations run. */
int orc_kmeans_1d(const double *x, int64_t n, int k, const double *init_centroids,
                  int32_t *assign_out, double *centroids_out) {
  double *cen = (double *)malloc(sizeof(double) * k);
  int32_t *prev = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
  int64_t *size = (int64_t *)malloc(sizeof(int64_t) * k);
  memcpy(cen, init_centroids, sizeof(double) * k);
  memset(assign_out, 0, sizeof(int32_t) * n);
  int it = 0, same;
  do {
    memcpy(prev, assign_out, sizeof(int32_t) * n);
    for (int64_t i = 0; i < n; ++i) {
      int index = 0;
      double dist = DBL_MAX;
      for (int j = 0; j < k; ++j) {
        double diff = x[i] - cen[j];
        double dt = sqrt(diff * diff);
        if (dt < dist) { dist = dt; index = j; }
      }
      assign_out[i] = index;
    }
    for (int j = 0; j < k; ++j) { cen[j] = 0.0; size[j] = 0; }
    for (int64_t i = 0; i < n; ++i) { cen[assign_out[i]] += x[i]; size[assign_out[i]] += 1; }
    for (int j = 0; j < k; ++j) cen[j] = cen[j] * (1.0 / (double)size[j]);
    ++it;
    same = memcmp(prev, assign_out, sizeof(int32_t) * n) == 0;
  } while (!same && it < 30);
  if (centroids_out) memcpy(centroids_out, cen, sizeof(double) * k);
  free(cen); free(prev); free(size);
  return it;
}

/* ------------------------------------------------------------------------- */
/* 1-D GMM EM (ExpectationMaximization1D.java)                               */
/* ------------------------------------------------------------------------- */

/* EM1D.initialize :21-59: weight = |cluster|/n, mean, population variance. */
void orc_em1d_initialize(const double *x, int64_t n, int k, const int32_t *assign,
                         double *w, double *mu, double *var) {
  for (int c = 0; c < k; ++c) {
    int64_t size = 0;
    double mean = 0.0, v = 0.0;
    for (int64_t i = 0; i < n; ++i) if (assign[i] == c) { mean += x[i]; ++size; }
    mean /= (double)size;
    for (int64_t i = 0; i < n; ++i) if (assign[i] == c) v += (x[i] - mean) * (x[i] - mean);
    v /= (double)size;
    w[c] = (double)size / (double)n;
    mu[c] = mean;
    var[c] = v;
  }
}

/* jMEF/UnivariateGaussian.java:197 (source parameters (mu, sigma^2)). */
double orc_jmef_density(double x, double mu, double var) {
  return exp(-(x - mu) * (x - mu) / (2.0 * var)) / sqrt(2.0 * ORC_PI * var);
}

/* jMEF/MixtureModel.java:78-83 */
static double jmef_mixture(double x, int k, const double *w, const double *mu, const double *var) {
  double cumul = 0.0;
  for (int i = 0; i < k; ++i) cumul += w[i] * orc_jmef_density(x, mu[i], var[i]);
  return cumul;
}

/* EM1D.logLikelihood :154-159 */
static double em1d_loglik(const double *x, int64_t n, int k, const double *w, const double *mu,
                          const double *var) {
  double value = 0.0;
  for (int64_t i = 0; i < n; ++i) value += log(jmef_mixture(x[i], k, w, mu, var));
  return value;
}

/* EM1D.run :68-145.  E-step :92-101, M-step :104-132 (variance around the NEW mean,
 * weight = sum/n), log-likelihood evaluated after the M-step :136, stop rule :141
 * with the threshold taken from the INITIAL log-likelihood :81.  The reference
 * hard-codes max_iters = 30 (:13) and rel_tol = 0.01 (:81).  Returns iterations run;
 * ll_series must hold max_iters+1 doubles. */
int orc_em1d_run(const double *x, int64_t n, int k, double *w, double *mu, double *var,
                 double *ll_series, int max_iters, double rel_tol) {
  double *p = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1) * k);
  int iterations = 0;
  double ll_new = em1d_loglik(x, n, k, w, mu, var);
  const double thr = fabs(ll_new) * rel_tol;
  double ll_old;
  if (ll_series) ll_series[0] = ll_new;
  do {
    ll_old = ll_new;
    for (int64_t row = 0; row < n; ++row) {
      double sum = 0.0;
      for (int col = 0; col < k; ++col) {
        double tmp = w[col] * orc_jmef_density(x[row], mu[col], var[col]);
        p[row * k + col] = tmp;
        sum += tmp;
      }
      for (int col = 0; col < k; ++col) p[row * k + col] /= sum;
    }
    for (int col = 0; col < k; ++col) {
      double sum = 0.0, m = 0.0, sigma = 0.0;
      for (int64_t row = 0; row < n; ++row) {
        double ww = p[row * k + col];
        sum += ww;
        m += x[row] * ww;
      }
      m /= sum;
      for (int64_t row = 0; row < n; ++row) {
        double diff = x[row] - m;
        sigma += p[row * k + col] * diff * diff;
      }
      sigma /= sum;
      mu[col] = m;
      var[col] = sigma;
      w[col] = sum / (double)n;
    }
    ++iterations;
    ll_new = em1d_loglik(x, n, k, w, mu, var);
    if (ll_series) ll_series[iterations] = ll_new;
  } while (fabs(ll_new - ll_old) > thr && iterations < max_iters);
  free(p);
  return iterations;
}

/* ------------------------------------------------------------------------- */
/* Decode-side densities                                                     */
/* ------------------------------------------------------------------------- */

/* org.apache.spark:spark-mllib_2.11:2.1.0 MultivariateGaussian.pdf for the 1x1 case
 * the reference constructs at Edge.scala:98-101 and evaluates at GU.scala:44.  The
 * dependency is NOT vendored under /root/reference; restated from its published
 * algorithm: pdf = exp(u + v*v*(-0.5)), u = -0.5*(d*ln(2 pi) + ln pdet(Sigma)),
 * v = sqrt(1/lambda) * (x - mu); an eigenvalue not above EPSILON*max(lambda)*d is
 * treated as zero (pseudo-determinant/pseudo-inverse), which for 1x1 means
 * var <= 0 or NaN gives the constant exp(-0.5*ln(2 pi)). */
double orc_mllib_pdf(double x, double mu, double var) {
  if (!(var > 0.0)) return exp(-0.5 * (log(2.0 * ORC_PI) + 0.0));
  double u = -0.5 * (log(2.0 * ORC_PI) + log(var));
  double v = sqrt(1.0 / var) * (x - mu);
  return exp(u + v * v * -0.5);
}

/* LearnedContinuousEdge.probability Edge.scala:114-118 -> gaussianMixturePdf GU.scala:43-45 */
double orc_mixture_pdf(const orc_model *m, int c, int j, double x) {
  const int64_t base = ((int64_t)c * m->N + j) * m->Kmax;
  double sum = 0.0;
  for (int k = 0; k < m->K[c]; ++k)
    sum += m->w[base + k] * orc_mllib_pdf(x, m->mu[base + k], m->var[base + k]);
  return sum;
}

/* LearnedDiscreteEdge.probability Edge.scala:69-75: 0 for an unseen symbol. */
static double cat_prob(const orc_model *m, int d, int j, int v) {
  if (v >= m->V[d]) return 0.0;
  return m->B[((int64_t)d * m->N + j) * m->Vmax + v];
}

/* emissionsSum of DNBC.scala:52-56 / :79-83: (0.0 + sum_d log P_d) + sum_c log gmm_c,
 * each `.sum` folding from 0.0 in variable order. */
static double emission_one(const orc_model *m, const orc_data *data, int64_t P, int64_t n, int j) {
  double sd = 0.0, sc = 0.0, e = 0.0;
  for (int d = 0; d < m->D; ++d) sd += log(cat_prob(m, d, j, data->disc[(int64_t)d * P + n]));
  for (int c = 0; c < m->C; ++c) sc += log(orc_mixture_pdf(m, c, j, data->cont[(int64_t)c * P + n]));
  e += sd;
  e += sc;
  return e;
}

void orc_emission_loglik(const orc_model *m, const orc_data *data, double *out) {
  const int64_t P = total_points(data);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < P; ++n)
    for (int j = 0; j < m->N; ++j) out[n * m->N + j] = emission_one(m, data, P, n, j);
}

/* ------------------------------------------------------------------------- */
/* Decode                                                                    */
/* ------------------------------------------------------------------------- */

/* Scala 2.11 TraversableOnce.maxBy with Ordering.Double: the first element is taken
 * unconditionally, later ones replace it only when strictly greater (`x > y`), so the
 * first maximum wins and a NaN never displaces anything.  `idx` lists the candidate
 * states in iteration order.  *tie is set when a later value equals the maximum. */
static int first_max(const double *val, const int *idx, int cnt, int *tie) {
  int best = idx[0];
  double bv = val[0];
  int t = 0;
  for (int q = 1; q < cnt; ++q) {
    if (val[q] > bv) { bv = val[q]; best = idx[q]; t = 0; }
    else if (val[q] == bv) t = 1;
  }
  if (tie) *tie |= t;
  return best;
}

/* One sequence.  REFERENCE mode follows viterbiInitialize DNBC.scala:49-60 and
 * viterbiCompute :67-91 including both quirks: the reported path is the
 * per-step argmax of delta (:75, :89), and delta_t(j) = e_t(j) + delta_{t-1}(i*) where
 * i* maximises delta_{t-1}(i) + ln a(i->j) but ln a is NOT added (`._2` at :85);
 * score = sum_j delta_{T-1}(j), 0.0 when T == 1 (:90, vcur stays empty).
 * BACKTRACE mode is the textbook Viterbi (adds ln a, back-pointers, score = max). */
static void decode_one(const orc_model *m, const orc_data *data, int64_t s, int mode,
                       const int *act, int na, double *lna, int32_t *path, double *score,
                       uint8_t *tie_out, double *work, int32_t *bp) {
  const int N = m->N;
  const int64_t P = total_points(data);
  const int64_t b = data->off[s], T = data->off[s + 1] - b;
  double *vprev = work, *vcur = work + N, *cand = work + 2 * N, *tmp = work + 3 * N;
  if (T <= 0) { if (score) score[s] = 0.0; return; }
  for (int q = 0; q < na; ++q) {
    int j = act[q];
    vprev[j] = emission_one(m, data, P, b, j) + log(m->pi[j]);
  }
  for (int64_t t = 1; t < T; ++t) {
    int tie0 = 0, tie1 = 0;
    if (mode == ORC_DECODE_REFERENCE) {
      for (int q = 0; q < na; ++q) tmp[q] = vprev[act[q]];
      path[b + t - 1] = first_max(tmp, act, na, &tie0);
      if (tie_out) tie_out[b + t - 1] |= (uint8_t)tie0;
    }
    for (int q = 0; q < na; ++q) {
      int j = act[q];
      for (int r = 0; r < na; ++r) cand[r] = vprev[act[r]] + lna[(int64_t)act[r] * N + j];
      int istar = first_max(cand, act, na, &tie1);
      double e = emission_one(m, data, P, b + t, j);
      if (mode == ORC_DECODE_REFERENCE) vcur[j] = e + vprev[istar];
      else { vcur[j] = e + (vprev[istar] + lna[(int64_t)istar * N + j]); bp[t * N + j] = istar; }
    }
    if (tie_out) tie_out[b + t] |= (uint8_t)(tie1 << 1);
    double *sw = vprev; vprev = vcur; vcur = sw;
  }
  {
    int tie0 = 0;
    for (int q = 0; q < na; ++q) tmp[q] = vprev[act[q]];
    int last = first_max(tmp, act, na, &tie0);
    path[b + T - 1] = last;
    if (tie_out) tie_out[b + T - 1] |= (uint8_t)tie0;
    if (mode == ORC_DECODE_REFERENCE) {
      double sc = 0.0;
      if (T > 1) for (int q = 0; q < na; ++q) sc += vprev[act[q]];
      if (score) score[s] = sc;
    } else {
      if (score) score[s] = vprev[last];
      for (int64_t t = T - 1; t >= 1; --t) path[b + t - 1] = bp[t * N + path[b + t]];
    }
  }
}

void orc_decode(const orc_model *m, const orc_data *data, int mode, int32_t *path_out,
                double *score_out, uint8_t *tie_out) {
  const int N = m->N;
  const int64_t P = total_points(data);
  int *act = (int *)malloc(sizeof(int) * N);
  int na = 0;
  for (int j = 0; j < N; ++j) if (!m->active || m->active[j]) act[na++] = j;
  double *lna = (double *)malloc(sizeof(double) * N * N);
  /* Math.log(transitions(hs).probability(hiddenState)) DNBC.scala:85; absent -> log 0 */
  for (int64_t q = 0; q < (int64_t)N * N; ++q) lna[q] = log(m->A[q]);
  if (tie_out) memset(tie_out, 0, P);
  int64_t Tmax = 0;
  for (int64_t s = 0; s < data->S; ++s)
    if (data->off[s + 1] - data->off[s] > Tmax) Tmax = data->off[s + 1] - data->off[s];
  if (na > 0) {
#pragma omp parallel
    {
      double *work = (double *)malloc(sizeof(double) * 4 * N);
      int32_t *bp = mode == ORC_DECODE_BACKTRACE
                        ? (int32_t *)malloc(sizeof(int32_t) * (size_t)(Tmax ? Tmax : 1) * N) : NULL;
#pragma omp for schedule(dynamic, 4)
      for (int64_t s = 0; s < data->S; ++s)
        decode_one(m, data, s, mode, act, na, lna, path_out, score_out, tie_out, work, bp);
      free(work); free(bp);
    }
  }
  free(act); free(lna);
}

/* Perf.scala:45-55 */
double orc_accuracy(const orc_data *data, const int32_t *path) {
  double acc = 0.0;
  for (int64_t s = 0; s < data->S; ++s) {
    int64_t b = data->off[s], e = data->off[s + 1], ok = 0;
    for (int64_t n = b; n < e; ++n) ok += data->labels[n] == path[n];
    acc += (double)ok / (double)(e - b);
  }
  return acc / (double)data->S * 100.0;
}

/* ------------------------------------------------------------------------- */
/* Baum-Welch (NEW: the reference has none)           */
/* ------------------------------------------------------------------------- */

int64_t orc_stats_size(const orc_model *m) {
  return 1 + (int64_t)m->N + (int64_t)m->N * m->N + m->N + (int64_t)m->D * m->N * m->Vmax +
         3 * (int64_t)m->C * m->N * m->Kmax;
}

/* log of component k of mixture (c, j) at x, natural log, including the weight */
static double log_component(const orc_model *m, int64_t base, int k, double x) {
  double var = m->var[base + k], d = x - m->mu[base + k];
  return log(m->w[base + k]) - 0.5 * log(2.0 * ORC_PI * var) - d * d / (2.0 * var);
}

/* robust (log-sum-exp) log emission used by the E-step; equals the reference's
 * emissionsSum wherever that does not underflow to log(0). */
static double em_log_emission(const orc_model *m, const orc_data *data, int64_t P, int64_t n, int j) {
  double e = 0.0;
  for (int d = 0; d < m->D; ++d) e += log(cat_prob(m, d, j, data->disc[(int64_t)d * P + n]));
  for (int c = 0; c < m->C; ++c) {
    const int64_t base = ((int64_t)c * m->N + j) * m->Kmax;
    double x = data->cont[(int64_t)c * P + n], mx = -INFINITY, lc[64];
    int K = m->K[c];
    for (int k = 0; k < K; ++k) { lc[k] = log_component(m, base, k, x); if (lc[k] > mx) mx = lc[k]; }
    if (mx == -INFINITY) { e += -INFINITY; continue; }
    double sum = 0.0;
    for (int k = 0; k < K; ++k) sum += exp(lc[k] - mx);
    e += mx + log(sum);
  }
  return e;
}

typedef struct { double *le, *alpha, *beta, *cs; int64_t cap; } fb_work;

static void fb_reserve(fb_work *w, int64_t T, int N) {
  if (T <= w->cap) return;
  free(w->le); free(w->alpha); free(w->beta); free(w->cs);
  w->le = (double *)malloc(sizeof(double) * T * N);
  w->alpha = (double *)malloc(sizeof(double) * T * N);
  w->beta = (double *)malloc(sizeof(double) * 2 * N);
  w->cs = (double *)malloc(sizeof(double) * T);
  w->cap = T;
}

/* Scaled forward-backward for one sequence; accumulates into st (layout in header)
 * and/or writes gamma.  b~_t(j) = exp(le_t(j) - max_j le_t(j)). */
static void fb_one(const orc_model *m, const orc_data *data, int64_t s, int clamp, fb_work *wk,
                   double *st, double *gamma_out) {
  const int N = m->N, D = m->D, C = m->C, Vmax = m->Vmax, Kmax = m->Kmax;
  const int64_t P = total_points(data);
  const int64_t b = data->off[s], T = data->off[s + 1] - b;
  if (T <= 0) return;
  fb_reserve(wk, T, N);
  double *le = wk->le, *al = wk->alpha, *cs = wk->cs;
  double ll = 0.0;
  /* emissions -> scaled probabilities in place */
  for (int64_t t = 0; t < T; ++t) {
    double mx = -INFINITY;
    for (int j = 0; j < N; ++j) {
      double v = em_log_emission(m, data, P, b + t, j);
      if (clamp && data->labels && data->labels[b + t] != j) v = -INFINITY;
      le[t * N + j] = v;
      if (v > mx) mx = v;
    }
    for (int j = 0; j < N; ++j) le[t * N + j] = exp(le[t * N + j] - mx);
    ll += mx;
  }
  /* forward */
  {
    double c0 = 0.0;
    for (int j = 0; j < N; ++j) { al[j] = m->pi[j] * le[j]; c0 += al[j]; }
    for (int j = 0; j < N; ++j) al[j] /= c0;
    cs[0] = c0; ll += log(c0);
  }
  for (int64_t t = 1; t < T; ++t) {
    double c = 0.0;
    for (int j = 0; j < N; ++j) {
      double a = 0.0;
      for (int i = 0; i < N; ++i) a += al[(t - 1) * N + i] * m->A[(int64_t)i * N + j];
      a *= le[t * N + j];
      al[t * N + j] = a; c += a;
    }
    for (int j = 0; j < N; ++j) al[t * N + j] /= c;
    cs[t] = c; ll += log(c);
  }
  if (st) st[0] += ll;
  double *S_pi = st ? st + 1 : NULL, *S_xi = st ? S_pi + N : NULL, *S_g = st ? S_xi + (int64_t)N * N : NULL;
  double *S_cat = st ? S_g + N : NULL, *S_gmm = st ? S_cat + (int64_t)D * N * Vmax : NULL;
  /* backward + statistics */
  double *bt = wk->beta, *bn = wk->beta + N;
  for (int j = 0; j < N; ++j) bt[j] = 1.0;
  for (int64_t t = T - 1; t >= 0; --t) {
    if (t < T - 1) {
      /* v_j = b~_{t+1}(j) beta_{t+1}(j) / c_{t+1};  beta_t(i) = sum_j a_ij v_j */
      for (int j = 0; j < N; ++j) bn[j] = le[(t + 1) * N + j] * bt[j] / cs[t + 1];
      if (st)
        for (int i = 0; i < N; ++i) {
          double ai = al[t * N + i];
          if (ai != 0.0)
            for (int j = 0; j < N; ++j) S_xi[(int64_t)i * N + j] += ai * m->A[(int64_t)i * N + j] * bn[j];
        }
      for (int i = 0; i < N; ++i) {
        double acc = 0.0;
        for (int j = 0; j < N; ++j) acc += m->A[(int64_t)i * N + j] * bn[j];
        bt[i] = acc;
      }
    }
    const int64_t n = b + t;
    for (int j = 0; j < N; ++j) {
      double g = al[t * N + j] * bt[j];
      if (gamma_out) gamma_out[n * N + j] = g;
      if (!st || g == 0.0) continue;
      if (t == 0) S_pi[j] += g;
      S_g[j] += g;
      for (int d = 0; d < D; ++d) {
        int v = data->disc[(int64_t)d * P + n];
        if (v < Vmax) S_cat[((int64_t)d * N + j) * Vmax + v] += g;
      }
      for (int c = 0; c < C; ++c) {
        const int64_t base = ((int64_t)c * N + j) * Kmax;
        double x = data->cont[(int64_t)c * P + n], mx = -INFINITY, lc[64], sum = 0.0;
        int K = m->K[c];
        for (int k = 0; k < K; ++k) { lc[k] = log_component(m, base, k, x); if (lc[k] > mx) mx = lc[k]; }
        if (mx == -INFINITY) continue;
        for (int k = 0; k < K; ++k) { lc[k] = exp(lc[k] - mx); sum += lc[k]; }
        for (int k = 0; k < K; ++k) {
          double r = g * lc[k] / sum, dx = x - m->mu[base + k];
          S_gmm[3 * (base + k) + 0] += r;
          S_gmm[3 * (base + k) + 1] += r * dx;
          S_gmm[3 * (base + k) + 2] += r * dx * dx;
        }
      }
    }
  }
}

void orc_estep(const orc_model *m, const orc_data *data, int clamp_labels, double *stats) {
  const int64_t sz = orc_stats_size(m);
  memset(stats, 0, sizeof(double) * sz);
  int nt = orc_num_threads();
  double *part = (double *)calloc((size_t)nt * sz, sizeof(double));
#pragma omp parallel
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    fb_work wk = {0, 0, 0, 0, 0};
#pragma omp for schedule(static)
    for (int64_t s = 0; s < data->S; ++s) fb_one(m, data, s, clamp_labels, &wk, part + (int64_t)tid * sz, NULL);
    free(wk.le); free(wk.alpha); free(wk.beta); free(wk.cs);
  }
  for (int t = 0; t < nt; ++t)
    for (int64_t q = 0; q < sz; ++q) stats[q] += part[(int64_t)t * sz + q];
  free(part);
}

void orc_posteriors(const orc_model *m, const orc_data *data, double *gamma_out) {
#pragma omp parallel
  {
    fb_work wk = {0, 0, 0, 0, 0};
#pragma omp for schedule(static)
    for (int64_t s = 0; s < data->S; ++s) fb_one(m, data, s, 0, &wk, NULL, gamma_out);
    free(wk.le); free(wk.alpha); free(wk.beta); free(wk.cs);
  }
}

/* M-step.  With one-hot posteriors it reduces term by term to
 * DiscreteEdge.learnFinalize Edge.scala:53-59 and the EM1D M-step EM1D.java:104-132
 * (variance around the new mean: s2/s0 - (s1/s0)^2 for moments centred on the old
 * mean).  A row/edge whose denominator is zero keeps its previous parameters. */
/* M-step with the product's variance floor: a new variance below floor_rel x (pooled variance of
 * that observed variable, from the same statistics) is raised to it; *floored counts them.  The
 * reference's EM1D has no floor (EM1D.java:120-124); floor_rel = 0 restates it exactly. */
void orc_mstep_floor(orc_model *m, const double *st, int64_t S, double floor_rel, int64_t *floored) {
  const int N = m->N, D = m->D, C = m->C, Vmax = m->Vmax, Kmax = m->Kmax;
  const double *S_pi = st + 1, *S_xi = S_pi + N, *S_g = S_xi + (int64_t)N * N;
  const double *S_cat = S_g + N, *S_gmm = S_cat + (int64_t)D * N * Vmax;
  int64_t nfl = 0;
  for (int j = 0; j < N; ++j) m->pi[j] = S_pi[j] / (double)S;
  for (int i = 0; i < N; ++i) {
    double sum = 0.0;
    for (int j = 0; j < N; ++j) sum += S_xi[(int64_t)i * N + j];
    if (sum > 0.0) for (int j = 0; j < N; ++j) m->A[(int64_t)i * N + j] = S_xi[(int64_t)i * N + j] / sum;
  }
  for (int d = 0; d < D; ++d)
    for (int j = 0; j < N; ++j) {
      const double *row = S_cat + ((int64_t)d * N + j) * Vmax;
      double sum = 0.0;
      for (int v = 0; v < m->V[d]; ++v) sum += row[v];
      if (sum > 0.0) for (int v = 0; v < m->V[d]; ++v) m->B[((int64_t)d * N + j) * Vmax + v] = row[v] / sum;
    }
  for (int c = 0; c < C; ++c) {
    /* pooled variance of variable c: x = z + mu_old */
    double n = 0.0, sx = 0.0, sxx = 0.0, pooled = 0.0;
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < m->K[c]; ++k) {
        const int64_t e = ((int64_t)c * N + j) * Kmax + k;
        const double mu = m->mu[e], s0 = S_gmm[3 * e], s1 = S_gmm[3 * e + 1], s2 = S_gmm[3 * e + 2];
        n += s0; sx += s1 + mu * s0; sxx += s2 + 2.0 * mu * s1 + mu * mu * s0;
      }
    if (n > 0.0) { const double mean = sx / n; pooled = sxx / n - mean * mean; }
    if (!(pooled > 0.0)) pooled = 0.0;
    const double vfloor = floor_rel * pooled;
    for (int j = 0; j < N; ++j) {
      const int64_t base = ((int64_t)c * N + j) * Kmax;
      double tot = 0.0;
      for (int k = 0; k < m->K[c]; ++k) tot += S_gmm[3 * (base + k)];
      if (!(tot > 0.0)) continue;
      for (int k = 0; k < m->K[c]; ++k) {
        double s0 = S_gmm[3 * (base + k)], s1 = S_gmm[3 * (base + k) + 1], s2 = S_gmm[3 * (base + k) + 2];
        m->w[base + k] = s0 / tot;
        if (s0 > 0.0) {
          double sh = s1 / s0;
          double v = s2 / s0 - sh * sh;
          if (!(v >= vfloor)) { v = vfloor; ++nfl; }
          m->mu[base + k] += sh;
          m->var[base + k] = v;
        }
      }
    }
  }
  if (floored) *floored = nfl;
}

void orc_mstep(orc_model *m, const double *st, int64_t S) { orc_mstep_floor(m, st, S, 1e-10, 0); }
