/*
 * dnbc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64 CPU restatement of the lucivpav/dnbc-scala learning / decoding hot path,
 * written from the reference's behaviour (file:line citations on every function
 * in dnbc_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (dnbc_scala_b200/, include/dnbc_b200.h) never links, imports or calls it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - supervised mle, 1-D k-means, 1-D GMM EM, emission densities, REFERENCE-mode
 *     decode: pinned against the reference's four data fixtures via the accuracy
 *     figures the reference publishes (README.md:9-15) and asserts
 *     (DynamicNaiveBayesianClassifierTest.scala:13,19,25,33).
 *   - forward-backward / Baum-Welch: PARITY UNPINNED -- the reference contains no
 *     Baum-Welch; tied to the reference only through the clamped-label
 *     bridge test (tests/test_oracle_bridge.py).
 *   - spark-mllib 2.1.0 MultivariateGaussian.pdf is an un-vendored dependency; its
 *     1x1 case is restated from its published formula (orc_mllib_pdf).
 *
 * Dense layout shared with the product C-ABI (include/dnbc_b200.h):
 *   pi   [N]                  initial probabilities
 *   A    [N][N]               A[i*N+j] = P(i -> j)
 *   B    [D][N][Vmax]         categorical emission P(symbol v | state j) of variable d
 *   w,mu,var [C][N][Kmax]     1-D Gaussian mixture per (continuous variable, state);
 *                             components k >= K[c] are padding (w = 0)
 *   active [N]                1 = state is in the decode state set
 *                             (reference: `transitions.keys`, DNBC.scala:51,77)
 *   observations: disc u8 [D][P], cont f64 [C][P], labels i32 [P], P = sum of lengths,
 *   point index n = off[s] + t; symbol id >= V[d] means "never seen" (probability 0,
 *   Edge.scala:70-73).
 */
#ifndef DNBC_ORACLE_H
#define DNBC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t N, D, C, Vmax, Kmax;
  const int32_t *V;        /* [D] alphabet size per discrete variable   */
  const int32_t *K;        /* [C] mixture size per continuous variable  */
  double *pi;              /* [N]                                       */
  double *A;               /* [N*N]                                     */
  double *B;               /* [D*N*Vmax]                                */
  double *w, *mu, *var;    /* [C*N*Kmax]                                */
  uint8_t *active;         /* [N]                                       */
} orc_model;

typedef struct {
  int64_t S;               /* number of sequences                        */
  const int64_t *off;      /* [S+1] point offsets                        */
  const uint8_t *disc;     /* [D][P]                                     */
  const double *cont;      /* [C][P]                                     */
  const int32_t *labels;   /* [P] or NULL                                */
} orc_data;

enum { ORC_DECODE_REFERENCE = 0, ORC_DECODE_BACKTRACE = 1 };
enum { ORC_QUIRK_TRANSITIONS = 1, ORC_DOUBLE_COUNT_FIRST = 2 };

/* ---- supervised learner (DNBC.scala:124-220, Edge.scala:45-63) ---- */
int orc_mle_counts(int N, int D, int Vmax, const orc_data *data, int flags,
                   int64_t *cnt_pi, int64_t *cnt_A, int64_t *cnt_B);
void orc_normalize_counts(const int64_t *cnt, int rows, int cols, double *prob,
                          uint8_t *row_nonempty);
/* gathers the points of edge (c, state) in learn order; returns count */
int64_t orc_edge_points(const orc_data *data, int c, int state, int flags,
                        double *out, int64_t cap);

/* ---- KMeans.java:20-123 with injectable initial centroids ---- */
int orc_kmeans_1d(const double *x, int64_t n, int k, const double *init_centroids,
                  int32_t *assign_out, double *centroids_out);
/* ---- EM1D.java:21-59 ---- */
void orc_em1d_initialize(const double *x, int64_t n, int k, const int32_t *assign,
                         double *w, double *mu, double *var);
/* ---- EM1D.java:68-145; ll_series[0] = initial LL, [i] = LL after iteration i ---- */
int orc_em1d_run(const double *x, int64_t n, int k, double *w, double *mu, double *var,
                 double *ll_series, int max_iters, double rel_tol);

/* ---- densities ---- */
double orc_jmef_density(double x, double mu, double var);   /* jMEF/UnivariateGaussian.java:197 */
double orc_mllib_pdf(double x, double mu, double var);      /* spark-mllib 2.1.0, 1x1 case      */
double orc_mixture_pdf(const orc_model *m, int c, int j, double x); /* GU.scala:43-45 */

/* ---- emission log-probabilities, natural log: out[n*N + j] (DNBC.scala:52-56,80-83) ---- */
void orc_emission_loglik(const orc_model *m, const orc_data *data, double *out);

/* ---- decode (DNBC.scala:24-91 REFERENCE; true Viterbi BACKTRACE) ----
 * path_out [P] int32, score_out [S] double, tie_out [P] uint8 (may be NULL):
 * tie bit0 = exact tie in the path argmax at this step, bit1 = exact tie in some
 * predecessor argmax of this step. */
void orc_decode(const orc_model *m, const orc_data *data, int mode, int32_t *path_out,
                double *score_out, uint8_t *tie_out);

/* ---- Baum-Welch (no reference counterpart) ----
 * stats layout (doubles), total orc_stats_size():
 *   [0]                     log-likelihood
 *   [1 .. 1+N)              pi-stat    sum_s gamma_{s,0}(j)
 *   [.. +N*N)               xi-stat    sum_t xi_t(i,j)
 *   [.. +N)                 gamma-sum  sum_t gamma_t(j)
 *   [.. +D*N*Vmax)          categorical sum_{t:o=v} gamma_t(j)
 *   [.. +3*C*N*Kmax)        GMM: s0 = sum r, s1 = sum r (x-mu_old), s2 = sum r (x-mu_old)^2
 */
int64_t orc_stats_size(const orc_model *m);
void orc_estep(const orc_model *m, const orc_data *data, int clamp_labels, double *stats);
void orc_mstep(orc_model *m, const double *stats, int64_t S);   /* floor_rel = 1e-10, the product's default */
void orc_mstep_floor(orc_model *m, const double *stats, int64_t S, double floor_rel, int64_t *floored);
/* gamma_out [P][N] optional posterior dump for tests */
void orc_posteriors(const orc_model *m, const orc_data *data, double *gamma_out);

/* Perf.scala:45-55: mean over sequences of per-sequence fraction correct, x100 */
double orc_accuracy(const orc_data *data, const int32_t *path);

int orc_num_threads(void);
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
