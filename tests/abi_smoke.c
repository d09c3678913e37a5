/*
 * abi_smoke.c -- a plain C11 consumer of include/dnbc_b200.h (test program, run by
 * tests/test_gpu_baseline_shapes.py::test_c_abi_smoke_program on a B200).
 *
 * Walks the entry points in the order the JNI glue (jni/dnbc_b200_jni.c) calls them for the
 * reference's API: DynamicNaiveBayesianClassifier.mle (...scala:124-161) = create -> upload ->
 * mle -> get_params; inferMostLikelyHiddenStates / score (:24-39) = upload -> decode; then the
 * additive Baum-Welch entry points and the library-owned communicator.  No Python, no torch:
 * the closest stand-in for the JVM call path available in an image without a JDK.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dnbc_b200.h"

#define CHECK(call)                                                                            \
  do {                                                                                         \
    int rc__ = (call);                                                                         \
    if (rc__ != DNBC_OK) {                                                                     \
      fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc__, dnbc_last_error(ctx)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define REQUIRE(cond)                                                      \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "%s:%d requirement failed: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

static uint64_t lcg_state = 88172645463325252ull;
static double uniform(void) {
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(lcg_state >> 11) / 9007199254740992.0;
}
static double gauss(void) { return sqrt(-2.0 * log(uniform() + 1e-300)) * cos(6.283185307179586 * uniform()); }

int main(void) {
  dnbc_ctx *ctx = NULL;
  enum { N = 3, D = 1, C = 1, V = 4, K = 2, S = 40, T = 50, P = S * T };
  const int32_t alphabet[D] = {V}, mixture[C] = {K};

  /* error convention: the reference's message for a hint < 1 (DynamicNaiveBayesianClassifier.scala:276) */
  const int32_t bad[C] = {0};
  dnbc_config cfg = {N, D, C, alphabet, bad, 0, 0};
  REQUIRE(dnbc_create(&cfg, &ctx) == DNBC_EINVAL && ctx == NULL);
  REQUIRE(strstr(dnbc_last_error(NULL), "at least one") != NULL);
  cfg.mixture = mixture;
  CHECK(dnbc_create(&cfg, &ctx));
  REQUIRE(dnbc_abi_version() == DNBC_B200_ABI_VERSION);

  /* a small labelled data set: a 3-state chain, one symbol variable, one two-component mixture */
  static int64_t off[S + 1];
  static uint8_t disc[D * P];
  static double cont[C * P];
  static int32_t labels[P];
  const double mean[N][K] = {{-6, -2}, {0, 3}, {7, 11}};
  for (int s = 0; s <= S; ++s) off[s] = (int64_t)s * T;
  for (int s = 0; s < S; ++s) {
    int h = (int)(uniform() * N);
    for (int t = 0; t < T; ++t) {
      if (t) h = uniform() < 0.8 ? h : (h + 1 + (int)(uniform() * (N - 1))) % N;
      const int n = s * T + t;
      labels[n] = h;
      disc[n] = (uint8_t)(uniform() < 0.7 ? h : (int)(uniform() * V));
      cont[n] = mean[h][uniform() < 0.5] + gauss();
    }
  }

  /* mle: upload (labels) -> dnbc_mle -> read the learned model back */
  CHECK(dnbc_upload(ctx, S, off, disc, cont, NULL, labels));
  int64_t n_seq = 0, n_pts = 0;
  CHECK(dnbc_data_shape(ctx, &n_seq, &n_pts));
  REQUIRE(n_seq == S && n_pts == P);
  static double ll_series[C * N * 31];
  static int32_t n_iters[C * N];
  CHECK(dnbc_mle(ctx, 0, NULL, 30, 0.01, ll_series, n_iters));
  double pi[N], A[N * N], B[D * N * V], w[C * N * K], mu[C * N * K], var[C * N * K];
  uint8_t active[N];
  CHECK(dnbc_get_params(ctx, pi, A, B, w, mu, var, active));
  double spi = 0;
  for (int j = 0; j < N; ++j) {
    spi += pi[j];
    double sa = 0, sb = 0, sw = 0;
    for (int k = 0; k < N; ++k) sa += A[j * N + k];
    for (int v = 0; v < V; ++v) sb += B[j * V + v];
    for (int k = 0; k < K; ++k) { sw += w[j * K + k]; REQUIRE(var[j * K + k] > 0); }
    REQUIRE(fabs(sa - 1) < 1e-12 && fabs(sb - 1) < 1e-12 && fabs(sw - 1) < 1e-9 && active[j] == 1);
    REQUIRE(A[j * N + j] > 0.6);                     /* the chain stays put 80 % of the time */
  }
  REQUIRE(fabs(spi - 1) < 1e-12);

  /* infer: decode the same sequences in the reference's mode and as a true Viterbi */
  static uint16_t path[P];
  static double score[S];
  static uint8_t tie[P];
  CHECK(dnbc_decode(ctx, DNBC_DECODE_REFERENCE, DNBC_PREC_F64, path, score, tie));
  int hit = 0;
  for (int n = 0; n < P; ++n) hit += path[n] == labels[n];
  REQUIRE(hit > P * 0.85);
  for (int s = 0; s < S; ++s) REQUIRE(isfinite(score[s]) && score[s] < 0);
  CHECK(dnbc_decode(ctx, DNBC_DECODE_BACKTRACE, DNBC_PREC_F32, path, score, NULL));
  hit = 0;
  for (int n = 0; n < P; ++n) hit += path[n] == labels[n];
  REQUIRE(hit > P * 0.9);

  /* Baum-Welch from the supervised model: the log-likelihood must not decrease */
  double ll[4];
  CHECK(dnbc_em_iterate(ctx, 4, ll));
  for (int i = 1; i < 4; ++i) REQUIRE(ll[i] >= ll[i - 1] - 1e-6 * fabs(ll[0]));

  /* the communicator the library owns: one rank here, so the all-reduce must be the identity */
  uint8_t id[DNBC_COMM_ID_BYTES];
  CHECK(dnbc_comm_unique_id(id));
  CHECK(dnbc_comm_init_rank(ctx, 1, 0, id));
  int nr = 0, rk = -1;
  CHECK(dnbc_comm_info(ctx, &nr, &rk));
  REQUIRE(nr == 1 && rk == 0);
  const int64_t nst = dnbc_stats_size(ctx);
  REQUIRE(nst == 1 + N + N * N + N + D * N * V + 3 * C * N * K);
  double *st0 = malloc(sizeof(double) * nst), *st1 = malloc(sizeof(double) * nst);
  CHECK(dnbc_em_estep(ctx, 0));
  CHECK(dnbc_em_get_stats(ctx, st0));
  CHECK(dnbc_em_allreduce(ctx));
  CHECK(dnbc_em_get_stats(ctx, st1));
  REQUIRE(memcmp(st0, st1, sizeof(double) * nst) == 0);
  REQUIRE(st0[0] >= ll[3] - 1e-6 * fabs(ll[0]));
  double ll2[2];
  CHECK(dnbc_em_iterate_sharded(ctx, 2, S, ll2));
  REQUIRE(fabs(ll2[0] - st0[0]) <= 1e-9 * fabs(st0[0]) && ll2[1] >= ll2[0] - 1e-6 * fabs(ll2[0]));
  CHECK(dnbc_comm_destroy(ctx));
  free(st0);
  free(st1);

  REQUIRE(dnbc_launch_count(ctx) > 20);
  printf("abi_smoke ok: %lld kernel launches, loglik %.6f -> %.6f\n", (long long)dnbc_launch_count(ctx), ll[0], ll2[1]);
  CHECK(dnbc_destroy(ctx));
  return 0;
}
