/*
 * dnbc_b200.h -- C ABI of the B200-native DNBC learn / decode hot path.
 *
 * This is the drop-in boundary under the reference's L2 API (lucivpav/dnbc-scala,
 * paths relative to the reference root):
 *   object DynamicNaiveBayesianClassifier.mle(sc, sequences, hints)
 *       core/src/main/scala/DynamicNaiveBayesianClassifier.scala:124-161
 *   inferMostLikelyHiddenStates(observedStates)            ...scala:24-28
 *   score(observedStates)                                  ...scala:35-39
 * The reference has no native boundary of its own (pure JVM); the JVM shim that a
 * maintainer adds keeps the String<->index dictionaries, flattens
 * `Iterable[Seq[State]]` into the dense arrays below, calls these entry points over
 * JNI, and rebuilds the four learned maps from dnbc_get_params.
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions across the boundary;
 *   - every call returns 0 on success, a DNBC_E* code otherwise; the message is at
 *     dnbc_last_error(ctx) (ctx may be NULL for dnbc_create failures);
 *   - the caller owns every host buffer; the library copies on upload and owns all
 *     device memory inside the context; outputs go into caller-provided buffers;
 *   - one host thread per context at a time; contexts are independent;
 *   - there is NO CPU fallback: without a CUDA device dnbc_create fails.
 *
 * Dense layouts (Vmax = max alphabet, Kmax = max mixture size, both fixed at create):
 *   pi   f64 [N]               A    f64 [N][N]  (row = from-state)
 *   B    f64 [D][N][Vmax]      w, mu, var  f64 [C][N][Kmax]  (k >= K[c] is padding)
 *   active u8 [N]              decode state set (reference `transitions.keys`, :51,:77)
 *   observations: point index n = offsets[s] + t, P = offsets[S]
 *     disc  u8  [D][P]   symbol ids; id >= V[d] (e.g. 255) = never-seen symbol, P = 0
 *                         (reference Edge.scala:70-73)
 *     cont  f64 [C][P] or f32 [C][P]
 *     labels i32 [P]     hidden-state ids (supervised learning / clamping), optional
 */
#ifndef DNBC_B200_H
#define DNBC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNBC_B200_ABI_VERSION 1

typedef struct dnbc_ctx dnbc_ctx;

enum {
  DNBC_OK = 0,
  DNBC_EINVAL = 1,      /* bad argument / shape (message mirrors the reference's text) */
  DNBC_ECUDA = 2,       /* CUDA runtime error                                           */
  DNBC_ENODEVICE = 3,   /* no usable GPU: the product has no CPU path                    */
  DNBC_ESTATE = 4,      /* call order (e.g. decode before upload)                       */
  DNBC_ENOMEM = 5,
  DNBC_ECOMM = 6        /* NCCL error                                                   */
};

/* decode modes.  REFERENCE reproduces viterbiCompute
 * (DynamicNaiveBayesianClassifier.scala:67-91) including its two quirks; BACKTRACE is
 * the textbook max-product recursion with an on-device back-pointer trace. */
enum { DNBC_DECODE_REFERENCE = 0, DNBC_DECODE_BACKTRACE = 1 };

/* arithmetic of the decode / emission path.  F32 is the throughput path; F64 mirrors
 * the reference's JVM doubles operation by operation and is what path-for-path parity
 * tests use. */
enum { DNBC_PREC_F32 = 0, DNBC_PREC_F64 = 1 };

/* supervised-learner switches */
enum {
  DNBC_QUIRK_TRANSITIONS = 1,   /* learnTransitions tests the wrong map (:204-216) */
  DNBC_DOUBLE_COUNT_FIRST = 2   /* `firstSequence ++ sequences` on a re-iterable (:129-133) */
};

typedef struct {
  int32_t n_states;             /* N */
  int32_t n_disc;               /* D */
  int32_t n_cont;               /* C */
  const int32_t *alphabet;      /* [D]  V[d] <= 254 */
  const int32_t *mixture;       /* [C]  K[c] >= 1 (the reference's continuousVariableHints) */
  int32_t device;               /* CUDA ordinal */
  int32_t flags;                /* reserved, 0 */
} dnbc_config;

/* ---- lifetime ---------------------------------------------------------------- */
int dnbc_abi_version(void);
int dnbc_create(const dnbc_config *cfg, dnbc_ctx **out);
int dnbc_destroy(dnbc_ctx *ctx);
const char *dnbc_last_error(const dnbc_ctx *ctx);

/* ---- model: replaces the 4-map constructor (...scala:14-17) -------------------- */
int dnbc_set_params(dnbc_ctx *ctx, const double *pi, const double *A, const double *B,
                    const double *w, const double *mu, const double *var,
                    const uint8_t *active /* NULL = all */);
int dnbc_get_params(dnbc_ctx *ctx, double *pi, double *A, double *B, double *w, double *mu,
                    double *var, uint8_t *active);

/* ---- observations: replaces Iterable[Seq[State]] / Seq[ObservedState] ---------- */
/* exactly one of cont_f64 / cont_f32 when C > 0; labels may be NULL. */
int dnbc_upload(dnbc_ctx *ctx, int64_t n_seq, const int64_t *offsets, const uint8_t *disc,
                const double *cont_f64, const float *cont_f32, const int32_t *labels);
/* synthetic data drawn on the device from the CURRENT parameters (Philox), following
 * the sampling loop of test-utils/src/main/scala/TestUtils.scala:171-185:
 * n_seq sequences of length seq_len, labels included. */
int dnbc_generate(dnbc_ctx *ctx, int64_t n_seq, int32_t seq_len, uint64_t seed);
/* copies sequences [first, first+count) back (any pointer may be NULL). */
int dnbc_download(dnbc_ctx *ctx, int64_t first, int64_t count, int64_t *offsets, uint8_t *disc,
                  double *cont_f64, int32_t *labels);
/* the fp32 observation planes as resident on the device: disc u8 [D][np], cont f32 [C][np] */
int dnbc_download_f32(dnbc_ctx *ctx, int64_t first, int64_t count, uint8_t *disc, float *cont_f32);
int dnbc_data_shape(const dnbc_ctx *ctx, int64_t *n_seq, int64_t *n_points);

/* ---- emission log-probabilities (K1): emissionsSum of ...scala:52-56, 79-83 ---- */
/* out [P][N] natural-log; f64 when precision == DNBC_PREC_F64 else f32. */
int dnbc_emission_logprob(dnbc_ctx *ctx, int precision, void *out);

/* ---- supervised learner: replaces the body of mle (...scala:145-160) ----------- */
/* counts as the reference accumulates them (Edge.scala:46-51); any pointer may be NULL */
int dnbc_supervised_counts(dnbc_ctx *ctx, int flags, int64_t *cnt_pi, int64_t *cnt_A,
                           int64_t *cnt_B);
/* full supervised mle on the device: counts -> normalisation (Edge.scala:53-59); per
 * (continuous variable, state) edge 1-D k-means (Tools/KMeans.java:20-123) from the given
 * initial centroids [C][N][Kmax] (NULL = deterministic quantile seeding), then
 * EM1D.initialize / EM1D.run (Tools/ExpectationMaximization1D.java:21-145) with the
 * reference's stop rule (max_iters 30, rel_tol 0.01 when passed as <= 0).
 * ll_series [C][N][max_iters+1] and n_iters [C][N] may be NULL.  The learned
 * parameters become the context's current parameters. */
int dnbc_mle(dnbc_ctx *ctx, int flags, const double *init_centroids, int max_iters,
             double rel_tol, double *ll_series, int32_t *n_iters);

/* ---- Baum-Welch EM (K1-K4; no reference counterpart) ---------------- */
/* packed fp64 sufficient statistics: [LL | pi N | xi N*N | gamma N | cat D*N*Vmax |
 * gmm C*N*Kmax*3 (s0, s1, s2 centred on the old mean)] */
int64_t dnbc_stats_size(const dnbc_ctx *ctx);
/* E-step over the resident sequences into the device statistics buffer */
int dnbc_em_estep(dnbc_ctx *ctx, int clamp_labels);
/* dnbc_upload + dnbc_em_estep in one call with the host->device copy of each chunk of sequences
 * overlapping the kernels of the previous chunk (the reference materialises every sequence on
 * the driver before learning, DynamicNaiveBayesianClassifier.scala:129-159).  The observations
 * stay resident afterwards; host buffers should be pinned.  labels may be NULL. */
int dnbc_em_estep_streamed(dnbc_ctx *ctx, int64_t n_seq, const int64_t *offsets, const uint8_t *disc,
                           const float *cont_f32, const int32_t *labels, int clamp_labels);
/* device address of the statistics buffer (test hook; the all-reduce is dnbc_em_allreduce) */
int dnbc_em_stats_device_ptr(dnbc_ctx *ctx, void **ptr);
int dnbc_em_get_stats(dnbc_ctx *ctx, double *stats);
int dnbc_em_set_stats(dnbc_ctx *ctx, const double *stats);
/* M-step from the statistics buffer; total_sequences = S summed over all shards */
int dnbc_em_mstep(dnbc_ctx *ctx, int64_t total_sequences);
/* Variance floor of the M-step.  The reference's per-edge EM has none (a collapsed component
 * gives var = 0 and NaN responsibilities, Tools/ExpectationMaximization1D.java:92-124); here a new
 * variance below rel_floor x (pooled variance of that observed variable) is raised to that value.
 * Default 1e-10; 0 disables.  dnbc_em_floored_components: how many components were raised since
 * dnbc_create. */
int dnbc_em_set_variance_floor(dnbc_ctx *ctx, double rel_floor);
int dnbc_em_floored_components(dnbc_ctx *ctx, int64_t *count);
/* n_iters x (E-step, M-step) on one GPU; loglik_out[i] = log-likelihood of the
 * parameters entering iteration i */
int dnbc_em_iterate(dnbc_ctx *ctx, int n_iters, double *loglik_out);
/* posterior state marginals gamma [P][N] f32 of the last E-step (test hook) */
int dnbc_em_get_posteriors(dnbc_ctx *ctx, float *gamma);

/* ---- multi-GPU: sequences shard over contexts, one statistics all-reduce per iteration ------
 * Replaces the reference's three Spark scatter/gather rounds through the driver
 * (DynamicNaiveBayesianClassifier.scala:231-232, 241-243, 250-252).  Every context holds its own
 * block of sequences; after the local E-step one ncclAllReduce(sum, ncclDouble) of the packed
 * statistics buffer, enqueued on the context's stream, leaves the global statistics on every GPU
 * and the M-step runs identically everywhere (no broadcast).  Decode needs no collective.
 *
 * Two ways to form the communicator, both owned by the library:
 *   (a) one process (or thread) per GPU: rank 0 calls dnbc_comm_unique_id and ships the 128 bytes
 *       to the other ranks by any means (the JVM shim: a Spark broadcast variable); every rank
 *       then calls dnbc_comm_init_rank on its context;
 *   (b) one process driving several GPUs (the reference's single driver thread): create one
 *       context per device and call dnbc_comm_init_all on the array; collectives for all of
 *       them are then issued together with dnbc_em_allreduce_group / dnbc_em_iterate_group.
 * Without a communicator (a single GPU) dnbc_em_allreduce is a no-op. */
#define DNBC_COMM_ID_BYTES 128
int dnbc_comm_unique_id(uint8_t id[DNBC_COMM_ID_BYTES]);
int dnbc_comm_init_rank(dnbc_ctx *ctx, int n_ranks, int rank, const uint8_t id[DNBC_COMM_ID_BYTES]);
int dnbc_comm_init_all(dnbc_ctx **ctxs, int n_ctx);
int dnbc_comm_info(const dnbc_ctx *ctx, int *n_ranks, int *rank);
int dnbc_comm_destroy(dnbc_ctx *ctx);
/* in-place all-reduce(sum, f64) of the statistics buffer on the context's stream (asynchronous) */
int dnbc_em_allreduce(dnbc_ctx *ctx);
/* the same for contexts formed with dnbc_comm_init_all (one ncclGroup over all of them) */
int dnbc_em_allreduce_group(dnbc_ctx **ctxs, int n_ctx);
/* n_iters x (local E-step, all-reduce, M-step with total_sequences = S over all shards): (a) is
 * called by every rank on its own context; (b) drives all contexts from the calling thread.
 * loglik_out[i] = GLOBAL log-likelihood of the parameters entering iteration i. */
int dnbc_em_iterate_sharded(dnbc_ctx *ctx, int n_iters, int64_t total_sequences, double *loglik_out);
int dnbc_em_iterate_group(dnbc_ctx **ctxs, int n_ctx, int n_iters, double *loglik_out);

/* ---- decode (K5): replaces inferMostLikelyHiddenStates / score ----------------- */
/* path_out u16 [P]; score_out f64 [S] (REFERENCE: sum_j delta_{T-1}(j), 0 for T == 1;
 * BACKTRACE: max_j delta_{T-1}(j)); tie_out u8 [P]: bit0 = the path argmax at this step
 * was (near-)tied, bit1 = a predecessor argmax of this step was (near-)tied (exact
 * equality in F64; relative margin in F32).  Any output may be NULL. */
int dnbc_decode(dnbc_ctx *ctx, int mode, int precision, uint16_t *path_out, double *score_out,
                uint8_t *tie_out);

/* ---- tuning / test options (none is needed in normal use; the library reads no environment) --- */
enum {
  /* cap on the points per E-step / decode chunk; 0 = sized from the free HBM.  Small values force
   * the multi-chunk pipelines on small inputs (tests). */
  DNBC_OPT_MAX_CHUNK_POINTS = 1,
  /* 1 = keep the register-resident recursions (N <= 64) on their large-input variant -- four
   * sequences per lane group, exact lane groups -- for inputs that fit one wave (tests). */
  DNBC_OPT_FB2_INTERLEAVE = 2,
  /* E-step path for N <= 16: 0 = automatic (the single-launch kernel of k_fused.cu while its estimated time
   * is below the general pipeline's launch-latency floor -- it is latency-optimal, not throughput-optimal), 1 = always
   * the fused kernel when the shape fits its shared memory, 2 = always the general pipeline. */
  DNBC_OPT_FUSED_ESTEP = 3,
  /* forward / backward recursions with one lane per sequence (even N <= 12, k_fbs.cu): 0 = automatic (chunks
   * that give every SM scheduler a warp of 32 sequences), 1 = whenever the state count allows it (tests),
   * 2 = never. */
  DNBC_OPT_FB_SEQ_LANES = 4
};
int dnbc_set_option(dnbc_ctx *ctx, int option, int64_t value);

/* ---- streams / timing hooks ---------------------------------------------------- */
int dnbc_synchronize(dnbc_ctx *ctx);
/* CUDA stream the library launches on (cudaStream_t as void*) */
int dnbc_stream(dnbc_ctx *ctx, void **stream);
/* number of kernels launched by this context since creation */
int64_t dnbc_launch_count(const dnbc_ctx *ctx);
/* device milliseconds spent per kernel family since the last reset: out[8] =
 * {emission, forward, backward, xi, stats_gmm, stats_cat, mstep, decode}; only
 * collected while profiling is enabled (adds events around every launch) */
int dnbc_profile_enable(dnbc_ctx *ctx, int on);
int dnbc_profile_read(dnbc_ctx *ctx, double *ms_out, int64_t *launches_out, int reset);

#ifdef __cplusplus
}
#endif
#endif
