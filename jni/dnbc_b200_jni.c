/*
 * dnbc_b200_jni.c -- JNI glue between the reference's JVM code and libdnbc_b200.so.
 *
 * Binds the natives of jni/NativeDnbc.scala (the class a maintainer adds next to
 * core/src/main/scala/DynamicNaiveBayesianClassifier.scala) to the C ABI of
 * include/dnbc_b200.h.  Pure marshalling: every function copies the JVM arrays into native
 * staging buffers, calls the dnbc_* entry point of the same name and returns its status; the
 * Scala side turns a non-zero status into `new Exception(nLastError(handle))`, the reference's
 * error convention (DynamicNaiveBayesianClassifier.scala:46, 148, 274, 276).
 *
 * JNI critical regions (GetPrimitiveArrayCritical) must not block: they stall the collector for
 * every thread of the Spark executor JVM.  So a region here only ever brackets one memcpy between
 * the JVM array and a malloc'ed staging buffer (copy_in / copy_out); cudaMalloc, kernels, stream
 * synchronisation and NCCL all run outside of it.
 *
 * Build (needs a JDK, which this repository's image does not have):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       jni/dnbc_b200_jni.c -Ldnbc_scala_b200/lib -ldnbc_b200 -o libdnbc_b200_jni.so
 * Here it is only type-checked against the C ABI with a minimal jni.h stub
 * (tests/test_cpu_oracle.py::test_jni_glue_type_checks_against_the_abi).
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dnbc_b200.h"

#define CTX(h) ((dnbc_ctx *)(intptr_t)(h))

/* native copy of a (possibly null) JVM primitive array; NULL for a null or empty array */
static void *copy_in(JNIEnv *env, jarray a, size_t elem) {
  if (!a) return NULL;
  const size_t bytes = (size_t)(*env)->GetArrayLength(env, a) * elem;
  void *buf = malloc(bytes ? bytes : 1);
  if (!buf) return NULL;
  void *p = (*env)->GetPrimitiveArrayCritical(env, a, NULL);
  if (p) {
    memcpy(buf, p, bytes);
    (*env)->ReleasePrimitiveArrayCritical(env, a, p, JNI_ABORT);
  }
  return buf;
}

/* zeroed native buffer the size of a (possibly null) JVM array, for outputs */
static void *alloc_out(JNIEnv *env, jarray a, size_t elem) {
  if (!a) return NULL;
  const size_t bytes = (size_t)(*env)->GetArrayLength(env, a) * elem;
  return calloc(bytes ? bytes : 1, 1);
}

/* copies a native output buffer back into its JVM array and frees it */
static void copy_out(JNIEnv *env, jarray a, void *buf, size_t elem) {
  if (!a || !buf) return;
  const size_t bytes = (size_t)(*env)->GetArrayLength(env, a) * elem;
  void *p = (*env)->GetPrimitiveArrayCritical(env, a, NULL);
  if (p) {
    memcpy(p, buf, bytes);
    (*env)->ReleasePrimitiveArrayCritical(env, a, p, 0);
  }
  free(buf);
}

JNIEXPORT jlong JNICALL Java_NativeDnbc_nCreate(JNIEnv *env, jobject self, jint n_states, jintArray alphabet,
                                                jintArray mixture, jint device) {
  (void)self;
  dnbc_config cfg = {0};
  cfg.n_states = n_states;
  cfg.n_disc = alphabet ? (*env)->GetArrayLength(env, alphabet) : 0;
  cfg.n_cont = mixture ? (*env)->GetArrayLength(env, mixture) : 0;
  int32_t *al = copy_in(env, alphabet, sizeof(jint)), *mx = copy_in(env, mixture, sizeof(jint));
  cfg.alphabet = al;
  cfg.mixture = mx;
  cfg.device = device;
  dnbc_ctx *ctx = NULL;
  const int rc = dnbc_create(&cfg, &ctx);
  free(mx);
  free(al);
  if (rc != DNBC_OK) {
    /* e.g. "Number of suggested normal distributions in a continuous variable must be at least one" (:276) */
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/Exception"), dnbc_last_error(NULL));
    return 0;
  }
  return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_NativeDnbc_nDestroy(JNIEnv *env, jobject self, jlong h) {
  (void)env; (void)self;
  dnbc_destroy(CTX(h));
}

JNIEXPORT jstring JNICALL Java_NativeDnbc_nLastError(JNIEnv *env, jobject self, jlong h) {
  (void)self;
  return (*env)->NewStringUTF(env, dnbc_last_error(CTX(h)));
}

JNIEXPORT jint JNICALL Java_NativeDnbc_nSetParams(JNIEnv *env, jobject self, jlong h, jdoubleArray pi, jdoubleArray a,
                                                  jdoubleArray b, jdoubleArray w, jdoubleArray mu, jdoubleArray var,
                                                  jbyteArray active) {
  (void)self;
  double *ppi = copy_in(env, pi, 8), *pa = copy_in(env, a, 8), *pb = copy_in(env, b, 8), *pw = copy_in(env, w, 8),
         *pmu = copy_in(env, mu, 8), *pv = copy_in(env, var, 8);
  uint8_t *pact = copy_in(env, active, 1);
  const int rc = dnbc_set_params(CTX(h), ppi, pa, pb, pw, pmu, pv, pact);
  free(pact); free(pv); free(pmu); free(pw); free(pb); free(pa); free(ppi);
  return rc;
}

JNIEXPORT jint JNICALL Java_NativeDnbc_nGetParams(JNIEnv *env, jobject self, jlong h, jdoubleArray pi, jdoubleArray a,
                                                  jdoubleArray b, jdoubleArray w, jdoubleArray mu, jdoubleArray var,
                                                  jbyteArray active) {
  (void)self;
  double *ppi = alloc_out(env, pi, 8), *pa = alloc_out(env, a, 8), *pb = alloc_out(env, b, 8), *pw = alloc_out(env, w, 8),
         *pmu = alloc_out(env, mu, 8), *pv = alloc_out(env, var, 8);
  uint8_t *pact = alloc_out(env, active, 1);
  const int rc = dnbc_get_params(CTX(h), ppi, pa, pb, pw, pmu, pv, pact);
  copy_out(env, active, pact, 1); copy_out(env, var, pv, 8); copy_out(env, mu, pmu, 8);
  copy_out(env, w, pw, 8); copy_out(env, b, pb, 8); copy_out(env, a, pa, 8); copy_out(env, pi, ppi, 8);
  return rc;
}

/* observations: offsets i64 [S+1], disc u8 [D][P], cont f64 [C][P], labels i32 [P] (null when decoding) */
JNIEXPORT jint JNICALL Java_NativeDnbc_nUpload(JNIEnv *env, jobject self, jlong h, jlong n_seq, jlongArray offsets,
                                               jbyteArray disc, jdoubleArray cont, jintArray labels) {
  (void)self;
  int64_t *po = copy_in(env, offsets, 8);
  uint8_t *pd = copy_in(env, disc, 1);
  double *pc = copy_in(env, cont, 8);
  int32_t *pl = copy_in(env, labels, 4);
  const int rc = dnbc_upload(CTX(h), n_seq, po, pd, pc, NULL, pl);   /* the library copies again, to the device */
  free(pl); free(pc); free(pd); free(po);
  return rc;
}

/* supervised learner = the body of `mle` (DynamicNaiveBayesianClassifier.scala:145-160 + learnFinalize :222-259) */
JNIEXPORT jint JNICALL Java_NativeDnbc_nMle(JNIEnv *env, jobject self, jlong h, jint flags, jdoubleArray init_centroids,
                                            jint max_iters, jdouble rel_tol) {
  (void)self;
  double *pc = copy_in(env, init_centroids, 8);
  const int rc = dnbc_mle(CTX(h), flags, pc, max_iters, rel_tol, NULL, NULL);
  free(pc);
  return rc;
}

/* inferMostLikelyHiddenStates / score (:24-39): path u16 [P] (as jshort), score f64 [S], tie u8 [P]; any may be null */
JNIEXPORT jint JNICALL Java_NativeDnbc_nDecode(JNIEnv *env, jobject self, jlong h, jint mode, jint precision,
                                               jshortArray path, jdoubleArray score, jbyteArray tie) {
  (void)self;
  uint16_t *pp = alloc_out(env, path, 2);
  double *ps = alloc_out(env, score, 8);
  uint8_t *pt = alloc_out(env, tie, 1);
  const int rc = dnbc_decode(CTX(h), mode, precision, pp, ps, pt);
  copy_out(env, tie, pt, 1); copy_out(env, score, ps, 8); copy_out(env, path, pp, 2);
  return rc;
}

/* Baum-Welch (new, additive): n_iters x (E-step, M-step) on this context's GPU */
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmIterate(JNIEnv *env, jobject self, jlong h, jint n_iters, jdoubleArray loglik) {
  (void)self;
  double *pl = alloc_out(env, loglik, 8);
  const int rc = dnbc_em_iterate(CTX(h), n_iters, pl);
  copy_out(env, loglik, pl, 8);
  return rc;
}

/* ---- multi-GPU learning: the library owns the NCCL communicator and the all-reduce ----------------
 * replaces the three Spark parallelize/map/collect rounds of learnFinalize (:231-232, 241-243, 250-252) */
JNIEXPORT jint JNICALL Java_NativeDnbc_nCommUniqueId(JNIEnv *env, jobject self, jbyteArray id) {
  (void)self;
  if (!id || (*env)->GetArrayLength(env, id) != DNBC_COMM_ID_BYTES) return DNBC_EINVAL;
  uint8_t *p = alloc_out(env, id, 1);
  const int rc = dnbc_comm_unique_id(p);
  copy_out(env, id, p, 1);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nCommInitRank(JNIEnv *env, jobject self, jlong h, jint n_ranks, jint rank,
                                                     jbyteArray id) {
  (void)self;
  if (!id || (*env)->GetArrayLength(env, id) != DNBC_COMM_ID_BYTES) return DNBC_EINVAL;
  uint8_t *p = copy_in(env, id, 1);
  const int rc = dnbc_comm_init_rank(CTX(h), n_ranks, rank, p);
  free(p);
  return rc;
}
/* one JVM driving several GPUs: handles = the dnbc_ctx of every device, in rank order */
static dnbc_ctx **ctx_array(JNIEnv *env, jlongArray handles, int *n) {
  *n = handles ? (*env)->GetArrayLength(env, handles) : 0;
  jlong *hs = copy_in(env, handles, sizeof(jlong));
  dnbc_ctx **ctxs = calloc(*n > 0 ? (size_t)*n : 1, sizeof(dnbc_ctx *));
  for (int i = 0; ctxs && hs && i < *n; ++i) ctxs[i] = CTX(hs[i]);
  free(hs);
  return ctxs;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nCommInitAll(JNIEnv *env, jobject self, jlongArray handles) {
  (void)self;
  int n = 0;
  dnbc_ctx **ctxs = ctx_array(env, handles, &n);
  const int rc = ctxs ? dnbc_comm_init_all(ctxs, n) : DNBC_ENOMEM;
  free(ctxs);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nCommDestroy(JNIEnv *env, jobject self, jlong h) {
  (void)env; (void)self;
  return dnbc_comm_destroy(CTX(h));
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmAllreduce(JNIEnv *env, jobject self, jlong h) {
  (void)env; (void)self;
  return dnbc_em_allreduce(CTX(h));
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmIterateSharded(JNIEnv *env, jobject self, jlong h, jint n_iters,
                                                         jlong total_sequences, jdoubleArray loglik) {
  (void)self;
  double *pl = alloc_out(env, loglik, 8);
  const int rc = dnbc_em_iterate_sharded(CTX(h), n_iters, total_sequences, pl);
  copy_out(env, loglik, pl, 8);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmIterateGroup(JNIEnv *env, jobject self, jlongArray handles, jint n_iters,
                                                       jdoubleArray loglik) {
  (void)self;
  int n = 0;
  dnbc_ctx **ctxs = ctx_array(env, handles, &n);
  double *pl = alloc_out(env, loglik, 8);
  const int rc = ctxs ? dnbc_em_iterate_group(ctxs, n, n_iters, pl) : DNBC_ENOMEM;
  copy_out(env, loglik, pl, 8);
  free(ctxs);
  return rc;
}

/* finer-grained pieces (an application that interleaves its own work between the phases) */
JNIEXPORT jlong JNICALL Java_NativeDnbc_nStatsSize(JNIEnv *env, jobject self, jlong h) {
  (void)env; (void)self;
  return (jlong)dnbc_stats_size(CTX(h));
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmEstep(JNIEnv *env, jobject self, jlong h, jboolean clamp_labels) {
  (void)env; (void)self;
  return dnbc_em_estep(CTX(h), clamp_labels ? 1 : 0);
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmGetStats(JNIEnv *env, jobject self, jlong h, jdoubleArray stats) {
  (void)self;
  double *p = alloc_out(env, stats, 8);
  const int rc = dnbc_em_get_stats(CTX(h), p);
  copy_out(env, stats, p, 8);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmSetStats(JNIEnv *env, jobject self, jlong h, jdoubleArray stats) {
  (void)self;
  double *p = copy_in(env, stats, 8);
  const int rc = dnbc_em_set_stats(CTX(h), p);
  free(p);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmMstep(JNIEnv *env, jobject self, jlong h, jlong total_sequences) {
  (void)env; (void)self;
  return dnbc_em_mstep(CTX(h), total_sequences);
}
