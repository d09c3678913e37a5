/*
 * dnbc_b200_jni.c -- JNI glue between the reference's JVM code and libdnbc_b200.so.
 *
 * Binds the natives of jni/NativeDnbc.scala (the class a maintainer adds next to
 * core/src/main/scala/DynamicNaiveBayesianClassifier.scala) to the C ABI of
 * include/dnbc_b200.h.  Pure marshalling: every function pins the JVM arrays, calls the
 * dnbc_* entry point of the same name and returns its status; the Scala side turns a non-zero
 * status into `new Exception(nLastError(handle))`, the reference's error convention
 * (DynamicNaiveBayesianClassifier.scala:46, 148, 274, 276).
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

#include "dnbc_b200.h"

#define CTX(h) ((dnbc_ctx *)(intptr_t)(h))

/* pinned views of (possibly null) JVM primitive arrays */
static void *pin(JNIEnv *env, jarray a) { return a ? (*env)->GetPrimitiveArrayCritical(env, a, NULL) : NULL; }
static void unpin(JNIEnv *env, jarray a, void *p, jint mode) {
  if (a) (*env)->ReleasePrimitiveArrayCritical(env, a, p, mode);
}

JNIEXPORT jlong JNICALL Java_NativeDnbc_nCreate(JNIEnv *env, jobject self, jint n_states, jintArray alphabet,
                                                jintArray mixture, jint device) {
  (void)self;
  dnbc_config cfg = {0};
  cfg.n_states = n_states;
  cfg.n_disc = alphabet ? (*env)->GetArrayLength(env, alphabet) : 0;
  cfg.n_cont = mixture ? (*env)->GetArrayLength(env, mixture) : 0;
  jint *al = (jint *)pin(env, alphabet), *mx = (jint *)pin(env, mixture);
  cfg.alphabet = (const int32_t *)al;
  cfg.mixture = (const int32_t *)mx;
  cfg.device = device;
  dnbc_ctx *ctx = NULL;
  const int rc = dnbc_create(&cfg, &ctx);
  unpin(env, mixture, mx, JNI_ABORT);
  unpin(env, alphabet, al, JNI_ABORT);
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
  double *ppi = pin(env, pi), *pa = pin(env, a), *pb = pin(env, b), *pw = pin(env, w), *pmu = pin(env, mu), *pv = pin(env, var);
  uint8_t *pact = pin(env, active);
  const int rc = dnbc_set_params(CTX(h), ppi, pa, pb, pw, pmu, pv, pact);
  unpin(env, active, pact, JNI_ABORT); unpin(env, var, pv, JNI_ABORT); unpin(env, mu, pmu, JNI_ABORT);
  unpin(env, w, pw, JNI_ABORT); unpin(env, b, pb, JNI_ABORT); unpin(env, a, pa, JNI_ABORT); unpin(env, pi, ppi, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_NativeDnbc_nGetParams(JNIEnv *env, jobject self, jlong h, jdoubleArray pi, jdoubleArray a,
                                                  jdoubleArray b, jdoubleArray w, jdoubleArray mu, jdoubleArray var,
                                                  jbyteArray active) {
  (void)self;
  double *ppi = pin(env, pi), *pa = pin(env, a), *pb = pin(env, b), *pw = pin(env, w), *pmu = pin(env, mu), *pv = pin(env, var);
  uint8_t *pact = pin(env, active);
  const int rc = dnbc_get_params(CTX(h), ppi, pa, pb, pw, pmu, pv, pact);
  /* mode 0: copy the results back into the JVM arrays */
  unpin(env, active, pact, 0); unpin(env, var, pv, 0); unpin(env, mu, pmu, 0);
  unpin(env, w, pw, 0); unpin(env, b, pb, 0); unpin(env, a, pa, 0); unpin(env, pi, ppi, 0);
  return rc;
}

/* observations: offsets i64 [S+1], disc u8 [D][P], cont f64 [C][P], labels i32 [P] (null when decoding) */
JNIEXPORT jint JNICALL Java_NativeDnbc_nUpload(JNIEnv *env, jobject self, jlong h, jlong n_seq, jlongArray offsets,
                                               jbyteArray disc, jdoubleArray cont, jintArray labels) {
  (void)self;
  int64_t *po = pin(env, offsets);
  uint8_t *pd = pin(env, disc);
  double *pc = pin(env, cont);
  int32_t *pl = pin(env, labels);
  const int rc = dnbc_upload(CTX(h), n_seq, po, pd, pc, NULL, pl);   /* the library copies: arrays are released on return */
  unpin(env, labels, pl, JNI_ABORT); unpin(env, cont, pc, JNI_ABORT); unpin(env, disc, pd, JNI_ABORT);
  unpin(env, offsets, po, JNI_ABORT);
  return rc;
}

/* supervised learner = the body of `mle` (DynamicNaiveBayesianClassifier.scala:145-160 + learnFinalize :222-259) */
JNIEXPORT jint JNICALL Java_NativeDnbc_nMle(JNIEnv *env, jobject self, jlong h, jint flags, jdoubleArray init_centroids,
                                            jint max_iters, jdouble rel_tol) {
  (void)self;
  double *pc = pin(env, init_centroids);
  const int rc = dnbc_mle(CTX(h), flags, pc, max_iters, rel_tol, NULL, NULL);
  unpin(env, init_centroids, pc, JNI_ABORT);
  return rc;
}

/* inferMostLikelyHiddenStates / score (:24-39): path u16 [P] (as jshort), score f64 [S], tie u8 [P]; any may be null */
JNIEXPORT jint JNICALL Java_NativeDnbc_nDecode(JNIEnv *env, jobject self, jlong h, jint mode, jint precision,
                                               jshortArray path, jdoubleArray score, jbyteArray tie) {
  (void)self;
  uint16_t *pp = pin(env, path);
  double *ps = pin(env, score);
  uint8_t *pt = pin(env, tie);
  const int rc = dnbc_decode(CTX(h), mode, precision, pp, ps, pt);
  unpin(env, tie, pt, 0); unpin(env, score, ps, 0); unpin(env, path, pp, 0);
  return rc;
}

/* Baum-Welch (new, additive): n_iters x (E-step, M-step) on this context's GPU */
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmIterate(JNIEnv *env, jobject self, jlong h, jint n_iters, jdoubleArray loglik) {
  (void)self;
  double *pl = pin(env, loglik);
  const int rc = dnbc_em_iterate(CTX(h), n_iters, pl);
  unpin(env, loglik, pl, 0);
  return rc;
}

/* multi-GPU learning from the JVM: one context per GPU, statistics summed by the caller */
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
  double *p = pin(env, stats);
  const int rc = dnbc_em_get_stats(CTX(h), p);
  unpin(env, stats, p, 0);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmSetStats(JNIEnv *env, jobject self, jlong h, jdoubleArray stats) {
  (void)self;
  double *p = pin(env, stats);
  const int rc = dnbc_em_set_stats(CTX(h), p);
  unpin(env, stats, p, JNI_ABORT);
  return rc;
}
JNIEXPORT jint JNICALL Java_NativeDnbc_nEmMstep(JNIEnv *env, jobject self, jlong h, jlong total_sequences) {
  (void)env; (void)self;
  return dnbc_em_mstep(CTX(h), total_sequences);
}
