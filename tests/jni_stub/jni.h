/* Minimal stand-in for <jni.h>: just the types and JNIEnv members jni/dnbc_b200_jni.c uses, so that the
 * glue can be TYPE-CHECKED against include/dnbc_b200.h in an image without a JDK.  Not a JNI
 * implementation; never linked. */
#ifndef DNBC_TEST_JNI_STUB_H
#define DNBC_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef int16_t jshort;
typedef double jdouble;
typedef uint8_t jboolean;
typedef jint jsize;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jbyteArray;
typedef jarray jshortArray;
typedef jarray jdoubleArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  jstring (*NewStringUTF)(JNIEnv *, const char *);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);
};
#endif
