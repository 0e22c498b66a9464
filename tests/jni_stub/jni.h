/* Minimal stand-in for <jni.h>: just enough declarations to type-check jni/mahout_b200_jni.c in an
 * image without a JDK (tests/test_abi.py::test_jni_glue_type_checks).  Not a real JNI header. */
#ifndef JNI_STUB_H
#define JNI_STUB_H
#include <stdint.h>
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef float jfloat;
typedef unsigned char jboolean;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jlongArray;
typedef jobject jdoubleArray;
typedef jobject jintArray;
typedef jobject jfloatArray;
typedef jint jsize;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*);
  void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint);
  jdouble* (*GetDoubleArrayElements)(JNIEnv*, jdoubleArray, jboolean*);
  void (*ReleaseDoubleArrayElements)(JNIEnv*, jdoubleArray, jdouble*, jint);
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
  void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jint, jint, const jlong*);
  jfloat* (*GetFloatArrayElements)(JNIEnv*, jfloatArray, jboolean*);
  void (*ReleaseFloatArrayElements)(JNIEnv*, jfloatArray, jfloat*, jint);
  jsize (*GetArrayLength)(JNIEnv*, jobject);
};
#endif
