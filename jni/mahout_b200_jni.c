/*
 * JNI glue: Java_org_apache_mahout_cf_taste_impl_common_NativeSketch_* -> include/mahout_b200.h.
 * 1:1, no logic.  Status codes map to exceptions the reference throws at the same places:
 *   MB200_ERR_BAD_ARG            -> IllegalArgumentException (Guava Preconditions, DoubleCountMinSketch.java:117)
 *   MB200_ERR_CM_DELTA/_EPSILON  -> IllegalArgumentException carrying the CMException text
 *   everything else              -> org.apache.mahout.cf.taste.common.TasteException
 *
 * NOT BUILT IN THIS REPOSITORY: the image has no jni.h.  Build where a JDK exists with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       jni/mahout_b200_jni.c -Lmahout_b200 -lmahout_b200 -o libmahout_b200_jni.so
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "mahout_b200.h"

#define JNI_FN(name) Java_org_apache_mahout_cf_taste_impl_common_NativeSketch_##name

static void throw_status(JNIEnv* env, int rc, mb200_ctx* ctx) {
  const char* cls = (rc == MB200_ERR_BAD_ARG || rc == MB200_ERR_CM_DELTA || rc == MB200_ERR_CM_EPSILON)
                        ? "java/lang/IllegalArgumentException"
                        : "org/apache/mahout/cf/taste/common/TasteException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), mb200_last_error(ctx));
}
#define CHECK(rc, ctx) do { int rc_ = (rc); if (rc_ != MB200_OK) { throw_status(env, rc_, (ctx)); } } while (0)

JNIEXPORT jlong JNICALL JNI_FN(createContext)(JNIEnv* env, jclass c, jint device) {
  mb200_ctx* ctx = NULL;
  CHECK(mb200_create(device, &ctx), NULL);
  return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL JNI_FN(destroyContext)(JNIEnv* env, jclass c, jlong ctx) {
  mb200_destroy((mb200_ctx*)(intptr_t)ctx);
}

JNIEXPORT void JNICALL JNI_FN(hashParams)(JNIEnv* env, jclass c, jlong seed, jint depth, jlongArray a, jlongArray b) {
  jlong* pa = (*env)->GetLongArrayElements(env, a, NULL);
  jlong* pb = (*env)->GetLongArrayElements(env, b, NULL);
  int rc = mb200_hash_params(seed, depth, (int64_t*)pa, (int64_t*)pb);
  (*env)->ReleaseLongArrayElements(env, a, pa, 0);
  (*env)->ReleaseLongArrayElements(env, b, pb, 0);
  CHECK(rc, NULL);
}

JNIEXPORT jlong JNICALL JNI_FN(createBank)(JNIEnv* env, jclass c, jlong ctx, jlong entities, jint depth, jint width,
                                           jlongArray a, jlongArray b, jint fracBits) {
  mb200_bank* bank = NULL;
  jlong* pa = (*env)->GetLongArrayElements(env, a, NULL);
  jlong* pb = (*env)->GetLongArrayElements(env, b, NULL);
  int rc = mb200_bank_create_params((mb200_ctx*)(intptr_t)ctx, entities, depth, width, (int64_t*)pa, (int64_t*)pb,
                                    fracBits, &bank);
  (*env)->ReleaseLongArrayElements(env, a, pa, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, b, pb, JNI_ABORT);
  CHECK(rc, (mb200_ctx*)(intptr_t)ctx);
  return (jlong)(intptr_t)bank;
}

JNIEXPORT void JNICALL JNI_FN(destroyBank)(JNIEnv* env, jclass c, jlong bank) {
  mb200_bank_destroy((mb200_bank*)(intptr_t)bank);
}

JNIEXPORT void JNICALL JNI_FN(update)(JNIEnv* env, jclass c, jlong bank, jobject entity, jobject key, jobject inc, jlong n) {
  const int64_t* e = entity ? (const int64_t*)(*env)->GetDirectBufferAddress(env, entity) : NULL;
  const int64_t* k = (const int64_t*)(*env)->GetDirectBufferAddress(env, key);
  const float* v = (const float*)(*env)->GetDirectBufferAddress(env, inc);
  CHECK(mb200_bank_update((mb200_bank*)(intptr_t)bank, e, k, v, n, MB200_MEM_HOST), NULL);
}

JNIEXPORT void JNICALL JNI_FN(updateOne)(JNIEnv* env, jclass c, jlong bank, jlong entity, jlong key, jdouble inc) {
  int64_t e = entity, k = key;
  double v = inc;
  CHECK(mb200_bank_update_f64((mb200_bank*)(intptr_t)bank, &e, &k, &v, 1, MB200_MEM_HOST), NULL);
}

JNIEXPORT jdouble JNICALL JNI_FN(query)(JNIEnv* env, jclass c, jlong bank, jlong entity, jlong key) {
  int64_t e = entity, k = key;
  double out = 0.0;
  CHECK(mb200_bank_query((mb200_bank*)(intptr_t)bank, &e, &k, 1, &out, MB200_MEM_HOST), NULL);
  return out;
}

JNIEXPORT jdouble JNICALL JNI_FN(cosine)(JNIEnv* env, jclass c, jlong bankA, jlong ea, jlong bankB, jlong eb) {
  int64_t a = ea, b = eb;
  double out = 0.0;
  CHECK(mb200_bank_cross_cosine((mb200_bank*)(intptr_t)bankA, &a, (mb200_bank*)(intptr_t)bankB, &b, 1, &out,
                                MB200_MEM_HOST), NULL);
  return out;
}

JNIEXPORT void JNICALL JNI_FN(cosineTopK)(JNIEnv* env, jclass c, jlong bank, jint k, jdouble threshold,
                                          jboolean excludeSelf, jint dtype, jint precision, jlongArray outIdx,
                                          jdoubleArray outSim, jintArray outCnt) {
  jlong* pi = (*env)->GetLongArrayElements(env, outIdx, NULL);
  jdouble* ps = (*env)->GetDoubleArrayElements(env, outSim, NULL);
  jint* pc = (*env)->GetIntArrayElements(env, outCnt, NULL);
  int rc = mb200_bank_cosine_topk((mb200_bank*)(intptr_t)bank, k, threshold, excludeSelf, dtype, precision,
                                  (int64_t*)pi, ps, (int32_t*)pc, MB200_MEM_HOST);
  (*env)->ReleaseLongArrayElements(env, outIdx, pi, 0);
  (*env)->ReleaseDoubleArrayElements(env, outSim, ps, 0);
  (*env)->ReleaseIntArrayElements(env, outCnt, pc, 0);
  CHECK(rc, NULL);
}

JNIEXPORT jobject JNICALL JNI_FN(allocPinned)(JNIEnv* env, jclass c, jlong bytes) {
  void* p = NULL;
  CHECK(mb200_host_alloc(bytes, &p), NULL);
  return p ? (*env)->NewDirectByteBuffer(env, p, bytes) : NULL;
}

JNIEXPORT void JNICALL JNI_FN(freePinned)(JNIEnv* env, jclass c, jobject buffer) {
  mb200_host_free((*env)->GetDirectBufferAddress(env, buffer));
}
/* ---- ingest (PreparePreferenceMatrixJob on the GPU) ------------------------------------------------ */
JNIEXPORT jlong JNICALL JNI_FN(parsePrefs)(JNIEnv* env, jclass c, jlong ctx, jobject text, jlong bytes,
                                           jboolean booleanData, jfloat ratingShift, jboolean transpose) {
  mb200_events* ev = NULL;
  CHECK(mb200_events_parse((mb200_ctx*)(intptr_t)ctx, (const char*)(*env)->GetDirectBufferAddress(env, text), bytes,
                           MB200_MEM_HOST, booleanData, ratingShift, transpose, &ev), (mb200_ctx*)(intptr_t)ctx);
  return (jlong)(intptr_t)ev;
}

JNIEXPORT jlong JNICALL JNI_FN(eventCount)(JNIEnv* env, jclass c, jlong events) {
  int64_t n = 0;
  CHECK(mb200_events_count((mb200_events*)(intptr_t)events, &n), NULL);
  return n;
}

JNIEXPORT void JNICALL JNI_FN(destroyEvents)(JNIEnv* env, jclass c, jlong events) {
  mb200_events_destroy((mb200_events*)(intptr_t)events);
}

JNIEXPORT jlong JNICALL JNI_FN(prepare)(JNIEnv* env, jclass c, jlong events, jint minPrefsPerUser) {
  mb200_prefs* p = NULL;
  CHECK(mb200_events_prepare((mb200_events*)(intptr_t)events, minPrefsPerUser, &p), NULL);
  return (jlong)(intptr_t)p;
}

/* info[0..2] = surviving events, items (matrix rows), users */
JNIEXPORT void JNICALL JNI_FN(prefsInfo)(JNIEnv* env, jclass c, jlong prefs, jlongArray info) {
  int64_t v[3] = {0, 0, 0};
  CHECK(mb200_prefs_info((mb200_prefs*)(intptr_t)prefs, &v[0], &v[1], &v[2]), NULL);
  (*env)->SetLongArrayRegion(env, info, 0, 3, (const jlong*)v);
}

JNIEXPORT void JNICALL JNI_FN(prefsTables)(JNIEnv* env, jclass c, jlong prefs, jlongArray itemId, jintArray indexValues) {
  jlong* pi = (*env)->GetLongArrayElements(env, itemId, NULL);
  jint* px = (*env)->GetIntArrayElements(env, indexValues, NULL);
  int rc = mb200_prefs_tables((mb200_prefs*)(intptr_t)prefs, (int64_t*)pi, (int32_t*)px);
  (*env)->ReleaseLongArrayElements(env, itemId, pi, 0);
  (*env)->ReleaseIntArrayElements(env, indexValues, px, 0);
  CHECK(rc, NULL);
}

/* K1 straight from the prepared (device-resident) events: entity = row, key = user, inc = pref */
JNIEXPORT void JNICALL JNI_FN(updateFromPrefs)(JNIEnv* env, jclass c, jlong bank, jlong prefs) {
  int64_t *row = NULL, *user = NULL, n = 0;
  float* pref = NULL;
  CHECK(mb200_prefs_info((mb200_prefs*)(intptr_t)prefs, &n, NULL, NULL), NULL);
  CHECK(mb200_prefs_columns((mb200_prefs*)(intptr_t)prefs, &row, &user, &pref), NULL);
  CHECK(mb200_bank_update((mb200_bank*)(intptr_t)bank, row, user, pref, n, MB200_MEM_DEVICE), NULL);
}

JNIEXPORT void JNICALL JNI_FN(destroyPrefs)(JNIEnv* env, jclass c, jlong prefs) {
  mb200_prefs_destroy((mb200_prefs*)(intptr_t)prefs);
}
/* clearBank, check, queryMany, read, cmDims follow the same pattern (one C call each). */

/* ---- the whole similarity phase over all GPUs of the process (csrc/job.cu) ------------------------------------ */
static void throw_multi(JNIEnv* env, int rc, mb200_multi* m) {
  const char* cls = rc == MB200_ERR_BAD_ARG ? "java/lang/IllegalArgumentException"
                                            : "org/apache/mahout/cf/taste/common/TasteException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), mb200_multi_last_error(m));
}

JNIEXPORT jlong JNICALL JNI_FN(createMulti)(JNIEnv* env, jclass c, jint numGpus) {
  mb200_multi* m = NULL;
  int rc = mb200_create_multi(numGpus, NULL, &m);
  if (rc != MB200_OK) throw_multi(env, rc, NULL);
  return (jlong)(intptr_t)m;
}

JNIEXPORT void JNICALL JNI_FN(destroyMulti)(JNIEnv* env, jclass c, jlong multi) {
  mb200_multi_destroy((mb200_multi*)(intptr_t)multi);
}

JNIEXPORT void JNICALL JNI_FN(jobItemSimilarity)(JNIEnv* env, jclass c, jlong multi, jlongArray row, jlongArray key,
                                                 jfloatArray value, jlong numRows, jint k, jdouble threshold, jint width,
                                                 jint depth, jlong seed, jboolean exactMeasure, jint fracBits, jint dtype,
                                                 jint precision, jlongArray outIdx, jdoubleArray outSim, jintArray outCnt,
                                                 jlongArray stats) {
  mb200_multi* m = (mb200_multi*)(intptr_t)multi;
  const jsize n = (*env)->GetArrayLength(env, row);
  jlong* prow = (*env)->GetLongArrayElements(env, row, NULL);
  jlong* pkey = (*env)->GetLongArrayElements(env, key, NULL);
  jfloat* pval = (*env)->GetFloatArrayElements(env, value, NULL);
  jlong* pidx = (*env)->GetLongArrayElements(env, outIdx, NULL);
  jdouble* psim = (*env)->GetDoubleArrayElements(env, outSim, NULL);
  jint* pcnt = (*env)->GetIntArrayElements(env, outCnt, NULL);
  const int64_t one = 1, zero = 0;
  mb200_job_params p;
  mb200_job_stats st;
  p.k = k;
  p.threshold = threshold;
  p.width = width;
  p.depth = depth;
  p.seed = seed;
  p.hash_a = exactMeasure ? &one : NULL;
  p.hash_b = exactMeasure ? &zero : NULL;
  p.frac_bits = fracBits;
  p.dtype = dtype;
  p.precision = precision;
  int rc = mb200_job_item_similarity(m, (const int64_t*)prow, (const int64_t*)pkey, (const float*)pval, n, numRows, &p,
                                     (int64_t*)pidx, (double*)psim, (int32_t*)pcnt, &st);
  (*env)->ReleaseLongArrayElements(env, row, prow, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, key, pkey, JNI_ABORT);
  (*env)->ReleaseFloatArrayElements(env, value, pval, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, outIdx, pidx, 0);
  (*env)->ReleaseDoubleArrayElements(env, outSim, psim, 0);
  (*env)->ReleaseIntArrayElements(env, outCnt, pcnt, 0);
  if (rc != MB200_OK) {
    throw_multi(env, rc, m);
    return;
  }
  if (stats != NULL && (*env)->GetArrayLength(env, stats) >= 9) {
    jlong w[9];
    w[0] = st.n_gpus;
    w[1] = st.events;
    w[2] = st.rows;
    w[3] = st.similarities_kept;
    w[4] = st.fallback_rows;
    w[5] = st.events_busiest_gpu;
    w[6] = (jlong)(st.route_s * 1e3);
    w[7] = (jlong)(st.build_s * 1e3);
    w[8] = (jlong)(st.cosine_s * 1e3);
    (*env)->SetLongArrayRegion(env, stats, 0, 9, w);
  }
}

JNIEXPORT void JNICALL JNI_FN(updateU8)(JNIEnv* env, jclass c, jlong bank, jobject entity, jobject key, jobject quanta,
                                        jlong n) {
  const uint32_t* e = entity ? (const uint32_t*)(*env)->GetDirectBufferAddress(env, entity) : NULL;
  const uint32_t* k = (const uint32_t*)(*env)->GetDirectBufferAddress(env, key);
  const uint8_t* q = (const uint8_t*)(*env)->GetDirectBufferAddress(env, quanta);
  CHECK(mb200_bank_update_u8((mb200_bank*)(intptr_t)bank, e, k, q, n, MB200_MEM_HOST), NULL);
}

