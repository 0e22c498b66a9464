/*
 * JNI surface of libmahout_b200.so for the sketch-similarity hot path.  One Java `native` method per
 * C-ABI entry point of include/mahout_b200.h; jni/mahout_b200_jni.c is the 1:1 glue.
 *
 * NOT COMPILED IN THIS REPOSITORY: the build image has no JDK (no javac, no jni.h).  The file is
 * the binding a maintainer adds to the reference tree (mr/src/main/java/...); see INTEGRATION.md.
 */
package org.apache.mahout.cf.taste.impl.common;

import java.nio.ByteBuffer;

public final class NativeSketch {

  static {
    // resolved through -Djava.library.path, which bin/mahout already derives from $JAVA_LIBRARY_PATH
    // (bin/mahout:338-340); same loading convention as the ViennaCL backend (Context.scala:61-65)
    System.loadLibrary("mahout_b200_jni");
  }

  private NativeSketch() { }

  public static final int MEM_HOST = 0;
  public static final int DTYPE_F16 = 0;
  public static final int DTYPE_BF16 = 1;
  public static final int PRECISION_TENSOR = 0;
  public static final int PRECISION_RESCORED = 1;
  public static final int PRECISION_CERTIFIED = 2;
  /** words of the stats array of jobItemSimilarity: gpus, events, rows, similarities kept, fallback rows,
   *  events on the busiest GPU, route / build / cosine milliseconds */
  public static final int JOB_STATS_WORDS = 9;

  public static int precisionOf(String name) {
    if ("tensor".equals(name)) {
      return PRECISION_TENSOR;
    }
    if ("certified".equals(name)) {
      return PRECISION_CERTIFIED;
    }
    if ("rescored".equals(name)) {
      return PRECISION_RESCORED;
    }
    throw new IllegalArgumentException("--precision must be rescored, certified or tensor");
  }

  /** mb200_create / mb200_destroy: returns the opaque context handle. Throws if there is no B200. */
  public static native long createContext(int device);
  public static native void destroyContext(long ctx);

  /** mb200_hash_params: a[i], b[i] of HashFunctionBuilder(seed) (HashFunctionBuilder.java:40-60). */
  public static native void hashParams(long seed, int depth, long[] a, long[] b);

  /** mb200_cm_dims: throws AbstractCountMinSketch.CMException-compatible IllegalArgumentException. */
  public static native int[] cmDims(double delta, double epsilon);

  /** mb200_bank_create_params / mb200_bank_destroy / mb200_bank_clear */
  public static native long createBank(long ctx, long entities, int depth, int width, long[] a, long[] b,
                                       int fracBits);
  public static native void destroyBank(long bank);
  public static native void clearBank(long bank);

  /**
   * mb200_bank_update: direct buffers over pinned memory (allocPinned) hold little-endian
   * int64 entity / int64 key / float32 increment arrays; entity may be null for a single sketch.
   */
  public static native void update(long bank, ByteBuffer entity, ByteBuffer key, ByteBuffer inc, long n);
  public static native void updateOne(long bank, long entity, long key, double inc);
  public static native void check(long bank);

  /** mb200_bank_query: DoubleCountMinSketch.get (DoubleCountMinSketch.java:94-103). */
  public static native double query(long bank, long entity, long key);
  public static native void queryMany(long bank, long[] entity, long[] key, double[] out);

  /** mb200_bank_read: counters of entities [e0, e1) as doubles, out[(e - e0)][i][j]. */
  public static native void read(long bank, long e0, long e1, double[] out);

  /** mb200_bank_cross_cosine: DoubleCountMinSketch.cosine (DoubleCountMinSketch.java:114-149). */
  public static native double cosine(long bankA, long entityA, long bankB, long entityB);

  /**
   * mb200_bank_cosine_topk: RowSimilarityJob semantics with the sketch cosine; outIdx / outSim are
   * [entities][k], outCnt [entities].
   */
  public static native void cosineTopK(long bank, int k, double threshold, boolean excludeSelf, int dtype,
                                       int precision, long[] outIdx, double[] outSim, int[] outCnt);

  /**
   * Ingest on the GPU.  parsePrefs = ToEntityPrefsMapper.map over one text split held in a direct
   * buffer (ToEntityPrefsMapper.java:56-76); prepare = idToIndex + ItemIDIndexReducer +
   * ToUserVectorsReducer (last pref of a (user, index) pair wins, minPrefsPerUser); updateFromPrefs feeds
   * the prepared, still device-resident events to mb200_bank_update (entity = matrix row, key = userID).
   */
  public static native long parsePrefs(long ctx, ByteBuffer text, long bytes, boolean booleanData, float ratingShift,
                                       boolean transpose);
  public static native long eventCount(long events);
  public static native void destroyEvents(long events);
  public static native long prepare(long events, int minPrefsPerUser);
  public static native void prefsInfo(long prefs, long[] info);
  public static native void prefsTables(long prefs, long[] itemId, int[] indexValues);
  public static native void updateFromPrefs(long bank, long prefs);
  public static native void destroyPrefs(long prefs);

  /**
   * mb200_create_multi / mb200_multi_destroy / mb200_job_item_similarity: phase 1 of ItemSimilarityJob.run
   * (ItemSimilarityJob.java:146-162) as one call over numGpus GPUs of this process (0 = all).  row / key / value are
   * the prepared events (dense rows, user keys, preferences); exactMeasure selects the identity hash family (one
   * counter column per key, width = number of distinct keys).  outIdx / outSim are [numRows][k], outCnt [numRows].
   */
  public static native long createMulti(int numGpus);
  public static native void destroyMulti(long multi);
  public static native void jobItemSimilarity(long multi, long[] row, long[] key, float[] value, long numRows, int k,
                                              double threshold, int width, int depth, long seed, boolean exactMeasure,
                                              int fracBits, int dtype, int precision, long[] outIdx, double[] outSim,
                                              int[] outCnt, long[] stats);

  /** mb200_bank_update_u8: the narrow wire format (uint32 entity / key, one byte of quanta per event). */
  public static native void updateU8(long bank, ByteBuffer entity, ByteBuffer key, ByteBuffer quanta, long n);

  /** mb200_host_alloc / mb200_host_free wrapped as a direct ByteBuffer. */
  public static native ByteBuffer allocPinned(long bytes);
  public static native void freePinned(ByteBuffer buffer);
}
