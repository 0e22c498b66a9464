/*
 * ItemSimilarityJob with phase 1 (RowSimilarityJob, 4 MapReduce jobs) replaced by the native call
 * sequence; phases 0 (PreparePreferenceMatrixJob) and 2 (MostSimilarItemPairs) and every flag are the
 * reference's (ItemSimilarityJob.java:97-179).  Additive flags: --sketchWidth --sketchDepth
 * --sketchSeed --precision.  NOT COMPILED HERE (no JDK / Hadoop jars in the build image).
 *
 * Sketch of the replaced phase (the rest of run() is unchanged and omitted):
 *
 *   if (shouldRunNextPhase(parsedArgs, currentPhase)) {
 *     // rating matrix rows (item index -> user vector) written by phase 0
 *     long ctx = NativeSketch.createContext(0);
 *     long bank = NativeSketch.createBank(ctx, numItems, depth, width, a, b, fracBits);
 *     for (Pair<IntWritable,VectorWritable> row : new SequenceFileDirIterable<>(ratingMatrix, ...)) {
 *       // entity = dense row of the item index, key = user column, increment = preference
 *       appendEvents(entityBuf, keyBuf, incBuf, row);            // pinned direct buffers
 *       if (full) NativeSketch.update(bank, entityBuf, keyBuf, incBuf, n);
 *     }
 *     NativeSketch.check(bank);
 *     NativeSketch.cosineTopK(bank, maxSimilarItemsPerItem, threshold, true,
 *                             NativeSketch.DTYPE_F16, NativeSketch.PRECISION_RESCORED, idx, sim, cnt);
 *     // similarity matrix rows in the format phase 2 reads (SequenceFile<IntWritable,VectorWritable>)
 *     writeSimilarityMatrix(similarityMatrixPath, idx, sim, cnt);
 *   }
 *
 * mahout_b200/itemsimilarity.py is the executable mirror of this job used by the parity tests.
 */
package org.apache.mahout.cf.taste.hadoop.similarity.item;

public final class NativeSketchItemSimilarityJob {
  private NativeSketchItemSimilarityJob() { }
}
