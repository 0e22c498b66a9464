/*
 * The measure class behind `-s SIMILARITY_SKETCH_COSINE` / `--similarityClassname
 * org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.NativeSketchCosineSimilarity`.
 *
 * RowSimilarityJob resolves --similarityClassname either as a VectorSimilarityMeasures enum name or as a class
 * name that it instantiates reflectively (RowSimilarityJob.java:124-130), so the class has to exist and implement
 * VectorSimilarityMeasure (VectorSimilarityMeasure.java:22-32) for the flag to keep working in every job that
 * accepts it.  Its five methods are the cosine measure's (measures/CosineSimilarity.java:24-49): when one of the
 * unmodified MapReduce jobs is run with this class name the result is the exact cosine.  The native jobs
 * (NativeSketchItemSimilarityJob, NativeSketchRowSimilarityJob) recognise the class and run the count-min-sketch
 * cosine -- min over the sketch rows of DoubleCountMinSketch.cosine (DoubleCountMinSketch.java:114-149) -- on the GPU
 * instead, with the sketch shape below.
 *
 * NOT COMPILED IN THIS REPOSITORY (no JDK in the build image); see INTEGRATION.md.
 */
package org.apache.mahout.math.hadoop.similarity.cooccurrence.measures;

import org.apache.mahout.math.Vector;

public class NativeSketchCosineSimilarity implements VectorSimilarityMeasure {

  /** configuration keys the native jobs read; the defaults are the ones the benchmarks use */
  public static final String SKETCH_WIDTH = NativeSketchCosineSimilarity.class.getName() + ".sketchWidth";
  public static final String SKETCH_DEPTH = NativeSketchCosineSimilarity.class.getName() + ".sketchDepth";
  public static final String SKETCH_SEED = NativeSketchCosineSimilarity.class.getName() + ".sketchSeed";
  public static final int DEFAULT_SKETCH_WIDTH = 4096;
  public static final int DEFAULT_SKETCH_DEPTH = 4;
  public static final long DEFAULT_SKETCH_SEED = 42L;

  /** true for the enum-style alias and for this class's name */
  public static boolean selects(String similarityClassname) {
    return "SIMILARITY_SKETCH_COSINE".equals(similarityClassname)
        || NativeSketchCosineSimilarity.class.getName().equals(similarityClassname);
  }

  @Override
  public Vector normalize(Vector vector) {
    return vector.normalize();
  }

  @Override
  public double norm(Vector vector) {
    return VectorSimilarityMeasure.NO_NORM;
  }

  @Override
  public double aggregate(double nonZeroValueA, double nonZeroValueB) {
    return nonZeroValueA * nonZeroValueB;
  }

  @Override
  public double similarity(double summedAggregations, double normA, double normB, int numberOfColumns) {
    return summedAggregations;
  }

  @Override
  public boolean consider(int numNonZeroEntriesA, int numNonZeroEntriesB, double maxValueA, double maxValueB,
      double threshold) {
    return numNonZeroEntriesB >= threshold / maxValueA && numNonZeroEntriesA >= threshold / maxValueB;
  }
}
