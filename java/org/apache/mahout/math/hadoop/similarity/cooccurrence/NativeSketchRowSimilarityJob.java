/*
 * `mahout rowsimilarity` with the similarity phase on the GPUs: same flags as RowSimilarityJob
 * (RowSimilarityJob.java:93-107) plus --sketchWidth / --sketchDepth / --sketchSeed / --numGpus / --precision.
 *
 * Input and output are RowSimilarityJob's: SequenceFile<IntWritable, VectorWritable> rows in, the top
 * --maxSimilaritiesPerRow similarities per row out (same key and index space, vectors of the input's cardinality).
 * The three MapReduce passes between them (normalisation + transposition, RowSimilarityJob.java:147-170;
 * co-occurrence + similarity, :172-200; symmetrisation + top-k, :202-216) are ONE native call:
 *
 *   rows  ->  (row, column, value) events  ->  NativeSketch.jobItemSimilarity  ->  top-k rows
 *
 * -s SIMILARITY_COSINE computes the exact measure (one counter column per matrix column), the sketch measure
 * (NativeSketchCosineSimilarity) the count-min cosine.  Every other measure is handed to the unmodified
 * RowSimilarityJob -- this class accelerates the cosine path and changes nothing else.  --maxObservationsPerRow /
 * --maxObservationsPerColumn are accepted; no down-sampling takes place (the dense contraction does not need it).
 *
 * NOT COMPILED IN THIS REPOSITORY (no JDK in the build image); see INTEGRATION.md.
 */
package org.apache.mahout.math.hadoop.similarity.cooccurrence;

import java.util.ArrayList;
import java.util.List;
import java.util.Map;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.IntWritable;
import org.apache.hadoop.io.SequenceFile;
import org.apache.hadoop.util.ToolRunner;
import org.apache.mahout.cf.taste.impl.common.NativeSketch;
import org.apache.mahout.common.AbstractJob;
import org.apache.mahout.common.Pair;
import org.apache.mahout.common.iterator.sequencefile.PathFilters;
import org.apache.mahout.common.iterator.sequencefile.PathType;
import org.apache.mahout.common.iterator.sequencefile.SequenceFileDirIterable;
import org.apache.mahout.math.RandomAccessSparseVector;
import org.apache.mahout.math.Vector;
import org.apache.mahout.math.VectorWritable;
import org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.NativeSketchCosineSimilarity;
import org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.VectorSimilarityMeasures;

public class NativeSketchRowSimilarityJob extends AbstractJob {

  public static void main(String[] args) throws Exception {
    ToolRunner.run(new NativeSketchRowSimilarityJob(), args);
  }

  /** the rows of a matrix flattened into the three event columns the native job takes */
  public static final class Events {
    public long[] row;      // dense row number 0..numRows-1 (position of the key among the sorted keys)
    public long[] key;      // column index (sketch measure) or dense column number (exact measure)
    public float[] value;
    public int[] rowKeys;   // dense row number -> the IntWritable key of the input
    public int numColumns;  // for the exact measure: number of distinct columns
  }

  @Override
  public int run(String[] args) throws Exception {
    addInputOption();
    addOutputOption();
    addOption("numberOfColumns", "r", "Number of columns in the input matrix", false);
    addOption("similarityClassname", "s", "Name of distributed similarity class to instantiate, alternatively use "
        + "one of the predefined similarities (" + VectorSimilarityMeasures.list() + ", SIMILARITY_SKETCH_COSINE)");
    addOption("maxSimilaritiesPerRow", "m", "Number of maximum similarities per row (default: 100)", "100");
    addOption("excludeSelfSimilarity", "ess", "compute similarity of rows to themselves?", String.valueOf(false));
    addOption("threshold", "tr", "discard row pairs with a similarity value below this", false);
    addOption("maxObservationsPerRow", null, "accepted for compatibility; no down-sampling takes place", "500");
    addOption("maxObservationsPerColumn", null, "accepted for compatibility; no down-sampling takes place", "500");
    addOption("randomSeed", null, "accepted for compatibility", false);
    addOption("sketchWidth", null, "width of the count-min sketches", String.valueOf(NativeSketchCosineSimilarity.DEFAULT_SKETCH_WIDTH));
    addOption("sketchDepth", null, "depth of the count-min sketches", String.valueOf(NativeSketchCosineSimilarity.DEFAULT_SKETCH_DEPTH));
    addOption("sketchSeed", null, "seed of the HashFunctionBuilder", String.valueOf(NativeSketchCosineSimilarity.DEFAULT_SKETCH_SEED));
    addOption("numGpus", null, "GPUs of this node to use (0 = all)", "0");
    addOption("precision", null, "rescored (similarities bit-equal to the Java loop), certified (exact top-k sets, "
        + "tensor-core values) or tensor", "rescored");

    Map<String, List<String>> parsedArgs = parseArguments(args);
    if (parsedArgs == null) {
      return -1;
    }
    String similarityClassname = getOption("similarityClassname");
    boolean sketch = NativeSketchCosineSimilarity.selects(similarityClassname);
    boolean exactCosine = "SIMILARITY_COSINE".equals(similarityClassname)
        || "org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.CosineSimilarity".equals(similarityClassname);
    if (!sketch && !exactCosine) {
      // not the path this class accelerates
      return ToolRunner.run(getConf(), new RowSimilarityJob(), args);
    }
    int maxSimilaritiesPerRow = Integer.parseInt(getOption("maxSimilaritiesPerRow"));
    boolean excludeSelfSimilarity = Boolean.parseBoolean(getOption("excludeSelfSimilarity"));
    double threshold = hasOption("threshold") ? Double.parseDouble(getOption("threshold")) : RowSimilarityJob.NO_THRESHOLD;
    if (!excludeSelfSimilarity) {
      throw new IllegalArgumentException("the native path excludes self similarity (ItemSimilarityJob's setting, "
          + "ItemSimilarityJob.java:157); use RowSimilarityJob for --excludeSelfSimilarity false");
    }

    Events ev = readRows(getInputPath(), getConf(), !sketch);
    int numRows = ev.rowKeys.length;
    long[] outIdx = new long[(long) numRows * maxSimilaritiesPerRow > Integer.MAX_VALUE ? 0 : numRows * maxSimilaritiesPerRow];
    if (outIdx.length == 0 && numRows > 0) {
      throw new IllegalArgumentException("rows x maxSimilaritiesPerRow exceeds a Java array; split the input");
    }
    double[] outSim = new double[outIdx.length];
    int[] outCnt = new int[numRows];
    long multi = NativeSketch.createMulti(Integer.parseInt(getOption("numGpus")));
    try {
      long[] stats = new long[NativeSketch.JOB_STATS_WORDS];
      NativeSketch.jobItemSimilarity(multi, ev.row, ev.key, ev.value, numRows, maxSimilaritiesPerRow,
          threshold == RowSimilarityJob.NO_THRESHOLD ? 0.0 : threshold,
          sketch ? Integer.parseInt(getOption("sketchWidth")) : Math.max(ev.numColumns, 1),
          sketch ? Integer.parseInt(getOption("sketchDepth")) : 1,
          Long.parseLong(getOption("sketchSeed")), !sketch, 1, NativeSketch.DTYPE_F16,
          NativeSketch.precisionOf(getOption("precision")), outIdx, outSim, outCnt, stats);
    } finally {
      NativeSketch.destroyMulti(multi);
    }
    writeRows(getOutputPath(), getConf(), ev.rowKeys, maxSimilaritiesPerRow, outIdx, outSim, outCnt);
    return 0;
  }

  /**
   * SequenceFile<IntWritable, VectorWritable> -> event columns.  Rows are numbered densely in ascending key
   * order (what TasteHadoopUtils.idToIndex keys look like after PreparePreferenceMatrixJob); for the exact measure
   * the columns are numbered densely as well, so that every column owns one counter.
   */
  static Events readRows(Path input, Configuration conf, boolean denseColumns) {
    List<Integer> keys = new ArrayList<Integer>();
    List<Vector> rows = new ArrayList<Vector>();
    long nnz = 0;
    for (Pair<IntWritable, VectorWritable> record
        : new SequenceFileDirIterable<IntWritable, VectorWritable>(input, PathType.LIST, PathFilters.partFilter(), conf)) {
      keys.add(record.getFirst().get());
      Vector v = record.getSecond().get();
      rows.add(v);
      nnz += v.getNumNondefaultElements();
    }
    Integer[] order = new Integer[keys.size()];
    for (int i = 0; i < order.length; i++) {
      order[i] = i;
    }
    final List<Integer> k = keys;
    java.util.Arrays.sort(order, new java.util.Comparator<Integer>() {
      @Override
      public int compare(Integer a, Integer b) {
        return k.get(a).compareTo(k.get(b));
      }
    });
    Events ev = new Events();
    ev.rowKeys = new int[order.length];
    ev.row = new long[(int) nnz];
    ev.key = new long[(int) nnz];
    ev.value = new float[(int) nnz];
    java.util.HashMap<Integer, Integer> columnNumber = new java.util.HashMap<Integer, Integer>();
    int at = 0;
    for (int r = 0; r < order.length; r++) {
      ev.rowKeys[r] = keys.get(order[r]);
      for (Vector.Element e : rows.get(order[r]).nonZeroes()) {
        long column = e.index();
        if (denseColumns) {
          Integer c = columnNumber.get(e.index());
          if (c == null) {
            c = columnNumber.size();
            columnNumber.put(e.index(), c);
          }
          column = c;
        }
        ev.row[at] = r;
        ev.key[at] = column;
        ev.value[at] = (float) e.get();
        at++;
      }
    }
    ev.numColumns = columnNumber.size();
    return ev;
  }

  /** top-k rows back into the input's key / index space, cardinality Integer.MAX_VALUE (Vectors.java:74) */
  static void writeRows(Path output, Configuration conf, int[] rowKeys, int k, long[] idx, double[] sim, int[] cnt)
    throws java.io.IOException {
    FileSystem fs = FileSystem.get(output.toUri(), conf);
    SequenceFile.Writer writer = SequenceFile.createWriter(fs, conf, new Path(output, "part-r-00000"),
        IntWritable.class, VectorWritable.class);
    try {
      IntWritable key = new IntWritable();
      VectorWritable value = new VectorWritable();
      for (int r = 0; r < rowKeys.length; r++) {
        if (cnt[r] == 0) {
          continue;   // the reducer never sees a row without similarities
        }
        Vector v = new RandomAccessSparseVector(Integer.MAX_VALUE, cnt[r]);
        for (int t = 0; t < cnt[r]; t++) {
          v.setQuick(rowKeys[(int) idx[r * k + t]], sim[r * k + t]);
        }
        key.set(rowKeys[r]);
        value.set(v);
        writer.append(key, value);
      }
    } finally {
      writer.close();
    }
  }
}
