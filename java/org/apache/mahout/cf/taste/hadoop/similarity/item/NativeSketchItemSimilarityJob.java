/*
 * `mahout itemsimilarity` with phase 1 on the GPUs: drop-in for ItemSimilarityJob
 * (cf/taste/hadoop/similarity/item/ItemSimilarityJob.java:97-179).  Same flags (:99-113) plus
 * --sketchWidth --sketchDepth --sketchSeed --numGpus --precision; same input (text `userID,itemID[,pref]`) and
 * output (`itemA<TAB>itemB<TAB>similarity`).
 *
 *   phase 0   PreparePreferenceMatrixJob, unmodified (:146-154): ITEMID_INDEX, USER_VECTORS, RATING_MATRIX,
 *             NUM_USERS in <tempDir>/prepareRatingMatrix
 *   phase 1   REPLACED (:156-172): the RATING_MATRIX (item index -> user vector) is read into event columns and
 *             handed to NativeSketchRowSimilarityJob's native call -- routing to the owner GPU, count-min
 *             sketch build, normalisation, all-pairs cosine with the top-k fused -- which writes
 *             <tempDir>/similarityMatrix in RowSimilarityJob's format
 *   phase 2   MostSimilarItemPairsMapper / Reducer, unmodified (:174-190): (minID, maxID) keys through the
 *             ITEMID_INDEX map, duplicates collapse, text output
 *
 * --startPhase / --endPhase keep their meaning, so the native phase also slots between phases run elsewhere.
 *
 * NOT COMPILED IN THIS REPOSITORY (no JDK in the build image); see INTEGRATION.md.  The C++ driver
 * mahout_b200/csrc/cli_itemsimilarity.cpp does the same three phases over the same C ABI and is what the tests run.
 */
package org.apache.mahout.cf.taste.hadoop.similarity.item;

import java.util.List;
import java.util.Map;
import java.util.concurrent.atomic.AtomicInteger;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.DoubleWritable;
import org.apache.hadoop.mapreduce.Job;
import org.apache.hadoop.mapreduce.lib.input.SequenceFileInputFormat;
import org.apache.hadoop.mapreduce.lib.output.TextOutputFormat;
import org.apache.hadoop.util.ToolRunner;
import org.apache.mahout.cf.taste.hadoop.EntityEntityWritable;
import org.apache.mahout.cf.taste.hadoop.preparation.PreparePreferenceMatrixJob;
import org.apache.mahout.common.AbstractJob;
import org.apache.mahout.math.hadoop.similarity.cooccurrence.NativeSketchRowSimilarityJob;
import org.apache.mahout.math.hadoop.similarity.cooccurrence.RowSimilarityJob;
import org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.NativeSketchCosineSimilarity;
import org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.VectorSimilarityMeasures;

public final class NativeSketchItemSimilarityJob extends AbstractJob {

  private static final int DEFAULT_MAX_SIMILAR_ITEMS_PER_ITEM = 100;   // ItemSimilarityJob.java:88
  private static final int DEFAULT_MAX_PREFS = 500;
  private static final int DEFAULT_MIN_PREFS_PER_USER = 1;

  public static void main(String[] args) throws Exception {
    ToolRunner.run(new NativeSketchItemSimilarityJob(), args);
  }

  @Override
  public int run(String[] args) throws Exception {
    addInputOption();
    addOutputOption();
    addOption("similarityClassname", "s", "SIMILARITY_SKETCH_COSINE (count-min sketch cosine), SIMILARITY_COSINE (exact) "
        + "or any of " + VectorSimilarityMeasures.list() + " (those run on the unmodified RowSimilarityJob)");
    addOption("maxSimilaritiesPerItem", "m", "try to cap the number of similar items per item to this number "
        + "(default: " + DEFAULT_MAX_SIMILAR_ITEMS_PER_ITEM + ')', String.valueOf(DEFAULT_MAX_SIMILAR_ITEMS_PER_ITEM));
    addOption("maxPrefs", "mppu", "accepted for compatibility; the native phase does not down-sample "
        + "(default: " + DEFAULT_MAX_PREFS + ')', String.valueOf(DEFAULT_MAX_PREFS));
    addOption("minPrefsPerUser", "mp", "ignore users with less preferences than this "
        + "(default: " + DEFAULT_MIN_PREFS_PER_USER + ')', String.valueOf(DEFAULT_MIN_PREFS_PER_USER));
    addOption("booleanData", "b", "Treat input as without pref values", String.valueOf(Boolean.FALSE));
    addOption("threshold", "tr", "discard item pairs with a similarity value below this", false);
    addOption("randomSeed", null, "accepted for compatibility", false);
    addOption("sketchWidth", null, "width of the count-min sketches", String.valueOf(NativeSketchCosineSimilarity.DEFAULT_SKETCH_WIDTH));
    addOption("sketchDepth", null, "depth of the count-min sketches", String.valueOf(NativeSketchCosineSimilarity.DEFAULT_SKETCH_DEPTH));
    addOption("sketchSeed", null, "seed of the HashFunctionBuilder", String.valueOf(NativeSketchCosineSimilarity.DEFAULT_SKETCH_SEED));
    addOption("numGpus", null, "GPUs of this node to use (0 = all)", "0");
    addOption("precision", null, "rescored, certified or tensor", "rescored");

    Map<String, List<String>> parsedArgs = parseArguments(args);
    if (parsedArgs == null) {
      return -1;
    }
    String similarityClassName = getOption("similarityClassname");
    int maxSimilarItemsPerItem = Integer.parseInt(getOption("maxSimilaritiesPerItem"));
    int minPrefsPerUser = Integer.parseInt(getOption("minPrefsPerUser"));
    boolean booleanData = Boolean.valueOf(getOption("booleanData"));
    double threshold = hasOption("threshold") ? Double.parseDouble(getOption("threshold")) : RowSimilarityJob.NO_THRESHOLD;

    Path similarityMatrixPath = getTempPath("similarityMatrix");
    Path prepPath = getTempPath("prepareRatingMatrix");
    AtomicInteger currentPhase = new AtomicInteger();

    if (shouldRunNextPhase(parsedArgs, currentPhase)) {
      ToolRunner.run(getConf(), new PreparePreferenceMatrixJob(), new String[] {
        "--input", getInputPath().toString(),
        "--output", prepPath.toString(),
        "--minPrefsPerUser", String.valueOf(minPrefsPerUser),
        "--booleanData", String.valueOf(booleanData),
        "--tempDir", getTempPath().toString(),
      });
    }

    if (shouldRunNextPhase(parsedArgs, currentPhase)) {
      // one native call instead of RowSimilarityJob's three MapReduce passes
      int rc = ToolRunner.run(getConf(), new NativeSketchRowSimilarityJob(), new String[] {
        "--input", new Path(prepPath, PreparePreferenceMatrixJob.RATING_MATRIX).toString(),
        "--output", similarityMatrixPath.toString(),
        "--similarityClassname", similarityClassName,
        "--maxSimilaritiesPerRow", String.valueOf(maxSimilarItemsPerItem),
        "--excludeSelfSimilarity", String.valueOf(Boolean.TRUE),
        "--threshold", String.valueOf(threshold),
        "--sketchWidth", getOption("sketchWidth"),
        "--sketchDepth", getOption("sketchDepth"),
        "--sketchSeed", getOption("sketchSeed"),
        "--numGpus", getOption("numGpus"),
        "--precision", getOption("precision"),
        "--tempDir", getTempPath().toString(),
      });
      if (rc != 0) {
        return -1;
      }
    }

    if (shouldRunNextPhase(parsedArgs, currentPhase)) {
      Job mostSimilarItems = prepareJob(similarityMatrixPath, getOutputPath(), SequenceFileInputFormat.class,
          ItemSimilarityJob.MostSimilarItemPairsMapper.class, EntityEntityWritable.class, DoubleWritable.class,
          ItemSimilarityJob.MostSimilarItemPairsReducer.class, EntityEntityWritable.class, DoubleWritable.class,
          TextOutputFormat.class);
      Configuration conf = mostSimilarItems.getConfiguration();
      conf.set(ItemSimilarityJob.ITEM_ID_INDEX_PATH_STR, new Path(prepPath, PreparePreferenceMatrixJob.ITEMID_INDEX).toString());
      conf.setInt(ItemSimilarityJob.MAX_SIMILARITIES_PER_ITEM, maxSimilarItemsPerItem);
      if (!mostSimilarItems.waitForCompletion(true)) {
        return -1;
      }
    }
    return 0;
  }
}
