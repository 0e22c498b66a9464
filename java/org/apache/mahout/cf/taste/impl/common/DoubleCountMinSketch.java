/*
 * Drop-in replacement of the fork's DoubleCountMinSketch
 * (mr/src/main/java/org/apache/mahout/cf/taste/impl/common/DoubleCountMinSketch.java): same package,
 * constructors, methods and exceptions; the counters live in HBM and every method is a native call.
 * NOT COMPILED HERE (no JDK in the build image) -- see INTEGRATION.md.
 */
package org.apache.mahout.cf.taste.impl.common;

import com.google.common.base.Preconditions;

public class DoubleCountMinSketch extends AbstractCountMinSketch implements AutoCloseable {

  private static final long CTX = NativeSketch.createContext(Integer.getInteger("mahout.b200.device", 0));
  private static final int FRAC_BITS = Integer.getInteger("mahout.b200.fracBits", 1);

  private long bank;

  public DoubleCountMinSketch(int width, int depth, HashFunctionBuilder hfBuilder) throws CMException {
    super(width, depth, hfBuilder);
    bank = allocate(hfBuilder);
  }

  public DoubleCountMinSketch(double delta, double epsilon, HashFunctionBuilder hfBuilder) throws CMException {
    super(delta, epsilon, hfBuilder);   // range checks + w = ceil(e/eps), d = ceil(ln(1/delta)) stay in Java
    bank = allocate(hfBuilder);
  }

  private long allocate(HashFunctionBuilder hfBuilder) {
    long[] a = new long[d];
    long[] b = new long[d];
    for (int i = 0; i < d; i++) {       // the builder stays the single source of the parameters
      a[i] = hfBuilder.getParamA(i);
      b[i] = hfBuilder.getParamB(i);
    }
    return NativeSketch.createBank(CTX, 1, d, w, a, b, FRAC_BITS);
  }

  /** C[i][h_i(key)] += increment for every row i (DoubleCountMinSketch.java:72-80). */
  public void update(long key, double increment) {
    insertedKeys.add(key);
    NativeSketch.updateOne(bank, 0, key, increment);
  }

  /** min_i C[i][h_i(key)] (DoubleCountMinSketch.java:94-103). */
  public double get(long key) {
    return NativeSketch.query(bank, 0, key);
  }

  /** min over rows of the per-row cosine, NaN if no row is comparable (:114-149). */
  public static double cosine(DoubleCountMinSketch a, DoubleCountMinSketch b) {
    Preconditions.checkArgument(a.w == b.w, "Widths of a (%s) and b (%s) must be the same", a.w, b.w);
    Preconditions.checkArgument(a.d == b.d, "Depths of a (%s) and b (%s) must be the same", a.d, b.d);
    return NativeSketch.cosine(a.bank, 0, b.bank, 0);
  }

  @Override
  public String toString() {
    double[] count = new double[w * d];
    NativeSketch.read(bank, 0, 1, count);
    StringBuilder builder = new StringBuilder();
    builder.append(System.lineSeparator());
    for (int i = 0; i < d; i++) {
      builder.append("| ");
      for (int j = 0; j < w; j++) {
        builder.append(count[j + i * w]);
        builder.append(" | ");
      }
      builder.append(System.lineSeparator());
    }
    return builder.toString();
  }

  @Override
  public void close() {
    if (bank != 0) {
      NativeSketch.destroyBank(bank);
      bank = 0;
    }
  }
}
