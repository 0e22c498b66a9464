/*
 * mahout_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the arithmetic of jalhajj/mahout's sketch-similarity
 * hot path, written from the reference's behaviour (file:line cites below are
 * into /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libmahout_b200.so) never links or calls it.
 *
 * PARITY PINNING.  The reference cannot be built here (no JVM, no jni.h, no
 * Maven repository), so `oracle/_ref` does not exist.  What the
 * reference's own tests pin is pinned (tests/test_oracle.py):
 *   - VectorSimilarityMeasuresTest.testCosineSimilarity  -> 0.769846046
 *   - ItemSimilarityJobTest.testCompleteJob              -> (1,3,~0.45) (2,3,~0.89)
 *   - ItemSimilarityJobTest.testMostSimilarItemsPairsMapper / Reducer
 *   - TasteHadoopUtilsTest (idToIndex range)
 * The fork's sketch classes (HashFunction*, *CountMinSketch, CosineCM) have NO
 * reference tests: for them parity is UNPINNED by the reference and rests on
 * the Java SE specification of java.util.Random / BigInteger (known answers
 * new Random(42).nextInt() == -1170105035, nextLong() == -5025562857975149833)
 * plus the collision-free cross-check against the pinned exact cosine.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __int128 i128;

/* ------------------------------------------------------------------------ */
/* java.util.Random (Java SE spec; used at HashFunctionBuilder.java:27,46-47) */
/* ------------------------------------------------------------------------ */
#define JR_MULT 0x5DEECE66DLL
#define JR_MASK ((1LL << 48) - 1)

static int64_t jr_scramble(int64_t seed) { return (seed ^ JR_MULT) & JR_MASK; }

static int32_t jr_next(int64_t *state, int bits) {
  *state = (int64_t)(((uint64_t)*state * (uint64_t)JR_MULT + 0xBULL) & (uint64_t)JR_MASK);
  return (int32_t)(*state >> (48 - bits));
}

static int64_t jr_next_long(int64_t *state) {
  int64_t hi = (int64_t)jr_next(state, 32);
  int64_t lo = (int64_t)jr_next(state, 32);
  return (int64_t)(((uint64_t)hi << 32) + (uint64_t)lo);
}

/* Math.abs(long): Long.MIN_VALUE stays negative (Java SE spec). */
static int64_t java_abs_long(int64_t v) {
  return v < 0 ? (int64_t)(0ULL - (uint64_t)v) : v;
}

int32_t orc_java_random_next_int(int64_t seed) {
  int64_t s = jr_scramble(seed);
  return jr_next(&s, 32);
}

int64_t orc_java_random_next_long(int64_t seed) {
  int64_t s = jr_scramble(seed);
  return jr_next_long(&s);
}

/* ------------------------------------------------------------------------ */
/* HashFunctionBuilder(seed).getHashFunction(i, w)  HashFunctionBuilder.java:23-28,40-60
 * For iteration index i the builder draws ra = abs(nextLong()), rb = abs(nextLong())
 * in that order from ONE Random(seed); parameters are shared by every sketch
 * built from the same builder.                                               */
/* ------------------------------------------------------------------------ */
#define ORC_PRIME 9223372036854775783LL /* HashFunctionBuilder.java:24 (2^63-25) */

void orc_hash_params(int64_t seed, int depth, int64_t *a, int64_t *b) {
  int64_t s = jr_scramble(seed);
  for (int i = 0; i < depth; i++) {
    a[i] = java_abs_long(jr_next_long(&s));
    b[i] = java_abs_long(jr_next_long(&s));
  }
}

/* HashFunction.hash(key)  HashFunction.java:31-34
 *   a.multiply(k).add(b).mod(bigPrime).mod(w).intValue()
 * BigInteger.mod always returns a non-negative value.                        */
int32_t orc_hash(int64_t a, int64_t b, int32_t w, int64_t key) {
  i128 v = (i128)a * (i128)key + (i128)b;
  i128 r = v % (i128)ORC_PRIME;
  if (r < 0) r += ORC_PRIME;
  return (int32_t)(r % (i128)w);
}

void orc_hash_many(int64_t a, int64_t b, int32_t w, const int64_t *keys, int64_t n, int32_t *out) {
  for (int64_t t = 0; t < n; t++) out[t] = orc_hash(a, b, w, keys[t]);
}

/* ------------------------------------------------------------------------ */
/* AbstractCountMinSketch(delta, epsilon)  AbstractCountMinSketch.java:69-83
 * returns 0 ok, 1 bad delta, 2 bad epsilon (CMException in the reference).   */
/* ------------------------------------------------------------------------ */
int orc_cm_dims(double delta, double epsilon, int32_t *width, int32_t *depth) {
  if (delta <= 0 || delta > exp(-1.0)) return 1;
  if (epsilon <= 0 || epsilon > exp(1.0)) return 2;
  *width = (int32_t)ceil(exp(1.0) / epsilon);
  *depth = (int32_t)ceil(log(1.0 / delta));
  return 0;
}

/* ------------------------------------------------------------------------ */
/* DoubleCountMinSketch: counters row-major count[j + i*w]
 * (DoubleCountMinSketch.java:62-64).                                         */
/* ------------------------------------------------------------------------ */

/* update(key, inc)  DoubleCountMinSketch.java:72-80 (sequential, in call order) */
void orc_cm_update(double *count, int32_t w, int32_t d, const int64_t *a, const int64_t *b,
                   const int64_t *keys, const double *incs, int64_t n) {
  for (int64_t t = 0; t < n; t++) {
    for (int i = 0; i < d; i++) {
      int32_t j = orc_hash(a[i], b[i], w, keys[t]);
      count[(int64_t)j + (int64_t)i * w] += incs[t];
    }
  }
}

/* get(key)  DoubleCountMinSketch.java:94-103 */
double orc_cm_get(const double *count, int32_t w, int32_t d, const int64_t *a, const int64_t *b,
                  int64_t key) {
  double estimate = DBL_MAX;
  for (int i = 0; i < d; i++) {
    int32_t j = orc_hash(a[i], b[i], w, key);
    double v = count[(int64_t)j + (int64_t)i * w];
    if (v < estimate) estimate = v;
  }
  return estimate;
}

/* cosine(a, b)  DoubleCountMinSketch.java:114-149
 * per row: AA, BB, AB summed in index order; rows with zero denominator are
 * skipped; min over rows; NaN when no row qualified.                         */
double orc_cm_cosine(const double *ca, const double *cb, int32_t w, int32_t d) {
  double min_cos = DBL_MAX;
  for (int i = 0; i < d; i++) {
    double va = 0.0, vb = 0.0, vab = 0.0;
    const double *ra = ca + (int64_t)i * w, *rb = cb + (int64_t)i * w;
    for (int j = 0; j < w; j++) {
      double xa = ra[j], xb = rb[j];
      va += xa * xa;
      vb += xb * xb;
      vab += xa * xb;
    }
    double den = sqrt(va) * sqrt(vb);
    if (den != 0) {
      double c = vab / den;
      min_cos = c < min_cos ? c : min_cos; /* Math.min; NaN cannot occur here */
    }
  }
  if (min_cos == DBL_MAX) return NAN;
  return min_cos;
}

/* AbstractSimilarity.normalizeWeightResult (unweighted)  AbstractSimilarity.java:313-330
 * as applied by CosineCM.userSimilarity (CosineCM.java:90-93).                */
double orc_clamp_similarity(double r) {
  if (isnan(r)) return r;
  if (r < -1.0) return -1.0;
  if (r > 1.0) return 1.0;
  return r;
}

/* ------------------------------------------------------------------------ */
/* Sketch bank: one DoubleCountMinSketch per entity, C[e][i][j], all built from
 * one HashFunctionBuilder (CosineCM.java:41-58 builds one sketch per entity
 * with cm.update(key, pref) in array order; pref is a float widened to double,
 * GenericUserPreferenceArray.java:54,141 / ToEntityPrefsMapper.java:73).
 * entity == NULL means a single sketch (E == 1).                              */
/* ------------------------------------------------------------------------ */
void orc_bank_update(double *bank, int64_t E, int32_t d, int32_t w, const int64_t *a,
                     const int64_t *b, const int64_t *entity, const int64_t *key, const float *inc,
                     int64_t n) {
  (void)E;
  for (int64_t t = 0; t < n; t++) {
    int64_t e = entity ? entity[t] : 0;
    double *c = bank + e * (int64_t)d * w;
    double x = (double)inc[t];
    for (int i = 0; i < d; i++) {
      int32_t j = orc_hash(a[i], b[i], w, key[t]);
      c[(int64_t)j + (int64_t)i * w] += x;
    }
  }
}

/* Same arithmetic, all host threads: events are split into `nthreads` slices,
 * each slice accumulates into a private bank which are then summed.  Equal to
 * the sequential result whenever every partial sum is exactly representable
 * (the precondition the product checks on ingest).  Only for E*d*w small
 * enough to replicate (the single-sketch benchmark configuration).            */
int orc_bank_update_mt(double *bank, int64_t E, int32_t d, int32_t w, const int64_t *a,
                       const int64_t *b, const int64_t *entity, const int64_t *key,
                       const float *inc, int64_t n, int nthreads) {
  int64_t cells = E * (int64_t)d * w;
  if (nthreads <= 1) {
    orc_bank_update(bank, E, d, w, a, b, entity, key, inc, n);
    return 1;
  }
  double *priv = (double *)calloc((size_t)cells * (size_t)nthreads, sizeof(double));
  if (!priv) return -1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int t = 0; t < nthreads; t++) {
    int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
    orc_bank_update(priv + (int64_t)t * cells, E, d, w, a, b, entity ? entity + lo : NULL, key + lo,
                    inc + lo, hi - lo);
  }
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t c = 0; c < cells; c++) {
    double s = bank[c];
    for (int t = 0; t < nthreads; t++) s += priv[(int64_t)t * cells + c];
    bank[c] = s;
  }
  free(priv);
  return nthreads;
}

void orc_bank_query(const double *bank, int32_t d, int32_t w, const int64_t *a, const int64_t *b,
                    const int64_t *entity, const int64_t *key, int64_t n, double *out) {
  for (int64_t t = 0; t < n; t++) {
    int64_t e = entity ? entity[t] : 0;
    out[t] = orc_cm_get(bank + e * (int64_t)d * w, w, d, a, b, key[t]);
  }
}

/* ------------------------------------------------------------------------ */
/* Top-k with the total order (similarity desc, index asc).
 * Reference: TopElementsQueue.java:26-59 + Vectors.topKElements (Vectors.java:58-79)
 * + RowSimilarityJob.UnsymmetrifyMapper (RowSimilarityJob.java:515-539): a
 * candidate enters only if candidate > heap-min, sentinels hold
 * Double.MIN_VALUE, so similarities <= Double.MIN_VALUE are never reported.
 * The reference breaks ties by hash-map iteration order (unspecified); the
 * north star fixes it: lower index wins.                                      */
/* ------------------------------------------------------------------------ */
typedef struct {
  double sim;
  int64_t idx;
} orc_elem;

static int elem_better(const orc_elem *x, const orc_elem *y) {
  if (x->sim != y->sim) return x->sim > y->sim;
  return x->idx < y->idx;
}

static int elem_cmp_desc(const void *p, const void *q) {
  const orc_elem *x = (const orc_elem *)p, *y = (const orc_elem *)q;
  if (elem_better(x, y)) return -1;
  if (elem_better(y, x)) return 1;
  return 0;
}

/* keep the k best of `cand` (n entries, modified in place); returns count */
static int64_t select_topk(orc_elem *cand, int64_t n, int k) {
  qsort(cand, (size_t)n, sizeof(orc_elem), elem_cmp_desc);
  return n < k ? n : k;
}

#define JAVA_DOUBLE_MIN_VALUE 4.9e-324 /* Double.MIN_VALUE == RowSimilarityJob.NO_THRESHOLD (:56) */

/* SimilarityReducer + top-k filter for one similarity value
 * (RowSimilarityJob.java:489-499): keep iff sim >= threshold; NaN never passes;
 * then the queue admits it only if > Double.MIN_VALUE.                        */
static int sim_admitted(double s, double threshold) {
  if (isnan(s)) return 0;
  if (!(s >= threshold)) return 0;
  return s > JAVA_DOUBLE_MIN_VALUE;
}

/* All-pairs sketch cosine + per-row top-k for rows [r0, r1) of a bank.
 * sim(r,c) = DoubleCountMinSketch.cosine(bank[r], bank[c]); the column index
 * reported is the entity index.  out_idx/out_sim are [(r1-r0)][k], out_cnt
 * [(r1-r0)].  Unused slots: idx -1, sim 0.                                     */
void orc_bank_cosine_topk(const double *bank, int64_t E, int32_t d, int32_t w, int64_t r0,
                          int64_t r1, int k, double threshold, int exclude_self, int nthreads,
                          int64_t *out_idx, double *out_sim, int32_t *out_cnt) {
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int64_t r = r0; r < r1; r++) {
    orc_elem *cand = (orc_elem *)malloc(sizeof(orc_elem) * (size_t)(E > 0 ? E : 1));
    int64_t n = 0;
    const double *cr = bank + r * (int64_t)d * w;
    for (int64_t c = 0; c < E; c++) {
      if (exclude_self && c == r) continue; /* similarities.setQuick(row, 0)  RowSimilarityJob.java:497-499 */
      double s = orc_cm_cosine(cr, bank + c * (int64_t)d * w, w, d);
      if (sim_admitted(s, threshold)) {
        cand[n].sim = s;
        cand[n].idx = c;
        n++;
      }
    }
    int64_t m = select_topk(cand, n, k);
    int64_t o = (r - r0) * k;
    for (int64_t t = 0; t < k; t++) {
      out_idx[o + t] = t < m ? cand[t].idx : -1;
      out_sim[o + t] = t < m ? cand[t].sim : 0.0;
    }
    out_cnt[r - r0] = (int32_t)m;
    free(cand);
  }
}

/* Dense cosine block sim[r][c] for r in [r0,r1), c in [0,E) (NaN preserved). */
void orc_bank_cosine_dense(const double *bank, int64_t E, int32_t d, int32_t w, int64_t r0,
                           int64_t r1, int nthreads, double *out) {
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int64_t r = r0; r < r1; r++)
    for (int64_t c = 0; c < E; c++)
      out[(r - r0) * E + c] =
          orc_cm_cosine(bank + r * (int64_t)d * w, bank + c * (int64_t)d * w, w, d);
}

/* ------------------------------------------------------------------------ */
/* Exact path: RowSimilarityJob with CosineSimilarity over a sparse row matrix
 * given in CSR form (rows = items, columns = users, values = float prefs).
 *   normalize: row / ||row||_2   CosineSimilarity.java:24-27, AbstractVector.java:208-210
 *   aggregate: a*b, similarity = sum   CosineSimilarity.java:34-42
 *   keep iff >= threshold, zero the diagonal   RowSimilarityJob.java:489-499
 *   per-row top-k over the symmetrised row   RowSimilarityJob.java:515-559
 * Down-sampling (RowSimilarityJob.java:288-315) is NOT restated: parity runs
 * use maxPrefs >= every row/column count, where the sample rate is exactly 1.
 * Dots are accumulated over shared columns in ascending column order.         */
/* ------------------------------------------------------------------------ */
void orc_rowsim_cosine_topk(int64_t nrows, int64_t ncols, const int64_t *rowptr,
                            const int32_t *colidx, const float *vals, int k, double threshold,
                            int exclude_self, int nthreads, int64_t *out_idx, double *out_sim,
                            int32_t *out_cnt) {
  if (nthreads < 1) nthreads = 1;
  int64_t nnz = rowptr[nrows];
  double *nv = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  for (int64_t r = 0; r < nrows; r++) {
    double ss = 0.0;
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) ss += (double)vals[p] * (double)vals[p];
    double nrm = sqrt(ss);
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) nv[p] = (double)vals[p] / nrm;
  }
#pragma omp parallel num_threads(nthreads)
  {
    double *dense = (double *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(double));
    orc_elem *cand = (orc_elem *)malloc(sizeof(orc_elem) * (size_t)(nrows > 0 ? nrows : 1));
#pragma omp for schedule(dynamic, 8)
    for (int64_t r = 0; r < nrows; r++) {
      for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) dense[colidx[p]] = nv[p];
      int64_t n = 0;
      for (int64_t c = 0; c < nrows; c++) {
        if (exclude_self && c == r) continue;
        double dot = 0.0;
        int any = 0;
        for (int64_t p = rowptr[c]; p < rowptr[c + 1]; p++) {
          double x = dense[colidx[p]];
          if (x != 0.0) {
            dot += x * nv[p];
            any = 1;
          }
        }
        /* pairs with no co-occurring column never reach SimilarityReducer */
        if (any && sim_admitted(dot, threshold)) {
          cand[n].sim = dot;
          cand[n].idx = c;
          n++;
        }
      }
      int64_t m = select_topk(cand, n, k);
      for (int64_t t = 0; t < k; t++) {
        out_idx[r * k + t] = t < m ? cand[t].idx : -1;
        out_sim[r * k + t] = t < m ? cand[t].sim : 0.0;
      }
      out_cnt[r] = (int32_t)m;
      for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) dense[colidx[p]] = 0.0;
    }
    free(dense);
    free(cand);
  }
  free(nv);
}

/* Plain cosine of two dense vectors the way VectorSimilarityMeasuresTest does it
 * (normalize both, sum products where both non-zero)
 * VectorSimilarityMeasuresTest.java:43-61.                                    */
double orc_exact_cosine(const double *x, const double *y, int64_t n) {
  double sx = 0.0, sy = 0.0;
  for (int64_t i = 0; i < n; i++) {
    sx += x[i] * x[i];
    sy += y[i] * y[i];
  }
  double nx = sqrt(sx), ny = sqrt(sy), dot = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double a = x[i] / nx, b = y[i] / ny;
    if (a != 0 && b != 0) dot += a * b;
  }
  return dot;
}

/* TasteHadoopUtils.idToIndex  TasteHadoopUtils.java:56-58
 *   0x7FFFFFFF & Longs.hashCode(id) % 0x7FFFFFFE   ('%' binds tighter than '&';
 *   Longs.hashCode(v) = (int)(v ^ (v >>> 32)); Java '%' truncates toward zero) */
int32_t orc_id_to_index(int64_t id) {
  int32_t h = (int32_t)((uint64_t)id ^ ((uint64_t)id >> 32));
  int32_t m = (int32_t)((int64_t)h % (int64_t)0x7FFFFFFE);
  return (int32_t)(0x7FFFFFFF & (uint32_t)m);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
