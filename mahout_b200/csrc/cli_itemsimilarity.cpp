// mahout_b200_itemsimilarity -- `mahout itemsimilarity` over libmahout_b200.so: the native host-side
// driver of the drop-in path (C++, C ABI only: what a Java ItemSimilarityJob would do through JNI).
//
// Mirrors cf/taste/hadoop/similarity/item/ItemSimilarityJob.java:97-232:
//   flags            --input/-i --output/-o --similarityClassname/-s --maxSimilaritiesPerItem/-m (100)
//                    --maxPrefs/-mppu (500, accepted; no down-sampling takes place) --minPrefsPerUser/-mp (1)
//                    --booleanData/-b --threshold/-tr --randomSeed --tempDir --startPhase --endPhase (:99-113)
//                    + --sketchWidth --sketchDepth --sketchSeed --fracBits --precision --device
//                    + --numGpus N (0 = every visible GPU): phase 1 runs as ONE mb200_job_item_similarity call over
//                      N GPUs of this process (items sharded by row mod N, events routed over NVLink)
//   phase 0          PreparePreferenceMatrixJob on the GPU (mb200_events_parse / mb200_events_prepare)
//   phase 1          RowSimilarityJob: K1 -> K2 -> K3 -> K5 (mb200_bank_update / mb200_bank_cosine_topk);
//                    -s SIMILARITY_COSINE is the exact measure (one counter column per user),
//                    -s SIMILARITY_SKETCH_COSINE the fork's count-min measure
//   phase 2          MostSimilarItemPairsMapper / Reducer (:181-232): (minID, maxID) keys, duplicates
//                    collapse, lines `itemA<TAB>itemB<TAB>similarity` ordered by (itemA, itemB)
// Exit status: 0 on success, -1 (255) on bad arguments or failure, like AbstractJob.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mahout_b200.h"

static const char* EXACT[] = {"SIMILARITY_COSINE",
                              "org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.CosineSimilarity"};
static const char* SKETCH[] = {"SIMILARITY_SKETCH_COSINE",
                               "org.apache.mahout.math.hadoop.similarity.cooccurrence.measures.NativeSketchCosineSimilarity"};

// Double.toString: shortest digits that round-trip; decimal notation for 1e-3 <= |v| < 1e7, otherwise
// "d.dddE[-]n"; always at least one digit behind the point.
static std::string java_double(double v) {
  if (v != v) return "NaN";
  if (v == 1.0 / 0.0) return "Infinity";
  if (v == -1.0 / 0.0) return "-Infinity";
  if (v == 0.0) return (1.0 / v < 0) ? "-0.0" : "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
  std::string s(buf, r.ptr);  // [-]d[.ddd]e[+-]xx
  const bool neg = s[0] == '-';
  if (neg) s.erase(0, 1);
  const size_t e = s.find('e');
  std::string digits = s.substr(0, e);
  const int exp10 = atoi(s.c_str() + e + 1);
  digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
  std::string out;
  if (exp10 >= -3 && exp10 < 7) {
    if (exp10 >= 0) {
      while ((int)digits.size() < exp10 + 1) digits.push_back('0');
      out = digits.substr(0, (size_t)exp10 + 1) + ".";
      const std::string frac = digits.substr((size_t)exp10 + 1);
      out += frac.empty() ? "0" : frac;
    } else {
      out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
  } else {
    out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
  }
  return neg ? "-" + out : out;
}

struct Args {
  std::string input, output, measure, precision = "rescored";
  int max_sim = 100, min_prefs = 1, width = 4096, depth = 4, frac_bits = 1, device = 0, num_gpus = -1;
  long long seed = 42;
  bool boolean_data = false, has_threshold = false;
  double threshold = 0.0;
};

static bool parse_args(int argc, char** argv, Args& a) {
  std::map<std::string, std::string> alias = {{"-i", "--input"},        {"-o", "--output"},      {"-s", "--similarityClassname"},
                                              {"-m", "--maxSimilaritiesPerItem"}, {"-mppu", "--maxPrefs"},
                                              {"-mp", "--minPrefsPerUser"}, {"-b", "--booleanData"}, {"-tr", "--threshold"}};
  for (int i = 1; i < argc; i++) {
    std::string k = argv[i];
    if (alias.count(k)) k = alias[k];
    if (i + 1 >= argc) {
      fprintf(stderr, "Missing value for option %s\n", k.c_str());
      return false;
    }
    const std::string v = argv[++i];
    if (k == "--input") a.input = v;
    else if (k == "--output") a.output = v;
    else if (k == "--similarityClassname") a.measure = v;
    else if (k == "--maxSimilaritiesPerItem") a.max_sim = atoi(v.c_str());
    else if (k == "--minPrefsPerUser") a.min_prefs = atoi(v.c_str());
    else if (k == "--booleanData") {
      std::string t = v;
      for (auto& c : t) c = (char)tolower(c);
      a.boolean_data = t == "true";  // Boolean.valueOf
    } else if (k == "--threshold") {
      a.threshold = atof(v.c_str());
      a.has_threshold = true;
    } else if (k == "--sketchWidth") a.width = atoi(v.c_str());
    else if (k == "--sketchDepth") a.depth = atoi(v.c_str());
    else if (k == "--sketchSeed") a.seed = atoll(v.c_str());
    else if (k == "--fracBits") a.frac_bits = atoi(v.c_str());
    else if (k == "--precision") a.precision = v;
    else if (k == "--device") a.device = atoi(v.c_str());
    else if (k == "--numGpus") a.num_gpus = atoi(v.c_str());
    else if (k == "--maxPrefs" || k == "--randomSeed" || k == "--tempDir" || k == "--startPhase" || k == "--endPhase") {
      // accepted for compatibility
    } else {
      fprintf(stderr, "Unexpected %s while processing Job-Specific Options\n", k.c_str());
      return false;
    }
  }
  if (a.input.empty() || a.output.empty() || a.measure.empty()) {
    fprintf(stderr, "Missing required option: --input, --output and --similarityClassname are required\n");
    return false;
  }
  return true;
}

#define TRY(call)                                                         \
  do {                                                                    \
    int rc_ = (call);                                                     \
    if (rc_ != MB200_OK) {                                                \
      fprintf(stderr, "itemsimilarity failed: %s\n", mb200_last_error(ctx)); \
      return -1;                                                          \
    }                                                                     \
  } while (0)

int main(int argc, char** argv) {
  Args a;
  if (!parse_args(argc, argv, a)) return -1;
  bool exact = false, known = false;
  for (const char* m : EXACT)
    if (a.measure == m) exact = known = true;
  for (const char* m : SKETCH)
    if (a.measure == m) known = true;
  if (!known) {
    fprintf(stderr, "itemsimilarity: the native path implements the cosine measure only (got %s)\n", a.measure.c_str());
    return -1;
  }
  if (a.max_sim <= 0) {
    fprintf(stderr, "maxSimilarItemsPerItem must be greater then 0!\n");
    return -1;
  }
  if (a.precision != "rescored" && a.precision != "tensor" && a.precision != "certified") {
    fprintf(stderr, "--precision must be rescored, certified or tensor\n");
    return -1;
  }
  const int precision = a.precision == "tensor" ? MB200_PRECISION_TENSOR
                        : a.precision == "certified" ? MB200_PRECISION_CERTIFIED : MB200_PRECISION_RESCORED;
  mb200_ctx* ctx = nullptr;
  mb200_multi* multi = nullptr;
  if (a.num_gpus >= 0) {
    // --numGpus: one process, N GPUs; the preparation phase runs on the first of them
    if (mb200_create_multi(a.num_gpus, nullptr, &multi) != MB200_OK) {
      fprintf(stderr, "itemsimilarity failed: %s\n", mb200_last_error(nullptr));
      return -1;
    }
    TRY(mb200_multi_ctx(multi, 0, &ctx));
  } else {
    TRY(mb200_create(a.device, &ctx));
  }

  // the input goes to page-locked memory, from there to the device once
  FILE* f = fopen(a.input.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "itemsimilarity failed: cannot open %s\n", a.input.c_str());
    return -1;
  }
  fseek(f, 0, SEEK_END);
  const long long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  void* text = nullptr;
  TRY(mb200_host_alloc(size > 0 ? size : 1, &text));
  const size_t got = size > 0 ? fread(text, 1, (size_t)size, f) : 0;
  fclose(f);
  mb200_events* ev = nullptr;
  TRY(mb200_events_parse(ctx, (const char*)text, (int64_t)got, MB200_MEM_HOST, a.boolean_data, 0.0f, 0, &ev));
  mb200_host_free(text);
  mb200_prefs* pm = nullptr;
  TRY(mb200_events_prepare(ev, a.min_prefs, &pm));
  mb200_events_destroy(ev);
  int64_t n = 0, num_items = 0, num_users = 0;
  TRY(mb200_prefs_info(pm, &n, &num_items, &num_users));

  std::vector<std::pair<std::pair<int64_t, int64_t>, double>> lines;
  if (num_items > 0 && n > 0) {
    std::vector<int64_t> item_id((size_t)num_items);
    TRY(mb200_prefs_tables(pm, item_id.data(), nullptr));
    int64_t *row = nullptr, *user = nullptr, *ucol = nullptr;
    float* pref = nullptr;
    TRY(mb200_prefs_columns(pm, &row, &user, &pref));
    TRY(mb200_prefs_user_columns(pm, &ucol));
    const int k = a.max_sim;
    std::vector<int64_t> idx((size_t)num_items * k);
    std::vector<double> sim((size_t)num_items * k);
    std::vector<int32_t> cnt((size_t)num_items);
    if (multi) {
      // phase 1 as one call over all GPUs: the prepared events go back to the host columns the job takes
      std::vector<int64_t> h_row((size_t)n), h_key((size_t)n);
      std::vector<float> h_pref((size_t)n);
      TRY(mb200_prefs_read(pm, h_row.data(), exact ? nullptr : h_key.data(), exact ? h_key.data() : nullptr, h_pref.data()));
      mb200_job_params jp;
      memset(&jp, 0, sizeof(jp));
      const int64_t one = 1, zero = 0;
      jp.k = k;
      jp.threshold = a.has_threshold ? a.threshold : 0.0;
      jp.width = exact ? (int32_t)std::max<int64_t>(num_users, 1) : a.width;
      jp.depth = exact ? 1 : a.depth;
      jp.seed = a.seed;
      jp.hash_a = exact ? &one : nullptr;
      jp.hash_b = exact ? &zero : nullptr;
      jp.frac_bits = a.frac_bits;
      jp.dtype = MB200_DTYPE_F16;
      jp.precision = precision;
      mb200_job_stats st;
      if (mb200_job_item_similarity(multi, h_row.data(), h_key.data(), h_pref.data(), n, num_items, &jp, idx.data(), sim.data(),
                                    cnt.data(), &st) != MB200_OK) {
        fprintf(stderr, "itemsimilarity failed: %s\n", mb200_multi_last_error(multi));
        return -1;
      }
      // RowSimilarityJob.Counters (RowSimilarityJob.java:84), as far as they exist on this path
      fprintf(stderr, "ROWS=%lld USED_OBSERVATIONS=%lld SIMILARITIES=%lld gpus=%d fallback_rows=%lld route=%.3fs build=%.3fs cosine=%.3fs\n",
              (long long)st.rows, (long long)st.events, (long long)st.similarities_kept, st.n_gpus, (long long)st.fallback_rows,
              st.route_s, st.build_s, st.cosine_s);
    } else {
    mb200_bank* bank = nullptr;
    if (exact) {
      const int64_t one = 1, zero = 0;
      TRY(mb200_bank_create_params(ctx, num_items, 1, (int32_t)std::max<int64_t>(num_users, 1), &one, &zero, a.frac_bits, &bank));
      TRY(mb200_bank_update(bank, row, ucol, pref, n, MB200_MEM_DEVICE));
    } else {
      TRY(mb200_bank_create(ctx, num_items, a.depth, a.width, a.seed, a.frac_bits, &bank));
      TRY(mb200_bank_update(bank, row, user, pref, n, MB200_MEM_DEVICE));
    }
    TRY(mb200_bank_check(bank));
    TRY(mb200_bank_cosine_topk(bank, k, a.has_threshold ? a.threshold : 0.0, 1, MB200_DTYPE_F16, precision, idx.data(),
                               sim.data(), cnt.data(), MB200_MEM_HOST));
    mb200_bank_destroy(bank);
    }
    // MostSimilarItemPairsMapper / Reducer
    std::map<std::pair<int64_t, int64_t>, double> pairs;
    for (int64_t r = 0; r < num_items; r++)
      for (int t = 0; t < cnt[(size_t)r]; t++) {
        const int64_t c = idx[(size_t)r * k + t];
        const int64_t ra = item_id[(size_t)r], ca = item_id[(size_t)c];
        pairs.emplace(ra < ca ? std::make_pair(ra, ca) : std::make_pair(ca, ra), sim[(size_t)r * k + t]);
      }
    lines.assign(pairs.begin(), pairs.end());
  }
  mb200_prefs_destroy(pm);
  FILE* out = fopen(a.output.c_str(), "w");
  if (!out) {
    fprintf(stderr, "itemsimilarity failed: cannot write %s\n", a.output.c_str());
    return -1;
  }
  for (auto& l : lines)
    fprintf(out, "%lld\t%lld\t%s\n", (long long)l.first.first, (long long)l.first.second, java_double(l.second).c_str());
  fclose(out);
  if (multi) mb200_multi_destroy(multi);
  else mb200_destroy(ctx);
  return 0;
}
