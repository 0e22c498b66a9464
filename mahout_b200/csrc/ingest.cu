// ingest.cu -- the step in front of the sketch path: text preference data ->
// device-resident event columns, and PreparePreferenceMatrixJob's bookkeeping on the GPU.
//
// Reference semantics:
//   ToEntityPrefsMapper.map (cf/taste/hadoop/ToEntityPrefsMapper.java:56-76): split on [\t,];
//     userID = Long.parseLong(tokens[0]), itemID = Long.parseLong(tokens[1]); pref =
//     tokens.length > 2 ? Float.parseFloat(tokens[2]) + ratingShift : 1.0f; booleanData ignores it.
//   TasteHadoopUtils.idToIndex (TasteHadoopUtils.java:56-58): 0x7FFFFFFF & Longs.hashCode(id) % 0x7FFFFFFE.
//   ItemIDIndexMapper / ItemIDIndexReducer (item/ItemIDIndex*.java): index -> minimum itemID, over all lines.
//   ToUserVectorsReducer.reduce (item/ToUserVectorsReducer.java:66-82): userVector.set(index, pref) -- the
//     last preference of a (user, index) pair wins; users with fewer than minPrefsPerUser entries are dropped.
//
// Kernels (all HBM-bound byte / integer work):
//   k_count_lines / k_parse_lines  tiles of 4 KB staged through shared memory with 16-byte loads; one
//                                  thread per line start parses the line; line -> output slot by a
//                                  two-level prefix sum, so the output keeps the input order
//   k_prep_insert / k_prep_mark / k_prep_flags   open-addressing hash tables in HBM keyed by user,
//                                  (user, index) and index: last-wins de-duplication (atomicMax of the
//                                  position), per-user counts, minimum itemID per index (atomicMin)
//   k_compact                      order-preserving compaction of the surviving events
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct mb200_events {
  mb200_ctx* ctx = nullptr;
  int64_t n = 0;
  long long* user = nullptr;
  long long* item = nullptr;
  float* pref = nullptr;
};

struct mb200_prefs {
  mb200_ctx* ctx = nullptr;
  int64_t n = 0, num_items = 0, num_users = 0;
  long long* row = nullptr;   // dense row number of the event's item index
  long long* user = nullptr;
  long long* ucol = nullptr;  // dense number 0..num_users-1 of the event's user (arbitrary but fixed order)
  float* pref = nullptr;
  std::vector<int64_t> item_id;       // row -> minimum itemID with that index
  std::vector<int32_t> index_values;  // row -> idToIndex value (ascending)
};

static void events_free(mb200_events* e) {
  if (!e) return;
  cudaFree(e->user);
  cudaFree(e->item);
  cudaFree(e->pref);
  delete e;
}

static void prefs_free(mb200_prefs* p) {
  if (!p) return;
  cudaFree(p->row);
  cudaFree(p->user);
  cudaFree(p->ucol);
  cudaFree(p->pref);
  delete p;
}

// ------------------------------------------------------------------------------------------------
// text -> events
// ------------------------------------------------------------------------------------------------
static constexpr int PT = 256;          // threads per CTA
static constexpr int PB = 16;           // bytes per thread
static constexpr int TILE = PT * PB;    // bytes per tile
static constexpr int LOOK = 496;        // look-ahead staged behind the tile (longer lines read global memory)

enum { PERR_NONE = 0, PERR_LONG = 1, PERR_TOKENS = 2, PERR_FLOAT = 3, PERR_FIXUPS = 4 };

struct ParseArgs {
  const unsigned char* text;
  long long bytes;
  long long num_tiles;
  int boolean_data, transpose;
  float rating_shift;
  const long long* tile_off;     // exclusive prefix sum of the tile line counts
  long long* user;
  long long* item;
  float* pref;
  unsigned long long* err_pos;   // smallest byte offset of a malformed line (ULLONG_MAX = none)
  int* err_code;
  // float tokens the device parser does not decide (long digit strings, exponents, hex, NaN...)
  long long* fix_line;           // output slot
  long long* fix_pos;            // byte offset of the token
  int* fix_len;
  int* fix_count;
  int fix_cap;
};

// One staged tile: s[o] = byte base + o for o in [0, TILE + LOOK), '\n' beyond the end of the buffer.
struct Tile {
  const unsigned char* g;
  long long n, base;
  const unsigned char* s;
  // byte at tile offset o >= 0; a virtual '\n' sits behind the end of the buffer
  __device__ __forceinline__ unsigned char at(int o) const {
    if (o < TILE + LOOK) return s[o];
    const long long pos = base + o;
    return pos < n ? __ldg(g + pos) : (unsigned char)'\n';
  }
  __device__ __forceinline__ bool eol(int o) const {
    const unsigned char c = at(o);
    return c == '\n' || (c == '\r' && at(o + 1) == '\n');
  }
};

// Stage bytes [base, base + TILE + LOOK) with 16-byte loads and stores; *s_prev = byte base - 1.
__device__ __forceinline__ void stage_tile(const unsigned char* g, long long n, long long base, unsigned char* s,
                                           unsigned char* s_prev) {
  if (threadIdx.x == 0) *s_prev = base > 0 ? __ldg(g + base - 1) : (unsigned char)'\n';
  const long long end = base + TILE + LOOK < n ? base + TILE + LOOK : n;
  const int have = (int)(end - base);
  int done = 0;
  if ((((uintptr_t)(g + base)) & 15) == 0) {
    const int vec = have >> 4;
    for (int v = threadIdx.x; v < vec; v += PT)
      reinterpret_cast<uint4*>(s)[v] = __ldg(reinterpret_cast<const uint4*>(g + base) + v);
    done = vec << 4;
  }
  for (int i = done + threadIdx.x; i < have; i += PT) s[i] = __ldg(g + base + i);
  for (int i = have + threadIdx.x; i < TILE + LOOK; i += PT) s[i] = '\n';
  __syncthreads();
}

// bit j of the result <=> byte j of the 16 bytes equals c
__device__ __forceinline__ unsigned eq_mask16(const uint4& x, unsigned c4) {
  const unsigned w[4] = {x.x, x.y, x.z, x.w};
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned t = __vcmpeq4(w[k], c4) & 0x01010101u;   // bits 0, 8, 16, 24
    m |= ((t * 0x10204080u) >> 28) << (4 * k);             // gathered into one nibble
  }
  return m;
}

// bit j <=> a non-empty line starts at byte j of this thread's 16 bytes.  A line starts behind a '\n'
// (or at offset 0 of the buffer) unless it is empty: "\n" or "\r\n" follow immediately.  Positions
// behind the end of the buffer hold '\n' padding and can never start a line.
__device__ __forceinline__ unsigned line_starts16(const unsigned char* s, unsigned char prev_byte) {
  const int o = threadIdx.x * PB;
  const uint4 x = *reinterpret_cast<const uint4*>(s + o);
  const unsigned nl = eq_mask16(x, 0x0A0A0A0Au), cr = eq_mask16(x, 0x0D0D0D0Du);
  const unsigned prev_nl = (threadIdx.x == 0 ? prev_byte : s[o - 1]) == '\n' ? 1u : 0u;
  const unsigned next_nl = s[o + PB] == '\n' ? 1u : 0u;   // LOOK >= 1 byte is always staged
  const unsigned eol = nl | (cr & ((nl >> 1) | (next_nl << 15)));
  return ((nl << 1) | prev_nl) & ~eol & 0xFFFFu;
}

__global__ void __launch_bounds__(PT) k_count_lines(const unsigned char* __restrict__ text, long long bytes,
                                                    long long num_tiles, long long* __restrict__ counts) {
  __shared__ __align__(16) unsigned char s[TILE + LOOK];
  __shared__ unsigned char s_prev;
  __shared__ int s_w[PT / 32];
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    stage_tile(text, bytes, tile * TILE, s, &s_prev);
    int c = __popc(line_starts16(s, s_prev));
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < PT / 32; w++) tot += s_w[w];
      counts[tile] = tot;
    }
    __syncthreads();
  }
}

// exclusive prefix sum of up to a few million tile counts; one CTA
__global__ void __launch_bounds__(1024) k_scan(const long long* __restrict__ in, long long n,
                                              long long* __restrict__ out, long long* __restrict__ total) {
  __shared__ long long s_sum[1024];
  const long long per = (n + 1023) / 1024;
  const long long lo = (long long)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  long long acc = 0;
  for (long long i = lo; i < hi; i++) acc += in[i];
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < 1024; t++) {
      const long long v = s_sum[t];
      s_sum[t] = run;
      run += v;
    }
    *total = run;
  }
  __syncthreads();
  long long run = s_sum[threadIdx.x];
  for (long long i = lo; i < hi; i++) {
    const long long v = in[i];
    out[i] = run;
    run += v;
  }
}

// Long.parseLong: optional sign, at least one digit, range checked.  Leaves q on the first byte after
// the digits.
__device__ __forceinline__ bool parse_long(const Tile& tv, int& q, long long* out) {
  unsigned char c = tv.at(q);
  bool neg = false;
  if (c == '-' || c == '+') {
    neg = c == '-';
    c = tv.at(++q);
  }
  if (c < '0' || c > '9') return false;
  // accumulate negatively so that Long.MIN_VALUE parses
  long long v = 0;
  const long long limit = neg ? LLONG_MIN : -LLONG_MAX;
  const long long multmin = limit / 10;
  while (c >= '0' && c <= '9') {
    const int dgt = c - '0';
    if (v < multmin) return false;
    v *= 10;
    if (v < limit + dgt) return false;
    v -= dgt;
    c = tv.at(++q);
  }
  *out = neg ? v : -v;
  return true;
}

__device__ __forceinline__ bool is_delim(unsigned char c) { return c == ',' || c == '\t'; }

// Float.parseFloat of the token [q, e) (already trimmed).  1 = value decided and correctly rounded,
// 0 = leave it to the host, -1 would be a syntax error -- also left to the host, which owns the message.
__device__ int parse_float_token(const Tile& tv, int q, int e, float* out) {
  if (q >= e) return 0;
  unsigned char c = tv.at(q);
  bool neg = false;
  if (c == '-' || c == '+') {
    neg = c == '-';
    q++;
  }
  unsigned long long mant = 0;
  int digits = 0, exp10 = 0;
  bool any = false, dot = false;
  for (; q < e; q++) {
    c = tv.at(q);
    if (c >= '0' && c <= '9') {
      any = true;
      if (mant == 0 && c == '0') {
        if (dot) exp10--;          // leading zeros after the point only move the exponent
        continue;
      }
      if (digits >= 15) return 0;  // more significant digits than one exact double: host
      mant = mant * 10 + (unsigned)(c - '0');
      digits++;
      if (dot) exp10--;
    } else if (c == '.' && !dot) {
      dot = true;
    } else {
      return 0;                    // exponent, suffix, hex, NaN, Infinity, garbage: host
    }
  }
  if (!any) return 0;
  if (exp10 < -22) return 0;
  static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  // mant < 10^15 < 2^53 and 10^k (k <= 22) are exact doubles: one correctly rounded division
  const double v = exp10 < 0 ? __ddiv_rn((double)mant, P10[-exp10]) : (double)mant;
  // double -> float rounds a second time; the two roundings can only disagree with a single one when
  // the double sits exactly on a float rounding boundary
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  if ((bits & 0x1FFFFFFFull) == 0x10000000ull) return 0;
  const float f = __double2float_rn(v);
  *out = neg ? -f : f;
  return 1;
}

__global__ void __launch_bounds__(PT) k_parse_lines(const ParseArgs a) {
  __shared__ __align__(16) unsigned char s[TILE + LOOK];
  __shared__ unsigned char s_prev;
  __shared__ int s_w[PT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const long long base = tile * TILE;
    stage_tile(a.text, a.bytes, base, s, &s_prev);
    Tile tv{a.text, a.bytes, base, s};
    unsigned flags = line_starts16(s, s_prev);
    // exclusive scan of the per-thread line counts over the CTA
    const int c = __popc(flags);
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; w++) wbase += s_w[w];
    long long slot = a.tile_off[tile] + wbase + inc - c;
    while (flags) {
      const int j = __ffs(flags) - 1;
      flags &= flags - 1;
      const int p = threadIdx.x * PB + j;   // tile offset of the line start
      int q = p;
      long long u = 0, it = 0;
      int err = PERR_NONE;
      float pref = 1.0f;
      if (!parse_long(tv, q, &u)) err = PERR_LONG;
      else if (!is_delim(tv.at(q))) err = tv.eol(q) ? PERR_TOKENS : PERR_LONG;
      else {
        q++;
        if (!parse_long(tv, q, &it)) err = tv.eol(q) ? PERR_TOKENS : PERR_LONG;
        else if (tv.eol(q)) {
          // two tokens: pref = 1.0
        } else if (!is_delim(tv.at(q))) {
          err = PERR_LONG;
        } else if (!a.boolean_data) {
          q++;
          int e = q;
          while (!tv.eol(e) && !is_delim(tv.at(e))) e++;
          int tq = q, te = e;           // String.trim(): strip bytes <= ' '
          while (tq < te && tv.at(tq) <= ' ') tq++;
          while (te > tq && tv.at(te - 1) <= ' ') te--;
          if (e == q) {
            // an empty third token is only legal when nothing but delimiters follows (String.split
            // drops trailing empty strings)
            int r = e;
            while (is_delim(tv.at(r))) r++;
            if (!tv.eol(r)) err = PERR_FLOAT;
          } else {
            float f;
            if (parse_float_token(tv, tq, te, &f) == 1) {
              pref = __fadd_rn(f, a.rating_shift);
            } else {
              const int k = atomicAdd(a.fix_count, 1);
              if (k < a.fix_cap) {
                a.fix_line[k] = slot;
                a.fix_pos[k] = base + tq;
                a.fix_len[k] = te - tq;
              } else {
                err = PERR_FIXUPS;
              }
              pref = 0.0f;
            }
          }
        }
      }
      if (err != PERR_NONE) {
        const unsigned long long pos = (unsigned long long)(base + p);
        const unsigned long long old = atomicMin(a.err_pos, pos);
        if (pos < old) *a.err_code = err;  // best effort: the code of the first line
      }
      a.user[slot] = a.transpose ? it : u;
      a.item[slot] = a.transpose ? u : it;
      a.pref[slot] = pref;
      slot++;
    }
    __syncthreads();
  }
}

// Java's Float.parseFloat grammar on the host for the tokens the device left undecided: strtof is
// correctly rounded; Java additionally allows a trailing f/F/d/D and spells the specials NaN / Infinity.
static bool host_parse_float(const char* tok, int len, float* out) {
  std::string t(tok, tok + len);
  if (t.empty()) return false;
  std::string body = t;
  const char last = body.back();
  if (last == 'f' || last == 'F' || last == 'd' || last == 'D') {
    // not a suffix when the token is a hex float without a binary exponent, which Java rejects anyway
    body.pop_back();
    if (body.empty()) return false;
  }
  const char* b = body.c_str();
  const char* unsigned_b = (*b == '+' || *b == '-') ? b + 1 : b;
  if (strcmp(unsigned_b, "NaN") == 0) {
    *out = nanf("");
    return true;
  }
  if (strcmp(unsigned_b, "Infinity") == 0) {
    *out = *b == '-' ? -INFINITY : INFINITY;
    return true;
  }
  // strtof also takes "inf", "nan(...)", leading blanks: Java does not
  for (const char* c = unsigned_b; *c; c++) {
    const bool ok = (*c >= '0' && *c <= '9') || *c == '.' || *c == 'e' || *c == 'E' || *c == '+' || *c == '-' ||
                    *c == 'x' || *c == 'X' || *c == 'p' || *c == 'P' || (*c >= 'a' && *c <= 'f') ||
                    (*c >= 'A' && *c <= 'F');
    if (!ok) return false;
  }
  if (!((*unsigned_b >= '0' && *unsigned_b <= '9') || *unsigned_b == '.')) return false;
  char* end = nullptr;
  const float v = strtof(b, &end);
  if (end == b || *end != '\0') return false;
  *out = v;
  return true;
}

// ------------------------------------------------------------------------------------------------
// preparation
// ------------------------------------------------------------------------------------------------
#define HT_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

// slot of `key` in an open-addressing table of mask+1 slots (+1 spare slot for the key that equals
// the empty marker)
__device__ __forceinline__ unsigned long long ht_insert(unsigned long long* keys, unsigned long long mask,
                                                        unsigned long long key) {
  if (key == HT_EMPTY) return mask + 1;
  unsigned long long s = mix64(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(keys + s, HT_EMPTY, key);
    if (prev == HT_EMPTY || prev == key) return s;
    s = (s + 1) & mask;
  }
}

// TasteHadoopUtils.idToIndex; Java: '%' binds tighter than '&' and truncates toward zero
__device__ __forceinline__ int id_to_index(long long id) {
  const int h = (int)((unsigned long long)id ^ ((unsigned long long)id >> 32));
  return 0x7FFFFFFF & (h % 0x7FFFFFFE);
}

__global__ void __launch_bounds__(256) k_fill64(long long* __restrict__ p, unsigned long long n, long long v) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void __launch_bounds__(256) k_id_to_index(const long long* __restrict__ ids, long long n,
                                                     int* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = id_to_index(ids[i]);
}

struct PrepArgs {
  const long long* user;
  const long long* item;
  const float* pref;
  long long n;
  unsigned long long mask;
  unsigned long long* ukeys;      // user table
  int* ucount;                    // surviving (de-duplicated) preferences per user slot
  unsigned long long* pkeys;      // (user slot, index) table
  unsigned long long* ppos;       // 1 + position of the last event of the pair
  unsigned long long* ikeys;      // index table
  long long* iminid;              // minimum itemID per index
  int* idx;                       // [n] item index of the event
  unsigned int* uslot;            // [n]
  unsigned long long* pslot;      // [n]
  unsigned char* keep;            // [n]
  int min_prefs;
};

__global__ void __launch_bounds__(256) k_prep_insert(const PrepArgs a) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
    const long long it = a.item[t];
    const int ix = id_to_index(it);
    a.idx[t] = ix;
    const unsigned long long su = ht_insert(a.ukeys, a.mask, (unsigned long long)a.user[t]);
    a.uslot[t] = (unsigned int)su;
    const unsigned long long sp = ht_insert(a.pkeys, a.mask, (su << 32) | (unsigned long long)(unsigned)ix);
    a.pslot[t] = sp;
    atomicMax(a.ppos + sp, (unsigned long long)t + 1ull);
    const unsigned long long si = ht_insert(a.ikeys, a.mask, (unsigned long long)(unsigned)ix);
    atomicMin(a.iminid + si, it);
  }
}

__global__ void __launch_bounds__(256) k_prep_mark(const PrepArgs a) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
    // userVector.set(index, 0.0) REMOVES the element (RandomAccessSparseVector.setQuick, :125-132): a pair whose
    // last preference is 0.0 -- e.g. rating + ratingShift == 0 -- neither counts toward minPrefsPerUser
    // (getNumNondefaultElements, ToUserVectorsReducer.java:75) nor reaches the item vectors
    const bool last = a.ppos[a.pslot[t]] == (unsigned long long)t + 1ull && a.pref[t] != 0.0f;
    a.keep[t] = last ? 1 : 0;
    if (last) atomicAdd(a.ucount + a.uslot[t], 1);
  }
}

__global__ void __launch_bounds__(256) k_prep_flags(const PrepArgs a) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x)
    if (a.keep[t] && a.ucount[a.uslot[t]] < a.min_prefs) a.keep[t] = 0;
}

// users that survive minPrefsPerUser get dense numbers 0..U-1 (the order is that of the atomics: any
// numbering serves -- the numbers only name counter columns of the exact measure)
__global__ void __launch_bounds__(256) k_number_users(const int* __restrict__ ucount, unsigned long long slots,
                                                      int min_prefs, int* __restrict__ unum,
                                                      unsigned long long* __restrict__ out) {
  for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots;
       s += (unsigned long long)gridDim.x * blockDim.x) {
    const bool ok = ucount[s] >= min_prefs && ucount[s] > 0;
    unum[s] = ok ? (int)atomicAdd(out, 1ull) : -1;
  }
}

// occupied slots of the index table -> (index, min itemID) pairs, unordered
__global__ void __launch_bounds__(256) k_collect_index(const unsigned long long* __restrict__ ikeys,
                                                       const long long* __restrict__ iminid, unsigned long long slots,
                                                       int* __restrict__ out_idx, long long* __restrict__ out_id,
                                                       unsigned long long* __restrict__ count) {
  for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < slots;
       s += (unsigned long long)gridDim.x * blockDim.x) {
    if (ikeys[s] != HT_EMPTY) {
      const unsigned long long k = atomicAdd(count, 1ull);
      out_idx[k] = (int)ikeys[s];
      out_id[k] = iminid[s];
    }
  }
}

static constexpr int CT = 256, CPT = 8;  // compaction: 2048 events per tile

__global__ void __launch_bounds__(CT) k_keep_counts(const unsigned char* __restrict__ keep, long long n,
                                                    long long num_tiles, long long* __restrict__ counts) {
  __shared__ int s_w[CT / 32];
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long t0 = tile * (CT * CPT) + (long long)threadIdx.x * CPT;
    int c = 0;
#pragma unroll
    for (int j = 0; j < CPT; j++)
      if (t0 + j < n && keep[t0 + j]) c++;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < CT / 32; w++) tot += s_w[w];
      counts[tile] = tot;
    }
    __syncthreads();
  }
}

// surviving events, in input order: row = rank of the event's index among the sorted distinct indexes
__global__ void __launch_bounds__(CT) k_compact(const PrepArgs a, long long num_tiles,
                                                const long long* __restrict__ tile_off,
                                                const float* __restrict__ pref, const int* __restrict__ sorted_idx,
                                                long long num_items, const int* __restrict__ unum,
                                                long long* __restrict__ out_row, long long* __restrict__ out_user,
                                                long long* __restrict__ out_ucol, float* __restrict__ out_pref) {
  __shared__ int s_w[CT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long t0 = tile * (CT * CPT) + (long long)threadIdx.x * CPT;
    int c = 0;
#pragma unroll
    for (int j = 0; j < CPT; j++)
      if (t0 + j < a.n && a.keep[t0 + j]) c++;
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; w++) wbase += s_w[w];
    long long slot = tile_off[tile] + wbase + inc - c;
    for (int j = 0; j < CPT; j++) {
      const long long t = t0 + j;
      if (t < a.n && a.keep[t]) {
        const int ix = a.idx[t];
        long long lo = 0, hi = num_items;
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if (__ldg(sorted_idx + mid) < ix) lo = mid + 1;
          else hi = mid;
        }
        out_row[slot] = lo;
        out_user[slot] = a.user[t];
        out_ucol[slot] = unum[a.uslot[t]];
        out_pref[slot] = pref[t];
        slot++;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Scratch {  // cudaMalloc'ed scratch released on scope exit
  std::vector<void*> ptrs;
  template <typename T>
  cudaError_t get(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) {
      ptrs.push_back(q);
      *p = (T*)q;
    }
    return e;
  }
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
};

static int grid_for(mb200_ctx* ctx, long long work_items, int per_sm) {
  const long long cap = (long long)ctx->num_sms * per_sm;
  return (int)std::max<long long>(1, std::min<long long>(work_items, cap));
}

extern "C" {

int mb200_id_to_index(mb200_ctx* ctx, const int64_t* ids, int64_t n, int32_t* out, int mem) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_id_to_index: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  if (n < 0 || (n > 0 && (!ids || !out))) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_id_to_index: bad arguments");
  if (n == 0) return MB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  Scratch sc;
  const long long* d_in = (const long long*)ids;
  int* d_out = out;
  if (mem == MB200_MEM_HOST) {
    long long* t = nullptr;
    MB_CUDA(ctx, sc.get(&t, (size_t)n));
    MB_CUDA(ctx, sc.get(&d_out, (size_t)n));
    MB_CUDA(ctx, cudaMemcpyAsync(t, ids, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    d_in = t;
  }
  k_id_to_index<<<grid_for(ctx, ceil_div64(n, 256), 16), 256, 0, ctx->stream>>>(d_in, n, d_out);
  ctx->launches++;
  MB_CUDA(ctx, cudaGetLastError());
  if (mem == MB200_MEM_HOST) MB_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

int mb200_events_parse(mb200_ctx* ctx, const char* text, int64_t bytes, int mem, int boolean_data,
                       float rating_shift, int transpose, mb200_events** out) {
  if (!ctx || !out) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_parse: ctx/out is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  *out = nullptr;
  if (bytes < 0 || (bytes > 0 && !text)) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_parse: bad text buffer");
  if (mem != MB200_MEM_HOST && mem != MB200_MEM_DEVICE)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_parse: mem must be MB200_MEM_HOST or MB200_MEM_DEVICE");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  mb200_events* ev = new mb200_events();
  ev->ctx = ctx;
  if (bytes == 0) {
    *out = ev;
    return MB200_OK;
  }
  Scratch sc;
  const unsigned char* d_text = (const unsigned char*)text;
  if (mem == MB200_MEM_HOST) {
    unsigned char* t = nullptr;
    cudaError_t e = sc.get(&t, (size_t)bytes);
    if (e != cudaSuccess) {
      events_free(ev);
      return mb200_fail(ctx, MB200_ERR_OOM, "mb200_events_parse: cannot stage %lld bytes of text: %s", (long long)bytes, cudaGetErrorString(e));
    }
    e = cudaMemcpyAsync(t, text, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
      events_free(ev);
      return mb200_fail(ctx, MB200_ERR_CUDA, "mb200_events_parse: H2D copy failed: %s", cudaGetErrorString(e));
    }
    d_text = t;
  }
  const long long num_tiles = ceil_div64(bytes, TILE);
  const int FIX_CAP = 1 << 20;
  long long *d_counts = nullptr, *d_off = nullptr, *d_total = nullptr, *d_fix_line = nullptr, *d_fix_pos = nullptr;
  unsigned long long* d_err = nullptr;
  int *d_errc = nullptr, *d_fix_len = nullptr, *d_fix_count = nullptr;
#define EV_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      events_free(ev);                                                                       \
      return mb200_fail(ctx, _e == cudaErrorMemoryAllocation ? MB200_ERR_OOM : MB200_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                        \
  } while (0)
  EV_TRY(sc.get(&d_counts, (size_t)num_tiles));
  EV_TRY(sc.get(&d_off, (size_t)num_tiles));
  EV_TRY(sc.get(&d_total, 1));
  EV_TRY(sc.get(&d_err, 1));
  EV_TRY(sc.get(&d_errc, 1));
  EV_TRY(sc.get(&d_fix_line, (size_t)FIX_CAP));
  EV_TRY(sc.get(&d_fix_pos, (size_t)FIX_CAP));
  EV_TRY(sc.get(&d_fix_len, (size_t)FIX_CAP));
  EV_TRY(sc.get(&d_fix_count, 1));
  const int grid = grid_for(ctx, num_tiles, 8);
  {
    ProfScope prof(ctx, MB200_K_PARSE);
    k_count_lines<<<grid, PT, 0, ctx->stream>>>(d_text, bytes, num_tiles, d_counts);
    k_scan<<<1, 1024, 0, ctx->stream>>>(d_counts, num_tiles, d_off, d_total);
  }
  ctx->launches += 2;
  long long total = 0;
  EV_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
  EV_TRY(cudaStreamSynchronize(ctx->stream));
  ev->n = total;
  if (total > 0) {
    EV_TRY(cudaMalloc(&ev->user, (size_t)total * 8));
    EV_TRY(cudaMalloc(&ev->item, (size_t)total * 8));
    EV_TRY(cudaMalloc(&ev->pref, (size_t)total * 4));
    EV_TRY(cudaMemsetAsync(d_err, 0xFF, 8, ctx->stream));
    EV_TRY(cudaMemsetAsync(d_errc, 0, 4, ctx->stream));
    EV_TRY(cudaMemsetAsync(d_fix_count, 0, 4, ctx->stream));
    ParseArgs a;
    a.text = d_text;
    a.bytes = bytes;
    a.num_tiles = num_tiles;
    a.boolean_data = boolean_data ? 1 : 0;
    a.transpose = transpose ? 1 : 0;
    a.rating_shift = rating_shift;
    a.tile_off = d_off;
    a.user = ev->user;
    a.item = ev->item;
    a.pref = ev->pref;
    a.err_pos = d_err;
    a.err_code = d_errc;
    a.fix_line = d_fix_line;
    a.fix_pos = d_fix_pos;
    a.fix_len = d_fix_len;
    a.fix_count = d_fix_count;
    a.fix_cap = FIX_CAP;
    {
      ProfScope prof(ctx, MB200_K_PARSE);
      k_parse_lines<<<grid, PT, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
    EV_TRY(cudaGetLastError());
    unsigned long long err_pos = 0;
    int err_code = 0, nfix = 0;
    EV_TRY(cudaMemcpyAsync(&err_pos, d_err, 8, cudaMemcpyDeviceToHost, ctx->stream));
    EV_TRY(cudaMemcpyAsync(&err_code, d_errc, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EV_TRY(cudaMemcpyAsync(&nfix, d_fix_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EV_TRY(cudaStreamSynchronize(ctx->stream));
    if (err_pos != 0xFFFFFFFFFFFFFFFFull) {
      events_free(ev);
      static const char* what[] = {"", "NumberFormatException: not a long", "ArrayIndexOutOfBoundsException: fewer than two tokens",
                                   "NumberFormatException: empty preference token", "too many irregular preference tokens"};
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_parse: malformed line at byte offset %llu (%s)", err_pos,
                        what[err_code >= 0 && err_code <= 4 ? err_code : 0]);
    }
    if (nfix > 0) {
      // tokens the device parser left undecided: Java-grammar parse on the host, patched in place
      std::vector<long long> fl((size_t)nfix), fp((size_t)nfix);
      std::vector<int> flen((size_t)nfix);
      EV_TRY(cudaMemcpy(fl.data(), d_fix_line, (size_t)nfix * 8, cudaMemcpyDeviceToHost));
      EV_TRY(cudaMemcpy(fp.data(), d_fix_pos, (size_t)nfix * 8, cudaMemcpyDeviceToHost));
      EV_TRY(cudaMemcpy(flen.data(), d_fix_len, (size_t)nfix * 4, cudaMemcpyDeviceToHost));
      std::vector<char> tok;
      for (int i = 0; i < nfix; i++) {
        tok.resize((size_t)flen[i]);
        if (mem == MB200_MEM_HOST) memcpy(tok.data(), text + fp[i], (size_t)flen[i]);
        else EV_TRY(cudaMemcpy(tok.data(), d_text + fp[i], (size_t)flen[i], cudaMemcpyDeviceToHost));
        float v = 0.f;
        if (!host_parse_float(tok.data(), flen[i], &v)) {
          events_free(ev);
          return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_parse: NumberFormatException: For input string: \"%.*s\" (byte offset %lld)",
                            flen[i] > 64 ? 64 : flen[i], tok.data(), fp[i]);
        }
        v = v + rating_shift;
        EV_TRY(cudaMemcpy(ev->pref + fl[i], &v, 4, cudaMemcpyHostToDevice));
      }
    }
  }
#undef EV_TRY
  *out = ev;
  return MB200_OK;
}

int mb200_events_create(mb200_ctx* ctx, const int64_t* user, const int64_t* item, const float* pref, int64_t n,
                        int mem, mb200_events** out) {
  if (!ctx || !out) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_create: ctx/out is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  *out = nullptr;
  if (n < 0 || (n > 0 && (!user || !item || !pref))) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_events_create: bad arguments");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  mb200_events* ev = new mb200_events();
  ev->ctx = ctx;
  ev->n = n;
  if (n > 0) {
    const cudaMemcpyKind kind = mem == MB200_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    cudaError_t e = cudaMalloc(&ev->user, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ev->item, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ev->pref, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ev->user, user, (size_t)n * 8, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ev->item, item, (size_t)n * 8, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ev->pref, pref, (size_t)n * 4, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      events_free(ev);
      return mb200_fail(ctx, e == cudaErrorMemoryAllocation ? MB200_ERR_OOM : MB200_ERR_CUDA, "mb200_events_create: %s", cudaGetErrorString(e));
    }
  }
  *out = ev;
  return MB200_OK;
}

int mb200_events_count(mb200_events* ev, int64_t* n) {
  if (!ev || !n) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_events_count: NULL argument");
  *n = ev->n;
  return MB200_OK;
}

int mb200_events_columns(mb200_events* ev, int64_t** user, int64_t** item, float** pref) {
  if (!ev) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_events_columns: events is NULL");
  if (user) *user = (int64_t*)ev->user;
  if (item) *item = (int64_t*)ev->item;
  if (pref) *pref = ev->pref;
  return MB200_OK;
}

int mb200_events_read(mb200_events* ev, int64_t* user, int64_t* item, float* pref) {
  if (!ev) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_events_read: events is NULL");
  mb200_ctx* ctx = ev->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ev->n == 0) return MB200_OK;
  if (user) MB_CUDA(ctx, cudaMemcpyAsync(user, ev->user, (size_t)ev->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (item) MB_CUDA(ctx, cudaMemcpyAsync(item, ev->item, (size_t)ev->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (pref) MB_CUDA(ctx, cudaMemcpyAsync(pref, ev->pref, (size_t)ev->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

int mb200_events_destroy(mb200_events* ev) {
  if (!ev) return MB200_OK;
  mb200_ctx* ctx = ev->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  events_free(ev);
  return MB200_OK;
}

int mb200_events_prepare(mb200_events* ev, int32_t min_prefs_per_user, mb200_prefs** out) {
  if (!ev || !out) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_events_prepare: NULL argument");
  mb200_ctx* ctx = ev->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  *out = nullptr;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long n = ev->n;
  if (n > (1LL << 30)) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "mb200_events_prepare: at most 2^30 events per call (got %lld)", n);
  mb200_prefs* pm = new mb200_prefs();
  pm->ctx = ctx;
  if (n == 0) {
    *out = pm;
    return MB200_OK;
  }
  unsigned long long slots = 1024;
  while (slots < 2ull * (unsigned long long)n) slots <<= 1;
  const unsigned long long mask = slots - 1;
  Scratch sc;
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.user = ev->user;
  a.item = ev->item;
  a.pref = ev->pref;
  a.n = n;
  a.mask = mask;
  a.min_prefs = min_prefs_per_user;
  unsigned long long *d_nusers = nullptr, *d_nidx = nullptr;
  long long *d_counts = nullptr, *d_off = nullptr, *d_total = nullptr, *d_uid = nullptr;
  int *d_uidx = nullptr, *d_sorted = nullptr, *d_unum = nullptr;
  const long long num_tiles = ceil_div64(n, CT * CPT);
#define PM_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      prefs_free(pm);                                                                        \
      return mb200_fail(ctx, _e == cudaErrorMemoryAllocation ? MB200_ERR_OOM : MB200_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                        \
  } while (0)
  PM_TRY(sc.get(&a.ukeys, (size_t)slots + 1));
  PM_TRY(sc.get(&a.ucount, (size_t)slots + 1));
  PM_TRY(sc.get(&a.pkeys, (size_t)slots + 1));
  PM_TRY(sc.get(&a.ppos, (size_t)slots + 1));
  PM_TRY(sc.get(&a.ikeys, (size_t)slots + 1));
  PM_TRY(sc.get(&a.iminid, (size_t)slots + 1));
  PM_TRY(sc.get(&a.idx, (size_t)n));
  PM_TRY(sc.get(&a.uslot, (size_t)n));
  PM_TRY(sc.get(&a.pslot, (size_t)n));
  PM_TRY(sc.get(&a.keep, (size_t)n));
  PM_TRY(sc.get(&d_nusers, 1));
  PM_TRY(sc.get(&d_unum, (size_t)slots + 1));
  PM_TRY(sc.get(&d_nidx, 1));
  PM_TRY(sc.get(&d_counts, (size_t)num_tiles));
  PM_TRY(sc.get(&d_off, (size_t)num_tiles));
  PM_TRY(sc.get(&d_total, 1));
  PM_TRY(cudaMemsetAsync(a.ukeys, 0xFF, (slots + 1) * 8, ctx->stream));
  PM_TRY(cudaMemsetAsync(a.pkeys, 0xFF, (slots + 1) * 8, ctx->stream));
  PM_TRY(cudaMemsetAsync(a.ikeys, 0xFF, (slots + 1) * 8, ctx->stream));
  PM_TRY(cudaMemsetAsync(a.ucount, 0, (slots + 1) * 4, ctx->stream));
  PM_TRY(cudaMemsetAsync(a.ppos, 0, (slots + 1) * 8, ctx->stream));
  k_fill64<<<grid_for(ctx, (long long)((slots + 1 + 255) / 256), 16), 256, 0, ctx->stream>>>(a.iminid, slots + 1, LLONG_MAX);
  PM_TRY(cudaMemsetAsync(d_nusers, 0, 8, ctx->stream));
  PM_TRY(cudaMemsetAsync(d_nidx, 0, 8, ctx->stream));
  const int grid = grid_for(ctx, ceil_div64(n, 256), 16);
  {
    ProfScope prof(ctx, MB200_K_PREPARE);
    k_prep_insert<<<grid, 256, 0, ctx->stream>>>(a);
    k_prep_mark<<<grid, 256, 0, ctx->stream>>>(a);
    k_prep_flags<<<grid, 256, 0, ctx->stream>>>(a);
    k_number_users<<<grid_for(ctx, (long long)((slots + 1 + 255) / 256), 16), 256, 0, ctx->stream>>>(a.ucount, slots + 1, min_prefs_per_user, d_unum, d_nusers);
  }
  ctx->launches += 4;
  PM_TRY(cudaGetLastError());
  // distinct indexes: collect, sort on the host (a few million 12-byte entries at most), send back
  PM_TRY(sc.get(&d_uidx, (size_t)std::min<unsigned long long>(slots + 1, (unsigned long long)n)));
  PM_TRY(sc.get(&d_uid, (size_t)std::min<unsigned long long>(slots + 1, (unsigned long long)n)));
  k_collect_index<<<grid_for(ctx, (long long)((slots + 1 + 255) / 256), 16), 256, 0, ctx->stream>>>(a.ikeys, a.iminid, slots + 1, d_uidx, d_uid, d_nidx);
  ctx->launches++;
  unsigned long long nusers = 0, nidx = 0;
  PM_TRY(cudaMemcpyAsync(&nusers, d_nusers, 8, cudaMemcpyDeviceToHost, ctx->stream));
  PM_TRY(cudaMemcpyAsync(&nidx, d_nidx, 8, cudaMemcpyDeviceToHost, ctx->stream));
  PM_TRY(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> h_idx((size_t)nidx);
  std::vector<long long> h_id((size_t)nidx);
  PM_TRY(cudaMemcpy(h_idx.data(), d_uidx, (size_t)nidx * 4, cudaMemcpyDeviceToHost));
  PM_TRY(cudaMemcpy(h_id.data(), d_uid, (size_t)nidx * 8, cudaMemcpyDeviceToHost));
  std::vector<size_t> order((size_t)nidx);
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return h_idx[x] < h_idx[y]; });
  pm->num_items = (int64_t)nidx;
  pm->num_users = (int64_t)nusers;
  pm->index_values.resize((size_t)nidx);
  pm->item_id.resize((size_t)nidx);
  for (size_t i = 0; i < order.size(); i++) {
    pm->index_values[i] = h_idx[order[i]];
    pm->item_id[i] = h_id[order[i]];
  }
  PM_TRY(sc.get(&d_sorted, (size_t)nidx));
  PM_TRY(cudaMemcpyAsync(d_sorted, pm->index_values.data(), (size_t)nidx * 4, cudaMemcpyHostToDevice, ctx->stream));
  // order-preserving compaction of the survivors
  const int cgrid = grid_for(ctx, num_tiles, 8);
  k_keep_counts<<<cgrid, CT, 0, ctx->stream>>>(a.keep, n, num_tiles, d_counts);
  k_scan<<<1, 1024, 0, ctx->stream>>>(d_counts, num_tiles, d_off, d_total);
  ctx->launches += 2;
  long long total = 0;
  PM_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
  PM_TRY(cudaStreamSynchronize(ctx->stream));
  pm->n = total;
  if (total > 0) {
    PM_TRY(cudaMalloc(&pm->row, (size_t)total * 8));
    PM_TRY(cudaMalloc(&pm->user, (size_t)total * 8));
    PM_TRY(cudaMalloc(&pm->pref, (size_t)total * 4));
    PM_TRY(cudaMalloc(&pm->ucol, (size_t)total * 8));
    {
      ProfScope prof(ctx, MB200_K_PREPARE);
      k_compact<<<cgrid, CT, 0, ctx->stream>>>(a, num_tiles, d_off, ev->pref, d_sorted, (long long)nidx, d_unum, pm->row, pm->user,
                                               pm->ucol, pm->pref);
    }
    ctx->launches++;
    PM_TRY(cudaGetLastError());
  }
  PM_TRY(cudaStreamSynchronize(ctx->stream));
#undef PM_TRY
  *out = pm;
  return MB200_OK;
}

int mb200_prefs_info(mb200_prefs* p, int64_t* n, int64_t* num_items, int64_t* num_users) {
  if (!p) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_prefs_info: prefs is NULL");
  if (n) *n = p->n;
  if (num_items) *num_items = p->num_items;
  if (num_users) *num_users = p->num_users;
  return MB200_OK;
}

int mb200_prefs_columns(mb200_prefs* p, int64_t** row, int64_t** user, float** pref) {
  if (!p) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_prefs_columns: prefs is NULL");
  if (row) *row = (int64_t*)p->row;
  if (user) *user = (int64_t*)p->user;
  if (pref) *pref = p->pref;
  return MB200_OK;
}

int mb200_prefs_user_columns(mb200_prefs* p, int64_t** ucol) {
  if (!p || !ucol) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_prefs_user_columns: NULL argument");
  *ucol = (int64_t*)p->ucol;
  return MB200_OK;
}

int mb200_prefs_read(mb200_prefs* p, int64_t* row, int64_t* user, int64_t* ucol, float* pref) {
  if (!p) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_prefs_read: prefs is NULL");
  mb200_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (p->n == 0) return MB200_OK;
  if (row) MB_CUDA(ctx, cudaMemcpyAsync(row, p->row, (size_t)p->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (user) MB_CUDA(ctx, cudaMemcpyAsync(user, p->user, (size_t)p->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (ucol) MB_CUDA(ctx, cudaMemcpyAsync(ucol, p->ucol, (size_t)p->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (pref) MB_CUDA(ctx, cudaMemcpyAsync(pref, p->pref, (size_t)p->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

int mb200_prefs_tables(mb200_prefs* p, int64_t* item_id, int32_t* index_values) {
  if (!p) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_prefs_tables: prefs is NULL");
  if (item_id && p->num_items) memcpy(item_id, p->item_id.data(), (size_t)p->num_items * 8);
  if (index_values && p->num_items) memcpy(index_values, p->index_values.data(), (size_t)p->num_items * 4);
  return MB200_OK;
}

int mb200_prefs_destroy(mb200_prefs* p) {
  if (!p) return MB200_OK;
  mb200_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  prefs_free(p);
  return MB200_OK;
}

}  // extern "C"
