// cosine.cu -- all-pairs sketch cosine + per-row top-k.
//
// Reference semantics:
//   DoubleCountMinSketch.cosine (DoubleCountMinSketch.java:114-149): per depth row i,
//     cos_i = AB / (sqrt(AA) * sqrt(BB)); rows with a zero denominator are skipped; the result is
//     the min over the remaining rows, NaN if none.
//   RowSimilarityJob with CosineSimilarity (RowSimilarityJob.java:478-501,515-559;
//     TopElementsQueue.java:26-59): keep sim >= threshold, drop the diagonal, per-row top-k where
//     a candidate must exceed Double.MIN_VALUE.  Ties: lower index wins (north star).
//
// Kernels:
//   K2 k_normalize   counters -> unit-norm 16-bit rows [d][E][ld] + validity bit masks (HBM-bound)
//   K3 k_cosine      TMA -> smem -> tcgen05.mma (FP32 accumulators in TMEM); the epilogue warps take
//                    the running min over depth (kept in TMEM) and select per-row candidates, so the
//                    similarity matrix never reaches HBM
//   K5 k_merge       merges the per-(row, column-chunk) candidate lists; k_rescore recomputes the
//                    kept candidates exactly from the integer counters (bit-equal to the reference's
//                    FP64 arithmetic) and certifies the top-k; k_exact_rows is the exact full-row
//                    path for rows that could not be certified
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "tc05.cuh"

using namespace tc05;

// ------------------------------------------------------------------------------------------------
// constants shared by host and device
// ------------------------------------------------------------------------------------------------
static constexpr int BM = 128;            // rows of A per tile (TMEM lanes)
static constexpr int BK = 64;             // K elements per pipeline stage (one 128-byte swizzle row)
static constexpr int BAND_MAX = 8192;      // columns above the cut the band pass can hold per row, at most (the job's
                                           // band_cap is 2048, 4096 or 8192: as much as 4 GB of buffers allow)
static constexpr int BAND_MIN = 2048;
static constexpr int LCAP = 512;           // capacity of one K3 candidate list: a thread appends a whole tile's worth
                                           // (<= BN / 2 entries) between two threshold raises, which run AFTER the
                                           // accumulator has gone back to the MMA warp
static constexpr int LARGE_MAX = 2048;     // candidates per row of the multi-pass selection (k beyond CAP)
static constexpr int CAP = 256;           // candidate-list capacity per (row, column chunk, half)
static constexpr int COS_THREADS = 384;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
static constexpr float F16_SCALE = 4096.0f;  // rows are stored as x/||x|| * 2^12 in FP16

struct CosParams {
  const int4* items;        // work items (m block, first tile, end tile, tile stride)
  int32_t num_items;
  int32_t total_tiles, tiles_per_block;
  int32_t depth, kblocks, stages;
  int64_t a_count;
  const uint32_t* a_valid;
  int64_t a_vw;
  const uint32_t* b_valid;
  int64_t b_vw;
  uint32_t a_id_mul, a_id_off, b_id_mul, b_id_add, b_id_base;
  int32_t exclude_self, nonstrict;
  float thr_init;
  int32_t ksel;
  uint2* lists;             // [item][half][128][LCAP] (value bits, index)
  int32_t* list_cnt;        // [item][half][128]
  float* list_bound;        // [item][half][128]: every candidate not in the list has value <= bound
  uint32_t* row_thr;        // [a rows padded]: best published lower bound of the row's ksel-th value (ordered uint)
  float* dense_out;
  int64_t dense_ld;
  float inv_scale2;
  uint32_t idesc;
  uint64_t policy_a, policy_b;  // L2 eviction priority of the A / B tile loads
  // pull-gather mode: block g of B is only readable once ready[g] == epoch (written by the copy
  // stream behind the DMA that brought the block from its owner); rot rotates the block order so
  // that every GPU starts with the blocks that arrive first
  const uint32_t* ready;
  uint32_t epoch;
  int32_t rot, num_blocks;
  uint32_t* abort_flag;
  // band pass (second chance of rows the candidate lists could not certify): the A rows are a compact copy of
  // those rows, a_ids[row] is the row's global index, row_cut[row] the fixed admission threshold (scaled tensor
  // value) below which a column is provably outside the row's top-k; every column above it is LISTED -- no
  // selection, no threshold raising, no lists: straight into the row's candidate buffer band_cand[row][band_cap]
  // through the cursor band_cnt[row] (a count beyond band_cap sends the row to the exact full-row path)
  const uint32_t* a_ids;
  const float* row_cut;
  uint32_t* band_cand;
  float* band_val;
  int32_t* band_cnt;
  int32_t band_cap;
  // k beyond the fused capacity (large_k_topk_locked): the candidates are taken 128 ranks at a time; row_ceil[row] is
  // the key (value desc, index asc) of the last candidate already taken -- only columns strictly after it qualify
  const unsigned long long* row_ceil;
};

// ------------------------------------------------------------------------------------------------
// K2: row norms + conversion
// ------------------------------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// One CTA per sketch row.  Rows of up to 256 x NORM_VPT counters are read ONCE with 16-byte loads and
// held in registers between the norm and the conversion (8 B read + 2 B written per counter: the
// kernel's whole HBM traffic); wider rows take the two-pass path.
static constexpr int NORM_VPT = 16;
template <typename OutT>
__global__ void __launch_bounds__(256) k_normalize(const long long* __restrict__ counters, long long E,
                                                   int d, int W, int ld, float scale,
                                                   OutT* __restrict__ out, uint32_t* __restrict__ valid,
                                                   long long vw, unsigned long long* __restrict__ flags) {
  __shared__ double red[8];
  const long long bid = blockIdx.x;
  const long long e = bid / d;
  const int i = (int)(bid % d);
  const long long* row = counters + (size_t)bid * W;
  OutT* o = out + ((size_t)i * E + e) * ld;
  const bool in_regs = (W & 1) == 0 && W <= 256 * NORM_VPT;
  long long v[NORM_VPT];
  double ss = 0.0;
  bool neg = false;
  if (in_regs) {
    const longlong2* row2 = reinterpret_cast<const longlong2*>(row);
    const int pairs = W >> 1;
#pragma unroll
    for (int t = 0; t < NORM_VPT / 2; t++) {
      const int j = t * 256 + threadIdx.x;
      longlong2 x = make_longlong2(0, 0);
      if (j < pairs) x = __ldg(row2 + j);
      v[2 * t] = x.x;
      v[2 * t + 1] = x.y;
      neg |= x.x < 0 || x.y < 0;
      ss += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
    }
  } else {
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      const long long xi = row[j];
      neg |= xi < 0;
      const double x = (double)xi;
      ss += x * x;
    }
  }
  if (neg) flags[FLAG_MIXED] = 1ull;
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < 8; w++) tot += red[w];
  const double nrm = sqrt(tot);
  // a counter of 1 quantum must stay non-zero in the 16-bit row (x / ||x|| * 2^12 > 2^-24): "tensor value 0
  // <=> exact similarity 0" is what lets non-negative data drop zero products unseen
  if (threadIdx.x == 0 && nrm >= 0x1.0p36) flags[FLAG_MIXED] = 1ull;
  const double inv = nrm > 0.0 ? (double)scale / nrm : 0.0;
  if (in_regs) {
    // ld is even and the row base is 4-byte aligned: two 16-bit outputs per store
    const int pairs_ld = ld >> 1, pairs = W >> 1;
#pragma unroll
    for (int t = 0; t < NORM_VPT / 2; t++) {
      const int j = t * 256 + threadIdx.x;
      if (j < pairs_ld) {
        const float a = j < pairs ? (float)((double)v[2 * t] * inv) : 0.0f;
        const float b = j < pairs ? (float)((double)v[2 * t + 1] * inv) : 0.0f;
        OutT two[2] = {to_out<OutT>(a), to_out<OutT>(b)};
        *reinterpret_cast<uint32_t*>(o + 2 * j) = *reinterpret_cast<const uint32_t*>(two);
      }
    }
    for (int j = (NORM_VPT / 2) * 256 + threadIdx.x; j < pairs_ld; j += 256)   // padding beyond the register tile
      *reinterpret_cast<uint32_t*>(o + 2 * j) = 0u;
  } else {
    for (int j = threadIdx.x; j < ld; j += blockDim.x) {
      const float x = j < W ? (float)((double)row[j] * inv) : 0.0f;
      o[j] = to_out<OutT>(x);
    }
  }
  if (threadIdx.x == 0 && nrm > 0.0) atomicOr(&valid[(size_t)i * vw + (e >> 5)], 1u << (e & 31));
}

// ------------------------------------------------------------------------------------------------
// candidate keys: (value desc, index asc) as one descending 64-bit order
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(uint32_t u) { return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ uint32_t ord2f(uint32_t o) { return o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu); }
__device__ __forceinline__ unsigned long long make_key(uint32_t vbits, uint32_t id) {
  return ((unsigned long long)f2ord(vbits) << 32) | (unsigned long long)(~id);
}
__device__ __forceinline__ uint2 key_entry(unsigned long long k) {
  return make_uint2(ord2f((uint32_t)(k >> 32)), ~(uint32_t)k);
}

// Bitonic sort, descending, of 32*NPL keys held NPL per lane; element e = s*32 + lane.
template <int NPL>
__device__ __forceinline__ void warp_sort_desc(unsigned long long (&key)[NPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * NPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int js = j >> 5;
#pragma unroll
        for (int s = 0; s < NPL; s++) {
          if ((s & js) == 0) {
            const int e = s * 32;  // lane bits do not reach bit k when k > 32... they never matter here
            const bool desc = (((e) & k) == 0);
            unsigned long long a = key[s], b = key[s | js];
            unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
            key[s] = desc ? hi : lo;
            key[s | js] = desc ? lo : hi;
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < NPL; s++) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[s], j);
          const int e = s * 32 + lane;
          const bool lower = (lane & j) == 0;
          const bool desc = ((e & k) == 0);
          const bool keep_max = (lower == desc);
          const unsigned long long mx = key[s] > other ? key[s] : other;
          const unsigned long long mn = key[s] > other ? other : key[s];
          key[s] = keep_max ? mx : mn;
        }
      }
    }
  }
}

// Warp-cooperative compaction of one candidate list: keep the ksel best of its n entries, sorted.
// Returns (to every lane) the value bits of the ksel-th entry, or 0xFFFFFFFF if n < ksel.
__device__ __noinline__ uint32_t warp_compact_list(uint2* list, int n, int ksel, int lane) {
  unsigned long long key[LCAP / 32];
#pragma unroll
  for (int s = 0; s < LCAP / 32; s++) {
    const int e = s * 32 + lane;
    key[s] = 0ull;
    if (e < n) {
      const uint2 x = __ldcg(list + e);
      key[s] = make_key(x.x, x.y);
    }
  }
  warp_sort_desc<LCAP / 32>(key, lane);
  const int keep = n < ksel ? n : ksel;
  uint32_t kth = 0xFFFFFFFFu;
#pragma unroll
  for (int s = 0; s < LCAP / 32; s++) {
    const int e = s * 32 + lane;
    if (e < keep) __stcg(list + e, key_entry(key[s]));
    // the ksel-th entry lives in slot s = (ksel-1)/32 of lane (ksel-1)%32
    const uint32_t v = ord2f((uint32_t)(key[s] >> 32));
    const uint32_t got = __shfl_sync(0xffffffffu, v, (ksel - 1) & 31);
    if (s == ((ksel - 1) >> 5) && n >= ksel) kth = got;
  }
  return kth;
}

// Warp-cooperative threshold raise of one candidate list (the common path; no sort): a bitwise
// radix search finds a value T with ksel <= #{v >= T} <= ksel + 32 (or the exact ksel-th value when
// ties make that impossible), entries below T are dropped and the survivors are packed to the front,
// unsorted.  Never cuts inside a group of equal values, so ties need no index rule here.
// Returns T in the ordered-uint domain; *new_n = survivors.
__device__ __noinline__ uint32_t warp_select_list(uint2* list, int n, int ksel, int lane, int* new_n) {
  uint32_t ord[LCAP / 32], idv[LCAP / 32];
#pragma unroll
  for (int s = 0; s < LCAP / 32; s++) {
    const int e = s * 32 + lane;
    ord[s] = 0u;
    idv[s] = 0u;
    if (e < n) {
      const uint2 x = __ldcg(list + e);
      ord[s] = f2ord(x.x);
      idv[s] = x.y;
    }
  }
  uint32_t T = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 0; bit--) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int s = 0; s < LCAP / 32; s++) c += (ord[s] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= ksel) {
      T = cand;
      if (c <= ksel + 32) break;
    }
  }
  __syncwarp();
  int base = 0;
#pragma unroll
  for (int s = 0; s < LCAP / 32; s++) {
    const bool keep = ord[s] >= T && ord[s] != 0u;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) __stcg(list + base + __popc(b & ((1u << lane) - 1u)), make_uint2(ord2f(ord[s]), idv[s]));
    base += __popc(b);
  }
  *new_n = base;
  return T;
}

// ------------------------------------------------------------------------------------------------
// K3: Sketch . Sketch^T on tcgen05 with fused min-over-depth and candidate selection
// ------------------------------------------------------------------------------------------------
// TMEM columns: ACC_STAGES accumulators of BN FP32 columns, then (HAS_MIN) the running min.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BN, int ACC_STAGES, bool HAS_MIN>
// Register cap instead of launch bounds (the two qualifiers exclude each other): the d = 1 kernel compiles to 128
// registers under it, the d > 1 kernel to 144, which leaves 16 K / 10 K of the SM's 64 K registers free.  The persistent
// CTAs hold every SM for the whole sweep, and whatever else must run beside them in the pull-gather mode gets only what
// they leave: with 138 / 150 registers (12.5 K / 7.9 K free) the fused form ran on every box, with the 161 - 168 the
// compiler took after the lists grew to 512 entries a peer block never arrived while K3 was running
// (profiles/r2_bench_n2_final_code_fused_timeout.log).  No variant in use spills under the cap.
__global__ void __maxnreg__(144)
k_cosine(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
         const CosParams p) {
  static_assert((ACC_STAGES + (HAS_MIN ? 1 : 0)) * BN <= 512, "TMEM budget");
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int MAX_STAGES = 8;
  constexpr int HALF = BN / 2;        // columns per epilogue thread
  constexpr int CHUNKS = HALF / 32;   // 32-column TMEM loads per thread per step

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[ACC_STAGES];
  __shared__ __align__(8) uint64_t bar_tempty[ACC_STAGES];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < ACC_STAGES; s++) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 8);  // one arrival per epilogue warp
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
        const int4 it = p.items[w];
        const int t0 = it.y, t1 = it.z;
        int g_seen = -1;
        for (int t = t0; t < t1; t += it.w) {
          const int gs = t / p.tiles_per_block;
          const int l0 = (t - gs * p.tiles_per_block) * BN;
          int g = gs + p.rot;
          if (g >= p.num_blocks) g -= p.num_blocks;
          if (p.ready != nullptr && g != g_seen) {
            // the block is still in flight from its owner: wait for the copy stream's signal
            const volatile uint32_t* flag = p.ready + g;
            if (*flag != p.epoch) {
              const long long t_wait = clock64();
              while (*flag != p.epoch) {
                __nanosleep(256);
                // ~4 s: report instead of hanging the GPU.  The sweep then runs on incomplete data, but the abort
                // word makes mb200_cosine_finish fail with MB200_ERR_CUDA before any result is handed out -- and
                // every other wait of this launch gives up at once (the word is sticky until finish clears it), so a
                // lost block costs one time-out, not one per work item
                if (clock64() - t_wait > 8000000000LL || *reinterpret_cast<volatile uint32_t*>(p.abort_flag) != 0u) {
                  atomicExch(p.abort_flag, 1u);
                  break;
                }
              }
            }
            __threadfence_system();
            asm volatile("fence.proxy.async;" ::: "memory");
            g_seen = g;
          }
          for (int dep = 0; dep < p.depth; dep++) {
            for (int kb = 0; kb < p.kblocks; kb++) {
              mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
              const uint32_t full = smem_u32(&bar_full[stage]);
              const uint32_t sa = smem_base + (uint32_t)stage * STAGE_BYTES;
              mbar_expect_tx(full, STAGE_BYTES);
              tma_load_3d(sa, &tmA, full, kb * BK, it.x * BM, dep, p.policy_a);
              tma_load_4d(sa + A_BYTES, &tmB, full, kb * BK, l0, dep, g, p.policy_b);
              if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    uint32_t q = 0;  // accumulation steps issued by this CTA
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
      const int4 it = p.items[w];
      const int t0 = it.y, t1 = it.z;
      for (int t = t0; t < t1; t += it.w) {
        for (int dep = 0; dep < p.depth; dep++, q++) {
          const uint32_t as = q % ACC_STAGES;
          const uint32_t aphase = (q / ACC_STAGES) & 1u;
          mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
          fence_after_sync();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = 0; kb < p.kblocks; kb++) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            fence_after_sync();
            if (lane == 0) {
              const uint32_t sa = smem_base + (uint32_t)stage * STAGE_BYTES;
              const uint64_t da = umma_desc_k128(sa);
              const uint64_t db = umma_desc_k128(sa + A_BYTES);
#pragma unroll
              for (int k = 0; k < BK / 16; k++) {
                // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc,
                         (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(smem_u32(&bar_empty[stage]));
              if (kb == p.kblocks - 1) umma_commit(smem_u32(&bar_tfull[as]));
            }
            __syncwarp();
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: min over depth + candidate selection =====================
    const int ew = warp - 4;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const int half = ew >> 2;      // column half of the tile this thread scans
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t q = 0;
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
      const int4 it = p.items[w];
      const int t0 = it.y, t1 = it.z;
      const long long grow = (long long)it.x * BM + row;
      const bool row_ok = grow < p.a_count;
      const bool band = p.row_cut != nullptr;
      const uint32_t my_id = (row_ok && p.exclude_self)
                                 ? (p.a_ids ? __ldg(p.a_ids + grow) : (uint32_t)grow * p.a_id_mul + p.a_id_off)
                                 : 0xFFFFFFFFu;
      uint32_t rv = 0;  // bit dep set <=> (my row, dep) has a non-zero norm
      if (row_ok) {
        for (int dep = 0; dep < p.depth; dep++)
          rv |= ((__ldg(p.a_valid + (size_t)dep * p.a_vw + (grow >> 5)) >> (grow & 31)) & 1u) << dep;
      }
      const size_t slot = ((size_t)w * 2 + half) * BM + row;
      uint2* list = p.lists + slot * LCAP;
      int cnt = 0;
      const unsigned long long ceil_key = (p.row_ceil != nullptr && row_ok) ? __ldg(p.row_ceil + grow) : ~0ull;
      float thr = (band && row_ok) ? __ldg(p.row_cut + grow) : p.thr_init;
      float bound = -INFINITY;
      uint32_t* my_row_thr = p.row_thr + (size_t)it.x * BM + row;
      // Threshold raises: warp-cooperative, and only AFTER the accumulator stage has been handed back (a raise costs
      // microseconds and the lanes of a warp tend to need theirs together: inside the chunk loop they stalled the
      // MMA warp -- K3 lost 17 % between k = 10 and k = 100).  Optional raises (the list holds `wm` entries: 2 * ksel,
      // so that a fresh threshold is published early) are rationed to two per accumulation step and warp, which fits
      // in the shadow of the next step's MMAs; a list that has no room left for a whole tile (HALF entries) is
      // raised at once, and at the end of the sweep every list is cut down to what the merge will want.
      const int wm = min(max(2 * p.ksel, 128), LCAP - HALF);
      auto raise_thresholds = [&](bool tile_done, bool sweep_done) {
        const bool must = (tile_done && cnt > LCAP - HALF) || (sweep_done && cnt > p.ksel + 32);
        unsigned need_must = __ballot_sync(0xffffffffu, must);
        unsigned need_opt = __ballot_sync(0xffffffffu, cnt >= wm && !must);
        int budget = 2;
        while (need_must != 0u || (budget > 0 && need_opt != 0u)) {
          int src;
          if (need_must != 0u) {
            src = __ffs(need_must) - 1;
            need_must &= need_must - 1;
          } else {
            src = __ffs(need_opt) - 1;
            need_opt &= need_opt - 1;
            budget--;
          }
          __syncwarp();
          uint2* l2 = (uint2*)__shfl_sync(0xffffffffu, (unsigned long long)list, src);
          const int n2 = __shfl_sync(0xffffffffu, cnt, src);
          int n3 = 0;
          const uint32_t T = warp_select_list(l2, n2, p.ksel, lane, &n3);
          uint32_t kth = 0xFFFFFFFFu;
          if (n3 > p.ksel + 32) {  // a large group of equal values: cut it by index (exact sort)
            __syncwarp();
            kth = warp_compact_list(l2, n3, p.ksel, lane);
          }
          if (lane == src) {
            if (kth != 0xFFFFFFFFu) {
              cnt = min(n3, p.ksel);
              const float kv = __uint_as_float(kth);
              bound = fmaxf(bound, kv);
              // ties at the k-th value may still win on the index unless the scan order is
              // index-monotone: admit them by stepping the threshold one ulp down
              const uint32_t ko = f2ord(kth);
              const uint32_t to = p.nonstrict && ko > 0 ? ko - 1 : ko;
              thr = fmaxf(thr, __uint_as_float(ord2f(to)));
              // other lists of the row scan other index ranges: they must keep admitting ties
              atomicMax(my_row_thr, ko > 0 ? ko - 1 : 0u);
            } else if (T != 0u) {
              cnt = n3;
              bound = fmaxf(bound, __uint_as_float(ord2f(T)));
              thr = fmaxf(thr, __uint_as_float(ord2f(T - 1)));  // x > thr  <=>  x >= T
              atomicMax(my_row_thr, T - 1);
            }
          }
          __syncwarp();
        }
      };
      for (int t = t0; t < t1; t += it.w) {
        const int gs = t / p.tiles_per_block;
        const int l0 = (t - gs * p.tiles_per_block) * BN + half * HALF;
        int g = gs + p.rot;
        if (g >= p.num_blocks) g -= p.num_blocks;
        {
          // lower bounds published by the other lists of this row (other half, other column chunks)
          const uint32_t gt = __ldcg(my_row_thr);
          if (gt != 0u) thr = fmaxf(thr, __uint_as_float(ord2f(gt)));
        }
        for (int dep = 0; dep < p.depth; dep++, q++) {
          const uint32_t as = q % ACC_STAGES;
          const uint32_t aphase = (q / ACC_STAGES) & 1u;
          // column validity of my 32-column chunks (uniform per warp), fetched ahead of the wait --
          // except in pull-gather mode, where the words arrive with the block
          uint32_t cm[CHUNKS];
          const uint32_t* bv = p.b_valid + ((size_t)g * p.depth + dep) * p.b_vw + (l0 >> 5);
          // (if the block has already landed they can be fetched ahead as well)
          // acquire at system scope: the validity words of the block (written by a peer's DMA, published by the
          // copy stream's flag write) must not be read ahead of the flag
          const bool early = p.ready == nullptr || ld_acquire_sys(p.ready + g) == p.epoch;
          if (early) {
#pragma unroll
            for (int c = 0; c < CHUNKS; c++) cm[c] = p.ready == nullptr ? __ldg(bv + c) : __ldcv(bv + c);
          }
          const bool my_valid = (rv >> dep) & 1u;
          const bool last = dep == p.depth - 1;
          mbar_wait(smem_u32(&bar_tfull[as]), aphase);
          fence_after_sync();
          if (!early) {
#pragma unroll
            for (int c = 0; c < CHUNKS; c++) cm[c] = __ldcv(bv + c);
          }
          const uint32_t acc_addr = tmem_base + lane_addr + as * BN + half * HALF;
          const uint32_t min_addr = tmem_base + lane_addr + ACC_STAGES * BN + half * HALF;
          // Candidate selection over one 32-column chunk of final values (executed per thread = per row).
          auto select_chunk = [&](int c, uint32_t (&v)[32]) {
            if (p.dense_out != nullptr && row_ok) {
              float* dst = p.dense_out + (size_t)grow * p.dense_ld + (size_t)t * BN + half * HALF + c * 32;
#pragma unroll
              for (int j = 0; j < 32; j++) dst[j] = __uint_as_float(v[j]) * p.inv_scale2;
            }
            const uint32_t id0 = (uint32_t)(l0 + c * 32) * p.b_id_mul + (uint32_t)g * p.b_id_add + p.b_id_base;
            if (band) {
              if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                  const float x = __uint_as_float(v[j]);
                  const uint32_t id = id0 + (uint32_t)j * p.b_id_mul;
                  if (x > thr && id != my_id) {
                    const int at = atomicAdd(p.band_cnt + grow, 1);
                    if (at < p.band_cap) {
                      p.band_cand[(size_t)grow * p.band_cap + at] = id;
                      p.band_val[(size_t)grow * p.band_cap + at] = x;
                    }
                  }
                }
              }
              return;
            }
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const float x = __uint_as_float(v[j]);
              if (x > thr) {  // NaN never passes
                const uint32_t id = id0 + (uint32_t)j * p.b_id_mul;
                if (id != my_id && make_key(v[j], id) < ceil_key) {
                  __stcg(list + cnt, make_uint2(v[j], id));
                  cnt++;
                }
              }
            }
          };
#pragma unroll
          for (int c = 0; c < CHUNKS; c++) {
            uint32_t v[32];
            uint32_t m[32];
            tmem_ld32(acc_addr + c * 32, v);
            if (HAS_MIN && dep > 0) tmem_ld32(min_addr + c * 32, m);
            tmem_wait_ld();
            const uint32_t cmask = my_valid ? cm[c] : 0u;
#pragma unroll
            for (int j = 0; j < 32; j++) {
              float x = ((cmask >> j) & 1u) ? __uint_as_float(v[j]) : __int_as_float(0x7FC00000);
              // fminf returns the non-NaN operand: NaN is the identity of the running min
              if (HAS_MIN && dep > 0) x = fminf(__uint_as_float(m[j]), x);
              v[j] = __float_as_uint(x);
            }
            // With a min buffer the final values are parked there too, so that the accumulator goes
            // back to the MMA warp before the (long, data-dependent) selection starts.
            if (HAS_MIN) tmem_st32(min_addr + c * 32, v);
            else select_chunk(c, v);
          }
          if (HAS_MIN) tmem_wait_st();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[as]));
          if (HAS_MIN && last) {
            // selection from the min buffer; this thread is its only reader and writer, and it will
            // not touch it again before depth 0 of the next tile
#pragma unroll
            for (int c = 0; c < CHUNKS; c++) {
              uint32_t v[32];
              tmem_ld32(min_addr + c * 32, v);
              tmem_wait_ld();
              select_chunk(c, v);
            }
          }
          if (!band) raise_thresholds(last, last && t + it.w >= t1);
        }
      }
      // the list stays unsorted (<= LCAP entries); K5 merges and orders
      p.list_cnt[slot] = cnt;
      p.list_bound[slot] = bound;
    }
  }

  __syncthreads();
  if (warp == 2) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// K5a: merge the candidate lists of one row (all column chunks, both halves); one warp per row.
// TENSOR precision: write the top-k straight from the tensor-core values.
// RESCORED: write the merged candidates for k_rescore.
// ------------------------------------------------------------------------------------------------
struct MergeParams {
  const uint2* lists;
  const int32_t* list_cnt;
  const float* list_bound;
  const uint32_t* row_thr;  // [rows padded] (see CosParams)
  const int32_t* slot_ptr;  // [num_m + 1]: the work items (= list slots) of row block m are
  const int32_t* slot_of;   // slot_of[slot_ptr[m] .. slot_ptr[m+1])
  int64_t a_count;
  int32_t ksel, k;
  double threshold;        // admitted iff sim >= threshold && sim > 0
  float inv_scale2;
  int32_t rescored;
  // carry state of a multi-push job: [rows][CAP] sorted keys (0 = empty) + bound
  const unsigned long long* carry_in;
  const float* carry_bound_in;
  unsigned long long* carry_out;
  float* carry_bound_out;
  int32_t emit;            // 0: only update the carry state
  // TENSOR outputs
  long long* out_idx;
  double* out_sim;
  int32_t* out_cnt;
  // RESCORED outputs
  uint32_t* cand_id;       // [a_count][CAP]
  float* cand_val;         // [a_count][CAP] scaled tensor values of the candidates (CERTIFIED), or NULL
  int32_t* cand_cnt;       // [a_count]
  float* cand_bound;       // [a_count] (scaled tensor value; -inf = nothing was ever dropped)
};

__global__ void __launch_bounds__(256) k_merge(const MergeParams p) {
  // survivors of the row-threshold filter wait here until CAP of them are pending; one 512-key sort
  // then folds them into the running best (so the cost does not grow with the number of lists)
  __shared__ unsigned long long s_pend[8][CAP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= p.a_count) return;
  const int m = (int)(r / BM), rr = (int)(r % BM);
  unsigned long long key[2 * CAP / 32];
#pragma unroll
  for (int s = 0; s < 2 * CAP / 32; s++) key[s] = 0ull;
  float bound = -INFINITY;
  if (p.carry_in != nullptr) {
#pragma unroll
    for (int u = 0; u < CAP / 32; u++) key[u] = p.carry_in[(size_t)r * CAP + u * 32 + lane];
    bound = p.carry_bound_in[r];
  }
  // Every value <= row_thr is below the row's ksel-th best (K3 proved it when it published the
  // threshold, and the publishing list's bound covers everything dropped for that reason).
  uint32_t rthr = p.row_thr[r];
  unsigned long long* pend = s_pend[wib];
  int npend = 0;
  bool have_best = p.carry_in != nullptr;
  auto fold = [&]() {
    if (!have_best) {
      // first fold of the row (the only one for most rows): nothing to merge with, sort the
      // pending keys alone -- 256 keys instead of 512
      unsigned long long k8[CAP / 32];
#pragma unroll
      for (int u = 0; u < CAP / 32; u++) {
        const int e = u * 32 + lane;
        k8[u] = e < npend ? pend[e] : 0ull;
      }
      __syncwarp();
      warp_sort_desc<CAP / 32>(k8, lane);
#pragma unroll
      for (int u = 0; u < CAP / 32; u++) {
        key[u] = k8[u];
        key[CAP / 32 + u] = 0ull;
      }
    } else {
#pragma unroll
      for (int u = 0; u < CAP / 32; u++) {
        const int e = u * 32 + lane;
        key[CAP / 32 + u] = e < npend ? pend[e] : 0ull;
      }
      __syncwarp();
      warp_sort_desc<2 * CAP / 32>(key, lane);
    }
    have_best = true;
    npend = 0;
    // truncate to ksel; the best value dropped here bounds everything dropped later as well
    unsigned long long kfirst = 0ull;  // first key beyond the cut (ksel < CAP always)
#pragma unroll
    for (int u = 0; u < CAP / 32; u++) {
      const unsigned long long kv = __shfl_sync(0xffffffffu, key[u], p.ksel & 31);
      if (u == (p.ksel >> 5)) kfirst = kv;
    }
    if (kfirst != 0ull) {
      const uint32_t ko = (uint32_t)(kfirst >> 32);
      bound = fmaxf(bound, __uint_as_float(ord2f(ko)));
      // from now on anything strictly below the dropped value cannot make the cut
      if (ko > 0u && ko - 1u > rthr) rthr = ko - 1u;
    }
#pragma unroll
    for (int u = 0; u < 2 * CAP / 32; u++)
      if (u * 32 + lane >= p.ksel) key[u] = 0ull;
  };
  for (int s = p.slot_ptr[m]; s < p.slot_ptr[m + 1]; s++) {
    const int w = p.slot_of[s];
    for (int h = 0; h < 2; h++) {
      const size_t slot = ((size_t)w * 2 + h) * BM + rr;
      const int n = p.list_cnt[slot];
      bound = fmaxf(bound, p.list_bound[slot]);
      const uint2* list = p.lists + slot * LCAP;
      for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        bool keep = false;
        uint2 x = make_uint2(0u, 0u);
        if (e < n) {
          x = __ldcg(list + e);
          keep = f2ord(x.x) > rthr;
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (b == 0u) continue;
        if (npend + 32 > CAP) fold();
        if (keep) pend[npend + __popc(b & ((1u << lane) - 1u))] = make_key(x.x, x.y);
        npend += __popc(b);
        __syncwarp();
      }
    }
  }
  fold();
  if (p.carry_out != nullptr) {
#pragma unroll
    for (int u = 0; u < CAP / 32; u++) p.carry_out[(size_t)r * CAP + u * 32 + lane] = key[u];
    if (lane == 0) p.carry_bound_out[r] = bound;
  }
  if (!p.emit) return;
  // key[0 .. CAP/32) now holds the row's best <= ksel candidates, sorted (value desc, index asc)
  if (p.rescored) {
    int n = 0;
#pragma unroll
    for (int u = 0; u < CAP / 32; u++) {
      const int e = u * 32 + lane;
      const bool have = key[u] != 0ull;
      p.cand_id[(size_t)r * CAP + e] = have ? ~(uint32_t)key[u] : 0xFFFFFFFFu;
      if (p.cand_val) p.cand_val[(size_t)r * CAP + e] = have ? __uint_as_float(ord2f((uint32_t)(key[u] >> 32))) : 0.0f;
      n += __popc(__ballot_sync(0xffffffffu, have));
    }
    if (lane == 0) {
      p.cand_cnt[r] = n;
      p.cand_bound[r] = bound;
    }
    return;
  }
  int admitted = 0;
#pragma unroll
  for (int u = 0; u < CAP / 32; u++) {
    const int e = u * 32 + lane;
    const uint2 x = key_entry(key[u]);
    const double sim = (double)(__uint_as_float(x.x) * p.inv_scale2);
    const bool ok = key[u] != 0ull && e < p.k && sim >= p.threshold && sim > 0.0;
    if (e < p.k) {
      p.out_idx[(size_t)r * p.k + e] = ok ? (long long)x.y : -1LL;
      p.out_sim[(size_t)r * p.k + e] = ok ? sim : 0.0;
    }
    admitted += __popc(__ballot_sync(0xffffffffu, ok));
  }
  if (lane == 0) p.out_cnt[r] = admitted;
}

// ------------------------------------------------------------------------------------------------
// K5b: exact re-score of the merged candidates.  One CTA per row.
//
// The reference sums xa*xa, xb*xb, xa*xb sequentially in FP64.  Counters are integers (in quanta),
// so while sum|xa*xb| < 2^53 every partial sum is an exactly representable integer and the
// order of summation cannot matter: the integer dot products below reproduce the reference's
// three sums bit for bit, and sqrt / mul / div are IEEE-correct on both sides.  Rows where that
// precondition fails are flagged and go through k_exact_rows (sequential FP64).
// ------------------------------------------------------------------------------------------------
struct RescoreParams {
  const long long* a_counters;  // [a_count][d][W]
  const long long* b_counters;  // [blocks][b_count][d][W]
  const long long* const* b_blocks_ptr;  // or: one pointer per block (peer-mapped banks), each [b_count][d][W]
  const int* const* b_blocks32;          // CERTIFIED: optional int32 copies of the blocks (INT_MIN = does not fit)
  int64_t a_count, b_count;
  int32_t d, W, blocks;
  uint32_t a_id_mul, a_id_off, b_id_mul, b_id_add;
  const uint32_t* cand_id;
  const int32_t* cand_cnt;
  const float* cand_bound;
  const float* cand_val;       // CERTIFIED: tensor values of the candidates
  float inv_scale2;
  float eps_rel;               // relative error bound of the tensor-core values
  float eps_abs;               // absolute error bound (similarity units); = eps_rel for mixed-sign counters
  int mixed;                   // counters may be negative (mb200_cosine_args.mixed_sign)
  int32_t k;
  double threshold;
  long long* out_idx;
  double* out_sim;
  int32_t* out_cnt;
  int32_t* row_flag;           // [a_count]: 1 = not certified (band pass, else exact full-row path), 2 = outside
                               // the exact-integer range (exact full-row path)
  int32_t* flag_count;
  float* row_cut;              // [a_count] scaled tensor value below which a column cannot be in the row's top-k
  float thr_floor;             // the job's admission threshold (scaled)
};

__device__ __forceinline__ void b_locate(const RescoreParams& p, uint32_t id, long long& g, long long& l) {
  if (p.b_id_add == 1 && p.b_id_mul != 1) {  // interleaved shards: id = l * G + g
    g = id % p.b_id_mul;
    l = id / p.b_id_mul;
  } else {                                   // contiguous blocks: id = l + g * b_count
    g = p.b_id_add ? id / p.b_id_add : 0;
    l = id - g * p.b_id_add;
  }
}

// first counter of sketch row (block g, local row l, depth i)
__device__ __forceinline__ const long long* b_row(const RescoreParams& p, long long g, long long l, int i) {
  const long long* base = p.b_blocks_ptr ? p.b_blocks_ptr[g] : p.b_counters + (size_t)g * p.b_count * p.d * p.W;
  return base + ((size_t)l * p.d + i) * p.W;
}

#define TWO53 9007199254740992.0
#define JAVA_MAX_DOUBLE 1.7976931348623157e308

// Ordering, certification and output of one row's re-scored candidates; executed by one warp.
template <int NPL>   // 32 * NPL >= n sorted entries
__device__ __forceinline__ void rescore_finish_t(const RescoreParams& p, long long r, int n, const double* s_min,
                                                 int bad_flag, int lane, bool certified_elsewhere) {
  // warp 0: order the candidates by (exact sim desc, index asc), certify, write the top-k
  double sim[NPL];
  uint32_t idv[NPL];
#pragma unroll
  for (int u = 0; u < NPL; u++) {
    const int c = u * 32 + lane;
    sim[u] = -INFINITY;
    idv[u] = 0xFFFFFFFFu;
    if (c < n) {
      const double s = s_min[c];
      idv[u] = p.cand_id[(size_t)r * CAP + c];
      // admitted iff not NaN (no comparable row), >= threshold and > Double.MIN_VALUE
      if (s != JAVA_MAX_DOUBLE && s >= p.threshold && s > 4.9e-324) sim[u] = s;
    }
  }
  // bitonic sort on (sim desc, id asc) with 96-bit keys
#pragma unroll
  for (int k = 2; k <= 32 * NPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int js = j >> 5;
#pragma unroll
        for (int s = 0; s < NPL; s++) {
          if ((s & js) == 0) {
            const bool desc = (((s * 32) & k) == 0);
            const bool a_first = sim[s] > sim[s | js] || (sim[s] == sim[s | js] && idv[s] <= idv[s | js]);
            if (a_first != desc) {
              double ts = sim[s]; sim[s] = sim[s | js]; sim[s | js] = ts;
              uint32_t ti = idv[s]; idv[s] = idv[s | js]; idv[s | js] = ti;
            }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < NPL; s++) {
          const double os = __shfl_xor_sync(0xffffffffu, sim[s], j);
          const uint32_t oi = __shfl_xor_sync(0xffffffffu, idv[s], j);
          const int e = s * 32 + lane;
          const bool lower = (lane & j) == 0;
          const bool desc = ((e & k) == 0);
          const bool mine_first = sim[s] > os || (sim[s] == os && idv[s] <= oi);
          const bool keep_first = (lower == desc);
          if (mine_first != keep_first) {
            sim[s] = os;
            idv[s] = oi;
          }
        }
      }
    }
  }
  int admitted = 0;
  double kth = -INFINITY;  // exact value of the k-th result (or -inf if fewer than k)
#pragma unroll
  for (int u = 0; u < NPL; u++) {
    const int e = u * 32 + lane;
    const bool ok = sim[u] > -INFINITY && e < p.k;
    if (e < p.k) {
      p.out_idx[(size_t)r * p.k + e] = ok ? (long long)idv[u] : -1LL;
      p.out_sim[(size_t)r * p.k + e] = ok ? sim[u] : 0.0;
    }
    admitted += __popc(__ballot_sync(0xffffffffu, ok));
    const double v = __shfl_sync(0xffffffffu, sim[u], (p.k - 1) & 31);
    if (u == ((p.k - 1) >> 5)) kth = v;
  }
  for (int e = 32 * NPL + lane; e < p.k; e += 32) {   // k beyond the sorted window: nothing there
    p.out_idx[(size_t)r * p.k + e] = -1LL;
    p.out_sim[(size_t)r * p.k + e] = 0.0;
  }
  if (lane == 0) {
    p.out_cnt[r] = admitted;
    // every candidate that is not in the merged list has tensor value <= bound, hence exact value
    // <= bound * (1 + eps) (+ eps absolute for values near zero).  The top-k is certain iff the
    // k-th exact value clears that.
    const float b = p.cand_bound[r];
    int flag = bad_flag;
    if (b > -INFINITY && !certified_elsewhere) {
      const double bs = (double)(b * p.inv_scale2);
      const double ub = bs + fabs(bs) * (double)p.eps_rel + (double)p.eps_abs;
      if (!(kth > ub) && flag == 0) flag = 1;
    }
    p.row_flag[r] = flag;
    if (flag) atomicAdd(p.flag_count, 1);
    if (flag == 1 && p.row_cut) {
      // The true k-th similarity is at least kth * (1 - eps) - eps_abs (kth may itself be a tensor value); a
      // column whose tensor value is below cut has an exact value below cut * (1 + eps) + eps_abs, which is
      // below that.  Fewer than k results so far: everything above the job's threshold is wanted.
      float cut = p.thr_floor;
      if (kth > -INFINITY) {
        const double c = (kth * (1.0 - 2.0 * (double)p.eps_rel) - 2.0 * (double)p.eps_abs) / (double)p.inv_scale2;
        cut = fmaxf(cut, nextafterf((float)c, -INFINITY));
      }
      p.row_cut[r] = nextafterf(cut, -INFINITY);
    }
  }
}

// the sort network is sized by the number of candidates (k = 50 keeps 64: a quarter of the CAP-wide sort)
__device__ __forceinline__ void rescore_finish(const RescoreParams& p, long long r, int n, const double* s_min,
                                               int bad_flag, int lane, bool certified_elsewhere = false) {
  if (n <= 64 && p.k <= 64) rescore_finish_t<2>(p, r, n, s_min, bad_flag, lane, certified_elsewhere);
  else if (n <= 128 && p.k <= 128) rescore_finish_t<4>(p, r, n, s_min, bad_flag, lane, certified_elsewhere);
  else rescore_finish_t<CAP / 32>(p, r, n, s_min, bad_flag, lane, certified_elsewhere);
}

#define RESCORE_SEG 4096  /* counters of one A row segment staged in shared memory */

__global__ void __launch_bounds__(256) k_rescore(const RescoreParams p) {
  __shared__ long long s_a[RESCORE_SEG];       // one segment of one depth row of A
  __shared__ double s_min[CAP];                // running min of cos_i per candidate
  __shared__ long long s_ab[CAP], s_bb[CAP];   // exact integer dot products of the current depth
  __shared__ double s_mag[CAP];
  __shared__ double s_red[8];
  __shared__ long long s_redi[8];
  __shared__ int s_bad;
  const long long r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.cand_cnt[r];
  for (int c = tid; c < CAP; c += blockDim.x) s_min[c] = JAVA_MAX_DOUBLE;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const long long* arow = p.a_counters + (size_t)r * p.d * p.W;
  for (int i = 0; i < p.d && n > 0; i++) {
    for (int c = tid; c < CAP; c += blockDim.x) {
      s_ab[c] = 0;
      s_bb[c] = 0;
      s_mag[c] = 0.0;
    }
    long long aa = 0;
    double amag = 0.0;
    int bad = 0;
    for (int j0 = 0; j0 < p.W; j0 += RESCORE_SEG) {
      const int seg = min(RESCORE_SEG, p.W - j0);
      __syncthreads();  // previous segment fully consumed
      for (int j = tid; j < seg; j += blockDim.x) {
        const long long x = arow[(size_t)i * p.W + j0 + j];
        s_a[j] = x;
        if (x >= (1LL << 31) || x <= -(1LL << 31)) bad = 1;
        aa += x * x;
        amag += (double)x * (double)x;
      }
      __syncthreads();
      for (int c = warp; c < n; c += 8) {
        const uint32_t id = p.cand_id[(size_t)r * CAP + c];
        long long g, l;
        b_locate(p, id, g, l);
        const long long* brow = b_row(p, g, l, i) + j0;
        long long bb = 0, ab = 0;
        double mag = 0.0;
        int badb = 0;
        for (int j = lane; j < seg; j += 32) {
          const long long y = __ldg(brow + j);
          const long long x = s_a[j];
          if (y >= (1LL << 31) || y <= -(1LL << 31)) badb = 1;
          bb += y * y;
          ab += x * y;
          mag += fabs((double)y * (double)y) + fabs((double)x * (double)y);
        }
        for (int o = 16; o > 0; o >>= 1) {
          bb += __shfl_xor_sync(0xffffffffu, bb, o);
          ab += __shfl_xor_sync(0xffffffffu, ab, o);
          mag += __shfl_xor_sync(0xffffffffu, mag, o);
          badb |= __shfl_xor_sync(0xffffffffu, badb, o);
        }
        if (lane == 0) {  // candidate c belongs to this warp for every segment: no race
          s_bb[c] += bb;
          s_ab[c] += ab;
          s_mag[c] += mag;
          if (badb) s_bad = 1;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      aa += __shfl_xor_sync(0xffffffffu, aa, o);
      amag += __shfl_xor_sync(0xffffffffu, amag, o);
    }
    if (lane == 0) {
      s_redi[warp] = aa;
      s_red[warp] = amag;
    }
    if (bad) s_bad = 1;
    __syncthreads();
    long long AA = 0;
    double AAm = 0.0;
    for (int w = 0; w < 8; w++) {
      AA += s_redi[w];
      AAm += s_red[w];
    }
    if (AAm >= TWO53 * 0.5 && tid == 0) s_bad = 1;
    const double sqa = sqrt((double)AA);
    for (int c = tid; c < n; c += blockDim.x) {
      if (s_mag[c] >= TWO53 * 0.5) s_bad = 1;
      const double den = __dmul_rn(sqa, sqrt((double)s_bb[c]));
      if (den != 0.0) {
        const double cs = __ddiv_rn((double)s_ab[c], den);
        s_min[c] = cs < s_min[c] ? cs : s_min[c];  // Math.min, no NaN possible here
      }
    }
    __syncthreads();
  }
  if (warp == 0) rescore_finish(p, r, n, s_min, s_bad ? 2 : 0, lane);
}

// ------------------------------------------------------------------------------------------------
// K5b, CERTIFIED form: the top-k SET is exact, the similarities are the tensor-core values (<= 2^-10
// relative) -- the contract of the north star at a fraction of the full re-score.  With value bounds
// L = v(1-eps), U = v(1+eps) a candidate is surely in when fewer than k others can beat it and
// surely out when k others surely do; in the sorted list only the entries
//     rank <  k with v <= v[k]   / (1 - 2 eps)      (v[k]   = the (k+1)-th tensor value)
//     rank >= k with v >= v[k-1] / (1 + 2 eps)      (v[k-1] = the k-th)
// are uncertain -- usually none or a handful -- and only those (plus entries within 2 eps of the
// threshold) are re-scored exactly from the int64 counters.  If the uncertain band reaches what was
// dropped earlier (bound) the row takes the exact full-row path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_certify(const RescoreParams p) {
  __shared__ long long s_a[RESCORE_SEG];
  __shared__ double s_fin[CAP];                // final value per candidate (tensor or exact)
  __shared__ double s_min[CAP];                // exact running min of the selected candidates
  __shared__ long long s_ab[CAP], s_bb[CAP];
  __shared__ int s_bmax[CAP];
  __shared__ int s_sel[CAP];                   // candidate slots that need the exact value
  __shared__ int s_redm[8];
  __shared__ long long s_redi[8];
  __shared__ int s_bad, s_nsel, s_flag;
  const long long r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.cand_cnt[r];
  const float* val = p.cand_val + (size_t)r * CAP;   // sorted descending
  if (tid == 0) {
    s_bad = 0;
    s_nsel = 0;
    s_flag = 0;
  }
  __syncthreads();
  const float e2 = 2.0f * p.eps_rel;
  const int k = p.k;
  const float vk = n > 0 ? val[min(k, n) - 1] : 0.0f;                 // k-th (or last) tensor value
  const float vk1 = n > k ? val[k] : -INFINITY;                       // (k+1)-th
  // with mixed-sign counters the error of a tensor value is absolute (eps * sum|x_i y_i| <= eps): ea2 widens
  // every band by it, and the positivity cut (sim > Double.MIN_VALUE) becomes a threshold to certify too
  const float ea2 = p.mixed ? 2.0f * p.eps_abs / p.inv_scale2 : 0.0f;
  const float band_hi = vk1 + fabsf(vk1) * (e2 / (1.0f - e2)) + ea2;
  const float band_lo = vk - fabsf(vk) * (e2 / (1.0f + e2)) - ea2;
  const float thr = (float)((p.threshold > 0.0 ? p.threshold : 0.0) / (double)p.inv_scale2);   // scaled domain
  for (int c = tid; c < CAP; c += blockDim.x) {
    s_min[c] = JAVA_MAX_DOUBLE;
    if (c < n) {
      const float v = val[c];
      s_fin[c] = (double)(v * p.inv_scale2);
      const bool near_cut = (c < k && v <= band_hi) || (c >= k && v >= band_lo);
      const bool near_thr = (p.threshold > 0.0 || p.mixed) && fabsf(v - thr) <= e2 * thr + ea2 + 1e-30f;
      if (near_cut || near_thr) s_sel[atomicAdd(&s_nsel, 1)] = c;
    }
  }
  if (tid == 0) {
    // everything dropped before the merge must be surely out
    const float b = p.cand_bound[r];
    if (b > -INFINITY && n >= k && !(b < band_lo)) s_flag = 1;
    if (b > -INFINITY && n < k) s_flag = 1;
  }
  __syncthreads();
  const int nsel = s_nsel;
  if (nsel == 0 && !s_flag) {
    // nothing is uncertain: the merged order (tensor value desc, index asc) is the answer
    if (warp == 0) {
      int admitted = 0;
      for (int e0 = 0; e0 < k; e0 += 32) {
        const int e = e0 + lane;
        const bool ok = e < n && e < k && s_fin[e < CAP ? e : 0] >= p.threshold && s_fin[e < CAP ? e : 0] > 4.9e-324;
        if (e < k) {
          p.out_idx[(size_t)r * k + e] = ok ? (long long)p.cand_id[(size_t)r * CAP + e] : -1LL;
          p.out_sim[(size_t)r * k + e] = ok ? s_fin[e] : 0.0;
        }
        admitted += __popc(__ballot_sync(0xffffffffu, ok));
      }
      if (lane == 0) {
        p.out_cnt[r] = admitted;
        p.row_flag[r] = 0;
      }
    }
    return;
  }
  const long long* arow = p.a_counters + (size_t)r * p.d * p.W;
  for (int i = 0; i < p.d && nsel > 0 && !s_flag; i++) {
    for (int c = tid; c < nsel; c += blockDim.x) {
      s_ab[c] = 0;
      s_bb[c] = 0;
      s_bmax[c] = 0;
    }
    long long aa = 0;
    int amax = 0;
    int bad = 0;
    for (int j0 = 0; j0 < p.W; j0 += RESCORE_SEG) {
      const int seg = min(RESCORE_SEG, p.W - j0);
      __syncthreads();
      for (int j = tid; j < seg; j += blockDim.x) {
        const long long x = arow[(size_t)i * p.W + j0 + j];
        s_a[j] = x;
        // counters beyond 31 bits send the row to the exact full-row path; below that the products
        // are single IMAD.WIDEs and the running maxima bound every partial sum
        if (x >= (1LL << 31) || x <= -(1LL << 31)) bad = 1;
        const int xi = (int)x;
        aa += (long long)xi * xi;
        const int ax = xi < 0 ? -xi : xi;
        amax = ax > amax ? ax : amax;
      }
      __syncthreads();
      // few candidates: several warps share one (slices of the segment), so that all 8 warps keep loads
      // in flight -- the rows of the undecided candidates may live in a peer's HBM, one NVLink hop away
      const int wpc = nsel >= 8 ? 1 : (nsel >= 4 ? 2 : (nsel >= 2 ? 4 : 8));
      const int per_round = 8 / wpc;
      for (int c0 = 0; c0 < nsel; c0 += per_round) {
        const int c = c0 + warp / wpc, part = warp % wpc;
        if (c >= nsel) continue;
        const uint32_t id = p.cand_id[(size_t)r * CAP + s_sel[c]];
        long long g, l;
        b_locate(p, id, g, l);
        const int lo = (int)((long long)seg * part / wpc), hi = (int)((long long)seg * (part + 1) / wpc);
        long long bb = 0, ab = 0;
        int bmax = 0, badb = 0;
        if (p.b_blocks32 != nullptr) {
          // int32 copies of the peers' banks: half the bytes over NVLink
          const int* brow32 = p.b_blocks32[g] + ((size_t)l * p.d + i) * p.W + j0;
          auto acc32 = [&](int yi, int j) {
            if (yi == INT_MIN) badb = 1;
            const int xi = (int)s_a[j];
            bb += (long long)yi * yi;
            ab += (long long)xi * yi;
            const int ay = yi < 0 ? -yi : yi;
            bmax = ay > bmax ? ay : bmax;
          };
          constexpr int UN4 = 4;
          int j = lo;
          if ((((uintptr_t)(brow32 + lo)) & 15) == 0) {
            for (; j + 128 * UN4 <= hi; j += 128 * UN4) {
              int4 y[UN4];
#pragma unroll
              for (int u = 0; u < UN4; u++) y[u] = __ldg(reinterpret_cast<const int4*>(brow32 + j + u * 128) + lane);
#pragma unroll
              for (int u = 0; u < UN4; u++) {
                const int e = j + u * 128 + 4 * lane;
                acc32(y[u].x, e);
                acc32(y[u].y, e + 1);
                acc32(y[u].z, e + 2);
                acc32(y[u].w, e + 3);
              }
            }
          }
#pragma unroll 4
          for (int jj = j + lane; jj < hi; jj += 32) acc32(__ldg(brow32 + jj), jj);
        } else {
        const long long* brow = b_row(p, g, l, i) + j0;
        auto acc = [&](long long y, int j) {
          if (y >= (1LL << 31) || y <= -(1LL << 31)) badb = 1;
          const int yi = (int)y, xi = (int)s_a[j];
          bb += (long long)yi * yi;
          ab += (long long)xi * yi;
          const int ay = yi < 0 ? -yi : yi;
          bmax = ay > bmax ? ay : bmax;
        };
        // the row may sit in a peer's HBM: 8 x 16-byte loads per lane are issued before the first use
        // (4 KB in flight per warp), otherwise the NVLink round trip bounds the kernel
        constexpr int UN = 8;
        int j = lo;
        if ((((uintptr_t)(brow + lo)) & 15) == 0) {
          for (; j + 64 * UN <= hi; j += 64 * UN) {
            longlong2 y[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) y[u] = __ldg(reinterpret_cast<const longlong2*>(brow + j + u * 64) + lane);
#pragma unroll
            for (int u = 0; u < UN; u++) {
              acc(y[u].x, j + u * 64 + 2 * lane);
              acc(y[u].y, j + u * 64 + 2 * lane + 1);
            }
          }
        }
#pragma unroll 4
        for (int jj = j + lane; jj < hi; jj += 32) acc(__ldg(brow + jj), jj);
        }
        for (int o = 16; o > 0; o >>= 1) {
          bb += __shfl_xor_sync(0xffffffffu, bb, o);
          ab += __shfl_xor_sync(0xffffffffu, ab, o);
          bmax = max(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
          badb |= __shfl_xor_sync(0xffffffffu, badb, o);
        }
        if (lane == 0) {
          atomicAdd((unsigned long long*)&s_bb[c], (unsigned long long)bb);
          atomicAdd((unsigned long long*)&s_ab[c], (unsigned long long)ab);
          atomicMax(&s_bmax[c], bmax);
          if (badb) s_bad = 1;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      aa += __shfl_xor_sync(0xffffffffu, aa, o);
      amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (lane == 0) {
      s_redi[warp] = aa;
      s_redm[warp] = amax;
    }
    if (bad) s_bad = 1;
    __syncthreads();
    long long AA = 0;
    int AM = 0;
    for (int w = 0; w < 8; w++) {
      AA += s_redi[w];
      AM = max(AM, s_redm[w]);
    }
    const double sqa = sqrt((double)AA);
    for (int c = tid; c < nsel; c += blockDim.x) {
      // every partial sum of the reference's FP64 loops stays an exact integer below 2^52
      const double big = (double)max(AM, s_bmax[c]);
      if (big * big * (double)p.W >= TWO53 * 0.5) s_bad = 1;
      const double den = __dmul_rn(sqa, sqrt((double)s_bb[c]));
      if (den != 0.0) {
        const double cs = __ddiv_rn((double)s_ab[c], den);
        s_min[c] = cs < s_min[c] ? cs : s_min[c];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  // (a row that is already known to be uncertified skipped the exact sums: its candidates keep their tensor values,
  // and the k-th of THOSE sets the band pass's cut -- dropping them instead would leave fewer than k values, no cut,
  // and every positive column of the row in the band)
  if (!s_flag)
    for (int c = tid; c < nsel; c += blockDim.x) s_fin[s_sel[c]] = s_min[c];   // JAVA_MAX_DOUBLE = not comparable
  __syncthreads();
  if (warp == 0) rescore_finish(p, r, n, s_fin, s_bad ? 2 : (s_flag ? 1 : 0), lane, true);
}

// ------------------------------------------------------------------------------------------------
// K5b, narrow form.  k_rowstats makes an int32 copy of the counters and, per sketch row, max|c| and
// the exact sum of squares; k_rescore32 then needs only the cross products, reads half the bytes and
// multiplies with one IMAD.WIDE per element.  Exactness precondition per (row, candidate, depth):
// max|a| * max|b| * W < 2^52 (and the same for the squares) -- checked from the row maxima.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rowstats(const long long* __restrict__ counters, int W,
                                                  int* __restrict__ narrow, unsigned short* __restrict__ narrow16,
                                                  unsigned long long* __restrict__ row_max,
                                                  long long* __restrict__ row_ss, long long* __restrict__ row_sum,
                                                  unsigned long long* __restrict__ global_max) {
  __shared__ unsigned long long s_mx[8];
  __shared__ long long s_ss[8], s_sm[8];
  const size_t rowi = blockIdx.x;
  const long long* row = counters + rowi * W;
  int* out = narrow + rowi * W;
  unsigned short* out16 = narrow16 + rowi * W;
  unsigned long long mx = 0;
  long long ss = 0, sm = 0;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    const long long x = row[j];
    const unsigned long long ax = x < 0 ? (unsigned long long)(-x) : (unsigned long long)x;
    mx = ax > mx ? ax : mx;
    const long long c = x > 2147483647LL ? 2147483647LL : (x < -2147483647LL ? -2147483647LL : x);
    out[j] = (int)c;
    // biased 16-bit copy (value + 2^15); only read when every counter is below 2^15 in magnitude
    out16[j] = (unsigned short)((ax < 32768ull ? (int)c : 0) + 32768);
    ss += c * c;  // exact whenever the row passes the magnitude test below
    sm += c;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = om > mx ? om : mx;
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_mx[threadIdx.x >> 5] = mx;
    s_ss[threadIdx.x >> 5] = ss;
    s_sm[threadIdx.x >> 5] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      mx = s_mx[w] > mx ? s_mx[w] : mx;
      ss += s_ss[w];
      sm += s_sm[w];
    }
    row_max[rowi] = mx;
    row_ss[rowi] = ss;
    row_sum[rowi] = sm;
    atomicMax(global_max, mx);
  }
}

struct Narrow {
  const int* a;                       // [a_count][d][W]
  const int* b;                       // [blocks][b_count][d][W]
  const unsigned long long* a_max;    // [a_count][d]
  const unsigned long long* b_max;
  const long long* a_ss;
  const long long* b_ss;
  // 16-bit form: counters + 2^15 as unsigned shorts, and the row sums that undo the bias
  const unsigned short* a16;
  const unsigned short* b16;
  const long long* a_sum;
  const long long* b_sum;
};

#define RESCORE32_SEG 8192

__global__ void __launch_bounds__(256) k_rescore32(const RescoreParams p, const Narrow nw) {
  __shared__ __align__(16) int s_a[RESCORE32_SEG];
  __shared__ double s_min[CAP];
  __shared__ long long s_ab[CAP];
  __shared__ int s_bad;
  const long long r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.cand_cnt[r];
  for (int c = tid; c < CAP; c += blockDim.x) s_min[c] = JAVA_MAX_DOUBLE;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const int* arow = nw.a + (size_t)r * p.d * p.W;
  const bool vec = (p.W & 3) == 0;
  for (int i = 0; i < p.d && n > 0; i++) {
    for (int c = tid; c < CAP; c += blockDim.x) s_ab[c] = 0;
    for (int j0 = 0; j0 < p.W; j0 += RESCORE32_SEG) {
      const int seg = min(RESCORE32_SEG, p.W - j0);
      __syncthreads();
      for (int j = tid; j < seg; j += blockDim.x) s_a[j] = arow[(size_t)i * p.W + j0 + j];
      __syncthreads();
      for (int c = warp; c < n; c += 8) {
        const uint32_t id = p.cand_id[(size_t)r * CAP + c];
        long long g, l;
        b_locate(p, id, g, l);
        const int* brow = nw.b + (((size_t)g * p.b_count + l) * p.d + i) * p.W + j0;
        long long ab = 0;
        if (vec) {
          const int4* b4 = reinterpret_cast<const int4*>(brow);
          const int4* a4 = reinterpret_cast<const int4*>(s_a);
#pragma unroll 4
          for (int j = lane; j < (seg >> 2); j += 32) {
            const int4 y = __ldg(b4 + j);
            const int4 x = a4[j];
            ab += (long long)x.x * y.x;
            ab += (long long)x.y * y.y;
            ab += (long long)x.z * y.z;
            ab += (long long)x.w * y.w;
          }
        } else {
          for (int j = lane; j < seg; j += 32) ab += (long long)s_a[j] * __ldg(brow + j);
        }
        for (int o = 16; o > 0; o >>= 1) ab += __shfl_xor_sync(0xffffffffu, ab, o);
        if (lane == 0) s_ab[c] += ab;
      }
    }
    __syncthreads();
    const double amax = (double)nw.a_max[(size_t)r * p.d + i];
    const long long AA = nw.a_ss[(size_t)r * p.d + i];
    const double sqa = sqrt((double)AA);
    for (int c = tid; c < n; c += blockDim.x) {
      const uint32_t id = p.cand_id[(size_t)r * CAP + c];
      long long g, l;
      b_locate(p, id, g, l);
      const size_t bi = ((size_t)g * p.b_count + l) * p.d + i;
      const double bmax = (double)nw.b_max[bi];
      const double big = fmax(amax, bmax);
      if (big * big * (double)p.W >= TWO53 * 0.5) s_bad = 1;   // sums might round in FP64: not this path
      const double den = __dmul_rn(sqa, sqrt((double)nw.b_ss[bi]));
      if (den != 0.0) {
        const double cs = __ddiv_rn((double)s_ab[c], den);
        s_min[c] = cs < s_min[c] ? cs : s_min[c];
      }
    }
    __syncthreads();
  }
  if (warp == 0) rescore_finish(p, r, n, s_min, s_bad ? 2 : 0, lane);
}

// K5b, 16-bit form (every |counter| < 2^15, W % 8 == 0): rows are read as biased unsigned shorts, two per
// 32-bit word, so a product is one unsigned IMAD.WIDE and the re-score moves a quarter of the int64
// bytes.  sum(x y) = sum((x+B)(y+B)) - B (sum x + sum y) - W B^2 with B = 2^15: all exact integers.
__global__ void __launch_bounds__(256) k_rescore16(const RescoreParams p, const Narrow nw) {
  __shared__ __align__(16) unsigned int s_a[RESCORE32_SEG / 2];   // one depth row of A, two counters per word
  __shared__ double s_min[CAP];
  __shared__ unsigned long long s_ab[CAP];
  const long long r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.cand_cnt[r];
  for (int c = tid; c < CAP; c += blockDim.x) s_min[c] = JAVA_MAX_DOUBLE;
  __syncthreads();
  const unsigned short* arow = nw.a16 + (size_t)r * p.d * p.W;
  for (int i = 0; i < p.d && n > 0; i++) {
    for (int c = tid; c < CAP; c += blockDim.x) s_ab[c] = 0;
    for (int j0 = 0; j0 < p.W; j0 += RESCORE32_SEG) {
      const int seg = min(RESCORE32_SEG, p.W - j0);   // multiple of 8
      __syncthreads();
      const uint4* a4g = reinterpret_cast<const uint4*>(arow + (size_t)i * p.W + j0);
      for (int j = tid; j < (seg >> 3); j += blockDim.x) reinterpret_cast<uint4*>(s_a)[j] = __ldg(a4g + j);
      __syncthreads();
      for (int c = warp; c < n; c += 8) {
        const uint32_t id = p.cand_id[(size_t)r * CAP + c];
        long long g, l;
        b_locate(p, id, g, l);
        const uint4* b4 = reinterpret_cast<const uint4*>(nw.b16 + (((size_t)g * p.b_count + l) * p.d + i) * p.W + j0);
        const uint4* a4 = reinterpret_cast<const uint4*>(s_a);
        unsigned long long ab = 0;
#pragma unroll 4
        for (int j = lane; j < (seg >> 3); j += 32) {
          const uint4 y = __ldg(b4 + j);
          const uint4 x = a4[j];
          const unsigned xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int q = 0; q < 4; q++) {
            ab += (unsigned long long)(xs[q] & 0xFFFFu) * (ys[q] & 0xFFFFu);
            ab += (unsigned long long)(xs[q] >> 16) * (ys[q] >> 16);
          }
        }
        for (int o = 16; o > 0; o >>= 1) ab += __shfl_xor_sync(0xffffffffu, ab, o);
        if (lane == 0) s_ab[c] += ab;
      }
    }
    __syncthreads();
    const long long AA = nw.a_ss[(size_t)r * p.d + i];
    const long long asum = nw.a_sum[(size_t)r * p.d + i];
    const double sqa = sqrt((double)AA);
    for (int c = tid; c < n; c += blockDim.x) {
      const uint32_t id = p.cand_id[(size_t)r * CAP + c];
      long long g, l;
      b_locate(p, id, g, l);
      const size_t bi = ((size_t)g * p.b_count + l) * p.d + i;
      const long long AB = (long long)s_ab[c] - 32768LL * (asum + nw.b_sum[bi]) - (long long)p.W * (1LL << 30);
      const double den = __dmul_rn(sqa, sqrt((double)nw.b_ss[bi]));
      if (den != 0.0) {
        const double cs = __ddiv_rn((double)AB, den);
        s_min[c] = cs < s_min[c] ? cs : s_min[c];
      }
    }
    __syncthreads();
  }
  if (warp == 0) rescore_finish(p, r, n, s_min, 0, lane);
}

// ------------------------------------------------------------------------------------------------
// K5c: exact full-row path for flagged rows: every column in sequential FP64, exactly the loop of
// DoubleCountMinSketch.cosine.  One CTA per flagged row, one thread per column (strided).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_exact_rows(const RescoreParams p, const int32_t* flagged_rows,
                                                    double inv_qa, double inv_qb, double* scratch,
                                                    int64_t total_cols) {
  const long long r = flagged_rows[blockIdx.x];
  const long long* arow = p.a_counters + (size_t)r * p.d * p.W;
  const uint32_t my_id = (uint32_t)r * p.a_id_mul + p.a_id_off;
  double* out = scratch + (size_t)blockIdx.x * total_cols;
  for (long long col = threadIdx.x; col < total_cols; col += blockDim.x) {
    const long long g = col / p.b_count, l = col % p.b_count;
    const long long* brow = b_row(p, g, l, 0);
    double mn = JAVA_MAX_DOUBLE;
    for (int i = 0; i < p.d; i++) {
      double va = 0.0, vb = 0.0, vab = 0.0;
      for (int j = 0; j < p.W; j++) {
        const double xa = (double)arow[(size_t)i * p.W + j] * inv_qa;
        const double xb = (double)brow[(size_t)i * p.W + j] * inv_qb;
        va = __dadd_rn(va, __dmul_rn(xa, xa));
        vb = __dadd_rn(vb, __dmul_rn(xb, xb));
        vab = __dadd_rn(vab, __dmul_rn(xa, xb));
      }
      const double den = __dmul_rn(sqrt(va), sqrt(vb));
      if (den != 0.0) {
        const double cs = __ddiv_rn(vab, den);
        mn = cs < mn ? cs : mn;
      }
    }
    (void)my_id;
    out[col] = mn == JAVA_MAX_DOUBLE ? nan("") : mn;
  }
}

// The same full-row values at memory speed: one WARP per (flagged row, column), coalesced 16-byte loads of both
// sketch rows, the three sums as exact integers.  While max|a| * max|b| * W < 2^52 every partial sum of the
// reference's sequential FP64 loop is an exactly representable integer, so the integer sums converted to double ARE
// that loop's results, whatever the order; a (column, depth) that fails the bound is recomputed by one lane with the
// sequential FP64 loop of k_exact_rows.  A CTA takes EXACT_COLS columns of one flagged row; the A row (d x W x 8 B)
// is read through L2 by every warp.
#define EXACT_COLS 256
__global__ void __launch_bounds__(256) k_exact_rows_fast(const RescoreParams p, const int32_t* flagged_rows,
                                                         double* scratch, int64_t total_cols) {
  const long long r = flagged_rows[blockIdx.x];
  const long long* arow = p.a_counters + (size_t)r * p.d * p.W;
  double* out = scratch + (size_t)blockIdx.x * total_cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long c0 = (long long)blockIdx.y * EXACT_COLS;
  const long long c1 = (c0 + EXACT_COLS) < total_cols ? (c0 + EXACT_COLS) : total_cols;
  const bool vec = (p.W & 1) == 0;
  for (long long col = c0 + warp; col < c1; col += 8) {
    const long long g = col / p.b_count, l = col % p.b_count;
    double mn = JAVA_MAX_DOUBLE;
    for (int i = 0; i < p.d; i++) {
      const long long* a = arow + (size_t)i * p.W;
      const long long* b = b_row(p, g, l, i);
      long long aa = 0, bb = 0, ab = 0;
      unsigned long long amax = 0, bmax = 0;
      auto acc = [&](long long x, long long y) {
        const unsigned long long ax = x < 0 ? (unsigned long long)(-x) : (unsigned long long)x;
        const unsigned long long ay = y < 0 ? (unsigned long long)(-y) : (unsigned long long)y;
        amax = ax > amax ? ax : amax;
        bmax = ay > bmax ? ay : bmax;
        aa += x * x;   // wraps only when the bound below fails, and then the value is not used
        bb += y * y;
        ab += x * y;
      };
      int j = 0;
      if (vec && ((((uintptr_t)a | (uintptr_t)b) & 15) == 0)) {
        const int pairs = p.W >> 1;
        constexpr int UN = 4;
        for (; j + 32 * UN <= pairs; j += 32 * UN) {
          longlong2 xa[UN], xb[UN];
#pragma unroll
          for (int u = 0; u < UN; u++) {
            xa[u] = __ldg(reinterpret_cast<const longlong2*>(a) + j + u * 32 + lane);
            xb[u] = __ldg(reinterpret_cast<const longlong2*>(b) + j + u * 32 + lane);
          }
#pragma unroll
          for (int u = 0; u < UN; u++) {
            acc(xa[u].x, xb[u].x);
            acc(xa[u].y, xb[u].y);
          }
        }
        j *= 2;
      }
      for (int jj = j + lane; jj < p.W; jj += 32) acc(__ldg(a + jj), __ldg(b + jj));
      for (int o = 16; o > 0; o >>= 1) {
        aa += __shfl_xor_sync(0xffffffffu, aa, o);
        bb += __shfl_xor_sync(0xffffffffu, bb, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, amax, o), ob = __shfl_xor_sync(0xffffffffu, bmax, o);
        amax = oa > amax ? oa : amax;
        bmax = ob > bmax ? ob : bmax;
      }
      double va, vb, vab;
      const double big = (double)(amax > bmax ? amax : bmax);
      if (big * big * (double)p.W < TWO53 * 0.5) {
        va = (double)aa;
        vb = (double)bb;
        vab = (double)ab;
      } else {
        // beyond the exact range: the reference's own summation order decides the roundings
        va = vb = vab = 0.0;
        if (lane == 0) {
          for (int jj = 0; jj < p.W; jj++) {
            const double xa = (double)a[jj], xb = (double)b[jj];
            va = __dadd_rn(va, __dmul_rn(xa, xa));
            vb = __dadd_rn(vb, __dmul_rn(xb, xb));
            vab = __dadd_rn(vab, __dmul_rn(xa, xb));
          }
        }
      }
      const double den = __dmul_rn(sqrt(va), sqrt(vb));
      if (den != 0.0) {
        const double cs = __ddiv_rn(vab, den);
        mn = cs < mn ? cs : mn;
      }
    }
    if (lane == 0) out[col] = mn == JAVA_MAX_DOUBLE ? nan("") : mn;
  }
}

// top-k of one flagged row's exact similarities (scratch), one CTA per row: k rounds of arg-max
__global__ void __launch_bounds__(256) k_exact_topk(const RescoreParams p, const int32_t* flagged_rows,
                                                    const double* scratch, int64_t total_cols,
                                                    int exclude_self) {
  __shared__ double s_v[256];
  __shared__ uint32_t s_i[256];
  const long long r = flagged_rows[blockIdx.x];
  const uint32_t my_id = exclude_self ? (uint32_t)r * p.a_id_mul + p.a_id_off : 0xFFFFFFFFu;
  const double* v = scratch + (size_t)blockIdx.x * total_cols;
  double last_v = INFINITY;
  uint32_t last_id = 0;
  bool first = true;
  int admitted = 0;
  for (int t = 0; t < p.k; t++) {
    double best = -INFINITY;
    uint32_t best_id = 0xFFFFFFFFu;
    for (long long col = threadIdx.x; col < total_cols; col += blockDim.x) {
      const long long g = col / p.b_count, l = col % p.b_count;
      const uint32_t id = (uint32_t)l * p.b_id_mul + (uint32_t)g * p.b_id_add;
      const double s = v[col];
      if (!(s >= p.threshold) || !(s > 4.9e-324) || id == my_id) continue;  // NaN fails the first test
      // strictly after (last_v, last_id) in the order (sim desc, id asc)
      if (!first && !(s < last_v || (s == last_v && id > last_id))) continue;
      if (s > best || (s == best && id < best_id)) {
        best = s;
        best_id = id;
      }
    }
    s_v[threadIdx.x] = best;
    s_i[threadIdx.x] = best_id;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const double ov = s_v[threadIdx.x + o];
        const uint32_t oi = s_i[threadIdx.x + o];
        if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x])) {
          s_v[threadIdx.x] = ov;
          s_i[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    best = s_v[0];
    best_id = s_i[0];
    __syncthreads();
    const bool ok = best > -INFINITY;
    if (threadIdx.x == 0) {
      p.out_idx[(size_t)r * p.k + t] = ok ? (long long)best_id : -1LL;
      p.out_sim[(size_t)r * p.k + t] = ok ? best : 0.0;
    }
    if (ok) {
      admitted++;
      last_v = best;
      last_id = best_id;
      first = false;
    } else {
      // nothing left: fill the rest
      for (int u = t + 1 + threadIdx.x; u < p.k; u += blockDim.x) {
        p.out_idx[(size_t)r * p.k + u] = -1LL;
        p.out_sim[(size_t)r * p.k + u] = 0.0;
      }
      break;
    }
  }
  if (threadIdx.x == 0) p.out_cnt[r] = admitted;
}

// ------------------------------------------------------------------------------------------------
// Band pass.  A row the candidate lists could not certify gets a second, targeted sweep instead of the exact
// full-row path: its normalised row is copied into a compact A operand (k_band_gather), K3 runs over those rows
// only, with a fixed per-row cut and no selection (every column above the cut is listed), and k_band_finish
// re-scores what was listed exactly -- integer dot products, as in k_certify -- and writes the row's top-k.
// The cost is that of K3 over the flagged rows plus a few hundred exact dot products per row, instead of one
// exact dot product per column.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_band_gather(const uint16_t* __restrict__ rows16, const uint32_t* __restrict__ valid,
                                                     long long a_count, long long a_vw, int d, int ld,
                                                     const int32_t* __restrict__ flagged, int nf, long long nf_vw,
                                                     uint32_t a_id_mul, uint32_t a_id_off, const float* __restrict__ row_cut,
                                                     uint16_t* __restrict__ out_rows, uint32_t* __restrict__ out_valid,
                                                     uint32_t* __restrict__ out_ids, float* __restrict__ out_cut) {
  const int r2 = blockIdx.x;  // compact row
  const long long r = flagged[r2];
  for (int i = 0; i < d; i++) {
    const uint4* src = reinterpret_cast<const uint4*>(rows16 + ((size_t)i * a_count + r) * ld);
    uint4* dst = reinterpret_cast<uint4*>(out_rows + ((size_t)i * nf + r2) * ld);
    for (int j = threadIdx.x; j < ld / 8; j += blockDim.x) dst[j] = __ldg(src + j);
    if (threadIdx.x == 0 && ((__ldg(valid + (size_t)i * a_vw + (r >> 5)) >> (r & 31)) & 1u))
      atomicOr(out_valid + (size_t)i * nf_vw + (r2 >> 5), 1u << (r2 & 31));
  }
  if (threadIdx.x == 0) {
    out_ids[r2] = (uint32_t)r * a_id_mul + a_id_off;
    out_cut[r2] = row_cut[r];
  }
}


struct BandParams {
  const int32_t* flagged;   // compact row -> row of the job
  int cap;                  // candidates a row can hold (row stride of cand / cval; the kernel's dynamic smem is 12 * cap)
  const uint32_t* cand;     // [nf][cap] column ids above the row's cut, all sweeps
  const int32_t* cand_cnt;  // [nf]
  const float* cval;        // optional [nf][cap] scaled tensor values of the candidates
  int certified;            // with cval: decide from the tensor values first, re-score only the undecided
  const float* bound;       // optional [nf]: largest scaled tensor value NOT among the candidates (-inf: none); the
                            // row is only accepted when its exact k-th similarity clears it (else flag 2)
  int nf;
};

// exact sums of one (A row, B row, depth) pair by one warp; false when the integers leave the exact range
__device__ __forceinline__ bool pair_sums_exact(const RescoreParams& p, const long long* __restrict__ a, long long g,
                                                long long l, int i, int lane, double& va, double& vb, double& vab) {
  long long aa = 0, bb = 0, ab = 0;
  unsigned long long amax = 0, bmax = 0;
  int bad = 0;
  auto acc = [&](long long x, long long y) {
    const unsigned long long ax = x < 0 ? (unsigned long long)(-x) : (unsigned long long)x;
    const unsigned long long ay = y < 0 ? (unsigned long long)(-y) : (unsigned long long)y;
    amax = ax > amax ? ax : amax;
    bmax = ay > bmax ? ay : bmax;
    aa += x * x;
    bb += y * y;
    ab += x * y;
  };
  if (p.b_blocks32 != nullptr) {
    const int* b = p.b_blocks32[g] + ((size_t)l * p.d + i) * p.W;
    int j = 0;
    if ((p.W & 3) == 0 && (((uintptr_t)b) & 15) == 0 && (((uintptr_t)a) & 15) == 0) {
      for (; j + 128 <= p.W; j += 128) {
        const int4 y = __ldg(reinterpret_cast<const int4*>(b + j) + lane);
        const longlong2 x0 = __ldg(reinterpret_cast<const longlong2*>(a + j) + 2 * lane);
        const longlong2 x1 = __ldg(reinterpret_cast<const longlong2*>(a + j) + 2 * lane + 1);
        if (y.x == INT_MIN || y.y == INT_MIN || y.z == INT_MIN || y.w == INT_MIN) bad = 1;
        acc(x0.x, y.x);
        acc(x0.y, y.y);
        acc(x1.x, y.z);
        acc(x1.y, y.w);
      }
    }
    for (int jj = j + lane; jj < p.W; jj += 32) {
      const int y = __ldg(b + jj);
      if (y == INT_MIN) bad = 1;
      acc(__ldg(a + jj), y);
    }
  } else {
    const long long* b = b_row(p, g, l, i);
    int j = 0;
    if ((p.W & 1) == 0 && ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0) {
      const int pairs = p.W >> 1;
      for (; j + 64 <= pairs; j += 64) {
        const longlong2 x0 = __ldg(reinterpret_cast<const longlong2*>(a) + j + lane);
        const longlong2 y0 = __ldg(reinterpret_cast<const longlong2*>(b) + j + lane);
        const longlong2 x1 = __ldg(reinterpret_cast<const longlong2*>(a) + j + 32 + lane);
        const longlong2 y1 = __ldg(reinterpret_cast<const longlong2*>(b) + j + 32 + lane);
        acc(x0.x, y0.x);
        acc(x0.y, y0.y);
        acc(x1.x, y1.x);
        acc(x1.y, y1.y);
      }
      j *= 2;
    }
    for (int jj = j + lane; jj < p.W; jj += 32) acc(__ldg(a + jj), __ldg(b + jj));
  }
  for (int o = 16; o > 0; o >>= 1) {
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
    const unsigned long long oa = __shfl_xor_sync(0xffffffffu, amax, o), ob = __shfl_xor_sync(0xffffffffu, bmax, o);
    amax = oa > amax ? oa : amax;
    bmax = ob > bmax ? ob : bmax;
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  const double big = (double)(amax > bmax ? amax : bmax);
  va = (double)aa;
  vb = (double)bb;
  vab = (double)ab;
  return !bad && big * big * (double)p.W < TWO53 * 0.5;
}

// one CTA per flagged row: collect the listed columns, re-score them exactly, order, write the top-k
// block-wide bitonic sort of s_val / s_id[0, np2) under (value desc, index asc)
__device__ __forceinline__ void band_sort(double* s_val, uint32_t* s_id, int np2) {
  for (int kk = 2; kk <= np2; kk <<= 1) {
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      for (int e = threadIdx.x; e < np2; e += blockDim.x) {
        const int o = e ^ jj;
        if (o > e) {
          const bool desc = (e & kk) == 0;
          const double ve = s_val[e], vo = s_val[o];
          const uint32_t ie = s_id[e], io = s_id[o];
          const bool e_first = ve > vo || (ve == vo && ie <= io);
          if (e_first != desc) {
            s_val[e] = vo;
            s_val[o] = ve;
            s_id[e] = io;
            s_id[o] = ie;
          }
        }
      }
      __syncthreads();
    }
  }
}

// One CTA per row: order what the band sweeps listed, decide as much as the tensor values allow (CERTIFIED with
// bp.cval: a candidate is surely IN when the (k+1)-th tensor value cannot reach it, surely OUT when it cannot reach
// the k-th), re-score the undecided rest exactly -- integer dot products -- and write the row's top-k.  RESCORED
// (and the large-k path, which has no tensor values in band order) re-scores every candidate.
#define BAND_IN_SHIFT 4.0   /* cosines are in [-1, 1]: a surely-in candidate sorts ahead of every exact value */
__global__ void __launch_bounds__(256) k_band_finish(const RescoreParams p, const BandParams bp) {
  extern __shared__ double s_band[];
  double* s_val = s_band;
  uint32_t* s_id = (uint32_t*)(s_band + bp.cap);
  __shared__ int s_n, s_bad, s_adm;
  const int r2 = blockIdx.x;
  const long long r = bp.flagged[r2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_n = bp.cand_cnt[r2];
    s_bad = 0;
    s_adm = 0;
  }
  __syncthreads();
  const int n = s_n;
  if (n > bp.cap) {
    // a tie group (or a row of near-equal similarities) too large to settle here: the exact full-row path
    if (tid == 0) p.row_flag[r] = 2;
    return;
  }
  const bool classify = bp.cval != nullptr && bp.certified;
  int np2 = 32;
  while (np2 < n) np2 <<= 1;
  for (int e = tid; e < np2; e += blockDim.x) {
    s_id[e] = e < n ? bp.cand[(size_t)r2 * bp.cap + e] : 0xFFFFFFFFu;
    s_val[e] = (e < n && classify) ? (double)(bp.cval[(size_t)r2 * bp.cap + e] * p.inv_scale2) : -INFINITY;
  }
  __syncthreads();
  double v_k = -INFINITY, v_k1 = -INFINITY;  // k-th and (k+1)-th tensor value
  if (classify) {
    band_sort(s_val, s_id, np2);
    if (n >= p.k) v_k = s_val[p.k - 1];
    if (n > p.k) v_k1 = s_val[p.k];
    __syncthreads();
  }
  const double er = (double)p.eps_rel, ea = (double)p.eps_abs;
  const double k1_hi = v_k1 + fabs(v_k1) * er + ea;   // the most the (k+1)-th can really be
  const double k_lo = v_k - fabs(v_k) * er - ea;      // the least the k-th can really be
  const double thr = p.threshold > 0.0 ? p.threshold : 0.0;
  const long long* arow = p.a_counters + (size_t)r * p.d * p.W;
  for (int c = warp; c < n; c += 8) {
    if (classify) {
      const double v = s_val[c];
      const double lo = v - fabs(v) * er - ea, hi = v + fabs(v) * er + ea;
      // the admission cut (threshold, or positivity when the error is absolute) must be clear as well
      const bool clear_of_thr = lo > thr || (thr == 0.0 && !p.mixed && v > 0.0);
      if (c < p.k && (n <= p.k || k1_hi < lo) && clear_of_thr) {
        if (lane == 0) s_val[c] = v + BAND_IN_SHIFT;   // surely in: keeps its tensor value
        continue;
      }
      if (c >= p.k && k_lo > hi) {
        if (lane == 0) s_val[c] = -INFINITY;           // surely out
        continue;
      }
    }
    long long g, l;
    b_locate(p, s_id[c], g, l);
    double mn = JAVA_MAX_DOUBLE;
    for (int i = 0; i < p.d; i++) {
      double va, vb, vab;
      if (!pair_sums_exact(p, arow + (size_t)i * p.W, g, l, i, lane, va, vb, vab)) s_bad = 1;
      const double den = __dmul_rn(sqrt(va), sqrt(vb));
      if (den != 0.0) {
        const double cs = __ddiv_rn(vab, den);
        mn = cs < mn ? cs : mn;
      }
    }
    // admitted iff not NaN (no comparable row), >= threshold and > Double.MIN_VALUE
    if (lane == 0) s_val[c] = (mn != JAVA_MAX_DOUBLE && mn >= p.threshold && mn > 4.9e-324) ? mn : -INFINITY;
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) p.row_flag[r] = 2;
    return;
  }
  // (surely in, by tensor value) ahead of (re-scored, by exact value): the first k are the row's top-k
  band_sort(s_val, s_id, np2);
  if (bp.bound != nullptr) {
    const float b = bp.bound[r2];
    const double kth = (p.k - 1 < np2) ? s_val[p.k - 1] : -INFINITY;
    const double bs = (double)(b * p.inv_scale2);
    if (b > -INFINITY && !(kth > bs + fabs(bs) * er + ea)) {
      if (tid == 0) p.row_flag[r] = 2;
      return;
    }
  }
  if (classify) {
    // back to plain values, everything beyond the k-th dropped, and the k results ordered by the value returned
    __syncthreads();
    for (int e = tid; e < np2; e += blockDim.x) {
      double v = s_val[e];
      if (v > 2.0) v -= BAND_IN_SHIFT;
      s_val[e] = e < p.k ? v : -INFINITY;
    }
    __syncthreads();
    band_sort(s_val, s_id, np2);
  }
  int admitted = 0;
  for (int e = tid; e < p.k; e += blockDim.x) {
    const bool ok = e < np2 && s_val[e] > -INFINITY;
    p.out_idx[(size_t)r * p.k + e] = ok ? (long long)s_id[e] : -1LL;
    p.out_sim[(size_t)r * p.k + e] = ok ? s_val[e] : 0.0;
    admitted += ok ? 1 : 0;
  }
  if (admitted) atomicAdd(&s_adm, admitted);
  __syncthreads();
  if (tid == 0) {
    p.out_cnt[r] = s_adm;
    p.row_flag[r] = 0;
  }
}

// rows whose flag equals `want` (want == 0: any non-zero flag)
__global__ void k_collect_flagged(const int32_t* row_flag, long long n, int32_t* flagged, int32_t* count, int want) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n && (want ? row_flag[r] == want : row_flag[r] != 0)) flagged[atomicAdd(count, 1)] = (int32_t)r;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn(mb200_ctx* ctx) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
    mb200_fail(ctx, MB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver (%s)",
               e != cudaSuccess ? cudaGetErrorString(e) : "symbol not found");
    return nullptr;
  }
  fn = (EncodeTiledFn)ptr;
  return fn;
}

static int make_tmap(mb200_ctx* ctx, CUtensorMap* tm, int dtype, const void* base, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn(ctx);
  if (!fn) return MB200_ERR_CUDA;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; i++) gs[i] = strides_bytes[i];
  CUresult r = fn(tm, dtype == MB200_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mb200_fail(ctx, MB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return MB200_OK;
}

// a workspace slot borrowed from the context (see Workspace in common.cuh)
struct DevBuf {
  void* p = nullptr;
  int alloc(Workspace& ws, size_t bytes) { return ws.get(bytes, &p); }
};

template <int BN, int ACC, bool HM>
static int launch_cosine(mb200_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, CosParams& p,
                         int grid) {
  const size_t stage_bytes = (size_t)(BM + BN) * BK * 2;
  int stages = (int)((ctx->smem_optin - 2048 - 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "cosine: not enough shared memory for 2 stages");
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  MB_CUDA(ctx, cudaFuncSetAttribute(k_cosine<BN, ACC, HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_cosine<BN, ACC, HM><<<grid, COS_THREADS, smem, ctx->stream>>>(tmA, tmB, p);
  MB_CUDA(ctx, cudaGetLastError());
  return MB200_OK;
}


extern "C" {

int64_t mb200_row_ld(int32_t width) { return ((int64_t)width + BK - 1) / BK * BK; }
int64_t mb200_valid_words(int64_t rows) { return (rows + 255) / 256 * 8; }

int mb200_cosine_last_band_rows(mb200_ctx* ctx, int64_t* rows) {
  if (!ctx || !rows) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_last_band_rows: NULL argument");
  *rows = ctx->last_band_rows;
  return MB200_OK;
}

int mb200_cosine_last_fallback_rows(mb200_ctx* ctx, int64_t* rows) {
  if (!ctx || !rows) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_last_fallback_rows: NULL argument");
  *rows = ctx->last_fallback_rows;
  return MB200_OK;
}

static int normalize_locked(mb200_bank* bk, int dtype, void* rows16, uint32_t* valid) {
  mb200_ctx* ctx = bk->ctx;
  if (dtype != MB200_DTYPE_F16 && dtype != MB200_DTYPE_BF16)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_bank_normalize: dtype must be MB200_DTYPE_F16 or _BF16");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t vw = mb200_valid_words(bk->E);
  const int ld = (int)mb200_row_ld(bk->W);
  MB_CUDA(ctx, cudaMemsetAsync(valid, 0, (size_t)bk->d * vw * sizeof(uint32_t), ctx->stream));
  const long long blocks = bk->E * (long long)bk->d;
  if (blocks > 0x7FFFFFFFLL) return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "mb200_bank_normalize: too many sketch rows");
  {
    ProfScope prof(ctx, MB200_K_NORMALIZE);
    if (dtype == MB200_DTYPE_F16)
      k_normalize<__half><<<(unsigned)blocks, 256, 0, ctx->stream>>>(bk->counters, bk->E, bk->d, bk->W, ld, F16_SCALE,
                                                                     (__half*)rows16, valid, vw, bk->flags);
    else
      k_normalize<__nv_bfloat16><<<(unsigned)blocks, 256, 0, ctx->stream>>>(bk->counters, bk->E, bk->d, bk->W, ld, 1.0f,
                                                                            (__nv_bfloat16*)rows16, valid, vw, bk->flags);
  }
  ctx->launches++;
  MB_CUDA(ctx, cudaGetLastError());
  return MB200_OK;
}

int mb200_bank_sign_info(mb200_bank* bk, int32_t* mixed_sign) {
  if (!bk || !mixed_sign) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_bank_sign_info: NULL argument");
  mb200_ctx* ctx = bk->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long mixed = 0;
  MB_CUDA(ctx, cudaMemcpyAsync(&mixed, bk->flags + FLAG_MIXED, sizeof(mixed), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *mixed_sign = mixed ? 1 : 0;
  return MB200_OK;
}

int mb200_bank_normalize(mb200_bank* bk, int dtype, void* rows16, uint32_t* valid) {
  if (!bk || !rows16 || !valid) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_bank_normalize: NULL argument");
  std::lock_guard<std::mutex> g(bk->ctx->mu);
  return normalize_locked(bk, dtype, rows16, valid);
}

#include <chrono>
static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define TRACE(label)                                                              \
  do {                                                                            \
    if (trace) {                                                                  \
      double t_ = now_ms();                                                       \
      fprintf(stderr, "[mb200 trace] %-18s +%.3f ms\n", label, t_ - t_last);      \
      t_last = t_;                                                                \
    }                                                                             \
  } while (0)

// ------------------------------------------------------------------------------------------------
// A cosine job: the A side and the top-k parameters are fixed at begin; every push runs K3 over one
// piece of the B side; the candidate lists of a push are merged lazily -- into the job's carry
// state when another push follows, straight into the outputs at finish (so the one-shot call
// begin + push + finish launches exactly one merge).
// ------------------------------------------------------------------------------------------------
struct mb200_cosine_job {
  mb200_ctx* ctx = nullptr;
  mb200_cosine_args a;           // begin arguments (A side, shape, k, threshold, ...)
  bool rescored = false;   // candidates are re-scored from the counters (RESCORED and CERTIFIED)
  bool certified = false;
  int ksel = 0, BN = 256, num_m = 0, ld = 0;
  float scale2 = 1.f, eps_rel = 0.f;
  uint32_t* row_thr = nullptr;               // [num_m * BM], job-owned
  unsigned long long* best = nullptr;        // [num_m * BM][CAP] carry state (allocated on the 2nd push)
  float* best_bound = nullptr;               // [num_m * BM]
  bool have_best = false;
  bool pending = false;                      // lists of the last push not merged yet
  MergeParams mp;                            // ... described here (context workspace memory)
  int pushes = 0;
  mb200_cosine_piece last_piece;             // the piece of the last push (re-swept by the band pass of a one-push job)
  // band job (see band_setup): compact A rows with explicit ids and fixed per-row cuts
  const uint32_t* band_ids = nullptr;
  const float* band_cut = nullptr;
  uint32_t* band_cand = nullptr;
  float* band_val = nullptr;
  int32_t* band_cnt = nullptr;
  int band_cap = 0;
  const unsigned long long* row_ceil = nullptr;  // large-k passes (see CosParams)
  // band phase of THIS job: after a finish that deferred its uncertified rows, pushes feed `band` and the next
  // finish completes it
  struct BandState* band = nullptr;
  bool band_pending = false;
  size_t ws_state = 0;                       // workspace slots of row_thr, best, best_bound
  size_t ws_base = 0, ws_next = 0;           // workspace slots [ws_base, ws_next) belong to the pending push
};

// the job's own state lives in context workspace slots [ws_state, ws_state + 3) (cudaMalloc / cudaFree
// per job cost up to hundreds of ms on a process holding tens of GB -- they synchronise the device)
struct BandState {
  mb200_cosine_job bj;        // the nested job over the compact rows
  RescoreParams rp;           // counters, outputs, flags of the main job's finish
  BandParams bp;
  mb200_cosine_args fin;      // the finish arguments (outputs, counters) of the main job
  int32_t* d_rows = nullptr;  // flagged rows
  int nband = 0;
  int64_t total_b = 0;
  size_t ws_mark = 0;         // workspace slots from here on are free for the exact path
};

static void job_free(mb200_cosine_job* j) {
  if (!j) return;
  if (j->ctx->active_job == j) j->ctx->active_job = nullptr;
  delete j->band;
  delete j;
}

// a context that is destroyed with a job still open takes the job with it (runtime.cu)
void mb200_job_release(mb200_cosine_job* j) { job_free(j); }

static int job_begin_locked(mb200_ctx* ctx, const mb200_cosine_args* a, size_t ws_base, mb200_cosine_job** out) {
  if (!a || !out) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_begin: args/out is NULL");
  *out = nullptr;
  if (ctx->active_job)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_begin: another cosine job is active on this context");
  if (!a->a_rows || !a->a_valid)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: NULL pointer argument");
  if (a->a_count <= 0 || a->depth <= 0 || a->depth > MB200_MAX_DEPTH || a->width <= 0)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: bad shape (a_count=%lld d=%d w=%d)",
                      (long long)a->a_count, a->depth, a->width);
  if (a->k <= 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: k must be positive (got %d)", a->k);
  if (a->dtype != MB200_DTYPE_F16 && a->dtype != MB200_DTYPE_BF16)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: bad dtype %d", a->dtype);
  if (a->precision != MB200_PRECISION_TENSOR && a->precision != MB200_PRECISION_RESCORED &&
      a->precision != MB200_PRECISION_CERTIFIED)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: bad precision %d", a->precision);
  if (a->a_id_mul <= 0 || a->a_id_off < 0 || (a->a_count - 1) * a->a_id_mul + a->a_id_off >= 0xFFFFFFFFLL)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: index mapping out of the 32-bit range");
  if (a->block_n != 0 && a->block_n != 128 && a->block_n != 256)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: block_n must be 0, 128 or 256");
  const bool rescored = a->precision != MB200_PRECISION_TENSOR;
  // candidates kept per row: k + margin, at most CAP - 64
  // re-scoring margin: the more candidates beyond k, the more often the k-th value clears what was dropped (rows
  // that do not certify cost a band pass).  Measured at 500000 x 500000, depth 1, k = 100 (profiles/
  // r2_margin_500k.txt): 160 kept -> 10 % of the rows take the band pass, step 2233 ms; 192 kept -> none, 2077 ms
  // (the epilogue's threshold raises no longer stall the MMA warp, so a longer list costs K3 ~1 %).
  int margin = rescored ? std::max(14, a->k >= 64 ? a->k : a->k / 4) : 0;
  if (const char* ev = getenv("MB200_MARGIN"))  // tuning override
    if (rescored && atoi(ev) >= 0) margin = atoi(ev);
  int ksel = (a->k + margin + 31) / 32 * 32;
  if (ksel > CAP - 64) ksel = CAP - 64;
  if (a->k > ksel)
    return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "mb200_cosine_topk: k = %d exceeds the fused top-k capacity %d", a->k, CAP - 64);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  mb200_cosine_job* j = new mb200_cosine_job();
  j->ctx = ctx;
  j->a = *a;
  j->rescored = rescored;
  j->certified = a->precision == MB200_PRECISION_CERTIFIED;
  j->ksel = ksel;
  j->BN = a->block_n == 128 ? 128 : 256;
  j->num_m = (int)((a->a_count + BM - 1) / BM);
  j->ld = (int)mb200_row_ld(a->width);
  const float scale = a->dtype == MB200_DTYPE_F16 ? F16_SCALE : 1.0f;
  j->scale2 = scale * scale;
  // error bound of a tensor-core similarity: two operand roundings (2 * 2^-11 for F16's 11-bit significand,
  // 2 * 2^-8 for BF16 -- the budget below keeps the factor 4 of the first release) plus the FP32 accumulation:
  // products of 16-bit operands are exact in FP32, one rounding/truncation (<= 2^-23 relative to the running
  // sum of |terms|) per K = 16 MMA step, ld / 16 steps, doubled for the alignment inside a step
  j->eps_rel = (a->dtype == MB200_DTYPE_F16 ? 0x1.0p-10f : 0x1.0p-6f) + (float)(j->ld / 16 + 16) * 0x1.0p-22f;
  j->ws_state = ws_base;
  j->ws_base = j->ws_next = ws_base + 3;
  const size_t rows_pad = (size_t)j->num_m * BM;
  {
    Workspace ws(ctx);
    ws.next = j->ws_state;
    void* q = nullptr;
    if (ws.get(rows_pad * sizeof(uint32_t), &q) != MB200_OK) {
      delete j;
      return MB200_ERR_OOM;
    }
    j->row_thr = (uint32_t*)q;
  }
  cudaError_t e = cudaMemsetAsync(j->row_thr, 0, rows_pad * sizeof(uint32_t), ctx->stream);
  if (e != cudaSuccess) {
    job_free(j);
    return mb200_fail(ctx, MB200_ERR_CUDA, "mb200_cosine_begin: %s", cudaGetErrorString(e));
  }
  ctx->active_job = j;
  ctx->last_fallback_rows = 0;
  *out = j;
  return MB200_OK;
}

// merge the pending lists: into the carry state (to_carry) or, at finish, into the outputs
static int job_merge_pending(mb200_cosine_job* j, bool to_carry, const mb200_cosine_args* outs, MergeParams* used) {
  mb200_ctx* ctx = j->ctx;
  MergeParams mp = j->mp;
  mp.carry_in = j->have_best ? j->best : nullptr;
  mp.carry_bound_in = j->have_best ? j->best_bound : nullptr;
  mp.carry_out = nullptr;
  mp.carry_bound_out = nullptr;
  mp.emit = to_carry ? 0 : 1;
  if (to_carry) {
    const size_t rows_pad = (size_t)j->num_m * BM;
    if (!j->best) {
      Workspace ws(ctx);
      ws.next = j->ws_state + 1;
      void *q0 = nullptr, *q1 = nullptr;
      MB_CHECK(ws.get(rows_pad * CAP * sizeof(unsigned long long), &q0));
      MB_CHECK(ws.get(rows_pad * sizeof(float), &q1));
      j->best = (unsigned long long*)q0;
      j->best_bound = (float*)q1;
    }
    mp.carry_out = j->best;
    mp.carry_bound_out = j->best_bound;
  } else {
    mp.out_idx = (long long*)outs->out_idx;
    mp.out_sim = outs->out_sim;
    mp.out_cnt = outs->out_cnt;
  }
  k_merge<<<(unsigned)((j->a.a_count + 7) / 8), 256, 0, ctx->stream>>>(mp);
  ctx->launches++;
  MB_CUDA(ctx, cudaGetLastError());
  if (to_carry) j->have_best = true;
  j->pending = false;
  if (used) *used = mp;
  return MB200_OK;
}

static int job_push_locked(mb200_cosine_job* j, const mb200_cosine_piece* pc) {
  mb200_ctx* ctx = j->ctx;
  const mb200_cosine_args* a = &j->a;
  if (!pc || !pc->b_rows || !pc->b_valid)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: NULL pointer argument");
  if (pc->b_count <= 0 || pc->b_blocks <= 0)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: bad shape (b_count=%lld blocks=%d)",
                      (long long)pc->b_count, pc->b_blocks);
  const int64_t max_id = (pc->b_count - 1) * pc->b_id_mul + (pc->b_blocks - 1) * pc->b_id_add + pc->b_id_base;
  if (pc->b_id_mul <= 0 || pc->b_id_add < 0 || pc->b_id_base < 0 || max_id >= 0xFFFFFFFFLL)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: index mapping out of the 32-bit range");
  if (pc->ready_flags && !ctx->gather_abort)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_push: ready_flags must come from mb200_gather_pull on this context");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool trace = getenv("MB200_TRACE") != nullptr;
  double t_last = now_ms();
  if (j->pending) MB_CHECK(job_merge_pending(j, true, nullptr, nullptr));  // before the workspace is reused

  const int BN = j->BN, ld = j->ld, num_m = j->num_m;
  const int tpb = (int)((pc->b_count + BN - 1) / BN);
  const int T = tpb * pc->b_blocks;
  if (a->dense_out && a->dense_ld < (int64_t)T * BN)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: dense_ld must be >= %lld", (long long)T * BN);
  // Work items (row block m, tile range [t0, t1)) are dealt round-robin to one CTA per SM; the CTAs of
  // one wave start their sweeps together, so items with the same tile range share their B tiles in L2.
  // A wave of items with c tiles costs about c + 3 tile times (3 = candidate-list warm-up), and
  // the merge kernel's work grows with the number of lists per row.  Two shapes are costed:
  //   uniform: every row block is cut into S column chunks;
  //   tail   : floor(num_m / P) full waves of whole-row sweeps, the remaining R < P row blocks cut
  //            into S chunks (fixes the quantisation of the last wave without paying the warm-up
  //            everywhere).
  const int P = ctx->num_sms;
  const int smax = std::min(T, 32);
  // One tile step (all depths of one 128 x BN tile) takes d * kblocks * t_mma on the tensor pipe.  In
  // the same time the wave pulls, from HBM, the A blocks of its P/s distinct row blocks (an A block is
  // re-read for every B tile: 148 x 4 MB does not fit L2) and s B tiles: with s = 1 that alone is
  // ~6 TB/s -- the HBM roofline, measured as 59 % DRAM utilisation already at s = 2
  // (profiles/r1_k_cosine_v3_ncu.txt).  Cutting the rows into s >= 2 column chunks lets s CTAs share
  // every A block through L2.
  const double t_mma_tile = (double)a->depth * (ld / BK) * (2.0 * BM * BN * BK) / (1500e12 / P);   // s per tile step
  auto dram_tile = [&](long long nblocks, int s) {
    const double rows = (double)std::min<long long>(nblocks, std::max(1, P / s));
    const double bytes = (rows * BM + (double)s * BN) * ld * 2.0 * a->depth;
    return bytes / 5.5e12;
  };
  auto wave_cost = [&](long long nblocks, int s) {
    const long long waves = (nblocks * s + P - 1) / P;
    const int ct = (T + s - 1) / s;
    const double slow = std::max(1.0, dram_tile(nblocks, s) / t_mma_tile);
    return (double)waves * ((double)ct * slow + 3.0) * (1.0 + 0.01 * s);
  };
  int S_main = 1, S_tail = 1, main_blocks = 0;
  {
    double best = 1e300;
    for (int s = 1; s <= smax; s++) {
      const double cost = wave_cost(num_m, s);
      if (cost < best * 0.999) {
        best = cost;
        S_main = S_tail = s;
        main_blocks = num_m;
      }
    }
    // tail shape: the largest number of row blocks that fills whole waves with sm chunks each, the
    // remaining R blocks cut finer
    for (int sm = 1; sm <= std::min(smax, 4); sm++) {
      const int per_wave = std::max(1, P / sm);
      const int full = num_m / per_wave * per_wave, R = num_m - full;
      if (full <= 0 || R <= 0) continue;
      for (int s = sm; s <= smax; s++) {
        const double cost = wave_cost(full, sm) + wave_cost(R, s);
        if (cost < best * 0.999) {
          best = cost;
          S_main = sm;
          S_tail = s;
          main_blocks = full;
        }
      }
    }
  }
  if (const char* ev = getenv("MB200_COS_CHUNKS")) {  // tuning override: uniform chunks
    const int s = atoi(ev);
    if (s >= 1 && s <= T) {
      S_main = S_tail = s;
      main_blocks = num_m;
    }
  }
  if (const char* ev = getenv("MB200_COS_MAIN")) {    // tuning override: chunks of the full waves
    const int sm = atoi(ev);
    if (sm >= 1 && sm <= std::min(T, 8)) {
      const int per_wave = std::max(1, P / sm);
      S_main = sm;
      main_blocks = num_m / per_wave * per_wave;
      const int R = num_m - main_blocks;
      double best = 1e300;
      S_tail = sm;
      for (int s = sm; s <= smax && R > 0; s++) {
        const double cost = wave_cost(R, s);
        if (cost < best * 0.999) {
          best = cost;
          S_tail = s;
        }
      }
    }
  }
  if (trace) fprintf(stderr, "[mb200 trace] plan: num_m=%d T=%d main=%d blocks x %d chunks, tail x %d chunks\n", num_m, T,
                     main_blocks, S_main, S_tail);
  std::vector<int4> items;
  std::vector<int32_t> slot_ptr((size_t)num_m + 1, 0), slot_of;
  {
    // item order inside a group: groups of Gm row blocks x all S chunks run together, chunk-major
    std::vector<std::vector<int32_t>> slots((size_t)num_m);
    const bool strided = pc->ready_flags != nullptr;
    auto emit = [&](int mb, int me, int S) {
      const int ct = (T + S - 1) / S;
      const int Sr = (T + ct - 1) / ct;
      const int Gm = std::max(1, P / Sr);
      for (int m0 = mb; m0 < me; m0 += Gm) {
        const int m1 = std::min(me, m0 + Gm);
        for (int sI = 0; sI < Sr; sI++)
          for (int m = m0; m < m1; m++) {
            slots[m].push_back((int32_t)items.size());
            // pull-gather: chunk sI takes every Sr-th tile, so that every CTA starts in the first
            // (local) block and follows the blocks in their order of arrival
            if (strided) items.push_back(make_int4(m, sI, T, Sr));
            else items.push_back(make_int4(m, sI * ct, std::min(T, (sI + 1) * ct), 1));
          }
      }
    };
    emit(0, main_blocks, S_main);
    emit(main_blocks, num_m, S_tail);
    for (int m = 0; m < num_m; m++) {
      slot_ptr[m + 1] = slot_ptr[m] + (int32_t)slots[m].size();
      slot_of.insert(slot_of.end(), slots[m].begin(), slots[m].end());
    }
  }
  const int num_items = (int)items.size();
  Workspace ws(ctx);
  ws.next = j->ws_base;
  DevBuf d_items, d_slot, d_sptr, d_lists, d_cnt, d_bound;
  MB_CHECK(d_items.alloc(ws, items.size() * sizeof(int4)));
  MB_CHECK(d_slot.alloc(ws, slot_of.size() * sizeof(int32_t)));
  MB_CHECK(d_sptr.alloc(ws, slot_ptr.size() * sizeof(int32_t)));
  const size_t nlists = (size_t)num_items * 2 * BM;
  MB_CHECK(d_lists.alloc(ws, nlists * LCAP * sizeof(uint2)));
  MB_CHECK(d_cnt.alloc(ws, nlists * sizeof(int32_t)));
  MB_CHECK(d_bound.alloc(ws, nlists * sizeof(float)));
  j->ws_next = ws.next;
  MB_CUDA(ctx, cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
  MB_CUDA(ctx, cudaMemcpyAsync(d_sptr.p, slot_ptr.data(), slot_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  MB_CUDA(ctx, cudaMemcpyAsync(d_slot.p, slot_of.data(), slot_of.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  TRACE("alloc+items");

  // tensor maps
  alignas(64) CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)ld, (uint64_t)a->a_count, (uint64_t)a->depth};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)a->a_count * ld * 2};
    const uint32_t box[3] = {BK, BM, 1};
    MB_CHECK(make_tmap(ctx, &tmA, a->dtype, a->a_rows, 3, dims, str, box));
  }
  {
    const uint64_t dims[4] = {(uint64_t)ld, (uint64_t)pc->b_count, (uint64_t)a->depth, (uint64_t)pc->b_blocks};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)pc->b_count * ld * 2, (uint64_t)a->depth * pc->b_count * ld * 2};
    const uint32_t box[4] = {BK, (uint32_t)BN, 1, 1};
    MB_CHECK(make_tmap(ctx, &tmB, a->dtype, pc->b_rows, 4, dims, str, box));
  }

  CosParams p;
  memset(&p, 0, sizeof(p));
  p.items = (const int4*)d_items.p;
  p.num_items = num_items;
  p.total_tiles = T;
  p.tiles_per_block = tpb;
  p.depth = a->depth;
  p.kblocks = ld / BK;
  p.a_count = a->a_count;
  p.a_valid = a->a_valid;
  p.a_vw = mb200_valid_words(a->a_count);
  p.b_valid = pc->b_valid;
  p.b_vw = mb200_valid_words(pc->b_count);
  p.a_id_mul = (uint32_t)a->a_id_mul;
  p.a_id_off = (uint32_t)a->a_id_off;
  p.b_id_mul = (uint32_t)pc->b_id_mul;
  p.b_id_add = (uint32_t)pc->b_id_add;
  p.b_id_base = (uint32_t)pc->b_id_base;
  p.num_blocks = pc->b_blocks;
  p.ready = pc->ready_flags;
  p.epoch = pc->ready_epoch;
  p.rot = pc->ready_flags ? (int32_t)(((pc->first_block % pc->b_blocks) + pc->b_blocks) % pc->b_blocks) : 0;
  p.abort_flag = ctx->gather_abort;
  p.exclude_self = a->exclude_self ? 1 : 0;
  // a list scans its columns in index order iff the blocks do not interleave
  p.nonstrict = ((pc->b_blocks > 1 && pc->b_id_add < pc->b_count * pc->b_id_mul) ||
                 (pc->ready_flags && pc->b_blocks > 1)) ? 1 : 0;
  {
    const double thr = a->threshold > 0.0 ? a->threshold : 0.0;
    const double slack = j->rescored ? (1.0 - (double)j->eps_rel) : (1.0 - 1e-6);
    // mixed-sign counters: a pair whose exact similarity clears the threshold (or is merely positive) may show
    // a tensor value up to eps lower
    const double abs_slack = (j->rescored && a->mixed_sign) ? (double)j->eps_rel : 0.0;
    p.thr_init = (float)((thr * slack - abs_slack) * (double)j->scale2);
    if (thr > 0.0 || abs_slack > 0.0) p.thr_init = nextafterf(p.thr_init, -INFINITY);
  }
  p.ksel = j->ksel;
  p.a_ids = j->band_ids;
  p.row_cut = j->band_cut;
  p.band_cand = j->band_cand;
  p.band_val = j->band_val;
  p.band_cnt = j->band_cnt;
  p.band_cap = j->band_cap;
  p.row_ceil = j->row_ceil;
  p.lists = (uint2*)d_lists.p;
  p.list_cnt = (int32_t*)d_cnt.p;
  p.list_bound = (float*)d_bound.p;
  p.row_thr = j->row_thr;
  p.dense_out = a->dense_out;
  p.dense_ld = a->dense_ld;
  p.inv_scale2 = 1.0f / j->scale2;
  p.idesc = umma_idesc_f16(a->dtype == MB200_DTYPE_BF16 ? 1 : 0, BM, BN);
  // L2 eviction priorities: measured on config 3 and at 1e5 items, evict_last(A) / evict_first(B) lose
  // against the default policy (profiles/r1_cosine_tuning.md), so both operands use evict_normal
  p.policy_a = L2_EVICT_NORMAL;
  p.policy_b = L2_EVICT_NORMAL;
  if (const char* ev = getenv("MB200_COS_HINTS")) {
    const int h = atoi(ev);
    if (h == 1 || h == 3) p.policy_a = L2_EVICT_LAST;
    if (h == 1 || h == 2) p.policy_b = L2_EVICT_FIRST;
  }
  const int grid = std::min(num_items, P);
  {
    ProfScope prof(ctx, MB200_K_COSINE);
    if (a->depth == 1) {
      if (BN == 256) MB_CHECK((launch_cosine<256, 2, false>(ctx, tmA, tmB, p, grid)));
      else MB_CHECK((launch_cosine<128, 2, false>(ctx, tmA, tmB, p, grid)));
    } else {
      if (BN == 256) MB_CHECK((launch_cosine<256, 1, true>(ctx, tmA, tmB, p, grid)));
      else MB_CHECK((launch_cosine<128, 2, true>(ctx, tmA, tmB, p, grid)));
    }
  }
  ctx->launches++;
  TRACE("launch K3");

  MergeParams& mp = j->mp;
  memset(&mp, 0, sizeof(mp));
  mp.lists = p.lists;
  mp.list_cnt = p.list_cnt;
  mp.list_bound = p.list_bound;
  mp.row_thr = p.row_thr;
  mp.slot_ptr = (const int32_t*)d_sptr.p;
  mp.slot_of = (const int32_t*)d_slot.p;
  mp.a_count = a->a_count;
  mp.ksel = j->ksel;
  mp.k = a->k;
  mp.threshold = a->threshold > 0.0 ? a->threshold : 0.0;
  mp.inv_scale2 = p.inv_scale2;
  mp.rescored = j->rescored ? 1 : 0;
  j->pending = true;
  j->pushes++;
  j->last_piece = *pc;
  return MB200_OK;
}

// Band pass, three steps (see k_band_gather).  setup: compact copy of the flagged rows + the nested job; push: one
// K3 sweep of the nested job over a piece of the B side + collection of what it listed; complete: exact re-scoring,
// ordering, output.  A one-push job whose piece is still resident runs all three inside its finish; a multi-push job
// (streamed B side) that asked to defer gets them spread over a second round of pushes.
static int band_setup(mb200_cosine_job* j, const mb200_cosine_args* fin, const RescoreParams& rp, const int32_t* flagged,
                      int nband, Workspace& ws, int64_t total_b) {
  mb200_ctx* ctx = j->ctx;
  const mb200_cosine_args* a = &j->a;
  const int ld = j->ld, nf = nband;
  const int64_t nf_vw = mb200_valid_words(nf);
  const int num_m = (nf + BM - 1) / BM;
  DevBuf d_arows, d_avalid, d_ids, d_cut, d_thr, d_cand, d_ccnt, d_cvals;
  int cap = BAND_MAX;  // 8 bytes per slot: 8192 slots up to 64 Ki band rows, 2048 beyond 128 Ki
  while (cap > BAND_MIN && (size_t)num_m * BM * cap * 8 > ((size_t)4 << 30)) cap >>= 1;
  MB_CHECK(d_arows.alloc(ws, (size_t)a->depth * nf * ld * 2));
  MB_CHECK(d_avalid.alloc(ws, (size_t)a->depth * nf_vw * sizeof(uint32_t)));
  MB_CHECK(d_ids.alloc(ws, (size_t)nf * sizeof(uint32_t)));
  MB_CHECK(d_cut.alloc(ws, (size_t)num_m * BM * sizeof(float)));
  MB_CHECK(d_thr.alloc(ws, (size_t)num_m * BM * sizeof(uint32_t)));
  MB_CHECK(d_cand.alloc(ws, (size_t)num_m * BM * cap * sizeof(uint32_t)));
  MB_CHECK(d_ccnt.alloc(ws, (size_t)num_m * BM * sizeof(int32_t)));
  MB_CHECK(d_cvals.alloc(ws, (size_t)num_m * BM * cap * sizeof(float)));
  MB_CUDA(ctx, cudaMemsetAsync(d_avalid.p, 0, (size_t)a->depth * nf_vw * sizeof(uint32_t), ctx->stream));
  MB_CUDA(ctx, cudaMemsetAsync(d_thr.p, 0, (size_t)num_m * BM * sizeof(uint32_t), ctx->stream));
  MB_CUDA(ctx, cudaMemsetAsync(d_ccnt.p, 0, (size_t)num_m * BM * sizeof(int32_t), ctx->stream));
  k_band_gather<<<nf, 256, 0, ctx->stream>>>((const uint16_t*)a->a_rows, a->a_valid, a->a_count, mb200_valid_words(a->a_count),
                                             a->depth, ld, flagged, nf, nf_vw, (uint32_t)a->a_id_mul, (uint32_t)a->a_id_off,
                                             rp.row_cut, (uint16_t*)d_arows.p, (uint32_t*)d_avalid.p, (uint32_t*)d_ids.p,
                                             (float*)d_cut.p);
  ctx->launches++;
  MB_CUDA(ctx, cudaGetLastError());
  delete j->band;
  BandState* bs = j->band = new BandState();
  mb200_cosine_job& bj = bs->bj;
  bj.ctx = ctx;
  bj.a = *a;
  bj.a.a_rows = d_arows.p;
  bj.a.a_valid = (const uint32_t*)d_avalid.p;
  bj.a.a_count = nf;
  bj.a.dense_out = nullptr;
  bj.rescored = j->rescored;
  bj.certified = j->certified;
  bj.ksel = j->ksel;
  bj.BN = j->BN;
  bj.num_m = num_m;
  bj.ld = ld;
  bj.scale2 = j->scale2;
  bj.eps_rel = j->eps_rel;
  bj.row_thr = (uint32_t*)d_thr.p;
  bj.band_ids = (const uint32_t*)d_ids.p;
  bj.band_cut = (const float*)d_cut.p;
  bj.band_cand = (uint32_t*)d_cand.p;
  bj.band_val = (float*)d_cvals.p;
  bj.band_cnt = (int32_t*)d_ccnt.p;
  bj.band_cap = cap;
  bj.ws_base = bj.ws_next = ws.next;
  bs->rp = rp;
  bs->fin = *fin;
  bs->d_rows = const_cast<int32_t*>(flagged);
  bs->nband = nband;
  bs->total_b = total_b;
  memset(&bs->bp, 0, sizeof(bs->bp));
  bs->bp.flagged = flagged;
  bs->bp.cap = cap;
  bs->bp.bound = nullptr;
  bs->bp.cval = (const float*)d_cvals.p;
  bs->bp.certified = j->certified ? 1 : 0;
  bs->bp.cand = (const uint32_t*)d_cand.p;
  bs->bp.cand_cnt = (const int32_t*)d_ccnt.p;
  bs->bp.nf = nf;
  return MB200_OK;
}

static int band_push(mb200_cosine_job* j, const mb200_cosine_piece* pc) {
  mb200_ctx* ctx = j->ctx;
  BandState* bs = j->band;
  bs->bj.pending = false;  // every sweep's lists are consumed right away (k_band_collect)
  MB_CHECK(job_push_locked(&bs->bj, pc));
  (void)ctx;
  bs->ws_mark = std::max(bs->ws_mark, bs->bj.ws_next);
  return MB200_OK;
}

static int band_complete(mb200_cosine_job* j) {
  mb200_ctx* ctx = j->ctx;
  BandState* bs = j->band;
  {
    ProfScope prof(ctx, MB200_K_RESCORE);
    const int smem = bs->bp.cap * 12;
    MB_CUDA(ctx, cudaFuncSetAttribute(k_band_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_band_finish<<<bs->bp.nf, 256, smem, ctx->stream>>>(bs->rp, bs->bp);
  }
  ctx->launches++;
  MB_CUDA(ctx, cudaGetLastError());
  ctx->last_band_rows = bs->nband;
  ctx->stat_band += bs->nband;
  return MB200_OK;
}

// exact full-row path for the rows whose flag is `want` (0: any flag)
static int exact_rows_pass(mb200_cosine_job* j, RescoreParams& rp, int32_t* d_rows, int want, int64_t total_b, Workspace& ws) {
  mb200_ctx* ctx = j->ctx;
  const mb200_cosine_args* a = &j->a;
  int32_t* d_count = rp.flag_count + 1;
  MB_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int32_t), ctx->stream));
  k_collect_flagged<<<(unsigned)((a->a_count + 255) / 256), 256, 0, ctx->stream>>>(rp.row_flag, a->a_count, d_rows, d_count, want);
  ctx->launches++;
  int32_t nexact = 0;
  MB_CUDA(ctx, cudaMemcpyAsync(&nexact, d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->last_fallback_rows = nexact;
  ctx->stat_fallback += nexact;
  if (ctx->max_fallback_rows >= 0 && nexact > ctx->max_fallback_rows)
    return mb200_fail(ctx, MB200_ERR_UNSUPPORTED,
                      "%d rows cannot be certified from their candidate lists and would take the exact full-row path "
                      "(limit MB200_OPT_MAX_FALLBACK_ROWS = %lld)", nexact, (long long)ctx->max_fallback_rows);
  if (nexact > 0) {
    DevBuf d_scratch;
    const int batch = (int)std::max<int64_t>(1, std::min<int64_t>(nexact, (1LL << 30) / (total_b * 8)));
    MB_CHECK(d_scratch.alloc(ws, (size_t)batch * total_b * sizeof(double)));
    for (int off = 0; off < nexact; off += batch) {
      const int m = std::min(batch, nexact - off);
      if (getenv("MB200_EXACT_ROWS_SEQ") != nullptr) {  // the loop-for-loop form, kept for cross-checks
        k_exact_rows<<<m, 256, 0, ctx->stream>>>(rp, d_rows + off, 1.0, 1.0, (double*)d_scratch.p, total_b);
      } else {
        dim3 grid((unsigned)m, (unsigned)((total_b + EXACT_COLS - 1) / EXACT_COLS));
        k_exact_rows_fast<<<grid, 256, 0, ctx->stream>>>(rp, d_rows + off, (double*)d_scratch.p, total_b);
      }
      k_exact_topk<<<m, 256, 0, ctx->stream>>>(rp, d_rows + off, (const double*)d_scratch.p, total_b, a->exclude_self ? 1 : 0);
      ctx->launches += 2;
      MB_CUDA(ctx, cudaGetLastError());
    }
  }
  return MB200_OK;
}

// second finish of a job in its band phase: complete the band pass, then the exact path for what it left
static int job_finish_band_phase(mb200_cosine_job* j) {
  mb200_ctx* ctx = j->ctx;
  BandState* bs = j->band;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CHECK(band_complete(j));
  Workspace ws(ctx);
  ws.next = bs->ws_mark;
  MB_CHECK(exact_rows_pass(j, bs->rp, bs->d_rows, 2, bs->total_b, ws));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

// fin carries the outputs and, for MB200_PRECISION_RESCORED, the resident counters with the index
// mapping of the whole B side (b_count, b_blocks, b_id_mul, b_id_add)
static int job_finish_locked(mb200_cosine_job* j, const mb200_cosine_args* fin) {
  mb200_ctx* ctx = j->ctx;
  const mb200_cosine_args* a = &j->a;
  const bool trace = getenv("MB200_TRACE") != nullptr;
  double t_last = now_ms();
  if (!fin || !fin->out_idx || !fin->out_sim || !fin->out_cnt)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: NULL pointer argument");
  if (!j->pending) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_finish: nothing was pushed");
  const bool rescored = j->rescored;
  if (rescored && (!fin->a_counters || (!fin->b_counters && !(j->certified && fin->b_counter_blocks))))
    return mb200_fail(ctx, MB200_ERR_BAD_ARG,
                      "mb200_cosine_topk: MB200_PRECISION_RESCORED needs a_counters and b_counters (CERTIFIED: or b_counter_blocks)");
  bool contiguous = false;
  if (rescored) {
    if (fin->b_count <= 0 || fin->b_blocks <= 0)
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: bad shape (b_count=%lld blocks=%d)",
                        (long long)fin->b_count, fin->b_blocks);
    contiguous = fin->b_id_mul == 1 && (fin->b_blocks == 1 || fin->b_id_add == fin->b_count);
    const bool interleaved = fin->b_id_add == 1 && fin->b_id_mul == fin->b_blocks && fin->b_blocks > 1;
    if (!contiguous && !interleaved)
      return mb200_fail(ctx, MB200_ERR_BAD_ARG,
                        "mb200_cosine_topk: B indices must be contiguous blocks (mul=1, add=b_count) or interleaved "
                        "shards (mul=b_blocks, add=1)");
  }
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  Workspace ws(ctx);
  ws.next = j->ws_next;
  const int64_t total_b = fin->b_count * fin->b_blocks;

  // merge (+ re-score)
  DevBuf d_cand, d_ccnt, d_cbound, d_flag, d_fcount, d_cval, d_rcut;
  if (rescored) {
    MB_CHECK(d_cand.alloc(ws, (size_t)a->a_count * CAP * sizeof(uint32_t)));
    MB_CHECK(d_ccnt.alloc(ws, (size_t)a->a_count * sizeof(int32_t)));
    MB_CHECK(d_cbound.alloc(ws, (size_t)a->a_count * sizeof(float)));
    MB_CHECK(d_flag.alloc(ws, (size_t)a->a_count * sizeof(int32_t)));
    MB_CHECK(d_fcount.alloc(ws, 2 * sizeof(int32_t)));
    MB_CHECK(d_rcut.alloc(ws, (size_t)a->a_count * sizeof(float)));
    MB_CHECK(d_cval.alloc(ws, j->certified ? (size_t)a->a_count * CAP * sizeof(float) : 4));
    MB_CUDA(ctx, cudaMemsetAsync(d_fcount.p, 0, 2 * sizeof(int32_t), ctx->stream));
    j->mp.cand_val = j->certified ? (float*)d_cval.p : nullptr;
    j->mp.cand_id = (uint32_t*)d_cand.p;
    j->mp.cand_cnt = (int32_t*)d_ccnt.p;
    j->mp.cand_bound = (float*)d_cbound.p;
  }
  ctx->last_fallback_rows = 0;
  ctx->stat_rows += a->a_count;
  {
    ProfScope prof(ctx, MB200_K_RESCORE);
    MergeParams mp;
    MB_CHECK(job_merge_pending(j, false, fin, &mp));
    if (rescored) {
      RescoreParams rp;
      memset(&rp, 0, sizeof(rp));
      rp.a_counters = (const long long*)fin->a_counters;
      rp.b_counters = (const long long*)fin->b_counters;
      DevBuf d_bptr, d_bptr32;
      if (j->certified && fin->b_counter_blocks) {
        // peer-mapped banks: the few uncertain candidates are read from their owners over NVLink
        MB_CHECK(d_bptr.alloc(ws, (size_t)fin->b_blocks * sizeof(void*)));
        MB_CUDA(ctx, cudaMemcpyAsync(d_bptr.p, fin->b_counter_blocks, (size_t)fin->b_blocks * sizeof(void*),
                                     cudaMemcpyHostToDevice, ctx->stream));
        rp.b_blocks_ptr = (const long long* const*)d_bptr.p;
        rp.b_counters = nullptr;
        if (fin->b_counter_blocks32) {
          MB_CHECK(d_bptr32.alloc(ws, (size_t)fin->b_blocks * sizeof(void*)));
          MB_CUDA(ctx, cudaMemcpyAsync(d_bptr32.p, fin->b_counter_blocks32, (size_t)fin->b_blocks * sizeof(void*),
                                       cudaMemcpyHostToDevice, ctx->stream));
          rp.b_blocks32 = (const int* const*)d_bptr32.p;
        }
      }
      rp.a_count = a->a_count;
      rp.b_count = fin->b_count;
      rp.d = a->depth;
      rp.W = a->width;
      rp.blocks = fin->b_blocks;
      rp.a_id_mul = (uint32_t)a->a_id_mul;
      rp.a_id_off = (uint32_t)a->a_id_off;
      rp.b_id_mul = (uint32_t)fin->b_id_mul;
      rp.b_id_add = contiguous ? (uint32_t)fin->b_count : 1u;
      rp.cand_id = mp.cand_id;
      rp.cand_cnt = mp.cand_cnt;
      rp.cand_bound = mp.cand_bound;
      rp.inv_scale2 = mp.inv_scale2;
      rp.eps_rel = j->eps_rel;
      rp.mixed = a->mixed_sign ? 1 : 0;
      rp.eps_abs = a->mixed_sign ? j->eps_rel : j->eps_rel * 1e-3f;
      rp.k = a->k;
      rp.threshold = mp.threshold;
      rp.out_idx = (long long*)fin->out_idx;
      rp.out_sim = fin->out_sim;
      rp.out_cnt = fin->out_cnt;
      rp.row_flag = (int32_t*)d_flag.p;
      rp.flag_count = (int32_t*)d_fcount.p;
      rp.row_cut = (float*)d_rcut.p;
      {
        const double thr = a->threshold > 0.0 ? a->threshold : 0.0;
        const double abs_slack = a->mixed_sign ? (double)j->eps_rel : 0.0;
        rp.thr_floor = (float)((thr * (1.0 - (double)j->eps_rel) - abs_slack) * (double)j->scale2);
      }
      rp.cand_val = mp.cand_val;
      if (j->certified) {
        if (trace) {
          cudaStreamSynchronize(ctx->stream);
          TRACE("K3+merge");
        }
        k_certify<<<(unsigned)a->a_count, 256, 0, ctx->stream>>>(rp);
        ctx->launches++;
      } else {
      // narrow (int32) copies + per-row statistics; the generic int64 kernel is kept for banks whose
      // counters do not fit 31 bits
      const bool same = fin->a_counters == fin->b_counters && fin->b_blocks == 1 && a->a_count == fin->b_count;
      const size_t rows_b = (size_t)total_b * a->depth, rows_a = (size_t)a->a_count * a->depth;
      DevBuf d_nb, d_bmax, d_bss, d_na, d_amax, d_ass, d_gmax, d_nb16, d_bsum, d_na16, d_asum;
      MB_CHECK(d_nb.alloc(ws, rows_b * a->width * sizeof(int)));
      MB_CHECK(d_bmax.alloc(ws, rows_b * 8));
      MB_CHECK(d_bss.alloc(ws, rows_b * 8));
      MB_CHECK(d_gmax.alloc(ws, 8));
      MB_CHECK(d_nb16.alloc(ws, rows_b * a->width * sizeof(unsigned short)));
      MB_CHECK(d_bsum.alloc(ws, rows_b * 8));
      MB_CUDA(ctx, cudaMemsetAsync(d_gmax.p, 0, 8, ctx->stream));
      k_rowstats<<<(unsigned)rows_b, 256, 0, ctx->stream>>>((const long long*)fin->b_counters, a->width, (int*)d_nb.p,
                                                             (unsigned short*)d_nb16.p, (unsigned long long*)d_bmax.p,
                                                             (long long*)d_bss.p, (long long*)d_bsum.p,
                                                             (unsigned long long*)d_gmax.p);
      ctx->launches++;
      Narrow nw;
      nw.b = (const int*)d_nb.p;
      nw.b_max = (const unsigned long long*)d_bmax.p;
      nw.b_ss = (const long long*)d_bss.p;
      nw.b16 = (const unsigned short*)d_nb16.p;
      nw.b_sum = (const long long*)d_bsum.p;
      if (same) {
        nw.a = nw.b;
        nw.a_max = nw.b_max;
        nw.a_ss = nw.b_ss;
        nw.a16 = nw.b16;
        nw.a_sum = nw.b_sum;
      } else {
        MB_CHECK(d_na.alloc(ws, rows_a * a->width * sizeof(int)));
        MB_CHECK(d_amax.alloc(ws, rows_a * 8));
        MB_CHECK(d_ass.alloc(ws, rows_a * 8));
        MB_CHECK(d_na16.alloc(ws, rows_a * a->width * sizeof(unsigned short)));
        MB_CHECK(d_asum.alloc(ws, rows_a * 8));
        k_rowstats<<<(unsigned)rows_a, 256, 0, ctx->stream>>>((const long long*)fin->a_counters, a->width, (int*)d_na.p,
                                                               (unsigned short*)d_na16.p, (unsigned long long*)d_amax.p,
                                                               (long long*)d_ass.p, (long long*)d_asum.p,
                                                               (unsigned long long*)d_gmax.p);
        ctx->launches++;
        nw.a = (const int*)d_na.p;
        nw.a_max = (const unsigned long long*)d_amax.p;
        nw.a_ss = (const long long*)d_ass.p;
        nw.a16 = (const unsigned short*)d_na16.p;
        nw.a_sum = (const long long*)d_asum.p;
      }
      unsigned long long gmax = 0;
      MB_CUDA(ctx, cudaMemcpyAsync(&gmax, d_gmax.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
      MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      TRACE("K3+merge+rowstats");
      if (gmax < (1ull << 15) && (a->width & 7) == 0 && getenv("MB200_RESCORE64") == nullptr &&
          getenv("MB200_RESCORE32") == nullptr)
        k_rescore16<<<(unsigned)a->a_count, 256, 0, ctx->stream>>>(rp, nw);
      else if (gmax < (1ull << 31) && getenv("MB200_RESCORE64") == nullptr)
        k_rescore32<<<(unsigned)a->a_count, 256, 0, ctx->stream>>>(rp, nw);
      else
        k_rescore<<<(unsigned)a->a_count, 256, 0, ctx->stream>>>(rp);
      ctx->launches++;
      }
      int32_t nflag = 0;
      MB_CUDA(ctx, cudaMemcpyAsync(&nflag, d_fcount.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      TRACE("rescore");
      ctx->last_band_rows = 0;
      ctx->last_fallback_rows = 0;
      if (nflag > 0) {
        rp.b_id_add = (uint32_t)fin->b_id_add;  // forward mapping for k_exact_*
        DevBuf d_rows;
        MB_CHECK(d_rows.alloc(ws, (size_t)nflag * sizeof(int32_t)));
        int32_t* d_count = rp.flag_count + 1;
        // Rows that are merely uncertified get a band pass -- here and now when the (single) piece of the B side is
        // still resident (the caller passes it again in fin), or spread over a second round of pushes when the
        // caller asked to defer them; otherwise, and for what the band pass cannot settle, the exact full-row path.
        const bool no_band = getenv("MB200_NO_BAND") != nullptr || j->band_cut != nullptr;
        const bool band_now = !no_band && j->pushes == 1 && fin->b_rows != nullptr && fin->b_valid != nullptr;
        const bool band_later = !no_band && !band_now && fin->defer_uncertified != 0;
        int32_t nband = 0;
        if (band_now || band_later) {
          MB_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int32_t), ctx->stream));
          k_collect_flagged<<<(unsigned)((a->a_count + 255) / 256), 256, 0, ctx->stream>>>(rp.row_flag, a->a_count,
                                                                                       (int32_t*)d_rows.p, d_count, 1);
          ctx->launches++;
          MB_CUDA(ctx, cudaMemcpyAsync(&nband, d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
          MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        if (nband > 0) {
          DevBuf d_brows;  // the band rows keep their own list: d_rows is reused by the exact path
          MB_CHECK(d_brows.alloc(ws, (size_t)nband * sizeof(int32_t)));
          MB_CUDA(ctx, cudaMemcpyAsync(d_brows.p, d_rows.p, (size_t)nband * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
          MB_CHECK(band_setup(j, fin, rp, (const int32_t*)d_brows.p, nband, ws, total_b));
          j->band->d_rows = (int32_t*)d_rows.p;
          if (band_later) {
            // everything else (band sweeps, completion, the exact path for flag-2 rows) happens in the second round
            j->band->ws_mark = ws.next;
            ctx->last_band_rows = nband;
            return MB200_BAND_PENDING;
          }
          mb200_cosine_piece pc = j->last_piece;
          pc.b_rows = fin->b_rows;
          pc.b_valid = fin->b_valid;
          pc.ready_flags = nullptr;  // every block of a pull-gather has landed by now
          pc.ready_epoch = 0;
          pc.first_block = 0;
          if (trace) {
            cudaStreamSynchronize(ctx->stream);
            TRACE("band setup");
          }
          MB_CHECK(band_push(j, &pc));
          if (trace) {
            cudaStreamSynchronize(ctx->stream);
            TRACE("band sweep");
          }
          MB_CHECK(band_complete(j));
          if (trace) {
            cudaStreamSynchronize(ctx->stream);
            TRACE("band finish");
          }
          ws.next = std::max(ws.next, j->band->ws_mark);
        }
        MB_CHECK(exact_rows_pass(j, rp, (int32_t*)d_rows.p, (band_now || band_later) ? 2 : 0, total_b, ws));
      }
    }
  }
  TRACE("launch K5");
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // results are complete on return
  TRACE("sync");
  if (ctx->gather_abort) {
    uint32_t ab = 0;
    MB_CUDA(ctx, cudaMemcpy(&ab, ctx->gather_abort, 4, cudaMemcpyDeviceToHost));
    if (ab) {
      cudaMemset(ctx->gather_abort, 0, 4);
      return mb200_fail(ctx, MB200_ERR_PULL_TIMEOUT, "mb200_cosine: a peer block never arrived (pull-gather timed out after ~4 s); results are invalid");
    }
  }
  return MB200_OK;
}

// ---- k beyond the fused capacity -------------------------------------------------------------------------------
// The reference takes any --maxSimilaritiesPerItem (ItemSimilarityJob.java:105).  The fused selection holds CAP - 64
// candidates per row; beyond that the candidates are taken LARGE_KP ranks at a time: pass p is an ordinary
// tensor-precision job that returns the next LARGE_KP columns after the last one already taken (row_ceil), in the
// exact order of the tensor values.  The collected candidates are then written out (TENSOR) or re-scored exactly,
// ordered and certified against the last value taken (k_band_finish); rows that do not certify take the exact
// full-row path.  Cost: ceil((k + margin) / 128) sweeps of K3.
static constexpr int LARGE_KP = 128;

__global__ void k_largek_init(unsigned long long* ceil, int32_t* ccnt, float* bound, long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) {
    ceil[r] = ~0ull;
    ccnt[r] = 0;
    bound[r] = -INFINITY;
  }
}

// one warp per row: the pass's results join the row's candidates; the last one becomes the ceiling of the next pass
__global__ void __launch_bounds__(256) k_largek_append(const long long* __restrict__ t_idx, const double* __restrict__ t_sim,
                                                       const int32_t* __restrict__ t_cnt, long long rows, float scale2,
                                                       uint32_t* __restrict__ cand, float* __restrict__ cval,
                                                       int32_t* __restrict__ ccnt, unsigned long long* __restrict__ ceil,
                                                       float* __restrict__ bound) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int n = t_cnt[r], off = ccnt[r];
  for (int t = lane; t < n; t += 32) {
    cand[(size_t)r * LARGE_MAX + off + t] = (uint32_t)t_idx[(size_t)r * LARGE_KP + t];
    cval[(size_t)r * LARGE_MAX + off + t] = (float)(t_sim[(size_t)r * LARGE_KP + t] * (double)scale2);  // exact: power of two
  }
  __syncwarp();
  if (lane == 0) {
    ccnt[r] = off + n;
    if (n == LARGE_KP) {
      const float v = (float)(t_sim[(size_t)r * LARGE_KP + n - 1] * (double)scale2);
      ceil[r] = make_key(__float_as_uint(v), (uint32_t)t_idx[(size_t)r * LARGE_KP + n - 1]);
      bound[r] = v;
    } else {
      ceil[r] = 0ull;       // the row is exhausted: nothing else clears the admission threshold
      bound[r] = -INFINITY;
    }
  }
}

__global__ void k_largek_emit(const uint32_t* __restrict__ cand, const float* __restrict__ cval, const int32_t* __restrict__ ccnt,
                              long long rows, int k, float inv_scale2, long long* __restrict__ out_idx,
                              double* __restrict__ out_sim, int32_t* __restrict__ out_cnt) {
  const long long r = blockIdx.x;
  const int n = min(ccnt[r], k);
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    out_idx[(size_t)r * k + t] = t < n ? (long long)cand[(size_t)r * LARGE_MAX + t] : -1LL;
    out_sim[(size_t)r * k + t] = t < n ? (double)(cval[(size_t)r * LARGE_MAX + t] * inv_scale2) : 0.0;
  }
  if (threadIdx.x == 0) out_cnt[r] = n;
}

__global__ void k_iota32(int32_t* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

static int job_push_locked(mb200_cosine_job* j, const mb200_cosine_piece* pc);
static int job_finish_locked(mb200_cosine_job* j, const mb200_cosine_args* fin);

static int large_k_topk_locked(mb200_ctx* ctx, const mb200_cosine_args* a, size_t ws_base) {
  const bool rescored = a->precision != MB200_PRECISION_TENSOR;
  const int want = a->k + (rescored ? std::max(64, a->k / 4) : 0);
  const int P = (want + LARGE_KP - 1) / LARGE_KP;
  if (P * LARGE_KP > LARGE_MAX)
    return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "mb200_cosine_topk: k = %d is beyond what the multi-pass selection holds (%d)",
                      a->k, LARGE_MAX - LARGE_MAX / 5);
  if (rescored && a->b_counter_blocks)
    return mb200_fail(ctx, MB200_ERR_UNSUPPORTED, "mb200_cosine_topk: k > %d needs the gathered b_counters (not b_counter_blocks)", CAP - 64);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long rows = a->a_count;
  Workspace ws(ctx);
  ws.next = ws_base;
  DevBuf d_cand, d_cval, d_ccnt, d_ceil, d_bound, d_tidx, d_tsim, d_tcnt, d_flag, d_fcount, d_rows;
  MB_CHECK(d_cand.alloc(ws, (size_t)rows * LARGE_MAX * sizeof(uint32_t)));
  MB_CHECK(d_cval.alloc(ws, (size_t)rows * LARGE_MAX * sizeof(float)));
  MB_CHECK(d_ccnt.alloc(ws, (size_t)rows * sizeof(int32_t)));
  MB_CHECK(d_ceil.alloc(ws, (size_t)rows * sizeof(unsigned long long)));
  MB_CHECK(d_bound.alloc(ws, (size_t)rows * sizeof(float)));
  MB_CHECK(d_tidx.alloc(ws, (size_t)rows * LARGE_KP * sizeof(long long)));
  MB_CHECK(d_tsim.alloc(ws, (size_t)rows * LARGE_KP * sizeof(double)));
  MB_CHECK(d_tcnt.alloc(ws, (size_t)rows * sizeof(int32_t)));
  MB_CHECK(d_flag.alloc(ws, (size_t)rows * sizeof(int32_t)));
  MB_CHECK(d_fcount.alloc(ws, 2 * sizeof(int32_t)));
  MB_CHECK(d_rows.alloc(ws, (size_t)rows * sizeof(int32_t)));
  const unsigned g256 = (unsigned)((rows + 255) / 256);
  k_largek_init<<<g256, 256, 0, ctx->stream>>>((unsigned long long*)d_ceil.p, (int32_t*)d_ccnt.p, (float*)d_bound.p, rows);
  const float scale = a->dtype == MB200_DTYPE_F16 ? F16_SCALE : 1.0f;
  const float scale2 = scale * scale;
  float eps_rel = 0.f;
  for (int p = 0; p < P; p++) {
    mb200_cosine_args sub = *a;
    sub.k = LARGE_KP;
    sub.precision = MB200_PRECISION_TENSOR;
    // a pair whose exact similarity clears the threshold may show a tensor value a little below it: the passes admit
    // with a lowered threshold, the exact one is applied after re-scoring
    if (rescored && a->threshold > 0.0) sub.threshold = a->threshold * (a->dtype == MB200_DTYPE_F16 ? 0.995 : 0.96);
    sub.dense_out = nullptr;
    sub.out_idx = (int64_t*)d_tidx.p;
    sub.out_sim = (double*)d_tsim.p;
    sub.out_cnt = (int32_t*)d_tcnt.p;
    mb200_cosine_job* j = nullptr;
    MB_CHECK(job_begin_locked(ctx, &sub, ws.next, &j));
    j->row_ceil = p > 0 ? (const unsigned long long*)d_ceil.p : nullptr;
    eps_rel = j->eps_rel;
    mb200_cosine_piece pc;
    memset(&pc, 0, sizeof(pc));
    pc.b_rows = a->b_rows;
    pc.b_valid = a->b_valid;
    pc.b_count = a->b_count;
    pc.b_blocks = a->b_blocks;
    pc.b_id_mul = a->b_id_mul;
    pc.b_id_add = a->b_id_add;
    int rc = job_push_locked(j, &pc);
    if (rc == MB200_OK) rc = job_finish_locked(j, &sub);
    if (rc != MB200_OK) cudaStreamSynchronize(ctx->stream);
    job_free(j);
    MB_CHECK(rc);
    k_largek_append<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(
        (const long long*)d_tidx.p, (const double*)d_tsim.p, (const int32_t*)d_tcnt.p, rows, scale2, (uint32_t*)d_cand.p,
        (float*)d_cval.p, (int32_t*)d_ccnt.p, (unsigned long long*)d_ceil.p, (float*)d_bound.p);
    ctx->launches++;
    MB_CUDA(ctx, cudaGetLastError());
  }
  ctx->last_band_rows = 0;
  ctx->last_fallback_rows = 0;
  if (!rescored) {
    k_largek_emit<<<(unsigned)rows, 128, 0, ctx->stream>>>((const uint32_t*)d_cand.p, (const float*)d_cval.p, (const int32_t*)d_ccnt.p,
                                                          rows, a->k, 1.0f / scale2, (long long*)a->out_idx, a->out_sim, a->out_cnt);
    ctx->launches++;
    MB_CUDA(ctx, cudaGetLastError());
    MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
  }
  // exact re-score of everything collected, certification against the last value taken
  const bool contiguous = a->b_id_mul == 1 && (a->b_blocks == 1 || a->b_id_add == a->b_count);
  RescoreParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.a_counters = (const long long*)a->a_counters;
  rp.b_counters = (const long long*)a->b_counters;
  rp.a_count = rows;
  rp.b_count = a->b_count;
  rp.d = a->depth;
  rp.W = a->width;
  rp.blocks = a->b_blocks;
  rp.a_id_mul = (uint32_t)a->a_id_mul;
  rp.a_id_off = (uint32_t)a->a_id_off;
  rp.b_id_mul = (uint32_t)a->b_id_mul;
  rp.b_id_add = contiguous ? (uint32_t)a->b_count : 1u;
  rp.inv_scale2 = 1.0f / scale2;
  rp.eps_rel = eps_rel;
  rp.mixed = a->mixed_sign ? 1 : 0;
  rp.eps_abs = a->mixed_sign ? eps_rel : eps_rel * 1e-3f;
  rp.k = a->k;
  rp.threshold = a->threshold > 0.0 ? a->threshold : 0.0;
  rp.out_idx = (long long*)a->out_idx;
  rp.out_sim = a->out_sim;
  rp.out_cnt = a->out_cnt;
  rp.row_flag = (int32_t*)d_flag.p;
  rp.flag_count = (int32_t*)d_fcount.p;
  MB_CUDA(ctx, cudaMemsetAsync(d_flag.p, 0, (size_t)rows * sizeof(int32_t), ctx->stream));
  MB_CUDA(ctx, cudaMemsetAsync(d_fcount.p, 0, 2 * sizeof(int32_t), ctx->stream));
  k_iota32<<<g256, 256, 0, ctx->stream>>>((int32_t*)d_rows.p, rows);
  BandParams bp;
  bp.flagged = (const int32_t*)d_rows.p;
  bp.cand = (const uint32_t*)d_cand.p;
  bp.cand_cnt = (const int32_t*)d_ccnt.p;
  bp.bound = (const float*)d_bound.p;
  bp.cval = nullptr;
  bp.certified = 0;
  bp.cap = LARGE_MAX;
  bp.nf = (int)rows;
  {
    ProfScope prof(ctx, MB200_K_RESCORE);
    MB_CUDA(ctx, cudaFuncSetAttribute(k_band_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, LARGE_MAX * 12));
    k_band_finish<<<(unsigned)rows, 256, LARGE_MAX * 12, ctx->stream>>>(rp, bp);
  }
  ctx->launches += 2;
  MB_CUDA(ctx, cudaGetLastError());
  // rows that did not certify (or left the exact-integer range): every column, exactly
  mb200_cosine_job fake;
  fake.ctx = ctx;
  fake.a = *a;
  rp.b_id_add = (uint32_t)a->b_id_add;  // forward mapping for k_exact_*
  MB_CHECK(exact_rows_pass(&fake, rp, (int32_t*)d_rows.p, 2, a->b_count * a->b_blocks, ws));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MB200_OK;
}

// the one-shot form: begin + one push of the whole B side + finish
static int cosine_topk_locked(mb200_ctx* ctx, const mb200_cosine_args* a, size_t ws_base) {
  if (!a) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: args is NULL");
  if (!a->b_rows || !a->b_valid || !a->out_idx || !a->out_sim || !a->out_cnt)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: NULL pointer argument");
  if (a->b_count <= 0 || a->b_blocks <= 0)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: bad shape (a_count=%lld b_count=%lld blocks=%d d=%d w=%d)",
                      (long long)a->a_count, (long long)a->b_count, a->b_blocks, a->depth, a->width);
  const bool contiguous = a->b_id_mul == 1 && (a->b_blocks == 1 || a->b_id_add == a->b_count);
  const bool interleaved = a->b_id_add == 1 && a->b_id_mul == a->b_blocks && a->b_blocks > 1;
  if (a->b_id_mul > 0 && a->b_id_add >= 0 && !contiguous && !interleaved)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG,
                      "mb200_cosine_topk: B indices must be contiguous blocks (mul=1, add=b_count) or interleaved "
                      "shards (mul=b_blocks, add=1)");
  if (a->precision != MB200_PRECISION_TENSOR && (!a->a_counters || !a->b_counters))
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_cosine_topk: MB200_PRECISION_RESCORED / _CERTIFIED need a_counters and b_counters");
  if (a->k > CAP - 64) return large_k_topk_locked(ctx, a, ws_base);
  mb200_cosine_job* j = nullptr;
  MB_CHECK(job_begin_locked(ctx, a, ws_base, &j));
  mb200_cosine_piece pc;
  pc.b_rows = a->b_rows;
  pc.b_valid = a->b_valid;
  pc.b_count = a->b_count;
  pc.b_blocks = a->b_blocks;
  pc.b_id_mul = a->b_id_mul;
  pc.b_id_add = a->b_id_add;
  pc.b_id_base = 0;
  pc.ready_flags = nullptr;
  pc.ready_epoch = 0;
  pc.first_block = 0;
  int rc = job_push_locked(j, &pc);
  if (rc == MB200_OK) rc = job_finish_locked(j, a);
  if (rc != MB200_OK) cudaStreamSynchronize(ctx->stream);
  job_free(j);
  return rc;
}

int mb200_bank_cosine_topk(mb200_bank* bk, int32_t k, double threshold, int exclude_self, int dtype,
                           int precision, int64_t* out_idx, double* out_sim, int32_t* out_cnt, int mem) {
  if (!bk) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_bank_cosine_topk: bank is NULL");
  mb200_ctx* ctx = bk->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (k <= 0 || !out_idx || !out_sim || !out_cnt)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_bank_cosine_topk: bad arguments (k=%d)", k);
  if (mem != MB200_MEM_HOST && mem != MB200_MEM_DEVICE)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_bank_cosine_topk: mem must be MB200_MEM_HOST or MB200_MEM_DEVICE");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t ld = mb200_row_ld(bk->W), vw = mb200_valid_words(bk->E);
  DevBuf rows, valid, o_idx, o_sim, o_cnt;
  Workspace ws(ctx);
  MB_CHECK(rows.alloc(ws, (size_t)bk->d * bk->E * ld * 2));
  MB_CHECK(valid.alloc(ws, (size_t)bk->d * vw * sizeof(uint32_t)));
  MB_CHECK(normalize_locked(bk, dtype, rows.p, (uint32_t*)valid.p));
  mb200_cosine_args a;
  memset(&a, 0, sizeof(a));
  a.a_rows = a.b_rows = rows.p;
  a.a_valid = a.b_valid = (const uint32_t*)valid.p;
  a.a_count = a.b_count = bk->E;
  a.a_id_mul = 1;
  a.a_id_off = 0;
  a.b_blocks = 1;
  a.b_id_mul = 1;
  a.b_id_add = bk->E;
  a.depth = bk->d;
  a.width = bk->W;
  a.dtype = dtype;
  a.precision = precision;
  a.k = k;
  a.threshold = threshold;
  a.exclude_self = exclude_self;
  a.a_counters = a.b_counters = (const int64_t*)bk->counters;
  if (precision != MB200_PRECISION_TENSOR) {
    unsigned long long mixed = 0;
    MB_CUDA(ctx, cudaMemcpyAsync(&mixed, bk->flags + FLAG_MIXED, sizeof(mixed), cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    a.mixed_sign = mixed ? 1 : 0;
  }
  if (mem == MB200_MEM_HOST) {
    MB_CHECK(o_idx.alloc(ws, (size_t)bk->E * k * 8));
    MB_CHECK(o_sim.alloc(ws, (size_t)bk->E * k * 8));
    MB_CHECK(o_cnt.alloc(ws, (size_t)bk->E * 4));
    a.out_idx = (int64_t*)o_idx.p;
    a.out_sim = (double*)o_sim.p;
    a.out_cnt = (int32_t*)o_cnt.p;
  } else {
    a.out_idx = out_idx;
    a.out_sim = out_sim;
    a.out_cnt = out_cnt;
  }
  MB_CHECK(cosine_topk_locked(ctx, &a, ws.next));
  if (mem == MB200_MEM_HOST) {
    MB_CUDA(ctx, cudaMemcpyAsync(out_idx, o_idx.p, (size_t)bk->E * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaMemcpyAsync(out_sim, o_sim.p, (size_t)bk->E * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaMemcpyAsync(out_cnt, o_cnt.p, (size_t)bk->E * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return MB200_OK;
}

int mb200_cosine_topk(mb200_ctx* ctx, const mb200_cosine_args* args) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_cosine_topk: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  return cosine_topk_locked(ctx, args, 0);
}

int mb200_cosine_begin(mb200_ctx* ctx, const mb200_cosine_args* args, mb200_cosine_job** job) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_cosine_begin: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  return job_begin_locked(ctx, args, 0, job);
}

int mb200_cosine_push(mb200_cosine_job* job, const mb200_cosine_piece* piece) {
  if (!job) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_cosine_push: job is NULL");
  std::lock_guard<std::mutex> g(job->ctx->mu);
  if (job->band != nullptr && job->band_pending) return band_push(job, piece);
  return job_push_locked(job, piece);
}

int mb200_cosine_finish(mb200_cosine_job* job, const mb200_cosine_args* fin) {
  if (!job) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_cosine_finish: job is NULL");
  mb200_ctx* ctx = job->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  int rc;
  if (job->band != nullptr && job->band_pending) {
    rc = job_finish_band_phase(job);
  } else {
    rc = job_finish_locked(job, fin);
    if (rc == MB200_BAND_PENDING) {
      job->band_pending = true;  // the job lives on: the caller pushes the B side once more, then finishes again
      return rc;
    }
  }
  if (rc != MB200_OK) cudaStreamSynchronize(ctx->stream);
  job_free(job);
  return rc;
}

__global__ void __launch_bounds__(256) k_narrow32(const long long* __restrict__ in, long long n, int* __restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long x = in[t];
    out[t] = (x > 2147483647LL || x < -2147483647LL) ? INT_MIN : (int)x;
  }
}

int mb200_bank_narrow32(mb200_bank* bk, int32_t* out) {
  if (!bk || !out) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_bank_narrow32: NULL argument");
  mb200_ctx* ctx = bk->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long n = bk->E * (long long)bk->d * bk->W;
  const long long want = (n + 255) / 256, cap = (long long)ctx->num_sms * 32;
  k_narrow32<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(bk->counters, n, (int*)out);
  ctx->launches++;
  MB_CUDA(ctx, cudaGetLastError());
  return MB200_OK;
}

// cuStreamWriteValue32 through the runtime's driver entry point (no -lcuda)
typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWriteValue32Fn get_write_value_fn(mb200_ctx* ctx) {
  static StreamWriteValue32Fn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
    mb200_fail(ctx, MB200_ERR_CUDA, "cuStreamWriteValue32 is not available from the driver");
    return nullptr;
  }
  fn = (StreamWriteValue32Fn)ptr;
  return fn;
}

int mb200_peer_alloc(mb200_ctx* ctx, int64_t bytes, void** ptr, void* ipc_handle) {
  if (!ctx || !ptr || !ipc_handle || bytes <= 0) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_peer_alloc: bad arguments");
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  *ptr = nullptr;
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return mb200_fail(ctx, MB200_ERR_OOM, "mb200_peer_alloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle, *ptr);
  if (e != cudaSuccess) {
    cudaFree(*ptr);
    *ptr = nullptr;
    return mb200_fail(ctx, MB200_ERR_CUDA, "mb200_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  return MB200_OK;
}

int mb200_peer_open(mb200_ctx* ctx, const void* ipc_handle, void** ptr) {
  if (!ctx || !ptr || !ipc_handle) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_peer_open: bad arguments");
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return mb200_fail(ctx, MB200_ERR_CUDA, "mb200_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return MB200_OK;
}

int mb200_peer_close(mb200_ctx* ctx, void* ptr) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_peer_close: ctx is NULL");
  if (!ptr) return MB200_OK;
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return MB200_OK;
}

int mb200_peer_free(mb200_ctx* ctx, void* ptr) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_peer_free: ctx is NULL");
  if (!ptr) return MB200_OK;
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MB_CUDA(ctx, cudaFree(ptr));
  return MB200_OK;
}

int mb200_gather_pull(mb200_ctx* ctx, void* staging_rows, uint32_t* staging_valid, const void* const* peer_rows,
                      const uint32_t* const* peer_valid, int32_t blocks, int32_t my_block, int64_t rows_bytes,
                      int64_t valid_bytes, const uint32_t** ready_flags, uint32_t* epoch) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_gather_pull: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!staging_rows || !staging_valid || !peer_rows || !peer_valid || !ready_flags || !epoch || blocks <= 0 ||
      blocks > MB200_MAX_BLOCKS || my_block < 0 || my_block >= blocks || rows_bytes <= 0 || valid_bytes <= 0)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_gather_pull: bad arguments (blocks=%d my_block=%d)", blocks, my_block);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  StreamWriteValue32Fn wv = get_write_value_fn(ctx);
  if (!wv) return MB200_ERR_CUDA;
  if (!ctx->gather_flags) {
    MB_CUDA(ctx, cudaMalloc(&ctx->gather_flags, (MB200_MAX_BLOCKS + 1) * sizeof(uint32_t)));
    MB_CUDA(ctx, cudaMemset(ctx->gather_flags, 0, (MB200_MAX_BLOCKS + 1) * sizeof(uint32_t)));
    ctx->gather_abort = ctx->gather_flags + MB200_MAX_BLOCKS;
    MB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->gather_ev, cudaEventDisableTiming));
  }
  const uint32_t ep = ++ctx->gather_epoch;
  // a time-out of an earlier launch whose job was aborted (never finished) must not poison this one
  MB_CUDA(ctx, cudaMemsetAsync(ctx->gather_abort, 0, sizeof(uint32_t), ctx->stream));
  // the pulls start once everything queued so far on the compute stream is done (the caller's
  // cross-rank barrier sits there), and run on the copy engines beside K3
  MB_CUDA(ctx, cudaEventRecord(ctx->gather_ev, ctx->stream));
  MB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->gather_ev, 0));
  for (int s = 0; s < blocks; s++) {
    const int b = (my_block + s) % blocks;
    if (!peer_rows[b] || !peer_valid[b]) return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_gather_pull: peer pointer %d is NULL", b);
    // The own block is copied on the COMPUTE stream, ahead of K3 (a local copy at HBM speed: 0.3 ms per GB): a
    // device-to-device copy inside one GPU may be carried out by SMs, and the persistent K3 CTAs leave none -- K3
    // starts in the own block, so a copy that waited for K3 to end would time every CTA out.  The peers' blocks go
    // through the copy engines over NVLink beside K3.
    cudaStream_t via = s == 0 ? ctx->stream : ctx->copy_stream;
    MB_CUDA(ctx, cudaMemcpyAsync((char*)staging_rows + (size_t)b * rows_bytes, peer_rows[b], (size_t)rows_bytes,
                                 cudaMemcpyDeviceToDevice, via));
    MB_CUDA(ctx, cudaMemcpyAsync((char*)staging_valid + (size_t)b * valid_bytes, peer_valid[b], (size_t)valid_bytes,
                                 cudaMemcpyDeviceToDevice, via));
    CUresult r = wv((CUstream)via, (CUdeviceptr)(uintptr_t)(ctx->gather_flags + b), ep, 0);
    if (r != CUDA_SUCCESS) return mb200_fail(ctx, MB200_ERR_CUDA, "cuStreamWriteValue32 failed with CUresult %d", (int)r);
  }
  *ready_flags = ctx->gather_flags;
  *epoch = ep;
  return MB200_OK;
}

int mb200_gather_pull_counters(mb200_ctx* ctx, void* const* dst_blocks, const void* const* src_blocks, int32_t blocks,
                               int32_t my_block, int64_t bytes_per_block) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_gather_pull_counters: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!dst_blocks || !src_blocks || blocks <= 0 || blocks > MB200_MAX_BLOCKS || my_block < 0 || my_block >= blocks ||
      bytes_per_block <= 0)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_gather_pull_counters: bad arguments (blocks=%d my_block=%d)", blocks, my_block);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->gather_flags)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_gather_pull_counters: call mb200_gather_pull first (the copies are ordered behind its row pulls)");
  // (the copy stream already waits for the caller's cross-rank barrier: mb200_gather_pull put that wait there)
  for (int s = 1; s < blocks; s++) {
    const int b = (my_block + s) % blocks;
    if (!dst_blocks[b] || !src_blocks[b])
      return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_gather_pull_counters: block pointer %d is NULL", b);
    MB_CUDA(ctx, cudaMemcpyAsync(dst_blocks[b], src_blocks[b], (size_t)bytes_per_block, cudaMemcpyDeviceToDevice, ctx->copy_stream));
  }
  return MB200_OK;
}

int mb200_gather_fence(mb200_ctx* ctx) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_gather_fence: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->fence_ev) MB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fence_ev, cudaEventDisableTiming));
  MB_CUDA(ctx, cudaEventRecord(ctx->fence_ev, ctx->copy_stream));
  MB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->fence_ev, 0));
  return MB200_OK;
}

int mb200_gather_wait(mb200_ctx* ctx) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_gather_wait: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return MB200_OK;
}

int mb200_cosine_abort(mb200_cosine_job* job) {
  if (!job) return MB200_OK;
  mb200_ctx* ctx = job->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  job_free(job);
  return MB200_OK;
}

}  // extern "C"
