/*
 * synth.h -- synthetic (user, item, pref) event generator used by bench.py and the tests.
 * NOT part of the reference-facing boundary (include/mahout_b200.h): it only exists so that the
 * benchmark can fill HBM with 10^9 Zipf-distributed events without a host round trip, and so
 * that the CPU oracle can regenerate any slice of the same stream (tests/synth_ref.py restates
 * the arithmetic below in numpy, bit for bit).
 *
 * Event t (global index, 0-based) of stream `seed`:
 *   r_s    = splitmix64_finalize(seed * 0x9E3779B97F4A7C15 + 4*t + s)          s = 0, 1, 2
 *   user   = 1 + r_0 mod users
 *   rank   = first i with cdf[i] >= (r_1 >> 11) * 2^-53    (clamped to items-1)
 *   item   = perm ? perm[rank] : rank + 1
 *   pref   = 0.5 * (1 + r_2 mod 10)                          in {0.5, 1.0, ..., 5.0}
 * `cdf` is the caller's float64 cumulative distribution over item ranks (device memory).
 */
#ifndef MB200_SYNTH_H
#define MB200_SYNTH_H
#include <stdint.h>

#include "../../include/mahout_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* all pointers are DEVICE pointers; out_user / out_pref may be NULL */
int mb200_synth_events(mb200_ctx* ctx, uint64_t seed, int64_t first, int64_t n, int64_t users,
                       const double* cdf, int64_t items, const int64_t* perm, int64_t* out_user,
                       int64_t* out_item, float* out_pref);

/* L2 atomic peak: `updates` RED.ADD.64 to uniformly random cells of a 2^cells_log2-word array (22 = 32 MiB,
 * L2-resident on B200); one warm-up launch, one timed.  The denominator bench.py quotes K1's reductions/s against. */
int mb200_bench_red64(mb200_ctx* ctx, int64_t cells_log2, int64_t updates, double* ms_out, double* updates_done);

#ifdef __cplusplus
}
#endif
#endif
