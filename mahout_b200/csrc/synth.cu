// synth.cu -- device-side synthetic event stream (bench / test support; see synth.h).
#include "common.cuh"
#include "synth.h"

__device__ __forceinline__ uint64_t splitmix_fin(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

__global__ void __launch_bounds__(256) k_synth(uint64_t base, long long first, long long n,
                                               unsigned long long users,
                                               const double* __restrict__ cdf, long long items,
                                               const long long* __restrict__ perm,
                                               long long* __restrict__ out_user,
                                               long long* __restrict__ out_item,
                                               float* __restrict__ out_pref) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const uint64_t t4 = (uint64_t)(first + i) * 4ull;
    const uint64_t r0 = splitmix_fin(base + t4), r1 = splitmix_fin(base + t4 + 1),
                   r2 = splitmix_fin(base + t4 + 2);
    const double u = (double)(r1 >> 11) * 0x1.0p-53;
    // lower bound: first index with cdf[idx] >= u
    long long lo = 0, hi = items;
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (__ldg(cdf + mid) < u) lo = mid + 1;
      else hi = mid;
    }
    if (lo >= items) lo = items - 1;
    if (out_user) out_user[i] = 1 + (long long)(r0 % users);
    out_item[i] = perm ? __ldg(perm + lo) : lo + 1;
    if (out_pref) out_pref[i] = 0.5f * (float)(1 + (int)(r2 % 10ull));
  }
}

extern "C" int mb200_synth_events(mb200_ctx* ctx, uint64_t seed, int64_t first, int64_t n,
                                  int64_t users, const double* cdf, int64_t items,
                                  const int64_t* perm, int64_t* out_user, int64_t* out_item,
                                  float* out_pref) {
  if (!ctx) return mb200_fail(nullptr, MB200_ERR_BAD_ARG, "mb200_synth_events: ctx is NULL");
  std::lock_guard<std::mutex> g(ctx->mu);
  if (n < 0 || users <= 0 || items <= 0 || !cdf || (n > 0 && !out_item))
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_synth_events: bad arguments");
  if (n == 0) return MB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  long long want = ceil_div64(n, 256);
  long long cap = (long long)ctx->num_sms * 16;
  k_synth<<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(
      seed * 0x9E3779B97F4A7C15ull, first, n, (unsigned long long)users, cdf, items,
      (const long long*)perm, (long long*)out_user, (long long*)out_item, out_pref);
  MB_CUDA(ctx, cudaGetLastError());
  return MB200_OK;
}

// ---- L2 atomic peak (bench support): RED.ADD.64 to uniformly random cells of an L2-resident array -------------
// What K1's miss path and bank-mode sparse path are bounded by once the counters sit in L2.  Every thread issues
// `per_thread` fire-and-forget 64-bit reductions at addresses from a counter-based generator (no loads).
__global__ void __launch_bounds__(512) k_red64_peak(unsigned long long* __restrict__ cells, unsigned long long mask,
                                                    int per_thread, unsigned long long seed) {
  unsigned long long x = seed + ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
#pragma unroll 4
  for (int i = 0; i < per_thread; i++) {
    x += 0x9E3779B97F4A7C15ull;
    const unsigned long long z = splitmix_fin(x);
    atomicAdd(cells + (z & mask), 1ull);
  }
}

extern "C" int mb200_bench_red64(mb200_ctx* ctx, int64_t cells_log2, int64_t updates, double* ms_out,
                                 double* updates_done) {
  if (!ctx || !ms_out || cells_log2 < 4 || cells_log2 > 32 || updates <= 0)
    return mb200_fail(ctx, MB200_ERR_BAD_ARG, "mb200_bench_red64: bad arguments");
  std::lock_guard<std::mutex> g(ctx->mu);
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned long long* cells = nullptr;
  const size_t n = (size_t)1 << cells_log2;
  MB_CUDA(ctx, cudaMalloc(&cells, n * 8));
  MB_CUDA(ctx, cudaMemsetAsync(cells, 0, n * 8, ctx->stream));
  const int grid = ctx->num_sms * 4, threads = 512;
  const int per_thread = (int)ceil_div64(updates, (int64_t)grid * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_red64_peak<<<grid, threads, 0, ctx->stream>>>(cells, n - 1, per_thread, 1);  // warm-up
  cudaEventRecord(e0, ctx->stream);
  k_red64_peak<<<grid, threads, 0, ctx->stream>>>(cells, n - 1, per_thread, 2);
  cudaEventRecord(e1, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(cells);
  ctx->launches += 2;
  if (e != cudaSuccess) return mb200_fail(ctx, MB200_ERR_CUDA, "mb200_bench_red64: %s", cudaGetErrorString(e));
  *ms_out = ms;
  if (updates_done) *updates_done = (double)grid * threads * per_thread;
  return MB200_OK;
}
