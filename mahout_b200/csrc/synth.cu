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
