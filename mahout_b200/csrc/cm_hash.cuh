// cm_hash.cuh -- the count-min hash family on 64-bit words.
//
// Reference: HashFunction.hash (HashFunction.java:31-34)
//     a.multiply(k).add(b).mod(bigPrime).mod(w).intValue(),  bigPrime = 2^63 - 25
// BigInteger arithmetic there; here a 64x64->128 product folded with 2^63 == 25 (mod p).
// a, b and k are first reduced to canonical residues in [0, p) -- BigInteger.mod is always
// non-negative, so ((a*k + b) mod p) == ((a mod p)*(k mod p) + (b mod p)) mod p, which also
// covers negative keys and the Math.abs(Long.MIN_VALUE) < 0 corner of the parameters.
//
// Compiles for the device (nvcc) and, for the CPU-side unit test of the folding arithmetic
// only (tests/host/hash_host_check.cpp), for the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CMH_FN __host__ __device__ __forceinline__
#else
#define CMH_FN static inline
#endif

#define CMH_P 0x7FFFFFFFFFFFFFE7ull  /* 2^63 - 25 = 9223372036854775783 (HashFunctionBuilder.java:24) */
#define CMH_M63 0x7FFFFFFFFFFFFFFFull

CMH_FN uint64_t cmh_mulhi(uint64_t x, uint64_t y) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(x, y);
#else
  return (uint64_t)(((unsigned __int128)x * (unsigned __int128)y) >> 64);
#endif
}

// signed 64-bit value -> residue in [0, p)
CMH_FN uint64_t cmh_residue(int64_t v) {
  if (v >= 0) {
    uint64_t u = (uint64_t)v;
    return u >= CMH_P ? u - CMH_P : u;
  }
  int64_t t = v + (int64_t)CMH_P;  // v < 0 < p: no overflow
  if (t < 0) t += (int64_t)CMH_P;  // only for v in [-2^63, -p)
  return (uint64_t)t;
}

// (a*k + b) mod p for a, k, b in [0, p)
CMH_FN uint64_t cmh_mul_add_mod(uint64_t a, uint64_t k, uint64_t b) {
  uint64_t lo = a * k, hi = cmh_mulhi(a, k);        // a*k < 2^126
  uint64_t L = lo & CMH_M63;
  uint64_t H = (hi << 1) | (lo >> 63);              // a*k = H*2^63 + L, H < 2^63
  uint64_t lo2 = H * 25ull, hi2 = cmh_mulhi(H, 25ull);  // 25*H < 2^68
  uint64_t L2 = lo2 & CMH_M63;
  uint64_t H2 = (hi2 << 1) | (lo2 >> 63);           // 25*H = H2*2^63 + L2, H2 < 32
  uint64_t s = L + L2;                              // < 2^64
  if (s >= CMH_P) s -= CMH_P;
  if (s >= CMH_P) s -= CMH_P;
  s += 25ull * H2;                                  // < p + 800
  if (s >= CMH_P) s -= CMH_P;
  s += b;                                           // < 2p < 2^64
  if (s >= CMH_P) s -= CMH_P;
  return s;
}

// (a*k + b) mod p for a, b in [0, p) and a key residue below 2^32 -- item and user IDs in practice.
// a*k < 2^95 is assembled from two 32x32->64 products and split at bit 63: a*k = xh * 2^63 + xl with
// xh < 2^32, so a*k == xl + 25 * xh (mod p), a value below 2^63 + 2^37.  "r >= p" is "bit 63 of r + 25", and
// r - p is then (r + 25) with that bit cleared: two such steps (before and after adding b) give the canonical
// residue in about half the instructions of the general path.
CMH_FN uint64_t cmh_mul_add_mod_small(uint64_t a, uint32_t k, uint64_t b) {
  const uint64_t p0 = (a & 0xFFFFFFFFull) * (uint64_t)k;
  const uint64_t p1 = (a >> 32) * (uint64_t)k;          // a < 2^63: p1 < 2^63
  const uint64_t lo = p0 + (p1 << 32);                  // low 64 bits of a*k
  const uint32_t hi = (uint32_t)(p1 >> 32) + (lo < p0 ? 1u : 0u);   // bits 64..94
  const uint32_t xh = (hi << 1) | (uint32_t)(lo >> 63); // (a*k) >> 63
  uint64_t r = (lo & CMH_M63) + (uint64_t)xh * 25ull;   // < 2^63 + 2^37
  uint64_t t = r + 25ull;
  if (t >> 63) r = t & CMH_M63;                         // r < p
  r += b;                                               // < 2p = 2^64 - 50
  t = r + 25ull;
  if (t >> 63) r = t & CMH_M63;
  return r;
}

// column of key residue `kr` in a row of width w; wmask = w-1 if w is a power of two else 0
CMH_FN uint32_t cmh_column(uint64_t a_res, uint64_t b_res, uint64_t kr, uint32_t w, uint32_t wmask) {
  uint64_t s = (kr >> 32) == 0 ? cmh_mul_add_mod_small(a_res, (uint32_t)kr, b_res)
                               : cmh_mul_add_mod(a_res, kr, b_res);
  return wmask ? (uint32_t)(s & (uint64_t)wmask) : (uint32_t)(s % (uint64_t)w);
}
