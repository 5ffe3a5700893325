// Counter-based Philox4x32-10 (Salmon et al., SC'11), written out by hand.
// Every random number of the library is a pure function of (seed, stream, counter) so a
// background matrix is identical on every GPU and independent of how the work is sharded
// (replaces R's global Mersenne-Twister state behind Rcpp::RNGScope,
// src/RcppExports.cpp:88,100 and unif_rand() in src/utils.cpp:72).
#pragma once
#include <stdint.h>

enum { CV_STREAM_BACKGROUND = 0, CV_STREAM_PROBSAMPLE = 1, CV_STREAM_BOOTSTRAP = 2, CV_STREAM_KNN = 3 };

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi,
                                                        uint32_t& lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(M0, c0, hi0, lo0);
    philox_mulhilo(M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

// 53-bit uniform in [0, 1) from two 32-bit words
__host__ __device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
