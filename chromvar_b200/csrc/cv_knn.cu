// get_background_peaks_alternative on the GPU (R/test_sets.R:291-330): background peaks drawn
// uniformly among the `window` nearest neighbours of each peak in the whitened
// (fragments per peak, GC bias) plane -- the sampler behind makePermutedSets (R/test_sets.R:246-289).
//
//   1. whitening as in getBackgroundPeaks but on the RAW per-peak totals (:298-301): cv_whiten(log = 0)
//   2. peaks sorted by the first whitened coordinate (radix sort of the fp64 keys)
//   3. exact k nearest neighbours, one warp per query: the warp walks outwards from the query's
//      position in the sorted order, 32 candidates per step and direction, keeps the best
//      `window` OTHER peaks in shared memory, and stops a direction once (x - xq)^2 alone exceeds
//      the current window-th squared distance.  Squared distances are dx*dx + dy*dy with every
//      operation rounded separately (no FMA) so a numpy restatement reproduces the ORDER exactly;
//      ties are broken by the peak index (nabor::knn leaves them unspecified).  The query itself
//      is never a candidate (the reference asks for window + 1 neighbours and drops the query, :308-312).
//   4. draws: Philox4x32-10(counter = (peak, iteration, CV_STREAM_KNN, 0), key = seed); u picks
//      rank floor(u * window) in the (distance, index)-sorted neighbour list.
#include <math.h>

#include <cub/device/device_radix_sort.cuh>
#include <vector>

#include "cv_internal.cuh"
#include "cv_philox.cuh"

#define KNN_WARPS 4

__global__ void k_knn_iota(int32_t* __restrict__ v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void k_knn_gather_y(const double* __restrict__ T, const int32_t* __restrict__ perm, int P,
                               double* __restrict__ ys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) ys[i] = T[(int64_t)P + perm[i]];
}

__device__ __forceinline__ bool knn_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

// Small windows (W <= 16; makePermutedSets uses 10): the list is kept exact after every insertion (the new
// worst of W entries is found each time), so the bound is always the tightest possible -- with a band of
// candidates only a few rounds wide that matters more than the cost per insertion.
// xs, ys: whitened coordinates in x-sorted order; perm: sorted position -> peak (0-based)
// out_bg: P x niter column-major, 1-based, row = peak;  out_nn: [P][window] 1-based (row = peak, sorted
// by (distance, index); -1 beyond the number of other peaks) or NULL
__global__ void __launch_bounds__(KNN_WARPS * 32)
k_knn_sample_small(const double* __restrict__ xs, const double* __restrict__ ys, const int32_t* __restrict__ perm,
             int P, int W, int Wcap, int niter, uint32_t seed_lo, uint32_t seed_hi,
             int32_t* __restrict__ out_bg, int32_t* __restrict__ out_nn) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* bd = reinterpret_cast<double*>(smem) + (size_t)warp * Wcap;
  int32_t* bi = reinterpret_cast<int32_t*>(smem + (size_t)KNN_WARPS * Wcap * 8) + (size_t)warp * Wcap;
  int32_t* bs = reinterpret_cast<int32_t*>(smem + (size_t)KNN_WARPS * Wcap * 12) + (size_t)warp * Wcap;
  const int q = blockIdx.x * KNN_WARPS + warp;
  if (q >= P) return;                                  // whole warp
  const double xq = xs[q], yq = ys[q];
  const int qi = perm[q];
  int m = 0;                                           // entries in the list
  double thr_d = INFINITY;                             // worst entry once m == W
  int thr_i = 0x7fffffff, thr_pos = 0;
  int l = q - 1, r = q + 1;
  bool ldone = l < 0, rdone = r >= P;

  auto refresh_worst = [&]() {
    double wd = -1.0;
    int wi = -1, wp = 0;
    for (int e = lane; e < m; e += 32) {
      const double d = bd[e];
      const int i = bi[e];
      if (knn_less(wd, wi, d, i)) { wd = d; wi = i; wp = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, wd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o), op = __shfl_xor_sync(0xffffffffu, wp, o);
      if (knn_less(wd, wi, od, oi)) { wd = od; wi = oi; wp = op; }
    }
    thr_d = wd; thr_i = wi; thr_pos = wp;
  };

  while (!ldone || !rdone) {
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      if (dir == 0 ? rdone : ldone) continue;
      const int pos = dir == 0 ? r + lane : l - lane;
      const bool valid = pos >= 0 && pos < P;
      double d2 = INFINITY, dx2 = INFINITY;
      int idx = 0x7fffffff;
      if (valid) {
        const double dx = __dsub_rn(xs[pos], xq), dy = __dsub_rn(ys[pos], yq);
        dx2 = __dmul_rn(dx, dx);
        d2 = __dadd_rn(dx2, __dmul_rn(dy, dy));
        idx = perm[pos];
      }
      // sorted by x: once the nearest candidate of this direction is farther in x alone than the
      // worst neighbour kept, nothing beyond it can enter the list
      const double dx2_first = __shfl_sync(0xffffffffu, dx2, 0);
      if (m == W && dx2_first > thr_d) {
        if (dir == 0) rdone = true; else ldone = true;
        continue;
      }
      unsigned mask = __ballot_sync(0xffffffffu, valid && (m < W || knn_less(d2, idx, thr_d, thr_i)));
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const double cd = __shfl_sync(0xffffffffu, d2, src);
        const int ci = __shfl_sync(0xffffffffu, idx, src);
        if (m < W) {
          if (lane == 0) { bd[m] = cd; bi[m] = ci; }
          ++m;
          __syncwarp();
          if (m == W) refresh_worst();
        } else if (knn_less(cd, ci, thr_d, thr_i)) {
          if (lane == 0) { bd[thr_pos] = cd; bi[thr_pos] = ci; }
          __syncwarp();
          refresh_worst();
        }
      }
      if (dir == 0) { r += 32; rdone = r >= P; } else { l -= 32; ldone = l < 0; }
    }
  }
  __syncwarp();
  // rank of every entry in (distance, index) order -> bs[rank] = peak
  for (int e = lane; e < m; e += 32) {
    const double d = bd[e];
    const int i = bi[e];
    int rank = 0;
    for (int f = 0; f < m; ++f) rank += knn_less(bd[f], bi[f], d, i) ? 1 : 0;
    bs[rank] = i;
  }
  __syncwarp();
  if (out_nn)
    for (int e = lane; e < W; e += 32) out_nn[(int64_t)qi * W + e] = e < m ? bs[e] + 1 : -1;
  for (int it = lane; it < niter; it += 32) {
    const Philox4 rr = philox4x32_10((uint32_t)qi, (uint32_t)it, CV_STREAM_KNN, 0u, seed_lo, seed_hi);
    int j = (int)(u01_53(rr.v[0], rr.v[1]) * (double)m);
    if (j >= m) j = m - 1;
    out_bg[(int64_t)it * P + qi] = bs[j] + 1;
  }
}

// (distance, index)-ascending bitonic sort of cap = 2^n entries by one warp (shared memory)
__device__ __forceinline__ void knn_sort(double* bd, int32_t* bi, int cap, int lane) {
  for (int k = 2; k <= cap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (cap >> 1); t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));        // insert a 0 at bit log2(j)
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const double dl = bd[lo], dh = bd[hi];
        const int il = bi[lo], ih = bi[hi];
        if (knn_less(dh, ih, dl, il) == up) { bd[lo] = dh; bd[hi] = dl; bi[lo] = ih; bi[hi] = il; }
      }
      __syncwarp();
    }
  }
}

// xs, ys: whitened coordinates in x-sorted order; perm: sorted position -> peak (0-based)
// out_bg: P x niter column-major, 1-based, row = peak;  out_nn: [P][window] 1-based (row = peak, sorted
// by (distance, index); -1 beyond the number of other peaks) or NULL
//
// Candidates that beat the current bound are APPENDED to a buffer of cap >= 2 W entries; when it is
// full the warp sorts it, keeps the best W and tightens the bound to the W-th.  (Keeping the list
// exact after every insertion -- find the new worst of W entries each time -- cost 4x more at W = 500.)
// The bound is only as fresh as the last sort, which lets a few more candidates in and a direction run a
// little longer; both are harmless: everything dropped was no better than a bound >= the final one.
__global__ void __launch_bounds__(KNN_WARPS * 32)
k_knn_sample(const double* __restrict__ xs, const double* __restrict__ ys, const int32_t* __restrict__ perm,
             int P, int W, int cap, int niter, uint32_t seed_lo, uint32_t seed_hi,
             int32_t* __restrict__ out_bg, int32_t* __restrict__ out_nn) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* bd = reinterpret_cast<double*>(smem) + (size_t)warp * cap;
  int32_t* bi = reinterpret_cast<int32_t*>(smem + (size_t)KNN_WARPS * cap * 8) + (size_t)warp * cap;
  const int q = blockIdx.x * KNN_WARPS + warp;
  if (q >= P) return;                                  // whole warp
  const double xq = xs[q], yq = ys[q];
  const int qi = perm[q];
  for (int e = lane; e < cap; e += 32) { bd[e] = INFINITY; bi[e] = 0x7fffffff; }
  __syncwarp();
  int m = 0;                                           // entries in the buffer
  double thr_d = INFINITY;                             // bound: the W-th best at the last sort
  int thr_i = 0x7fffffff;
  int l = q - 1, r = q + 1;
  bool ldone = l < 0, rdone = r >= P;

  auto compact = [&]() {                               // best W first, the rest cleared
    knn_sort(bd, bi, cap, lane);
    if (m > W) {
      for (int e = W + lane; e < cap; e += 32) { bd[e] = INFINITY; bi[e] = 0x7fffffff; }
      m = W;
    }
    __syncwarp();
    if (m == W) { thr_d = bd[W - 1]; thr_i = bi[W - 1]; }
  };

  while (!ldone || !rdone) {
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      if (dir == 0 ? rdone : ldone) continue;
      const int pos = dir == 0 ? r + lane : l - lane;
      const bool valid = pos >= 0 && pos < P;
      double d2 = INFINITY, dx2 = INFINITY;
      int idx = 0x7fffffff;
      if (valid) {
        const double dx = __dsub_rn(xs[pos], xq), dy = __dsub_rn(ys[pos], yq);
        dx2 = __dmul_rn(dx, dx);
        d2 = __dadd_rn(dx2, __dmul_rn(dy, dy));
        idx = perm[pos];
      }
      // sorted by x: once the nearest candidate of this direction is farther in x alone than the
      // bound, nothing beyond it can enter the list
      const double dx2_first = __shfl_sync(0xffffffffu, dx2, 0);
      if (dx2_first > thr_d) {
        if (dir == 0) rdone = true; else ldone = true;
        continue;
      }
      const bool take = valid && knn_less(d2, idx, thr_d, thr_i);
      const unsigned mask = __ballot_sync(0xffffffffu, take);
      if (mask) {
        if (m + 32 > cap) compact();                  // (cap >= W + 32: room for a full round afterwards)
        const bool take2 = take && knn_less(d2, idx, thr_d, thr_i);      // the bound may just have tightened
        const unsigned mask2 = __ballot_sync(0xffffffffu, take2);
        if (take2) {
          const int at = m + __popc(mask2 & ((1u << lane) - 1u));
          bd[at] = d2;
          bi[at] = idx;
        }
        m += __popc(mask2);
        __syncwarp();
      }
      if (dir == 0) { r += 32; rdone = r >= P; } else { l -= 32; ldone = l < 0; }
    }
  }
  __syncwarp();
  knn_sort(bd, bi, cap, lane);
  if (m > W) m = W;
  if (out_nn)
    for (int e = lane; e < W; e += 32) out_nn[(int64_t)qi * W + e] = e < m ? bi[e] + 1 : -1;
  for (int it = lane; it < niter; it += 32) {
    const Philox4 rr = philox4x32_10((uint32_t)qi, (uint32_t)it, CV_STREAM_KNN, 0u, seed_lo, seed_hi);
    int j = (int)(u01_53(rr.v[0], rr.v[1]) * (double)m);
    if (j >= m) j = m - 1;
    out_bg[(int64_t)it * P + qi] = bi[j] + 1;
  }
}

extern "C" int cv_background_peaks_knn(cv_handle* h, const double* bias, int niter, int window,
                                       int32_t* out_bg, double* out_trans, int32_t* out_nn) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (niter < 1 || window < 1) CV_FAIL(h, CV_ERR_INVALID, "niterations >= 1 and window >= 1 required");
  if (window > 2048) CV_FAIL(h, CV_ERR_UNSUPPORTED, "window > 2048 is not supported");
  if (!out_bg) CV_FAIL(h, CV_ERR_INVALID, "out_bg is NULL");
  double mm[4];
  CV_TRY(cv_whiten(h, bias, 0, mm));
  const int64_t P = h->P;
  if (window > P - 1) CV_FAIL(h, CV_ERR_INVALID, "window (%d) exceeds the number of other peaks (%lld)", window, (long long)(P - 1));
  const double* d_T = h->bk[1].as<double>();
  // buffers: 2 perm (sorted -> peak), 4 iota, 5 sorted x, 6 sorted y, 7 radix-sort scratch, 11 out, 8 nn
  CV_TRY(cv_ensure(h, h->bk[2], (size_t)P * 4));
  CV_TRY(cv_ensure(h, h->bk[4], (size_t)P * 4));
  CV_TRY(cv_ensure(h, h->bk[5], (size_t)P * 8));
  CV_TRY(cv_ensure(h, h->bk[6], (size_t)P * 8));
  CV_TRY(cv_ensure(h, h->bk[11], (size_t)P * niter * 4));
  if (out_nn) CV_TRY(cv_ensure(h, h->bk[8], (size_t)P * window * 4));
  int32_t* d_perm = h->bk[2].as<int32_t>();
  int32_t* d_iota = h->bk[4].as<int32_t>();
  double* d_xs = h->bk[5].as<double>();
  double* d_ys = h->bk[6].as<double>();
  int32_t* d_out = h->bk[11].as<int32_t>();
  int32_t* d_nn = out_nn ? h->bk[8].as<int32_t>() : nullptr;
  CV_LAUNCH(h, k_knn_iota, (unsigned)((P + 255) / 256), 256, 0, d_iota, (int)P);
  size_t tmp_bytes = 0;
  CV_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_T, d_xs, d_iota, d_perm, (int)P, 0, 64, h->stream));
  CV_TRY(cv_ensure(h, h->bk[7], tmp_bytes));
  CV_CUDA(h, cub::DeviceRadixSort::SortPairs(h->bk[7].p, tmp_bytes, d_T, d_xs, d_iota, d_perm, (int)P, 0, 64, h->stream));
  h->launches += 4;                                    // the radix sort's own kernels
  CV_LAUNCH(h, k_knn_gather_y, (unsigned)((P + 255) / 256), 256, 0, d_T, d_perm, (int)P, d_ys);
  if (window <= 16 || h->opt_knn_insertion) {           // ("knn_insertion" = 1: the exact-insertion kernel at any window)
    const int Wcap = (window + 31) / 32 * 32;
    const size_t smem = (size_t)KNN_WARPS * Wcap * 16;
    if (smem > 48 * 1024)
      CV_CUDA(h, cudaFuncSetAttribute(k_knn_sample_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CV_LAUNCH(h, k_knn_sample_small, (unsigned)((P + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem,
              d_xs, d_ys, d_perm, (int)P, window, Wcap, niter, (uint32_t)h->seed, (uint32_t)(h->seed >> 32), d_out, d_nn);
  } else {
    // buffer: a power of two >= 2 W with room for at least W / 2 appended candidates (+ one round of 32)
    // between two sorts -- at W = 32 a buffer of 64 entries was sorted after every round (9.2 vs 4.6 ms)
    int Wcap = 64;
    while (Wcap < 2 * window || Wcap - window - 32 < window / 2) Wcap <<= 1;
    const size_t smem = (size_t)KNN_WARPS * Wcap * 12;
    if (smem > 48 * 1024)
      CV_CUDA(h, cudaFuncSetAttribute(k_knn_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CV_LAUNCH(h, k_knn_sample, (unsigned)((P + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem, d_xs, d_ys, d_perm,
              (int)P, window, Wcap, niter, (uint32_t)h->seed, (uint32_t)(h->seed >> 32), d_out, d_nn);
  }
  CV_CUDA(h, cudaEventRecord(h->ev[7], h->stream));
  CV_CUDA(h, cudaMemcpyAsync(out_bg, d_out, (size_t)P * niter * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out_trans)
    CV_CUDA(h, cudaMemcpyAsync(out_trans, d_T, (size_t)P * 2 * 8, cudaMemcpyDeviceToHost, h->stream));
  if (out_nn)
    CV_CUDA(h, cudaMemcpyAsync(out_nn, d_nn, (size_t)P * window * 4, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
  h->st.ms_background = ms;
  return CV_OK;
}
