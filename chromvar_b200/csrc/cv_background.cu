// getBackgroundPeaks on the GPU (reference: get_background_peaks_core, R/background_peaks.R:116-156;
// bg_sample_helper / ProbSampleReplace / euc_dist, src/utils.cpp:38-117).
//
//   K2  front half: log10 intensity x GC bias -> sample covariance -> 2x2 Cholesky ->
//       forward-solve (whitening, not mean-centred) -> bs x bs grid -> nearest grid point
//       -> stable counting sort of the peaks by bin (bin density, bin start, peaks by bin)
//   K3  per source bin: weights dnorm(dist(bin_i, bin_j), 0, w) over the occupied destination
//       bins, normalised -> one CDF row per occupied source bin
//   K4  sampling: for every (peak, iteration) a Philox4x32-10 draw keyed by
//       (seed, peak, iteration): destination bin by CDF search, then a peak uniformly inside
//       that bin.  This is the generative statement of what bg_sample_helper's per-bin
//       probability vector p[k] = bin_p[bm[k], i] / density[bm[k]] (src/utils.cpp:89-91) draws.
//
// All reductions are order-fixed (two-stage, final stage on the host) so every GPU given the
// same totals, bias and seed produces the bit-identical matrix -- no broadcast is needed.
#include <math.h>

#include <vector>

#include "cv_internal.cuh"
#include "cv_philox.cuh"

#define RED_BLOCKS 512
#define RED_THREADS 256

struct GridParams {
  double min1, by1, max1, min2, by2, max2;
  int bs;
};

// The front half's scalars stay on the device (no host round trip between its kernels): means,
// Cholesky factor of the 2 x 2 covariance, ranges of the whitened coordinates, and the flags the host
// reads once at the end of the call.
struct FrontState {
  double mx, my, r11, r12, r22, min1, max1, min2, max2;
  uint32_t flags;           // 1: a peak without fragments, 2: covariance not positive definite
  uint32_t pad;
};
enum { FRONT_ZERO_PEAK = 1u, FRONT_NOT_PD = 2u };

__device__ __forceinline__ GridParams grid_from_state(const FrontState* st, int bs) {
  GridParams g;
  g.bs = bs;
  g.min1 = st->min1; g.max1 = st->max1; g.min2 = st->min2; g.max2 = st->max2;
  g.by1 = (g.max1 - g.min1) / (double)(bs - 1);   // seq(min, max, length.out = bs) (:134-137)
  g.by2 = (g.max2 - g.min2) / (double)(bs - 1);
  return g;
}

__device__ __forceinline__ double grid_line(double frm, double by, double to, int x, int n) {
  // base::seq(from, to, length.out = n): from + (0:(n-1)) * by with the last element = to
  if (x == n - 1) return to;
  return __dadd_rn(frm, __dmul_rn((double)x, by));
}

// block tree reduction (fixed order) of NV values per thread
template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double* sh /*[NV][RED_THREADS]*/) {
  for (int i = 0; i < NV; ++i) sh[i * RED_THREADS + threadIdx.x] = v[i];
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int i = 0; i < NV; ++i)
        sh[i * RED_THREADS + threadIdx.x] += sh[i * RED_THREADS + threadIdx.x + o];
    __syncthreads();
  }
  for (int i = 0; i < NV; ++i) v[i] = sh[i * RED_THREADS];
}

// intensity axis: log10(fragments per peak) for getBackgroundPeaks (R/background_peaks.R:128), the raw
// total for get_background_peaks_alternative (:188)
template <bool LOG>
__device__ __forceinline__ double intensity(unsigned long long f) { return LOG ? log10((double)f) : (double)f; }

// partial[b][0..1] = sum intensity, sum bias over block b's strided elements
template <bool LOG>
__global__ void __launch_bounds__(RED_THREADS)
k_front_sums(const unsigned long long* __restrict__ fpp, const double* __restrict__ bias, int P,
             double* __restrict__ partial, FrontState* st) {
  __shared__ double sh[2 * RED_THREADS];
  double v[2] = {0.0, 0.0};
  bool zero = false;
  for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < P; i += gridDim.x * RED_THREADS) {
    if (fpp[i] == 0ull) zero = true;
    v[0] += intensity<LOG>(fpp[i]);
    v[1] += bias[i];
  }
  if (zero) atomicOr(&st->flags, FRONT_ZERO_PEAK);
  block_reduce_sum<2>(v, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x * 2] = v[0]; partial[blockIdx.x * 2 + 1] = v[1]; }
}

// the final stages of the reductions, one thread, blocks summed in index order (what the host loop
// did before): stage 0 means, 1 covariance -> chol (upper-triangular R with t(R) R = cov), 2 ranges
__global__ void k_front_finish(const double* __restrict__ partial, int G, int P, int stage, FrontState* st) {
  // one warp: the partials are fetched coalesced into shared memory, then lane 0 adds them in index
  // order (a chain of dependent global loads took 40-100 us here -- a fifth of the whole sampler)
  __shared__ double sp[RED_BLOCKS * 4];
  const int width = stage == 0 ? 2 : (stage == 1 ? 3 : 4);
  for (int i = threadIdx.x; i < G * width; i += blockDim.x) sp[i] = partial[i];
  __syncthreads();
  if (threadIdx.x || blockIdx.x) return;
  if (stage == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < G; ++i) { a += sp[i * 2]; b += sp[i * 2 + 1]; }
    st->mx = a / (double)P;
    st->my = b / (double)P;
  } else if (stage == 1) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < G; ++i) { a += sp[i * 3]; b += sp[i * 3 + 1]; c += sp[i * 3 + 2]; }
    const double c11 = a / (double)(P - 1), c12 = b / (double)(P - 1), c22 = c / (double)(P - 1);   // stats::cov, n-1 (:130)
    const double r11 = sqrt(c11);
    const double r12 = c12 / r11;
    const double r22 = sqrt(c22 - r12 * r12);
    st->r11 = r11; st->r12 = r12; st->r22 = r22;
    if (!(r11 > 0.0) || !(r22 > 0.0)) atomicOr(&st->flags, FRONT_NOT_PD);
  } else {
    double m0 = sp[0], m1 = sp[1], m2 = sp[2], m3 = sp[3];
    for (int b = 1; b < G; ++b) {
      m0 = fmin(m0, sp[b * 4 + 0]); m1 = fmax(m1, sp[b * 4 + 1]);
      m2 = fmin(m2, sp[b * 4 + 2]); m3 = fmax(m3, sp[b * 4 + 3]);
    }
    st->min1 = m0; st->max1 = m1; st->min2 = m2; st->max2 = m3;
  }
}

template <bool LOG>
__global__ void __launch_bounds__(RED_THREADS)
k_front_cov(const unsigned long long* __restrict__ fpp, const double* __restrict__ bias, int P,
            const FrontState* __restrict__ st, double* __restrict__ partial) {
  __shared__ double sh[3 * RED_THREADS];
  double v[3] = {0.0, 0.0, 0.0};
  const double mx = st->mx, my = st->my;
  for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < P; i += gridDim.x * RED_THREADS) {
    double dx = intensity<LOG>(fpp[i]) - mx, dy = bias[i] - my;
    v[0] += dx * dx; v[1] += dx * dy; v[2] += dy * dy;
  }
  block_reduce_sum<3>(v, sh);
  if (threadIdx.x == 0)
    for (int i = 0; i < 3; ++i) partial[blockIdx.x * 3 + i] = v[i];
}

// trans = t(forwardsolve(t(R), t(norm_mat)))  (:131): y1 = x / r11, y2 = (b - r12*y1) / r22
// T is P x 2 column-major.  partial[b] = (min1, max1, min2, max2)
template <bool LOG>
__global__ void __launch_bounds__(RED_THREADS)
k_front_transform(const unsigned long long* __restrict__ fpp, const double* __restrict__ bias,
                  int P, const FrontState* __restrict__ st, double* __restrict__ T,
                  double* __restrict__ partial) {
  __shared__ double sh[4 * RED_THREADS];
  const double r11 = st->r11, r12 = st->r12, r22 = st->r22;
  double mn1 = INFINITY, mx1 = -INFINITY, mn2 = INFINITY, mx2 = -INFINITY;
  for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < P; i += gridDim.x * RED_THREADS) {
    double y1 = intensity<LOG>(fpp[i]) / r11;
    double y2 = __ddiv_rn(__dsub_rn(bias[i], __dmul_rn(r12, y1)), r22);
    T[i] = y1; T[(int64_t)P + i] = y2;
    mn1 = fmin(mn1, y1); mx1 = fmax(mx1, y1); mn2 = fmin(mn2, y2); mx2 = fmax(mx2, y2);
  }
  sh[threadIdx.x] = mn1; sh[RED_THREADS + threadIdx.x] = mx1;
  sh[2 * RED_THREADS + threadIdx.x] = mn2; sh[3 * RED_THREADS + threadIdx.x] = mx2;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + o]);
      sh[RED_THREADS + threadIdx.x] = fmax(sh[RED_THREADS + threadIdx.x], sh[RED_THREADS + threadIdx.x + o]);
      sh[2 * RED_THREADS + threadIdx.x] = fmin(sh[2 * RED_THREADS + threadIdx.x], sh[2 * RED_THREADS + threadIdx.x + o]);
      sh[3 * RED_THREADS + threadIdx.x] = fmax(sh[3 * RED_THREADS + threadIdx.x], sh[3 * RED_THREADS + threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int i = 0; i < 4; ++i) partial[blockIdx.x * 4 + i] = sh[i * RED_THREADS];
}

// nearest grid line on one axis; ties -> lower index (first minimum, like which.min)
__device__ __forceinline__ int nearest_line(double t, double frm, double by, double to, int n) {
  if (n == 1 || !(by > 0.0)) return 0;
  int g = (int)floor((t - frm) / by);
  int lo = max(0, min(n - 1, g - 1)), hi = max(0, min(n - 1, g + 2));
  int best = lo;
  double bd = fabs(t - grid_line(frm, by, to, lo, n));
  for (int x = lo + 1; x <= hi; ++x) {
    double d = fabs(t - grid_line(frm, by, to, x, n));
    if (d < bd) { bd = d; best = x; }
  }
  return best;
}

// bin_membership (0-based) = nabor::knn(bin_data, trans, k = 1) on the product grid (:139-147):
// bin id = x * bs + y with x the intensity axis.
__global__ void k_membership(const double* __restrict__ T, int P, const FrontState* __restrict__ st, int bs,
                             int32_t* __restrict__ memb0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const GridParams g = grid_from_state(st, bs);
  int x = nearest_line(T[i], g.min1, g.by1, g.max1, g.bs);
  int y = nearest_line(T[(int64_t)P + i], g.min2, g.by2, g.max2, g.bs);
  memb0[i] = x * g.bs + y;
}

// ---- stable counting sort of the peaks by bin ------------------------------------------------
// One warp ranks 1024 consecutive peaks: rank inside the block's bin group, in peak order.
#define RANK_CHUNK 1024
__global__ void __launch_bounds__(32)
k_bin_rank(const int32_t* __restrict__ memb0, int P, int nb, int32_t* __restrict__ local_rank,
           uint32_t* __restrict__ blockhist) {
  extern __shared__ uint32_t hist[];
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x;
  for (int b = lane; b < nb; b += 32) hist[b] = 0;
  __syncwarp();
  int base = blockIdx.x * RANK_CHUNK;
  for (int r = 0; r < RANK_CHUNK / 32; ++r) {
    int i = base + r * 32 + lane;
    bool act = i < P;
    int b = act ? memb0[i] : -1 - lane;          // inactive lanes get unique keys
    unsigned m = __match_any_sync(full, b);
    int leader = __ffs(m) - 1;
    uint32_t old = 0;
    if (act && lane == leader) { old = hist[b]; hist[b] = old + __popc(m); }
    old = __shfl_sync(full, old, leader);
    if (act) local_rank[i] = (int32_t)(old + __popc(m & ((1u << lane) - 1u)));
    __syncwarp();
  }
  for (int b = lane; b < nb; b += 32) blockhist[(int64_t)blockIdx.x * nb + b] = hist[b];
}

// per bin: exclusive scan over the blocks (in place), total -> density
__global__ void k_bin_offsets(uint32_t* __restrict__ blockhist, int nblk, int nb,
                              uint32_t* __restrict__ density) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t run = 0;
  int k = 0;
  for (; k + 8 <= nblk; k += 8) {                  // eight independent loads in flight, then the in-place prefix
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = blockhist[(int64_t)(k + u) * nb + b];
#pragma unroll
    for (int u = 0; u < 8; ++u) { blockhist[(int64_t)(k + u) * nb + b] = run; run += v[u]; }
  }
  for (; k < nblk; ++k) {
    uint32_t v = blockhist[(int64_t)k * nb + b];
    blockhist[(int64_t)k * nb + b] = run;
    run += v;
  }
  density[b] = run;
}

__global__ void __launch_bounds__(1024)
k_scan_small_u32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) out[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void k_bin_scatter(const int32_t* __restrict__ memb0, const int32_t* __restrict__ local_rank,
                              const uint32_t* __restrict__ blockoff, const uint32_t* __restrict__ start,
                              int P, int nb, int32_t* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int b = memb0[i];
  sorted[start[b] + blockoff[(int64_t)(i / RANK_CHUNK) * nb + b] + local_rank[i]] = i;
}

// ---- K3: bin weights ----------------------------------------------------------------------------
// stats::dnorm(x, 0, sigma) following R's nmath/dnorm.c (plain form below 5 sigma, split
// exponent beyond, 0 past the underflow bound)
__device__ __forceinline__ double dnorm0(double x, double sigma) {
  const double M_1_SQRT_2PI_ = 0.398942280401432677939946059934;
  double z = fabs(x / sigma);
  if (z < 5.0) return M_1_SQRT_2PI_ * exp(-0.5 * z * z) / sigma;
  if (z > 38.56804181549334) return 0.0;   // sqrt(-2 ln2 (DBL_MIN_EXP + 1 - DBL_MANT_DIG))
  double x1 = ldexp(rint(ldexp(z, 16)), -16);
  double x2 = z - x1;
  return M_1_SQRT_2PI_ / sigma * (exp(-0.5 * x1 * x1) * exp((-0.5 * x2 - x1) * x2));
}

// One CTA per source bin i: wt[j] = weight(i, j) * [density[j] > 0], prefix-summed into the
// CDF row cdf[i][.]; optionally the normalised probabilities.  The running sum is not bit-equal to
// the total, so the last occupied bin (and every empty bin behind it) is written as exactly 1: a
// draw u1 in [0, 1) can then never fall past the last occupied bin.
// weight(i, j) = binp[j + i*nb] when an explicit bin_p matrix is given (cv_bg_sample_helper),
// else dnorm(euc_dist(bin_i, bin_j), 0, w) on the grid (R/background_peaks.R:144-145).
#define CDF_THREADS 256
__global__ void __launch_bounds__(CDF_THREADS)
k_bin_cdf(const double* __restrict__ binp, const FrontState* __restrict__ fst, int bs, double w, int nb,
          const uint32_t* __restrict__ density, double* __restrict__ cdf,
          double* __restrict__ prob, int all_rows) {
  __shared__ double sh[CDF_THREADS];
  const int i = blockIdx.x;
  GridParams g = {};
  g.bs = 1;
  if (!binp) g = grid_from_state(fst, bs);
  if (!all_rows && density[i] == 0) return;
  const int chunk = (nb + CDF_THREADS - 1) / CDF_THREADS;
  const int j0 = threadIdx.x * chunk, j1 = min(nb, j0 + chunk);
  double xi = 0, yi = 0;
  if (!binp) {
    xi = grid_line(g.min1, g.by1, g.max1, i / g.bs, g.bs);
    yi = grid_line(g.min2, g.by2, g.max2, i % g.bs, g.bs);
  }
  auto weight = [&](int j) -> double {
    if (density[j] == 0) return 0.0;
    if (binp) return binp[(int64_t)i * nb + j];
    double dx = xi - grid_line(g.min1, g.by1, g.max1, j / g.bs, g.bs);
    double dy = yi - grid_line(g.min2, g.by2, g.max2, j % g.bs, g.bs);
    double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // euc_dist, src/utils.cpp:107-117
    return dnorm0(d, w);
  };
  double local = 0.0;
  for (int j = j0; j < j1; ++j) local += weight(j);
  sh[threadIdx.x] = local;
  __syncthreads();
  for (int o = 1; o < CDF_THREADS; o <<= 1) {
    double t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0.0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  const double total = sh[CDF_THREADS - 1];
  // last bin with a positive weight (occupied and within dnorm's range)
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = -1;
  __syncthreads();
  int my_last = -1;
  for (int j = j0; j < j1; ++j)
    if (weight(j) > 0.0) my_last = j;
  if (my_last >= 0) atomicMax(&s_last, my_last);
  __syncthreads();
  const int last = s_last;
  double run = sh[threadIdx.x] - local;
  for (int j = j0; j < j1; ++j) {
    double wt = weight(j);
    run += wt;
    cdf[(int64_t)i * nb + j] = j >= last ? 1.0 : run / total;
    if (prob) prob[(int64_t)i * nb + j] = wt / total;
  }
}

// one draw: background peak (0-based) of peak p in iteration it -- Philox counter (p, it), CDF search over
// the destination bins of p's bin, uniform peak inside the bin
__device__ __forceinline__ int bg_draw(int p, int it, const int32_t* __restrict__ memb0, int nb,
                                       const double* __restrict__ cdf, const uint32_t* __restrict__ density,
                                       const uint32_t* __restrict__ start, const int32_t* __restrict__ sorted,
                                       uint32_t seed_lo, uint32_t seed_hi) {
  Philox4 r = philox4x32_10((uint32_t)p, (uint32_t)it, CV_STREAM_BACKGROUND, 0u, seed_lo, seed_hi);
  double u1 = u01_53(r.v[0], r.v[1]), u2 = u01_53(r.v[2], r.v[3]);
  const double* row = cdf + (int64_t)memb0[p] * nb;
  int lo = 0, hi = nb - 1;              // first j with row[j] > u1
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(row + mid) > u1) hi = mid; else lo = mid + 1;
  }
  uint32_t cnt = density[lo];
  while (cnt == 0u && lo > 0) cnt = density[--lo];      // cannot happen with the clamped CDF; cheap to keep safe
  uint32_t within = (uint32_t)(u2 * (double)cnt);
  if (within >= cnt) within = cnt ? cnt - 1 : 0u;
  return sorted[start[lo] + within];
}

// ---- K4: the draws ----------------------------------------------------------------------------------
// out: P x niter column-major, 1-based.  HBM-bound on the 4*P*niter bytes written.
__global__ void __launch_bounds__(256)
k_bg_sample(const int32_t* __restrict__ memb0, int P, int nb, const double* __restrict__ cdf,
            const uint32_t* __restrict__ density, const uint32_t* __restrict__ start,
            const int32_t* __restrict__ sorted, int niter, uint32_t seed_lo, uint32_t seed_hi,
            int32_t* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;    // grid: (peaks / 256, iterations) -- no 64-bit division per draw
  if (p >= P) return;
  for (int it = blockIdx.y; it < niter; it += gridDim.y)
    out[(int64_t)it * P + p] = bg_draw(p, it, memb0, nb, cdf, density, start, sorted, seed_lo, seed_hi) + 1;
}

__global__ void k_u32_to_f64(const uint32_t* __restrict__ in, double* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

// buffers in h->bk[]: 0 bias, 1 T, 2 memb0, 3 partial, 4 local_rank, 5 blockhist, 6 density,
// 7 start, 8 sorted, 9 cdf, 10 prob, 11 out, 12 density f64
static int sample_from_bins(cv_handle* h, int64_t P, const int32_t* d_memb0, int64_t nb,
                            const double* d_binp, const FrontState* d_state, int bs, double w, int niter,
                            int32_t* host_out, double* host_binprob, double* host_density) {
  const int nblk = (int)((P + RANK_CHUNK - 1) / RANK_CHUNK);
  CV_TRY(cv_ensure(h, h->bk[4], (size_t)P * 4));
  CV_TRY(cv_ensure(h, h->bk[5], (size_t)nblk * nb * 4));
  CV_TRY(cv_ensure(h, h->bk[6], (size_t)nb * 4));
  CV_TRY(cv_ensure(h, h->bk[7], (size_t)(nb + 1) * 4));
  CV_TRY(cv_ensure(h, h->bk[8], (size_t)P * 4));
  CV_TRY(cv_ensure(h, h->bk[9], (size_t)nb * nb * 8));
  CV_TRY(cv_ensure(h, h->bk[11], (size_t)P * niter * 4));
  if (nb * 4 > 200 * 1024) CV_FAIL(h, CV_ERR_UNSUPPORTED, "bs^2 = %lld bins do not fit the ranking kernel", (long long)nb);
  CV_CUDA(h, cudaFuncSetAttribute(k_bin_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(nb * 4)));
  CV_LAUNCH(h, k_bin_rank, nblk, 32, (size_t)nb * 4, d_memb0, (int)P, (int)nb, h->bk[4].as<int32_t>(),
            h->bk[5].as<uint32_t>());
  CV_LAUNCH(h, k_bin_offsets, (unsigned)((nb + 127) / 128), 128, 0, h->bk[5].as<uint32_t>(), nblk, (int)nb,
            h->bk[6].as<uint32_t>());
  CV_LAUNCH(h, k_scan_small_u32, 1, 1024, 0, h->bk[6].as<uint32_t>(), h->bk[7].as<uint32_t>(), (int)nb);
  CV_LAUNCH(h, k_bin_scatter, (unsigned)((P + 255) / 256), 256, 0, d_memb0, h->bk[4].as<int32_t>(),
            h->bk[5].as<uint32_t>(), h->bk[7].as<uint32_t>(), (int)P, (int)nb, h->bk[8].as<int32_t>());
  double* d_prob = nullptr;
  if (host_binprob) {
    CV_TRY(cv_ensure(h, h->bk[10], (size_t)nb * nb * 8));
    d_prob = h->bk[10].as<double>();
  }
  CV_LAUNCH(h, k_bin_cdf, (unsigned)nb, CDF_THREADS, 0, d_binp, d_state, bs, w, (int)nb, h->bk[6].as<uint32_t>(),
            h->bk[9].as<double>(), d_prob, host_binprob ? 1 : 0);
  const int64_t total = P * niter;
  // (a variant that stages the CDF row of each source bin in shared memory was measured: equal at
  // P = 5e5, 1.7x slower at P = 1e5, 5x slower at P = 2000 where bins hold a handful of peaks)
  CV_LAUNCH(h, k_bg_sample, dim3((unsigned)((P + 255) / 256), (unsigned)std::min(niter, 65535)), 256, 0, d_memb0, (int)P, (int)nb,
            h->bk[9].as<double>(), h->bk[6].as<uint32_t>(), h->bk[7].as<uint32_t>(),
            h->bk[8].as<int32_t>(), niter, (uint32_t)(h->seed & 0xffffffffu), (uint32_t)(h->seed >> 32),
            h->bk[11].as<int32_t>());
  CV_CUDA(h, cudaEventRecord(h->ev[7], h->stream));
  if (host_out)
    CV_CUDA(h, cudaMemcpyAsync(host_out, h->bk[11].p, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
  if (host_binprob)
    CV_CUDA(h, cudaMemcpyAsync(host_binprob, d_prob, (size_t)nb * nb * 8, cudaMemcpyDeviceToHost, h->stream));
  if (host_density) {
    CV_TRY(cv_ensure(h, h->bk[12], (size_t)nb * 8));
    CV_LAUNCH(h, k_u32_to_f64, (unsigned)((nb + 255) / 256), 256, 0, h->bk[6].as<uint32_t>(), h->bk[12].as<double>(), (int)nb);
    CV_CUDA(h, cudaMemcpyAsync(host_density, h->bk[12].p, (size_t)nb * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

int cv_sample_background(cv_handle* h, int64_t P, const int32_t* d_memb0, int64_t nb,
                         const double* d_binp_colmajor, const double* d_bin_xy, double w, int niter,
                         int32_t* host_out, double* host_binprob, double* host_density) {
  (void)d_bin_xy;
  return sample_from_bins(h, P, d_memb0, nb, d_binp_colmajor, nullptr, 1, w, niter, host_out, host_binprob,
                          host_density);
}

static double sum_partials(const std::vector<double>& v, int stride, int off, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[(size_t)i * stride + off];
  return s;
}

// Whitening shared by cv_background_peaks (log10 intensity) and cv_background_peaks_knn (raw totals):
// state / input checks, bias -> h->bk[0], T = t(forwardsolve(t(chol(cov(norm_mat))), t(norm_mat))) -> h->bk[1]
// (P x 2 column-major), mm = (min1, max1, min2, max2).  Records h->ev[6] after the bias upload.
int cv_whiten(cv_handle* h, const double* bias, int log_intensity, double* mm) {
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  if (!h->totals_committed)
    CV_FAIL(h, CV_ERR_STATE, "per-peak totals handed out for a cross-shard reduction; call cv_peak_totals_commit first");
  if (!bias) CV_FAIL(h, CV_ERR_INVALID, "bias is NULL");
  const int64_t P = h->P;
  if (P < 2) CV_FAIL(h, CV_ERR_INVALID, "need at least two peaks");
  const unsigned long long* fpp = h->peak_tot.as<unsigned long long>();
  for (int64_t i = 0; i < P; ++i)
    if (!(bias[i] == bias[i]) || isinf(bias[i])) CV_FAIL(h, CV_ERR_INVALID, "bias holds a non-finite value at peak %lld", (long long)(i + 1));

  CV_TRY(cv_ensure(h, h->bk[0], (size_t)P * 8));
  CV_TRY(cv_ensure(h, h->bk[1], (size_t)P * 2 * 8));
  CV_TRY(cv_ensure(h, h->bk[3], (size_t)RED_BLOCKS * 4 * 8 + sizeof(FrontState)));
  double* d_bias = h->bk[0].as<double>();
  double* d_T = h->bk[1].as<double>();
  double* d_part = h->bk[3].as<double>();
  FrontState* d_st = reinterpret_cast<FrontState*>(d_part + RED_BLOCKS * 4);
  CV_CUDA(h, cudaMemcpyAsync(d_bias, bias, (size_t)P * 8, cudaMemcpyHostToDevice, h->stream));
  CV_CUDA(h, cudaEventRecord(h->ev[6], h->stream));
  CV_CUDA(h, cudaMemsetAsync(d_st, 0, sizeof(FrontState), h->stream));
  const int G = (int)std::min<int64_t>(RED_BLOCKS, (P + RED_THREADS - 1) / RED_THREADS);

  // sums -> means -> covariance -> chol -> forward solve + ranges: every scalar stays on the device
  // (k_front_finish), so the seven launches queue without a host round trip in between
  if (log_intensity) CV_LAUNCH(h, k_front_sums<true>, G, RED_THREADS, 0, fpp, d_bias, (int)P, d_part, d_st);
  else CV_LAUNCH(h, k_front_sums<false>, G, RED_THREADS, 0, fpp, d_bias, (int)P, d_part, d_st);
  CV_LAUNCH(h, k_front_finish, 1, 256, 0, d_part, G, (int)P, 0, d_st);
  if (log_intensity) CV_LAUNCH(h, k_front_cov<true>, G, RED_THREADS, 0, fpp, d_bias, (int)P, d_st, d_part);
  else CV_LAUNCH(h, k_front_cov<false>, G, RED_THREADS, 0, fpp, d_bias, (int)P, d_st, d_part);
  CV_LAUNCH(h, k_front_finish, 1, 256, 0, d_part, G, (int)P, 1, d_st);
  if (log_intensity) CV_LAUNCH(h, k_front_transform<true>, G, RED_THREADS, 0, fpp, d_bias, (int)P, d_st, d_T, d_part);
  else CV_LAUNCH(h, k_front_transform<false>, G, RED_THREADS, 0, fpp, d_bias, (int)P, d_st, d_T, d_part);
  CV_LAUNCH(h, k_front_finish, 1, 256, 0, d_part, G, (int)P, 2, d_st);
  if (mm) {                                              // a caller that needs the ranges on the host
    CV_TRY(cv_whiten_verdict(h, log_intensity));
    FrontState hs;
    CV_CUDA(h, cudaMemcpy(&hs, d_st, sizeof(hs), cudaMemcpyDeviceToHost));
    mm[0] = hs.min1; mm[1] = hs.max1; mm[2] = hs.min2; mm[3] = hs.max2;
  }
  return CV_OK;
}

// reads the front half's flags (one stream synchronisation): the reference's stop() conditions
int cv_whiten_verdict(cv_handle* h, int log_intensity) {
  FrontState hs;
  const FrontState* d_st = reinterpret_cast<const FrontState*>(h->bk[3].as<double>() + RED_BLOCKS * 4);
  CV_CUDA(h, cudaMemcpyAsync(&hs, d_st, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  // (:122-125, :183-184) every peak needs a fragment
  if (hs.flags & FRONT_ZERO_PEAK) CV_FAIL(h, CV_ERR_INVALID, "All peaks must have at least one fragment in one sample");
  if (hs.flags & FRONT_NOT_PD)
    CV_FAIL(h, CV_ERR_INVALID, "covariance of (%sfragments, bias) is not positive definite (chol would fail)",
            log_intensity ? "log10 " : "");
  return CV_OK;
}

extern "C" int cv_background_peaks(cv_handle* h, const double* bias, int niter, double w, int bs,
                                   int32_t* out_bg, double* out_trans, int32_t* out_membership,
                                   double* out_density, double* out_binprob) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (niter < 1 || bs < 2 || !(w > 0.0)) CV_FAIL(h, CV_ERR_INVALID, "niterations >= 1, bs >= 2 and w > 0 required");
  if (bs > 200) CV_FAIL(h, CV_ERR_UNSUPPORTED, "bs > 200 is not supported");
  const int64_t P = h->P;
  const int64_t nb = (int64_t)bs * bs;
  CV_TRY(cv_whiten(h, bias, 1, nullptr));
  CV_TRY(cv_ensure(h, h->bk[2], (size_t)P * 4));
  double* d_T = h->bk[1].as<double>();
  int32_t* d_memb = h->bk[2].as<int32_t>();
  const FrontState* d_st = reinterpret_cast<const FrontState*>(h->bk[3].as<double>() + RED_BLOCKS * 4);
  CV_LAUNCH(h, k_membership, (unsigned)((P + 255) / 256), 256, 0, d_T, (int)P, d_st, bs, d_memb);
  // (a zero peak makes log10 -inf and everything behind it garbage, but nothing out of bounds: bins are
  // clamped to the grid; the verdict below discards the result)
  CV_TRY(sample_from_bins(h, P, d_memb, nb, nullptr, d_st, bs, w, niter, out_bg, out_binprob, out_density));
  CV_TRY(cv_whiten_verdict(h, 1));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
  h->st.ms_background = ms;

  if (out_trans)
    CV_CUDA(h, cudaMemcpyAsync(out_trans, d_T, (size_t)P * 2 * 8, cudaMemcpyDeviceToHost, h->stream));
  if (out_membership)
    CV_CUDA(h, cudaMemcpyAsync(out_membership, d_memb, (size_t)P * 4, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  if (out_membership)
    for (int64_t i = 0; i < P; ++i) out_membership[i] += 1;   // 1-based like nn.idx
  return CV_OK;
}

// =========================================================================================================
// getPermutedData (R/background_peaks.R:178-203): a permuted assay = reorder_columns(counts, bg) with
// bg = getBackgroundPeaks(object, niterations = ncol(object)): column c of the result is
// counts[bg[, c], c].  The P x N index matrix never exists: row r of column c draws its source peak
// on the fly (the sampler's counter is (peak, iteration = c), so this IS the matrix getBackgroundPeaks
// would return), looks the source up in a dense copy of counts column c and keeps the nonzeros.
// Two passes over the draws (count, then fill in row order); one CTA per column at a time.
// =========================================================================================================
#define PERM_THREADS 512
struct PermTables {
  const int32_t* memb0; const double* cdf; const uint32_t* density; const uint32_t* start; const int32_t* sorted;
  int nb; uint32_t seed_lo, seed_hi;
};

template <bool FILL>
__global__ void __launch_bounds__(PERM_THREADS)
k_perm_gather(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx, const uint32_t* __restrict__ val,
              int P, int N, PermTables t, uint32_t* __restrict__ dense /*[gridDim.x][P], zero*/,
              uint32_t* __restrict__ colcnt, const uint32_t* __restrict__ outptr, int32_t* __restrict__ out_row,
              double* __restrict__ out_x) {
  __shared__ uint32_t s_warp[PERM_THREADS / 32];
  __shared__ uint32_t s_base;
  uint32_t* d = dense + (int64_t)blockIdx.x * P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = blockIdx.x; c < N; c += gridDim.x) {
    const int64_t e0 = colptr[c], e1 = colptr[c + 1];
    for (int64_t e = e0 + threadIdx.x; e < e1; e += PERM_THREADS) d[rowidx[e]] = val[e];
    if (threadIdx.x == 0) s_base = FILL ? outptr[c] : 0u;
    __syncthreads();
    uint32_t total = 0;
    for (int r0 = 0; r0 < P; r0 += PERM_THREADS) {
      const int r = r0 + threadIdx.x;
      uint32_t v = 0u;
      if (r < P) v = d[bg_draw(r, c, t.memb0, t.nb, t.cdf, t.density, t.start, t.sorted, t.seed_lo, t.seed_hi)];
      const unsigned m = __ballot_sync(0xffffffffu, v != 0u);
      if (!FILL) {
        if (lane == 0) total += __popc(m);
      } else {
        // ordered compaction of this tile of rows: warp counts -> offsets -> writes
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, all = 0;
        for (int w = 0; w < PERM_THREADS / 32; ++w) {
          const uint32_t n = s_warp[w];
          if (w < warp) before += n;
          all += n;
        }
        if (v != 0u) {
          const uint32_t pos = s_base + before + __popc(m & ((1u << lane) - 1u));
          out_row[pos] = r;
          out_x[pos] = (double)v;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += all;
      }
    }
    if (!FILL) {
      if (lane == 0) s_warp[warp] = total;
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < PERM_THREADS / 32; ++w) s += s_warp[w];
        colcnt[c] = s;
      }
    }
    __syncthreads();
    for (int64_t e = e0 + threadIdx.x; e < e1; e += PERM_THREADS) d[rowidx[e]] = 0u;     // leave the scratch zeroed
    __syncthreads();
  }
}

__global__ void k_u32_prefix_to_i64(const uint32_t* __restrict__ in, int64_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i];
}

// buffers: bk[13] dense scratch, bk[14] column counts / offsets (u32[N+1] x 2), bk[15] colptr i64; results in tmp1 / tmp2
extern "C" int cv_permuted_counts(cv_handle* h, const double* bias, double w, int bs, int64_t* out_colptr,
                                  int64_t* out_nnz) {
  if (!h || !out_colptr || !out_nnz) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  // the sampler's tables for this (counts, bias, w, bs): whitening, bins, CDF rows (one iteration drawn on the way)
  CV_TRY(cv_background_peaks(h, bias, 1, w, bs, nullptr, nullptr, nullptr, nullptr, nullptr));
  const int64_t P = h->P, N = h->N;
  const int grid = (int)std::min<int64_t>(N, (int64_t)h->num_sms * 2);
  CV_TRY(cv_ensure(h, h->bk[13], (size_t)grid * P * 4));
  CV_TRY(cv_ensure(h, h->bk[14], (size_t)(N + 1) * 4 * 2));
  CV_TRY(cv_ensure(h, h->bk[15], (size_t)(N + 1) * 8));
  uint32_t* cnt = h->bk[14].as<uint32_t>();
  uint32_t* off = cnt + (N + 1);
  CV_CUDA(h, cudaMemsetAsync(h->bk[13].p, 0, (size_t)grid * P * 4, h->stream));
  CV_CUDA(h, cudaMemsetAsync(cnt, 0, (size_t)(N + 1) * 4, h->stream));
  PermTables t;
  t.memb0 = h->bk[2].as<int32_t>(); t.cdf = h->bk[9].as<double>(); t.density = h->bk[6].as<uint32_t>();
  t.start = h->bk[7].as<uint32_t>(); t.sorted = h->bk[8].as<int32_t>(); t.nb = bs * bs;
  t.seed_lo = (uint32_t)(h->seed & 0xffffffffu); t.seed_hi = (uint32_t)(h->seed >> 32);
  CV_LAUNCH(h, k_perm_gather<false>, grid, PERM_THREADS, 0, h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(),
            h->val.as<uint32_t>(), (int)P, (int)N, t, h->bk[13].as<uint32_t>(), cnt, nullptr, nullptr, nullptr);
  CV_TRY(cv_exclusive_scan_u32_async(h, cnt, off, N));
  CV_LAUNCH(h, k_u32_prefix_to_i64, (unsigned)((N + 1 + 255) / 256), 256, 0, off, h->bk[15].as<int64_t>(), (int)(N + 1));
  CV_CUDA(h, cudaMemcpyAsync(out_colptr, h->bk[15].p, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  const int64_t nnz = out_colptr[N];
  if (nnz >= (1ll << 32)) CV_FAIL(h, CV_ERR_UNSUPPORTED, "permuted counts hold >= 2^32 nonzeros");
  *out_nnz = nnz;
  CV_TRY(cv_ensure(h, h->tmp1, (size_t)std::max<int64_t>(nnz, 1) * 4));
  CV_TRY(cv_ensure(h, h->tmp2, (size_t)std::max<int64_t>(nnz, 1) * 8));
  CV_LAUNCH(h, k_perm_gather<true>, grid, PERM_THREADS, 0, h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(),
            h->val.as<uint32_t>(), (int)P, (int)N, t, h->bk[13].as<uint32_t>(), cnt, off, h->tmp1.as<int32_t>(),
            h->tmp2.as<double>());
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

extern "C" int cv_permuted_counts_fetch(cv_handle* h, int64_t nnz, int32_t* out_rowidx, double* out_x) {
  if (!h || nnz < 0) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (h->tmp1.cap < (size_t)nnz * 4 || h->tmp2.cap < (size_t)nnz * 8) CV_FAIL(h, CV_ERR_STATE, "no permuted counts: call cv_permuted_counts first");
  if (out_rowidx && nnz) CV_CUDA(h, cudaMemcpyAsync(out_rowidx, h->tmp1.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out_x && nnz) CV_CUDA(h, cudaMemcpyAsync(out_x, h->tmp2.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}
