// GPU equivalents of the reference's five registered native routines
// (src/RcppExports.cpp:59-119 -> src/utils.cpp):
//   row_sds            src/utils.cpp:9-25     cv_row_sds
//   row_sds_perm       src/utils.cpp:28-33    cv_row_sds_perm   (batched over bootstrap replicates)
//   ProbSampleReplace  src/utils.cpp:38-77    cv_prob_sample_replace
//   bg_sample_helper   src/utils.cpp:83-99    cv_bg_sample_helper
//   euc_dist           src/utils.cpp:107-117  cv_euc_dist
// The two samplers draw from the same distributions as the reference but from Philox streams
// instead of R's Mersenne-Twister (RNG parity with R is impossible by construction).
#include <math.h>

#include "cv_internal.cuh"
#include "cv_philox.cuh"

struct W3 {
  double n, mean, m2;
};
__device__ __forceinline__ void w3_add(W3& w, double x) {
  w.n += 1.0;
  double d = x - w.mean;
  w.mean += d / w.n;
  w.m2 += d * (x - w.mean);
}
__device__ __forceinline__ W3 w3_merge(const W3& a, const W3& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  W3 r;
  r.n = a.n + b.n;
  double d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}

__device__ __forceinline__ double sd_finish(const W3& w, double nonfinite, int na_rm, int64_t N) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (na_rm) return w.n > 1.0 ? sqrt(w.m2 / (w.n - 1.0)) : nan;   // :18-24
  if (nonfinite > 0.0) return nan;                                 // arma::stddev over NaN / Inf
  return N > 1 ? sqrt(w.m2 / (w.n - 1.0)) : 0.0;                  // :16
}

// Block = 32 rows (x) x 8 column lanes (y).  x is K x N column-major, so for a fixed column the
// 32 rows of a block are one contiguous 256-byte segment.  rep < 0: plain row_sds; otherwise
// the columns are resampled with replacement: draw d of replicate rep takes column
// floor(u * N), u = Philox(counter = (d, rep), stream BOOTSTRAP)  (RcppArmadillo::sample,
// src/utils.cpp:30).
#define SD_LANES 8
__global__ void __launch_bounds__(32 * SD_LANES)
k_row_sds(const double* __restrict__ x, int64_t K, int64_t N, int na_rm, int n_rep,
          uint32_t seed_lo, uint32_t seed_hi, double* __restrict__ out) {
  __shared__ double sh[SD_LANES][32][4];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * 32 + tx;
  const int rep = n_rep > 0 ? (int)blockIdx.y : -1;
  W3 w = {0.0, 0.0, 0.0};
  double nonfinite = 0.0;
  for (int64_t d = ty; d < N; d += SD_LANES) {
    int64_t c = d;
    if (rep >= 0) {
      Philox4 r = philox4x32_10((uint32_t)d, (uint32_t)rep, CV_STREAM_BOOTSTRAP, (uint32_t)(d >> 32),
                                seed_lo, seed_hi);
      c = (int64_t)(u01_53(r.v[0], r.v[1]) * (double)N);
      if (c >= N) c = N - 1;
    }
    if (k < K) {
      double v = x[c * K + k];
      if (isfinite(v)) w3_add(w, v); else nonfinite += 1.0;
    }
  }
  sh[ty][tx][0] = w.n; sh[ty][tx][1] = w.mean; sh[ty][tx][2] = w.m2; sh[ty][tx][3] = nonfinite;
  __syncthreads();
  if (ty == 0 && k < K) {
    W3 tot = {0.0, 0.0, 0.0};
    double nf = 0.0;
    for (int l = 0; l < SD_LANES; ++l) {
      W3 b = {sh[l][tx][0], sh[l][tx][1], sh[l][tx][2]};
      tot = w3_merge(tot, b);
      nf += sh[l][tx][3];
    }
    out[(int64_t)(rep < 0 ? 0 : rep) * K + k] = sd_finish(tot, nf, na_rm, N);
  }
}

static int row_sds_impl(cv_handle* h, int64_t K, int64_t N, const double* x, int na_rm, int n_rep,
                        double* out_sd) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (K < 1 || N < 1 || !x || !out_sd) CV_FAIL(h, CV_ERR_INVALID, "row_sds: empty input");
  if (n_rep > 65535) CV_FAIL(h, CV_ERR_UNSUPPORTED, "at most 65535 bootstrap replicates per call");
  const int reps = n_rep > 0 ? n_rep : 1;
  CV_TRY(cv_ensure(h, h->bk[13], (size_t)K * N * 8));
  CV_TRY(cv_ensure(h, h->bk[14], (size_t)reps * K * 8));
  CV_CUDA(h, cudaMemcpyAsync(h->bk[13].p, x, (size_t)K * N * 8, cudaMemcpyHostToDevice, h->stream));
  dim3 grid((unsigned)((K + 31) / 32), (unsigned)reps);
  CV_LAUNCH(h, k_row_sds, grid, 32 * SD_LANES, 0, h->bk[13].as<double>(), K, N, na_rm, n_rep,
            (uint32_t)(h->seed & 0xffffffffu), (uint32_t)(h->seed >> 32), h->bk[14].as<double>());
  CV_CUDA(h, cudaMemcpyAsync(out_sd, h->bk[14].p, (size_t)reps * K * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

extern "C" int cv_row_sds(cv_handle* h, int64_t K, int64_t N, const double* x, int na_rm,
                          double* out_sd) {
  return row_sds_impl(h, K, N, x, na_rm, 0, out_sd);
}

extern "C" int cv_row_sds_perm(cv_handle* h, int64_t K, int64_t N, const double* x, int na_rm,
                               int n_rep, double* out_sd) {
  if (h && n_rep < 1) CV_FAIL(h, CV_ERR_INVALID, "n_rep must be >= 1");
  return row_sds_impl(h, K, N, x, na_rm, n_rep, out_sd);
}

// ---- ProbSampleReplace ------------------------------------------------------------------------
// CDF of prob / sum(prob) (single block, sequential chunks), then `size` independent draws by
// CDF search.  Same distribution as the reference's Walker alias table.
__global__ void __launch_bounds__(1024)
k_cdf_build(const double* __restrict__ prob, int64_t n, double* __restrict__ cdf) {
  __shared__ double sh[1024];
  __shared__ double carry;
  if (threadIdx.x == 0) carry = 0.0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    int64_t i = base + threadIdx.x;
    double v = i < n ? prob[i] : 0.0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      double t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0.0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) cdf[i] = carry + sh[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
}

__global__ void k_cdf_draw(const double* __restrict__ cdf, int64_t n, int64_t size, uint32_t seed_lo,
                           uint32_t seed_hi, int32_t* __restrict__ out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= size) return;
  Philox4 r = philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), CV_STREAM_PROBSAMPLE, 0u, seed_lo, seed_hi);
  const double total = cdf[n - 1];
  double u = u01_53(r.v[0], r.v[1]) * total;
  int64_t lo = 0, hi = n - 1;          // first index with cdf > u
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
  }
  out[k] = (int32_t)(lo + 1);          // 1-based (:76)
}

extern "C" int cv_prob_sample_replace(cv_handle* h, int64_t size, int64_t n, const double* prob,
                                      int32_t* out) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (size < 0 || n < 1 || !prob || (size > 0 && !out)) CV_FAIL(h, CV_ERR_INVALID, "ProbSampleReplace: bad arguments");
  if (n > 0x7fffffff) CV_FAIL(h, CV_ERR_UNSUPPORTED, "n must be < 2^31");
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    if (!(prob[i] >= 0.0) || isinf(prob[i])) CV_FAIL(h, CV_ERR_INVALID, "probabilities must be finite and non-negative");
    s += prob[i];
  }
  if (!(s > 0.0)) CV_FAIL(h, CV_ERR_INVALID, "probabilities sum to zero");
  if (size == 0) return CV_OK;
  CV_TRY(cv_ensure(h, h->bk[13], (size_t)n * 8));
  CV_TRY(cv_ensure(h, h->bk[14], (size_t)n * 8));
  CV_TRY(cv_ensure(h, h->bk[15], (size_t)size * 4));
  CV_CUDA(h, cudaMemcpyAsync(h->bk[13].p, prob, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  CV_LAUNCH(h, k_cdf_build, 1, 1024, 0, h->bk[13].as<double>(), n, h->bk[14].as<double>());
  CV_LAUNCH(h, k_cdf_draw, (unsigned)((size + 255) / 256), 256, 0, h->bk[14].as<double>(), n, size,
            (uint32_t)(h->seed & 0xffffffffu), (uint32_t)(h->seed >> 32), h->bk[15].as<int32_t>());
  CV_CUDA(h, cudaMemcpyAsync(out, h->bk[15].p, (size_t)size * 4, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

// ---- bg_sample_helper ---------------------------------------------------------------------------
extern "C" int cv_bg_sample_helper(cv_handle* h, int64_t P, const int32_t* bin_membership,
                                   int64_t nb, const double* bin_p, const double* bin_density,
                                   int niter, int32_t* out) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (P < 1 || nb < 1 || niter < 1 || !bin_membership || !bin_p || !out)
    CV_FAIL(h, CV_ERR_INVALID, "bg_sample_helper: bad arguments");
  for (int64_t i = 0; i < P; ++i)
    if (bin_membership[i] < 0 || bin_membership[i] >= nb)
      CV_FAIL(h, CV_ERR_INVALID, "bin_membership out of range at peak %lld", (long long)(i + 1));
  (void)bin_density;   // recomputed from bin_membership (the reference passes tabulate2 of it)
  CV_TRY(cv_ensure(h, h->bk[2], (size_t)P * 4));
  CV_TRY(cv_ensure(h, h->bk[13], (size_t)nb * nb * 8));
  CV_CUDA(h, cudaMemcpyAsync(h->bk[2].p, bin_membership, (size_t)P * 4, cudaMemcpyHostToDevice, h->stream));
  CV_CUDA(h, cudaMemcpyAsync(h->bk[13].p, bin_p, (size_t)nb * nb * 8, cudaMemcpyHostToDevice, h->stream));
  return cv_sample_background(h, P, h->bk[2].as<int32_t>(), nb, h->bk[13].as<double>(), nullptr, 0.0,
                              niter, out, nullptr, nullptr);
}

// ---- euc_dist -------------------------------------------------------------------------------------
__global__ void k_euc_dist(const double* __restrict__ x, int64_t n, int64_t d, double* __restrict__ out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t i = blockIdx.y;
  if (j >= n) return;
  double s = 0.0;
  for (int64_t c = 0; c < d; ++c) {
    double t = x[c * n + i] - x[c * n + j];
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  out[i * n + j] = sqrt(s);
}

extern "C" int cv_euc_dist(cv_handle* h, int64_t n, int64_t d, const double* x, double* out) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (n < 1 || d < 1 || !x || !out) CV_FAIL(h, CV_ERR_INVALID, "euc_dist: bad arguments");
  if (n > 65535) CV_FAIL(h, CV_ERR_UNSUPPORTED, "euc_dist: n > 65535");
  CV_TRY(cv_ensure(h, h->bk[13], (size_t)n * d * 8));
  CV_TRY(cv_ensure(h, h->bk[14], (size_t)n * n * 8));
  CV_CUDA(h, cudaMemcpyAsync(h->bk[13].p, x, (size_t)n * d * 8, cudaMemcpyHostToDevice, h->stream));
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
  CV_LAUNCH(h, k_euc_dist, grid, 256, 0, h->bk[13].as<double>(), n, d, h->bk[14].as<double>());
  CV_CUDA(h, cudaMemcpyAsync(out, h->bk[14].p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

// =========================================================================================================
// deviationsCovariability (R/kmers.R:51-57): covs = cov(t(z)) (pairs of annotations over the cells,
// two-pass, n - 1), then row i divided by row_sds(z)[i]^2.  NaN / NA anywhere in rows i or j makes the
// pair NaN, as stats::cov(use = "everything") does.  z stays in HBM: [K][N] row-major as the deviation
// kernels leave it, or a K x N column-major host matrix uploaded first.  fp64 SYRK, 64 x 64 output tile
// per CTA, 4 x 4 per thread, the N axis staged through shared memory 16 cells at a time.
// =========================================================================================================
#define COV_TILE 64
#define COV_KC 16

__global__ void __launch_bounds__(256)
k_row_means(const double* __restrict__ z, int64_t K, int64_t N, int64_t ld_row, int64_t ld_col, double* __restrict__ mean) {
  __shared__ double sh[256];
  const int64_t k = blockIdx.x;
  double s = 0.0;
  for (int64_t c = threadIdx.x; c < N; c += 256) s += z[k * ld_row + c * ld_col];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[k] = sh[0] / (double)N;
}

__global__ void __launch_bounds__(256)
k_cov_tiles(const double* __restrict__ z, int64_t K, int64_t N, int64_t ld_row, int64_t ld_col,
            const double* __restrict__ mean, double* __restrict__ cov /* K x K */) {
  __shared__ double sa[COV_KC][COV_TILE + 1], sb[COV_KC][COV_TILE + 1];
  const int64_t i0 = (int64_t)blockIdx.y * COV_TILE, j0 = (int64_t)blockIdx.x * COV_TILE;
  if (j0 > i0) return;                                   // lower triangle; mirrored on the way out
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4] = {};
  for (int64_t c0 = 0; c0 < N; c0 += COV_KC) {
    for (int t = threadIdx.x; t < COV_TILE * COV_KC; t += 256) {
      // row-major z: consecutive threads read consecutive cells of one row; column-major: consecutive rows
      const int r = ld_col == 1 ? t / COV_KC : t % COV_TILE, cc = ld_col == 1 ? t % COV_KC : t / COV_TILE;
      const int64_t c = c0 + cc;
      const int64_t ia = i0 + r, jb = j0 + r;
      sa[cc][r] = (ia < K && c < N) ? z[ia * ld_row + c * ld_col] - mean[ia] : 0.0;
      sb[cc][r] = (jb < K && c < N) ? z[jb * ld_row + c * ld_col] - mean[jb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < COV_KC; ++cc) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = sa[cc][ty * 4 + u]; b[u] = sb[cc][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
  const double inv = 1.0 / (double)(N - 1);
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < K && j < K) {
        const double cv = acc[u][v] * inv;
        cov[i + j * K] = cv;
        cov[j + i * K] = cv;
      }
    }
}

// covs / matrix(vars^2, byrow = FALSE): entry (i, j) divided by var_i = covs[i, i]
__global__ void k_cov_normalise(const double* __restrict__ cov, int64_t K, double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * K) return;
  const int64_t i = idx % K;
  out[idx] = cov[idx] / cov[i + i * K];
}

extern "C" int cv_deviations_covariability(cv_handle* h, int64_t K, int64_t N, const double* z, double* out) {
  if (!h || !out) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  const double* dz;
  int64_t ld_row, ld_col;
  if (z) {
    if (K < 1 || N < 2) CV_FAIL(h, CV_ERR_INVALID, "covariability needs K >= 1 annotations and N >= 2 cells");
    CV_TRY(cv_ensure(h, h->tmp0, (size_t)K * N * 8));
    CV_CUDA(h, cudaMemcpyAsync(h->tmp0.p, z, (size_t)K * N * 8, cudaMemcpyHostToDevice, h->stream));
    dz = h->tmp0.as<double>(); ld_row = 1; ld_col = K;                  // K x N column-major
  } else {
    if (!h->have_results) CV_FAIL(h, CV_ERR_STATE, "no resident z: call cv_deviations first or pass z");
    K = h->K; N = h->N;
    if (N < 2) CV_FAIL(h, CV_ERR_INVALID, "covariability needs N >= 2 cells");
    dz = h->z_rm.as<double>(); ld_row = N; ld_col = 1;                  // [K][N] row-major, resident
  }
  CV_TRY(cv_ensure(h, h->tmp1, (size_t)K * 8));
  CV_TRY(cv_ensure(h, h->tmp2, (size_t)K * K * 8 * 2));
  double* mean = h->tmp1.as<double>();
  double* cov = h->tmp2.as<double>();
  double* nrm = cov + K * K;
  CV_LAUNCH(h, k_row_means, (unsigned)K, 256, 0, dz, K, N, ld_row, ld_col, mean);
  const unsigned nt = (unsigned)((K + COV_TILE - 1) / COV_TILE);
  CV_LAUNCH(h, k_cov_tiles, dim3(nt, nt), 256, 0, dz, K, N, ld_row, ld_col, mean, cov);
  CV_LAUNCH(h, k_cov_normalise, (unsigned)((K * K + 255) / 256), 256, 0, cov, K, nrm);
  CV_CUDA(h, cudaMemcpyAsync(out, nrm, (size_t)K * K * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}
