// compute_deviations_core on the GPU: the whole bplapply fan-out over compute_deviations_single
// (R/compute_deviations.R:355-408, 451-537) as
//   k_bg_prepare   background matrix -> [P][B] 0-based, validated            (:365, :493-494)
//   k_anno_stats   expected-fraction table E[k][0..B], overlap, matches      (:488, :502-503, :506)
//   k_contract     observed + B background sums for one (annotation, cell tile) in shared
//                  memory (row-gather SpMM over the cell-tiled counts layout) with the fused
//                  fp64 epilogue: raw deviations, mean / sd over iterations, bias-corrected
//                  deviation, z, NA mask, and the per-annotation variability partials
//                  (:486-520; src/utils.cpp:9-25)
//   k_var_merge    variability partials -> row SD of z
//   k_transpose    [K][N] -> K x N column-major (ABI layout)
#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <numeric>

#include "cv_internal.cuh"

#define CV_ENTRY_VAL_BITS 20
#define CV_ENTRY_VAL_MASK 0xFFFFFu
#define CONTRACT_THREADS 512
#define STATS_THREADS 512

enum { FLAG_BADIDX = 8u, FLAG_ZEROPEAK = 16u };

static double wall_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// =========================================================================================
// input preparation
// =========================================================================================
// bg_raw: P x B column-major with index_base  ->  bgT[p][i] 0-based (row-major).
__global__ void __launch_bounds__(256)
k_bg_prepare(const int32_t* __restrict__ bg_raw, int P, int B, int base,
             int32_t* __restrict__ bgT, uint32_t* __restrict__ flags) {
  __shared__ int32_t tile[32][33];
  int p0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  bool bad = false;
  for (int r = ty; r < 32; r += 8) {   // r: iteration within tile, tx: peak within tile
    int p = p0 + tx, i = i0 + r;
    int32_t v = 0;
    if (p < P && i < B) {
      v = bg_raw[(int64_t)i * P + p] - base;
      if (v < 0 || v >= P) { bad = true; v = 0; }
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {   // r: peak within tile, tx: iteration within tile
    int p = p0 + r, i = i0 + tx;
    if (p < P && i < B) bgT[(int64_t)p * B + i] = tile[tx][r];
  }
  if (bad) atomicOr(flags, FLAG_BADIDX);
}

__global__ void k_anno_prepare(const int32_t* __restrict__ in, int64_t n, int P, int base,
                               int32_t* __restrict__ out, uint32_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t v = in[i] - base;
  if (v < 0 || v >= P) { atomicOr(flags, FLAG_BADIDX); v = 0; }
  out[i] = v;
}

// One CTA per annotation (strided): expected-fraction table, overlap, matches.
//   etab[k][0] = sum_{p in set} e[p]                     (tf_vec %*% expectation, :488)
//   etab[k][i] = sum_{p in set} e[bg[p,i]]               (sample_mat %*% expectation, :502)
//   overlap[k] = mean_i( sum_{p in set} mult_set(bg[p,i]) ) / #set      (:506, :477-478)
//   matches[k] = #set / P                                (:533)
// mark: int32[gridDim.x][P], all zero on entry and on exit.
__global__ void __launch_bounds__(STATS_THREADS)
k_anno_stats(const int64_t* __restrict__ annoptr, const int32_t* __restrict__ annoidx,
             const int32_t* __restrict__ bgT, const double* __restrict__ e, int P, int B, int K,
             int32_t* __restrict__ mark, double* __restrict__ etab, double* __restrict__ overlap,
             double* __restrict__ matches) {
  extern __shared__ double s_part[];   // [nwarps][I]
  __shared__ unsigned long long s_ov[STATS_THREADS / 32];
  const int I = B + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = STATS_THREADS / 32;
  int32_t* mk = mark + (int64_t)blockIdx.x * P;
  for (int k = blockIdx.x; k < K; k += gridDim.x) {
    int64_t beg = annoptr[k], end = annoptr[k + 1];
    for (int64_t idx = beg + threadIdx.x; idx < end; idx += STATS_THREADS)
      atomicAdd(mk + annoidx[idx], 1);
    __syncthreads();
    unsigned long long ov = 0;
    for (int j0 = 0; j0 < I; j0 += 32) {
      int j = j0 + lane;
      double se = 0.0;
      if (j < I) {
        for (int64_t idx = beg + warp; idx < end; idx += nwarps) {
          int p = __ldg(annoidx + idx);
          int q = (j == 0) ? p : __ldg(bgT + (int64_t)p * B + (j - 1));
          se += __ldg(e + q);
          if (j > 0) ov += (unsigned long long)mk[q];
        }
        s_part[warp * I + j] = se;
      }
    }
    ov = warp_sum<unsigned long long>(ov);
    if (lane == 0) s_ov[warp] = ov;
    __syncthreads();
    for (int j = threadIdx.x; j < I; j += STATS_THREADS) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += s_part[w * I + j];
      etab[(int64_t)k * I + j] = s;
    }
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int w = 0; w < nwarps; ++w) tot += s_ov[w];
      double cnt = (double)(end - beg);
      matches[k] = cnt / (double)P;
      overlap[k] = (end > beg) ? ((double)tot / (double)B) / cnt : cv_na_real();
    }
    for (int64_t idx = beg + threadIdx.x; idx < end; idx += STATS_THREADS) mk[annoidx[idx]] = 0;
    __syncthreads();
  }
}

// =========================================================================================
// K5b + K6 + K7: the gather-contraction with fused epilogue.
// One work item = (annotation k, cell tile t).  Shared memory holds acc[B+1][W] integer sums:
// row 0 = observed, row i = background iteration i.  A warp takes one peak p of the set at a
// time: its lanes fetch the B+1 source rows q_j (q_0 = p, q_j = bg[p][j]) and their extents
// in tile t's row layout, then every row's entries (cell_in_tile, count) are loaded
// coalesced, 32 per warp instruction, and added into acc[j][cell] with shared-memory atomics.
// =========================================================================================
struct ContractParams {
  int P, N, K, B, W, T;
  int64_t n_items;
  const int64_t* annoptr;
  const int32_t* annoidx;
  const int32_t* bgT;
  const uint32_t* rowptr;    // [T*P + 1]
  const uint32_t* entries;
  const double* etab;        // [K][B+1]
  const double* fps;         // [N]
  const int32_t* order;      // [K] annotations by decreasing set size
  uint32_t* work_ctr;
  double* dev_rm;            // [K][N]
  double* z_rm;              // [K][N]
  double* varpart;           // [K][T][4]  n_finite, mean, M2, n_nonfinite
  long long* obs;            // [K][N] or null
  long long* samp;           // [K][B][N] or null
  unsigned long long* upd_ctr;   // visited entries (instrumentation), may be null
};

struct Welford {
  double n, mean, m2;
};
__device__ __forceinline__ void welford_add(Welford& w, double x) {
  w.n += 1.0;
  double d = x - w.mean;
  w.mean += d / w.n;
  w.m2 += d * (x - w.mean);
}
__device__ __forceinline__ Welford welford_merge(const Welford& a, const Welford& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Welford r;
  r.n = a.n + b.n;
  double d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}

template <typename AccT>
__device__ __forceinline__ void acc_add(AccT* p, uint32_t v);
template <>
__device__ __forceinline__ void acc_add<int>(int* p, uint32_t v) { atomicAdd(p, (int)v); }
template <>
__device__ __forceinline__ void acc_add<long long>(long long* p, uint32_t v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

template <typename AccT>
__global__ void __launch_bounds__(CONTRACT_THREADS, 1)
k_contract(const ContractParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AccT* acc = reinterpret_cast<AccT*>(smem_raw);
  __shared__ int s_item;
  __shared__ double s_winv[1024 + 1];            // 1 / E[k][i]; B <= 1024
  __shared__ double s_w[CONTRACT_THREADS / 32][4];
  const unsigned full = 0xffffffffu;
  const int I = prm.B + 1, W = prm.W, B = prm.B, P = prm.P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = CONTRACT_THREADS / 32;
  constexpr int U = 8;
  uint32_t visited = 0;   // entries visited by this thread (instrumentation; < 2^32 per thread)

  for (;;) {
    if (threadIdx.x == 0) s_item = (int)atomicAdd(prm.work_ctr, 1u);
    __syncthreads();
    const int64_t item = s_item;
    if (item >= prm.n_items) break;
    const int t = (int)(item / prm.K);
    const int k = prm.order[item - (int64_t)t * prm.K];
    const int64_t beg = prm.annoptr[k], end = prm.annoptr[k + 1];
    const int c0 = t * W;
    const int ncell = min(W, prm.N - c0);

    // ---- zero the accumulators, stage 1/E ------------------------------------------------
    {
      uint4* a4 = reinterpret_cast<uint4*>(acc);
      const int n4 = (int)(((size_t)I * W * sizeof(AccT)) >> 4);
      for (int i = threadIdx.x; i < n4; i += CONTRACT_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
      for (int i = threadIdx.x; i < I; i += CONTRACT_THREADS)
        s_winv[i] = 1.0 / prm.etab[(int64_t)k * I + i];
    }
    __syncthreads();

    // ---- gather-contraction ---------------------------------------------------------------
    const uint32_t* __restrict__ rp = prm.rowptr + (int64_t)t * P;
    const uint32_t* __restrict__ ent_base = prm.entries;
    for (int64_t idx = beg + warp; idx < end; idx += NW) {
      const int p = __ldg(prm.annoidx + idx);
      const int32_t* __restrict__ bgrow = prm.bgT + (int64_t)p * B;
      for (int j0 = 0; j0 < I; j0 += 32) {
        const int j = j0 + lane;
        uint32_t s = 0, e = 0;
        if (j < I) {
          const int q = (j == 0) ? p : __ldg(bgrow + (j - 1));
          s = __ldg(rp + q);
          e = __ldg(rp + q + 1);
        }
        const int nj = min(32, I - j0);
        AccT* accj = acc + (size_t)j0 * W;
        for (int jj = 0; jj < nj; jj += U) {
          uint32_t ent[U];
          uint32_t longrows = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t sj = __shfl_sync(full, s, jj + u);
            const uint32_t ej = __shfl_sync(full, e, jj + u);
            const uint32_t x = sj + lane;
            ent[u] = (x < ej) ? __ldg(ent_base + x) : 0u;
            longrows |= (ej - sj > 32u) ? (1u << u) : 0u;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (ent[u]) {
              acc_add<AccT>(accj + (size_t)(jj + u) * W + (ent[u] >> CV_ENTRY_VAL_BITS),
                            ent[u] & CV_ENTRY_VAL_MASK);
              ++visited;
            }
          }
          if (longrows) {   // warp-uniform: rows with more than 32 entries in this tile
            for (int u = 0; u < U; ++u) {
              if (!(longrows & (1u << u))) continue;
              const uint32_t sj = __shfl_sync(full, s, jj + u);
              const uint32_t ej = __shfl_sync(full, e, jj + u);
              for (uint32_t x = sj + 32u + lane; x < ej; x += 32u) {
                const uint32_t en = __ldg(ent_base + x);
                acc_add<AccT>(accj + (size_t)(jj + u) * W + (en >> CV_ENTRY_VAL_BITS),
                              en & CV_ENTRY_VAL_MASK);
                ++visited;
              }
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- fused epilogue (fp64 from exact integer sums) -----------------------------------
    Welford wf = {0.0, 0.0, 0.0};
    double nonfinite = 0.0;
    const bool empty = (end == beg);
    const double E0 = prm.etab[(int64_t)k * I];
    for (int c = threadIdx.x; c < ncell; c += CONTRACT_THREADS) {
      const int64_t o = (int64_t)k * prm.N + c0 + c;
      double devv, zv;
      if (empty) {
        devv = zv = cv_na_real();                       // :458-461
      } else {
        const double fpsv = prm.fps[c0 + c];
        const double expected = E0 * fpsv;              // :488
        const double rf = 1.0 / fpsv;
        const double obs_dev = ((double)acc[c] - expected) / expected;   // :489
        double sum = 0.0;
        for (int i = 1; i <= B; ++i) {
          // (sampled - sampled_expected) / sampled_expected, sampled_expected = E_i * fps (:502-504)
          const double d = (double)acc[(size_t)i * W + c] * s_winv[i] * rf - 1.0;
          sum += d;
        }
        const double mean = sum / (double)B;            // :511
        double ss = 0.0;
        for (int i = 1; i <= B; ++i) {
          const double d = (double)acc[(size_t)i * W + c] * s_winv[i] * rf - 1.0;
          ss += (d - mean) * (d - mean);
        }
        const double sd = sqrt(ss / (double)(B - 1));   // :512
        devv = obs_dev - mean;                          // :514
        zv = devv / sd;                                 // :515
        if (expected < 1.0) devv = zv = cv_na_real();   // :509, :517-520 (threshold = 1)
      }
      prm.dev_rm[o] = devv;
      prm.z_rm[o] = zv;
      if (isfinite(zv)) welford_add(wf, zv); else nonfinite += 1.0;
      if (prm.obs) prm.obs[o] = (long long)acc[c];
      if (prm.samp)
        for (int i = 1; i <= B; ++i)
          prm.samp[((int64_t)k * B + (i - 1)) * prm.N + c0 + c] = (long long)acc[(size_t)i * W + c];
    }
    // variability partial of this (k, t): deterministic tree merge
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Welford b;
      b.n = __shfl_xor_sync(full, wf.n, o);
      b.mean = __shfl_xor_sync(full, wf.mean, o);
      b.m2 = __shfl_xor_sync(full, wf.m2, o);
      // merge in a lane-symmetric order so both partners get bit-identical results
      wf = (lane & o) ? welford_merge(b, wf) : welford_merge(wf, b);
      nonfinite += __shfl_xor_sync(full, nonfinite, o);
    }
    if (lane == 0) {
      s_w[warp][0] = wf.n; s_w[warp][1] = wf.mean; s_w[warp][2] = wf.m2; s_w[warp][3] = nonfinite;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      Welford tot = {0.0, 0.0, 0.0};
      double nf = 0.0;
      for (int w = 0; w < NW; ++w) {
        Welford b = {s_w[w][0], s_w[w][1], s_w[w][2]};
        tot = welford_merge(tot, b);
        nf += s_w[w][3];
      }
      double* vp = prm.varpart + ((int64_t)k * prm.T + t) * 4;
      vp[0] = tot.n; vp[1] = tot.mean; vp[2] = tot.m2; vp[3] = nf;
    }
    __syncthreads();
  }
  if (prm.upd_ctr) {
    unsigned long long v64 = warp_sum<unsigned long long>((unsigned long long)visited);
    if (lane == 0 && v64) atomicAdd(prm.upd_ctr, v64);
  }
}

// min(rowSums) > 0 check of compute_deviations_core (R/compute_deviations.R:362-363)
__global__ void k_check_zero_peaks(const unsigned long long* __restrict__ peak_tot, int P,
                                   uint32_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P && peak_tot[i] == 0ull) atomicOr(flags, 16u);
}

// varpart[K][T][4] -> varfinal[K][4], variab[K] = sqrt(M2 / (n - 1)) over finite entries
// (row_sds(z, na_rm = TRUE), src/utils.cpp:18-24: NaN when <= 1 finite value)
__global__ void k_var_merge(const double* __restrict__ varpart, int K, int T,
                            double* __restrict__ varfinal, double* __restrict__ variab) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  Welford tot = {0.0, 0.0, 0.0};
  double nf = 0.0;
  for (int t = 0; t < T; ++t) {
    const double* vp = varpart + ((int64_t)k * T + t) * 4;
    Welford b = {vp[0], vp[1], vp[2]};
    tot = welford_merge(tot, b);
    nf += vp[3];
  }
  varfinal[k * 4 + 0] = tot.n; varfinal[k * 4 + 1] = tot.mean;
  varfinal[k * 4 + 2] = tot.m2; varfinal[k * 4 + 3] = nf;
  variab[k] = tot.n > 1.0 ? sqrt(tot.m2 / (tot.n - 1.0)) : __longlong_as_double(0x7ff8000000000000LL);
}

// in: R rows x C columns, row stride ld_in -> out: C rows x R columns, row stride R
// (a block of columns of a [R][N] row-major matrix -> the same columns of the R x N column-major one)
template <typename T>
__global__ void __launch_bounds__(256)
k_transpose(const T* __restrict__ in, T* __restrict__ out, int64_t R, int64_t C, int64_t ld_in) {
  __shared__ T tile[32][33];
  int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (r0 + r < R && c0 + tx < C) tile[r][tx] = in[(r0 + r) * ld_in + c0 + tx];
  __syncthreads();
  for (int c = ty; c < 32; c += 8)
    if (c0 + c < C && r0 + tx < R) out[(c0 + c) * R + r0 + tx] = tile[tx][c];
}

// =========================================================================================
// host side
// =========================================================================================
int cv_choose_tile(cv_handle* h, int B, int acc_bytes, int* W_out) {
  const int I = B + 1;
  // static shared memory of k_contract: s_winv (8200 B) + s_w + s_item, rounded up
  const int budget = h->max_smem_optin - 10 * 1024;
  int W = budget / (I * acc_bytes);
  W = (W / 32) * 32;
  if (W > 4096) W = 4096;
  int64_t nround = ((h->N + 31) / 32) * 32;
  if (W > nround) W = (int)nround;
  if (W < 32)
    CV_FAIL(h, CV_ERR_UNSUPPORTED,
            "niterations = %d needs more shared memory than one SM has for a 32-cell tile", B);
  *W_out = W;
  return CV_OK;
}

// host-side validation + H2D of annotations / background / expectation.  `sync` = 0 leaves the
// copies in flight on `stream_for_copies` (the streamed call orders its kernels behind an event).
static int upload_annotations(cv_handle* h, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                              int index_base, const int32_t* bg, int B, const double* expectation,
                              cudaStream_t st) {
  h->have_anno = false;
  h->have_results = false;
  if (K < 1 || !annoptr) CV_FAIL(h, CV_ERR_INVALID, "Annotations are not valid");
  if (K > 0x7fffffff) CV_FAIL(h, CV_ERR_UNSUPPORTED, "K must be < 2^31");
  if (index_base != 0 && index_base != 1) CV_FAIL(h, CV_ERR_INVALID, "index_base must be 0 or 1");
  if (B < 1 || !bg) CV_FAIL(h, CV_ERR_INVALID, "background_peaks must have at least one column");
  if (B > 1024) CV_FAIL(h, CV_ERR_UNSUPPORTED, "niterations > 1024 is not supported");
  if (annoptr[0] != 0) CV_FAIL(h, CV_ERR_INVALID, "annoptr must start at 0");
  for (int64_t k = 0; k < K; ++k)
    if (annoptr[k + 1] < annoptr[k]) CV_FAIL(h, CV_ERR_INVALID, "annoptr is not monotone");
  const int64_t nnzA = annoptr[K];
  if (nnzA > 0 && !annoidx) CV_FAIL(h, CV_ERR_INVALID, "annoidx is NULL");
  const int64_t P = h->P;
  h->K = K; h->B = B; h->nnzA = nnzA; h->index_base = index_base;
  h->h_annoptr.assign(annoptr, annoptr + K + 1);
  CV_TRY(cv_ensure(h, h->annoptr, (size_t)(K + 1) * 8));
  CV_TRY(cv_ensure(h, h->anno_raw, (size_t)nnzA * 4));
  CV_TRY(cv_ensure(h, h->annoidx, (size_t)nnzA * 4));
  CV_TRY(cv_ensure(h, h->bg_raw, (size_t)P * B * 4));
  CV_TRY(cv_ensure(h, h->expect, (size_t)P * 8));
  CV_CUDA(h, cudaMemcpyAsync(h->annoptr.p, annoptr, (size_t)(K + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nnzA)
    CV_CUDA(h, cudaMemcpyAsync(h->anno_raw.p, annoidx, (size_t)nnzA * 4, cudaMemcpyHostToDevice, st));
  CV_CUDA(h, cudaMemcpyAsync(h->bg_raw.p, bg, (size_t)P * B * 4, cudaMemcpyHostToDevice, st));
  if (expectation)
    CV_CUDA(h, cudaMemcpyAsync(h->expect.p, expectation, (size_t)P * 8, cudaMemcpyHostToDevice, st));
  // annotations by decreasing set size (longest-processing-time-first work order)
  h->h_order.resize((size_t)K);
  std::iota(h->h_order.begin(), h->h_order.end(), 0);
  std::stable_sort(h->h_order.begin(), h->h_order.end(), [&](int32_t a, int32_t b) {
    return (annoptr[a + 1] - annoptr[a]) > (annoptr[b + 1] - annoptr[b]);
  });
  CV_TRY(cv_ensure(h, h->order, (size_t)K * 4));
  CV_CUDA(h, cudaMemcpyAsync(h->order.p, h->h_order.data(), (size_t)K * 4, cudaMemcpyHostToDevice, st));
  h->st.h2d_bytes = (int64_t)((K + 1) * 8 + nnzA * 4 + P * B * 4 + (expectation ? P * 8 : 0));
  return CV_OK;
}

extern "C" int cv_deviations_upload(cv_handle* h, int64_t K, const int64_t* annoptr,
                                    const int32_t* annoidx, int index_base, const int32_t* bg,
                                    int B, const double* expectation) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  h->replay_valid = false;
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  if (!expectation) CV_FAIL(h, CV_ERR_INVALID, "expectation is NULL");
  double t0 = wall_ms();
  CV_TRY(upload_annotations(h, K, annoptr, annoidx, index_base, bg, B, expectation, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->st.ms_h2d = wall_ms() - t0;
  h->have_anno = true;
  return CV_OK;
}

template <typename AccT>
static int launch_contract(cv_handle* h, const ContractParams& prm, size_t smem) {
  CV_CUDA(h, cudaFuncSetAttribute(k_contract<AccT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = prm.n_items < h->num_sms ? prm.n_items : h->num_sms;
  CV_LAUNCH(h, k_contract<AccT>, (unsigned)grid, CONTRACT_THREADS, smem, prm);
  return CV_OK;
}

// What a call streams between host and device while the kernels run (all optional).
struct DevIO {
  // counts arriving chunk by chunk (cv_deviations_csc): host CSC arrays, colptr as int64
  bool stream_counts = false;
  const CscView* view = nullptr;
  const int64_t* h_colptr = nullptr;
  int x_type = 0;
  size_t xb = 0;
  bool narrow = false;               // values are being narrowed to bytes on host threads (h->narrower -> h->stage8)
  bool stage_ri = false;             // ... and the row indices staged in page-locked memory alongside (h->stage_ri)
  bool land_out = false;             // results land in page-locked memory (h->stage_out) and are copied on by host threads
  int64_t x_bytes_sent = 0;          // bytes of values that actually crossed PCIe
  // results leaving chunk by chunk (K x N column-major: a cell chunk is a contiguous slab)
  double* out_dev = nullptr;
  double* out_z = nullptr;
  bool results_streamed = false;     // set by deviations_run when out_dev / out_z were written
  int soft_fail = 0;                 // set when the dense path turned out not to be exact
};

// Chunk [c0, c1) of the counts: H2D of its row indices and values, K1 on what arrived (validation,
// u32 values, per-peak / per-cell totals, fragments per sample), and a snapshot of the K1 flag
// words into pinned host memory -- all on the upload stream, so it runs ahead of the GEMMs.
static int enqueue_counts_upload(cv_handle* h, DevIO& io, int64_t c0, int64_t c1, int chunk_index,
                                 cudaEvent_t done) {
  const int64_t e0 = io.h_colptr[c0], e1 = io.h_colptr[c1];
  int x_type = io.x_type;
  if (e1 > e0) {
    bool narrowed = false;
    if (io.stage_ri) {
      // pageable caller memory: this chunk's slices (row indices copied, values narrowed) are waited
      // for first, then both travel from page-locked memory as asynchronous DMA
      narrowed = cv_narrower_wait(h->narrower, e0, e1) != 0;
      CV_CUDA(h, cudaMemcpyAsync(h->rowidx.as<int32_t>() + e0, h->stage_ri + e0, (size_t)(e1 - e0) * 4,
                                 cudaMemcpyHostToDevice, h->s_up));
    } else {
      CV_TRY(cv_csc_view_copy(h, *io.view, e0, e1, h->rowidx.as<int32_t>(), nullptr, h->s_up));
      // the row indices are on their way while the last slices of this chunk finish narrowing
      narrowed = io.narrow && cv_narrower_wait(h->narrower, e0, e1);
    }
    // xraw is staging only (K1 of this chunk reads it on the same stream before the next chunk's
    // copy lands), so a byte chunk and a chunk in the caller's type may share it.
    if (narrowed) {
      CV_CUDA(h, cudaMemcpyAsync(h->xraw.as<char>() + (size_t)e0, h->stage8 + e0, (size_t)(e1 - e0),
                                 cudaMemcpyHostToDevice, h->s_up));
      x_type = CV_U8;
      io.x_bytes_sent += e1 - e0;
    } else {
      CV_TRY(cv_csc_view_copy(h, *io.view, e0, e1, nullptr, h->xraw.as<char>(), h->s_up));
      io.x_bytes_sent += (e1 - e0) * (int64_t)io.xb;
    }
  }
  cudaStream_t compute = h->stream;
  h->stream = h->s_up;                                   // CV_LAUNCH targets h->stream
  int rc = cv_counts_range(h, x_type, e0, e1, c0, c1);
  h->stream = compute;
  CV_TRY(rc);
  CV_CUDA(h, cudaMemcpyAsync(h->pinned + 4 * chunk_index, h->flags.p, 16, cudaMemcpyDeviceToHost, h->s_up));
  CV_CUDA(h, cudaEventRecord(done, h->s_up));
  return CV_OK;
}

struct LandCopy { cv_narrower* nw; void* dst; const void* src; size_t n; };
static void CUDART_CB land_copy_cb(void* p) {
  LandCopy* c = reinterpret_cast<LandCopy*>(p);
  cv_narrower_copy(c->nw, c->dst, c->src, c->n);          // no CUDA call in here: only queues the pieces
  delete c;
}

static int enqueue_result_download(cv_handle* h, const DevIO& io, int64_t c0, int64_t c1, cudaEvent_t ready) {
  const int64_t K = h->K;
  CV_CUDA(h, cudaStreamWaitEvent(h->s_dn, ready, 0));
  if (io.land_out) {
    // pageable destination: DMA into the page-locked landing buffer, then the host threads copy the
    // slab on into the caller's matrix (in pieces, in parallel) as soon as it has landed
    const size_t slab = (size_t)(c1 - c0) * K * 8, mat = (size_t)h->N * K * 8;
    double* outs[2] = {io.out_dev, io.out_z};
    const double* srcs[2] = {h->dev_cm.as<double>(), h->z_cm.as<double>()};
    for (int m = 0; m < 2; ++m) {
      if (!outs[m]) continue;
      char* land = h->stage_out + (size_t)m * mat + (size_t)c0 * K * 8;
      CV_CUDA(h, cudaMemcpyAsync(land, srcs[m] + c0 * K, slab, cudaMemcpyDeviceToHost, h->s_dn));
      LandCopy* c = new LandCopy{h->narrower, outs[m] + c0 * K, land, slab};
      cudaError_t e = cudaLaunchHostFunc(h->s_dn, land_copy_cb, c);
      if (e != cudaSuccess) { delete c; CV_CUDA(h, e); }
    }
    return CV_OK;
  }
  if (io.out_dev)
    CV_CUDA(h, cudaMemcpyAsync(io.out_dev + c0 * K, h->dev_cm.as<double>() + c0 * K, (size_t)(c1 - c0) * K * 8,
                               cudaMemcpyDeviceToHost, h->s_dn));
  if (io.out_z)
    CV_CUDA(h, cudaMemcpyAsync(io.out_z + c0 * K, h->z_cm.as<double>() + c0 * K, (size_t)(c1 - c0) * K * 8,
                               cudaMemcpyDeviceToHost, h->s_dn));
  return CV_OK;
}

// [K][c0:c1) row-major -> the same cells of the K x N column-major ABI matrices
static int enqueue_transposes(cv_handle* h, int64_t c0, int64_t c1) {
  const int64_t K = h->K, N = h->N;
  dim3 grid((unsigned)((c1 - c0 + 31) / 32), (unsigned)((K + 31) / 32));
  CV_LAUNCH(h, k_transpose<double>, grid, 256, 0, h->dev_rm.as<double>() + c0, h->dev_cm.as<double>() + c0 * K, K,
            c1 - c0, N);
  CV_LAUNCH(h, k_transpose<double>, grid, 256, 0, h->z_rm.as<double>() + c0, h->z_cm.as<double>() + c0 * K, K,
            c1 - c0, N);
  return CV_OK;
}

static cudaEvent_t pool_event(cv_handle* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}

// ---- path selection: dense tcgen05 GEMM vs. cell-tiled row-gather SpMM -----------------------
// sparse cost: (B+1) * nnzA * (nnz / P) entry updates at the measured rate of k_contract;
// dense cost: the padded u8 GEMM (x digit planes) at a sustained tensor rate.
static int choose_path(cv_handle* h, uint32_t max_val, int64_t want_chunk, bool must_be_dense, DensePlan* pl,
                       int* use_dense) {
  *use_dense = 0;
  std::string why;
  const int feasible = cv_dense_plan(h, max_val, want_chunk, pl, &why);
  if (h->opt_path == 2 || must_be_dense) {
    if (!feasible) CV_FAIL(h, CV_ERR_UNSUPPORTED, "dense path requested but not applicable: %s", why.c_str());
    *use_dense = 1;
  } else if (h->opt_path == 0 && feasible) {
    // rate of the row-gather kernel in entry updates per ms: the last sparse run on this handle measured
    // it (deviations_run), 2.7e8 (config 2, round 1) until then
    const double updates = (double)(h->B + 1) * (double)h->nnzA * ((double)h->nnz / (double)h->P);
    const double sparse_ms = updates / h->sparse_rate + 0.3;
    *use_dense = cv_dense_cost_ms(h, *pl) < sparse_ms;
  }
  return CV_OK;
}

// Cell-chunk boundaries of the dense path (multiples of 256 cells, each <= pl.chunk_cells).
static std::vector<int64_t> chunk_bounds(cv_handle* h, const DensePlan& pl, int64_t N, bool streaming, bool host_io) {
  const int64_t chunk = pl.chunk_cells;
  std::vector<int64_t> b;
  const int64_t U = (N + 255) / 256, maxU = chunk / 256;
  const int n_min = (int)((U + maxU - 1) / maxU);
  // chunks exist to overlap host copies with the GEMMs; a call whose inputs and results stay on the
  // device takes as few as the operand budget allows (one launch ramp, one set of overflow columns)
  const int n = host_io ? std::max<int>(n_min, (int)std::min<int64_t>(3, U)) : n_min;
  if (!host_io && h->opt_chunk <= 0) {
    const int64_t per = ((U + n_min - 1) / n_min) * 256;           // equal chunks
    for (int64_t c = 0; c < N; c += per) b.push_back(c);
    b.push_back(N);
    return b;
  }
  if (h->opt_chunk > 0 || !pl.pair || n != 3 || U < 6 || U > 4096) {
    for (int64_t c = 0; c < N; c += chunk) b.push_back(c);
    b.push_back(N);
    return b;
  }
  const int64_t ipu = pl.fp4 ? pl.n_sg : pl.n_groups;            // work items per 256-cell unit
  const int64_t pairs = std::max(1, h->num_sms / 2);
  auto waves = [&](int64_t u) { return (u * ipu + pairs - 1) / pairs; };
  const int64_t minU = std::max<int64_t>(1, U / 6);
  int64_t best = -1, ba = 0, bb = 0;
  for (int64_t a = minU; a <= maxU; ++a)
    for (int64_t m = minU; m <= maxU; ++m) {
      const int64_t c = U - a - m;
      if (c < minU || c > maxU) continue;
      // counts arriving from the host: a chunk's upload has to fit behind the previous chunk's GEMM
      // (PCIe moves a unit in about two thirds of the time the GEMM needs for it)
      if (streaming && (2 * m > 3 * a + 4 || 2 * c > 3 * m + 4)) continue;
      const int64_t cost = (waves(a) + waves(m) + waves(c)) * 4096 + a * 8 + (m > c ? 1 : 0);   // fewest waves, then the smallest first chunk
      if (best < 0 || cost < best) { best = cost; ba = a; bb = m; }
    }
  if (best < 0) {
    for (int64_t c = 0; c < N; c += chunk) b.push_back(c);
    b.push_back(N);
    return b;
  }
  b = {0, ba * 256, (ba + bb) * 256, N};
  if (streaming && h->opt_tail_chunk && U >= 24) {
    // results leave for the host one chunk behind the GEMMs, so the copy of the LAST chunk's results
    // is exposed at the end of the call: a short fourth chunk (an eighth of the third, whole waves
    // kept where possible) shortens that tail by more than one more round of per-chunk launches costs
    const int64_t c3 = U - ba - bb;
    int64_t best_d = 0, best_cost = -1;
    for (int64_t d = std::max<int64_t>(1, c3 / 12); d <= std::max<int64_t>(1, c3 / 3); ++d) {
      const int64_t cost = (waves(c3 - d) + waves(d)) * 4096 + d * 1160;   // a wave ~ 0.23 ms, a unit of results ~ 0.065 ms on the wire
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_d = d; }
    }
    if (best_d > 0 && c3 - best_d >= 1) b = {0, ba * 256, (ba + bb) * 256, (U - best_d) * 256, N};
  }
  return b;
}

// Every kernel of the path on resident (or arriving) inputs.
static int deviations_run(cv_handle* h, int rebuild_counts_layout, int keep_sums, DevIO* io) {
  const int64_t P = h->P, N = h->N, K = h->K;
  const int B = h->B, I = B + 1;
  const bool streaming = io && io->stream_counts;
  h->have_results = false;
  if (!h->replaying) h->replay_valid = false;            // re-established at the end of a clean checked run
  cv_spans_reset(h);

  // accumulator width of the row-gather path: |sum| <= (largest set) * (largest count)
  int64_t set_max = 0;
  for (int64_t k = 0; k < K; ++k) set_max = std::max(set_max, h->h_annoptr[k + 1] - h->h_annoptr[k]);

  int use_dense = 0, dense_chunks_used = 0;
  bool use_gather = false;
  bool verdict_clean = false;
  uint32_t verdict_hf[16] = {};
  DensePlan pl;
  {
    // about three GEMM launches per call: host copies overlap them chunk by chunk (config 2, e2m1
    // kernel: e2e 20.2 / 19.6 / 22.4 ms with 4 / 3 / 2 chunks, kernels alone 17.5 / 17.0 / 16.6 ms;
    // profiles/r01p); buffers for up to half the cells, boundaries from chunk_bounds()
    const bool host_io_plan = streaming || (io && (io->out_dev || io->out_z));
    const int64_t want_chunk = host_io_plan ? std::max<int64_t>((N + 1) / 2, 1024) : 0;
    // few cells (bulk data): gathered counts rows instead of a stacked operand (cv_dense_gather.cu)
    use_gather = !streaming && !h->cols_dup && (h->replaying ? h->replay_gather : cv_gather_applicable(h) != 0);
    if (!use_gather) CV_TRY(choose_path(h, streaming ? 0u : h->max_val, want_chunk, streaming, &pl, &use_dense));
  }

  // ---- stage 1: validation and per-annotation tables ---------------------------------------------
  uint32_t* dflags = h->flags.as<uint32_t>() + CV_DFLAGS;
  cv_span_begin(h, CV_SPAN_ANNO);
  CV_CUDA(h, cudaMemsetAsync(dflags, 0, 32, h->stream));
  if (h->nnzA)
    CV_LAUNCH(h, k_anno_prepare, (unsigned)((h->nnzA + 255) / 256), 256, 0, h->anno_raw.as<int32_t>(),
              h->nnzA, (int)P, h->index_base, h->annoidx.as<int32_t>(), dflags);
  cv_span_end(h);
  CV_TRY(cv_ensure(h, h->etab, (size_t)K * I * 8));
  CV_TRY(cv_ensure(h, h->overlap, (size_t)K * 8));
  CV_TRY(cv_ensure(h, h->matches, (size_t)K * 8));
  CV_TRY(cv_ensure(h, h->work_ctr, 64));
  CV_TRY(cv_ensure(h, h->dev_rm, (size_t)K * N * 8));
  CV_TRY(cv_ensure(h, h->z_rm, (size_t)K * N * 8));
  CV_TRY(cv_ensure(h, h->dev_cm, (size_t)K * N * 8));
  CV_TRY(cv_ensure(h, h->z_cm, (size_t)K * N * 8));
  CV_TRY(cv_ensure(h, h->varfinal, (size_t)K * 4 * 8));
  CV_TRY(cv_ensure(h, h->variab, (size_t)K * 8));
  if (keep_sums) {
    CV_TRY(cv_ensure(h, h->obs, (size_t)K * N * 8));
    CV_TRY(cv_ensure(h, h->samp, (size_t)K * B * N * 8));
  }
  h->kept_sums = keep_sums != 0;
  CV_CUDA(h, cudaMemsetAsync(h->work_ctr.p, 0, 64, h->stream));

  // ---- stage 2: operand layout + contraction + fused epilogue -----------------------------------
  int T = 0;
  bool wide = false;
  if (use_gather) {
    DensePlan plg;
    plg.I = I;
    plg.G = 1;
    // bit columns + per-annotation tables (auxiliary stream; the expectation table rides on the GEMM when
    // the cell tile has spare columns for its digits), then the operands
    CV_TRY(cv_dense_build_rows(h, plg, 0, cv_gather_etab_in_gemm(h) ? 2 : 1));
    CV_TRY(cv_gather_prepare(h));
    uint32_t hf[16] = {};
    if (!h->replaying) {
      CV_CUDA(h, cudaMemcpyAsync(hf, h->flags.p, 64, cudaMemcpyDeviceToHost, h->stream));
      CV_CUDA(h, cudaStreamSynchronize(h->stream));
      if (hf[CV_DFLAGS] & FLAG_BADIDX)
        CV_FAIL(h, CV_ERR_INVALID, "Annotations are not valid / background_peaks index outside 1..nrow(counts)");
      if (hf[CV_DFLAGS + 3] || (hf[CV_DFLAGS + 4] & 2u)) {
        // a list repeats a peak (bit columns cannot hold the multiplicity), or an expectation outside the
        // fixed-point range of the digit columns: the general paths take over
        CV_CUDA(h, cudaStreamSynchronize(h->s_aux));
        h->aux_pending = false;
        CV_CUDA(h, cudaMemsetAsync(dflags + 1, 0, 12, h->stream));
        use_gather = false;
        const int64_t want_chunk = (io && (io->out_dev || io->out_z)) ? std::max<int64_t>((N + 1) / 2, 1024) : 0;
        CV_TRY(choose_path(h, h->max_val, want_chunk, false, &pl, &use_dense));
      }
    }
    if (use_gather) {
      CV_TRY(cv_gather_run(h, keep_sums));
      cv_span_begin(h, CV_SPAN_FINALIZE);
      CV_TRY(enqueue_transposes(h, 0, N));
      cv_span_end(h);
      T = (int)((N + 127) / 128);
      dense_chunks_used = 1;
      verdict_clean = !keep_sums;
    }
  }
  if (use_dense) {
    const int64_t chunk = pl.chunk_cells;
    // chunk boundaries: uniform when the caller fixed the chunk ("cell_chunk"); otherwise about three
    // chunks whose work items fill whole waves of CTA pairs (config 2: 40 units of 256 cells x 97
    // super-groups on 74 pairs: 8 + 16 + 16 units = 53 waves, 14 + 13 + 13 = 55), the small one first
    std::vector<int64_t> bounds = chunk_bounds(h, pl, N, streaming, streaming || (io && (io->out_dev || io->out_z)));
    const int n_chunks = (int)bounds.size() - 1;
    dense_chunks_used = n_chunks;
    std::vector<cudaEvent_t> ev_up((size_t)n_chunks), ev_done((size_t)n_chunks), ev_dn((size_t)n_chunks);
    for (int i = 0; i < n_chunks; ++i) { ev_up[i] = pool_event(h); ev_done[i] = pool_event(h); ev_dn[i] = pool_event(h); }
    const bool trace = h->opt_trace != 0;                 // CV_TRACE, read once in cv_create
    cudaEvent_t ev_t0 = pool_event(h), ev_prep = pool_event(h);
    CV_CUDA(h, cudaEventRecord(ev_t0, h->stream));         // origin of the timeline; orders the upload stream
    auto lo = [&](int i) { return bounds[(size_t)i]; };
    auto hi = [&](int i) { return bounds[(size_t)i + 1]; };
    CV_TRY(cv_dense_prepare(h, pl, 0));                       // operand M + tables: cell independent
    if (trace) cudaEventRecord(ev_prep, h->stream);
    if (streaming) {
      if (n_chunks > CV_PINNED_WORDS / 4) CV_FAIL(h, CV_ERR_UNSUPPORTED, "too many cell chunks (%d)", n_chunks);
      // the zeroing of the totals / flags (cv_counts_begin, compute stream) precedes the first K1
      CV_CUDA(h, cudaStreamWaitEvent(h->s_up, ev_t0, 0));
      CV_TRY(enqueue_counts_upload(h, *io, lo(0), hi(0), 0, ev_up[0]));
    }
    // multiplicity bounds of the u8 operand before any GEMM runs (one short host sync; the first
    // counts chunk is already in flight)
    uint32_t hf[16];
    if (h->replaying) {
      memcpy(hf, h->replay_hf, 64);                          // same inputs, same kernels: same words
    } else {
      CV_CUDA(h, cudaMemcpyAsync(hf, h->flags.p, 64, cudaMemcpyDeviceToHost, h->stream));
      CV_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    bool clean = !streaming && !keep_sums && hf[CV_DFLAGS + 3] == 0;   // may this call's verdict be replayed?
    memcpy(verdict_hf, hf, 64);
    if (hf[CV_DFLAGS] & FLAG_BADIDX)
      CV_FAIL(h, CV_ERR_INVALID, "Annotations are not valid / background_peaks index outside 1..nrow(counts)");
    if (hf[CV_DFLAGS + 3]) {
      // a list repeats a peak: bit columns cannot hold the multiplicity, rebuild the rows by scatter
      CV_CUDA(h, cudaStreamSynchronize(h->s_aux));           // the first attempt's table kernels
      h->aux_pending = false;
      CV_CUDA(h, cudaMemsetAsync(dflags + 1, 0, 12, h->stream));
      cv_dense_plan_set_fp4(&pl, 0);                         // the e2m1 rows exist only in the gather formulation
      CV_TRY(cv_dense_prepare(h, pl, 1));
      CV_CUDA(h, cudaMemcpyAsync(hf, h->flags.p, 64, cudaMemcpyDeviceToHost, h->stream));
      CV_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    int soft = 0;
    int rc = cv_dense_check(h, pl, hf, streaming ? 0 : 1, &soft);
    if (rc != CV_OK) {
      if (soft && h->opt_path != 2 && !streaming) use_dense = 0;     // exact fallback below
      else { if (io) io->soft_fail = soft; return rc; }
    }
    if (use_dense && pl.fp4 &&
        ((!streaming && !((double)h->max_cell_tot * (double)h->dn_maxmult < 16777216.0)) ||
         (double)hf[CV_DFLAGS + 7] > 0.7 * (double)pl.Hcap || (hf[CV_DFLAGS + 5] & 4u))) {
      // sums may not fit a float exactly, or the operand alone (multiplicities outside the e2m1
      // alphabet on most peaks: dense k-mer annotations) already fills the overflow reserve, or an
      // entry passed 15: u8 operands / s32 accumulators instead.  Only the operand rows are redone
      // (bit columns, inverse background map and the per-annotation tables do not depend on the format).
      clean = false;
      CV_CUDA(h, cudaMemsetAsync(dflags + 5, 0, 12, h->stream));
      cv_dense_plan_set_fp4(&pl, 0);
      CV_TRY(cv_ensure(h, h->dn_M, (size_t)pl.m_bytes));
      CV_TRY(cv_ensure(h, h->dn_C, (size_t)pl.c_bytes));
      cv_span_begin(h, CV_SPAN_LAYOUT);
      CV_TRY(cv_dense_rebuild_rows_u8(h, pl));
      cv_span_end(h);
    }
    verdict_clean = clean && use_dense;
    if (use_dense) {
      const bool dl = io && (io->out_dev || io->out_z);
      int planes_max = pl.planes;
      for (int i = 0; i < n_chunks; ++i) {
        DensePlan pli = pl;
        pli.cell_tot_bound = h->max_cell_tot;              // resident counts: known since K1
        if (streaming) {
          // K1 of this chunk has run on the upload stream: its verdict and the largest count so far
          // decide how many u8 digit planes this chunk's GEMM needs
          CV_CUDA(h, cudaEventSynchronize(ev_up[i]));
          const uint32_t* kf = h->pinned + 4 * i;
          if (kf[0]) return cv_counts_verdict(h, kf);            // counts_check failed: message from K1
          pli.planes = 1;
          for (uint32_t v = kf[1] >> 8; v; v >>= 8) pli.planes++;
          pli.cell_tot_bound = (uint64_t)kf[2] | ((uint64_t)kf[3] << 32);   // running maximum incl. this chunk
          planes_max = std::max(planes_max, pli.planes);
          CV_CUDA(h, cudaStreamWaitEvent(h->stream, ev_up[i], 0));
        }
        {
          const int rcc = cv_dense_chunk(h, pli, lo(i), hi(i), keep_sums);
          if (rcc != CV_OK) {
            if (h->f4_veto && !streaming && i == 0) {         // the first chunk does not fit the e2m1 operands: u8 from the start
              const int rc2 = deviations_run(h, rebuild_counts_layout, keep_sums, io);
              h->f4_veto = false;
              return rc2;
            }
            if (h->f4_veto && io) io->soft_fail = 1;         // the caller retries with u8
            return rcc;
          }
        }
        cv_span_begin(h, CV_SPAN_FINALIZE);
        CV_TRY(enqueue_transposes(h, lo(i), hi(i)));
        cv_span_end(h);
        CV_CUDA(h, cudaEventRecord(ev_done[i], h->stream));
        // host copies: next chunk in, previous chunk out (one behind, so that a blocking copy of
        // pageable memory still leaves a queued chunk for the GPU)
        if (streaming && i + 1 < n_chunks)
          CV_TRY(enqueue_counts_upload(h, *io, lo(i + 1), hi(i + 1), i + 1, ev_up[i + 1]));
        if (dl && i > 0) {
          CV_TRY(enqueue_result_download(h, *io, lo(i - 1), hi(i - 1), ev_done[i - 1]));
          if (trace) cudaEventRecord(ev_dn[i - 1], h->s_dn);
        }
      }
      pl.planes = planes_max;
      if (dl) {
        CV_TRY(enqueue_result_download(h, *io, lo(n_chunks - 1), hi(n_chunks - 1), ev_done[n_chunks - 1]));
        if (trace) cudaEventRecord(ev_dn[n_chunks - 1], h->s_dn);
        io->results_streamed = true;
      }
      if (trace) {
        // timeline of the call (ms since its first kernel): CV_TRACE=1
        cudaStreamSynchronize(h->stream);
        cudaStreamSynchronize(h->s_dn);
        float t = 0.f;
        cudaEventElapsedTime(&t, ev_t0, ev_prep);
        fprintf(stderr, "[cv trace] operand M ready %.2f ms; chunks of %lld cells\n", t, (long long)chunk);
        for (int i = 0; i < n_chunks; ++i) {
          float a = -1.f, b = -1.f, c = -1.f;
          if (streaming) cudaEventElapsedTime(&a, ev_t0, ev_up[i]);
          cudaEventElapsedTime(&b, ev_t0, ev_done[i]);
          if (dl) cudaEventElapsedTime(&c, ev_t0, ev_dn[i]);
          fprintf(stderr, "[cv trace] chunk %d: counts on device %.2f, computed %.2f, results on host %.2f\n", i, a, b, c);
        }
        cudaGetLastError();
      }
      T = (int)((N + 127) / 128);
    }
  }
  if (!use_dense && !use_gather) {
    if (h->aux_pending) {                                    // a dense attempt was abandoned
      CV_CUDA(h, cudaStreamSynchronize(h->s_aux));
      h->aux_pending = false;
    }
    wide = (double)set_max * (double)h->max_val >= 2147483647.0;
    const int acc_bytes = wide ? 8 : 4;
    // background matrix -> [P][B] 0-based, expected-fraction table, overlap, matches
    cv_span_begin(h, CV_SPAN_ANNO);
    CV_TRY(cv_ensure(h, h->bgT, (size_t)P * B * 4));
    {
      dim3 grid((unsigned)((P + 31) / 32), (unsigned)((B + 31) / 32));
      CV_LAUNCH(h, k_bg_prepare, grid, 256, 0, h->bg_raw.as<int32_t>(), (int)P, B, h->index_base,
                h->bgT.as<int32_t>(), dflags);
    }
    {
      int grid = (int)std::min<int64_t>(K, (int64_t)h->num_sms * 2);
      size_t mark_bytes = (size_t)grid * P * 4;
      if (h->mark.cap < mark_bytes) {
        CV_TRY(cv_ensure(h, h->mark, mark_bytes));
        CV_CUDA(h, cudaMemsetAsync(h->mark.p, 0, h->mark.cap, h->stream));
      }
      size_t smem = (size_t)(STATS_THREADS / 32) * I * 8;
      CV_CUDA(h, cudaFuncSetAttribute(k_anno_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CV_LAUNCH(h, k_anno_stats, grid, STATS_THREADS, smem, h->annoptr.as<int64_t>(),
                h->annoidx.as<int32_t>(), h->bgT.as<int32_t>(), h->expect.as<double>(), (int)P, B,
                (int)K, h->mark.as<int32_t>(), h->etab.as<double>(), h->overlap.as<double>(),
                h->matches.as<double>());
    }
    cv_span_end(h);
    CV_CUDA(h, cudaMemsetAsync(h->work_ctr.p, 0, 64, h->stream));
    int W = 0;
    CV_TRY(cv_choose_tile(h, B, acc_bytes, &W));
    cv_span_begin(h, CV_SPAN_LAYOUT);
    if (rebuild_counts_layout || !h->layout_valid || h->W != W) CV_TRY(cv_build_layout(h, W));
    cv_span_end(h);
    T = h->T;
    CV_TRY(cv_ensure(h, h->varpart, (size_t)K * T * 4 * 8));
    ContractParams prm;
    prm.P = (int)P; prm.N = (int)N; prm.K = (int)K; prm.B = B; prm.W = W; prm.T = T;
    prm.n_items = (int64_t)T * K;
    prm.annoptr = h->annoptr.as<int64_t>();
    prm.annoidx = h->annoidx.as<int32_t>();
    prm.bgT = h->bgT.as<int32_t>();
    prm.rowptr = h->rowptr.as<uint32_t>();
    prm.entries = h->entries.as<uint32_t>();
    prm.etab = h->etab.as<double>();
    prm.fps = h->fps.as<double>();
    prm.order = h->order.as<int32_t>();
    prm.work_ctr = h->work_ctr.as<uint32_t>();
    prm.dev_rm = h->dev_rm.as<double>();
    prm.z_rm = h->z_rm.as<double>();
    prm.varpart = h->varpart.as<double>();
    prm.obs = keep_sums ? h->obs.as<long long>() : nullptr;
    prm.samp = keep_sums ? h->samp.as<long long>() : nullptr;
    prm.upd_ctr = reinterpret_cast<unsigned long long*>(h->work_ctr.as<uint32_t>() + 2);
    const size_t smem = (size_t)I * W * acc_bytes;
    cv_span_begin(h, CV_SPAN_CONTRACT);
    if (wide) CV_TRY(launch_contract<long long>(h, prm, smem));
    else CV_TRY(launch_contract<int>(h, prm, smem));
    cv_span_end(h);
    cv_span_begin(h, CV_SPAN_FINALIZE);
    CV_TRY(enqueue_transposes(h, 0, N));
    cv_span_end(h);
  }

  // ---- stage 3: totals, zero-peak check, variability merge -----------------------------------------
  if (h->aux_pending) {                                  // overlap / matches from the auxiliary stream
    CV_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_aux[3], 0));
    h->aux_pending = false;
  }
  cv_span_begin(h, CV_SPAN_FINALIZE);
  if (streaming) CV_TRY(cv_counts_finish_async(h));       // every K1 chunk was waited for above
  if (h->opt_check_zero)
    CV_LAUNCH(h, k_check_zero_peaks, (unsigned)((P + 255) / 256), 256, 0,
              h->peak_tot.as<unsigned long long>(), (int)P, dflags);
  CV_LAUNCH(h, k_var_merge, (unsigned)((K + 127) / 128), 128, 0, h->varpart.as<double>(), (int)K, T,
            h->varfinal.as<double>(), h->variab.as<double>());
  cv_span_end(h);
  uint32_t hf[16];
  unsigned long long upd = 0;
  if (h->replaying) {
    // no host in the loop: the flag words travel to page-locked memory behind the kernels and
    // cv_settle() looks at them (and at the timed spans) once somebody needs the outcome
    CV_CUDA(h, cudaMemcpyAsync(h->pinned + CV_PINNED_WORDS - 32, h->flags.p, 64, cudaMemcpyDeviceToHost, h->stream));
    h->st.path = 1;
    h->st.n_cell_tiles = dense_chunks_used;
    h->have_results = true;
    h->async_pending = true;
    return CV_OK;
  }
  CV_CUDA(h, cudaMemcpyAsync(hf, h->flags.p, 64, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaMemcpyAsync(&upd, h->work_ctr.as<uint32_t>() + 2, 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->s_dn));
  if (streaming) {
    CV_TRY(cv_counts_verdict(h, hf));                 // counts_check on what arrived
    int soft = 0;
    int rc = cv_dense_check(h, pl, hf, 1, &soft);     // planes / accumulator bound, now that K1 is complete
    if (rc != CV_OK) { io->soft_fail = soft; io->results_streamed = false; return rc; }
  }
  if (hf[CV_DFLAGS] & FLAG_ZEROPEAK)
    CV_FAIL(h, CV_ERR_INVALID, "All peaks must have at least one fragment in one sample");
  if (hf[CV_DFLAGS] & FLAG_BADIDX)
    CV_FAIL(h, CV_ERR_INVALID, "Annotations are not valid / background_peaks index outside 1..nrow(counts)");
  if (use_dense && pl.fp4 && hf[CV_DFLAGS + 5]) {
    // the e2m1 operands could not hold this problem (overflow columns beyond the reserve, or a
    // column listing a peak twice): the results are void, redo with the u8 kernel
    h->f4_veto = true;
    if (streaming) {
      io->soft_fail = 1;
      io->results_streamed = false;
      CV_FAIL(h, CV_ERR_UNSUPPORTED, "counts not representable on the e2m1 path (flags %u)", hf[CV_DFLAGS + 5]);
    }
    const int rc2 = deviations_run(h, rebuild_counts_layout, keep_sums, io);
    h->f4_veto = false;
    return rc2;
  }
  double ms[CV_SPAN_N];
  cv_spans_collect(h, ms);
  h->st.ms_counts_layout = ms[CV_SPAN_LAYOUT] + ms[CV_SPAN_K1];
  h->st.ms_anno_prep = ms[CV_SPAN_ANNO];
  h->st.ms_contract = ms[CV_SPAN_CONTRACT];
  h->st.ms_finalize = ms[CV_SPAN_FINALIZE];
  h->st.contract_updates = (int64_t)upd;
  h->st.few_cells = use_gather ? 1 : 0;
  if (use_gather) {
    h->st.contract_bytes = 0;
    h->st.contract_flops = h->dn_flops;
    h->st.path = 1;
    h->st.acc_bytes = 4;
    h->st.cell_tile = 256;
    h->st.n_cell_tiles = 1;
    h->st.planes = 2;
    h->st.epilogue = 64;
    h->st.operand_bits = 8;
    h->st.overflow_columns = 0;
  } else if (use_dense) {
    h->st.contract_bytes = 0;
    h->st.contract_flops = h->dn_flops;
    if (pl.fp4)                                            // + the overflow columns of every chunk
      h->st.contract_flops += 2.0 * (double)pl.rowsM4 * (double)std::min<int64_t>(pl.chunk_cells, N) * (double)hf[CV_DFLAGS + 6];
    h->st.path = 1;
    h->st.acc_bytes = 4;
    h->st.cell_tile = pl.pair ? 256 : 128;
    h->st.n_cell_tiles = dense_chunks_used;
    h->st.planes = pl.planes;
    h->st.operand_bits = pl.fp4 ? 4 : 8;
    h->st.overflow_columns = pl.fp4 ? (int32_t)hf[CV_DFLAGS + 6] : 0;
  } else {
    // gather model: every visited entry is a 4-byte load, plus the
    // compulsory one-time traffic (layout, annotation lists, background rows, outputs)
    h->st.contract_bytes = (int64_t)upd * 4 + h->n_entries * 4 + (int64_t)T * P * 4 +
                           h->nnzA * 4 * T + (int64_t)P * B * 4 + 2 * K * N * 8;
    if (upd > 10000000ull && ms[CV_SPAN_CONTRACT] > 0.05)      // feeds the cost model of later calls
      h->sparse_rate = (double)upd / ms[CV_SPAN_CONTRACT];
    h->st.contract_flops = 0;
    h->st.path = 0;
    h->st.acc_bytes = wide ? 8 : 4;
    h->st.planes = 0;
    h->st.epilogue = 64;
    h->st.operand_bits = 0;
    h->st.overflow_columns = 0;
  }
  h->have_results = true;
  // a clean dense run on resident inputs (no fallback taken, nothing flagged): its verdict can be replayed
  // (not under an e2m1 veto either: the veto ends with this call, a replay would plan e2m1 operands again)
  h->replay_valid = verdict_clean && (use_dense || use_gather) && !h->f4_veto && !hf[CV_DFLAGS + 5] &&
                    !(hf[CV_DFLAGS] & (FLAG_ZEROPEAK | FLAG_BADIDX));
  h->replay_gather = use_gather;
  if (h->replay_valid) memcpy(h->replay_hf, verdict_hf, 64);
  return CV_OK;
}

// Waits for a pending replay (cv_deviations_compute_async) and collects what the checked call
// collects inline: the flag words and the timed spans.
int cv_settle(cv_handle* h) {
  if (!h->async_pending) return CV_OK;
  h->async_pending = false;
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  const uint32_t* hf = h->pinned + CV_PINNED_WORDS - 32;
  if ((hf[CV_DFLAGS] & (FLAG_ZEROPEAK | FLAG_BADIDX)) || hf[CV_DFLAGS + 5] || hf[CV_DFLAGS + 3]) {
    h->have_results = false;
    h->replay_valid = false;
    CV_FAIL(h, CV_ERR_STATE, "replayed call raised flags the checked call did not (resident inputs changed?)");
  }
  double ms[CV_SPAN_N];
  cv_spans_collect(h, ms);
  h->st.ms_counts_layout = ms[CV_SPAN_LAYOUT] + ms[CV_SPAN_K1];
  h->st.ms_anno_prep = ms[CV_SPAN_ANNO];
  h->st.ms_contract = ms[CV_SPAN_CONTRACT];
  h->st.ms_finalize = ms[CV_SPAN_FINALIZE];
  return CV_OK;
}

extern "C" int cv_deviations_wait(cv_handle* h) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  return cv_settle(h);
}

// cv_deviations_compute with the host out of the loop.  When the last fully checked call ran on the
// same resident inputs and options and took no fallback, every kernel of the path is enqueued again
// -- same work, same results -- but the host neither waits for the flag words between the stages
// (their values are known) nor for the end of the call, so consecutive calls queue up behind each
// other and a descheduled host thread no longer leaves the GPU idle.  Otherwise it is the checked
// call itself.  cv_deviations_wait (or any call that reads results or changes inputs) settles it.
extern "C" int cv_deviations_compute_async(cv_handle* h, int rebuild_counts_layout) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (!h->replay_valid || h->opt_path == 1) return cv_deviations_compute(h, rebuild_counts_layout, 0);
  if (!h->have_counts || !h->have_anno) CV_FAIL(h, CV_ERR_STATE, "counts / annotations not set");
  h->replaying = true;
  const int rc = deviations_run(h, rebuild_counts_layout, 0, nullptr);
  h->replaying = false;
  if (rc != CV_OK) { h->replay_valid = false; h->async_pending = false; }
  return rc;
}

extern "C" int cv_deviations_compute(cv_handle* h, int rebuild_counts_layout, int keep_sums) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  if (!h->have_anno) CV_FAIL(h, CV_ERR_STATE, "annotations / background not uploaded");
  if (!h->totals_committed)
    CV_FAIL(h, CV_ERR_STATE, "per-peak totals handed out for a cross-shard reduction; call cv_peak_totals_commit first");
  return deviations_run(h, rebuild_counts_layout, keep_sums, nullptr);
}

static int download_results(cv_handle* h, double* out_dev, double* out_z, double* out_matches,
                            double* out_overlap, double* out_variability, int64_t* out_observed,
                            int64_t* out_sampled) {
  if (!h->have_results) CV_FAIL(h, CV_ERR_STATE, "no results: call cv_deviations_compute first");
  const int64_t N = h->N, K = h->K;
  const int B = h->B;
  double t0 = wall_ms();
  int64_t bytes = 0;
  if (out_dev) { CV_CUDA(h, cudaMemcpyAsync(out_dev, h->dev_cm.p, (size_t)K * N * 8, cudaMemcpyDeviceToHost, h->stream)); bytes += K * N * 8; }
  if (out_z) { CV_CUDA(h, cudaMemcpyAsync(out_z, h->z_cm.p, (size_t)K * N * 8, cudaMemcpyDeviceToHost, h->stream)); bytes += K * N * 8; }
  if (out_matches) { CV_CUDA(h, cudaMemcpyAsync(out_matches, h->matches.p, (size_t)K * 8, cudaMemcpyDeviceToHost, h->stream)); bytes += K * 8; }
  if (out_overlap) { CV_CUDA(h, cudaMemcpyAsync(out_overlap, h->overlap.p, (size_t)K * 8, cudaMemcpyDeviceToHost, h->stream)); bytes += K * 8; }
  if (out_variability) { CV_CUDA(h, cudaMemcpyAsync(out_variability, h->variab.p, (size_t)K * 8, cudaMemcpyDeviceToHost, h->stream)); bytes += K * 8; }
  if (out_observed || out_sampled) {
    if (!h->kept_sums) CV_FAIL(h, CV_ERR_STATE, "integer sums were not kept (keep_sums = 0)");
    if (out_observed) {
      // device holds [K][N]; ABI wants K x N column-major
      CV_TRY(cv_ensure(h, h->tmp1, (size_t)K * N * 8));
      dim3 grid((unsigned)((N + 31) / 32), (unsigned)((K + 31) / 32));
      CV_LAUNCH(h, k_transpose<long long>, grid, 256, 0, h->obs.as<long long>(), h->tmp1.as<long long>(), K, N, N);
      CV_CUDA(h, cudaMemcpyAsync(out_observed, h->tmp1.p, (size_t)K * N * 8, cudaMemcpyDeviceToHost, h->stream));
      bytes += K * N * 8;
    }
    if (out_sampled) {
      CV_CUDA(h, cudaMemcpyAsync(out_sampled, h->samp.p, (size_t)K * B * N * 8, cudaMemcpyDeviceToHost, h->stream));
      bytes += K * B * N * 8;
    }
  }
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->st.ms_d2h = wall_ms() - t0;
  h->st.d2h_bytes = bytes;
  return CV_OK;
}

extern "C" int cv_deviations_download(cv_handle* h, double* out_dev, double* out_z,
                                      double* out_matches, double* out_overlap,
                                      double* out_variability, int64_t* out_observed,
                                      int64_t* out_sampled) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  return download_results(h, out_dev, out_z, out_matches, out_overlap, out_variability, out_observed, out_sampled);
}

// counts resident: H2D of the annotations, then the kernels with the K x N results leaving chunk
// by chunk while later chunks are still being computed
extern "C" int cv_deviations(cv_handle* h, int64_t K, const int64_t* annoptr,
                             const int32_t* annoidx, int index_base, const int32_t* bg, int B,
                             const double* expectation, double* out_dev, double* out_z,
                             double* out_matches, double* out_overlap, double* out_variability,
                             int64_t* out_observed, int64_t* out_sampled) {
  if (!h) return CV_ERR_INVALID;
  CV_TRY(cv_deviations_upload(h, K, annoptr, annoidx, index_base, bg, B, expectation));
  double h2d_ms = h->st.ms_h2d;
  int64_t h2d_b = h->st.h2d_bytes;
  if (!h->totals_committed)
    CV_FAIL(h, CV_ERR_STATE, "per-peak totals handed out for a cross-shard reduction; call cv_peak_totals_commit first");
  DevIO io;
  io.out_dev = out_dev;
  io.out_z = out_z;
  CV_TRY(deviations_run(h, 0, (out_observed || out_sampled) ? 1 : 0, &io));
  CV_TRY(download_results(h, io.results_streamed ? nullptr : out_dev, io.results_streamed ? nullptr : out_z,
                          out_matches, out_overlap, out_variability, out_observed, out_sampled));
  if (io.results_streamed) h->st.d2h_bytes += 2 * K * h->N * 8;
  h->st.ms_h2d = h2d_ms;
  h->st.h2d_bytes = h2d_b;
  return CV_OK;
}

// compute_deviations_core (R/compute_deviations.R:355-408) as ONE call on host data: counts
// (CSC as in dgCMatrix), annotations, background, expectation in; deviations / z out.  The counts
// travel in cell chunks on their own stream while earlier chunks are in the GEMM, and finished
// chunks of the K x N results travel back while later chunks are computed; afterwards the counts
// are resident exactly as after cv_set_counts_csc.
extern "C" int cv_deviations_csc(cv_handle* h, int64_t P, int64_t N, const void* colptr, int colptr_type,
                                 const int32_t* rowidx, const void* x, int x_type, int64_t K,
                                 const int64_t* annoptr, const int32_t* annoidx, int index_base,
                                 const int32_t* bg, int B, const double* expectation, double* out_dev,
                                 double* out_z, double* out_matches, double* out_overlap,
                                 double* out_variability) {
  if (!h) return CV_ERR_INVALID;
  CscView v;
  CV_TRY(cv_csc_view_init(h, &v, P, N, colptr, colptr_type, rowidx, x, x_type));
  return cv_deviations_view(h, v, K, annoptr, annoidx, index_base, bg, B, expectation, out_dev, out_z, out_matches,
                            out_overlap, out_variability);
}

int cv_deviations_view(cv_handle* h, const CscView& v, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                       int index_base, const int32_t* bg, int B, const double* expectation, double* out_dev,
                       double* out_z, double* out_matches, double* out_overlap, double* out_variability) {
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  h->replay_valid = false;
  const int64_t P = v.P, N = v.N;
  const int x_type = v.x_type;
  // the streamed pipeline needs the expectation up front and the dense path; otherwise (or when
  // the dense path turns out not to be exact for these inputs) the call is the plain sequence
  // (few cells: the plain sequence, whose resident call takes the gathered-rows path of cv_dense_gather.cu)
  bool try_stream = expectation != nullptr && h->opt_path != 1 && K >= 1 && annoptr && B >= 1 && B <= 255 &&
                    !(N <= 256 && h->opt_gather != 0 && (double)K * (B + 1) * (double)P > 1.0e9);
  int64_t nnz = 0;
  if (try_stream) {
    nnz = v.cp[(size_t)N];
    bool ok = cv_csc_view_check(h, v) == CV_OK && nnz > 0 && annoptr[0] == 0 && annoptr[K] >= 0;
    if (ok) {
      // same cost model as the resident call, on the host-known shape (largest count unknown yet:
      // one u8 plane assumed, verified once the counts have passed K1)
      h->have_counts = false;
      h->have_anno = false;
      h->P = P; h->N = N; h->nnz = nnz; h->K = K; h->B = B; h->nnzA = annoptr[K];
      DensePlan pl;
      int use_dense = 0;
      const int opt = h->opt_path;
      int rc = choose_path(h, 0u, std::max<int64_t>((N + 1) / 2, 1024), false, &pl, &use_dense);
      ok = rc == CV_OK && use_dense && (opt == 2 || cv_dense_cost_ms(h, pl) > 2.0);   // tiny problems: plain sequence
    }
    try_stream = ok;
  }
  if (try_stream) {
    const size_t xb = v.xb;
    {
      double t0 = wall_ms();
      h->have_counts = false;
      h->P = P; h->N = N; h->nnz = nnz;
      CV_TRY(cv_ensure(h, h->colptr, (size_t)(N + 1) * 8));
      CV_TRY(cv_ensure(h, h->rowidx, (size_t)nnz * 4));
      CV_TRY(cv_ensure(h, h->xraw, (size_t)nnz * xb));
      CV_TRY(cv_counts_begin(h));
      // small inputs first, on the compute stream (the operand build needs them immediately)
      CV_CUDA(h, cudaMemcpyAsync(h->colptr.p, v.cp.data(), (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, h->stream));
      int rc = upload_annotations(h, K, annoptr, annoidx, index_base, bg, B, expectation, h->stream);
      if (rc != CV_OK) return rc;
      int64_t h2d_b = h->st.h2d_bytes + (N + 1) * 8 + nnz * 4;
      h->have_anno = true;
      DevIO io;
      io.stream_counts = true;
      io.view = &v;
      io.h_colptr = v.cp.data();
      io.x_type = x_type;
      io.xb = xb;
      io.out_dev = out_dev;
      io.out_z = out_z;
      // wide value types (R: double) are narrowed to bytes by host threads while the annotations
      // travel and the operand is built; each chunk's upload waits for its own slices only
      struct NarrowGuard {                      // the caller's x must not be read after the call returns
        cv_handle* h;
        ~NarrowGuard() { if (h->narrower) cv_narrower_finish(h->narrower); }
      } narrow_guard{h};
      // host threads between the caller's memory and the DMA engines ("host_narrow"): always when asked
      // for (n >= 1), by default (-1) when the caller's memory is PAGEABLE (R's vectors are): row
      // indices staged, values narrowed, results landed in page-locked memory.  Pinned caller memory
      // travels directly by default.
      auto pageable = [](const void* p) {
        cudaPointerAttributes a;
        if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return p != nullptr; }
        return a.type == cudaMemoryTypeUnregistered;
      };
      const bool one_seg = v.segs.size() == 1;
      const bool in_pageable = one_seg && pageable(v.segs[0].x) && pageable(v.segs[0].rowidx);
      const bool want_workers = h->opt_host_narrow >= 1 || (h->opt_host_narrow < 0 && in_pageable && nnz >= (1 << 20));
      auto ensure_pinned = [&](void** buf, size_t* cap, size_t bytes) {
        if (*cap >= bytes) return true;
        if (*buf) cudaFreeHost(*buf);
        *buf = nullptr;
        *cap = 0;
        if (cudaHostAlloc(buf, bytes, cudaHostAllocDefault) == cudaSuccess) { *cap = bytes; return true; }
        cudaGetLastError();                     // no page-locked memory to spare: direct copies
        return false;
      };
      if (want_workers && !h->narrower) {
        int ndev = 1;
        cudaGetDeviceCount(&ndev);
        const int hw = (int)std::thread::hardware_concurrency();
        // staging threads per engine: the host's cores shared among the visible GPUs, one left for the caller,
        // at most 12 (pageable e2e, config 2, 16-core host: 31.5 / 26.5-27.4 / 24.3 / 23.9 ms with 3 / 6 / 10 / 14)
        const int t = h->opt_host_narrow > 1 ? h->opt_host_narrow : std::max(1, std::min(12, hw / std::max(ndev, 1) - 1));
        h->narrower = cv_narrower_create(t);
      }
      if (want_workers && xb > 1 && one_seg) {
        size_t cap8 = h->stage8_cap;
        void* b8 = h->stage8;
        if (ensure_pinned(&b8, &cap8, (size_t)nnz)) {
          h->stage8 = (uint8_t*)b8; h->stage8_cap = cap8;
          const int32_t* ri = nullptr;
          if (in_pageable) {
            size_t capr = h->stage_ri_cap * 4;
            void* br = h->stage_ri;
            if (ensure_pinned(&br, &capr, (size_t)nnz * 4)) { ri = v.segs[0].rowidx; io.stage_ri = true; }
            h->stage_ri = (int32_t*)br; h->stage_ri_cap = capr / 4;
          }
          cv_narrower_start(h->narrower, v.segs[0].x, x_type, nnz, h->stage8, ri, h->stage_ri);
          io.narrow = true;
        } else { h->stage8 = (uint8_t*)b8; h->stage8_cap = cap8; }
      }
      if (want_workers && (out_dev || out_z) && (pageable(out_dev) || pageable(out_z)) &&
          (size_t)K * (size_t)N * 16 <= ((size_t)2 << 30)) {
        size_t capo = h->stage_out_cap;
        void* bo = h->stage_out;
        if (ensure_pinned(&bo, &capo, (size_t)K * (size_t)N * 16)) io.land_out = true;
        h->stage_out = (char*)bo; h->stage_out_cap = capo;
      }
      rc = deviations_run(h, 0, 0, &io);
      if (h->narrower) cv_narrower_finish(h->narrower);      // incl. the last slabs of the results
      if (rc == CV_OK) {
        CV_TRY(download_results(h, io.results_streamed ? nullptr : out_dev, io.results_streamed ? nullptr : out_z,
                                out_matches, out_overlap, out_variability, nullptr, nullptr));
        h->st.ms_h2d = wall_ms() - t0;          // whole call: copies and kernels overlap
        h->st.h2d_bytes = h2d_b + io.x_bytes_sent;
        h->st.d2h_bytes += 2 * K * N * 8;
        return CV_OK;
      }
      cudaStreamSynchronize(h->s_up);
      cudaStreamSynchronize(h->s_dn);
      cudaStreamSynchronize(h->stream);
      if (!io.soft_fail) return rc;             // a real input error: same message as the plain sequence
      h->have_counts = false;                   // not exact on the dense path: redo as the plain sequence
    }
  }
  CV_TRY(cv_set_counts_view(h, v));
  std::vector<double> e_default;
  if (!expectation) {
    e_default.resize((size_t)P);
    CV_TRY(cv_expectations(h, 0, nullptr, 0, e_default.data()));
    expectation = e_default.data();
  }
  const int rc_plain = cv_deviations(h, K, annoptr, annoidx, index_base, bg, B, expectation, out_dev, out_z, out_matches,
                                     out_overlap, out_variability, nullptr, nullptr);
  h->f4_veto = false;
  return rc_plain;
}

extern "C" int cv_variability_partials_device(cv_handle* h, void** dptr, int64_t* n_elems) {
  if (!h || !dptr) return CV_ERR_INVALID;
  CV_TRY(cv_settle(h));
  if (!h->have_results) CV_FAIL(h, CV_ERR_STATE, "no results");
  *dptr = h->varfinal.p;
  if (n_elems) *n_elems = h->K * 4;
  return CV_OK;
}

// parts[n_parts][K][4] (device memory: an all-gather of every shard's varfinal, rank order) -> variab[K]
__global__ void __launch_bounds__(128)
k_var_merge_gathered(const double* __restrict__ parts, int n_parts, int K, double* __restrict__ variab) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  Welford tot = {0.0, 0.0, 0.0};
  for (int s = 0; s < n_parts; ++s) {
    const double* q = parts + ((int64_t)s * K + k) * 4;
    Welford b = {q[0], q[1], q[2]};
    tot = welford_merge(tot, b);
  }
  variab[k] = tot.n > 1.0 ? sqrt(tot.m2 / (tot.n - 1.0)) : __longlong_as_double(0x7ff8000000000000LL);
}

// Reduction 2 of the multi-process layout with the host out of the loop: the caller all-gathers the
// K x 4 partial tables on the device (NCCL, stream-ordered behind the step's kernels) and this merges
// them into the handle's variability vector on the same stream.
extern "C" int cv_variability_merge_device(cv_handle* h, const void* d_parts, int n_parts) {
  if (!h || !d_parts || n_parts < 1) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (!h->have_results) CV_FAIL(h, CV_ERR_STATE, "no results");
  CV_LAUNCH(h, k_var_merge_gathered, (unsigned)((h->K + 127) / 128), 128, 0, reinterpret_cast<const double*>(d_parts),
            n_parts, (int)h->K, h->variab.as<double>());
  return CV_OK;
}

// host-side merge of per-shard (n, mean, M2, n_nonfinite) tables, Chan et al. parallel formula
extern "C" int cv_merge_variability(const double* parts, int n_parts, int64_t K, int na_rm,
                                    double* out_sd) {
  if (!parts || !out_sd || n_parts < 1 || K < 0) return CV_ERR_INVALID;
  const double nan = __builtin_nan("");
  for (int64_t k = 0; k < K; ++k) {
    double n = 0, mean = 0, m2 = 0, nf = 0;
    for (int s = 0; s < n_parts; ++s) {
      const double* q = parts + ((int64_t)s * K + k) * 4;
      nf += q[3];
      if (q[0] == 0) continue;
      if (n == 0) { n = q[0]; mean = q[1]; m2 = q[2]; continue; }
      double tot = n + q[0], d = q[1] - mean;
      mean += d * (q[0] / tot);
      m2 += q[2] + d * d * (n * q[0] / tot);
      n = tot;
    }
    if (!na_rm && nf > 0) out_sd[k] = nan;          // arma::stddev over a row holding NaN
    else out_sd[k] = n > 1 ? sqrt(m2 / (n - 1)) : nan;
  }
  return CV_OK;
}
