// Handle lifecycle, counts upload, K1 (per-peak / per-cell / total fragment sums) and the
// cell-tiled row layout of the counts matrix that the gather-contraction reads.
//
// Replaces (reference, chromVAR 1.5.0):
//   getFragmentsPerPeak / getFragmentsPerSample / getTotalFragments  R/helper_functions.R:48-89
//   counts_check                                                     R/utils.R:3-12
//   compute_expectations_core                                        R/compute_deviations.R:415-448
#include <math_constants.h>
#include <stdlib.h>

#include "cv_internal.cuh"

static std::string g_create_err;

int cv_ensure(cv_handle* h, DevBuf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.cap >= bytes) return CV_OK;
  if (b.p) {
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + (bytes >> 4);  // a little slack so slowly growing inputs do not realloc
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc(&b.p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    b.p = nullptr;
    CV_FAIL(h, CV_ERR_OOM, "device allocation of %zu bytes failed: %s", bytes,
            cudaGetErrorString(e));
  }
  b.cap = want;
  return CV_OK;
}

void cv_release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

int cv_sync_check(cv_handle* h) {
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

extern "C" int cv_version(void) { return 200; }

extern "C" int cv_create(int device, uint64_t seed, cv_handle** out) {
  if (!out) return CV_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_err = std::string("no CUDA device available (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                   "); libchromvar_b200 has no CPU fallback";
    return CV_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_err = "device ordinal out of range";
    return CV_ERR_INVALID;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return CV_ERR_CUDA;
  }
  if (prop.major != 10) {
    char b[256];
    snprintf(b, sizeof(b),
             "device %d is sm_%d%d; libchromvar_b200 is built for sm_100a (B200) only", device,
             prop.major, prop.minor);
    g_create_err = b;
    return CV_ERR_UNSUPPORTED;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return CV_ERR_CUDA;
  }
  cv_handle* h = new cv_handle();
  h->device = device;
  h->seed = seed;
  h->opt_trace = getenv("CV_TRACE") != nullptr;
  h->opt_sync_launch = getenv("CV_SYNC_LAUNCH") != nullptr;
  h->num_sms = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  // the compute stream outranks the helpers: where its kernels and the auxiliary stream's table builders
  // (or the upload stream's K1) have blocks pending at the same time, the critical path is served first
  // (CV_STREAM_PRIORITY=0 in the environment: all streams alike)
  int prio_lo = 0, prio_hi = 0;
  const char* sp = getenv("CV_STREAM_PRIORITY");
  if (!(sp && sp[0] == '0')) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);      // lo = least, hi = greatest (numerically smaller)
  if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
    g_create_err = "cudaStreamCreate failed";
    delete h;
    return CV_ERR_CUDA;
  }
  if (cudaStreamCreateWithPriority(&h->s_up, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->s_aux, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->s_dn, cudaStreamNonBlocking, prio_lo) != cudaSuccess) {
    g_create_err = "cudaStreamCreate failed";
    delete h;
    return CV_ERR_CUDA;
  }
  for (int i = 0; i < 8; ++i) cudaEventCreate(&h->ev[i]);
  for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&h->ev_aux[i], cudaEventDisableTiming);
  if (cudaHostAlloc((void**)&h->pinned, CV_PINNED_WORDS * 4, cudaHostAllocDefault) != cudaSuccess) {
    g_create_err = "cudaHostAlloc failed";
    cudaGetLastError();
    delete h;
    return CV_ERR_OOM;
  }
  *out = h;
  return CV_OK;
}

extern "C" void cv_destroy(cv_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  h->async_pending = false;
  cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = {&h->colptr, &h->rowidx, &h->val, &h->xraw, &h->peak_tot, &h->peak_loc,
                    &h->cell_tot, &h->fps, &h->flags, &h->tcnt, &h->rowptr, &h->entries,
                    &h->scan_tmp, &h->annoptr, &h->anno_raw, &h->annoidx, &h->bg_raw, &h->bgT, &h->expect,
                    &h->etab, &h->overlap, &h->matches, &h->mark, &h->order, &h->work_ctr,
                    &h->dev_rm, &h->z_rm, &h->dev_cm, &h->z_cm, &h->varpart, &h->varfinal,
                    &h->variab, &h->obs, &h->samp, &h->tmp0, &h->tmp1, &h->tmp2, &h->dn_M, &h->dn_C, &h->dn_S,
                    &h->dn_abits, &h->dn_abitsT, &h->dn_invptr, &h->dn_invsrc, &h->dn_E2, &h->dn_ov, &h->dn_winv,
                    &h->f4_Tq, &h->f4_Sq, &h->f4_ncol, &h->f4_off, &h->f4_colq, &h->f4_colt, &h->f4_info, &h->f4_sync};
  for (DevBuf* b : bufs) cv_release(*b);
  for (DevBuf& b : h->bk) cv_release(b);
  for (int i = 0; i < 8; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < 4; ++i)
    if (h->ev_aux[i]) cudaEventDestroy(h->ev_aux[i]);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->narrower) cv_narrower_destroy(h->narrower);
  if (h->stage8) cudaFreeHost(h->stage8);
  if (h->stage_ri) cudaFreeHost(h->stage_ri);
  if (h->stage_out) cudaFreeHost(h->stage_out);
  if (h->s_up) cudaStreamDestroy(h->s_up);
  if (h->s_dn) cudaStreamDestroy(h->s_dn);
  if (h->s_aux) cudaStreamDestroy(h->s_aux);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" const char* cv_last_error(const cv_handle* h) {
  return h ? h->err.c_str() : g_create_err.c_str();
}

extern "C" int cv_set_seed(cv_handle* h, uint64_t seed) {
  if (!h) return CV_ERR_INVALID;
  h->seed = seed;
  return CV_OK;
}

extern "C" int cv_set_option(cv_handle* h, const char* key, int64_t value) {
  if (!h || !key) return CV_ERR_INVALID;
  CV_TRY(cv_settle(h));
  h->replay_valid = false;                    // options change which kernels run
  if (!strcmp(key, "path")) {
    if (value < 0 || value > 2) CV_FAIL(h, CV_ERR_INVALID, "path must be 0 (auto), 1 (row-gather SpMM) or 2 (dense tcgen05)");
    h->opt_path = (int)value;
    return CV_OK;
  }
  if (!strcmp(key, "dense_cta_pair")) { h->opt_pair = value < 0 ? -1 : (value != 0); return CV_OK; }
  if (!strcmp(key, "cell_chunk")) {
    if (value < 0) CV_FAIL(h, CV_ERR_INVALID, "cell_chunk must be >= 0");
    h->opt_chunk = value;
    return CV_OK;
  }
  if (!strcmp(key, "dense_rows_scatter")) { h->opt_rows_scatter = value != 0; return CV_OK; }
  if (!strcmp(key, "dense_epilogue")) {
    if (value < 0 || value > 2) CV_FAIL(h, CV_ERR_INVALID, "dense_epilogue must be 0 (auto), 1 (float pairs) or 2 (fp64)");
    h->opt_epilogue = (int)value;
    return CV_OK;
  }
  if (!strcmp(key, "dense_lockstep")) { h->opt_lockstep = value == 2 ? 2 : (value != 0); return CV_OK; }
  if (!strcmp(key, "dense_fp4")) { h->opt_fp4 = value < 0 ? -1 : (value ? 1 : 0); return CV_OK; }
  if (!strcmp(key, "dense_band")) { h->opt_band = (int)value; return CV_OK; }
  if (!strcmp(key, "knn_insertion")) { h->opt_knn_insertion = value != 0; return CV_OK; }
  if (!strcmp(key, "debug_no_tma")) { h->opt_debug_no_tma = (int)value; return CV_OK; }
  if (!strcmp(key, "small_n_gather")) { h->opt_gather = value < 0 ? -1 : (value != 0); return CV_OK; }
  if (!strcmp(key, "check_zero_peaks")) { h->opt_check_zero = value != 0; return CV_OK; }
  if (!strcmp(key, "tail_chunk")) { h->opt_tail_chunk = value != 0; return CV_OK; }
  if (!strcmp(key, "host_narrow")) {
    // -1 (default): host threads stage / narrow PAGEABLE caller memory (cv_hostnarrow.cpp); 0 never; n >= 1
    // always, n = 1 with the default number of threads, n > 1 with n threads.  Measured for PINNED caller
    // memory (profiles/r01u_summary.md): neutral on one GPU (the call is compute-bound), a loss on an
    // 8-GPU host with 4 cores per GPU (the conversion competes for cores) -- hence off for pinned memory
    h->opt_host_narrow = (int)(value < 0 ? -1 : value);
    if (h->narrower && value > 1 && cv_narrower_threads(h->narrower) != (int)value) {
      cv_narrower_destroy(h->narrower);
      h->narrower = nullptr;
    }
    return CV_OK;
  }
  CV_FAIL(h, CV_ERR_INVALID, "unknown option %s", key);
}

// ---- timed spans ----------------------------------------------------------------------------------------
static cudaEvent_t span_event(cv_handle* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}
void cv_spans_reset(cv_handle* h) {
  h->ev_used = 0;
  h->spans.clear();
}
void cv_span_begin(cv_handle* h, int cat) {
  cv_handle::Span sp;
  sp.a = span_event(h);
  sp.b = span_event(h);
  sp.cat = cat;
  cudaEventRecord(sp.a, h->stream);
  h->spans.push_back(sp);
}
void cv_span_end(cv_handle* h) { cudaEventRecord(h->spans.back().b, h->stream); }
void cv_spans_collect(cv_handle* h, double* ms_by_cat) {
  for (int i = 0; i < CV_SPAN_N; ++i) ms_by_cat[i] = 0.0;
  for (const auto& sp : h->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) ms_by_cat[sp.cat] += ms;
    else cudaGetLastError();
  }
}

extern "C" void* cv_stream(cv_handle* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int cv_stats(cv_handle* h, cv_stage_stats* out) {
  if (!h || !out) return CV_ERR_INVALID;
  CV_TRY(cv_settle(h));
  h->st.kernel_launches = h->launches;
  *out = h->st;
  return CV_OK;
}

// =========================================================================================
// device helpers
// =========================================================================================
// flags[0] bit0: negative / non-finite / non-integer value; bit1: row index out of range;
//          bit2: value >= 2^32
// flags[1]: max value
enum { FLAG_BADVAL = 1u, FLAG_BADROW = 2u, FLAG_HUGE = 4u, FLAG_BADIDX = 8u, FLAG_ZEROPEAK = 16u, FLAG_FRACTION = 32u };

// column of nonzero e: largest c with colptr[c] <= e   (searched in [lo, hi])
__device__ __forceinline__ int find_col(const int64_t* __restrict__ colptr, int lo, int hi,
                                        int64_t e) {
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(colptr + mid) <= e) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Column lookup for a warp holding consecutive nonzeros: two full searches (first and last
// lane), the rest search the (usually empty) range in between.
__device__ __forceinline__ int find_col_warp(const int64_t* __restrict__ colptr, int N,
                                             int64_t e, int64_t nnz, bool active) {
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x & 31;
  int64_t e_first = __shfl_sync(full, e, 0);
  int64_t e_last = e_first + 31 < nnz ? e_first + 31 : nnz - 1;
  int c = 0;
  if (lane == 0) c = find_col(colptr, 0, N - 1, e_first);
  if (lane == 31) c = find_col(colptr, 0, N - 1, e_last);
  int c_lo = __shfl_sync(full, c, 0);
  int c_hi = __shfl_sync(full, c, 31);
  if (c_lo == c_hi) return c_lo;
  return active ? find_col(colptr, c_lo, c_hi, e) : c_lo;
}

template <typename XT>
__device__ __forceinline__ uint32_t load_count(const XT* __restrict__ x, int64_t e,
                                               uint32_t& bad);
template <>
__device__ __forceinline__ uint32_t load_count<double>(const double* __restrict__ x, int64_t e,
                                                       uint32_t& bad) {
  double v = x[e];
  if (!(v >= 0.0) || v == CUDART_INF) { bad |= FLAG_BADVAL; return 0; }  // also catches NaN
  if (v != floor(v)) { bad |= FLAG_FRACTION; return 0; }
  if (v >= 4294967296.0) { bad |= FLAG_HUGE; return 0; }
  return (uint32_t)v;
}
template <>
__device__ __forceinline__ uint32_t load_count<float>(const float* __restrict__ x, int64_t e,
                                                      uint32_t& bad) {
  float v = x[e];
  if (!(v >= 0.0f) || v == CUDART_INF_F) { bad |= FLAG_BADVAL; return 0; }
  if (v != floorf(v)) { bad |= FLAG_FRACTION; return 0; }
  if (v >= 4294967296.0f) { bad |= FLAG_HUGE; return 0; }
  return (uint32_t)v;
}
template <>
__device__ __forceinline__ uint32_t load_count<int32_t>(const int32_t* __restrict__ x,
                                                        int64_t e, uint32_t& bad) {
  int32_t v = x[e];
  if (v < 0) { bad |= FLAG_BADVAL; return 0; }
  return (uint32_t)v;
}
template <>
__device__ __forceinline__ uint32_t load_count<uint8_t>(const uint8_t* __restrict__ x,
                                                        int64_t e, uint32_t& bad) {
  return (uint32_t)x[e];
}

// =========================================================================================
// K1: one pass over the CSC nonzeros -> validated uint32 values, per-peak totals (atomics
// spread over P addresses), per-cell totals (warp-aggregated), max value.
// HBM-bound: reads nnz*(4 + sizeof(XT)), writes nnz*4.
// =========================================================================================
template <typename XT>
__global__ void __launch_bounds__(256)
k_counts_sums(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
              const XT* __restrict__ x, int64_t e_begin, int64_t nnz /* end of the range */, int P, int N,
              uint32_t* __restrict__ val, unsigned long long* __restrict__ peak_tot,
              unsigned long long* __restrict__ cell_tot, uint32_t* __restrict__ flags) {
  const unsigned full = 0xffffffffu;
  int lane = threadIdx.x & 31;
  uint32_t bad = 0, vmax = 0;
  bool unsorted = false, dup = false;
  int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t base = e_begin + warp_id * 32; base < nnz; base += warps_total * 32) {
    int64_t e = base + lane;
    bool active = e < nnz;
    int c = find_col_warp(colptr, N, active ? e : nnz - 1, nnz, active);
    uint32_t v = 0;
    if (active) {
      v = load_count<XT>(x, e, bad);
      int r = rowidx[e];
      // flags[4]: some column lists its rows out of order (legal, just rare: R's dgCMatrix and scipy keep
      // them sorted); the per-range builders of the dense path then scan whole columns instead of searching
      if (e > __ldg(colptr + c)) {
        const int rp = rowidx[e - 1];
        if (rp > r) unsorted = true;
        if (rp == r) dup = true;                          // adjacent duplicates (sorted input); unsorted input counts as suspect
      }
      if (r < 0 || r >= P) { bad |= FLAG_BADROW; v = 0; }
      else if (v) atomicAdd(peak_tot + r, (unsigned long long)v);
      val[e] = v;
      vmax = max(vmax, v);
    }
    int c0 = __shfl_sync(full, c, 0);
    bool uniform = __all_sync(full, c == c0);
    if (uniform) {
      unsigned long long s = warp_sum<unsigned long long>(v);
      if (lane == 0 && s) atomicAdd(cell_tot + c0, s);
    } else if (active && v) {
      atomicAdd(cell_tot + c, (unsigned long long)v);
    }
  }
  vmax = max(vmax, __shfl_xor_sync(full, vmax, 16));
  vmax = max(vmax, __shfl_xor_sync(full, vmax, 8));
  vmax = max(vmax, __shfl_xor_sync(full, vmax, 4));
  vmax = max(vmax, __shfl_xor_sync(full, vmax, 2));
  vmax = max(vmax, __shfl_xor_sync(full, vmax, 1));
  bad |= __shfl_xor_sync(full, bad, 16);
  bad |= __shfl_xor_sync(full, bad, 8);
  bad |= __shfl_xor_sync(full, bad, 4);
  bad |= __shfl_xor_sync(full, bad, 2);
  bad |= __shfl_xor_sync(full, bad, 1);
  if (__any_sync(full, unsorted) && lane == 0) atomicOr(flags + 4, 1u);
  if (__any_sync(full, dup) && lane == 0) atomicOr(flags + 4, 2u);
  if (lane == 0) {
    if (bad) atomicOr(flags + 0, bad);
    if (vmax) atomicMax(flags + 1, vmax);
  }
}

// grand total (slot P) + fps (double) + zero-peak check; single block, deterministic
__global__ void __launch_bounds__(1024)
k_totals_finish(unsigned long long* __restrict__ peak_tot, int P) {
  __shared__ unsigned long long sh[32];
  unsigned long long s = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) s += peak_tot[i];
  s = warp_sum<unsigned long long>(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0ull;
    s = warp_sum<unsigned long long>(s);
    if (threadIdx.x == 0) peak_tot[P] = s;
  }
}

// fragments per sample as doubles + the largest cell total (bounds the s32 accumulators)
__global__ void k_u64_to_f64(const unsigned long long* __restrict__ in, double* __restrict__ out,
                             int64_t n, unsigned long long* __restrict__ vmax) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = i < n ? in[i] : 0ull;
  if (i < n) out[i] = (double)v;
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  if ((threadIdx.x & 31) == 0 && v) atomicMax(vmax, v);
}

// dense column-major -> CSC (rows within a column land in arbitrary order; nothing downstream
// depends on the order).  Pass 1 counts nonzeros per column, pass 2 writes them.
// Dense P x N input -> CSC with the rows of every column in increasing order (what dgCMatrix and
// scipy guarantee, and what the per-range builders of the dense path search in): per-(column, block of
// 256 rows) nonzero counts, a scan over the blocks of each column, then an ordered fill.
template <typename XT>
__global__ void __launch_bounds__(256)
k_dense_count(const XT* __restrict__ x, int64_t P, int64_t N, uint32_t* __restrict__ blkcnt) {
  __shared__ uint32_t s_w[8];
  const int64_t c = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool nz = r < P && x[c * P + r] != (XT)0;
  const unsigned m = __ballot_sync(0xffffffffu, nz);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    blkcnt[c * gridDim.x + blockIdx.x] = t;
  }
}

// one CTA per column: exclusive scan of its block counts (in place), column total -> cnt[c]
__global__ void __launch_bounds__(1024)
k_dense_blkscan(uint32_t* __restrict__ blkcnt, int nblk, unsigned long long* __restrict__ cnt) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  uint32_t* row = blkcnt + (int64_t)blockIdx.x * nblk;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblk ? row[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblk) row[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt[blockIdx.x] = (unsigned long long)carry;
}

template <typename XT>
__global__ void __launch_bounds__(256)
k_dense_fill(const XT* __restrict__ x, int64_t P, int64_t N, const int64_t* __restrict__ colptr,
             const uint32_t* __restrict__ blkoff, int32_t* __restrict__ rowidx, XT* __restrict__ xout) {
  __shared__ uint32_t s_w[8];
  const int64_t c = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const XT v = r < P ? x[c * P + r] : (XT)0;
  const bool nz = v != (XT)0;
  const unsigned m = __ballot_sync(0xffffffffu, nz);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_w[warp] = __popc(m);
  __syncthreads();
  if (nz) {
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    const int64_t pos = colptr[c] + blkoff[c * gridDim.x + blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
    rowidx[pos] = (int32_t)r;
    xout[pos] = v;
  }
}

// exclusive scan of u64 counts into int64 colptr (N+1) -- small (N elements), single block
__global__ void __launch_bounds__(1024)
k_scan_small_u64(const unsigned long long* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  __shared__ unsigned long long sh[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    int64_t i = base + threadIdx.x;
    unsigned long long v = i < n ? in[i] : 0ull;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0ull;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) out[i] = (int64_t)(carry + sh[threadIdx.x] - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = (int64_t)carry;
}

__global__ void k_widen_i32(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// =========================================================================================
// Cell-tiled row layout ("tile CSR"): for each tile of W consecutive cells, a CSR over peaks
// whose entries are (cell_in_tile << 20 | count).  Counts >= 2^20 are split over several
// entries, zero counts are dropped.
//   pass A: tcnt[t*P + q] += pieces(v)        (atomics spread over T*P addresses)
//   scan  : rowptr = exclusive_scan(tcnt)
//   pass B: slot = rowptr[..] + atomicSub(tcnt[..]) ; write entries
// =========================================================================================
#define CV_ENTRY_VAL_BITS 20
#define CV_ENTRY_VAL_MASK 0xFFFFFu

__device__ __forceinline__ uint32_t entry_pieces(uint32_t v) {
  return (v + (CV_ENTRY_VAL_MASK - 1)) / CV_ENTRY_VAL_MASK;   // 0 for v == 0
}

__global__ void __launch_bounds__(256)
k_layout_count(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
               const uint32_t* __restrict__ val, int64_t nnz, int P, int N, int W,
               uint32_t* __restrict__ tcnt) {
  int lane = threadIdx.x & 31;
  int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t base = warp_id * 32; base < nnz; base += warps_total * 32) {
    int64_t e = base + lane;
    bool active = e < nnz;
    int c = find_col_warp(colptr, N, active ? e : nnz - 1, nnz, active);
    if (active) {
      uint32_t pieces = entry_pieces(val[e]);
      if (pieces) atomicAdd(tcnt + (int64_t)(c / W) * P + rowidx[e], pieces);
    }
  }
}

__global__ void __launch_bounds__(256)
k_layout_fill(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
              const uint32_t* __restrict__ val, int64_t nnz, int P, int N, int W,
              uint32_t* __restrict__ tcnt, const uint32_t* __restrict__ rowptr,
              uint32_t* __restrict__ entries) {
  int lane = threadIdx.x & 31;
  int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t base = warp_id * 32; base < nnz; base += warps_total * 32) {
    int64_t e = base + lane;
    bool active = e < nnz;
    int c = find_col_warp(colptr, N, active ? e : nnz - 1, nnz, active);
    if (active) {
      uint32_t v = val[e];
      uint32_t pieces = entry_pieces(v);
      if (pieces) {
        int t = c / W;
        int64_t idx = (int64_t)t * P + rowidx[e];
        uint32_t old = atomicSub(tcnt + idx, pieces);
        uint32_t slot = rowptr[idx] + (old - pieces);
        uint32_t cl = (uint32_t)(c - t * W) << CV_ENTRY_VAL_BITS;
        while (v) {
          uint32_t piece = v > CV_ENTRY_VAL_MASK ? CV_ENTRY_VAL_MASK : v;
          entries[slot++] = cl | piece;
          v -= piece;
        }
      }
    }
  }
}

// ---- exclusive scan u32 (n up to ~1e8): block-local scan + block totals + add ------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
             unsigned long long* __restrict__ block_sums) {
  __shared__ uint32_t wsum[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? in[base + i] : 0u;
    s += v[i];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int i = 0; i < w; ++i) woff += wsum[i];
  uint32_t run = woff + inc - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == SCAN_THREADS - 1) block_sums[blockIdx.x] = (unsigned long long)run;
}

// scan of the block totals (single block, sequential over chunks of 1024)
__global__ void __launch_bounds__(1024)
k_scan_sums(unsigned long long* __restrict__ sums, int64_t nb, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long sh[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += 1024) {
    int64_t i = base + threadIdx.x;
    unsigned long long v = i < nb ? sums[i] : 0ull;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0ull;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_add(uint32_t* __restrict__ out, int64_t n, const unsigned long long* __restrict__ sums,
           const unsigned long long* __restrict__ total) {
  uint32_t off = (uint32_t)sums[blockIdx.x];
  int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) out[base + i] += off;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (uint32_t)(*total);
}

// out has n+1 slots; asynchronous (the total stays on the device behind the block sums)
int cv_exclusive_scan_u32_async(cv_handle* h, const uint32_t* in, uint32_t* out, int64_t n) {
  int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (nb < 1) nb = 1;
  CV_TRY(cv_ensure(h, h->scan_tmp, (size_t)(nb + 1) * 8));
  unsigned long long* sums = h->scan_tmp.as<unsigned long long>();
  CV_LAUNCH(h, k_scan_block, (unsigned)nb, SCAN_THREADS, 0, in, out, n, sums);
  CV_LAUNCH(h, k_scan_sums, 1, 1024, 0, sums, nb, sums + nb);
  CV_LAUNCH(h, k_scan_add, (unsigned)nb, SCAN_THREADS, 0, out, n, sums, sums + nb);
  return CV_OK;
}

// out has n+1 slots; returns total through *total_host
static int exclusive_scan_u32(cv_handle* h, const uint32_t* in, uint32_t* out, int64_t n,
                              unsigned long long* total_host) {
  CV_TRY(cv_exclusive_scan_u32_async(h, in, out, n));
  int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (nb < 1) nb = 1;
  CV_CUDA(h, cudaMemcpyAsync(total_host, h->scan_tmp.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

static unsigned grid_for(cv_handle* h, int64_t items, int per_block) {
  int64_t g = (items + per_block - 1) / per_block;
  int64_t cap = (int64_t)h->num_sms * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

int cv_build_layout(cv_handle* h, int W) {
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  if (W < 32 || W > 4096) CV_FAIL(h, CV_ERR_UNSUPPORTED, "cell tile %d outside [32,4096]", W);
  int T = (int)((h->N + W - 1) / W);
  int64_t n = (int64_t)T * h->P;
  CV_TRY(cv_ensure(h, h->tcnt, (size_t)n * 4));
  CV_TRY(cv_ensure(h, h->rowptr, (size_t)(n + 1) * 4));
  CV_CUDA(h, cudaMemsetAsync(h->tcnt.p, 0, (size_t)n * 4, h->stream));
  unsigned grid = grid_for(h, h->nnz, 256 * 4);
  if (h->nnz > 0)
    CV_LAUNCH(h, k_layout_count, grid, 256, 0, h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(),
              h->val.as<uint32_t>(), h->nnz, (int)h->P, (int)h->N, W, h->tcnt.as<uint32_t>());
  unsigned long long total = 0;
  CV_TRY(exclusive_scan_u32(h, h->tcnt.as<uint32_t>(), h->rowptr.as<uint32_t>(), n, &total));
  if (total >= (1ull << 32))
    CV_FAIL(h, CV_ERR_UNSUPPORTED, "counts layout needs %llu entries (>= 2^32) on one device",
            total);
  h->n_entries = (int64_t)total;
  CV_TRY(cv_ensure(h, h->entries, (size_t)(total + 64) * 4));
  if (h->nnz > 0)
    CV_LAUNCH(h, k_layout_fill, grid, 256, 0, h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(),
              h->val.as<uint32_t>(), h->nnz, (int)h->P, (int)h->N, W, h->tcnt.as<uint32_t>(),
              h->rowptr.as<uint32_t>(), h->entries.as<uint32_t>());
  h->W = W;
  h->T = T;
  h->layout_valid = true;
  h->st.cell_tile = W;
  h->st.n_cell_tiles = T;
  return CV_OK;
}

// =========================================================================================
// counts upload
// =========================================================================================
static size_t type_bytes(int t) {
  switch (t) {
    case CV_I32: return 4;
    case CV_I64: return 8;
    case CV_F64: return 8;
    case CV_U8: return 1;
    case CV_F32: return 4;
  }
  return 0;
}

static double wall_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// K1 on the device-resident CSC (colptr int64, rowidx, xraw of x_type), in pieces so that the
// streamed call can run it per uploaded cell chunk:
//   cv_counts_begin         buffers + zeroed totals / flags
//   cv_counts_range         nonzeros [e0, e1) = cells [c0, c1): validation, val, totals, fps
//   cv_counts_finish_async  grand total
//   cv_counts_verdict       host-side reading of the flag words (after a stream sync)
int cv_counts_begin(cv_handle* h) {
  int64_t P = h->P, N = h->N, nnz = h->nnz;
  CV_TRY(cv_ensure(h, h->val, (size_t)nnz * 4));
  CV_TRY(cv_ensure(h, h->peak_tot, (size_t)(P + 1) * 8));
  CV_TRY(cv_ensure(h, h->cell_tot, (size_t)N * 8));
  CV_TRY(cv_ensure(h, h->fps, (size_t)N * 8));
  CV_TRY(cv_ensure(h, h->flags, 64));
  CV_CUDA(h, cudaMemsetAsync(h->peak_tot.p, 0, (size_t)(P + 1) * 8, h->stream));
  CV_CUDA(h, cudaMemsetAsync(h->cell_tot.p, 0, (size_t)N * 8, h->stream));
  CV_CUDA(h, cudaMemsetAsync(h->flags.p, 0, 32, h->stream));
  h->have_counts = false;
  return CV_OK;
}

int cv_counts_range(cv_handle* h, int x_type, int64_t e0, int64_t e1, int64_t c0, int64_t c1) {
  int64_t P = h->P, N = h->N;
  unsigned grid = grid_for(h, e1 - e0, 256 * 4);
  auto pt = h->peak_tot.as<unsigned long long>();
  auto ct = h->cell_tot.as<unsigned long long>();
  if (e1 > e0) {
    switch (x_type) {
      case CV_F64:
        CV_LAUNCH(h, k_counts_sums<double>, grid, 256, 0, h->colptr.as<int64_t>(),
                  h->rowidx.as<int32_t>(), h->xraw.as<double>(), e0, e1, (int)P, (int)N,
                  h->val.as<uint32_t>(), pt, ct, h->flags.as<uint32_t>());
        break;
      case CV_F32:
        CV_LAUNCH(h, k_counts_sums<float>, grid, 256, 0, h->colptr.as<int64_t>(),
                  h->rowidx.as<int32_t>(), h->xraw.as<float>(), e0, e1, (int)P, (int)N,
                  h->val.as<uint32_t>(), pt, ct, h->flags.as<uint32_t>());
        break;
      case CV_I32:
        CV_LAUNCH(h, k_counts_sums<int32_t>, grid, 256, 0, h->colptr.as<int64_t>(),
                  h->rowidx.as<int32_t>(), h->xraw.as<int32_t>(), e0, e1, (int)P, (int)N,
                  h->val.as<uint32_t>(), pt, ct, h->flags.as<uint32_t>());
        break;
      case CV_U8:
        CV_LAUNCH(h, k_counts_sums<uint8_t>, grid, 256, 0, h->colptr.as<int64_t>(),
                  h->rowidx.as<int32_t>(), h->xraw.as<uint8_t>(), e0, e1, (int)P, (int)N,
                  h->val.as<uint32_t>(), pt, ct, h->flags.as<uint32_t>());
        break;
      default:
        CV_FAIL(h, CV_ERR_INVALID, "unsupported value type %d", x_type);
    }
  }
  if (c1 > c0)
    CV_LAUNCH(h, k_u64_to_f64, (unsigned)((c1 - c0 + 255) / 256), 256, 0, ct + c0, h->fps.as<double>() + c0,
              c1 - c0, reinterpret_cast<unsigned long long*>(h->flags.as<uint32_t>() + 2));
  return CV_OK;
}

int cv_counts_finish_async(cv_handle* h) {
  CV_LAUNCH(h, k_totals_finish, 1, 1024, 0, h->peak_tot.as<unsigned long long>(), (int)h->P);
  return CV_OK;
}

int cv_counts_verdict(cv_handle* h, const uint32_t* flags) {
  if (flags[0] & FLAG_BADROW) CV_FAIL(h, CV_ERR_INVALID, "counts: row index outside [0, P)");
  if (flags[0] & FLAG_HUGE) CV_FAIL(h, CV_ERR_UNSUPPORTED, "counts: value >= 2^32");
  if (flags[0] & FLAG_BADVAL)       // counts_check, R/utils.R:3-12: min(assays(object)$counts) >= 0
    CV_FAIL(h, CV_ERR_INVALID, "counts must be finite and non-negative: min(assays(object)$counts) >= 0 is not TRUE (counts_check)");
  if (flags[0] & FLAG_FRACTION)     // a restriction of this engine, not of the reference
    CV_FAIL(h, CV_ERR_UNSUPPORTED,
            "counts hold fractional values: this engine computes exact integer sums and accepts whole-number counts only "
            "(the reference also runs on fractional counts, e.g. depth-normalised matrices; round them or use the R path)");
  h->max_val = flags[1];
  h->max_cell_tot = (uint64_t)flags[2] | ((uint64_t)flags[3] << 32);
  h->have_counts = true;
  h->layout_valid = false;
  h->totals_committed = true;  // single shard until the host says otherwise
  h->have_results = false;
  return CV_OK;
}

static int counts_sums(cv_handle* h, int x_type) {
  CV_TRY(cv_counts_begin(h));
  CV_TRY(cv_counts_range(h, x_type, 0, h->nnz, 0, h->N));
  CV_TRY(cv_counts_finish_async(h));
  uint32_t flags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CV_CUDA(h, cudaMemcpyAsync(flags, h->flags.p, 32, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->cols_dup = flags[4] != 0u;                     // a duplicate row index, or rows out of order (not examined further)
  return cv_counts_verdict(h, flags);
}

// One shard's counts as the host holds them: colptr rebased to the shard (int64), nonzeros in one or
// several segments (a shard may span several column blocks of the caller, cv_multi.cu).
int cv_csc_view_init(cv_handle* h, CscView* v, int64_t P, int64_t N, const void* colptr, int colptr_type,
                     const int32_t* rowidx, const void* x, int x_type) {
  if (P <= 0 || N <= 0 || !colptr) CV_FAIL(h, CV_ERR_INVALID, "counts: empty matrix");
  if (P > 0x7fffffff || N > 0x7fffffff) CV_FAIL(h, CV_ERR_UNSUPPORTED, "P, N must be < 2^31");
  if (colptr_type != CV_I32 && colptr_type != CV_I64)
    CV_FAIL(h, CV_ERR_INVALID, "colptr_type must be CV_I32 or CV_I64");
  const size_t xb = type_bytes(x_type);
  if (!xb || x_type == CV_I64) CV_FAIL(h, CV_ERR_INVALID, "unsupported x_type %d", x_type);
  v->P = P; v->N = N; v->x_type = x_type; v->xb = xb;
  v->cp.resize((size_t)N + 1);
  if (colptr_type == CV_I32) for (int64_t i = 0; i <= N; ++i) v->cp[(size_t)i] = ((const int32_t*)colptr)[i];
  else memcpy(v->cp.data(), colptr, (size_t)(N + 1) * 8);
  const int64_t nnz = v->cp[(size_t)N];
  if (v->cp[0] != 0 || nnz < 0) CV_FAIL(h, CV_ERR_INVALID, "colptr must start at 0 and be monotone");
  if (nnz > 0 && (!rowidx || !x)) CV_FAIL(h, CV_ERR_INVALID, "rowidx / x are NULL");
  v->segs.clear();
  if (nnz > 0) v->segs.push_back(CscSeg{0, nnz, rowidx, (const char*)x});
  return CV_OK;
}

int cv_csc_view_check(cv_handle* h, const CscView& v) {
  const int64_t N = v.N;
  if ((int64_t)v.cp.size() != N + 1 || v.cp[0] != 0) CV_FAIL(h, CV_ERR_INVALID, "colptr must start at 0 and be monotone");
  for (int64_t i = 0; i < N; ++i)
    if (v.cp[(size_t)i + 1] < v.cp[(size_t)i]) CV_FAIL(h, CV_ERR_INVALID, "colptr is not monotone at column %lld", (long long)i);
  if (v.cp[(size_t)N] >= (1ll << 32)) CV_FAIL(h, CV_ERR_UNSUPPORTED, "nnz >= 2^32 on one handle: shard the cells");
  int64_t at = 0;
  for (const CscSeg& sg : v.segs) {
    if (sg.e0 != at || sg.e1 <= sg.e0 || !sg.rowidx || !sg.x) CV_FAIL(h, CV_ERR_INVALID, "counts: segments do not tile the nonzeros");
    at = sg.e1;
  }
  if (at != v.cp[(size_t)N]) CV_FAIL(h, CV_ERR_INVALID, "counts: segments do not tile the nonzeros");
  return CV_OK;
}

// H2D of nonzeros [e0, e1) of the view (row indices and / or values) into the device arrays at the same
// element offsets; one copy per segment touched
int cv_csc_view_copy(cv_handle* h, const CscView& v, int64_t e0, int64_t e1, int32_t* d_rowidx, char* d_x,
                     cudaStream_t st) {
  for (const CscSeg& sg : v.segs) {
    const int64_t a = std::max(e0, sg.e0), b = std::min(e1, sg.e1);
    if (b <= a) continue;
    if (d_rowidx)
      CV_CUDA(h, cudaMemcpyAsync(d_rowidx + a, sg.rowidx + (a - sg.e0), (size_t)(b - a) * 4, cudaMemcpyHostToDevice, st));
    if (d_x)
      CV_CUDA(h, cudaMemcpyAsync(d_x + (size_t)a * v.xb, sg.x + (size_t)(a - sg.e0) * v.xb, (size_t)(b - a) * v.xb,
                                 cudaMemcpyHostToDevice, st));
  }
  return CV_OK;
}

int cv_set_counts_view(cv_handle* h, const CscView& v) {
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  h->replay_valid = false;
  h->have_counts = false;
  CV_TRY(cv_csc_view_check(h, v));
  const int64_t P = v.P, N = v.N, nnz = v.cp[(size_t)N];
  h->P = P; h->N = N; h->nnz = nnz;
  double t0 = wall_ms();
  CV_TRY(cv_ensure(h, h->colptr, (size_t)(N + 1) * 8));
  CV_TRY(cv_ensure(h, h->rowidx, (size_t)nnz * 4));
  CV_TRY(cv_ensure(h, h->xraw, (size_t)nnz * v.xb));
  CV_CUDA(h, cudaMemcpyAsync(h->colptr.p, v.cp.data(), (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, h->stream));
  CV_TRY(cv_csc_view_copy(h, v, 0, nnz, h->rowidx.as<int32_t>(), h->xraw.as<char>(), h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->st.ms_h2d = wall_ms() - t0;
  h->st.h2d_bytes = (int64_t)((N + 1) * 8 + nnz * (4 + (int64_t)v.xb));
  CV_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
  CV_TRY(counts_sums(h, v.x_type));
  CV_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
  CV_CUDA(h, cudaEventSynchronize(h->ev[1]));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->st.ms_counts_layout = ms;
  return CV_OK;
}

extern "C" int cv_set_counts_csc(cv_handle* h, int64_t P, int64_t N, const void* colptr,
                                 int colptr_type, const int32_t* rowidx, const void* x,
                                 int x_type) {
  if (!h) return CV_ERR_INVALID;
  h->have_counts = false;
  CscView v;
  CV_TRY(cv_csc_view_init(h, &v, P, N, colptr, colptr_type, rowidx, x, x_type));
  return cv_set_counts_view(h, v);
}

template <typename XT>
static int dense_to_csc(cv_handle* h, const XT* xd, int64_t P, int64_t N) {
  CV_TRY(cv_ensure(h, h->tmp1, (size_t)(N + 1) * 8));
  CV_TRY(cv_ensure(h, h->colptr, (size_t)(N + 1) * 8));
  auto cnt = h->tmp1.as<unsigned long long>();
  CV_CUDA(h, cudaMemsetAsync(cnt, 0, (size_t)(N + 1) * 8, h->stream));
  dim3 grid((unsigned)((P + 255) / 256), (unsigned)N);
  if (N > 65535) CV_FAIL(h, CV_ERR_UNSUPPORTED, "dense counts input with more than 65535 columns; pass CSC");
  CV_TRY(cv_ensure(h, h->tmp0, (size_t)grid.x * (size_t)N * 4));
  uint32_t* blk = h->tmp0.as<uint32_t>();
  CV_LAUNCH(h, k_dense_count<XT>, grid, 256, 0, xd, P, N, blk);
  CV_LAUNCH(h, k_dense_blkscan, (unsigned)N, 1024, 0, blk, (int)grid.x, cnt);
  CV_LAUNCH(h, k_scan_small_u64, 1, 1024, 0, cnt, h->colptr.as<int64_t>(), N);
  int64_t nnz = 0;
  CV_CUDA(h, cudaMemcpyAsync(&nnz, h->colptr.as<int64_t>() + N, 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  if (nnz >= (1ll << 32)) CV_FAIL(h, CV_ERR_UNSUPPORTED, "nnz >= 2^32 on one handle");
  h->nnz = nnz;
  CV_TRY(cv_ensure(h, h->rowidx, (size_t)nnz * 4));
  CV_TRY(cv_ensure(h, h->xraw, (size_t)nnz * sizeof(XT)));
  CV_LAUNCH(h, k_dense_fill<XT>, grid, 256, 0, xd, P, N, h->colptr.as<int64_t>(), blk,
            h->rowidx.as<int32_t>(), h->xraw.as<XT>());
  return CV_OK;
}

extern "C" int cv_set_counts_dense(cv_handle* h, int64_t P, int64_t N, const void* x, int x_type) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  CV_TRY(cv_settle(h));
  h->replay_valid = false;
  h->have_counts = false;
  if (P <= 0 || N <= 0 || !x) CV_FAIL(h, CV_ERR_INVALID, "counts: empty matrix");
  if (P > 0x7fffffff || N > 0x7fffffff) CV_FAIL(h, CV_ERR_UNSUPPORTED, "P, N must be < 2^31");
  size_t xb = type_bytes(x_type);
  if (x_type != CV_F64 && x_type != CV_I32 && x_type != CV_F32)
    CV_FAIL(h, CV_ERR_INVALID, "dense counts: x_type must be CV_F64, CV_F32 or CV_I32");
  h->P = P; h->N = N;
  double t0 = wall_ms();
  CV_TRY(cv_ensure(h, h->tmp2, (size_t)P * N * xb));
  CV_CUDA(h, cudaMemcpyAsync(h->tmp2.p, x, (size_t)P * N * xb, cudaMemcpyHostToDevice, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->st.ms_h2d = wall_ms() - t0;
  h->st.h2d_bytes = (int64_t)(P * N * xb);
  CV_CUDA(h, cudaEventRecord(h->ev[0], h->stream));
  if (x_type == CV_F64) CV_TRY(dense_to_csc<double>(h, h->tmp2.as<double>(), P, N));
  else if (x_type == CV_F32) CV_TRY(dense_to_csc<float>(h, h->tmp2.as<float>(), P, N));
  else CV_TRY(dense_to_csc<int32_t>(h, h->tmp2.as<int32_t>(), P, N));
  CV_TRY(counts_sums(h, x_type));
  CV_CUDA(h, cudaEventRecord(h->ev[1], h->stream));
  CV_CUDA(h, cudaEventSynchronize(h->ev[1]));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->st.ms_counts_layout = ms;
  cv_release(h->tmp2);
  return CV_OK;
}

// =========================================================================================
// fragments / totals / expectations
// =========================================================================================
extern "C" int cv_fragments(cv_handle* h, double* out_per_peak, double* out_per_sample,
                            double* out_total) {
  if (!h) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  int64_t P = h->P, N = h->N;
  if (out_per_peak || out_total) {
    std::vector<unsigned long long> tot((size_t)P + 1);
    CV_CUDA(h, cudaMemcpyAsync(tot.data(), h->peak_tot.p, (size_t)(P + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    CV_CUDA(h, cudaStreamSynchronize(h->stream));
    if (out_per_peak)
      for (int64_t i = 0; i < P; ++i) out_per_peak[i] = (double)tot[(size_t)i];
    if (out_total) *out_total = (double)tot[(size_t)P];
  }
  if (out_per_sample) {
    CV_CUDA(h, cudaMemcpyAsync(out_per_sample, h->fps.p, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
    CV_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return CV_OK;
}

extern "C" int cv_peak_totals_device(cv_handle* h, void** dptr, int64_t* n_elems) {
  if (!h || !dptr) return CV_ERR_INVALID;
  CV_TRY(cv_settle(h));
  h->replay_valid = false;                    // the totals are about to be rewritten by the caller
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  // keep the shard-local totals: the norm/group expectation variants are per-shard quantities
  CV_TRY(cv_ensure(h, h->peak_loc, (size_t)(h->P + 1) * 8));
  CV_CUDA(h, cudaMemcpyAsync(h->peak_loc.p, h->peak_tot.p, (size_t)(h->P + 1) * 8, cudaMemcpyDeviceToDevice, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->totals_committed = false;
  *dptr = h->peak_tot.p;
  if (n_elems) *n_elems = h->P + 1;
  return CV_OK;
}

extern "C" int cv_peak_totals_commit(cv_handle* h) {
  if (!h) return CV_ERR_INVALID;
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  h->totals_committed = true;
  return CV_OK;
}

// e_p = rowSum_p / total   (R/compute_deviations.R:424-425)
__global__ void k_expect_default(const unsigned long long* __restrict__ peak_tot, int P,
                                 double* __restrict__ e) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) e[i] = (double)peak_tot[i] / (double)peak_tot[P];
}

// norm / group variants: one pass over the nonzeros accumulating into acc[g*P + row] (fp64
// atomics; the reference sums in storage order, differences are O(1e-16) relative).
//   norm=1: add v / fps[col]   (:419-422, rowSums of column-normalised counts)
//   norm=0: add v              (:436-441), divided by the group's total afterwards
__global__ void __launch_bounds__(256)
k_expect_accum(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
               const uint32_t* __restrict__ val, int64_t nnz, int P, int N,
               const double* __restrict__ fps, const int32_t* __restrict__ group, int norm,
               double* __restrict__ acc, double* __restrict__ gtot) {
  int lane = threadIdx.x & 31;
  int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t base = warp_id * 32; base < nnz; base += warps_total * 32) {
    int64_t e = base + lane;
    bool active = e < nnz;
    int c = find_col_warp(colptr, N, active ? e : nnz - 1, nnz, active);
    if (active) {
      uint32_t v = val[e];
      if (v) {
        int g = group ? group[c] : 0;
        double add = norm ? (double)v / fps[c] : (double)v;
        atomicAdd(acc + (int64_t)g * P + rowidx[e], add);
        if (!norm) atomicAdd(gtot + g, (double)v);
      }
    }
  }
}

__global__ void k_expect_finish(const double* __restrict__ acc, const double* __restrict__ gtot,
                                int P, int G, int norm, double* __restrict__ e) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  double s = 0;
  for (int g = 0; g < G; ++g) {
    double v = acc[(int64_t)g * P + i];
    if (!norm) v /= gtot[g];
    s += v;
  }
  e[i] = s / (double)G;   // rowMeans over groups (:445); G == 1 without groups
}

// debugging aid (CV_SYNC_LAUNCH): every CV_LAUNCH waits for its kernel, a failure names it
int cv_sync_after_launch(cv_handle* h, const char* kernel) {
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaGetLastError();
    CV_FAIL(h, CV_ERR_CUDA, "kernel %s failed: %s", kernel, cudaGetErrorString(e));
  }
  return CV_OK;
}

// norm / group variants, first half: this engine's cells accumulated into acc[G*P] (+ gtot[G]) in
// tmp1; nothing is synchronised.  group: labels of THIS engine's cells (host).
int cv_expect_accumulate(cv_handle* h, int norm, const int32_t* group, int G) {
  cudaSetDevice(h->device);
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  const int64_t P = h->P, N = h->N;
  if (G < 1) CV_FAIL(h, CV_ERR_INVALID, "n_groups must be >= 1");
  if (group)
    for (int64_t i = 0; i < N; ++i)
      if (group[i] < 0 || group[i] >= G) CV_FAIL(h, CV_ERR_INVALID, "group label out of range at cell %lld", (long long)i);
  CV_TRY(cv_ensure(h, h->tmp1, (size_t)G * P * 8 + (size_t)G * 8));
  double* acc = h->tmp1.as<double>();
  double* gtot = acc + (size_t)G * P;
  CV_CUDA(h, cudaMemsetAsync(acc, 0, (size_t)G * P * 8 + (size_t)G * 8, h->stream));
  int32_t* dgroup = nullptr;
  if (group) {
    CV_TRY(cv_ensure(h, h->tmp2, (size_t)N * 4));
    dgroup = h->tmp2.as<int32_t>();
    CV_CUDA(h, cudaMemcpyAsync(dgroup, group, (size_t)N * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (h->nnz > 0)
    CV_LAUNCH(h, k_expect_accum, grid_for(h, h->nnz, 1024), 256, 0, h->colptr.as<int64_t>(),
              h->rowidx.as<int32_t>(), h->val.as<uint32_t>(), h->nnz, (int)P, (int)N,
              h->fps.as<double>(), dgroup, norm, acc, gtot);
  return CV_OK;
}

// second half: rowMeans over the groups of acc (this engine's tmp1, or a buffer summed over engines)
int cv_expect_finish(cv_handle* h, const double* acc, int G, int norm, double* out_e) {
  cudaSetDevice(h->device);
  const int64_t P = h->P;
  CV_TRY(cv_ensure(h, h->tmp0, (size_t)P * 8));
  CV_LAUNCH(h, k_expect_finish, (unsigned)((P + 255) / 256), 256, 0, acc, acc + (size_t)G * P, (int)P, G, norm,
            h->tmp0.as<double>());
  CV_CUDA(h, cudaMemcpyAsync(out_e, h->tmp0.p, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

extern "C" int cv_expectations(cv_handle* h, int norm, const int32_t* group, int n_groups,
                               double* out_e) {
  if (!h || !out_e) return CV_ERR_INVALID;
  cudaSetDevice(h->device);
  if (!h->have_counts) CV_FAIL(h, CV_ERR_STATE, "counts not set");
  int64_t P = h->P;
  if (!norm && !group) {
    if (!h->totals_committed)
      CV_FAIL(h, CV_ERR_STATE, "per-peak totals were handed out for a cross-shard reduction; call cv_peak_totals_commit first");
    CV_TRY(cv_ensure(h, h->tmp0, (size_t)P * 8));
    CV_LAUNCH(h, k_expect_default, (unsigned)((P + 255) / 256), 256, 0,
              h->peak_tot.as<unsigned long long>(), (int)P, h->tmp0.as<double>());
    CV_CUDA(h, cudaMemcpyAsync(out_e, h->tmp0.p, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
    CV_CUDA(h, cudaStreamSynchronize(h->stream));
    return CV_OK;
  }
  const int G = group ? n_groups : 1;
  CV_TRY(cv_expect_accumulate(h, norm, group, G));
  return cv_expect_finish(h, h->tmp1.as<double>(), G, norm, out_e);
}
