// cv_multi.cu -- one process, several GPUs behind the C ABI.
//
// The reference fans compute_deviations_core out over annotations with bplapply inside ONE R process
// (R/compute_deviations.R:387-391).  Here the fan-out is over cells: the caller hands the counts over
// as one or several CSC column blocks (a dgCMatrix's @p is int32, so an atlas is a list of blocks),
// the library cuts them into one contiguous cell shard per device (balanced by nonzeros), runs the
// single-device engine (cv_handle) of every shard on its own host thread, and performs the two
// cross-shard reductions of the path itself, over NVLink peer memory:
//
//   1. per-peak fragment totals  int64[P+1]  sum over shards   (k_sum_peers: every device reads its
//      peers' shard-local totals directly; feeds computeExpectations and the sampler)
//   2. per-annotation variability partials (n, mean, M2, n_nonfinite)[K] -> Chan merge in device
//      order (k_var_merge_peers on the first device)
//
// Without peer access between some pair of devices the peers' buffers are staged with
// cudaMemcpyPeerAsync first ("multi_p2p" = 0 forces that route).  Annotations, background matrix and
// expectation are replicated; results land in the caller's K x N matrices, each shard writing its
// own slab of columns.
#include <algorithm>
#include <thread>

#include "cv_internal.cuh"

#define CV_MAX_DEVICES 16

struct PeerPtrs { const void* p[CV_MAX_DEVICES]; };

struct cv_multi {
  int n = 0;
  cv_handle* h[CV_MAX_DEVICES] = {};
  int dev[CV_MAX_DEVICES] = {};
  bool p2p = false;               // every ordered pair of devices has peer access enabled
  int opt_p2p = 1;
  std::string err;
  int64_t P = 0, N = 0;
  int64_t bounds[CV_MAX_DEVICES + 1] = {};
  bool have_counts = false;
  int64_t K = 0;
  DevBuf stage[CV_MAX_DEVICES];   // per device: staged copies of the peers' buffers (no-P2P route)
  DevBuf sd0;                     // first device: merged row SDs
  DevBuf flag0;                   // first device: zero-peak flag
  DevBuf acc0;                    // first device: expectation accumulators summed over the shards
};

static thread_local std::string g_multi_create_err;

#define CVM_FAIL(m, code, ...)                 \
  do {                                         \
    char _b[640];                              \
    snprintf(_b, sizeof(_b), __VA_ARGS__);     \
    (m)->err = _b;                             \
    return (code);                             \
  } while (0)

// dst[i] = sum over devices of src_d[i]  (unsigned 64-bit; src pointers may be peer memory)
__global__ void __launch_bounds__(256)
k_sum_peers(unsigned long long* __restrict__ dst, PeerPtrs src, int n, int64_t len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  unsigned long long s = 0ull;
  for (int d = 0; d < n; ++d) s += reinterpret_cast<const unsigned long long*>(src.p[d])[i];
  dst[i] = s;
}

// parts_d[k][4] = (n_finite, mean, M2, n_nonfinite) of shard d -> sd[k] over all shards, merged in
// device order (Chan et al.); NaN when <= 1 finite value (row_sds(z, TRUE), src/utils.cpp:18-24)
__global__ void __launch_bounds__(128)
k_var_merge_peers(PeerPtrs parts, int n, int K, double* __restrict__ sd) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double cnt = 0.0, mean = 0.0, m2 = 0.0;
  for (int d = 0; d < n; ++d) {
    const double* q = reinterpret_cast<const double*>(parts.p[d]) + 4 * (int64_t)k;
    const double qn = q[0], qm = q[1], q2 = q[2];
    if (qn == 0.0) continue;
    if (cnt == 0.0) { cnt = qn; mean = qm; m2 = q2; continue; }
    const double tot = cnt + qn, dl = qm - mean;
    mean += dl * (qn / tot);
    m2 += q2 + dl * dl * (cnt * qn / tot);
    cnt = tot;
  }
  sd[k] = cnt > 1.0 ? sqrt(m2 / (cnt - 1.0)) : __longlong_as_double(0x7ff8000000000000LL);
}

// dst[i] = sum over devices of src_d[i]  (fp64, device order)
__global__ void __launch_bounds__(256)
k_sum_peers_f64(double* __restrict__ dst, PeerPtrs src, int n, int64_t len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int d = 0; d < n; ++d) s += reinterpret_cast<const double*>(src.p[d])[i];
  dst[i] = s;
}

__global__ void k_any_zero_u64(const unsigned long long* __restrict__ v, int64_t n, uint32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && v[i] == 0ull) atomicOr(flag, 1u);
}

// ---- helpers -------------------------------------------------------------------------------------------
static int fail_from(cv_multi* m, int i, int rc) {
  char b[700];
  snprintf(b, sizeof(b), "device %d: %s", m->dev[i], m->h[i]->err.c_str());
  m->err = m->n > 1 ? std::string(b) : m->h[i]->err;
  return rc;
}

// runs fn(i) for every device on its own host thread (inline for one device); first failure wins
template <typename F>
static int for_each_device(cv_multi* m, int n_active, F fn) {
  int rcs[CV_MAX_DEVICES] = {};
  if (n_active == 1) {
    rcs[0] = fn(0);
  } else {
    std::thread th[CV_MAX_DEVICES];
    for (int i = 0; i < n_active; ++i)
      th[i] = std::thread([&, i]() {
        cudaSetDevice(m->dev[i]);
        rcs[i] = fn(i);
      });
    for (int i = 0; i < n_active; ++i) th[i].join();
  }
  for (int i = 0; i < n_active; ++i)
    if (rcs[i] != CV_OK) return fail_from(m, i, rcs[i]);
  return CV_OK;
}

static int n_active(const cv_multi* m) {
  int a = 0;
  for (int i = 0; i < m->n; ++i)
    if (m->bounds[i + 1] > m->bounds[i]) a = i + 1;
  return std::max(a, 1);
}

// Pointers device i uses to read buffer `get(d)` of every device d: the peers' own memory with P2P,
// staged copies otherwise (enqueued on device i's stream).
template <typename G>
static int peer_view(cv_multi* m, int i, int na, size_t bytes, G get, PeerPtrs* out) {
  cv_handle* hi = m->h[i];
  const bool direct = m->p2p && m->opt_p2p;
  if (!direct) CV_TRY(cv_ensure(hi, m->stage[i], bytes * (size_t)na));
  for (int d = 0; d < na; ++d) {
    if (d == i || direct || m->dev[d] == m->dev[i]) { out->p[d] = get(d); continue; }
    char* dst = m->stage[i].as<char>() + (size_t)d * bytes;
    CV_CUDA(hi, cudaMemcpyPeerAsync(dst, m->dev[i], get(d), m->dev[d], bytes, hi->stream));
    out->p[d] = dst;
  }
  return CV_OK;
}

// reduction 1: every device ends up with the global per-peak totals (slot P = grand total)
static int reduce_peak_totals(cv_multi* m) {
  const int na = n_active(m);
  const int64_t len = m->P + 1;
  // shard-local totals are kept (the norm / group expectation variants are per-shard quantities)
  for (int i = 0; i < na; ++i) {
    cv_handle* h = m->h[i];
    cudaSetDevice(m->dev[i]);
    auto keep_local = [&]() -> int {
      CV_TRY(cv_ensure(h, h->peak_loc, (size_t)len * 8));
      CV_CUDA(h, cudaMemcpyAsync(h->peak_loc.p, h->peak_tot.p, (size_t)len * 8, cudaMemcpyDeviceToDevice, h->stream));
      return CV_OK;
    };
    const int rc = keep_local();
    if (rc != CV_OK) return fail_from(m, i, rc);
  }
  for (int i = 0; i < na; ++i) {
    cudaSetDevice(m->dev[i]);
    cudaError_t e = cudaStreamSynchronize(m->h[i]->stream);
    if (e != cudaSuccess) CVM_FAIL(m, CV_ERR_CUDA, "device %d: %s", m->dev[i], cudaGetErrorString(e));
  }
  if (na > 1) {
    for (int i = 0; i < na; ++i) {
      cv_handle* h = m->h[i];
      cudaSetDevice(m->dev[i]);
      auto sum = [&]() -> int {
        PeerPtrs src;
        CV_TRY(peer_view(m, i, na, (size_t)len * 8, [&](int d) { return (const void*)m->h[d]->peak_loc.p; }, &src));
        CV_LAUNCH(h, k_sum_peers, (unsigned)((len + 255) / 256), 256, 0, h->peak_tot.as<unsigned long long>(), src, na, len);
        return CV_OK;
      };
      const int rc = sum();
      if (rc != CV_OK) return fail_from(m, i, rc);
    }
    for (int i = 0; i < na; ++i) {
      cudaSetDevice(m->dev[i]);
      cudaError_t e = cudaStreamSynchronize(m->h[i]->stream);
      if (e != cudaSuccess) CVM_FAIL(m, CV_ERR_CUDA, "device %d: per-peak totals reduction failed: %s", m->dev[i], cudaGetErrorString(e));
    }
  }
  for (int i = 0; i < na; ++i) {
    m->h[i]->totals_committed = true;
    m->h[i]->replay_valid = false;
  }
  return CV_OK;
}

// min(rowSums(counts)) > 0 on the GLOBAL totals (R/compute_deviations.R:362-363)
static int check_zero_peaks_global(cv_multi* m) {
  cv_handle* h = m->h[0];
  cudaSetDevice(m->dev[0]);
  CV_TRY(cv_ensure(h, m->flag0, 64));
  CV_CUDA(h, cudaMemsetAsync(m->flag0.p, 0, 4, h->stream));
  CV_LAUNCH(h, k_any_zero_u64, (unsigned)((m->P + 255) / 256), 256, 0, h->peak_tot.as<unsigned long long>(), m->P,
            m->flag0.as<uint32_t>());
  uint32_t f = 0;
  CV_CUDA(h, cudaMemcpyAsync(&f, m->flag0.p, 4, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  if (f) CV_FAIL(h, CV_ERR_INVALID, "All peaks must have at least one fragment in one sample");
  return CV_OK;
}

// reduction 2: global row SD of z from the shards' partial tables
static int merge_variability(cv_multi* m, int64_t K, double* out_sd) {
  const int na = n_active(m);
  cv_handle* h = m->h[0];
  cudaSetDevice(m->dev[0]);
  CV_TRY(cv_ensure(h, m->sd0, (size_t)K * 8));
  PeerPtrs parts;
  CV_TRY(peer_view(m, 0, na, (size_t)K * 32, [&](int d) { return (const void*)m->h[d]->varfinal.p; }, &parts));
  CV_LAUNCH(h, k_var_merge_peers, (unsigned)((K + 127) / 128), 128, 0, parts, na, (int)K, m->sd0.as<double>());
  CV_CUDA(h, cudaMemcpyAsync(out_sd, m->sd0.p, (size_t)K * 8, cudaMemcpyDeviceToHost, h->stream));
  CV_CUDA(h, cudaStreamSynchronize(h->stream));
  return CV_OK;
}

// ---- column blocks -> one view per device ------------------------------------------------------------
struct Shards {
  std::vector<int64_t> gcp;                 // global colptr over all blocks
  std::vector<int64_t> block_e0;            // first global nonzero of each block
  CscView view[CV_MAX_DEVICES];
};

static int build_shards(cv_multi* m, int64_t P, const cv_csc_block* blocks, int n_blocks, Shards* S) {
  if (P <= 0 || P > 0x7fffffff) CVM_FAIL(m, CV_ERR_INVALID, "counts: P must be in [1, 2^31)");
  if (!blocks || n_blocks < 1) CVM_FAIL(m, CV_ERR_INVALID, "counts: no column blocks");
  int64_t N = 0;
  for (int b = 0; b < n_blocks; ++b) {
    if (blocks[b].n_cols < 0 || (blocks[b].n_cols > 0 && !blocks[b].colptr))
      CVM_FAIL(m, CV_ERR_INVALID, "counts: column block %d is empty or has no colptr", b);
    if (blocks[b].colptr_type != CV_I32 && blocks[b].colptr_type != CV_I64)
      CVM_FAIL(m, CV_ERR_INVALID, "counts: column block %d: colptr_type must be CV_I32 or CV_I64", b);
    if (blocks[b].x_type != blocks[0].x_type)
      CVM_FAIL(m, CV_ERR_INVALID, "counts: all column blocks must use the same value type");
    N += blocks[b].n_cols;
  }
  if (N <= 0) CVM_FAIL(m, CV_ERR_INVALID, "counts: empty matrix");
  const int x_type = blocks[0].x_type;
  size_t xb = x_type == CV_F64 ? 8 : (x_type == CV_F32 || x_type == CV_I32) ? 4 : x_type == CV_U8 ? 1 : 0;
  if (!xb) CVM_FAIL(m, CV_ERR_INVALID, "counts: unsupported x_type %d", x_type);
  S->gcp.assign((size_t)N + 1, 0);
  S->block_e0.assign((size_t)n_blocks + 1, 0);
  int64_t c = 0, e = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const cv_csc_block& bl = blocks[b];
    S->block_e0[(size_t)b] = e;
    int64_t first = 0, prev = 0;
    for (int64_t i = 0; i <= bl.n_cols; ++i) {
      const int64_t v = bl.n_cols == 0 ? 0 : (bl.colptr_type == CV_I32 ? (int64_t)((const int32_t*)bl.colptr)[i] : ((const int64_t*)bl.colptr)[i]);
      if (i == 0) { first = v; prev = v; if (bl.n_cols == 0) break; continue; }
      if (v < prev) CVM_FAIL(m, CV_ERR_INVALID, "counts: colptr of block %d is not monotone at column %lld", b, (long long)(i - 1));
      S->gcp[(size_t)(c + i)] = e + (v - first);
      prev = v;
    }
    if (first != 0) CVM_FAIL(m, CV_ERR_INVALID, "counts: colptr of block %d must start at 0", b);
    const int64_t nnz_b = prev - first;
    if (nnz_b > 0 && (!bl.rowidx || !bl.x)) CVM_FAIL(m, CV_ERR_INVALID, "counts: block %d: rowidx / x are NULL", b);
    c += bl.n_cols;
    e += nnz_b;
  }
  S->block_e0[(size_t)n_blocks] = e;
  const int64_t nnz = e;
  // contiguous cell shards balanced by nonzeros
  const int n = (int)std::min<int64_t>(m->n, N);
  m->bounds[0] = 0;
  for (int i = 1; i < n; ++i) {
    const double target = (double)nnz * i / n;
    int64_t cut = std::lower_bound(S->gcp.begin(), S->gcp.end(), (int64_t)target) - S->gcp.begin();
    cut = std::min<int64_t>(std::max<int64_t>(cut, m->bounds[i - 1] + 1), N - (n - i));
    m->bounds[i] = cut;
  }
  for (int i = n; i <= m->n; ++i) m->bounds[i] = N;
  if (n < 1) m->bounds[1] = N;
  for (int i = 0; i < m->n; ++i) {
    CscView& v = S->view[i];
    const int64_t c0 = m->bounds[i], c1 = m->bounds[i + 1];
    v.P = P; v.N = c1 - c0; v.x_type = x_type; v.xb = xb;
    v.cp.resize((size_t)(c1 - c0) + 1);
    const int64_t e_lo = S->gcp[(size_t)c0], e_hi = S->gcp[(size_t)c1];
    for (int64_t j = c0; j <= c1; ++j) v.cp[(size_t)(j - c0)] = S->gcp[(size_t)j] - e_lo;
    v.segs.clear();
    for (int b = 0; b < n_blocks; ++b) {
      const int64_t b0 = S->block_e0[(size_t)b], b1 = S->block_e0[(size_t)b + 1];
      const int64_t a = std::max(e_lo, b0), z = std::min(e_hi, b1);
      if (z <= a) continue;
      v.segs.push_back(CscSeg{a - e_lo, z - e_lo, blocks[b].rowidx + (a - b0), (const char*)blocks[b].x + (size_t)(a - b0) * xb});
    }
  }
  m->P = P;
  m->N = N;
  return CV_OK;
}

// ---- ABI ------------------------------------------------------------------------------------------------
extern "C" int cv_create_multi(const int* devices, int n_dev, uint64_t seed, cv_multi** out) {
  if (!out) return CV_ERR_INVALID;
  *out = nullptr;
  if (n_dev < 1 || n_dev > CV_MAX_DEVICES) {
    g_multi_create_err = "n_dev must be in [1, 16]";
    return CV_ERR_INVALID;
  }
  cv_multi* m = new cv_multi();
  m->n = n_dev;
  for (int i = 0; i < n_dev; ++i) {
    m->dev[i] = devices ? devices[i] : i;
    int rc = cv_create(m->dev[i], seed, &m->h[i]);
    if (rc != CV_OK) {
      g_multi_create_err = std::string(cv_last_error(nullptr));
      for (int q = 0; q < i; ++q) cv_destroy(m->h[q]);
      delete m;
      return rc;
    }
  }
  // peer access between every ordered pair (NVLink / NVSwitch on a B200 box)
  m->p2p = n_dev > 1;
  for (int i = 0; i < n_dev && m->p2p; ++i)
    for (int j = 0; j < n_dev; ++j) {
      if (m->dev[i] == m->dev[j]) continue;        // several shards on one device (tests): plain device memory
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]) != cudaSuccess || !can) { m->p2p = false; break; }
      cudaSetDevice(m->dev[i]);
      cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { m->p2p = false; break; }
      cudaGetLastError();
    }
  cudaGetLastError();
  m->bounds[0] = 0;
  *out = m;
  return CV_OK;
}

extern "C" void cv_destroy_multi(cv_multi* m) {
  if (!m) return;
  for (int i = 0; i < m->n; ++i) {
    cudaSetDevice(m->dev[i]);
    cv_release(m->stage[i]);
    if (i == 0) { cv_release(m->sd0); cv_release(m->flag0); cv_release(m->acc0); }
    cv_destroy(m->h[i]);
  }
  delete m;
}

extern "C" const char* cv_multi_last_error(const cv_multi* m) {
  return m ? m->err.c_str() : g_multi_create_err.c_str();
}

extern "C" int cv_multi_n_devices(const cv_multi* m) { return m ? m->n : 0; }

extern "C" cv_handle* cv_multi_handle(cv_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->h[i] : nullptr; }

extern "C" int cv_multi_peer_access(const cv_multi* m) { return m ? (m->p2p && m->opt_p2p ? 1 : 0) : 0; }

extern "C" int cv_multi_set_seed(cv_multi* m, uint64_t seed) {
  if (!m) return CV_ERR_INVALID;
  for (int i = 0; i < m->n; ++i) cv_set_seed(m->h[i], seed);
  return CV_OK;
}

extern "C" int cv_multi_set_option(cv_multi* m, const char* key, int64_t value) {
  if (!m || !key) return CV_ERR_INVALID;
  if (!strcmp(key, "multi_p2p")) { m->opt_p2p = value != 0; return CV_OK; }
  for (int i = 0; i < m->n; ++i) {
    cudaSetDevice(m->dev[i]);
    int rc = cv_set_option(m->h[i], key, value);
    if (rc != CV_OK) return fail_from(m, i, rc);
  }
  return CV_OK;
}

extern "C" int cv_multi_shards(const cv_multi* m, int64_t* bounds) {
  if (!m || !bounds) return CV_ERR_INVALID;
  for (int i = 0; i <= m->n; ++i) bounds[i] = m->bounds[i];
  return CV_OK;
}

extern "C" int cv_multi_set_counts(cv_multi* m, int64_t P, const cv_csc_block* blocks, int n_blocks) {
  if (!m) return CV_ERR_INVALID;
  m->have_counts = false;
  Shards S;
  CV_TRY(build_shards(m, P, blocks, n_blocks, &S));
  const int na = n_active(m);
  CV_TRY(for_each_device(m, na, [&](int i) { return cv_set_counts_view(m->h[i], S.view[i]); }));
  CV_TRY(reduce_peak_totals(m));
  m->have_counts = true;
  return CV_OK;
}

// A base matrix (column-major P x N): every device takes a contiguous slab of columns and converts it
// to CSC itself (cv_set_counts_dense); shards are equal numbers of cells, the nonzeros are not known
// before the conversion.
extern "C" int cv_multi_set_counts_dense(cv_multi* m, int64_t P, int64_t N, const void* x, int x_type) {
  if (!m) return CV_ERR_INVALID;
  m->have_counts = false;
  if (P <= 0 || N <= 0 || !x) CVM_FAIL(m, CV_ERR_INVALID, "counts: empty matrix");
  const size_t xb = x_type == CV_F64 ? 8 : (x_type == CV_F32 || x_type == CV_I32) ? 4 : 0;
  if (!xb) CVM_FAIL(m, CV_ERR_INVALID, "dense counts: x_type must be CV_F64, CV_F32 or CV_I32");
  const int n = (int)std::min<int64_t>(m->n, N);
  for (int i = 0; i <= m->n; ++i) m->bounds[i] = i < n ? N * i / n : N;
  m->P = P;
  m->N = N;
  const int na = n_active(m);
  CV_TRY(for_each_device(m, na, [&](int i) {
    const int64_t c0 = m->bounds[i], c1 = m->bounds[i + 1];
    return cv_set_counts_dense(m->h[i], P, c1 - c0, (const char*)x + (size_t)P * (size_t)c0 * xb, x_type);
  }));
  CV_TRY(reduce_peak_totals(m));
  m->have_counts = true;
  return CV_OK;
}

extern "C" int cv_multi_fragments(cv_multi* m, double* out_per_peak, double* out_per_sample, double* out_total) {
  if (!m) return CV_ERR_INVALID;
  if (!m->have_counts) CVM_FAIL(m, CV_ERR_STATE, "counts not set");
  const int na = n_active(m);
  if (out_per_peak || out_total) {
    int rc = cv_fragments(m->h[0], out_per_peak, nullptr, out_total);       // global after reduction 1
    if (rc != CV_OK) return fail_from(m, 0, rc);
  }
  if (out_per_sample)
    for (int i = 0; i < na; ++i) {
      int rc = cv_fragments(m->h[i], nullptr, out_per_sample + m->bounds[i], nullptr);
      if (rc != CV_OK) return fail_from(m, i, rc);
    }
  return CV_OK;
}

// computeExpectations, all three variants (R/compute_deviations.R:412-447).  The default is a function
// of the reduced totals; norm / group are sums over cells of per-cell terms: every device accumulates
// its shard (group: labels of ALL cells, the device reads its slice), the first device adds the
// accumulators in device order and takes the row means.
extern "C" int cv_multi_expectations(cv_multi* m, int norm, const int32_t* group, int n_groups, double* out_e) {
  if (!m || !out_e) return CV_ERR_INVALID;
  if (!m->have_counts) CVM_FAIL(m, CV_ERR_STATE, "counts not set");
  const int na = n_active(m);
  if ((!norm && !group) || na == 1) {
    int rc = cv_expectations(m->h[0], norm, group, n_groups, out_e);
    return rc == CV_OK ? CV_OK : fail_from(m, 0, rc);
  }
  const int G = group ? n_groups : 1;
  if (G < 1) CVM_FAIL(m, CV_ERR_INVALID, "n_groups must be >= 1");
  CV_TRY(for_each_device(m, na, [&](int i) -> int {
    cv_handle* h = m->h[i];
    CV_TRY(cv_expect_accumulate(h, norm, group ? group + m->bounds[i] : nullptr, G));
    CV_CUDA(h, cudaStreamSynchronize(h->stream));
    return CV_OK;
  }));
  cv_handle* h = m->h[0];
  cudaSetDevice(m->dev[0]);
  const int64_t len = (int64_t)G * m->P + G;
  auto sum = [&]() -> int {
    CV_TRY(cv_ensure(h, m->acc0, (size_t)len * 8));
    PeerPtrs src;
    CV_TRY(peer_view(m, 0, na, (size_t)len * 8, [&](int d) { return (const void*)m->h[d]->tmp1.p; }, &src));
    CV_LAUNCH(h, k_sum_peers_f64, (unsigned)((len + 255) / 256), 256, 0, m->acc0.as<double>(), src, na, len);
    return cv_expect_finish(h, m->acc0.as<double>(), G, norm, out_e);
  };
  const int rc = sum();
  return rc == CV_OK ? CV_OK : fail_from(m, 0, rc);
}

extern "C" int cv_multi_background_peaks(cv_multi* m, const double* bias, int niter, double w, int bs, int32_t* out_bg) {
  if (!m) return CV_ERR_INVALID;
  if (!m->have_counts) CVM_FAIL(m, CV_ERR_STATE, "counts not set");
  // cell independent: a pure function of (global totals, bias, seed) -- one device draws it
  int rc = cv_background_peaks(m->h[0], bias, niter, w, bs, out_bg, nullptr, nullptr, nullptr, nullptr);
  return rc == CV_OK ? CV_OK : fail_from(m, 0, rc);
}

static int finish_call(cv_multi* m, int64_t K, double* out_variability) {
  if (out_variability) {
    int rc = merge_variability(m, K, out_variability);
    if (rc != CV_OK) return fail_from(m, 0, rc);
  }
  return CV_OK;
}

extern "C" int cv_multi_deviations(cv_multi* m, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                                   int index_base, const int32_t* bg, int B, const double* expectation,
                                   double* out_dev, double* out_z, double* out_matches, double* out_overlap,
                                   double* out_variability) {
  if (!m) return CV_ERR_INVALID;
  if (!m->have_counts) CVM_FAIL(m, CV_ERR_STATE, "counts not set");
  if (!expectation) CVM_FAIL(m, CV_ERR_INVALID, "expectation is NULL");
  const int na = n_active(m);
  // the totals every device holds are global: the zero-peak check of each device's call is the global one
  CV_TRY(for_each_device(m, na, [&](int i) {
    cv_handle* h = m->h[i];
    const int64_t c0 = m->bounds[i];
    return cv_deviations(h, K, annoptr, annoidx, index_base, bg, B, expectation, out_dev ? out_dev + K * c0 : nullptr,
                         out_z ? out_z + K * c0 : nullptr, i == 0 ? out_matches : nullptr, i == 0 ? out_overlap : nullptr,
                         nullptr, nullptr, nullptr);
  }));
  return finish_call(m, K, out_variability);
}

extern "C" int cv_multi_deviations_csc(cv_multi* m, int64_t P, const cv_csc_block* blocks, int n_blocks, int64_t K,
                                       const int64_t* annoptr, const int32_t* annoidx, int index_base,
                                       const int32_t* bg, int B, const double* expectation, double* out_dev,
                                       double* out_z, double* out_matches, double* out_overlap,
                                       double* out_variability) {
  if (!m) return CV_ERR_INVALID;
  if (!expectation) {
    // computeExpectations' default needs the global totals before any deviation: plain sequence
    std::vector<double> e((size_t)std::max<int64_t>(P, 1));
    CV_TRY(cv_multi_set_counts(m, P, blocks, n_blocks));
    CV_TRY(cv_multi_expectations(m, 0, nullptr, 0, e.data()));
    return cv_multi_deviations(m, K, annoptr, annoidx, index_base, bg, B, e.data(), out_dev, out_z, out_matches,
                               out_overlap, out_variability);
  }
  m->have_counts = false;
  Shards S;
  CV_TRY(build_shards(m, P, blocks, n_blocks, &S));
  const int na = n_active(m);
  // every device streams its own shard (H2D of the counts, kernels, D2H of its slab of the results
  // overlapped); a shard may hold no fragment of some peak, the check runs on the reduced totals
  int rc = for_each_device(m, na, [&](int i) {
    cv_handle* h = m->h[i];
    const int64_t c0 = m->bounds[i];
    const int keep = h->opt_check_zero;
    h->opt_check_zero = 0;
    const int r = cv_deviations_view(h, S.view[i], K, annoptr, annoidx, index_base, bg, B, expectation,
                                     out_dev ? out_dev + K * c0 : nullptr, out_z ? out_z + K * c0 : nullptr,
                                     i == 0 ? out_matches : nullptr, i == 0 ? out_overlap : nullptr, nullptr);
    h->opt_check_zero = keep;
    return r;
  });
  if (rc != CV_OK) return rc;
  CV_TRY(reduce_peak_totals(m));
  m->have_counts = true;
  if (m->h[0]->opt_check_zero) {
    rc = check_zero_peaks_global(m);
    if (rc != CV_OK) return fail_from(m, 0, rc);
  }
  return finish_call(m, K, out_variability);
}

extern "C" int cv_multi_stats(cv_multi* m, int i, cv_stage_stats* out) {
  if (!m || i < 0 || i >= m->n) return CV_ERR_INVALID;
  return cv_stats(m->h[i], out);
}
