// Internal declarations shared by the translation units of libchromvar_b200.so.
// Nothing here crosses the C ABI (include/chromvar_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/chromvar_b200.h"

#define CV_NUM_SMS_FALLBACK 148
#define CV_PINNED_WORDS 4096

// ---- error plumbing -------------------------------------------------------------------
#define CV_FAIL(h, code, ...)                                \
  do {                                                       \
    char _b[512];                                            \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                   \
    (h)->err = _b;                                           \
    return (code);                                           \
  } while (0)

#define CV_CUDA(h, expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      cudaGetLastError();                                                              \
      CV_FAIL(h, (_e == cudaErrorMemoryAllocation ? CV_ERR_OOM : CV_ERR_CUDA),         \
              "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
              cudaGetErrorString(_e));                                                 \
    }                                                                                  \
  } while (0)

#define CV_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != CV_OK) return _rc; \
  } while (0)

// kernel launch wrapper: counts launches (cv_stats.kernel_launches) and checks the launch
#define CV_LAUNCH(h, kernel, grid, block, smem, ...)                         \
  do {                                                                       \
    kernel<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);           \
    (h)->launches++;                                                         \
    CV_CUDA(h, cudaGetLastError());                                          \
    if ((h)->opt_sync_launch) CV_TRY(cv_sync_after_launch(h, #kernel));      \
  } while (0)

// ---- grow-only device buffer -------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct cv_handle {
  int device = 0;
  int num_sms = CV_NUM_SMS_FALLBACK;
  int max_smem_optin = 0;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;      // compute
  cudaStream_t s_up = nullptr;        // host -> device copies of the streamed call
  cudaStream_t s_dn = nullptr;        // device -> host copies overlapping the next chunk
  cudaStream_t s_aux = nullptr;       // per-annotation tables beside the operand build
  bool aux_pending = false;           // ev[7] (overlap / matches ready) still has to be joined
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_aux[4] = {};         // operand build: inputs ready, bit columns, table, overlap
  uint32_t* pinned = nullptr;         // CV_PINNED_WORDS of page-locked host memory (flag snapshots)
  struct cv_narrower* narrower = nullptr;   // host threads converting x to bytes ahead of the uploads (cv_hostnarrow.cpp)
  uint8_t* stage8 = nullptr;          // page-locked staging of the narrowed values
  size_t stage8_cap = 0;
  int32_t* stage_ri = nullptr;        // page-locked staging of the row indices (pageable caller memory)
  size_t stage_ri_cap = 0;            // in elements
  char* stage_out = nullptr;          // page-locked landing buffer of the results (pageable caller memory)
  size_t stage_out_cap = 0;
  // cv_deviations_compute_async: a resident call whose host-side verdicts (index checks, exactness
  // bounds, e2m1 feasibility) are known from the last fully checked call on the SAME resident inputs
  // is enqueued without any host synchronisation; cv_deviations_wait settles it
  bool replay_valid = false;          // verdict below belongs to the current resident inputs and options
  uint32_t replay_hf[16] = {};        // flag words the checked call read after the operand build
  bool replaying = false;             // deviations_run is enqueueing a replay
  bool replay_gather = false;         // the checked call took the few-cells path (cv_dense_gather.cu)
  bool async_pending = false;         // a replay is in flight: flags / timings not collected yet
  int opt_tail_chunk = 1;             // "tail_chunk": streamed call ends with a short fourth cell chunk
  int opt_host_narrow = -1;           // "host_narrow": -1 (default) host threads stage / narrow PAGEABLE caller memory, 0 never, n >= 1 always
  // timed spans of the current call (CUDA events on `stream`), summed per category afterwards
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { cudaEvent_t a, b; int cat; };
  std::vector<Span> spans;
  std::string err;
  int64_t launches = 0;

  // ---- counts (raw input format, device resident) ----
  int64_t P = 0, N = 0, nnz = 0;
  bool have_counts = false;
  bool totals_committed = false;   // per-peak totals are global (single shard or after commit)
  DevBuf colptr;     // int64[N+1]
  DevBuf rowidx;     // int32[nnz]
  DevBuf val;        // uint32[nnz]   (validated integer counts)
  DevBuf xraw;       // staging for the uploaded x / dense input
  DevBuf peak_tot;   // int64[P+1], slot P = grand total (this shard, or global after commit)
  DevBuf peak_loc;   // int64[P+1] local copy kept when totals get all-reduced
  DevBuf cell_tot;   // int64[N]
  DevBuf fps;        // double[N]  fragments per sample
  DevBuf flags;      // uint32[16]: [0..7] counts validation (K1), [8..15] deviations call (CV_DFLAGS)
  uint32_t max_val = 0;
  uint64_t max_cell_tot = 0;

  // ---- cell-tiled row layout of the counts (derived) ----
  bool layout_valid = false;
  int W = 0, T = 0;  // cells per tile, number of tiles
  int64_t n_entries = 0;
  DevBuf tcnt;       // uint32[T*P]   per (tile,row) entry counts, reused as fill cursor
  DevBuf rowptr;     // uint32[T*P+1]
  DevBuf entries;    // uint32[n_entries] : cell_in_tile << 20 | value
  DevBuf scan_tmp;

  // ---- annotations / background / expectation (raw + derived) ----
  bool have_anno = false;
  int64_t K = 0, nnzA = 0;
  int B = 0;
  std::vector<int64_t> h_annoptr;
  std::vector<int32_t> h_order;
  DevBuf annoptr;    // int64[K+1]
  DevBuf anno_raw;   // int32[nnzA] as uploaded (index_base)
  DevBuf annoidx;    // int32[nnzA] 0-based
  DevBuf bg_raw;     // int32[P*B] as uploaded (col-major, index_base)
  int index_base = 1;
  DevBuf bgT;        // int32[P][B] 0-based
  DevBuf expect;     // double[P]
  DevBuf etab;       // double[K][B+1]
  DevBuf overlap;    // double[K]
  DevBuf matches;    // double[K]
  DevBuf mark;       // int32[grid][P] scratch for the overlap count
  DevBuf order;      // int32[K] annotations by decreasing set size
  DevBuf work_ctr;   // uint32 dynamic work counters

  // ---- results (device) ----
  bool have_results = false;
  DevBuf dev_rm, z_rm;     // double[K][N] row-major
  DevBuf dev_cm, z_cm;     // double K x N col-major (ABI layout)
  DevBuf varpart;          // double[K][T][4]
  DevBuf varfinal;         // double[K][4]
  DevBuf variab;           // double[K]
  DevBuf obs, samp;        // int64 parity outputs
  bool kept_sums = false;

  // dense tensor-core path (cv_dense.cu)
  DevBuf dn_M;       // u8 [n_groups*256][Ppad] stacked annotation x iteration operand
  DevBuf dn_C;       // u8 [chunk cells][Ppad] densified counts of the current cell chunk / digit plane
  DevBuf dn_S;       // int64 [K][B+1][chunk cells] exact sums when counts need several u8 planes
  DevBuf dn_abits;   // u32 [ceil(K/32)][P] annotation membership bits
  DevBuf dn_abitsT;  // the same bits peak-major: u32 [ceil(K/1024)][P][32] (a peak's 32 words = one 128-byte line)
  DevBuf dn_invptr;  // u32 [B*P + 1] inverse background map: sources of (iteration, peak)
  DevBuf dn_invsrc;  // i32 [B*P]
  DevBuf dn_E2;      // f64 [P][B+1] e[p], e[bg[p, j]]
  DevBuf dn_ov;      // u64 [K] background overlap counts
  DevBuf dn_winv;    // f64 [K] 2^-F_k, then u32 [K][B+1] W[k][j] = round(2^F_k / E[k][j])
  // e2m1 variant (cv_dense_fp4.cu); dn_M / dn_C then hold the nibble-packed operands
  DevBuf f4_Tq;      // u32 [P] terms of the largest operand entry of peak q (0 = none outside S)
  DevBuf f4_Sq;      // u32 [P] terms of the largest count of peak q in the current cell chunk
  DevBuf f4_ncol;    // u32 [P] overflow columns of peak q in the current chunk
  DevBuf f4_off;     // u32 [P+1] exclusive scan of f4_ncol
  DevBuf f4_colq;    // i32 [Hcap] peak of overflow column x
  DevBuf f4_colt;    // u32 [Hcap] multiplicity term of overflow column x
  DevBuf f4_info;    // u32 [16]: k-blocks and overflow columns of the current chunk
  DevBuf f4_sync;    // u32 [65536] producer lockstep counters of one GEMM launch
  int opt_lockstep = 1;        // e2m1 GEMM: TMA producers of all CTAs start each sub-tile together
  int opt_fp4 = -1;            // -1 / 1 use the e2m1 kernel when the problem allows it, 0 never
  bool f4_veto = false;        // this problem turned out not to fit the e2m1 path: retry with u8
  bool dn_fixed_ok = true;
  uint64_t dn_maxmult = 1;     // bound on the entries of M (set by cv_dense_check)
  int opt_epilogue = 0;        // dense fused epilogue: 0 by tile length, 1 fixed point / fp32, 2 fp64
  int opt_rows_scatter = 0;   // build the operand rows with the per-annotation scatter kernel only
  double dn_flops = 0;
  int opt_path = 0;  // 0 auto, 1 force row-gather SpMM, 2 force dense tcgen05
  int opt_pair = -1; // dense GEMM kernel: 1 CTA pair (tcgen05 cta_group::2), 0 single CTA, -1 by context
  int64_t opt_chunk = 0;       // dense path: cap on cells per GEMM launch (0 = memory budget)
  int opt_band = 0;            // dense path: groups per raster band (0 = default)
  int opt_debug_no_tma = 0;    // experiment only: GEMM without operand loads (garbage results)
  int opt_check_zero = 1;      // reject peaks without fragments (off for shards of a larger matrix)
  int opt_gather = -1;         // "small_n_gather": -1 few-cells path when it applies and the stacked operand would be large, 0 never, 1 whenever it applies
  bool cols_dup = false;       // K1: some column of the counts lists a peak twice
  int opt_knn_insertion = 0;   // "knn_insertion": the small-window kNN kernel at every window (comparison / tests)
  int opt_sync_launch = 0;     // CV_SYNC_LAUNCH in the environment at cv_create: wait after every kernel, name the one that failed
  int opt_trace = 0;           // CV_TRACE in the environment at cv_create: timeline of the streamed call on stderr
  double sparse_rate = 2.7e8;  // row-gather kernel, entry updates per ms: measured by the last sparse run (choose_path)

  // background sampler / compat-routine buffers (cv_background.cu, cv_compat.cu)
  DevBuf bk[16];

  // generic scratch
  DevBuf tmp0, tmp1, tmp2;

  cv_stage_stats st = {};
};

// One shard's counts as the host holds them (cv_counts.cu): colptr rebased to the shard, the
// nonzeros [e0, e1) of segment i live at rowidx[e - e0], x[(e - e0) * xb]
struct CscSeg { int64_t e0, e1; const int32_t* rowidx; const char* x; };
struct CscView {
  int64_t P = 0, N = 0;
  std::vector<int64_t> cp;
  std::vector<CscSeg> segs;
  int x_type = 0;
  size_t xb = 0;
};
int cv_csc_view_init(cv_handle* h, CscView* v, int64_t P, int64_t N, const void* colptr, int colptr_type,
                     const int32_t* rowidx, const void* x, int x_type);
int cv_csc_view_check(cv_handle* h, const CscView& v);
int cv_csc_view_copy(cv_handle* h, const CscView& v, int64_t e0, int64_t e1, int32_t* d_rowidx, char* d_x, cudaStream_t st);
int cv_set_counts_view(cv_handle* h, const CscView& v);
int cv_sync_after_launch(cv_handle* h, const char* kernel);
int cv_expect_accumulate(cv_handle* h, int norm, const int32_t* group, int G);
int cv_expect_finish(cv_handle* h, const double* acc, int G, int norm, double* out_e);
// cv_deviations.cu: cv_deviations_csc on a view (expectation may be NULL)
int cv_deviations_view(cv_handle* h, const CscView& v, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                       int index_base, const int32_t* bg, int B, const double* expectation, double* out_dev,
                       double* out_z, double* out_matches, double* out_overlap, double* out_variability);

int cv_ensure(cv_handle* h, DevBuf& b, size_t bytes);
void cv_release(DevBuf& b);
int cv_sync_check(cv_handle* h);

// cv_counts.cu
int cv_build_layout(cv_handle* h, int W);
// cv_background.cu: shared sampler back half (bin CDF rows + Philox draws)
// cv_hostnarrow.cpp
struct cv_narrower;
cv_narrower* cv_narrower_create(int threads);
int cv_narrower_threads(const cv_narrower* nw);
void cv_narrower_destroy(cv_narrower* nw);
void cv_narrower_start(cv_narrower* nw, const void* x, int x_type, int64_t n, uint8_t* out, const int32_t* ri,
                       int32_t* ri_out);
void cv_narrower_copy(cv_narrower* nw, void* dst, const void* src, size_t n);
void cv_narrower_copies_wait(cv_narrower* nw);
int cv_narrower_wait(cv_narrower* nw, int64_t e0, int64_t e1);
void cv_narrower_finish(cv_narrower* nw);
int cv_settle(cv_handle* h);            // cv_deviations.cu: waits for a pending replay and collects its verdict
int cv_whiten(cv_handle* h, const double* bias, int log_intensity, double* mm);   // mm == NULL: no host synchronisation
int cv_whiten_verdict(cv_handle* h, int log_intensity);
int cv_sample_background(cv_handle* h, int64_t P, const int32_t* d_memb0, int64_t nb,
                         const double* d_binp_colmajor, const double* d_bin_xy, double w,
                         int niter, int32_t* host_out, double* host_binprob, double* host_density);
// flag words of a deviations call live behind the K1 words
#define CV_DFLAGS 8
// timed-span categories (cv_stage_stats)
enum { CV_SPAN_LAYOUT = 0, CV_SPAN_ANNO = 1, CV_SPAN_CONTRACT = 2, CV_SPAN_FINALIZE = 3, CV_SPAN_K1 = 4, CV_SPAN_N = 5 };
void cv_span_begin(cv_handle* h, int cat);
void cv_span_end(cv_handle* h);
void cv_spans_reset(cv_handle* h);
void cv_spans_collect(cv_handle* h, double* ms_by_cat /*CV_SPAN_N*/);   // after a stream sync

// cv_dense.cu: shape of the dense tensor-core path for the current problem
struct DensePlan {
  int I = 0, G = 0, n_groups = 0;
  int planes = 1;            // u8 digit planes of the counts (1 when max count <= 255)
  int pair = 1;              // CTA-pair kernel
  int64_t Ppad = 0, rowsM = 0;
  int64_t chunk_cells = 0;   // cells per GEMM launch, multiple of 256
  int64_t m_bytes = 0, c_bytes = 0, sums_bytes = 0;
  uint64_t cell_tot_bound = ~0ull;   // largest cell total of the cells a launch covers (exactness of fp32 sums)
  // e2m1 variant: annotations per 256- / 208-row sub-tile, super-groups of 464 rows, peak columns
  // padded to 256, reserved overflow columns, bytes per operand row, operand sizes of both variants
  int fp4 = 0, Ga = 0, Gb = 0, n_sg = 0;
  int64_t Ppad4 = 0, Hcap = 0, rowb = 0, rowsM4 = 0;
  int64_t m_bytes_u8 = 0, c_bytes_u8 = 0;
};
int cv_f4_plan(cv_handle* h, DensePlan* pl);
void cv_dense_plan_set_fp4(DensePlan* pl, int on);
int cv_f4_build_rows(cv_handle* h, const DensePlan& pl);
struct DenseParams;
int cv_f4_chunk(cv_handle* h, const DensePlan& pl, const DenseParams& prm, int64_t c0, int64_t c1, int64_t rows,
                int fixed_point_epilogue, int check_now);
int cv_dense_plan(cv_handle* h, uint32_t max_val, int64_t want_chunk, DensePlan* pl, std::string* why);
double cv_dense_cost_ms(cv_handle* h, const DensePlan& pl);
int cv_dense_prepare(cv_handle* h, const DensePlan& pl, int scatter_rows);
int cv_dense_chunk(cv_handle* h, const DensePlan& pl, int64_t c0, int64_t c1, int keep_sums);
int cv_dense_check(cv_handle* h, const DensePlan& pl, const uint32_t* hflags, int with_counts, int* soft_fail);
// cv_dense_build.cu
int cv_dense_build_rows(cv_handle* h, const DensePlan& pl, int scatter, int tables_only = 0);
int cv_dense_finalize_sums(cv_handle* h, const DenseParams& prm);
// cv_dense_gather.cu: the contraction for few cells (gathered counts rows, no stacked operand)
int cv_gather_applicable(const cv_handle* h);
int cv_gather_etab_in_gemm(const cv_handle* h);
int cv_gather_prepare(cv_handle* h);
int cv_gather_run(cv_handle* h, int keep_sums);
int cv_dense_rebuild_rows_u8(cv_handle* h, const DensePlan& pl);
int cv_dense_build_cells(cv_handle* h, const DensePlan& pl, int64_t c0, int64_t rows, int shift);
int cv_exclusive_scan_u32_async(cv_handle* h, const uint32_t* in, uint32_t* out, int64_t n);
// cv_counts.cu: K1 in pieces so the streamed call can run it per uploaded cell chunk
int cv_counts_begin(cv_handle* h);
int cv_counts_range(cv_handle* h, int x_type, int64_t e0, int64_t e1, int64_t c0, int64_t c1);
int cv_counts_finish_async(cv_handle* h);
int cv_counts_verdict(cv_handle* h, const uint32_t* hflags);
// cv_deviations.cu
int cv_choose_tile(cv_handle* h, int B, int acc_bytes, int* W_out);

// ---- small device helpers -----------------------------------------------------------------
__device__ __forceinline__ double cv_na_real() {
  return __longlong_as_double((long long)CV_NA_REAL_BITS);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
