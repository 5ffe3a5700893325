/*
 * chromvar_b200.h -- C ABI of libchromvar_b200.so, the B200-native deviation engine.
 *
 * This is the drop-in boundary for chromVAR's hot path
 *   computeExpectations -> getBackgroundPeaks -> compute_deviations_core -> computeVariability.
 * Plain C: pointers and sizes only, no R / Rcpp / torch types.  Every entry point cites the
 * reference interface it replaces (paths relative to the chromVAR 1.5.0 tree).  The reference
 * binds native code through `.Call` into registered routines (src/RcppExports.cpp:59-137,
 * R/RcppExports.R:20-38); INTEGRATION.md shows the `.Call` glue a maintainer adds for each
 * function below.
 *
 * Conventions
 *   - Matrices are column-major (R layout).  z / deviations are K x N (annotation = row).
 *   - Peak indices crossing the ABI carry an explicit `index_base` (R passes 1).
 *   - Caller allocates every output; the library keeps no host pointer after return.
 *   - Every call returns 0 (CV_OK) or an error code; cv_last_error() gives the message.  Nothing
 *     throws or longjmps across the ABI.  There is no CPU fallback: without a usable sm_100
 *     device cv_create fails.
 *   - NA: entries the reference sets to NA (R/compute_deviations.R:458-461,517-520) are written
 *     as R's NA_real_ bit pattern 0x7FF00000000007A2.
 *   - One cv_handle = one GPU.  Cells (columns of the counts matrix) shard across GPUs.  Two ways
 *     to use several GPUs:
 *       (a) ONE process, cv_multi (section "several GPUs behind one call" below): the library shards
 *           the caller's column blocks, runs one host thread per device and performs the two
 *           cross-shard reductions itself over NVLink peer memory.  This is what the R shim binds
 *           (R is single-process; the reference's own fan-out is bplapply inside one process,
 *           R/compute_deviations.R:387-391).
 *       (b) one process per GPU (torchrun): the two reductions are exposed as device buffers so the
 *           host runs NCCL on them (cv_peak_totals_device, cv_variability_partials_device,
 *           cv_variability_merge_device).
 */
#ifndef CHROMVAR_B200_H
#define CHROMVAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cv_handle cv_handle;

enum {
  CV_OK = 0,
  CV_ERR_INVALID = 1,     /* bad argument / input fails the reference's own checks */
  CV_ERR_CUDA = 2,        /* CUDA runtime error (message holds cudaGetErrorString) */
  CV_ERR_OOM = 3,         /* device or pinned-host allocation failed */
  CV_ERR_STATE = 4,       /* call order: e.g. deviations before counts were set */
  CV_ERR_UNSUPPORTED = 5  /* shape outside what the kernels cover (message says which) */
};

/* element types for the polymorphic inputs */
enum { CV_I32 = 0, CV_I64 = 1, CV_F64 = 2, CV_U8 = 3, CV_F32 = 4 };

/* R's NA_real_ payload */
#define CV_NA_REAL_BITS 0x7FF00000000007A2ULL

/* ---- lifecycle ------------------------------------------------------------------------ */
/* device: CUDA ordinal.  seed: keys the Philox streams of the samplers (R shim derives it from
 * R's RNG so set.seed() keeps run-to-run reproducibility; the reference uses Rcpp::RNGScope,
 * src/RcppExports.cpp:88,100). */
int cv_create(int device, uint64_t seed, cv_handle** out);
void cv_destroy(cv_handle* h);
const char* cv_last_error(const cv_handle* h); /* h == NULL: last error of a failed cv_create */
int cv_version(void);
int cv_set_seed(cv_handle* h, uint64_t seed);
/* options:
 *   "path"             0 auto (cost model), 1 row-gather SpMM, 2 dense tcgen05 (fails if the problem
 *                      cannot be represented exactly in u8 operands / s32 accumulators)
 *   "dense_cta_pair"   GEMM kernel: 1 or -1 (default) tcgen05 cta_group::2 (256-cell tiles on two
 *                      SMs), 0 single CTA (128-cell tiles)
 *   "dense_epilogue"   arithmetic of the fused epilogue: 2 fp64 throughout; 1 the per-iteration
 *                      statistics without fp64 (mean over iterations in exact 64-bit fixed point,
 *                      variance in fp32 of exact integer differences, fp64 once per annotation;
 *                      needs sums < 2^24; dev agrees with fp64 to ~1e-9, sd to ~1e-6 relative);
 *                      0 (default) fp64 when a tile is long enough to hide it (>= 1100 k-blocks of
 *                      128 peaks) or the problem is tiny, the fixed-point form otherwise
 *   "dense_fp4"        -1 (default) / 1: the dense contraction runs on 4-bit e2m1 operands (tcgen05
 *                      kind::mxf4, f32 accumulators holding exact integers < 2^24, twice the MAC rate
 *                      of u8) whenever the problem can be represented exactly that way -- values
 *                      outside {0,1,2,3,4,6} are split into terms in extra "overflow" columns -- and
 *                      falls back to u8 otherwise; 0: never
 *   "dense_lockstep"   1 (default): in the e2m1 kernel the TMA producers of all CTAs start each sub-tile
 *                      together (L2 then serves all but the first request for an operand block); 0: free running
 *   "dense_rows_scatter" 1 builds the operand rows with the per-annotation scatter kernel only
 *   "cell_chunk"       cap on the cells per GEMM launch of the dense path (0 = memory budget)
 *   "small_n_gather"   few-cells contraction (N <= 256 cells, counts < 65536, no annotation listing a peak twice):
 *                      per iteration the counts rows are gathered by the background indices into an MN-major
 *                      tcgen05 operand, nothing is stacked.  -1 (default): when the stacked operand would
 *                      exceed 1 GB; 0: never; 1: whenever it applies
 *   "knn_insertion"    cv_background_peaks_knn: 1 keeps the neighbour list exact after every insertion at any
 *                      window (the kernel windows <= 16 always use); 0 (default): larger windows append
 *                      candidates to a buffer and sort it when full.  Same result either way.
 *   "check_zero_peaks" 1 (default) reject peaks without fragments (R/compute_deviations.R:362-363);
 *                      0 for a shard of a larger matrix (the caller checks the reduced totals)
 *   "tail_chunk"       cv_deviations_csc: 1 (default) a short fourth cell chunk ends the call so that little
 *                      of the results is still on its way to the host when the last GEMM finishes; 0: three chunks
 *   "host_narrow"      cv_deviations_csc: host threads between the caller's memory and the DMA engines.
 *                      -1 (default): when the caller's arrays are PAGEABLE (R's vectors are) the row indices
 *                      are staged and the values narrowed to bytes into page-locked memory slice by slice
 *                      ahead of the uploads, and the results land in page-locked memory and are copied on
 *                      in parallel pieces -- instead of the driver's single-threaded bounce copies; pinned
 *                      caller memory travels directly.  0: never.  1: always (also for pinned memory: fewer
 *                      bytes on a shared PCIe fabric); n > 1: with n threads.  Chunks holding anything but
 *                      integers 0..255 travel as they are.  Same results bit for bit. */
int cv_set_option(cv_handle* h, const char* key, int64_t value);
/* the CUDA stream all of this handle's kernels run on (a cudaStream_t), for event timing */
void* cv_stream(cv_handle* h);

/* ---- counts (assay "counts"; replaces the dgCMatrix / matrix the reference passes around) -- */
/* CSC as in Matrix::dgCMatrix: colptr (N+1, @p, CV_I32 or CV_I64), rowidx (@i, 0-based int32),
 * x (@x; CV_F64 as R stores it, or CV_I32 / CV_U8).  Values must be finite, >= 0 and
 * integer-valued (counts_check, R/utils.R:3-12).  nnz < 2^32 per handle. */
int cv_set_counts_csc(cv_handle* h, int64_t P, int64_t N, const void* colptr, int colptr_type,
                      const int32_t* rowidx, const void* x, int x_type);
/* dense column-major P x N (base matrix / dgeMatrix input; bulk ATAC, config 4) */
int cv_set_counts_dense(cv_handle* h, int64_t P, int64_t N, const void* x, int x_type);

/* getFragmentsPerPeak / getFragmentsPerSample / getTotalFragments
 * (R/helper_functions.R:48-58, 64-74, 79-89).  Any output may be NULL. */
int cv_fragments(cv_handle* h, double* out_per_peak /*P*/, double* out_per_sample /*N*/,
                 double* out_total /*1*/);
/* Multi-GPU: device pointer to this shard's per-peak totals, int64[P+1] (slot P = grand total).
 * The host all-reduces (sum) that buffer across shards (NCCL) and then calls
 * cv_peak_totals_commit so expectations / background sampling use the global totals. */
int cv_peak_totals_device(cv_handle* h, void** dptr, int64_t* n_elems);
int cv_peak_totals_commit(cv_handle* h);

/* computeExpectations -> compute_expectations_core (R/compute_deviations.R:415-448).
 * norm: 0/1.  group: NULL or N labels in 0..n_groups-1 (factor codes - 1). out_e: P doubles.
 * (norm / group variants are per-shard quantities; with sharded cells only the default
 * variant is defined across shards.) */
int cv_expectations(cv_handle* h, int norm, const int32_t* group, int n_groups, double* out_e);

/* ---- getBackgroundPeaks -> get_background_peaks_core (R/background_peaks.R:116-156) -------- */
/* bias: P doubles.  out_bg: P x niter column-major, 1-based peak indices (the reference returns
 * the same matrix as doubles).  Optional debug outputs (NULL to skip):
 *   out_trans P x 2 col-major (trans_norm_mat), out_membership P (1-based bin of each peak),
 *   out_density bs^2, out_binprob bs^2 x bs^2 row-major: row i = destination-bin distribution
 *   of a peak in bin i (what src/utils.cpp:83-99 induces). */
int cv_background_peaks(cv_handle* h, const double* bias, int niter, double w, int bs,
                        int32_t* out_bg, double* out_trans, int32_t* out_membership,
                        double* out_density, double* out_binprob);

/* ---- get_background_peaks_alternative (R/test_sets.R:291-330), the sampler of makePermutedSets
 * (R/test_sets.R:246-289): each background peak is drawn uniformly (with replacement) among the
 * `window` nearest OTHER peaks in the whitened (fragments per peak, bias) plane -- exact k nearest
 * neighbours (nabor::knn(k = window + 1, eps = 0) minus the query, :308-312); distance ties are
 * broken by the peak index (the reference leaves them to the kd-tree).
 * out_bg: P x niter column-major, 1-based.  Optional (NULL to skip): out_trans P x 2 col-major
 * (tmp_vals, :301); out_nn [P][window] row-major, 1-based neighbours in (distance, index) order. */
int cv_background_peaks_knn(cv_handle* h, const double* bias, int niter, int window,
                            int32_t* out_bg, double* out_trans, int32_t* out_nn);

/* ---- computeDeviations -> compute_deviations_core (R/compute_deviations.R:355-408), the whole
 * bplapply fan-out over compute_deviations_single (:451-537) in one call -------------------- */
/* annoptr: K+1 offsets into annoidx; annoidx: peak indices (index_base 0 or 1; duplicates keep
 * their multiplicity like sparseMatrix does, :480-484).  bg: P x B col-major int32 peak indices
 * (same index_base).  expectation: P doubles.
 * Outputs: out_dev, out_z K x N col-major; out_matches, out_overlap K (fractionMatches,
 * fractionBackgroundOverlap, :398-401); out_variability K or NULL (row SD of z over this
 * handle's cells, finite entries only = row_sds(z, TRUE), src/utils.cpp:9-25);
 * parity outputs or NULL: out_observed K x N col-major int64 (observed, :486),
 * out_sampled [K][B][N] int64, N fastest (sampled, :501). */
int cv_deviations(cv_handle* h, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                  int index_base, const int32_t* bg, int B, const double* expectation,
                  double* out_dev, double* out_z, double* out_matches, double* out_overlap,
                  double* out_variability, int64_t* out_observed, int64_t* out_sampled);

/* compute_deviations_core as ONE call on host data (what the R shim's .Call passes): the counts
 * matrix (arguments as cv_set_counts_csc), annotations / background / expectation (as
 * cv_deviations; expectation == NULL means computeExpectations' default, :357-359) in, results out.
 * Counts travel to the device in cell chunks while earlier chunks are already in the GEMM, and
 * finished chunks of the K x N results travel back while later chunks are computed.  Afterwards
 * the counts are resident exactly as after cv_set_counts_csc. */
int cv_deviations_csc(cv_handle* h, int64_t P, int64_t N, const void* colptr, int colptr_type,
                      const int32_t* rowidx, const void* x, int x_type, int64_t K,
                      const int64_t* annoptr, const int32_t* annoidx, int index_base,
                      const int32_t* bg, int B, const double* expectation, double* out_dev,
                      double* out_z, double* out_matches, double* out_overlap,
                      double* out_variability);

/* The same call split in three so a caller (bench.py `value`, repeated-set callers such as
 * R/variability_synergy.R:173-177) can keep inputs resident in HBM:
 *   upload   : H2D of annotations / background / expectation (raw input formats only)
 *   compute  : every kernel of the path on resident inputs; results stay on the device.
 *              rebuild_counts_layout != 0 also redoes the counts layout pass (what a fresh
 *              cv_set_counts_* triggers) so a timed step covers all derived work.
 *   download : D2H of the results (any pointer may be NULL). */
int cv_deviations_upload(cv_handle* h, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                         int index_base, const int32_t* bg, int B, const double* expectation);
int cv_deviations_compute(cv_handle* h, int rebuild_counts_layout, int keep_sums);
/* cv_deviations_compute(h, rebuild, 0) with the host out of the loop.  When the last fully checked
 * call (cv_deviations / cv_deviations_compute) ran on the SAME resident inputs and options without
 * taking a fallback, every kernel of the path is enqueued again -- same work, bit-identical results --
 * but the host waits neither for the flag words between the stages (their values are known from the
 * checked call) nor for the end of the call: consecutive calls queue up on the device and a
 * descheduled host thread no longer leaves the GPU idle.  Without such a verdict it IS the checked
 * call.  cv_deviations_wait -- or any entry point that reads results, reports statistics or changes
 * inputs / options -- waits for the queued calls and verifies their flag words. */
int cv_deviations_compute_async(cv_handle* h, int rebuild_counts_layout);
int cv_deviations_wait(cv_handle* h);
int cv_deviations_download(cv_handle* h, double* out_dev, double* out_z, double* out_matches,
                           double* out_overlap, double* out_variability, int64_t* out_observed,
                           int64_t* out_sampled);
/* Multi-GPU computeVariability: device pointer to double[K][4] = (n_finite, mean, M2,
 * n_nonfinite) of this shard's z rows; shards are merged with cv_merge_variability. */
int cv_variability_partials_device(cv_handle* h, void** dptr, int64_t* n_elems);
/* merge n_parts partial tables (host memory, each K x 4) -> sd[K] (n-1 denominator) */
int cv_merge_variability(const double* parts, int n_parts, int64_t K, int na_rm, double* out_sd);

/* Multi-process computeVariability without the host: `d_parts` is a DEVICE buffer holding n_parts
 * tables of K x 4 doubles back to back (an NCCL all-gather of cv_variability_partials_device, in rank
 * order); the merged row SDs replace this handle's variability (cv_deviations_download's
 * out_variability).  Enqueued on the handle's stream, no host synchronisation. */
int cv_variability_merge_device(cv_handle* h, const void* d_parts, int n_parts);

/* ---- several GPUs behind one call (replaces the bplapply fan-out, R/compute_deviations.R:387-391) ----
 * One process, n_dev devices, one host thread per device inside the library.  The counts arrive as
 * column blocks (a dgCMatrix holds < 2^31 nonzeros -- @p is int32 -- so an atlas is a list of
 * blocks); the library cuts the concatenated columns into one contiguous cell shard per
 * device, balanced by nonzeros, and each shard may span several blocks.  Per-peak totals are
 * summed and the per-annotation variability partials merged across devices by the library's own
 * kernels over peer memory (staged device-to-device copies where peer access is unavailable).
 * Replicated inputs: annotations, background matrix, expectation.  Outputs are the caller's K x N
 * (N = all cells) column-major matrices; every shard writes its own slab of columns.
 * Per-shard limits are those of a cv_handle (nonzeros < 2^32 per device). */
typedef struct cv_multi cv_multi;
typedef struct cv_csc_block {
  int64_t n_cols;          /* cells in this block */
  const void* colptr;      /* n_cols + 1 offsets, starting at 0 (a dgCMatrix's @p) */
  int colptr_type;         /* CV_I32 or CV_I64 */
  const int32_t* rowidx;   /* @i, 0-based */
  const void* x;           /* @x */
  int x_type;              /* CV_F64 / CV_F32 / CV_I32 / CV_U8, the same for every block */
} cv_csc_block;
/* devices: n_dev CUDA ordinals (NULL = 0 .. n_dev-1); the same seed keys every device's sampler */
int cv_create_multi(const int* devices, int n_dev, uint64_t seed, cv_multi** out);
void cv_destroy_multi(cv_multi* m);
const char* cv_multi_last_error(const cv_multi* m);   /* m == NULL: a failed cv_create_multi */
int cv_multi_n_devices(const cv_multi* m);
int cv_multi_peer_access(const cv_multi* m);          /* 1: reductions read peer memory directly */
cv_handle* cv_multi_handle(cv_multi* m, int i);       /* device i's engine (statistics, options) */
int cv_multi_set_seed(cv_multi* m, uint64_t seed);
/* cv_set_option on every device; "multi_p2p" 0 stages the peers' buffers instead of direct peer loads */
int cv_multi_set_option(cv_multi* m, const char* key, int64_t value);
/* counts resident on the devices, per-peak totals reduced (getFragmentsPer*, computeExpectations,
 * getBackgroundPeaks then see the whole matrix).  bounds: n_dev + 1 cell offsets of the shards. */
int cv_multi_set_counts(cv_multi* m, int64_t P, const cv_csc_block* blocks, int n_blocks);
int cv_multi_shards(const cv_multi* m, int64_t* bounds);
int cv_multi_fragments(cv_multi* m, double* out_per_peak, double* out_per_sample, double* out_total);
/* a base matrix, column-major P x N (x_type CV_F64 / CV_F32 / CV_I32): each device converts its slab of columns */
int cv_multi_set_counts_dense(cv_multi* m, int64_t P, int64_t N, const void* x, int x_type);
/* computeExpectations over ALL cells, arguments as cv_expectations (group: one 0-based label per cell) */
int cv_multi_expectations(cv_multi* m, int norm, const int32_t* group, int n_groups, double* out_e);
int cv_multi_background_peaks(cv_multi* m, const double* bias, int niter, double w, int bs, int32_t* out_bg);
/* compute_deviations_core on resident counts / as ONE call on host data (arguments as cv_deviations /
 * cv_deviations_csc; expectation == NULL in the _csc form means computeExpectations' default).
 * out_variability: row SD of z over ALL cells. */
int cv_multi_deviations(cv_multi* m, int64_t K, const int64_t* annoptr, const int32_t* annoidx,
                        int index_base, const int32_t* bg, int B, const double* expectation,
                        double* out_dev, double* out_z, double* out_matches, double* out_overlap,
                        double* out_variability);
int cv_multi_deviations_csc(cv_multi* m, int64_t P, const cv_csc_block* blocks, int n_blocks, int64_t K,
                            const int64_t* annoptr, const int32_t* annoidx, int index_base,
                            const int32_t* bg, int B, const double* expectation, double* out_dev,
                            double* out_z, double* out_matches, double* out_overlap,
                            double* out_variability);

/* ---- the reference's five registered native routines (src/RcppExports.cpp:59-119) --------- */
/* row_sds(x, na_rm)  src/utils.cpp:9-25.  x: K x N col-major. */
int cv_row_sds(cv_handle* h, int64_t K, int64_t N, const double* x, int na_rm, double* out_sd);
/* row_sds_perm(x, na_rm)  src/utils.cpp:28-33, batched: n_rep bootstrap replicates at once
 * (R/compute_variability.R:57-60 calls it bootstrap_samples times).  out: n_rep x K row-major. */
int cv_row_sds_perm(cv_handle* h, int64_t K, int64_t N, const double* x, int na_rm, int n_rep,
                    double* out_sd);
/* ProbSampleReplace(size, prob)  src/utils.cpp:38-77.  out: size 1-based indices. */
int cv_prob_sample_replace(cv_handle* h, int64_t size, int64_t n, const double* prob,
                           int32_t* out);
/* bg_sample_helper(bin_membership, bin_p, bin_density, niterations)  src/utils.cpp:83-99.
 * bin_membership: P 0-based bins; bin_p: nb x nb col-major; bin_density: nb.
 * out: P x niter col-major, 1-based. */
int cv_bg_sample_helper(cv_handle* h, int64_t P, const int32_t* bin_membership, int64_t nb,
                        const double* bin_p, const double* bin_density, int niter, int32_t* out);
/* euc_dist(x)  src/utils.cpp:107-117.  x: n x d col-major -> out n x n. */
int cv_euc_dist(cv_handle* h, int64_t n, int64_t d, const double* x, double* out);

/* ---- next rows of SURVEY 8f on the device --------------------------------------------------- */
/* getPermutedData (R/background_peaks.R:178-203): ONE permuted assay of the resident counts =
 * reorder_columns(counts, getBackgroundPeaks(object, niterations = ncol(object))), i.e. column c
 * of the result is counts[bg[, c], c].  The P x N index matrix is never materialised: the draws
 * (same Philox counters as cv_background_peaks: (peak, iteration = c)) happen inside the gather.
 * Two calls because the result's size is only known afterwards: cv_permuted_counts fills
 * out_colptr (N + 1, 0-based offsets) and *out_nnz; cv_permuted_counts_fetch copies the row indices
 * (0-based, increasing within a column) and values of that result. */
int cv_permuted_counts(cv_handle* h, const double* bias, double w, int bs, int64_t* out_colptr, int64_t* out_nnz);
int cv_permuted_counts_fetch(cv_handle* h, int64_t nnz, int32_t* out_rowidx, double* out_x);
/* deviationsCovariability (R/kmers.R:51-57): cov(t(z)) with row i divided by row_sds(z)[i]^2.
 * z: K x N column-major host matrix, or NULL = the z of the last deviations call on this handle,
 * still in HBM (K, N then come from the handle).  out: K x K column-major. */
int cv_deviations_covariability(cv_handle* h, int64_t K, int64_t N, const double* z, double* out);

/* ---- instrumentation ------------------------------------------------------------------- */
typedef struct cv_stage_stats {
  double ms_counts_layout;   /* row/col sums + cell-tiled row layout build (K1) */
  double ms_anno_prep;       /* background transpose + expected-fraction table + overlap */
  double ms_contract;        /* gather-contraction + fused epilogue (K5+K6+K7): dominant kernel */
  double ms_finalize;        /* variability merge + K x N transposes */
  double ms_h2d, ms_d2h;     /* host<->device copies of the last call (wall clock) */
  double ms_background;      /* last cv_background_peaks, device time */
  double contract_flops;     /* dense path: 2 * rows * cells * peaks of the padded GEMM (0 on the sparse path) */
  int64_t contract_updates;  /* (B+1) * sum_k sum_{p in set_k} nnz(counts row) actually visited */
  int64_t contract_bytes;    /* algorithmic bytes of the gather model, DESIGN.md section 4 */
  int64_t h2d_bytes, d2h_bytes;
  int64_t kernel_launches;   /* kernels launched by this handle since creation */
  int32_t path;              /* 0 = cell-tiled row-gather SpMM, 1 = dense tcgen05 */
  int32_t cell_tile;         /* row-gather: W; dense: cells per tile (256 = CTA pair, 128) */
  int32_t n_cell_tiles;      /* row-gather: tiles; dense: cell chunks (GEMM launches per plane) */
  int32_t acc_bytes;         /* accumulator width used (4 or 8) */
  int32_t planes;            /* dense path: u8 digit planes of the counts (1 when max count <= 255) */
  int32_t epilogue;          /* dense path: 32 = fixed-point / fp32 epilogue, 64 = fp64 epilogue */
  int32_t operand_bits;      /* dense path: 8 = u8 operands (tcgen05 kind::i8), 4 = e2m1 operands (kind::mxf4) */
  int32_t overflow_columns;  /* e2m1 operands: extra columns holding the terms of values outside {0,1,2,3,4,6}, all chunks */
  int32_t few_cells;         /* 1: the few-cells contraction ran (counts rows gathered per iteration, no stacked operand) */
  int32_t reserved;
} cv_stage_stats;
int cv_stats(cv_handle* h, cv_stage_stats* out);
int cv_multi_stats(cv_multi* m, int i, cv_stage_stats* out);   /* device i of a cv_multi */

/* Host-only: the conversion cv_deviations_csc applies to wide value types before the upload
 * ("host_narrow" option; a dgCMatrix's x slot is double, R/helper_functions.R:50-88).  x_type CV_F64 /
 * CV_F32 / CV_I32.  Returns 1 when every value is an integer in 0..255 and out[0, n) holds them,
 * 0 when some value is not (out is then unspecified), < 0 on bad arguments.  No GPU involved. */
int cv_host_narrow(const void* x, int x_type, int64_t n, uint8_t* out, int threads);

/* Diagnostic (no device needed): how the e2m1 variant of the dense path writes the integer v -- its
 * terms (each in {1,2,3,4,6}, summing to v; 5 = 4+1, 13 = 6+6+1) and their 4-bit e2m1 codes.  Returns
 * the number of terms (1 for v = 0), or -1 when max_terms is too small.  terms / codes may be NULL. */
int cv_e2m1_terms(uint32_t v, uint32_t* terms, uint32_t* codes, int max_terms);

#ifdef __cplusplus
}
#endif
#endif /* CHROMVAR_B200_H */
