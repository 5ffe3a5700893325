/*
 * .Call glue between R and libchromvar_b200.so -- the binding a chromVAR maintainer adds next to
 * src/RcppExports.cpp (which registers _chromVAR_row_sds, _chromVAR_bg_sample_helper, ... the
 * same way, src/RcppExports.cpp:121-137).  Plain R C API, no Rcpp needed.
 *
 * Everything goes through ONE cv_multi object (include/chromvar_b200.h, "several GPUs behind one
 * call"): R is a single process, so the library itself drives all the GPUs named in cvb_init().
 * The counts are passed as LISTS of dgCMatrix slots -- one element for an ordinary matrix, several
 * for an atlas whose nonzeros exceed what one dgCMatrix can index (@p is int32).
 *
 * R and its headers are absent from this repository's image: the file is compiled and exercised
 * against the stand-in runtime in rpkg/rstub/ (tests/test_gpu_rglue.py drives every registered
 * routine through the registered table).  In an R package: PKG_CPPFLAGS = -I<repo>/include,
 * PKG_LIBS = -L<repo>/chromvar_b200/lib -lchromvar_b200 (rpkg/src/Makevars).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>
#include "chromvar_b200.h"

static cv_multi *M = NULL;

static void need_engine(void) {
  if (!M) Rf_error("chromVARb200: call cvb_init() first");
}
static void check(int rc) {
  if (rc != CV_OK) Rf_error("%s", cv_multi_last_error(M));       /* stop(<library message>) */
}
static void check0(int rc) {                                     /* calls on device 0's own engine */
  if (rc != CV_OK) Rf_error("%s", cv_last_error(cv_multi_handle(M, 0)));
}

/* cvb_init(devices, seed): devices = integer vector of CUDA ordinals; seed is drawn in R
 * (sample.int) so set.seed() still governs results (the reference: Rcpp::RNGScope,
 * src/RcppExports.cpp:88,100). */
SEXP cvb_init(SEXP devices, SEXP seed) {
  if (M) { cv_destroy_multi(M); M = NULL; }
  int n = (int) XLENGTH(devices);
  if (n < 1) Rf_error("cvb_init: no device given");
  int *dev = (int *) R_alloc(n, sizeof(int));
  for (int k = 0; k < n; ++k)
    dev[k] = TYPEOF(devices) == REALSXP ? (int) REAL(devices)[k] : INTEGER(devices)[k];
  int rc = cv_create_multi(dev, n, (uint64_t) Rf_asReal(seed), &M);
  if (rc != CV_OK) { M = NULL; Rf_error("%s", cv_multi_last_error(NULL)); }
  return R_NilValue;
}

SEXP cvb_shutdown(void) {
  if (M) { cv_destroy_multi(M); M = NULL; }
  return R_NilValue;
}

SEXP cvb_n_devices(void) {
  SEXP out = PROTECT(Rf_allocVector(INTSXP, 1));
  INTEGER(out)[0] = M ? cv_multi_n_devices(M) : 0;
  UNPROTECT(1);
  return out;
}

SEXP cvb_set_seed(SEXP seed) {
  need_engine();
  check(cv_multi_set_seed(M, (uint64_t) Rf_asReal(seed)));
  return R_NilValue;
}

SEXP cvb_set_option(SEXP key, SEXP value) {
  need_engine();
  if (TYPEOF(key) != STRSXP || XLENGTH(key) != 1) Rf_error("cvb_set_option: key must be a string");
  check(cv_multi_set_option(M, CHAR(STRING_ELT(key, 0)), (int64_t) Rf_asReal(value)));
  return R_NilValue;
}

/* lists of the @p (integer), @i (integer), @x (double) slots of the column blocks -> cv_csc_block[] */
static cv_csc_block *blocks_from(SEXP plist, SEXP ilist, SEXP xlist, int *n_blocks) {
  if (TYPEOF(plist) != VECSXP || TYPEOF(ilist) != VECSXP || TYPEOF(xlist) != VECSXP)
    Rf_error("counts: p, i, x must be lists of the dgCMatrix slots of the column blocks");
  int nb = (int) XLENGTH(plist);
  if (nb < 1 || XLENGTH(ilist) != nb || XLENGTH(xlist) != nb) Rf_error("counts: p, i, x lists differ in length");
  cv_csc_block *b = (cv_csc_block *) R_alloc(nb, sizeof(cv_csc_block));
  for (int k = 0; k < nb; ++k) {
    SEXP p = VECTOR_ELT(plist, k), i = VECTOR_ELT(ilist, k), x = VECTOR_ELT(xlist, k);
    if (TYPEOF(p) != INTSXP || XLENGTH(p) < 1) Rf_error("counts: block %d: @p must be a non-empty integer vector", k + 1);
    if (TYPEOF(i) != INTSXP) Rf_error("counts: block %d: @i must be integer", k + 1);
    if (TYPEOF(x) != REALSXP && TYPEOF(x) != INTSXP) Rf_error("counts: block %d: @x must be numeric", k + 1);
    if (XLENGTH(i) != XLENGTH(x)) Rf_error("counts: block %d: @i and @x differ in length", k + 1);
    if (TYPEOF(x) != TYPEOF(VECTOR_ELT(xlist, 0))) Rf_error("counts: all blocks must store @x in the same type");
    b[k].n_cols = (int64_t) XLENGTH(p) - 1;
    b[k].colptr = INTEGER(p);
    b[k].colptr_type = CV_I32;
    if ((R_xlen_t) INTEGER(p)[XLENGTH(p) - 1] != XLENGTH(i)) Rf_error("counts: block %d: @p does not end at length(@i)", k + 1);
    b[k].rowidx = INTEGER(i);
    b[k].x = TYPEOF(x) == REALSXP ? (const void *) REAL(x) : (const void *) INTEGER(x);
    b[k].x_type = TYPEOF(x) == REALSXP ? CV_F64 : CV_I32;
  }
  *n_blocks = nb;
  return b;
}

/* counts resident on the GPU(s): cvb_set_counts(list(@p...), list(@i...), list(@x...), nrow) */
SEXP cvb_set_counts(SEXP plist, SEXP ilist, SEXP xlist, SEXP nrow) {
  need_engine();
  int nb = 0;
  cv_csc_block *b = blocks_from(plist, ilist, xlist, &nb);
  check(cv_multi_set_counts(M, (int64_t) Rf_asInteger(nrow), b, nb));
  return R_NilValue;
}

/* base matrix input (bulk data).  One device: uploaded as it is.  Several: the shim converts to
 * dgCMatrix first (as(x, "dgCMatrix")) -- a dense matrix of a few hundred samples is a one-GPU job. */
SEXP cvb_set_counts_dense(SEXP x) {
  need_engine();
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  if (Rf_isNull(dim) || XLENGTH(dim) != 2) Rf_error("counts must be a matrix");
  /* column-major like R's matrix: every device takes a slab of columns */
  if (TYPEOF(x) == INTSXP)
    check(cv_multi_set_counts_dense(M, INTEGER(dim)[0], INTEGER(dim)[1], INTEGER(x), CV_I32));
  else if (TYPEOF(x) == REALSXP)
    check(cv_multi_set_counts_dense(M, INTEGER(dim)[0], INTEGER(dim)[1], REAL(x), CV_F64));
  else
    Rf_error("counts must be numeric");
  return R_NilValue;
}

/* getFragmentsPerPeak / PerSample / Total in one call: list(peak, sample, total) */
SEXP cvb_fragments(SEXP P, SEXP N) {
  need_engine();
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SEXP a = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(P)));
  SEXP b = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(N)));
  SEXP c = PROTECT(Rf_allocVector(REALSXP, 1));
  if (cv_multi_n_devices(M) == 1) check0(cv_fragments(cv_multi_handle(M, 0), REAL(a), REAL(b), REAL(c)));
  else check(cv_multi_fragments(M, REAL(a), REAL(b), REAL(c)));
  SET_VECTOR_ELT(out, 0, a); SET_VECTOR_ELT(out, 1, b); SET_VECTOR_ELT(out, 2, c);
  UNPROTECT(4);
  return out;
}

/* compute_expectations_core: group = integer codes of as.factor(group) (1-based) or NULL */
SEXP cvb_expectations(SEXP P, SEXP norm, SEXP group, SEXP nlevels) {
  need_engine();
  SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(P)));
  const int nrm = Rf_asLogical(norm) == 1;
  int32_t *g = NULL;
  if (!Rf_isNull(group)) {
    if (TYPEOF(group) != INTSXP) Rf_error("group must be the integer codes of a factor");
    R_xlen_t n = XLENGTH(group);
    g = (int32_t *) R_alloc(n, sizeof(int32_t));
    for (R_xlen_t k = 0; k < n; ++k) g[k] = INTEGER(group)[k] - 1;
  }
  /* norm / group: sums over cells of per-cell terms, accumulated per shard and added on the first device */
  check(cv_multi_expectations(M, nrm, g, g ? Rf_asInteger(nlevels) : 0, REAL(out)));
  UNPROTECT(1);
  return out;
}

static SEXP int_matrix_as_double(const int32_t *src, R_xlen_t P, int B) {
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int) P, B));
  double *o = REAL(out);
  for (R_xlen_t k = 0; k < P * (R_xlen_t) B; ++k) o[k] = (double) src[k];
  UNPROTECT(1);
  return out;
}

/* get_background_peaks_core: returns the P x niterations matrix, storage mode double like
 * bg_sample_helper's arma::umat -> REALSXP (R/RcppExports.R:32-34) */
SEXP cvb_background_peaks(SEXP bias, SEXP niter, SEXP w, SEXP bs) {
  need_engine();
  R_xlen_t P = XLENGTH(bias);
  int B = Rf_asInteger(niter);
  if (B < 1) Rf_error("niterations must be >= 1");
  int32_t *tmp = (int32_t *) R_alloc((size_t) P * B, sizeof(int32_t));
  check(cv_multi_background_peaks(M, REAL(bias), B, Rf_asReal(w), Rf_asInteger(bs), tmp));
  return int_matrix_as_double(tmp, P, B);
}

/* get_background_peaks_alternative (R/test_sets.R:291-330): P x niterations matrix (double, like the
 * do.call(rbind, ...) of sample() results the reference returns).  Cell independent: device 0,
 * whose per-peak totals are the global ones. */
SEXP cvb_background_peaks_knn(SEXP bias, SEXP niter, SEXP window) {
  need_engine();
  R_xlen_t P = XLENGTH(bias);
  int B = Rf_asInteger(niter);
  if (B < 1) Rf_error("niterations must be >= 1");
  int32_t *tmp = (int32_t *) R_alloc((size_t) P * B, sizeof(int32_t));
  check0(cv_background_peaks_knn(cv_multi_handle(M, 0), REAL(bias), B, Rf_asInteger(window), tmp, NULL, NULL));
  return int_matrix_as_double(tmp, P, B);
}

/* background_peaks as R holds it: a P x B matrix of 1-based indices, integer or (what
 * getBackgroundPeaks returns) double */
static const int32_t *bg_as_int(SEXP bg, int *B, R_xlen_t *P) {
  SEXP bdim = Rf_getAttrib(bg, R_DimSymbol);
  if (Rf_isNull(bdim) || XLENGTH(bdim) != 2) Rf_error("background_peaks must be a matrix");
  *P = INTEGER(bdim)[0];
  *B = INTEGER(bdim)[1];
  if (TYPEOF(bg) == INTSXP) return INTEGER(bg);
  if (TYPEOF(bg) != REALSXP) Rf_error("background_peaks must be numeric");
  R_xlen_t n = XLENGTH(bg);
  int32_t *out = (int32_t *) R_alloc(n, sizeof(int32_t));
  const double *src = REAL(bg);
  for (R_xlen_t k = 0; k < n; ++k) {
    const double v = src[k];
    if (!(v >= 1.0 && v <= 2147483647.0) || v != (double) (int32_t) v)
      Rf_error("background_peaks must hold whole peak indices");
    out[k] = (int32_t) v;
  }
  return out;
}

static int64_t *annoptr_as_i64(SEXP annoptr, int64_t *K) {
  *K = (int64_t) XLENGTH(annoptr) - 1;
  if (*K < 1) Rf_error("Annotations are not valid");
  int64_t *ptr = (int64_t *) R_alloc(*K + 1, sizeof(int64_t));
  for (int64_t k = 0; k <= *K; ++k)
    ptr[k] = TYPEOF(annoptr) == REALSXP ? (int64_t) REAL(annoptr)[k] : (int64_t) INTEGER(annoptr)[k];
  return ptr;
}

static SEXP result_list(int64_t K, int n, SEXP *dev, SEXP *z, SEXP *m, SEXP *o, SEXP *v) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
  *dev = PROTECT(Rf_allocMatrix(REALSXP, (int) K, n));
  *z = PROTECT(Rf_allocMatrix(REALSXP, (int) K, n));
  *m = PROTECT(Rf_allocVector(REALSXP, K));
  *o = PROTECT(Rf_allocVector(REALSXP, K));
  *v = PROTECT(Rf_allocVector(REALSXP, K));
  SET_VECTOR_ELT(out, 0, *dev); SET_VECTOR_ELT(out, 1, *z); SET_VECTOR_ELT(out, 2, *m);
  SET_VECTOR_ELT(out, 3, *o); SET_VECTOR_ELT(out, 4, *v);
  return out;                                            /* 6 objects protected */
}

/* compute_deviations_core on resident counts: annoptr (K+1 offsets), annoidx (integer, 1-based),
 * bg (P x B, 1-based), expectation (double P).
 * Returns list(deviations K x N, z K x N, matches K, overlap K, variability K). */
SEXP cvb_deviations(SEXP annoptr, SEXP annoidx, SEXP bg, SEXP expectation, SEXP N) {
  need_engine();
  int64_t K;
  int64_t *ptr = annoptr_as_i64(annoptr, &K);
  int B;
  R_xlen_t P;
  const int32_t *bgi = bg_as_int(bg, &B, &P);
  if (TYPEOF(annoidx) != INTSXP) Rf_error("Annotations are not valid");
  if (XLENGTH(expectation) != P) Rf_error("length(expectation) == nrow(counts_mat) is not TRUE");
  SEXP dev, z, m, o, v;
  SEXP out = result_list(K, Rf_asInteger(N), &dev, &z, &m, &o, &v);
  check(cv_multi_deviations(M, K, ptr, INTEGER(annoidx), 1, bgi, B, REAL(expectation), REAL(dev), REAL(z),
                            REAL(m), REAL(o), REAL(v)));
  UNPROTECT(6);
  return out;
}

/* compute_deviations_core in ONE library call on R's own buffers: the dgCMatrix slots of the counts
 * (lists: one element per column block) plus everything cvb_deviations takes.  Every device streams
 * its shard of the cells in while earlier chunks are in the GEMM and streams its slab of
 * deviations / z back; the counts stay resident afterwards (later cvb_expectations /
 * cvb_background_peaks calls reuse them).  expectation may be NULL: computeExpectations' default
 * (R/compute_deviations.R:357-359). */
SEXP cvb_deviations_csc(SEXP plist, SEXP ilist, SEXP xlist, SEXP nrow, SEXP annoptr, SEXP annoidx, SEXP bg,
                        SEXP expectation) {
  need_engine();
  int nb = 0;
  cv_csc_block *blocks = blocks_from(plist, ilist, xlist, &nb);
  int64_t n = 0;
  for (int k = 0; k < nb; ++k) n += blocks[k].n_cols;
  if (n > 2147483647) Rf_error("more than 2^31 - 1 cells");
  int64_t K;
  int64_t *ptr = annoptr_as_i64(annoptr, &K);
  int B;
  R_xlen_t P;
  const int32_t *bgi = bg_as_int(bg, &B, &P);
  if (P != (R_xlen_t) Rf_asInteger(nrow)) Rf_error("nrow(counts_mat) == nrow(background_peaks) is not TRUE");
  if (TYPEOF(annoidx) != INTSXP) Rf_error("Annotations are not valid");
  if (!Rf_isNull(expectation) && XLENGTH(expectation) != P)
    Rf_error("length(expectation) == nrow(counts_mat) is not TRUE");
  SEXP dev, z, m, o, v;
  SEXP out = result_list(K, (int) n, &dev, &z, &m, &o, &v);
  check(cv_multi_deviations_csc(M, (int64_t) P, blocks, nb, K, ptr, INTEGER(annoidx), 1, bgi, B,
                                Rf_isNull(expectation) ? NULL : REAL(expectation), REAL(dev), REAL(z), REAL(m),
                                REAL(o), REAL(v)));
  UNPROTECT(6);
  return out;
}

/* row_sds / row_sds_perm (src/utils.cpp:9-33; R/RcppExports.R:20-26) */
SEXP cvb_row_sds(SEXP x, SEXP na_rm) {
  need_engine();
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  if (Rf_isNull(dim) || XLENGTH(dim) != 2) Rf_error("x must be a matrix");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, INTEGER(dim)[0]));
  check0(cv_row_sds(cv_multi_handle(M, 0), INTEGER(dim)[0], INTEGER(dim)[1], REAL(x), Rf_asLogical(na_rm) == 1, REAL(out)));
  UNPROTECT(1);
  return out;
}

SEXP cvb_row_sds_perm(SEXP x, SEXP na_rm, SEXP n_rep) {
  need_engine();
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  if (Rf_isNull(dim) || XLENGTH(dim) != 2) Rf_error("x must be a matrix");
  int R = Rf_asInteger(n_rep), K = INTEGER(dim)[0];
  if (R < 1) Rf_error("bootstrap_samples > 0 && all_whole(bootstrap_samples) is not TRUE");
  double *tmp = (double *) R_alloc((size_t) R * K, sizeof(double));
  check0(cv_row_sds_perm(cv_multi_handle(M, 0), K, INTEGER(dim)[1], REAL(x), Rf_asLogical(na_rm) == 1, R, tmp));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, K));       /* rbind of the replicates */
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) REAL(out)[(size_t) k * R + r] = tmp[(size_t) r * K + k];
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
  {"cvb_init", (DL_FUNC) &cvb_init, 2},
  {"cvb_shutdown", (DL_FUNC) &cvb_shutdown, 0},
  {"cvb_n_devices", (DL_FUNC) &cvb_n_devices, 0},
  {"cvb_set_seed", (DL_FUNC) &cvb_set_seed, 1},
  {"cvb_set_option", (DL_FUNC) &cvb_set_option, 2},
  {"cvb_set_counts", (DL_FUNC) &cvb_set_counts, 4},
  {"cvb_set_counts_dense", (DL_FUNC) &cvb_set_counts_dense, 1},
  {"cvb_fragments", (DL_FUNC) &cvb_fragments, 2},
  {"cvb_expectations", (DL_FUNC) &cvb_expectations, 4},
  {"cvb_background_peaks", (DL_FUNC) &cvb_background_peaks, 4},
  {"cvb_background_peaks_knn", (DL_FUNC) &cvb_background_peaks_knn, 3},
  {"cvb_deviations", (DL_FUNC) &cvb_deviations, 5},
  {"cvb_deviations_csc", (DL_FUNC) &cvb_deviations_csc, 8},
  {"cvb_row_sds", (DL_FUNC) &cvb_row_sds, 2},
  {"cvb_row_sds_perm", (DL_FUNC) &cvb_row_sds_perm, 3},
  {NULL, NULL, 0}
};

void R_init_chromVARb200(DllInfo *dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
