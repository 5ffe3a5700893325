/*
 * .Call glue between R and libchromvar_b200.so -- the binding a chromVAR maintainer adds next to
 * src/RcppExports.cpp (which registers _chromVAR_row_sds, _chromVAR_bg_sample_helper, ... the
 * same way, src/RcppExports.cpp:121-137).  Plain R C API, no Rcpp needed.
 *
 * NOT compiled in this repository's image (R and its headers are absent); it is the reference-side
 * stub INTEGRATION.md describes.  Build inside the R package with
 *   PKG_CPPFLAGS = -I<repo>/include     PKG_LIBS = -L<repo>/chromvar_b200/lib -lchromvar_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "chromvar_b200.h"

static cv_handle *H = NULL;

static void check(int rc) {
  if (rc != CV_OK) Rf_error("%s", cv_last_error(H));   /* stop(cv_last_error()) */
}

/* cvb_init(device, seed): seed is drawn in R (sample.int) so set.seed() still governs results */
SEXP cvb_init(SEXP device, SEXP seed) {
  if (H) { cv_destroy(H); H = NULL; }
  int rc = cv_create(Rf_asInteger(device), (uint64_t) Rf_asReal(seed), &H);
  if (rc != CV_OK) Rf_error("%s", cv_last_error(NULL));
  return R_NilValue;
}

SEXP cvb_set_seed(SEXP seed) { check(cv_set_seed(H, (uint64_t) Rf_asReal(seed))); return R_NilValue; }

/* counts: dgCMatrix slots @p (integer), @i (integer), @x (double), @Dim */
SEXP cvb_set_counts_csc(SEXP p, SEXP i, SEXP x, SEXP dim) {
  check(cv_set_counts_csc(H, INTEGER(dim)[0], INTEGER(dim)[1], INTEGER(p), CV_I32, INTEGER(i),
                          REAL(x), CV_F64));
  return R_NilValue;
}

SEXP cvb_set_counts_dense(SEXP x) {
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  if (TYPEOF(x) == INTSXP)
    check(cv_set_counts_dense(H, INTEGER(dim)[0], INTEGER(dim)[1], INTEGER(x), CV_I32));
  else
    check(cv_set_counts_dense(H, INTEGER(dim)[0], INTEGER(dim)[1], REAL(x), CV_F64));
  return R_NilValue;
}

/* getFragmentsPerPeak / PerSample / Total in one call: list(peak, sample, total) */
SEXP cvb_fragments(SEXP P, SEXP N) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SEXP a = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(P)));
  SEXP b = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(N)));
  SEXP c = PROTECT(Rf_allocVector(REALSXP, 1));
  check(cv_fragments(H, REAL(a), REAL(b), REAL(c)));
  SET_VECTOR_ELT(out, 0, a); SET_VECTOR_ELT(out, 1, b); SET_VECTOR_ELT(out, 2, c);
  UNPROTECT(4);
  return out;
}

/* compute_expectations_core: group = integer codes of as.factor(group) (1-based) or NULL */
SEXP cvb_expectations(SEXP P, SEXP norm, SEXP group, SEXP nlevels) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(P)));
  int32_t *g = NULL;
  if (!Rf_isNull(group)) {
    R_xlen_t n = XLENGTH(group);
    g = (int32_t *) R_alloc(n, sizeof(int32_t));
    for (R_xlen_t k = 0; k < n; ++k) g[k] = INTEGER(group)[k] - 1;
  }
  check(cv_expectations(H, Rf_asLogical(norm), g, Rf_asInteger(nlevels), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* get_background_peaks_core: returns the P x niterations matrix, storage mode double like
 * bg_sample_helper's arma::umat -> REALSXP (R/RcppExports.R:32-34) */
SEXP cvb_background_peaks(SEXP bias, SEXP niter, SEXP w, SEXP bs) {
  R_xlen_t P = XLENGTH(bias);
  int B = Rf_asInteger(niter);
  int32_t *tmp = (int32_t *) R_alloc((size_t) P * B, sizeof(int32_t));
  check(cv_background_peaks(H, REAL(bias), B, Rf_asReal(w), Rf_asInteger(bs), tmp, NULL, NULL, NULL, NULL));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int) P, B));
  double *o = REAL(out);
  for (R_xlen_t k = 0; k < P * (R_xlen_t) B; ++k) o[k] = (double) tmp[k];
  UNPROTECT(1);
  return out;
}

/* get_background_peaks_alternative (R/test_sets.R:291-330): P x niterations matrix (double, like the
 * do.call(rbind, ...) of sample() results the reference returns) */
SEXP cvb_background_peaks_knn(SEXP bias, SEXP niter, SEXP window) {
  R_xlen_t P = XLENGTH(bias);
  int B = Rf_asInteger(niter);
  int32_t *tmp = (int32_t *) R_alloc((size_t) P * B, sizeof(int32_t));
  check(cv_background_peaks_knn(H, REAL(bias), B, Rf_asInteger(window), tmp, NULL, NULL));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int) P, B));
  double *o = REAL(out);
  for (R_xlen_t k = 0; k < P * (R_xlen_t) B; ++k) o[k] = (double) tmp[k];
  UNPROTECT(1);
  return out;
}

/* compute_deviations_core: annoptr (double, K+1 offsets), annoidx (integer, 1-based),
 * bg (integer matrix P x B, 1-based), expectation (double P).
 * Returns list(deviations K x N, z K x N, matches K, overlap K, variability K). */
SEXP cvb_deviations(SEXP annoptr, SEXP annoidx, SEXP bg, SEXP expectation, SEXP N) {
  int64_t K = (int64_t) XLENGTH(annoptr) - 1;
  int n = Rf_asInteger(N);
  int64_t *ptr = (int64_t *) R_alloc(K + 1, sizeof(int64_t));
  for (int64_t k = 0; k <= K; ++k) ptr[k] = (int64_t) REAL(annoptr)[k];
  SEXP bdim = Rf_getAttrib(bg, R_DimSymbol);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
  SEXP dev = PROTECT(Rf_allocMatrix(REALSXP, (int) K, n));
  SEXP z = PROTECT(Rf_allocMatrix(REALSXP, (int) K, n));
  SEXP m = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP o = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP v = PROTECT(Rf_allocVector(REALSXP, K));
  check(cv_deviations(H, K, ptr, INTEGER(annoidx), 1, INTEGER(bg), INTEGER(bdim)[1], REAL(expectation),
                      REAL(dev), REAL(z), REAL(m), REAL(o), REAL(v), NULL, NULL));
  SET_VECTOR_ELT(out, 0, dev); SET_VECTOR_ELT(out, 1, z); SET_VECTOR_ELT(out, 2, m);
  SET_VECTOR_ELT(out, 3, o); SET_VECTOR_ELT(out, 4, v);
  UNPROTECT(6);
  return out;
}

/* compute_deviations_core in ONE library call on R's own buffers: dgCMatrix slots of the counts
 * plus everything cvb_deviations takes.  The library streams the counts to the GPU in cell chunks
 * while earlier chunks are in the GEMM and streams finished chunks of deviations / z back; the
 * counts stay resident afterwards (later cvb_expectations / cvb_background_peaks calls reuse them).
 * expectation may be NULL: computeExpectations' default (R/compute_deviations.R:357-359). */
SEXP cvb_deviations_csc(SEXP p, SEXP i, SEXP x, SEXP dim, SEXP annoptr, SEXP annoidx, SEXP bg,
                        SEXP expectation) {
  int64_t K = (int64_t) XLENGTH(annoptr) - 1;
  int P = INTEGER(dim)[0], n = INTEGER(dim)[1];
  int64_t *ptr = (int64_t *) R_alloc(K + 1, sizeof(int64_t));
  for (int64_t k = 0; k <= K; ++k) ptr[k] = (int64_t) REAL(annoptr)[k];
  SEXP bdim = Rf_getAttrib(bg, R_DimSymbol);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
  SEXP dev = PROTECT(Rf_allocMatrix(REALSXP, (int) K, n));
  SEXP z = PROTECT(Rf_allocMatrix(REALSXP, (int) K, n));
  SEXP m = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP o = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP v = PROTECT(Rf_allocVector(REALSXP, K));
  check(cv_deviations_csc(H, P, n, INTEGER(p), CV_I32, INTEGER(i), REAL(x), CV_F64, K, ptr, INTEGER(annoidx), 1,
                          INTEGER(bg), INTEGER(bdim)[1], Rf_isNull(expectation) ? NULL : REAL(expectation),
                          REAL(dev), REAL(z), REAL(m), REAL(o), REAL(v)));
  SET_VECTOR_ELT(out, 0, dev); SET_VECTOR_ELT(out, 1, z); SET_VECTOR_ELT(out, 2, m);
  SET_VECTOR_ELT(out, 3, o); SET_VECTOR_ELT(out, 4, v);
  UNPROTECT(6);
  return out;
}

/* row_sds / row_sds_perm (src/utils.cpp:9-33) */
SEXP cvb_row_sds(SEXP x, SEXP na_rm) {
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, INTEGER(dim)[0]));
  check(cv_row_sds(H, INTEGER(dim)[0], INTEGER(dim)[1], REAL(x), Rf_asLogical(na_rm), REAL(out)));
  UNPROTECT(1);
  return out;
}

SEXP cvb_row_sds_perm(SEXP x, SEXP na_rm, SEXP n_rep) {
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  int R = Rf_asInteger(n_rep), K = INTEGER(dim)[0];
  double *tmp = (double *) R_alloc((size_t) R * K, sizeof(double));
  check(cv_row_sds_perm(H, K, INTEGER(dim)[1], REAL(x), Rf_asLogical(na_rm), R, tmp));
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, R, K));       /* rbind of the replicates */
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) REAL(out)[(size_t) k * R + r] = tmp[(size_t) r * K + k];
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
  {"cvb_init", (DL_FUNC) &cvb_init, 2},
  {"cvb_set_seed", (DL_FUNC) &cvb_set_seed, 1},
  {"cvb_set_counts_csc", (DL_FUNC) &cvb_set_counts_csc, 4},
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
