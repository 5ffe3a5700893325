/*
 * Stand-in for R's <Rinternals.h>: the slice of the R C API that rpkg/src/cv_rglue.c uses, declared
 * with R's own names and semantics and implemented by rstub.c.  TEST INFRASTRUCTURE -- R and its
 * headers are not installed in this image, so the .Call glue is compiled and exercised against
 * this (like oracle/shim/ stands in for RcppArmadillo).  Written from R's documented API
 * ("Writing R Extensions", sections 5.9-5.10); nothing is copied from R's sources.
 */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef struct rstub_sexp* SEXP;
typedef enum { FALSE = 0, TRUE } Rboolean;

#define NILSXP 0
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19

extern SEXP R_NilValue;
extern SEXP R_DimSymbol;
extern SEXP R_NamesSymbol;
extern int R_NaInt;
extern double R_NaReal;
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
#define NA_REAL R_NaReal

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_protect(SEXP s);
void Rf_unprotect(int n);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)

int TYPEOF(SEXP s);
R_xlen_t XLENGTH(SEXP s);
int LENGTH(SEXP s);
int Rf_length(SEXP s);
double* REAL(SEXP s);
int* INTEGER(SEXP s);
int* LOGICAL(SEXP s);
SEXP VECTOR_ELT(SEXP list, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP list, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP s, R_xlen_t i);
const char* CHAR(SEXP charsxp);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkString(const char* s);
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v);

Rboolean Rf_isNull(SEXP s);
Rboolean Rf_isReal(SEXP s);
Rboolean Rf_isInteger(SEXP s);
int Rf_asInteger(SEXP s);
double Rf_asReal(SEXP s);
int Rf_asLogical(SEXP s);
SEXP Rf_getAttrib(SEXP s, SEXP name);
SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP val);

#if defined(__GNUC__)
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char* fmt, ...);
#endif
void Rf_warning(const char* fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
