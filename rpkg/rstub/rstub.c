/*
 * rstub.c -- a miniature stand-in for the R runtime, enough to LOAD and CALL rpkg/src/cv_rglue.c the
 * way R does: R_init_<pkg>() registers the .Call table, `.Call("name", ...)` looks the routine up
 * there, checks the argument count, runs it with R-style vectors (SEXP) as arguments and turns
 * Rf_error() into an R-level error.  TEST INFRASTRUCTURE (R is not installed in this image); the
 * harness side (rstub_*) is driven from Python through ctypes (tests/test_gpu_rglue.py).
 *
 * What it checks that a blind reading of the glue cannot: symbol registration and arity, argument
 * marshalling (integer / double / list / dim attributes), PROTECT / UNPROTECT balance, that errors
 * leave through Rf_error with the library's message, and that results have R's shapes.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "include/R.h"
#include "include/Rinternals.h"
#include "include/R_ext/Rdynload.h"

struct rstub_sexp {
  int type;
  R_xlen_t len;
  void* data;            /* REAL / INTEGER / LOGICAL payload, SEXP[] for lists and string vectors, char[] for CHARSXP */
  SEXP dim;              /* the "dim" attribute (INTSXP) or NULL */
  SEXP names;
  struct rstub_sexp* next;   /* allocation list */
};
struct rstub_dllinfo { int dynamic_symbols; };

static struct rstub_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL};
static struct rstub_sexp dimsym_obj = {CHARSXP, 3, (void*)"dim", NULL, NULL, NULL};
static struct rstub_sexp namessym_obj = {CHARSXP, 5, (void*)"names", NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
SEXP R_DimSymbol = &dimsym_obj;
SEXP R_NamesSymbol = &namessym_obj;
int R_NaInt = (int)0x80000000;
double R_NaReal;

static SEXP all_objects = NULL;
static long n_live = 0;
static int protect_depth = 0, protect_underflow = 0;
static struct transient { struct transient* next; } *transients = NULL;
static jmp_buf* error_jmp = NULL;
static char error_msg[1024];
static const R_CallMethodDef* call_table = NULL;
static struct rstub_dllinfo the_dll = {1};

static void init_na(void) {
  union { double d; uint64_t u; } v;
  v.u = 0x7FF00000000007A2ULL;           /* R's NA_real_ payload (1954) */
  R_NaReal = v.d;
}

/* ---- the R API the glue uses ------------------------------------------------------------------- */
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) {
  size_t es = 0;
  switch (type) {
    case REALSXP: es = sizeof(double); break;
    case INTSXP: case LGLSXP: es = sizeof(int); break;
    case VECSXP: case STRSXP: es = sizeof(SEXP); break;
    case CHARSXP: es = 1; break;
    default: Rf_error("rstub: allocVector of unsupported type %u", type);
  }
  if (n < 0) Rf_error("rstub: negative length");
  SEXP s = (SEXP)calloc(1, sizeof(struct rstub_sexp));
  if (!s) Rf_error("rstub: out of memory");
  s->type = (int)type;
  s->len = n;
  s->data = calloc((size_t)n + 1, es);
  if (!s->data) { free(s); Rf_error("cannot allocate vector of size %.1f Gb", (double)n * (double)es / 1073741824.0); }
  if (type == VECSXP || type == STRSXP)
    for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  s->next = all_objects;
  all_objects = s;
  ++n_live;
  return s;
}
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol) {
  if (nrow < 0 || ncol < 0) Rf_error("negative extents to matrix");
  SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
  SEXP d = Rf_allocVector(INTSXP, 2);
  INTEGER(d)[0] = nrow;
  INTEGER(d)[1] = ncol;
  s->dim = d;
  return s;
}
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) {
  protect_depth -= n;
  if (protect_depth < 0) { protect_underflow = 1; protect_depth = 0; }
}
int TYPEOF(SEXP s) { return s->type; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int LENGTH(SEXP s) { return (int)s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
double* REAL(SEXP s) {
  if (s->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type-%d object", s->type);
  return (double*)s->data;
}
int* INTEGER(SEXP s) {
  if (s->type != INTSXP && s->type != LGLSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type-%d object", s->type);
  return (int*)s->data;
}
int* LOGICAL(SEXP s) {
  if (s->type != LGLSXP) Rf_error("LOGICAL() can only be applied to a 'logical', not a type-%d object", s->type);
  return (int*)s->data;
}
SEXP VECTOR_ELT(SEXP list, R_xlen_t i) {
  if (list->type != VECSXP) Rf_error("VECTOR_ELT() can only be applied to a 'list', not a type-%d object", list->type);
  if (i < 0 || i >= list->len) Rf_error("rstub: VECTOR_ELT index out of range");
  return ((SEXP*)list->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP list, R_xlen_t i, SEXP v) {
  if (list->type != VECSXP) Rf_error("SET_VECTOR_ELT() can only be applied to a 'list'");
  if (i < 0 || i >= list->len) Rf_error("rstub: SET_VECTOR_ELT index out of range");
  ((SEXP*)list->data)[i] = v;
  return v;
}
SEXP STRING_ELT(SEXP s, R_xlen_t i) {
  if (s->type != STRSXP) Rf_error("STRING_ELT() can only be applied to a 'character vector', not a type-%d object", s->type);
  if (i < 0 || i >= s->len) Rf_error("rstub: STRING_ELT index out of range");
  return ((SEXP*)s->data)[i];
}
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) {
  if (s->type != STRSXP || i < 0 || i >= s->len) Rf_error("rstub: bad SET_STRING_ELT");
  ((SEXP*)s->data)[i] = v;
}
const char* CHAR(SEXP c) {
  if (c->type != CHARSXP) Rf_error("CHAR() can only be applied to a 'CHARSXP'");
  return (const char*)c->data;
}
SEXP Rf_mkChar(const char* str) {
  const size_t n = strlen(str);
  SEXP c = Rf_allocVector(CHARSXP, (R_xlen_t)n);
  memcpy(c->data, str, n + 1);
  return c;
}
SEXP Rf_mkString(const char* str) {
  SEXP s = Rf_allocVector(STRSXP, 1);
  SET_STRING_ELT(s, 0, Rf_mkChar(str));
  return s;
}
Rboolean Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP ? TRUE : FALSE; }
Rboolean Rf_isReal(SEXP s) { return s->type == REALSXP ? TRUE : FALSE; }
Rboolean Rf_isInteger(SEXP s) { return s->type == INTSXP ? TRUE : FALSE; }
int Rf_asInteger(SEXP s) {
  if (s->len < 1) return NA_INTEGER;
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) {
    const double v = ((double*)s->data)[0];
    if (v != v || v >= 2147483648.0 || v <= -2147483649.0) return NA_INTEGER;
    return (int)v;
  }
  return NA_INTEGER;
}
double Rf_asReal(SEXP s) {
  if (s->len < 1) return NA_REAL;
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) {
    const int v = ((int*)s->data)[0];
    return v == NA_INTEGER ? NA_REAL : (double)v;
  }
  return NA_REAL;
}
int Rf_asLogical(SEXP s) {
  if (s->len < 1) return NA_LOGICAL;
  if (s->type == LGLSXP || s->type == INTSXP) {
    const int v = ((int*)s->data)[0];
    return v == NA_INTEGER ? NA_LOGICAL : (v != 0);
  }
  if (s->type == REALSXP) {
    const double v = ((double*)s->data)[0];
    return v != v ? NA_LOGICAL : (v != 0.0);
  }
  return NA_LOGICAL;
}
SEXP Rf_getAttrib(SEXP s, SEXP name) {
  if (name == R_DimSymbol) return s->dim ? s->dim : R_NilValue;
  if (name == R_NamesSymbol) return s->names ? s->names : R_NilValue;
  return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP val) {
  if (name == R_DimSymbol) s->dim = Rf_isNull(val) ? NULL : val;
  else if (name == R_NamesSymbol) s->names = Rf_isNull(val) ? NULL : val;
  else Rf_error("rstub: attribute not supported");
  return val;
}
char* R_alloc(size_t n, int size) {
  struct transient* t = (struct transient*)malloc(sizeof(struct transient) + n * (size_t)size + 16);
  if (!t) Rf_error("cannot allocate memory block of size %.1f Gb", (double)n * size / 1073741824.0);
  t->next = transients;
  transients = t;
  return (char*)(t + 1);
}
void R_CheckUserInterrupt(void) {}
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_msg, sizeof(error_msg), fmt, ap);
  va_end(ap);
  if (error_jmp) longjmp(*error_jmp, 1);
  fprintf(stderr, "rstub: Rf_error outside .Call: %s\n", error_msg);
  abort();
}
void Rf_warning(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "Warning message: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
int R_registerRoutines(DllInfo* info, const R_CMethodDef* const c, const R_CallMethodDef* const call,
                       const R_FortranMethodDef* const f, const R_ExternalMethodDef* const e) {
  (void)info; (void)c; (void)f; (void)e;
  call_table = call;
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) {
  Rboolean old = info->dynamic_symbols ? TRUE : FALSE;
  info->dynamic_symbols = value;
  return old;
}
Rboolean R_forceSymbols(DllInfo* info, Rboolean value) { (void)info; (void)value; return FALSE; }

/* ---- harness side (ctypes) ----------------------------------------------------------------------- */
typedef void (*init_fn)(DllInfo*);
/* what library(pkg) does after dlopen: run R_init_<pkg>; returns the number of registered .Call routines */
int rstub_load(init_fn init) {
  init_na();
  call_table = NULL;
  the_dll.dynamic_symbols = 1;
  init(&the_dll);
  int n = 0;
  if (call_table) while (call_table[n].name) ++n;
  return n;
}
int rstub_dynamic_symbols(void) { return the_dll.dynamic_symbols; }
const char* rstub_routine_name(int i) { return call_table[i].name; }
int rstub_routine_nargs(int i) { return call_table[i].numArgs; }

SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_new(int type, R_xlen_t n) {
  jmp_buf jb;
  error_jmp = &jb;
  if (setjmp(jb)) { error_jmp = NULL; return NULL; }
  SEXP s = Rf_allocVector((SEXPTYPE)type, n);
  error_jmp = NULL;
  return s;
}
SEXP rstub_new_string(const char* str) {
  jmp_buf jb;
  error_jmp = &jb;
  if (setjmp(jb)) { error_jmp = NULL; return NULL; }
  SEXP s = Rf_mkString(str);
  error_jmp = NULL;
  return s;
}
void rstub_set_dim(SEXP s, int nrow, int ncol) {
  SEXP d = rstub_new(INTSXP, 2);
  ((int*)d->data)[0] = nrow;
  ((int*)d->data)[1] = ncol;
  s->dim = d;
}
int rstub_type(SEXP s) { return s->type; }
R_xlen_t rstub_len(SEXP s) { return s->len; }
void* rstub_data(SEXP s) { return s->data; }
int rstub_get_dim(SEXP s, int* nrow, int* ncol) {
  if (!s->dim) return 0;
  *nrow = ((int*)s->dim->data)[0];
  *ncol = ((int*)s->dim->data)[1];
  return 1;
}
SEXP rstub_elt(SEXP list, R_xlen_t i) { return ((SEXP*)list->data)[i]; }
void rstub_set_elt(SEXP list, R_xlen_t i, SEXP v) { ((SEXP*)list->data)[i] = v; }
long rstub_live_objects(void) { return n_live; }

/* frees every vector allocated so far (the harness keeps no SEXP across this) */
void rstub_gc(void) {
  while (all_objects) {
    SEXP nx = all_objects->next;
    free(all_objects->data);
    free(all_objects);
    all_objects = nx;
  }
  n_live = 0;
}

static void free_transients(void) {
  while (transients) {
    struct transient* nx = transients->next;
    free(transients);
    transients = nx;
  }
}

/* .Call(name, args...): NULL + message in errbuf when the routine raised an R error (or is not
 * registered / called with the wrong number of arguments, as R reports it).
 * status: 0 ok, 1 R error, 2 PROTECT stack imbalance after a normal return */
SEXP rstub_dot_call(const char* name, int nargs, SEXP* a, char* errbuf, int errlen, int* status) {
  *status = 1;
  if (errlen > 0) errbuf[0] = 0;
  const R_CallMethodDef* r = NULL;
  if (call_table)
    for (const R_CallMethodDef* t = call_table; t->name; ++t)
      if (!strcmp(t->name, name)) { r = t; break; }
  if (!r) {
    snprintf(errbuf, (size_t)errlen, "C symbol name \"%s\" not in DLL for package", name);
    return NULL;
  }
  if (r->numArgs != nargs) {
    snprintf(errbuf, (size_t)errlen, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, r->numArgs, name);
    return NULL;
  }
  jmp_buf jb;
  error_jmp = &jb;
  protect_depth = 0;
  protect_underflow = 0;
  SEXP res = NULL;
  if (setjmp(jb)) {
    error_jmp = NULL;
    free_transients();
    snprintf(errbuf, (size_t)errlen, "%s", error_msg);
    return NULL;
  }
  DL_FUNC f = r->fun;
  switch (nargs) {
    case 0: res = ((SEXP(*)(void))f)(); break;
    case 1: res = ((SEXP(*)(SEXP))f)(a[0]); break;
    case 2: res = ((SEXP(*)(SEXP, SEXP))f)(a[0], a[1]); break;
    case 3: res = ((SEXP(*)(SEXP, SEXP, SEXP))f)(a[0], a[1], a[2]); break;
    case 4: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3]); break;
    case 5: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 7: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
    case 8: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
    case 9: res = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
    default:
      error_jmp = NULL;
      snprintf(errbuf, (size_t)errlen, "rstub: too many arguments");
      return NULL;
  }
  error_jmp = NULL;
  free_transients();
  if (protect_depth != 0 || protect_underflow) {
    snprintf(errbuf, (size_t)errlen, "stack imbalance in '.Call', %d then %d", 0, protect_depth);
    *status = 2;
    return res;
  }
  *status = 0;
  return res;
}
