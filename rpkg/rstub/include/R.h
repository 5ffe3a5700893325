/* Stand-in for R's <R.h> (see Rinternals.h in this directory): transient allocation. */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif
/* memory released when the current .Call returns (normally or through Rf_error) */
char* R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
#ifdef __cplusplus
}
#endif
#endif
