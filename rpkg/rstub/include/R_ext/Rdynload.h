/* Stand-in for R's <R_ext/Rdynload.h> (see Rinternals.h in this directory): routine registration. */
#ifndef RSTUB_RDYNLOAD_H
#define RSTUB_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct {
  const char* name;
  DL_FUNC fun;
  int numArgs;
} R_CallMethodDef;
typedef R_CallMethodDef R_CMethodDef;
typedef R_CallMethodDef R_FortranMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
typedef struct rstub_dllinfo DllInfo;
int R_registerRoutines(DllInfo* info, const R_CMethodDef* const croutines, const R_CallMethodDef* const callRoutines,
                       const R_FortranMethodDef* const fortranRoutines,
                       const R_ExternalMethodDef* const externalRoutines);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
Rboolean R_forceSymbols(DllInfo* info, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif
