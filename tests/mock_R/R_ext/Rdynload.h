/* mock of <R_ext/Rdynload.h>: see ../Rinternals.h (test infrastructure) */
#ifndef MOCK_RDYNLOAD_H
#define MOCK_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef R_CallMethodDef R_CMethodDef;
typedef struct mock_dllinfo DllInfo;
int R_registerRoutines(DllInfo* info, const R_CMethodDef* const cRoutines, const R_CallMethodDef* const callRoutines,
                       const void* const fortranRoutines, const void* const externalRoutines);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
Rboolean R_forceSymbols(DllInfo* info, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif
