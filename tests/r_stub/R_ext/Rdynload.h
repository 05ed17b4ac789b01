/* see ../R.h */
#ifndef R_STUB_RDYNLOAD_H
#define R_STUB_RDYNLOAD_H
#include "../Rinternals.h"
typedef void* (*DL_FUNC)(void);
typedef struct _DllInfo DllInfo;
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char* name; DL_FUNC fun; int numArgs; void* types; } R_CMethodDef;
typedef R_CMethodDef R_FortranMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
int R_registerRoutines(DllInfo* info, const R_CMethodDef* c, const R_CallMethodDef* call, const R_FortranMethodDef* f,
                       const R_ExternalMethodDef* ext);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#endif
