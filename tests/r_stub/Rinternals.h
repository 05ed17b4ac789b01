/* see R.h in this directory */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
double* REAL(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asLogical(SEXP x);
Rboolean Rf_isNull(SEXP x);
int Rf_length(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_mkChar(const char* s);
SEXP Rf_ScalarReal(double x);
SEXP Rf_ScalarInteger(int x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#endif
