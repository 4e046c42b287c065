/* Functional mock of the parts of R's C API that r_shim/drjacoby_gpu_shim.c uses (implemented in rmock.c).
 * TEST ONLY: this image has no R, so the shim is compiled against these declarations -- which follow the public
 * R API signatures -- and EXECUTED against rmock.c by tests/test_gpu_r_shim.py, which builds the `args` list as
 * R/main.R:348-386 does and walks the returned list as R/main.R:411-482 does. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define FALSE 0
#define TRUE 1
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define CHARSXP 9
#define NA_INTEGER (-2147483647 - 1)
#define NA_LOGICAL NA_INTEGER
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
int TYPEOF(SEXP);
R_xlen_t XLENGTH(SEXP);
int LENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_mkChar(const char*);
SEXP Rf_mkString(const char*);
const char* CHAR(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isReal(SEXP), Rf_isInteger(SEXP), Rf_isLogical(SEXP), Rf_isString(SEXP), Rf_isNull(SEXP), Rf_isNewList(SEXP);
SEXP Rf_coerceVector(SEXP, unsigned int);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarLogical(int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
char* R_alloc(size_t, int);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
#define getAttrib Rf_getAttrib
#define setAttrib Rf_setAttrib
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define isReal Rf_isReal
#define isInteger Rf_isInteger
#define isLogical Rf_isLogical
#define isString Rf_isString
#define isNull Rf_isNull
#define isNewList Rf_isNewList
#define coerceVector Rf_coerceVector
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define mkNamed Rf_mkNamed
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define ScalarReal Rf_ScalarReal
#define ScalarInteger Rf_ScalarInteger
#define ScalarLogical Rf_ScalarLogical
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define error Rf_error
#define warning Rf_warning
#endif
