/* Minimal mock of the parts of R's C API that r_shim/drjacoby_gpu_shim.c uses.  TEST ONLY: lets gcc
 * type-check the shim in an image that has no R.  Declarations follow the public R API signatures. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define FALSE 0
#define TRUE 1
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol;
SEXP Rf_getAttrib(SEXP, SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char* CHAR(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isReal(SEXP), Rf_isInteger(SEXP), Rf_isLogical(SEXP);
SEXP Rf_coerceVector(SEXP, unsigned int);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
void Rf_error(const char*, ...) __attribute__((noreturn));
char* R_alloc(size_t, int);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
#define getAttrib Rf_getAttrib
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define isReal Rf_isReal
#define isInteger Rf_isInteger
#define isLogical Rf_isLogical
#define coerceVector Rf_coerceVector
#define allocVector Rf_allocVector
#define mkNamed Rf_mkNamed
#define ScalarReal Rf_ScalarReal
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define error Rf_error
#endif
