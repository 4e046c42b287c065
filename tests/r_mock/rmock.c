/* rmock.c -- a small functional stand-in for libR's vector API (TEST ONLY; see Rinternals.h).
 *
 * Objects live in an arena that rmock_reset() frees (R's garbage collector is not modelled: PROTECT / UNPROTECT only
 * keep a balance counter that the tests assert on).  Rf_error() long-jumps to the frame set up by rmock_call1(),
 * exactly the control flow of a real .Call (R's error unwinds out of the C routine).  R_alloc() memory belongs to
 * the arena too.  R_CheckUserInterrupt() raises an "interrupt" (long jump out of R_ToplevelExec, which then returns
 * FALSE) when the test has armed rmock_interrupt_after(n): n-th check from now.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
  int type;
  R_xlen_t length;
  void* data;      /* double / int / SEXP array, or char* for CHARSXP */
  SEXP names;      /* names attribute (STRSXP) or NULL */
  SEXP dim;        /* dim attribute (INTSXP) or NULL */
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL};
static struct SEXPREC names_sym = {-1, 0, NULL, NULL, NULL};
static struct SEXPREC dim_sym = {-2, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec;
SEXP R_NamesSymbol = &names_sym;
SEXP R_DimSymbol = &dim_sym;

/* ---- arena */
static void** arena = NULL;
static size_t arena_n = 0, arena_cap = 0;
static int protect_balance = 0, protect_max = 0;
static long interrupt_countdown = -1;
static int in_toplevel = 0;
static jmp_buf toplevel_jmp;
static jmp_buf call_jmp;
static int call_active = 0;
static char error_message[1024];
static const R_CallMethodDef* registered = NULL;
static int dynamic_symbols = 1;

static void* arena_alloc(size_t bytes) {
  void* p = calloc(1, bytes ? bytes : 1);
  if (!p) { fprintf(stderr, "rmock: out of memory\n"); abort(); }
  if (arena_n == arena_cap) {
    arena_cap = arena_cap ? arena_cap * 2 : 1024;
    arena = (void**)realloc(arena, arena_cap * sizeof(void*));
  }
  arena[arena_n++] = p;
  return p;
}

void rmock_reset(void) {
  for (size_t i = 0; i < arena_n; ++i) free(arena[i]);
  arena_n = 0;
  protect_balance = protect_max = 0;
  interrupt_countdown = -1;
}
size_t rmock_live_objects(void) { return arena_n; }
int rmock_protect_balance(void) { return protect_balance; }
void rmock_interrupt_after(long n) { interrupt_countdown = n; }

/* ---- vectors */
static size_t elt_size(unsigned int type) {
  switch (type) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case STRSXP: case VECSXP: return sizeof(SEXP);
    default: return 0;
  }
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  if (n < 0 || !elt_size(type)) Rf_error("rmock: allocVector(type %u, length %ld)", type, (long)n);
  SEXP s = (SEXP)arena_alloc(sizeof(struct SEXPREC));
  s->type = (int)type; s->length = n;
  s->data = arena_alloc((size_t)n * elt_size(type));
  if (type == VECSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  if (type == STRSXP) { SEXP e = Rf_mkChar(""); for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = e; }
  return s;
}
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
  SEXP d = Rf_allocVector(INTSXP, 2);
  INTEGER(d)[0] = nrow; INTEGER(d)[1] = ncol;
  s->dim = d;
  return s;
}
SEXP Rf_mkChar(const char* c) {
  SEXP s = (SEXP)arena_alloc(sizeof(struct SEXPREC));
  s->type = CHARSXP; s->length = (R_xlen_t)strlen(c);
  s->data = arena_alloc(strlen(c) + 1);
  memcpy(s->data, c, strlen(c) + 1);
  return s;
}
SEXP Rf_mkString(const char* c) { SEXP s = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(s, 0, Rf_mkChar(c)); return s; }
SEXP Rf_mkNamed(unsigned int type, const char** names) {
  R_xlen_t n = 0;
  while (names[n][0]) ++n;
  SEXP s = Rf_allocVector(type, n), nm = Rf_allocVector(STRSXP, n);
  for (R_xlen_t i = 0; i < n; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  s->names = nm;
  return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }

int TYPEOF(SEXP s) { return s->type; }
R_xlen_t XLENGTH(SEXP s) { return s->length; }
int LENGTH(SEXP s) { return (int)s->length; }
static void need(SEXP s, int type, const char* what) {
  if (!s || s->type != type) Rf_error("rmock: %s applied to an object of type %d", what, s ? s->type : -99);
}
double* REAL(SEXP s) { need(s, REALSXP, "REAL()"); return (double*)s->data; }
int* INTEGER(SEXP s) { if (s->type != INTSXP && s->type != LGLSXP) need(s, INTSXP, "INTEGER()"); return (int*)s->data; }
int* LOGICAL(SEXP s) { need(s, LGLSXP, "LOGICAL()"); return (int*)s->data; }
const char* CHAR(SEXP s) { need(s, CHARSXP, "CHAR()"); return (const char*)s->data; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) {
  need(s, VECSXP, "VECTOR_ELT()");
  if (i < 0 || i >= s->length) Rf_error("rmock: VECTOR_ELT index %ld out of bounds (length %ld)", (long)i, (long)s->length);
  return ((SEXP*)s->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) {
  need(s, VECSXP, "SET_VECTOR_ELT()");
  if (i < 0 || i >= s->length) Rf_error("rmock: SET_VECTOR_ELT index %ld out of bounds (length %ld)", (long)i, (long)s->length);
  ((SEXP*)s->data)[i] = v;
  return v;
}
SEXP STRING_ELT(SEXP s, R_xlen_t i) {
  need(s, STRSXP, "STRING_ELT()");
  if (i < 0 || i >= s->length) Rf_error("rmock: STRING_ELT index %ld out of bounds (length %ld)", (long)i, (long)s->length);
  return ((SEXP*)s->data)[i];
}
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) {
  need(s, STRSXP, "SET_STRING_ELT()");
  need(v, CHARSXP, "SET_STRING_ELT() value");
  if (i < 0 || i >= s->length) Rf_error("rmock: SET_STRING_ELT index out of bounds");
  ((SEXP*)s->data)[i] = v;
}
SEXP Rf_getAttrib(SEXP s, SEXP sym) {
  if (s == R_NilValue) return R_NilValue;
  if (sym == R_NamesSymbol) return s->names ? s->names : R_NilValue;
  if (sym == R_DimSymbol) return s->dim ? s->dim : R_NilValue;
  return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP sym, SEXP v) {
  if (sym == R_NamesSymbol) { need(v, STRSXP, "names<-"); if (v->length != s->length) Rf_error("rmock: names of the wrong length"); s->names = v; }
  else if (sym == R_DimSymbol) s->dim = v;
  return v;
}
Rboolean Rf_isReal(SEXP s) { return s->type == REALSXP; }
Rboolean Rf_isInteger(SEXP s) { return s->type == INTSXP; }
Rboolean Rf_isLogical(SEXP s) { return s->type == LGLSXP; }
Rboolean Rf_isString(SEXP s) { return s->type == STRSXP; }
Rboolean Rf_isNull(SEXP s) { return s->type == NILSXP; }
Rboolean Rf_isNewList(SEXP s) { return s->type == NILSXP || s->type == VECSXP; }

int Rf_asInteger(SEXP s) {
  if (s->length < 1) return NA_INTEGER;
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) { double v = ((double*)s->data)[0]; return (v != v) ? NA_INTEGER : (int)v; }
  return NA_INTEGER;
}
int Rf_asLogical(SEXP s) {
  if (s->length < 1) return NA_LOGICAL;
  if (s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == INTSXP) { int v = ((int*)s->data)[0]; return v == NA_INTEGER ? NA_LOGICAL : v != 0; }
  if (s->type == REALSXP) { double v = ((double*)s->data)[0]; return (v != v) ? NA_LOGICAL : v != 0.0; }
  return NA_LOGICAL;
}
double Rf_asReal(SEXP s) {
  if (s->length < 1) return 0.0 / 0.0;
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) { int v = ((int*)s->data)[0]; return v == NA_INTEGER ? 0.0 / 0.0 : (double)v; }
  return 0.0 / 0.0;
}
SEXP Rf_coerceVector(SEXP s, unsigned int type) {
  if ((unsigned)s->type == type) return s;  /* R returns the object itself when nothing has to change */
  if (s->type != REALSXP && s->type != INTSXP && s->type != LGLSXP) Rf_error("rmock: cannot coerce type %d to %u", s->type, type);
  SEXP out = Rf_allocVector(type, s->length);
  out->names = s->names;
  for (R_xlen_t i = 0; i < s->length; ++i) {
    double v;
    if (s->type == REALSXP) v = ((double*)s->data)[i];
    else { int q = ((int*)s->data)[i]; v = q == NA_INTEGER ? 0.0 / 0.0 : (double)q; }
    if (type == REALSXP) ((double*)out->data)[i] = v;
    else if (type == INTSXP) ((int*)out->data)[i] = (v != v) ? NA_INTEGER : (int)v;  /* truncation towards zero, as R */
    else if (type == LGLSXP) ((int*)out->data)[i] = (v != v) ? NA_LOGICAL : v != 0.0;
    else Rf_error("rmock: cannot coerce to type %u", type);
  }
  return out;
}
SEXP Rf_protect(SEXP s) { if (++protect_balance > protect_max) protect_max = protect_balance; return s; }
void Rf_unprotect(int n) {
  protect_balance -= n;
  if (protect_balance < 0) { fprintf(stderr, "rmock: UNPROTECT below zero (stack imbalance)\n"); abort(); }
}
char* R_alloc(size_t n, int size) { return (char*)arena_alloc(n * (size_t)size); }

void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_message, sizeof(error_message), fmt, ap);
  va_end(ap);
  if (!call_active) { fprintf(stderr, "rmock: Rf_error outside a call: %s\n", error_message); abort(); }
  protect_balance = 0;  /* R resets the protect stack when an error unwinds to top level */
  longjmp(call_jmp, 1);
}
void Rf_warning(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
void R_CheckUserInterrupt(void) {
  if (interrupt_countdown > 0 && --interrupt_countdown == 0) {
    interrupt_countdown = -1;
    if (in_toplevel) longjmp(toplevel_jmp, 1);  /* the user pressed Esc: R unwinds to the enclosing top-level context */
    Rf_error("rmock: interrupt outside R_ToplevelExec (would long-jump across the caller's frames)");
  }
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {
  volatile Rboolean ok = TRUE;
  jmp_buf saved;
  const int was = in_toplevel;
  memcpy(&saved, &toplevel_jmp, sizeof(jmp_buf));
  in_toplevel = 1;
  if (setjmp(toplevel_jmp) == 0) fun(data);
  else ok = FALSE;
  in_toplevel = was;
  memcpy(&toplevel_jmp, &saved, sizeof(jmp_buf));
  return ok;
}

int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)dll; (void)c; (void)f; (void)e;
  registered = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int v) { (void)dll; dynamic_symbols = v; return 1; }

/* ---- what the test driver (Python, ctypes) uses: .Call by registered name, builders, accessors */
typedef SEXP (*call1_fn)(SEXP);
/* .Call("name", arg): looks the routine up in the registration table (R_useDynamicSymbols(dll, FALSE) means nothing
 * else is callable), runs it under an error frame; returns NULL and fills errbuf if it raised an R error */
SEXP rmock_dot_call(const char* name, SEXP arg, char* errbuf, int errlen) {
  call1_fn volatile fn = NULL;
  for (const R_CallMethodDef* d = registered; d && d->name; ++d)
    if (strcmp(d->name, name) == 0 && d->numArgs == 1) fn = (call1_fn)d->fun;
  if (!fn) { snprintf(errbuf, (size_t)errlen, "rmock: \"%s\" not available for .Call()", name); return NULL; }
  SEXP volatile out = NULL;
  call_active = 1;
  if (setjmp(call_jmp) == 0) out = fn(arg);
  else { snprintf(errbuf, (size_t)errlen, "%s", error_message); out = NULL; }
  call_active = 0;
  return out;
}
int rmock_dynamic_symbols(void) { return dynamic_symbols; }
SEXP rmock_real(const double* v, long n) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, sizeof(double) * (size_t)n); return s; }
SEXP rmock_int(const int* v, long n) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP rmock_lgl(const int* v, long n) { SEXP s = Rf_allocVector(LGLSXP, n); if (n) memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP rmock_matrix(const double* v, int nrow, int ncol) { SEXP s = Rf_allocMatrix(REALSXP, nrow, ncol); memcpy(s->data, v, sizeof(double) * (size_t)nrow * ncol); return s; }
SEXP rmock_str(const char* c) { return Rf_mkString(c); }
SEXP rmock_list(long n) { return Rf_allocVector(VECSXP, n); }
SEXP rmock_nil(void) { return R_NilValue; }
void rmock_set_names(SEXP s, const char** names) {
  SEXP nm = Rf_allocVector(STRSXP, s->length);
  for (R_xlen_t i = 0; i < s->length; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  s->names = nm;
}
void rmock_list_set(SEXP l, long i, SEXP v) { SET_VECTOR_ELT(l, i, v); }
int rmock_type(SEXP s) { return s->type; }
long rmock_length(SEXP s) { return (long)s->length; }
SEXP rmock_elt(SEXP s, long i) { return ((SEXP*)s->data)[i]; }
void* rmock_data(SEXP s) { return s->data; }
const char* rmock_name(SEXP s, long i) { return s->names ? (const char*)((SEXP*)s->names->data)[i]->data : ""; }
int rmock_nrow(SEXP s) { return s->dim ? ((int*)s->dim->data)[0] : -1; }
int rmock_ncol(SEXP s) { return s->dim ? ((int*)s->dim->data)[1] : -1; }
