/*
 * drjacoby_gpu_shim.c -- the .Call shim that binds libdrjacoby_b200.so into the reference R package.
 *
 * It replaces, for the GPU path, the Rcpp-generated entry
 *     RcppExport SEXP _drjacoby_main_cpp(SEXP argsSEXP)            src/RcppExports.cpp:15-24
 * and is registered the same way (src/RcppExports.cpp:25-33).  Difference in granularity: the
 * reference calls main_cpp once per chain from deploy_chain (R/main.R:583-610); this entry receives
 * ALL chains in one call -- `args$args_params` exactly as built at R/main.R:348-366 plus
 *     init_mat   numeric matrix d x chains (R/main.R:322-328)
 *     chains     integer
 *     seed       numeric (Philox key) ; rng "philox" | "replay" ; replay numeric vector or NULL
 *     model      list(builtin = "example") or list(source = "<CUDA text>", loglike = "loglike", logprior = "logprior")
 * and returns a list with one element per chain, each the 9-field list of src/main.cpp:268-276, so the
 * post-processing of R/main.R:411-577 runs unchanged.
 *
 * NOT compiled in this repository's image (no R headers); tests/test_r_shim_syntax.py compiles it
 * against a minimal mock of the R API to keep it syntactically honest.  Build on a machine with R:
 *     R CMD SHLIB drjacoby_gpu_shim.c -I../include -L../drjacoby_b200 -ldrjacoby_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>

#include "drjacoby_b200.h"

static SEXP list_get(SEXP list, const char* name) {
  SEXP names = getAttrib(list, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

static SEXP list_need(SEXP list, const char* name) {
  SEXP v = list_get(list, name);
  if (v == R_NilValue) error("drjacoby_b200: args_params$%s is missing", name);
  return v;
}

/* named list of numeric vectors -> drj_list (offsets/lengths/values owned by R_alloc) */
static void flatten(SEXP lst, int skip_block, drj_list* out) {
  R_xlen_t n = (lst == R_NilValue) ? 0 : XLENGTH(lst);
  SEXP names = (n > 0) ? getAttrib(lst, R_NamesSymbol) : R_NilValue;
  int64_t* off = (int64_t*)R_alloc(n + 1, sizeof(int64_t));
  int64_t* len = (int64_t*)R_alloc(n + 1, sizeof(int64_t));
  const char** nm = (const char**)R_alloc(n + 1, sizeof(char*));
  int64_t total = 0;
  int k = 0;
  for (R_xlen_t i = 0; i < n; ++i) {
    const char* name = (names != R_NilValue) ? CHAR(STRING_ELT(names, i)) : "";
    if (skip_block && strcmp(name, "block") == 0) continue; /* misc$block is the reserved slot, R/main.R:339-342 */
    SEXP v = VECTOR_ELT(lst, i);
    if (!(isReal(v) || isInteger(v) || isLogical(v)))
      error("drjacoby_b200: element '%s' must be numeric to be used on the GPU path", name);
    off[k] = total; len[k] = XLENGTH(v); nm[k] = name;
    total += len[k];
    ++k;
  }
  double* vals = (double*)R_alloc(total + 1, sizeof(double));
  k = 0;
  for (R_xlen_t i = 0; i < n; ++i) {
    const char* name = (names != R_NilValue) ? CHAR(STRING_ELT(names, i)) : "";
    if (skip_block && strcmp(name, "block") == 0) continue;
    SEXP v = PROTECT(coerceVector(VECTOR_ELT(lst, i), REALSXP));
    memcpy(vals + off[k], REAL(v), sizeof(double) * (size_t)len[k]);
    UNPROTECT(1);
    ++k;
  }
  out->n_items = k; out->names = nm; out->offsets = off; out->lengths = len; out->values = vals;
}

/* Rcpp::checkUserInterrupt() (src/main.cpp:156,216) long-jumps out of C on Esc; that must not happen
 * across live CUDA state, so the check runs inside R_ToplevelExec and an interrupt is handed back to
 * the library as "stop", which unwinds normally and returns DRJ_ERR_INTERRUPT. */
static void check_interrupt(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }
static int progress_cb(void* user, int phase, int done, int total) {
  (void)user; (void)phase; (void)done; (void)total;
  return R_ToplevelExec(check_interrupt, NULL) == FALSE;
}

static SEXP chain_matrix_list(const double* base, int rungs, int iters) {
  /* vector<vector<double>> [rungs][iters] -> list of numeric vectors */
  SEXP out = PROTECT(allocVector(VECSXP, rungs));
  for (int r = 0; r < rungs; ++r) {
    SEXP v = PROTECT(allocVector(REALSXP, iters));
    memcpy(REAL(v), base + (size_t)r * iters, sizeof(double) * (size_t)iters);
    SET_VECTOR_ELT(out, r, v);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}

static SEXP chain_theta_list(const double* base, int rungs, int iters, int d) {
  /* vector<vector<vector<double>>> [rungs][iters][d] -> list of lists of numeric vectors */
  SEXP out = PROTECT(allocVector(VECSXP, rungs));
  for (int r = 0; r < rungs; ++r) {
    SEXP lr = PROTECT(allocVector(VECSXP, iters));
    for (int t = 0; t < iters; ++t) {
      SEXP v = PROTECT(allocVector(REALSXP, d));
      memcpy(REAL(v), base + ((size_t)r * iters + t) * d, sizeof(double) * (size_t)d);
      SET_VECTOR_ELT(lr, t, v);
      UNPROTECT(1);
    }
    SET_VECTOR_ELT(out, r, lr);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}

SEXP drjgpu_main(SEXP args) {
  SEXP p = list_need(args, "args_params");
  const int d = (int)XLENGTH(list_need(p, "theta_min"));
  const int chains = asInteger(list_need(p, "chains"));
  const int rungs = asInteger(list_need(p, "rungs"));
  const int burnin = asInteger(list_need(p, "burnin"));
  const int samples = asInteger(list_need(p, "samples"));
  const int save_hot = asLogical(list_need(p, "save_hot_draws"));

  drj_config c;
  memset(&c, 0, sizeof(c));
  c.abi_version = DRJ_ABI_VERSION;
  c.device = 0;
  c.d = d; c.chains = chains; c.rungs = rungs; c.burnin = burnin; c.samples = samples;
  c.n_block = asInteger(list_need(p, "n_block"));
  c.coupling_on = asLogical(list_need(p, "coupling_on"));
  c.save_hot_draws = save_hot;
  c.store_pt = 1;
  c.target_acceptance = asReal(list_need(p, "target_acceptance"));
  c.theta_min = REAL(list_need(p, "theta_min"));
  c.theta_max = REAL(list_need(p, "theta_max"));
  /* trans_type arrives as a double vector (2*is.finite(min) + is.finite(max), R/main.R:293) */
  SEXP tt = PROTECT(coerceVector(list_need(p, "trans_type"), INTSXP));
  c.trans_type = INTEGER(tt);
  SEXP sk = list_need(p, "skip_param");
  uint8_t* skip = (uint8_t*)R_alloc(d, 1);
  for (int i = 0; i < d; ++i) skip[i] = (uint8_t)(LOGICAL(sk)[i] != 0);
  c.skip_param = skip;
  /* df_params$block: list of numeric vectors, 1-based -> CSR */
  SEXP bl = list_need(p, "block");
  int32_t* bptr = (int32_t*)R_alloc(d + 1, sizeof(int32_t));
  int nnz = 0;
  bptr[0] = 0;
  for (int i = 0; i < d; ++i) { nnz += (int)XLENGTH(VECTOR_ELT(bl, i)); bptr[i + 1] = nnz; }
  int32_t* bidx = (int32_t*)R_alloc(nnz + 1, sizeof(int32_t));
  for (int i = 0, k = 0; i < d; ++i) {
    SEXP b = PROTECT(coerceVector(VECTOR_ELT(bl, i), INTSXP));
    for (R_xlen_t j = 0; j < XLENGTH(b); ++j) bidx[k++] = INTEGER(b)[j];
    UNPROTECT(1);
  }
  c.block_ptr = bptr; c.block_idx = bidx;
  c.beta_raised = REAL(list_need(p, "beta_raised"));
  /* init_mat is d x chains column-major == [chains][d] row-major */
  c.theta_init = REAL(list_need(p, "init_mat"));
  c.seed = (uint64_t)asReal(list_need(p, "seed"));
  SEXP replay = list_get(p, "replay");
  c.rng_mode = (replay != R_NilValue) ? DRJ_RNG_REPLAY : DRJ_RNG_PHILOX;
  c.replay_uniforms = (replay != R_NilValue) ? REAL(replay) : NULL;

  drj_list data, misc;
  flatten(list_need(p, "x"), 0, &data);
  flatten(list_get(p, "misc"), 1, &misc);

  /* model: R closures never reach this point (the R wrapper stops first); only names / CUDA text */
  SEXP m = list_need(p, "model");
  drj_model_desc md;
  memset(&md, 0, sizeof(md));
  SEXP mb = list_get(m, "builtin"), ms = list_get(m, "source");
  if (mb != R_NilValue) md.builtin = CHAR(STRING_ELT(mb, 0));
  if (ms != R_NilValue) {
    md.cuda_source = CHAR(STRING_ELT(ms, 0));
    md.loglike_name = CHAR(STRING_ELT(list_need(m, "loglike"), 0));
    md.logprior_name = CHAR(STRING_ELT(list_need(m, "logprior"), 0));
  }
  SEXP pn = getAttrib(list_need(p, "theta_min"), R_NamesSymbol);
  const char** pnames = (const char**)R_alloc(d, sizeof(char*));
  for (int i = 0; i < d; ++i) pnames[i] = (pn != R_NilValue) ? CHAR(STRING_ELT(pn, i)) : "";
  md.param_names = pnames; md.data_names = data.names; md.misc_names = misc.names;
  md.n_data_items = data.n_items; md.n_misc_items = misc.n_items; md.d = d; md.n_block = c.n_block;

  drj_error err;
  drj_model* model = NULL;
  if (drj_model_create(&md, c.device, &model, &err) != DRJ_OK) { UNPROTECT(1); error("%s", err.message); }

  const int thr = save_hot ? rungs : 1;
  drj_output o;
  memset(&o, 0, sizeof(o));
  o.loglike_burnin = (double*)R_alloc((size_t)chains * rungs * burnin, sizeof(double));
  o.logprior_burnin = (double*)R_alloc((size_t)chains * rungs * burnin, sizeof(double));
  o.theta_burnin = (double*)R_alloc((size_t)chains * thr * burnin * d, sizeof(double));
  o.loglike_sampling = (double*)R_alloc((size_t)chains * rungs * samples, sizeof(double));
  o.logprior_sampling = (double*)R_alloc((size_t)chains * rungs * samples, sizeof(double));
  o.theta_sampling = (double*)R_alloc((size_t)chains * thr * samples * d, sizeof(double));
  o.mc_accept_burnin = (int32_t*)R_alloc((size_t)chains * (rungs > 1 ? rungs - 1 : 1), sizeof(int32_t));
  o.mc_accept_sampling = (int32_t*)R_alloc((size_t)chains * (rungs > 1 ? rungs - 1 : 1), sizeof(int32_t));
  o.t_diff = (double*)R_alloc(chains, sizeof(double));

  int rc = drj_run_mcmc(&c, model, &data, &misc, &o, progress_cb, NULL, &err);
  drj_model_destroy(model);
  if (rc != DRJ_OK) { UNPROTECT(1); error("%s", err.message); } /* Rcpp::stop -> Rf_error, src/RcppExports.cpp:16,22 */

  static const char* fields[] = {"loglike_burnin", "logprior_burnin", "theta_burnin", "loglike_sampling",
                                 "logprior_sampling", "theta_sampling", "mc_accept_burnin", "mc_accept_sampling",
                                 "t_diff", ""};
  SEXP out = PROTECT(allocVector(VECSXP, chains));
  for (int ch = 0; ch < chains; ++ch) {
    SEXP one = PROTECT(mkNamed(VECSXP, fields));
    SET_VECTOR_ELT(one, 0, chain_matrix_list(o.loglike_burnin + (size_t)ch * rungs * burnin, rungs, burnin));
    SET_VECTOR_ELT(one, 1, chain_matrix_list(o.logprior_burnin + (size_t)ch * rungs * burnin, rungs, burnin));
    SET_VECTOR_ELT(one, 2, chain_theta_list(o.theta_burnin + (size_t)ch * thr * burnin * d, thr, burnin, d));
    SET_VECTOR_ELT(one, 3, chain_matrix_list(o.loglike_sampling + (size_t)ch * rungs * samples, rungs, samples));
    SET_VECTOR_ELT(one, 4, chain_matrix_list(o.logprior_sampling + (size_t)ch * rungs * samples, rungs, samples));
    SET_VECTOR_ELT(one, 5, chain_theta_list(o.theta_sampling + (size_t)ch * thr * samples * d, thr, samples, d));
    SEXP mb1 = PROTECT(allocVector(INTSXP, rungs - 1)), mb2 = PROTECT(allocVector(INTSXP, rungs - 1));
    for (int i = 0; i < rungs - 1; ++i) {
      INTEGER(mb1)[i] = o.mc_accept_burnin[(size_t)ch * (rungs - 1) + i];
      INTEGER(mb2)[i] = o.mc_accept_sampling[(size_t)ch * (rungs - 1) + i];
    }
    SET_VECTOR_ELT(one, 6, mb1);
    SET_VECTOR_ELT(one, 7, mb2);
    SET_VECTOR_ELT(one, 8, ScalarReal(o.t_diff[ch]));
    SET_VECTOR_ELT(out, ch, one);
    UNPROTECT(3);
  }
  UNPROTECT(2);
  return out;
}

static const R_CallMethodDef CallEntries[] = {{"drjgpu_main", (DL_FUNC)&drjgpu_main, 1}, {NULL, NULL, 0}};

void R_init_drjacobygpu(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
