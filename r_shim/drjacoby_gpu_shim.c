/*
 * drjacoby_gpu_shim.c -- the .Call shim that binds libdrjacoby_b200.so into the reference R package.
 *
 * It replaces, for the GPU path, the Rcpp-generated entry
 *     RcppExport SEXP _drjacoby_main_cpp(SEXP argsSEXP)            src/RcppExports.cpp:15-24
 * and is registered the same way (src/RcppExports.cpp:25-33).  Difference in granularity: the
 * reference calls main_cpp once per chain from deploy_chain (R/main.R:583-610), chains spread over a PSOCK
 * cluster (R/main.R:392-402); these entries receive ALL chains in one call and spread them over the GPUs of the
 * box inside the library (drj_run_mcmc_multi / drj_multi_*: one host thread per GPU; the worker threads never
 * touch the R API, the progress / interrupt hook runs on the calling thread).
 *
 *   drjgpu_main(args)     -- every draw comes back: a list with one element per chain, each the 9-field list of
 *                            src/main.cpp:268-276, so the post-processing of R/main.R:411-577 runs unchanged.
 *   drjgpu_summary(args)  -- for runs whose draws cannot be returned whole (one REALSXP per stored iteration and
 *                            rung is what the reference's format costs): the draws stay on the GPUs; returns the
 *                            on-device reductions (per-chain means / variances, autocovariance sums of the
 *                            concatenated chains, deviance moments, quantiles) plus, if args_params$thin is given,
 *                            every thin-th draw per chain in the 9-field format.
 *
 * `args$args_params` exactly as built at R/main.R:348-366 plus
 *     init_mat   numeric matrix d x chains (R/main.R:322-328)
 *     chains     integer
 *     seed       numeric (Philox key) ; replay numeric vector or NULL (the uniforms a CPU run would draw)
 *     model      list(builtin = "example") or list(source = "<CUDA text>", loglike = "loglike", logprior = "logprior")
 *     devices    integer vector of CUDA ordinals, or NULL = every visible GPU
 *     (summary)  max_lag, probs (quantile probabilities, <= 8), thin
 *
 * NOT compiled against a real R in this repository's image (no R headers); tests/test_gpu_r_shim.py compiles it
 * against a functional mock of the R API (tests/r_mock) and EXECUTES it on the GPU.  Build on a machine with R:
 *     R CMD SHLIB drjacoby_gpu_shim.c -I../include -L../drjacoby_b200 -ldrjacoby_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>

#include "drjacoby_b200.h"

static SEXP list_get(SEXP list, const char* name) {
  if (list == R_NilValue) return R_NilValue;
  SEXP names = getAttrib(list, R_NamesSymbol);
  if (names == R_NilValue) return R_NilValue;
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
    if (len[k] > 0) memcpy(vals + off[k], REAL(v), sizeof(double) * (size_t)len[k]);
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
    if (iters > 0) memcpy(REAL(v), base + (size_t)r * iters, sizeof(double) * (size_t)iters);
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

/* everything both entries need from args$args_params; all storage is R_alloc'ed or owned by the R objects of `args`
 * (which the caller keeps alive for the duration of the .Call) */
typedef struct shim_run {
  drj_config c;
  drj_list data, misc;
  drj_model* model;
  const int32_t* devices;
  int n_devices;
  int nprotect;
} shim_run;

static void parse_args(SEXP args, shim_run* r) {
  memset(r, 0, sizeof(*r));
  SEXP p = list_need(args, "args_params");
  drj_config* c = &r->c;
  const int d = (int)XLENGTH(list_need(p, "theta_min"));
  c->abi_version = DRJ_ABI_VERSION;
  c->device = 0;
  c->d = d;
  c->chains = asInteger(list_need(p, "chains"));
  c->rungs = asInteger(list_need(p, "rungs"));
  c->burnin = asInteger(list_need(p, "burnin"));
  c->samples = asInteger(list_need(p, "samples"));
  c->n_block = asInteger(list_need(p, "n_block"));
  c->coupling_on = asLogical(list_need(p, "coupling_on"));
  c->save_hot_draws = asLogical(list_need(p, "save_hot_draws"));
  c->store_pt = 1;
  c->target_acceptance = asReal(list_need(p, "target_acceptance"));
  SEXP tmin = PROTECT(coerceVector(list_need(p, "theta_min"), REALSXP));
  SEXP tmax = PROTECT(coerceVector(list_need(p, "theta_max"), REALSXP));
  r->nprotect += 2;
  if ((int)XLENGTH(tmax) != d) error("drjacoby_b200: theta_min and theta_max differ in length");
  c->theta_min = REAL(tmin);
  c->theta_max = REAL(tmax);
  /* trans_type arrives as a double vector (2*is.finite(min) + is.finite(max), R/main.R:293) */
  SEXP tt = PROTECT(coerceVector(list_need(p, "trans_type"), INTSXP));
  r->nprotect += 1;
  if ((int)XLENGTH(tt) != d) error("drjacoby_b200: trans_type must have one entry per parameter");
  c->trans_type = INTEGER(tt);
  SEXP sk = PROTECT(coerceVector(list_need(p, "skip_param"), LGLSXP));
  r->nprotect += 1;
  if ((int)XLENGTH(sk) != d) error("drjacoby_b200: skip_param must have one entry per parameter");
  uint8_t* skip = (uint8_t*)R_alloc(d, 1);
  for (int i = 0; i < d; ++i) skip[i] = (uint8_t)(LOGICAL(sk)[i] != 0);
  c->skip_param = skip;
  /* df_params$block: list of numeric vectors, 1-based -> CSR */
  SEXP bl = list_need(p, "block");
  if ((int)XLENGTH(bl) != d) error("drjacoby_b200: block must have one entry per parameter");
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
  c->block_ptr = bptr; c->block_idx = bidx;
  SEXP br = PROTECT(coerceVector(list_need(p, "beta_raised"), REALSXP));
  r->nprotect += 1;
  if ((int)XLENGTH(br) != c->rungs) error("drjacoby_b200: beta_raised must have one entry per rung");
  c->beta_raised = REAL(br);
  /* init_mat is d x chains column-major == [chains][d] row-major */
  SEXP im = PROTECT(coerceVector(list_need(p, "init_mat"), REALSXP));
  r->nprotect += 1;
  if (XLENGTH(im) != (R_xlen_t)d * c->chains) error("drjacoby_b200: init_mat must be a d x chains matrix");
  c->theta_init = REAL(im);
  c->seed = (uint64_t)asReal(list_need(p, "seed"));
  SEXP replay = list_get(p, "replay");
  c->rng_mode = (replay != R_NilValue) ? DRJ_RNG_REPLAY : DRJ_RNG_PHILOX;
  c->replay_uniforms = (replay != R_NilValue) ? REAL(replay) : NULL;
  if (replay != R_NilValue) {
    int d_free = 0;
    for (int i = 0; i < d; ++i) d_free += skip[i] ? 0 : 1;
    const double U = 3.0 * d_free * c->rungs + ((c->coupling_on && c->rungs > 1) ? (c->rungs - 1) : 0);
    if ((double)XLENGTH(replay) != (double)c->chains * (c->burnin - 1 + c->samples) * U)
      error("drjacoby_b200: replay must hold chains * (burnin - 1 + samples) * U uniforms (U = %g per iteration)", U);
  }

  flatten(list_need(p, "x"), 0, &r->data);
  flatten(list_get(p, "misc"), 1, &r->misc);

  /* devices: NULL = every visible GPU */
  SEXP dv = list_get(p, "devices");
  if (dv != R_NilValue) {
    SEXP di = PROTECT(coerceVector(dv, INTSXP));
    r->nprotect += 1;
    r->n_devices = (int)XLENGTH(di);
    int32_t* devs = (int32_t*)R_alloc(r->n_devices + 1, sizeof(int32_t));
    for (int i = 0; i < r->n_devices; ++i) devs[i] = INTEGER(di)[i];
    r->devices = devs;
  }

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
  if (mb == R_NilValue && ms == R_NilValue) error("drjacoby_b200: model needs `builtin` or `source` (R functions cannot run on the GPU path)");
  SEXP pn = getAttrib(list_need(p, "theta_min"), R_NamesSymbol);
  const char** pnames = (const char**)R_alloc(d, sizeof(char*));
  for (int i = 0; i < d; ++i) pnames[i] = (pn != R_NilValue) ? CHAR(STRING_ELT(pn, i)) : "";
  md.param_names = pnames; md.data_names = r->data.names; md.misc_names = r->misc.names;
  md.n_data_items = r->data.n_items; md.n_misc_items = r->misc.n_items; md.d = d; md.n_block = c->n_block;

  drj_error err;
  if (drj_model_create(&md, r->devices ? r->devices[0] : 0, &r->model, &err) != DRJ_OK) error("%s", err.message);
}

/* the per-chain lists of src/main.cpp:268-276 from chain-major buffers holding nb / ns stored iterations per series */
static SEXP chains_to_r(const drj_config* c, const drj_output* o, int nb, int ns) {
  static const char* fields[] = {"loglike_burnin", "logprior_burnin", "theta_burnin", "loglike_sampling",
                                 "logprior_sampling", "theta_sampling", "mc_accept_burnin", "mc_accept_sampling",
                                 "t_diff", ""};
  const int chains = c->chains, rungs = c->rungs, d = c->d, thr = c->save_hot_draws ? rungs : 1;
  SEXP out = PROTECT(allocVector(VECSXP, chains));
  for (int ch = 0; ch < chains; ++ch) {
    SEXP one = PROTECT(mkNamed(VECSXP, fields));
    SET_VECTOR_ELT(one, 0, chain_matrix_list(o->loglike_burnin + (size_t)ch * rungs * nb, rungs, nb));
    SET_VECTOR_ELT(one, 1, chain_matrix_list(o->logprior_burnin + (size_t)ch * rungs * nb, rungs, nb));
    SET_VECTOR_ELT(one, 2, chain_theta_list(o->theta_burnin + (size_t)ch * thr * nb * d, thr, nb, d));
    SET_VECTOR_ELT(one, 3, chain_matrix_list(o->loglike_sampling + (size_t)ch * rungs * ns, rungs, ns));
    SET_VECTOR_ELT(one, 4, chain_matrix_list(o->logprior_sampling + (size_t)ch * rungs * ns, rungs, ns));
    SET_VECTOR_ELT(one, 5, chain_theta_list(o->theta_sampling + (size_t)ch * thr * ns * d, thr, ns, d));
    SEXP mb1 = PROTECT(allocVector(INTSXP, rungs - 1)), mb2 = PROTECT(allocVector(INTSXP, rungs - 1));
    for (int i = 0; i < rungs - 1; ++i) {
      INTEGER(mb1)[i] = o->mc_accept_burnin[(size_t)ch * (rungs - 1) + i];
      INTEGER(mb2)[i] = o->mc_accept_sampling[(size_t)ch * (rungs - 1) + i];
    }
    SET_VECTOR_ELT(one, 6, mb1);
    SET_VECTOR_ELT(one, 7, mb2);
    SET_VECTOR_ELT(one, 8, ScalarReal(o->t_diff[ch]));
    SET_VECTOR_ELT(out, ch, one);
    UNPROTECT(3);
  }
  UNPROTECT(1);
  return out;
}

static void alloc_draws(const drj_config* c, drj_output* o, int nb, int ns) {
  const size_t chains = (size_t)c->chains, rungs = (size_t)c->rungs, d = (size_t)c->d, thr = c->save_hot_draws ? rungs : 1;
  memset(o, 0, sizeof(*o));
  o->loglike_burnin = (double*)R_alloc(chains * rungs * nb + 1, sizeof(double));
  o->logprior_burnin = (double*)R_alloc(chains * rungs * nb + 1, sizeof(double));
  o->theta_burnin = (double*)R_alloc(chains * thr * nb * d + 1, sizeof(double));
  o->loglike_sampling = (double*)R_alloc(chains * rungs * ns + 1, sizeof(double));
  o->logprior_sampling = (double*)R_alloc(chains * rungs * ns + 1, sizeof(double));
  o->theta_sampling = (double*)R_alloc(chains * thr * ns * d + 1, sizeof(double));
  o->mc_accept_burnin = (int32_t*)R_alloc(chains * (rungs > 1 ? rungs - 1 : 1), sizeof(int32_t));
  o->mc_accept_sampling = (int32_t*)R_alloc(chains * (rungs > 1 ? rungs - 1 : 1), sizeof(int32_t));
  o->accept_burnin = (int32_t*)R_alloc(chains * rungs, sizeof(int32_t));
  o->accept_sampling = (int32_t*)R_alloc(chains * rungs, sizeof(int32_t));
  o->t_diff = (double*)R_alloc(chains, sizeof(double));
}

SEXP drjgpu_main(SEXP args) {
  shim_run r;
  parse_args(args, &r);
  drj_output o;
  alloc_draws(&r.c, &o, r.c.burnin, r.c.samples);
  drj_error err;
  int rc = drj_run_mcmc_multi(&r.c, r.devices, r.n_devices, r.model, &r.data, &r.misc, &o, progress_cb, NULL, &err);
  drj_model_destroy(r.model);
  if (rc != DRJ_OK) error("%s", err.message); /* Rcpp::stop -> Rf_error, src/RcppExports.cpp:16,22 (the library has returned: nothing is live) */
  SEXP out = chains_to_r(&r.c, &o, r.c.burnin, r.c.samples);
  UNPROTECT(r.nprotect);
  return out;
}

SEXP drjgpu_summary(SEXP args) {
  shim_run r;
  parse_args(args, &r);
  SEXP p = list_need(args, "args_params");
  const drj_config* c = &r.c;
  const int d = c->d, chains = c->chains, rungs = c->rungs;
  SEXP s_lag = list_get(p, "max_lag"), s_probs = list_get(p, "probs"), s_thin = list_get(p, "thin");
  const int L = (s_lag != R_NilValue) ? asInteger(s_lag) : 0;
  const int thin = (s_thin != R_NilValue) ? asInteger(s_thin) : 0;
  int nq = 0;
  const double* probs = NULL;
  if (s_probs != R_NilValue) {
    SEXP pr = PROTECT(coerceVector(s_probs, REALSXP));
    r.nprotect += 1;
    nq = (int)XLENGTH(pr);
    probs = REAL(pr);
  }
  if (L < 0 || thin < 0) { drj_model_destroy(r.model); error("drjacoby_b200: max_lag and thin must be non-negative"); }

  static const char* fields[] = {"chain_mean", "chain_var", "autocov_sum", "deviance_mean_var", "quantiles", "accept_burnin",
                                 "accept_sampling", "mc_accept_burnin", "mc_accept_sampling", "t_diff", "chains", ""};
  SEXP out = PROTECT(mkNamed(VECSXP, fields));
  r.nprotect += 1;
  /* matrices are filled in R's column-major order: chain_mean[d x chains] etc. take the library's row-major [chains][d] as is */
  SEXP cm = PROTECT(allocMatrix(REALSXP, d, chains)), cv = PROTECT(allocMatrix(REALSXP, d, chains));
  SEXP ac = PROTECT(allocMatrix(REALSXP, L + 1, d)), dev = PROTECT(allocVector(REALSXP, 2));
  SEXP qu = PROTECT(allocMatrix(REALSXP, nq, d));
  r.nprotect += 5;
  SET_VECTOR_ELT(out, 0, cm); SET_VECTOR_ELT(out, 1, cv); SET_VECTOR_ELT(out, 2, ac); SET_VECTOR_ELT(out, 3, dev); SET_VECTOR_ELT(out, 4, qu);

  drj_error err;
  drj_multi* m = NULL;
  int rc = drj_multi_create(c, r.devices, r.n_devices, r.model, &r.data, &r.misc, &m, &err);
  /* burn-in then sampling in ~1 % batches, the interrupt check between batches on this (the R main) thread */
  for (int phase = 0; phase < 2 && rc == DRJ_OK; ++phase) {
    const int total = phase == 0 ? c->burnin - 1 : c->samples;
    const int step = total > 100 ? (total + 99) / 100 : (total > 0 ? total : 1);
    for (int done = 0; done < total && rc == DRJ_OK; done += step) {
      rc = drj_multi_advance(m, (total - done < step) ? total - done : step, NULL, &err);
      if (rc == DRJ_OK && progress_cb(NULL, phase, done, total)) {
        rc = DRJ_ERR_INTERRUPT;
        strcpy(err.message, "interrupted");
      }
    }
  }
  drj_output o;
  memset(&o, 0, sizeof(o));
  int nb = 0, ns = 0;
  if (rc == DRJ_OK) {
    drj_summary sm;
    memset(&sm, 0, sizeof(sm));
    sm.max_lag = L;
    sm.chain_mean = REAL(cm); sm.chain_var = REAL(cv); sm.autocov_sum = REAL(ac); sm.deviance_mean_var = REAL(dev);
    sm.n_quantiles = nq; sm.quantile_probs = probs; sm.quantiles = nq ? REAL(qu) : NULL;
    rc = drj_multi_summary(m, &sm, &err);
  }
  if (rc == DRJ_OK) {
    if (thin > 0) {
      nb = (c->burnin + thin - 1) / thin; ns = (c->samples + thin - 1) / thin;
      alloc_draws(c, &o, nb, ns);
    } else {  /* counters only: the draws stay on the devices */
      o.mc_accept_burnin = (int32_t*)R_alloc((size_t)chains * (rungs > 1 ? rungs - 1 : 1), sizeof(int32_t));
      o.mc_accept_sampling = (int32_t*)R_alloc((size_t)chains * (rungs > 1 ? rungs - 1 : 1), sizeof(int32_t));
      o.accept_burnin = (int32_t*)R_alloc((size_t)chains * rungs, sizeof(int32_t));
      o.accept_sampling = (int32_t*)R_alloc((size_t)chains * rungs, sizeof(int32_t));
      o.t_diff = (double*)R_alloc(chains, sizeof(double));
    }
    rc = drj_multi_download_thinned(m, thin > 0 ? thin : 1, &o, &err);
  }
  drj_multi_destroy(m);
  drj_model_destroy(r.model);
  if (rc != DRJ_OK) error("%s", err.message);

  SEXP ab = PROTECT(allocMatrix(INTSXP, rungs, chains)), as = PROTECT(allocMatrix(INTSXP, rungs, chains));
  SEXP mb = PROTECT(allocMatrix(INTSXP, rungs > 1 ? rungs - 1 : 0, chains)), ms = PROTECT(allocMatrix(INTSXP, rungs > 1 ? rungs - 1 : 0, chains));
  SEXP td = PROTECT(allocVector(REALSXP, chains));
  r.nprotect += 5;
  for (int i = 0; i < chains * rungs; ++i) { INTEGER(ab)[i] = o.accept_burnin[i]; INTEGER(as)[i] = o.accept_sampling[i]; }
  for (int i = 0; i < chains * (rungs - 1); ++i) { INTEGER(mb)[i] = o.mc_accept_burnin[i]; INTEGER(ms)[i] = o.mc_accept_sampling[i]; }
  for (int i = 0; i < chains; ++i) REAL(td)[i] = o.t_diff[i];
  SET_VECTOR_ELT(out, 5, ab); SET_VECTOR_ELT(out, 6, as); SET_VECTOR_ELT(out, 7, mb); SET_VECTOR_ELT(out, 8, ms); SET_VECTOR_ELT(out, 9, td);
  if (thin > 0) SET_VECTOR_ELT(out, 10, chains_to_r(c, &o, nb, ns));
  UNPROTECT(r.nprotect);
  return out;
}

static const R_CallMethodDef CallEntries[] = {{"drjgpu_main", (DL_FUNC)&drjgpu_main, 1},
                                              {"drjgpu_summary", (DL_FUNC)&drjgpu_summary, 1},
                                              {NULL, NULL, 0}};

void R_init_drjacobygpu(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
