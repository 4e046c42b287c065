/*
 * sampler.c -- ORACLE (test infrastructure, see drj_oracle.h).
 *
 * Restatement of the reference's MCMC core, one chain at a time, rungs serially, exactly in the
 * reference's order of operations and RNG consumption:
 *   particle_init        src/Particle.cpp:8-44     (Particle::init, theta_to_phi :78-99)
 *   particle_init_like   src/Particle.h:59-79      (Particle::init_like)
 *   particle_update      src/Particle.h:82-171     (Particle::update; propose_phi Particle.cpp:48-50,
 *                                                   phi_prop_to_theta_prop :55-74, get_adjustment :103-124)
 *   coupling             src/main.cpp:282-344
 *   run_chain            src/main.cpp:84-278       (run_mcmc<T1,T2>: burn-in rep 1..burnin-1,
 *                                                   sampling rep 0..samples-1, store-then-couple)
 *   seq_sum              src/misc.h:20-27
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "drj_oracle.h"
#include "models.h"

typedef struct {
  double *theta, *theta_prop, *phi, *phi_prop, *bw;
  int* bw_index;
  double bw_stepsize;
  double *loglike_block, *loglike_prop_block;
  double loglike, logprior;
  int accept_count;
  double beta_raised;
} particle;

typedef struct {
  int mode;
  ora_rmt* mt;
  const double* replay; /* this chain's [T][U] block */
  int64_t U;            /* uniforms per iteration */
  int d_free;
  int rungs;
  uint64_t seed;
  int64_t chain; /* global chain id */
} chain_rng;

static double seq_sum(const double* v, int n) {
  double ret = 0;
  for (int i = 0; i < n; ++i) ret += v[i];
  return ret;
}

/* three uniforms for the update of parameter `param` (free slot `k`) of `rung` at iteration `t` */
static void draw_update(chain_rng* g, int t, int rung, int param, int k, double u[3], int need_norm) {
  if (g->mode == ORA_RNG_RMT) {
    if (need_norm) {
      u[0] = ora_rmt_unif_rand(g->mt);
      u[1] = ora_rmt_unif_rand(g->mt);
    }
    /* u[2] is drawn later, after the user functions (Particle.h:139) -- order is immaterial
       here because the oracle's models never touch the RNG */
    u[2] = ora_rmt_unif_rand(g->mt);
  } else if (g->mode == ORA_RNG_REPLAY) {
    const double* p = g->replay + (int64_t)(t - 1) * g->U + 3 * ((int64_t)rung * g->d_free + k);
    u[0] = p[0]; u[1] = p[1]; u[2] = p[2];
  } else {
    double w[4];
    ora_philox_uniforms(g->seed, g->chain, (uint32_t)t, (uint32_t)rung, (uint32_t)param, 0u, w);
    u[0] = w[0]; u[1] = w[1]; u[2] = w[2];
  }
}

static double draw_coupling(chain_rng* g, int t, int pair) {
  if (g->mode == ORA_RNG_RMT) return ora_rmt_unif_rand(g->mt);
  if (g->mode == ORA_RNG_REPLAY)
    return g->replay[(int64_t)(t - 1) * g->U + 3 * (int64_t)g->d_free * g->rungs + pair];
  double w[4];
  ora_philox_uniforms(g->seed, g->chain, (uint32_t)t, 0u, (uint32_t)pair, 1u, w);
  return w[0];
}

static int theta_to_phi(const ora_config* c, particle* p) {
  for (int i = 0; i < c->d; ++i) {
    switch (c->trans_type[i]) {
      case 0: p->phi[i] = p->theta[i]; break;
      case 1: p->phi[i] = log(c->theta_max[i] - p->theta[i]); break;
      case 2: p->phi[i] = log(p->theta[i] - c->theta_min[i]); break;
      case 3: p->phi[i] = log(p->theta[i] - c->theta_min[i]) - log(c->theta_max[i] - p->theta[i]); break;
      default: return ORA_ERR_TRANS_TYPE;
    }
  }
  return ORA_OK;
}

static int particle_init(const ora_config* c, particle* p, const double* theta_init, double beta_raised) {
  int d = c->d, B = c->n_block;
  p->theta = malloc(sizeof(double) * d);
  p->theta_prop = malloc(sizeof(double) * d);
  p->phi = malloc(sizeof(double) * d);
  p->phi_prop = calloc(d, sizeof(double));
  p->bw = malloc(sizeof(double) * d);
  p->bw_index = malloc(sizeof(int) * d);
  p->loglike_block = calloc(B, sizeof(double));
  p->loglike_prop_block = calloc(B, sizeof(double));
  p->beta_raised = beta_raised;
  memcpy(p->theta, theta_init, sizeof(double) * d);
  memcpy(p->theta_prop, theta_init, sizeof(double) * d);
  int rc = theta_to_phi(c, p);
  for (int i = 0; i < d; ++i) { p->bw[i] = 1.0; p->bw_index[i] = 1; }
  p->bw_stepsize = 1.0;
  p->loglike = 0; p->logprior = 0;
  p->accept_count = 0;
  return rc;
}

static void particle_free(particle* p) {
  free(p->theta); free(p->theta_prop); free(p->phi); free(p->phi_prop); free(p->bw);
  free(p->bw_index); free(p->loglike_block); free(p->loglike_prop_block);
}

static void set_err(ora_error* e, int code, int64_t chain, int rung, int param, int iter,
                    const double* theta, int d, const char* msg) {
  e->code = code; e->chain = chain; e->rung = rung; e->param = param; e->iteration = iter;
  for (int i = 0; i < 64; ++i) e->theta[i] = (i < d) ? theta[i] : 0.0;
  snprintf(e->message, sizeof(e->message), "%s", msg);
}

static int particle_init_like(const ora_config* c, const ora_model* m, const ora_list* data,
                              const ora_list* misc, particle* p) {
  for (int i = 0; i < c->d; ++i)
    for (int j = c->block_ptr[i]; j < c->block_ptr[i + 1]; ++j) {
      int this_block = c->block_idx[j];
      p->loglike_block[this_block - 1] = m->loglike(p->theta, c->d, data, misc, this_block);
    }
  p->loglike = seq_sum(p->loglike_block, c->n_block);
  p->logprior = m->logprior(p->theta, c->d, misc);
  if (p->loglike == -INFINITY || p->logprior == -INFINITY) return ORA_ERR_INIT_NEGINF;
  return ORA_OK;
}

static int particle_update(const ora_config* c, const ora_model* m, const ora_list* data,
                           const ora_list* misc, particle* p, chain_rng* g, int t, int rung,
                           int* bad_param) {
  int d = c->d, B = c->n_block;
  memcpy(p->theta_prop, p->theta, sizeof(double) * d);
  memcpy(p->phi_prop, p->phi, sizeof(double) * d);
  int k = -1;
  for (int i = 0; i < d; ++i) {
    if (c->skip_param[i]) continue;
    ++k;
    /* propose_phi: R::rnorm(phi[i], bw[i]) */
    double mu = p->phi[i], sigma = p->bw[i];
    int degenerate = (isnan(mu) || !isfinite(sigma) || sigma < 0) ? 2 : ((sigma == 0 || !isfinite(mu)) ? 1 : 0);
    double u[3];
    draw_update(g, t, rung, i, k, u, degenerate == 0);
    if (degenerate == 2) p->phi_prop[i] = NAN;
    else if (degenerate == 1) p->phi_prop[i] = mu;
    else {
      const double BIG = 134217728.0;
      double uu = (int)(BIG * u[0]) + u[1];
      p->phi_prop[i] = mu + sigma * ora_qnorm(uu / BIG);
    }
    /* phi_prop_to_theta_prop */
    double tmin = c->theta_min[i], tmax = c->theta_max[i];
    switch (c->trans_type[i]) {
      case 0: p->theta_prop[i] = p->phi_prop[i]; break;
      case 1: p->theta_prop[i] = tmax - exp(p->phi_prop[i]); break;
      case 2: p->theta_prop[i] = exp(p->phi_prop[i]) + tmin; break;
      case 3: p->theta_prop[i] = (tmax * exp(p->phi_prop[i]) + tmin) / (1 + exp(p->phi_prop[i])); break;
      default: return ORA_ERR_TRANS_TYPE;
    }
    /* get_adjustment */
    double adj = 0;
    switch (c->trans_type[i]) {
      case 0: break;
      case 1: adj = log(tmax - p->theta_prop[i]) - log(tmax - p->theta[i]); break;
      case 2: adj = log(p->theta_prop[i] - tmin) - log(p->theta[i] - tmin); break;
      case 3:
        adj = log(tmax - p->theta_prop[i]) + log(p->theta_prop[i] - tmin) - log(tmax - p->theta[i]) -
              log(p->theta[i] - tmin);
        break;
    }
    /* block-wise likelihood of the proposal */
    memcpy(p->loglike_prop_block, p->loglike_block, sizeof(double) * B);
    for (int j = c->block_ptr[i]; j < c->block_ptr[i + 1]; ++j) {
      int this_block = c->block_idx[j];
      p->loglike_prop_block[this_block - 1] = m->loglike(p->theta_prop, d, data, misc, this_block);
    }
    double loglike_prop = seq_sum(p->loglike_prop_block, B);
    double logprior_prop = m->logprior(p->theta_prop, d, misc);
    if (isnan(loglike_prop) || loglike_prop == INFINITY) { *bad_param = i; return ORA_ERR_LOGLIKE_BAD; }
    if (isnan(logprior_prop) || logprior_prop == INFINITY) { *bad_param = i; return ORA_ERR_LOGPRIOR_BAD; }
    /* Metropolis-Hastings ratio */
    double MH;
    if (p->beta_raised == 0.0) MH = (logprior_prop - p->logprior) + adj;
    else MH = p->beta_raised * (loglike_prop - p->loglike) + (logprior_prop - p->logprior) + adj;
    int accept = (log(u[2]) < MH);
    if (accept) {
      p->theta[i] = p->theta_prop[i];
      p->phi[i] = p->phi_prop[i];
      memcpy(p->loglike_block, p->loglike_prop_block, sizeof(double) * B);
      p->loglike = loglike_prop;
      p->logprior = logprior_prop;
      p->bw[i] = exp(log(p->bw[i]) + p->bw_stepsize * (1 - c->target_acceptance) / sqrt((double)p->bw_index[i]));
      p->bw_index[i]++;
      p->accept_count++;
    } else {
      p->theta_prop[i] = p->theta[i];
      p->phi_prop[i] = p->phi[i];
      p->bw[i] = exp(log(p->bw[i]) - p->bw_stepsize * c->target_acceptance / sqrt((double)p->bw_index[i]));
      p->bw_index[i]++;
    }
  }
  return ORA_OK;
}

static void swap_ptr(double** a, double** b) { double* t = *a; *a = *b; *b = t; }
static void swap_dbl(double* a, double* b) { double t = *a; *a = *b; *b = t; }

static void coupling(particle* pv, int rungs, int* mc_accept, chain_rng* g, int t) {
  for (int i = 0; i < rungs - 1; ++i) {
    double ll1 = pv[i].loglike, ll2 = pv[i + 1].loglike;
    double b1 = pv[i].beta_raised, b2 = pv[i + 1].beta_raised;
    double acceptance;
    if (b1 == 0.0) acceptance = (ll1 * b2) - (ll2 * b2);
    else acceptance = (ll2 * b1 + ll1 * b2) - (ll1 * b1 + ll2 * b2);
    double u = draw_coupling(g, t, i);
    if (log(u) < acceptance) {
      swap_ptr(&pv[i].theta, &pv[i + 1].theta);
      swap_ptr(&pv[i].phi, &pv[i + 1].phi);
      swap_ptr(&pv[i].loglike_block, &pv[i + 1].loglike_block);
      swap_dbl(&pv[i].loglike, &pv[i + 1].loglike);
      swap_dbl(&pv[i].logprior, &pv[i + 1].logprior);
      mc_accept[i]++;
    }
  }
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int run_chain(const ora_config* c, const ora_model* m, const ora_list* data, const ora_list* misc,
                     int ch, ora_output* out, ora_error* err) {
  double t0 = now_s();
  int d = c->d, R = c->rungs, burnin = c->burnin, samples = c->samples;
  int theta_rungs = c->save_hot_draws ? R : 1;
  int d_free = 0;
  for (int i = 0; i < d; ++i) d_free += c->skip_param[i] ? 0 : 1;
  int64_t T = (int64_t)burnin - 1 + samples;
  chain_rng g;
  g.mode = c->rng_mode; g.mt = c->rmt; g.d_free = d_free; g.rungs = R;
  g.U = 3 * (int64_t)d_free * R + ((c->coupling_on && R > 1) ? (R - 1) : 0);
  g.replay = c->replay_uniforms ? c->replay_uniforms + (int64_t)ch * T * g.U : NULL;
  g.seed = c->seed; g.chain = c->chain_offset + ch;

  double* ll_b = out->loglike_burnin + (int64_t)ch * R * burnin;
  double* lp_b = out->logprior_burnin + (int64_t)ch * R * burnin;
  double* th_b = out->theta_burnin + (int64_t)ch * theta_rungs * burnin * d;
  double* ll_s = out->loglike_sampling + (int64_t)ch * R * samples;
  double* lp_s = out->logprior_sampling + (int64_t)ch * R * samples;
  double* th_s = out->theta_sampling + (int64_t)ch * theta_rungs * samples * d;
  int* mc_b = out->mc_accept_burnin + (int64_t)ch * (R - 1);
  int* mc_s = out->mc_accept_sampling + (int64_t)ch * (R - 1);
  for (int i = 0; i < R - 1; ++i) { mc_b[i] = 0; mc_s[i] = 0; }

  particle* pv = calloc(R, sizeof(particle));
  int rc = ORA_OK;
  for (int r = 0; r < R && rc == ORA_OK; ++r) {
    rc = particle_init(c, &pv[r], c->theta_init + (int64_t)ch * d, c->beta_raised[r]);
    if (rc != ORA_OK) { set_err(err, rc, g.chain, r, -1, 0, pv[r].theta, d, "trans_type invalid"); break; }
    rc = particle_init_like(c, m, data, misc, &pv[r]);
    if (rc != ORA_OK)
      set_err(err, rc, g.chain, r, -1, 0, pv[r].theta, d,
              "Starting values result in -Inf in likelihood or prior. Consider setting inital values in the parameters data.frame.");
  }
  if (rc != ORA_OK) goto done;

  /* first stored row = initial values (main.cpp:128-137) */
  for (int r = 0; r < R; ++r) {
    ll_b[(int64_t)r * burnin] = pv[r].loglike;
    lp_b[(int64_t)r * burnin] = pv[r].logprior;
    if (c->save_hot_draws) memcpy(th_b + ((int64_t)r * burnin) * d, pv[r].theta, sizeof(double) * d);
  }
  if (!c->save_hot_draws) memcpy(th_b, pv[R - 1].theta, sizeof(double) * d);

  int t = 0; /* global update index, 1-based */
  for (int rep = 1; rep < burnin; ++rep) {
    ++t;
    for (int r = 0; r < R; ++r) {
      int bad = -1;
      rc = particle_update(c, m, data, misc, &pv[r], &g, t, r, &bad);
      if (rc != ORA_OK) {
        set_err(err, rc, g.chain, r, bad, t, pv[r].theta_prop, d,
                rc == ORA_ERR_LOGLIKE_BAD ? "NA, NaN or Inf in likelihood" : "NA, NaN or Inf in prior");
        goto done;
      }
      ll_b[(int64_t)r * burnin + rep] = pv[r].loglike;
      lp_b[(int64_t)r * burnin + rep] = pv[r].logprior;
      if (c->save_hot_draws) memcpy(th_b + ((int64_t)r * burnin + rep) * d, pv[r].theta, sizeof(double) * d);
    }
    if (!c->save_hot_draws) memcpy(th_b + (int64_t)rep * d, pv[R - 1].theta, sizeof(double) * d);
    if (c->coupling_on) coupling(pv, R, mc_b, &g, t);
  }
  for (int r = 0; r < R; ++r) {
    if (out->accept_burnin) out->accept_burnin[(int64_t)ch * R + r] = pv[r].accept_count;
    pv[r].accept_count = 0; /* main.cpp:208-210; bw and bw_index carry over */
  }
  for (int rep = 0; rep < samples; ++rep) {
    ++t;
    for (int r = 0; r < R; ++r) {
      int bad = -1;
      rc = particle_update(c, m, data, misc, &pv[r], &g, t, r, &bad);
      if (rc != ORA_OK) {
        set_err(err, rc, g.chain, r, bad, t, pv[r].theta_prop, d,
                rc == ORA_ERR_LOGLIKE_BAD ? "NA, NaN or Inf in likelihood" : "NA, NaN or Inf in prior");
        goto done;
      }
      ll_s[(int64_t)r * samples + rep] = pv[r].loglike;
      lp_s[(int64_t)r * samples + rep] = pv[r].logprior;
      if (c->save_hot_draws) memcpy(th_s + ((int64_t)r * samples + rep) * d, pv[r].theta, sizeof(double) * d);
    }
    if (!c->save_hot_draws) memcpy(th_s + (int64_t)rep * d, pv[R - 1].theta, sizeof(double) * d);
    if (c->coupling_on) coupling(pv, R, mc_s, &g, t);
  }
  for (int r = 0; r < R; ++r) {
    if (out->accept_sampling) out->accept_sampling[(int64_t)ch * R + r] = pv[r].accept_count;
    if (out->bw_final) memcpy(out->bw_final + ((int64_t)ch * R + r) * d, pv[r].bw, sizeof(double) * d);
  }
done:
  for (int r = 0; r < R; ++r)
    if (pv[r].theta) particle_free(&pv[r]);
  free(pv);
  if (out->t_diff) out->t_diff[ch] = now_s() - t0;
  return rc;
}

/* chains are independent (R/main.R:392-402): one chain at a time per worker thread, dynamic
 * hand-out -- the analogue of parallel::clusterApplyLB over the host cores */
typedef struct {
  const ora_config* c; const ora_model* m; const ora_list* data; const ora_list* misc;
  ora_output* out; ora_error* errs; int next;
} pool_job;

static void* pool_worker(void* arg) {
  pool_job* j = (pool_job*)arg;
  for (;;) {
    int ch = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (ch >= j->c->chains) break;
    run_chain(j->c, j->m, j->data, j->misc, ch, j->out, &j->errs[ch]);
  }
  return NULL;
}

int ora_run_mcmc(const ora_config* c, const char* model_name, const ora_list* data, const ora_list* misc,
                 ora_output* out, ora_error* err) {
  ora_error dummy;
  if (!err) err = &dummy;
  memset(err, 0, sizeof(*err));
  const ora_model* m = ora_find_model(model_name);
  if (!m) {
    err->code = ORA_ERR_MODEL;
    snprintf(err->message, sizeof(err->message), "oracle: unknown model '%s'", model_name);
    return ORA_ERR_MODEL;
  }
  if (c->d < 1 || c->rungs < 1 || c->chains < 1 || c->burnin < 1 || c->samples < 1) {
    err->code = ORA_ERR_ARG;
    snprintf(err->message, sizeof(err->message), "oracle: bad sizes");
    return ORA_ERR_ARG;
  }
  if (c->rng_mode == ORA_RNG_RMT) {
    /* serial, one continuing stream (lapply over chains, R/main.R:401) */
    for (int ch = 0; ch < c->chains; ++ch) {
      int rc = run_chain(c, m, data, misc, ch, out, err);
      if (rc != ORA_OK) return rc;
    }
    return ORA_OK;
  }
  ora_error* errs = calloc(c->chains, sizeof(ora_error));
  int nthreads = c->n_threads > 0 ? c->n_threads : 1;
  if (nthreads > c->chains) nthreads = c->chains;
  pool_job job = {c, m, data, misc, out, errs, 0};
  if (nthreads <= 1) {
    pool_worker(&job);
  } else {
    pthread_t* th = malloc(sizeof(pthread_t) * nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, pool_worker, &job);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
  }
  int rc = ORA_OK;
  for (int ch = 0; ch < c->chains; ++ch)
    if (errs[ch].code != ORA_OK) { *err = errs[ch]; rc = errs[ch].code; break; }
  free(errs);
  return rc;
}
