/*
 * drj_oracle.h -- CPU ORACLE for the drjacoby MCMC sampling core.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker / the CPU baseline.  The product (drjacoby_b200/) never links or imports it.
 *
 * What it is: a plain-C restatement of the reference's algorithm
 *   src/Particle.h:59-171   (init_like, update)
 *   src/Particle.cpp:8-124  (init, propose_phi, phi_prop_to_theta_prop, theta_to_phi, get_adjustment)
 *   src/main.cpp:84-344     (run_mcmc loop, storage order, coupling)
 *   src/misc.h:20-27        (sequential sum)
 * plus the third-party arithmetic the reference links but does not vendor (R's nmath and default
 * RNG, base R, version unpinned by the reference: DESCRIPTION:41 "R (>= 3.1.0)"), restated from
 * their published algorithms (MT19937, Wichura AS241, Loader's saddle-point expansions).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this oracle against every exact
 * number the reference publishes for this path (docs/articles/example.html:275-330,359,384,
 * 520-555: six head() rows, 20 acceptance rates, Rhat, ESS) and the weak pins in
 * docs/articles/metropolis_coupling.html:284,287 and docs/articles/checks_double_well.html:203,206.
 * The reference itself (Rcpp) cannot be compiled here: no R, Rcpp.h or Rmath.h in this image.
 */
#ifndef DRJ_ORACLE_H
#define DRJ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Named list of numeric vectors flattened (the reference's `data` / `misc` Rcpp::List). */
typedef struct ora_list {
  int32_t n_items;
  const int64_t* offsets; /* [n_items] offset of item k in values */
  const int64_t* lengths; /* [n_items] */
  const double* values;
} ora_list;

enum { ORA_RNG_PHILOX = 0, ORA_RNG_REPLAY = 1, ORA_RNG_RMT = 2 };

/* R's default RNG state (Mersenne-Twister, Inversion). */
typedef struct ora_rmt {
  uint32_t mt[624];
  int32_t mti;
} ora_rmt;

/* Same field meaning as System (src/System.h:10-53) but batched over chains. */
typedef struct ora_config {
  int32_t d;
  int32_t n_block;
  int32_t chains;
  int32_t rungs;
  int32_t burnin;
  int32_t samples;
  int32_t coupling_on;
  int32_t save_hot_draws;
  double target_acceptance;
  const double* theta_min;    /* [d] */
  const double* theta_max;    /* [d] */
  const int32_t* trans_type;  /* [d] 0..3, R/main.R:293 */
  const uint8_t* skip_param;  /* [d] */
  const int32_t* block_ptr;   /* [d+1] CSR of df_params$block */
  const int32_t* block_idx;   /* 1-based block ids */
  const double* beta_raised;  /* [rungs] */
  const double* theta_init;   /* [chains][d] */
  int32_t rng_mode;           /* ORA_RNG_* */
  int32_t n_threads;          /* chains are run one per OpenMP thread (not in RMT mode) */
  uint64_t seed;              /* Philox key */
  int64_t chain_offset;       /* global id of chain 0 of this call (Philox counter) */
  const double* replay_uniforms; /* [chains][T][U], SURVEY 3.5 */
  ora_rmt* rmt;               /* RMT mode: stream continues across chains (lapply, R/main.R:401) */
} ora_config;

/* Same 9 fields as the list returned at src/main.cpp:268-276, chain-major, plus accept counts. */
typedef struct ora_output {
  double* loglike_burnin;    /* [chains][rungs][burnin] */
  double* logprior_burnin;   /* [chains][rungs][burnin] */
  double* theta_burnin;      /* [chains][theta_rungs][burnin][d] */
  double* loglike_sampling;  /* [chains][rungs][samples] */
  double* logprior_sampling; /* [chains][rungs][samples] */
  double* theta_sampling;    /* [chains][theta_rungs][samples][d] */
  int32_t* mc_accept_burnin;   /* [chains][rungs-1] */
  int32_t* mc_accept_sampling; /* [chains][rungs-1] */
  int32_t* accept_burnin;      /* [chains][rungs]  accept_count at end of burn-in (main.cpp:195) */
  int32_t* accept_sampling;    /* [chains][rungs]  accept_count at end of sampling (main.cpp:255) */
  double* t_diff;              /* [chains] seconds */
  double* bw_final;            /* [chains][rungs][d] nullable */
} ora_output;

enum {
  ORA_OK = 0,
  ORA_ERR_INIT_NEGINF = 1,   /* Particle.h:75-78 */
  ORA_ERR_LOGLIKE_BAD = 2,   /* Particle.h:121-124 */
  ORA_ERR_LOGPRIOR_BAD = 3,  /* Particle.h:125-128 */
  ORA_ERR_TRANS_TYPE = 4,    /* Particle.cpp:71,95,120 */
  ORA_ERR_MODEL = 5,
  ORA_ERR_ARG = 6
};

typedef struct ora_error {
  int32_t code;
  int32_t rung;
  int32_t param;
  int32_t iteration;
  int64_t chain;
  double theta[64];
  char message[256];
} ora_error;

/* run all chains of the config; returns ORA_* (first error in chain order) */
int ora_run_mcmc(const ora_config* cfg, const char* model_name, const ora_list* data,
                 const ora_list* misc, ora_output* out, ora_error* err);

/* evaluate a model's loglike / logprior once (used by density/model unit tests) */
double ora_model_loglike(const char* model_name, int d, const double* theta, const ora_list* data,
                         const ora_list* misc, int block);
double ora_model_logprior(const char* model_name, int d, const double* theta, const ora_list* misc);

/* R default RNG restatement */
void ora_rmt_set_seed(ora_rmt* s, uint32_t seed);
double ora_rmt_unif_rand(ora_rmt* s);
double ora_rmt_norm_rand(ora_rmt* s);
void ora_rmt_runif(ora_rmt* s, int n, double* out);
void ora_rmt_rnorm(ora_rmt* s, int n, double mean, double sd, double* out);

/* Philox4x32-10 */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the uniforms the sampler draws for (chain, iteration, rung, slot, tag) */
void ora_philox_uniforms(uint64_t seed, int64_t chain, uint32_t iteration, uint32_t rung,
                         uint32_t slot, uint32_t tag, double u[4]);

/* nmath restatements */
double ora_qnorm(double p);
double ora_dnorm(double x, double mu, double sigma, int give_log);
double ora_dlnorm(double x, double meanlog, double sdlog, int give_log);
double ora_dgamma(double x, double shape, double scale, int give_log);
double ora_dbeta(double x, double a, double b, int give_log);
double ora_dpois(double x, double lambda, int give_log);
double ora_dunif(double x, double a, double b, int give_log);
double ora_dexp(double x, double scale, int give_log);
double ora_dbinom(double x, double n, double p, int give_log);
double ora_stirlerr(double n);
double ora_bd0(double x, double np);

#ifdef __cplusplus
}
#endif
#endif
