/*
 * models.c -- ORACLE (test infrastructure, see drj_oracle.h).
 *
 * CPU restatements of the user likelihood / prior functions that the reference ships and that
 * BASELINE.json's configs name.  Each follows the cited reference file operation for operation
 * (per-datum R::dnorm, sequential `ret +=` accumulation).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "drj_oracle.h"
#include "models.h"

static const double* item(const ora_list* l, int k) { return l->values + l->offsets[k]; }
static int64_t item_len(const ora_list* l, int k) { return l->lengths[k]; }

/* --- inst/extdata/example_loglike_logprior.cpp:5-35 (K1, K5): params (mu, sigma), data x */
static double example_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)misc; (void)block;
  double mu = p[0], sigma = p[1];
  const double* x = item(data, 0);
  int64_t n = item_len(data, 0);
  double ret = 0.0;
  for (int64_t i = 0; i < n; ++i) ret += ora_dnorm(x[i], mu, sigma, 1);
  return ret;
}
static double example_logprior(const double* p, int d, const ora_list* misc) {
  (void)d; (void)misc;
  return -log(20.0) + ora_dlnorm(p[1], 0.0, 1.0, 1);
}

/* Same model with the reference's per-call overhead that matters at large n: the whole data vector
 * is copied on EVERY likelihood call (std::vector<double> x = Rcpp::as<...>(data["x"]),
 * inst/extdata/example_loglike_logprior.cpp:12).  Used by bench.py as the "faithful overhead" baseline. */
static double example_copy_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)misc; (void)block;
  double mu = p[0], sigma = p[1];
  int64_t n = item_len(data, 0);
  double* x = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  memcpy(x, item(data, 0), sizeof(double) * (size_t)n);
  double ret = 0.0;
  for (int64_t i = 0; i < n; ++i) ret += ora_dnorm(x[i], mu, sigma, 1);
  free(x);
  return ret;
}

/* --- inst/extdata/coupling_loglike_logprior.cpp:5-37: params (alpha, beta, epsilon), data x */
static double coupling_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)misc; (void)block;
  double mean = p[0] * p[0] * p[1] + p[2];
  const double* x = item(data, 0);
  int64_t n = item_len(data, 0);
  double ret = 0.0;
  for (int64_t i = 0; i < n; ++i) ret += ora_dnorm(x[i], mean, 1.0, 1);
  return ret;
}
static double coupling_logprior(const double* p, int d, const ora_list* misc) {
  (void)d; (void)misc;
  return -log(20.0) - log(10.0) + ora_dnorm(p[2], 0.0, 1.0, 1);
}

/* --- inst/extdata/blocks_loglike_logprior.cpp:5-58: params mu_1..mu_G, sigma, phi, tau
 *     (G = d-3 groups; the shipped file has G = 6, blocks 1..G data groups, block G+1 global) */
static double blocks_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)misc;
  int G = d - 3;
  double sigma = p[G], phi = p[G + 1], tau = p[G + 2];
  double ret = 0.0;
  if (block == G + 1) {
    for (int i = 0; i < G; ++i) ret += ora_dnorm(p[i], phi, tau, 1);
  } else {
    const double* x = item(data, block - 1);
    int64_t n = item_len(data, block - 1);
    for (int64_t i = 0; i < n; ++i) ret += ora_dnorm(x[i], p[block - 1], sigma, 1);
  }
  return ret;
}
static double blocks_logprior(const double* p, int d, const ora_list* misc) {
  (void)misc;
  int G = d - 3;
  return ora_dgamma(p[G], 0.01, 100.0, 1) + ora_dnorm(p[G + 1], 0, 1000, 1) +
         ora_dgamma(p[G + 2], 0.01, 100.0, 1);
}

/* --- inst/extdata/checks/doublewell_loglike_logprior.cpp:5-16 (K2): params (mu, gamma) */
static double doublewell_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)data; (void)misc; (void)block;
  double mu = p[0], gamma = p[1];
  return -gamma * (mu * mu - 1.0) * (mu * mu - 1.0);
}
static double zero_logprior(const double* p, int d, const ora_list* misc) {
  (void)p; (void)d; (void)misc;
  return 0.0;
}

/* --- inst/extdata/checks/multilevel_loglike_logprior.cpp:5-33 (K4): params mu_1..mu_G (G = d),
 *     blocks 1..G data groups (sd 1), block G+1 global N(0,1) */
static double multilevel_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)misc;
  int G = d;
  double ret = 0.0;
  if (block == G + 1) {
    for (int i = 0; i < G; ++i) ret += ora_dnorm(p[i], 0, 1.0, 1);
  } else {
    const double* x = item(data, block - 1);
    int64_t n = item_len(data, block - 1);
    for (int64_t i = 0; i < n; ++i) ret += ora_dnorm(x[i], p[block - 1], 1.0, 1);
  }
  return ret;
}

/* --- inst/extdata/checks/normal_loglike_logprior.cpp:5-20: params (mu), data x */
static double normal_mu_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)misc; (void)block;
  const double* x = item(data, 0);
  int64_t n = item_len(data, 0);
  double ret = 0.0;
  for (int64_t i = 0; i < n; ++i) ret += ora_dnorm(x[i], p[0], 1.0, 1);
  return ret;
}

/* --- inst/extdata/checks/returnprior_loglike_logprior.cpp:5-18:
 *     params (real_line, neg_line, pos_line, unit_interval) */
static double zero_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)p; (void)d; (void)data; (void)misc; (void)block;
  return 0.0;
}
static double returnprior_logprior(const double* p, int d, const ora_list* misc) {
  (void)d; (void)misc;
  return ora_dnorm(p[0], 0.0, 1.0, 1) + ora_dgamma(-p[1], 5.0, 5.0, 1) + ora_dgamma(p[2], 5.0, 5.0, 1) +
         ora_dbeta(p[3], 3.0, 3.0, 1);
}

/* --- tests/testthat/test_input_files/loglike_logprior.cpp:6-50: (mu, sigma) with underflow clamp */
static double testfile_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  double ret = example_loglike(p, d, data, misc, block);
  if (!isfinite(ret)) ret = -(DBL_MAX / 100.0);
  return ret;
}
static double testfile_logprior(const double* p, int d, const ora_list* misc) {
  (void)d; (void)misc;
  double ret = ora_dnorm(p[0], 6, 0.1, 1) + ora_dnorm(p[1], 1, 0.1, 1);
  if (!isfinite(ret)) ret = -(DBL_MAX / 100.0);
  return ret;
}

/* --- vignettes/getting_model_fits.Rmd:72-104 (K3): params (N_0, r, N_max),
 *     data items (pop, time); 100-step discrete logistic map, Poisson observation */
static double logistic_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)misc; (void)block;
  double N_0 = p[0], r = p[1], N_max = p[2];
  double K = N_max * r / (r - 1);
  double x[100];
  x[0] = N_0;
  for (int i = 1; i < 100; ++i) x[i] = r * x[i - 1] * (1 - x[i - 1] / K);
  const double* pop = item(data, 0);
  const double* time = item(data, 1);
  int64_t n = item_len(data, 0);
  double ret = 0.0;
  for (int64_t j = 0; j < n; ++j) ret += ora_dpois(pop[j], x[(int)time[j] - 1], 1);
  return ret;
}
static double logistic_logprior(const double* p, int d, const ora_list* misc) {
  (void)d; (void)misc;
  return ora_dunif(p[0], 1, 100, 1) + ora_dunif(p[1], 1, 2, 1) + ora_dunif(p[2], 50, 200, 1);
}

/* --- tests/testthat/test-Inf_checks_work.R:14-94: functions returning a constant (misc[0]) */
static double const_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)p; (void)d; (void)data; (void)block;
  return item(misc, 0)[0];
}
static double const_logprior(const double* p, int d, const ora_list* misc) {
  (void)p; (void)d;
  return item(misc, 1)[0];
}
/* --- tests/testthat/test-Inf_checks_work.R:162-169, 221-228: -Inf over half the domain */
static double halfdomain_loglike(const double* p, int d, const ora_list* data, const ora_list* misc, int block) {
  (void)d; (void)data; (void)misc; (void)block;
  return (p[0] > 0.5) ? -INFINITY : 2.0;
}
static double halfdomain_logprior(const double* p, int d, const ora_list* misc) {
  (void)d; (void)misc;
  return (p[0] > 0.5) ? -INFINITY : 2.0;
}

static const ora_model MODELS[] = {
    {"example", example_loglike, example_logprior},
    {"example_exact", example_loglike, example_logprior},
    {"example_copy", example_copy_loglike, example_logprior},
    {"coupling", coupling_loglike, coupling_logprior},
    {"blocks", blocks_loglike, blocks_logprior},
    {"doublewell", doublewell_loglike, zero_logprior},
    {"multilevel", multilevel_loglike, zero_logprior},
    {"normal_mu", normal_mu_loglike, zero_logprior},
    {"returnprior", zero_loglike, returnprior_logprior},
    {"testfile", testfile_loglike, testfile_logprior},
    {"testfile_nullprior", testfile_loglike, zero_logprior},
    {"testfile_nulllike", zero_loglike, testfile_logprior},
    {"logistic", logistic_loglike, logistic_logprior},
    {"const", const_loglike, const_logprior},
    {"halfdomain_like", halfdomain_loglike, zero_logprior},
    {"halfdomain_prior", zero_loglike, halfdomain_logprior},
};

const ora_model* ora_find_model(const char* name) {
  for (unsigned i = 0; i < sizeof(MODELS) / sizeof(MODELS[0]); ++i)
    if (strcmp(MODELS[i].name, name) == 0) return &MODELS[i];
  return 0;
}

double ora_model_loglike(const char* model_name, int d, const double* theta, const ora_list* data,
                         const ora_list* misc, int block) {
  const ora_model* m = ora_find_model(model_name);
  return m ? m->loglike(theta, d, data, misc, block) : NAN;
}
double ora_model_logprior(const char* model_name, int d, const double* theta, const ora_list* misc) {
  const ora_model* m = ora_find_model(model_name);
  return m ? m->logprior(theta, d, misc) : NAN;
}
