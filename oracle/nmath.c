/*
 * nmath.c -- ORACLE (test infrastructure, see drj_oracle.h).
 *
 * Restatement of the R nmath density functions the reference's model files call through the
 * `R::` namespace (not vendored by the reference; base R, version unpinned):
 *   R::dnorm   inst/extdata/example_loglike_logprior.cpp:17, blocks_...:26,36,54, coupling_...:20,34
 *   R::dlnorm  inst/extdata/example_loglike_logprior.cpp:31
 *   R::dgamma  inst/extdata/blocks_loglike_logprior.cpp:53,55; checks/returnprior_...:13-14
 *   R::dbeta   inst/extdata/checks/returnprior_loglike_logprior.cpp:15
 *   dpois      vignettes/getting_model_fits.Rmd:93 ; dunif :101-103
 * Algorithms: dnorm/dlnorm closed forms with R's edge-case order; dpois/dgamma/dbeta/dbinom via
 * Catherine Loader's saddle-point expansion (stirlerr, bd0), "Fast and accurate computation of
 * binomial probabilities" (2000).  Checked against scipy.stats in tests/test_oracle_nmath.py.
 */
#include <float.h>
#include <math.h>

#include "drj_oracle.h"

#define LN_SQRT_2PI 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#define LN_2PI 1.837877066409345483560659472811     /* log(2*pi) */
#define TWO_PI 6.283185307179586476925286766559

static double d0(int give_log) { return give_log ? -INFINITY : 0.0; }
static double d1(int give_log) { return give_log ? 0.0 : 1.0; }
static double dexp_(double x, int give_log) { return give_log ? x : exp(x); }
static double dfexp(double f, double x, int give_log) {
  return give_log ? -0.5 * log(f) + x : exp(x) / sqrt(f);
}

double ora_dnorm(double x, double mu, double sigma, int give_log) {
  if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
  if (sigma < 0) return NAN;
  if (!isfinite(sigma)) return d0(give_log);
  if (!isfinite(x) && mu == x) return NAN;
  if (sigma == 0) return (x == mu) ? INFINITY : d0(give_log);
  x = (x - mu) / sigma;
  if (!isfinite(x)) return d0(give_log);
  x = fabs(x);
  if (x >= 2 * sqrt(DBL_MAX)) return d0(give_log);
  if (give_log) return -(LN_SQRT_2PI + 0.5 * x * x + log(sigma));
  return 0.398942280401432677939946059934 * exp(-0.5 * x * x) / sigma;
}

double ora_dlnorm(double x, double meanlog, double sdlog, int give_log) {
  if (isnan(x) || isnan(meanlog) || isnan(sdlog)) return x + meanlog + sdlog;
  if (sdlog < 0) return NAN;
  if (!isfinite(x) && log(x) == meanlog) return NAN;
  if (sdlog == 0) return (log(x) == meanlog) ? INFINITY : d0(give_log);
  if (x <= 0) return d0(give_log);
  double y = (log(x) - meanlog) / sdlog;
  return give_log ? -(LN_SQRT_2PI + 0.5 * y * y + log(x * sdlog))
                  : 0.398942280401432677939946059934 * exp(-0.5 * y * y) / (x * sdlog);
}

/* log(n!) - log( sqrt(2*pi*n)*(n/e)^n ), tabulated at half-integers up to 15 */
double ora_stirlerr(double n) {
  static const double S0 = 0.083333333333333333333;        /* 1/12 */
  static const double S1 = 0.00277777777777777777778;      /* 1/360 */
  static const double S2 = 0.00079365079365079365079365;   /* 1/1260 */
  static const double S3 = 0.000595238095238095238095238;  /* 1/1680 */
  static const double S4 = 0.0008417508417508417508417508; /* 1/1188 */
  static const double sferr_halves[31] = {
      0.0,
      0.1534264097200273452913839393,
      0.08106146679532725821967026359,
      0.05481412105191765389613870235,
      0.04134069595540929409382208141,
      0.03316287351993628748511050974,
      0.02767792568499833914878929275,
      0.02374616365629749597133027909,
      0.02079067210376509311152277177,
      0.01848845053267318523077935748,
      0.01664469118982119216319486537,
      0.01513497322191737887351383688,
      0.01387612882307074799874572702,
      0.01281046524292022692425065528,
      0.01189670994589177009505572412,
      0.0111045597582069173266307552,
      0.01041126526197209649747856713,
      0.009799416126158803298390373402,
      0.009255462182712732917728636633,
      0.008768700134139385462955047269,
      0.00833056343336287125646931866,
      0.007934114564314020547249562491,
      0.007573675487951840794972024212,
      0.007244554301320383179546196602,
      0.006942840107209529865664152663,
      0.006665247032707682442356180895,
      0.006408994188004207068439631083,
      0.006171712263039457647534604798,
      0.005951370112758847735624416046,
      0.005746216513010115682026102477,
      0.00555473355196280137103868996};
  double nn;
  if (n <= 15.0) {
    nn = n + n;
    if (nn == (int)nn) return sferr_halves[(int)nn];
    return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - LN_SQRT_2PI;
  }
  nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* deviance part: x log(x/np) + np - x, stable for x near np */
double ora_bd0(double x, double np) {
  if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    if (fabs(s) < DBL_MIN) return s;
    double ej = 2 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / ((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / np) + np - x;
}

static double dpois_raw(double x, double lambda, int give_log) {
  if (lambda == 0) return (x == 0) ? d1(give_log) : d0(give_log);
  if (!isfinite(lambda)) return d0(give_log);
  if (x < 0) return d0(give_log);
  if (x <= lambda * DBL_MIN) return dexp_(-lambda, give_log);
  if (lambda < x * DBL_MIN) {
    if (!isfinite(x)) return d0(give_log);
    return dexp_(-lambda + x * log(lambda) - lgamma(x + 1), give_log);
  }
  return dfexp(TWO_PI * x, -ora_stirlerr(x) - ora_bd0(x, lambda), give_log);
}

double ora_dpois(double x, double lambda, int give_log) {
  if (isnan(x) || isnan(lambda)) return x + lambda;
  if (lambda < 0) return NAN;
  if (fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x))) return d0(give_log); /* non-integer x */
  if (x < 0 || !isfinite(x)) return d0(give_log);
  x = nearbyint(x);
  return dpois_raw(x, lambda, give_log);
}

double ora_dgamma(double x, double shape, double scale, int give_log) {
  if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
  if (shape < 0 || scale <= 0) return NAN;
  if (x < 0) return d0(give_log);
  if (shape == 0) return (x == 0) ? INFINITY : d0(give_log);
  if (x == 0) {
    if (shape < 1) return INFINITY;
    if (shape > 1) return d0(give_log);
    return give_log ? -log(scale) : 1 / scale;
  }
  double pr;
  if (shape < 1) {
    pr = dpois_raw(shape, x / scale, give_log);
    return give_log ? pr + (isfinite(shape / x) ? log(shape / x) : log(shape) - log(x))
                    : pr * shape / x;
  }
  pr = dpois_raw(shape - 1, x / scale, give_log);
  return give_log ? pr - log(scale) : pr / scale;
}

static double dbinom_raw(double x, double n, double p, double q, int give_log) {
  if (p == 0) return (x == 0) ? d1(give_log) : d0(give_log);
  if (q == 0) return (x == n) ? d1(give_log) : d0(give_log);
  double lc;
  if (x == 0) {
    if (n == 0) return d1(give_log);
    lc = (p < 0.1) ? -ora_bd0(n, n * q) - n * p : n * log(q);
    return dexp_(lc, give_log);
  }
  if (x == n) {
    lc = (q < 0.1) ? -ora_bd0(n, n * p) - n * q : n * log(p);
    return dexp_(lc, give_log);
  }
  if (x < 0 || x > n) return d0(give_log);
  lc = ora_stirlerr(n) - ora_stirlerr(x) - ora_stirlerr(n - x) - ora_bd0(x, n * p) -
       ora_bd0(n - x, n * q);
  double lf = LN_2PI + log(x) + log1p(-x / n);
  return dexp_(lc - 0.5 * lf, give_log);
}

double ora_dbinom(double x, double n, double p, int give_log) {
  if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
  if (p < 0 || p > 1 || n < 0 || fabs(n - nearbyint(n)) > 1e-7 * fmax(1.0, fabs(n))) return NAN;
  if (fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x))) return d0(give_log);
  if (x < 0 || !isfinite(x)) return d0(give_log);
  n = nearbyint(n);
  x = nearbyint(x);
  return dbinom_raw(x, n, p, 1 - p, give_log);
}

static double lbeta_(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

double ora_dbeta(double x, double a, double b, int give_log) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (a < 0 || b < 0) return NAN;
  if (x < 0 || x > 1) return d0(give_log);
  if (a == 0 || b == 0 || !isfinite(a) || !isfinite(b)) {
    if (a == 0 && b == 0) return (x == 0 || x == 1) ? INFINITY : d0(give_log);
    if (a == 0 || a / b == 0) return (x == 0) ? INFINITY : d0(give_log);
    if (b == 0 || b / a == 0) return (x == 1) ? INFINITY : d0(give_log);
    return (x == 0.5) ? INFINITY : d0(give_log);
  }
  if (x == 0) {
    if (a > 1) return d0(give_log);
    if (a < 1) return INFINITY;
    return give_log ? log(b) : b;
  }
  if (x == 1) {
    if (b > 1) return d0(give_log);
    if (b < 1) return INFINITY;
    return give_log ? log(a) : a;
  }
  double lval;
  if (a <= 2 || b <= 2)
    lval = (a - 1) * log(x) + (b - 1) * log1p(-x) - lbeta_(a, b);
  else
    lval = log(a + b - 1) + dbinom_raw(a - 1, a + b - 2, x, 1 - x, 1);
  return dexp_(lval, give_log);
}

double ora_dunif(double x, double a, double b, int give_log) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (b <= a) return NAN;
  if (a <= x && x <= b) return give_log ? -log(b - a) : 1.0 / (b - a);
  return d0(give_log);
}

double ora_dexp(double x, double scale, int give_log) {
  if (isnan(x) || isnan(scale)) return x + scale;
  if (scale <= 0.0) return NAN;
  if (x < 0.0) return d0(give_log);
  return give_log ? (-x / scale) - log(scale) : exp(-x / scale) / scale;
}
