// drj_models.cuh -- built-in model library: CUDA restatements of every likelihood / prior the
// reference ships (inst/extdata/*.cpp, inst/extdata/checks/*.cpp, tests/testthat/test_input_files,
// vignettes/getting_model_fits.Rmd) written against the device model contract of drj_sampler.cuh.
// Parameters arrive as `const double* theta` in df_params order (the reference passes a named
// NumericVector, src/main.cpp:11); `data` / `misc` are the flattened named lists; the current block
// (misc["block"], src/Particle.h:110) is the `block` argument.
#pragma once
#include "drj_sampler.cuh"
#include "drj_lanes.cuh"
#include "drj_spec.cuh"

// per-datum form of  sum_i R::dnorm(x_i, mean, sd, log=TRUE) = -(n (log sqrt(2 pi) + log sd) + sum_i (x_i - mean)^2 / (2 sd^2)):
// per datum  d = x - mean (1 DADD), acc = d*d + acc (1 DFMA); the scale 1/(2 sd^2) and the constants are
// applied once in finish().  Two FP64-pipe instructions per datum, each with at most two distinct
// source registers: the earlier form z = fma(x, 1/sd, -mean/sd) read three, and when ptxas put both
// loop constants in the same register bank that copy of the loop ran ~2x slower (profiles/README.md).
struct DrjNormConsts {
  double mean, half_inv_var, lognorm;
};
__device__ __forceinline__ bool drj_norm_prepare_l(double mean, double sd, double logsd, DrjNormConsts& k) {
  k.mean = mean;
  k.half_inv_var = 0.5 / (sd * sd);
  k.lognorm = DRJ_LN_SQRT_2PI + logsd;
  // outside this range R::dnorm's edge-case branches matter: use the exact serial evaluation
  return (sd > 1e-150) && (sd < 1e150) && (fabs(mean) < 1e150);
}
__device__ __forceinline__ bool drj_norm_prepare(double mean, double sd, DrjNormConsts& k) {
  // (log(1) is exactly +0: where sd is the literal 1 the compiler drops the logarithm, which it cannot fold itself)
  return drj_norm_prepare_l(mean, sd, sd == 1.0 ? 0.0 : log(sd), k);
}
__device__ __forceinline__ double drj_norm_datum(const DrjNormConsts& k, double x, double acc) {
  const double d = x - k.mean;
  return fma(d, d, acc);
}
__device__ __forceinline__ double drj_norm_finish(const DrjNormConsts& k, double acc, long long n) {
  return -((double)n * k.lognorm + k.half_inv_var * acc);
}
__device__ __forceinline__ double drj_norm_serial(const double* x, long long n, double mean, double sd) {
  double ret = 0.0;
  for (long long i = 0; i < n; ++i) ret += R::dnorm(x[i], mean, sd, 1);
  return ret;
}

// inst/extdata/example_loglike_logprior.cpp:5-35 -- params (mu, sigma), data$x  (K1, K5)
struct DrjModelExample {
  static constexpr int D = 2, NBLOCK = 1;
  static constexpr bool DATUM_FORM = true;
  typedef DrjNormConsts Consts;
  __device__ static __forceinline__ int datum_item(int) { return 0; }
  __device__ static __forceinline__ bool prepare(const double* th, const DrjList&, int, Consts& k) {
    return drj_norm_prepare(th[0], th[1], k);
  }
  __device__ static __forceinline__ double datum(const Consts& k, double x, double acc) { return drj_norm_datum(k, x, acc); }
  __device__ static __forceinline__ double finish(const Consts& k, double acc, long long n) { return drj_norm_finish(k, acc, n); }
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int) {
    return drj_norm_serial(data.item(0), data.len(0), th[0], th[1]);
  }
  __device__ static __forceinline__ double logprior(const double* th, const DrjList&) {
    return -2.99573227355399099343 + R::dlnorm(th[1], 0.0, 1.0, 1);
  }
  // sigma's lower bound is 0 in the reference's parameter table: log(sigma) is the sampler's Jacobian logarithm (DrjHints)
  static constexpr bool HINTS = true;
  __device__ static __forceinline__ bool prepare_h(const double* th, const DrjList&, int, Consts& k, const double* loglo, unsigned mask) {
    return drj_norm_prepare_l(th[0], th[1], (mask & 2u) ? loglo[1] : log(th[1]), k);
  }
  __device__ static __forceinline__ double logprior_h(const double* th, const DrjList&, const double* loglo, unsigned mask) {
    return -2.99573227355399099343 + ((mask & 2u) ? R::dlnorm_l(th[1], loglo[1], 0.0, 1.0, 1) : R::dlnorm(th[1], 0.0, 1.0, 1));
  }
};

// same model, likelihood evaluated exactly as the reference does (serial R::dnorm per datum)
struct DrjModelExampleExact : DrjModelExample {
  static constexpr bool DATUM_FORM = false;
};

// inst/extdata/coupling_loglike_logprior.cpp:5-37 -- params (alpha, beta, epsilon), data$x
struct DrjModelCoupling {
  static constexpr int D = 3, NBLOCK = 1;
  static constexpr bool DATUM_FORM = true;
  typedef DrjNormConsts Consts;
  __device__ static __forceinline__ int datum_item(int) { return 0; }
  __device__ static __forceinline__ bool prepare(const double* th, const DrjList&, int, Consts& k) {
    return drj_norm_prepare(th[0] * th[0] * th[1] + th[2], 1.0, k);
  }
  __device__ static __forceinline__ double datum(const Consts& k, double x, double acc) { return drj_norm_datum(k, x, acc); }
  __device__ static __forceinline__ double finish(const Consts& k, double acc, long long n) { return drj_norm_finish(k, acc, n); }
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int) {
    return drj_norm_serial(data.item(0), data.len(0), th[0] * th[0] * th[1] + th[2], 1.0);
  }
  __device__ static __forceinline__ double logprior(const double* th, const DrjList&) {
    return -2.99573227355399099343 - 2.30258509299404568402 + R::dnorm(th[2], 0.0, 1.0, 1);
  }
};

// inst/extdata/checks/normal_loglike_logprior.cpp:5-20 -- params (mu), data$x, sd 1
struct DrjModelNormalMu {
  static constexpr int D = 1, NBLOCK = 1;
  static constexpr bool DATUM_FORM = true;
  typedef DrjNormConsts Consts;
  __device__ static __forceinline__ int datum_item(int) { return 0; }
  __device__ static __forceinline__ bool prepare(const double* th, const DrjList&, int, Consts& k) {
    return drj_norm_prepare(th[0], 1.0, k);
  }
  __device__ static __forceinline__ double datum(const Consts& k, double x, double acc) { return drj_norm_datum(k, x, acc); }
  __device__ static __forceinline__ double finish(const Consts& k, double acc, long long n) { return drj_norm_finish(k, acc, n); }
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int) {
    return drj_norm_serial(data.item(0), data.len(0), th[0], 1.0);
  }
  __device__ static __forceinline__ double logprior(const double*, const DrjList&) { return 0.0; }
};

// inst/extdata/checks/doublewell_loglike_logprior.cpp:5-16 -- params (mu, gamma), no data  (K2)
struct DrjModelDoubleWell {
  static constexpr int D = 2, NBLOCK = 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static __forceinline__ double loglike(const double* th, const DrjList&, const DrjList&, int) {
    const double mu = th[0], gamma = th[1];
    return -gamma * (mu * mu - 1.0) * (mu * mu - 1.0);
  }
  __device__ static __forceinline__ double logprior(const double*, const DrjList&) { return 0.0; }
};

// inst/extdata/checks/multilevel_loglike_logprior.cpp:5-33 generalised to G groups (the shipped
// file has G = 5; K4 uses G = 50): params mu_1..mu_G, blocks 1..G = data groups, block G+1 = global
template <int G>
struct DrjModelMultilevel {
  static constexpr int D = G, NBLOCK = G + 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int block) {
    double ret = 0.0;
    if (block == G + 1) {
#pragma unroll 10
      for (int i = 0; i < G; ++i) ret += R::dnorm(th[i], 0.0, 1.0, 1);
    } else {
      const double* x = data.item(block - 1);
      const long long n = data.len(block - 1);
      const double mu = th[block - 1];
      for (long long i = 0; i < n; ++i) ret += R::dnorm(x[i], mu, 1.0, 1);
    }
    return ret;
  }
  __device__ static __forceinline__ double logprior(const double*, const DrjList&) { return 0.0; }
  // lane-split form (drj_lanes.cuh): lane l of LANES adds the terms l, l + LANES, ... of the block, in order
#ifndef DRJ_MULTILEVEL_LANES
#define DRJ_MULTILEVEL_LANES 8  // K4 (4096 x 16, G = 50) measured: 4 lanes 4.6e9 (8 warps per SM: 3 KB of state per particle), 8 lanes 5.2e9, 16 lanes 3.1e9
#endif
  static constexpr int LANES = DRJ_MULTILEVEL_LANES;
  __device__ static double loglike_lanes(const double* th, const DrjList& data, const DrjList&, int block, int lane) {
    double ret = 0.0;
    if (block == G + 1) {
      // regular arguments: dnorm(z, 0, 1, log) = -(log sqrt(2 pi) + z^2 / 2), the same operations as R::dnorm's regular
      // path, with ONE test for the whole lane instead of one per term; anything else takes R::dnorm term by term
      constexpr int NT = (G + LANES - 1) / LANES;
      bool regular = true;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const int i = lane + k * LANES;
        if (i < G) {
          const double z = th[i];
          regular = regular && (fabs(z) < 2.6815615859885194e+154);
          ret += -(DRJ_LN_SQRT_2PI + 0.5 * z * z);
        }
      }
      if (!regular) {
        ret = 0.0;
        for (int i = lane; i < G; i += LANES) ret += R::dnorm(th[i], 0.0, 1.0, 1);
      }
    } else {
      const double* x = data.item(block - 1);
      const long long n = data.len(block - 1);
      const double mu = th[block - 1];
      for (long long i = lane; i < n; i += LANES) ret += R::dnorm(x[i], mu, 1.0, 1);
    }
    return ret;
  }
};

// inst/extdata/blocks_loglike_logprior.cpp:5-58 with G groups (shipped: G = 6):
// params mu_1..mu_G, sigma, phi, tau
template <int G>
struct DrjModelBlocks {
  static constexpr int D = G + 3, NBLOCK = G + 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int block) {
    const double sigma = th[G], phi = th[G + 1], tau = th[G + 2];
    double ret = 0.0;
    if (block == G + 1) {
      for (int i = 0; i < G; ++i) ret += R::dnorm(th[i], phi, tau, 1);
    } else {
      const double* x = data.item(block - 1);
      const long long n = data.len(block - 1);
      const double mu = th[block - 1];
      for (long long i = 0; i < n; ++i) ret += R::dnorm(x[i], mu, sigma, 1);
    }
    return ret;
  }
  __device__ static double logprior(const double* th, const DrjList&) {
    return R::dgamma(th[G], 0.01, 100.0, 1) + R::dnorm(th[G + 1], 0.0, 1000.0, 1) + R::dgamma(th[G + 2], 0.01, 100.0, 1);
  }
};

// inst/extdata/checks/returnprior_loglike_logprior.cpp:5-18:
// params (real_line, neg_line, pos_line, unit_interval)
struct DrjModelReturnPrior {
  static constexpr int D = 4, NBLOCK = 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static __forceinline__ double loglike(const double*, const DrjList&, const DrjList&, int) { return 0.0; }
  __device__ static double logprior(const double* th, const DrjList&) {
    return R::dnorm(th[0], 0.0, 1.0, 1) + R::dgamma(-th[1], 5.0, 5.0, 1) + R::dgamma(th[2], 5.0, 5.0, 1) +
           R::dbeta(th[3], 3.0, 3.0, 1);
  }
};

// tests/testthat/test_input_files/loglike_logprior.cpp:6-50 -- (mu, sigma) with the underflow clamp
template <bool LIKE, bool PRIOR>
struct DrjModelTestFile {
  static constexpr int D = 2, NBLOCK = 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int) {
    if (!LIKE) return 0.0;
    double ret = drj_norm_serial(data.item(0), data.len(0), th[0], th[1]);
    if (!isfinite(ret)) ret = -(DRJ_DBL_MAX / 100.0);
    return ret;
  }
  __device__ static double logprior(const double* th, const DrjList&) {
    if (!PRIOR) return 0.0;
    double ret = R::dnorm(th[0], 6.0, 0.1, 1) + R::dnorm(th[1], 1.0, 0.1, 1);
    if (!isfinite(ret)) ret = -(DRJ_DBL_MAX / 100.0);
    return ret;
  }
};

// one step of the discrete logistic map x <- r x (1 - x / K) (vignettes/getting_model_fits.Rmd:80-83), the body of a
// 100-step dependent chain: written as r x - (r / K) x^2 -- two independent products and one fma, a dependent depth of
// two FP64 instructions instead of the three of the literal form (and no division on the chain: c = r / K once).  A few
// ulp per step on a map that contracts onto its fixed point, far inside the 1e-10 parity gate (tests: logistic_k3, the
// K3 full-size case, the getting_model_fits vignette's acceptance rates).
__device__ __forceinline__ double drj_logistic_step(double x, double r, double c) {
  return fma(-(c * x), x, r * x);
}

// vignettes/getting_model_fits.Rmd:72-104 -- params (N_0, r, N_max), data (pop, time)  (K3)
struct DrjModelLogistic {
  static constexpr int D = 3, NBLOCK = 1;
  static constexpr bool DATUM_FORM = false;
  // The Poisson terms' data-only parts, once per thread block (DrjCtaInit): of R::dpois(pop_j, x, log) the Stirling error of
  // pop_j (three to five FP64 divisions) and log(2 pi pop_j) do not depend on the parameters.  [0, 32): stirlerr,
  // [32, 64): the logarithm, [64, 96): the rounded count, or -1 where the count is not a regular one (or there are
  // more than 32 observations) and R::dpois itself is called.
  static constexpr bool CTA_INIT = true;
  static constexpr int MAXOBS = 32;
  __device__ static __forceinline__ double* pois_consts() {
    __shared__ double s_pc[3 * MAXOBS];
    return s_pc;
  }
  __device__ static __forceinline__ void cta_init(const DrjList& data, const DrjList&) {
    double* const pc = pois_consts();
    const double* pop = data.item(0);
    const int n = (int)data.len(0);
    for (int j = threadIdx.x; j < MAXOBS; j += blockDim.x) {
      double st = 0.0, lf = 0.0, xr = -1.0;
      if (j < n && n <= MAXOBS) {
        const double x = pop[j];
        if (!isnan(x) && !drj_nmath::nonint(x) && x >= 0.0 && isfinite(x)) {
          xr = nearbyint(x);
          st = drj_nmath::stirlerr(xr);
          lf = log(DRJ_2PI * xr);
        }
      }
      pc[j] = st; pc[MAXOBS + j] = lf; pc[2 * MAXOBS + j] = xr;
    }
  }
  __device__ static __forceinline__ double pois_term(const double* pop, int j, double lambda) {
    const double* const pc = pois_consts();
    const double xr = j < MAXOBS ? pc[2 * MAXOBS + j] : -1.0;
    if (xr >= 0.0) return R::dpois_pre(xr, lambda, pc[j], pc[MAXOBS + j]);
    return R::dpois(pop[j], lambda, 1);
  }
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int) {
    const double N_0 = th[0], r = th[1], N_max = th[2];
    const double K = N_max * r / (r - 1.0);
    const double c = r / K;
    const double* pop = data.item(0);
    const double* time = data.item(1);
    const int n = (int)data.len(0);
    // observation times are increasing in the shipped data; walk the 100-step map once and pick
    // off the observed steps (falls back to re-walking if a time goes backwards)
    double ret = 0.0;
    double x = N_0;
    int step = 1;  // x == x[step], 1-based as in the R code
    for (int j = 0; j < n; ++j) {
      const int tj = (int)time[j];
      if (tj < step) { x = N_0; step = 1; }
      for (; step < tj; ++step) x = drj_logistic_step(x, r, c);
      ret += pois_term(pop, j, x);
    }
    return ret;
  }
  __device__ static double logprior(const double* th, const DrjList&) {
    // R::dunif(x, a, b, log) with the three constant widths' logarithms written out (CUDA's log() is inline code the
    // compiler does not fold: three logarithms of constants per update were 4 % of the K3 run, round-2 ncu)
    const bool in = th[0] >= 1.0 && th[0] <= 100.0 && th[1] >= 1.0 && th[1] <= 2.0 && th[2] >= 50.0 && th[2] <= 200.0;
    if (in) return -4.59511985013458992685 + -0.0 + -5.01063529409625575354;  // -(log 99 + log 1 + log 150)
    return R::dunif(th[0], 1.0, 100.0, 1) + R::dunif(th[1], 1.0, 2.0, 1) + R::dunif(th[2], 50.0, 200.0, 1);
  }
  // lane-split form (drj_lanes.cuh): the 100-step map is a serial recurrence every lane walks (same instructions),
  // keeping the state at ITS observation times j = lane, lane + LANES, ...; the Poisson terms -- 20 x ~150
  // instructions, 3/4 of the serial evaluation -- are then evaluated LANES at a time.
#ifndef DRJ_LOGISTIC_LANES
#define DRJ_LOGISTIC_LANES 4  // K3 (1024 x 10) measured: 2 lanes 2.8e8, 4 lanes 5.2e8, 8 lanes 3.8e8, 16 lanes 3.0e8 updates/s; one thread 2.4e8
#endif
  static constexpr int LANES = DRJ_LOGISTIC_LANES;
  __device__ static double loglike_lanes(const double* th, const DrjList& data, const DrjList&, int, int lane) {
    const double N_0 = th[0], r = th[1], N_max = th[2];
    const double K = N_max * r / (r - 1.0);
    const double c = r / K;
    const double* pop = data.item(0);
    const double* time = data.item(1);
    const int n = (int)data.len(0);
    constexpr int MAXQ = 32 / LANES;  // observations per lane kept in registers (n <= 32); more: the serial walk per term below
    double xs[MAXQ];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) xs[q] = 0.0;
    double x = N_0;
    int step = 1;
    bool monotone = n <= MAXQ * LANES;
    for (int j = 0; j < n && monotone; ++j) {  // uniform over the lanes
      const int tj = (int)time[j];
      if (tj < step) { monotone = false; break; }
#pragma unroll 4
      for (; step < tj; ++step) x = drj_logistic_step(x, r, c);
      // (one store with a run-time index by the lane that owns observation j: the select over all MAXQ slots that kept xs
      // in registers was 13 % of the K3 run's instructions, round-2 ncu; the array now lives in local memory, L1-resident)
      if ((j & (LANES - 1)) == lane) xs[j / LANES] = x;
    }
    double ret = 0.0;
    if (monotone) {
#pragma unroll
      for (int q = 0; q < MAXQ; ++q) {
        const int j = lane + q * LANES;
        if (j < n) ret += pois_term(pop, j, xs[q]);
      }
    } else {  // times that go backwards, or a long series: every lane walks to each of its own observations
      for (int j = lane; j < n; j += LANES) {
        const int tj = (int)time[j];
        x = N_0;
        for (step = 1; step < tj; ++step) x = drj_logistic_step(x, r, c);
        ret += pois_term(pop, j, x);
      }
    }
    return ret;
  }
};

// tests/testthat/test-Inf_checks_work.R:14-94 -- constant returns: loglike = misc[[1]], logprior = misc[[2]]
struct DrjModelConst {
  static constexpr int D = 1, NBLOCK = 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static __forceinline__ double loglike(const double*, const DrjList&, const DrjList& misc, int) { return misc.scalar(0); }
  __device__ static __forceinline__ double logprior(const double*, const DrjList& misc) { return misc.scalar(1); }
};

// tests/testthat/test-Inf_checks_work.R:162-169 and :221-228 -- -Inf over half the domain
template <bool IN_LIKE>
struct DrjModelHalfDomain {
  static constexpr int D = 1, NBLOCK = 1;
  static constexpr bool DATUM_FORM = false;
  __device__ static __forceinline__ double loglike(const double* th, const DrjList&, const DrjList&, int) {
    if (!IN_LIKE) return 0.0;
    return (th[0] > 0.5) ? -DRJ_INF : 2.0;
  }
  __device__ static __forceinline__ double logprior(const double* th, const DrjList&) {
    if (IN_LIKE) return 0.0;
    return (th[0] > 0.5) ? -DRJ_INF : 2.0;
  }
};
