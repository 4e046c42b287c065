// drj_device.cuh -- device-side building blocks of the drjacoby_b200 sampler (sm_100a).
//
// Compiled two ways from the same text: by nvcc into libdrjacoby_b200.so (built-in models) and by
// NVRTC at run time together with a user's model source (the GPU analogue of the reference's
// inst/templates/cpp_template.cpp).  Hence: no standard headers, no host code in this file.
//
// Contents
//   DrjList            flattened named list of numeric vectors = the reference's `data` / `misc`
//                      Rcpp::List arguments (src/main.cpp:11-12)
//   drj_philox4x32_10  counter-based RNG, one call per parameter update / coupling pair
//   drj_qnorm          Wichura AS241 (what R::rnorm's inversion uses, src/Particle.cpp:49)
//   namespace R        dnorm, dlnorm, dgamma, dbeta, dpois, dunif, dexp, dbinom with R's argument
//                      order and edge-case behaviour, so bodies written against cpp_template port
//                      by changing the signature only (model files call R::dnorm etc.,
//                      inst/extdata/example_loglike_logprior.cpp:17,31)
#pragma once

#define DRJ_INF (__longlong_as_double(0x7ff0000000000000LL))
#define DRJ_NAN (__longlong_as_double(0x7ff8000000000000LL))
#define DRJ_DBL_MIN 2.2250738585072014e-308
#define DRJ_DBL_MAX 1.7976931348623157e+308
#define DRJ_LN_SQRT_2PI 0.918938533204672741780329736406
#define DRJ_LN_2PI 1.837877066409345483560659472811
#define DRJ_2PI 6.283185307179586476925286766559

// ----------------------------------------------------------------------------- lists
struct DrjList {
  const double* values;
  const long long* offsets;
  const long long* lengths;
  int n_items;
  __device__ __forceinline__ const double* item(int k) const { return values + offsets[k]; }
  __device__ __forceinline__ long long len(int k) const { return lengths[k]; }
  __device__ __forceinline__ double scalar(int k) const { return values[offsets[k]]; }
};

// ----------------------------------------------------------------------------- Philox4x32-10
// Stream layout (DESIGN.md "RNG"; mirrored by oracle/rng.c for parity):
//   key     = (seed_lo, seed_hi)
//   counter = (chain_lo32, iteration, rung << 16 | slot, tag | chain_hi << 8)
//   tag 0 = parameter update (slot = parameter index), tag 1 = coupling (slot = pair index)
struct DrjU4 { unsigned int x, y, z, w; };

__device__ __forceinline__ DrjU4 drj_philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2,
                                                   unsigned int c3, unsigned int k0, unsigned int k1) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    unsigned int hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    unsigned int hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    unsigned int n0 = hi1 ^ c1 ^ k0;
    unsigned int n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  DrjU4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
  return r;
}

// 32-bit word -> uniform in (0,1), exactly as R's unif_rand()/fixup do
__device__ __forceinline__ double drj_u32_to_unif(unsigned int w) {
  double x = (double)w * 2.3283064365386963e-10;
  if (x <= 0.0) return 0.5 * 2.328306437080797e-10;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * 2.328306437080797e-10;
  return x;
}

// ----------------------------------------------------------------------------- qnorm (AS241)
// (A branch-free variant -- both rationals evaluated for every draw, result selected -- was measured in round 2: no gain
// on the throughput shapes (a warp of 32 draws nearly always needs the tail branch too: n = 100 sweep 1.05e10 either
// way), and 8 % slower on a single chain, whose dependent chain then always contains the tail's log and sqrt.)
__device__ __forceinline__ double drj_qnorm(double p) {
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q *
          (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
               45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
            133.14166789178437745) * r + 3.387132872796366608) /
          (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
               21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
            42.313330701600911252) * r + 1.0);
    return val;
  }
  r = (q < 0.0) ? p : (1.0 - p);
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r +
                0.24178072517745061177) * r + 1.27045825245236838258) * r + 3.64784832476320460504) * r +
             5.7694972214606914055) * r + 4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r +
                0.0151986665636164571966) * r + 0.14810397642748007459) * r + 0.68976733498510000455) * r +
             1.6763848301838038494) * r + 2.05319162663775882187) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r +
                0.0012426609473880784386) * r + 0.026532189526576123093) * r + 0.29656057182850489123) * r +
             1.7848265399172913358) * r + 5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r +
                1.8463183175100546818e-5) * r + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r +
             0.13692988092273580531) * r + 0.59983220655588793769) * r + 1.0);
  }
  return (q < 0.0) ? -val : val;
}


// R::rnorm(mu, sigma) by inversion from two 32-bit-resolution uniforms (norm_rand, INVERSION), in two steps: the
// standard normal draw (independent of the chain's state: the LANES kernels compute it ahead, one draw per lane) and
// the location-scale step with R's argument checks
__device__ __forceinline__ double drj_norm_rand(double u1, double u2) {
  const double BIG = 134217728.0;  // 2^27
  const double u = (double)(int)(BIG * u1) + u2;
  return drj_qnorm(u / BIG);
}
__device__ __forceinline__ double drj_rnorm_from(double mu, double sigma, double z) {
  if (isnan(mu) || !isfinite(sigma) || sigma < 0.0) return DRJ_NAN;
  if (sigma == 0.0 || !isfinite(mu)) return mu;
  return __dadd_rn(mu, __dmul_rn(sigma, z));  // mu + sigma * norm_rand(), product rounded as in R
}
__device__ __forceinline__ double drj_rnorm(double mu, double sigma, double u1, double u2) {
  if (isnan(mu) || !isfinite(sigma) || sigma < 0.0) return DRJ_NAN;
  if (sigma == 0.0 || !isfinite(mu)) return mu;
  return __dadd_rn(mu, __dmul_rn(sigma, drj_norm_rand(u1, u2)));
}

// ----------------------------------------------------------------------------- nmath
namespace drj_nmath {
__device__ __forceinline__ double d0(int lg) { return lg ? -DRJ_INF : 0.0; }
__device__ __forceinline__ double d1(int lg) { return lg ? 0.0 : 1.0; }
__device__ __forceinline__ double dexp_(double x, int lg) { return lg ? x : exp(x); }
__device__ __forceinline__ double dfexp(double f, double x, int lg) {
  return lg ? -0.5 * log(f) + x : exp(x) / sqrt(f);
}
__device__ __forceinline__ bool nonint(double x) {
  return fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x));
}

// log(n!) - log(sqrt(2 pi n) (n/e)^n) at n = 0, 0.5, ..., 15
__device__ const double DRJ_SFERR_HALVES[31] = {
      0.0,
      0.1534264097200273452913839393, 0.08106146679532725821967026359,
      0.05481412105191765389613870235, 0.04134069595540929409382208141,
      0.03316287351993628748511050974, 0.02767792568499833914878929275,
      0.02374616365629749597133027909, 0.02079067210376509311152277177,
      0.01848845053267318523077935748, 0.01664469118982119216319486537,
      0.01513497322191737887351383688, 0.01387612882307074799874572702,
      0.01281046524292022692425065528, 0.01189670994589177009505572412,
      0.0111045597582069173266307552, 0.01041126526197209649747856713,
      0.009799416126158803298390373402, 0.009255462182712732917728636633,
      0.008768700134139385462955047269, 0.00833056343336287125646931866,
      0.007934114564314020547249562491, 0.007573675487951840794972024212,
      0.007244554301320383179546196602, 0.006942840107209529865664152663,
      0.006665247032707682442356180895, 0.006408994188004207068439631083,
      0.006171712263039457647534604798, 0.005951370112758847735624416046,
      0.005746216513010115682026102477, 0.00555473355196280137103868996};
// stirlerr: half-integer table below 15, series above (Loader 2000)
__device__ inline double stirlerr(double n) {
  const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778,
               S2 = 0.00079365079365079365079365, S3 = 0.000595238095238095238095238,
               S4 = 0.0008417508417508417508417508;
  double nn;
  if (n <= 15.0) {
    nn = n + n;
    if (nn == (double)(int)nn) return DRJ_SFERR_HALVES[(int)nn];
    return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - DRJ_LN_SQRT_2PI;
  }
  nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

// a / (2 j + 1), correctly rounded, without the division: for the divisors the series of bd0 uses, the correctly rounded
// reciprocal y is a constant, and with it q = RN(a y), r = a - b q (exact, one FMA), q' = RN(q + r y) IS RN(a / b)
// (Markstein's quotient step; checked against exact rational arithmetic, tools/markstein_check.py, and on the device
// against the division itself, tests/test_gpu_nmath.py).  CUDA's own division does the same after refining its
// reciprocal at run time: ~15 instructions and a branch against 3.  Outside a wide regular range of a (no underflow of the
// residual, no overflow) and for j >= 32 the division itself.
__device__ const double DRJ_RCP_ODD[32] = {
    0x1.0000000000000p+0, 0x1.5555555555555p-2, 0x1.999999999999ap-3, 0x1.2492492492492p-3,
    0x1.c71c71c71c71cp-4, 0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, 0x1.1111111111111p-4,
    0x1.e1e1e1e1e1e1ep-5, 0x1.af286bca1af28p-5, 0x1.8618618618618p-5, 0x1.642c8590b2164p-5,
    0x1.47ae147ae147bp-5, 0x1.2f684bda12f68p-5, 0x1.1a7b9611a7b96p-5, 0x1.0842108421084p-5,
    0x1.f07c1f07c1f08p-6, 0x1.d41d41d41d41dp-6, 0x1.bacf914c1bad0p-6, 0x1.a41a41a41a41ap-6,
    0x1.8f9c18f9c18fap-6, 0x1.7d05f417d05f4p-6, 0x1.6c16c16c16c17p-6, 0x1.5c9882b931057p-6,
    0x1.4e5e0a72f0539p-6, 0x1.4141414141414p-6, 0x1.3521cfb2b78c1p-6, 0x1.29e4129e4129ep-6,
    0x1.1f7047dc11f70p-6, 0x1.15b1e5f75270dp-6, 0x1.0c9714fbcda3bp-6, 0x1.0410410410410p-6};
__device__ __forceinline__ double drj_quot_odd(double a, double b, int j) {  // b = 2 j + 1, j < 32, a in the regular range
  const double y = DRJ_RCP_ODD[j];
  const double q = a * y;
  const double r = fma(-b, q, a);
  return fma(r, y, q);
}
__device__ __forceinline__ double drj_div_odd(double a, int j) {
  const double b = (double)((j << 1) + 1);
  if (j < 32 && fabs(a) > 1e-280 && fabs(a) < 1e280) return drj_quot_odd(a, b, j);
  return a / b;
}

// x log(x/np) + np - x, stable near x == np
__device__ inline double bd0(double x, double np) {
  if (!isfinite(x) || !isfinite(np) || np == 0.0) return DRJ_NAN;
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    if (fabs(s) < DRJ_DBL_MIN) return s;
    double ej = 2 * x * v;
    v = v * v;
    // (the terms only shrink, and the loop ends when a term falls below half an ulp of s: with s and the first term in
    // a regular range every dividend of the first 31 terms is, and the test is made once, not per term)
    const bool regular = fabs(s) > 1e-250 && fabs(ej) < 1e250;
    double bj = 1.0;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      bj += 2.0;  // 2 j + 1
      double s1 = s + ((regular && j < 32) ? drj_quot_odd(ej, bj, j) : ej / bj);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / np) + np - x;
}

__device__ inline double dpois_raw(double x, double lambda, int lg) {
  if (lambda == 0) return (x == 0) ? d1(lg) : d0(lg);
  if (!isfinite(lambda)) return d0(lg);
  if (x < 0) return d0(lg);
  if (x <= lambda * DRJ_DBL_MIN) return dexp_(-lambda, lg);
  if (lambda < x * DRJ_DBL_MIN) {
    if (!isfinite(x)) return d0(lg);
    return dexp_(-lambda + x * log(lambda) - lgamma(x + 1), lg);
  }
  return dfexp(DRJ_2PI * x, -stirlerr(x) - bd0(x, lambda), lg);
}

__device__ inline double dbinom_raw(double x, double n, double p, double q, int lg) {
  if (p == 0) return (x == 0) ? d1(lg) : d0(lg);
  if (q == 0) return (x == n) ? d1(lg) : d0(lg);
  double lc;
  if (x == 0) {
    if (n == 0) return d1(lg);
    lc = (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
    return dexp_(lc, lg);
  }
  if (x == n) {
    lc = (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
    return dexp_(lc, lg);
  }
  if (x < 0 || x > n) return d0(lg);
  lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
  double lf = DRJ_LN_2PI + log(x) + log1p(-x / n);
  return dexp_(lc - 0.5 * lf, lg);
}
}  // namespace drj_nmath

// ----------------------------------------------------------------------------- loop-invariant logarithm and reciprocal
// A likelihood written the reference's way -- for (i) ll += R::dnorm(x[i], mu, sigma, true) -- asks for log(sigma) and a
// division by sigma once per TERM.  Both are inline code with a rare-case branch inside, which the compiler can
// neither hoist out of the user's loop nor overlap between terms: measured on the README example's model, 100 data
// points, 24.1 us per update, of which 14 us the repeated logarithm and 6.5 us the division.  As calls of functions
// that are not inlined and have no side effects they ARE loop-invariant code to the compiler (hoisted out of the
// user's loop; removed where the result is unused, e.g. a literal sigma of 1).  R::dnorm then forms
// z = (x - mu) * (1 / sigma): the product differs from R's quotient by at most an ulp of z (1e-16 relative, six orders
// inside the parity gate; -DDRJ_DNORM_EXACT_DIV keeps the quotient).
__device__ __noinline__ double drj_log_invariant(double x) { return log(x); }
__device__ __noinline__ double drj_rcp_invariant(double x) { return 1.0 / x; }

namespace R {
__device__ __forceinline__ double dnorm(double x, double mu, double sigma, int lg) {
  {
    const double logsig = drj_log_invariant(sigma);
    const double inv = drj_rcp_invariant(sigma);
    // (everything formed before the test, so that neither call sinks into a conditional block and stays in the loop)
#ifdef DRJ_DNORM_EXACT_DIV
    const double zm = (x - mu) / sigma;
#else
    const double zm = (x - mu) * inv;
#endif
    const double val = -(DRJ_LN_SQRT_2PI + 0.5 * zm * zm + logsig);
    // With sigma a normal number whose reciprocal is normal too, this expression IS R's answer for every x and mu: |z|
    // beyond 2 sqrt(DBL_MAX) or infinite gives z^2 = Inf and the value -Inf (R_D__0); a NaN anywhere, or Inf - Inf,
    // gives NaN.  So the test depends on sigma alone: loop-invariant in the user's loop, no data-dependent branch per term.
    if (lg && sigma != 1.0 && sigma >= 2.2250738585072014e-308 && sigma <= 1e300) return val;
  }
  // Regular arguments first, behind ONE fused test (a NaN anywhere fails a comparison); R's edge cases after it.  The
  // regular path is the last line of R's dnorm4 with the same operations: the d = 50 multilevel model evaluates 55 of
  // these per update, and the chain of seven separate tests and branches was 2/3 of each term's instructions (round-2 ncu).
  const double z = (x - mu) / sigma;
  if (sigma > 0.0 && sigma < DRJ_INF && fabs(z) < 2.6815615859885194e+154) {  // 2*sqrt(DBL_MAX)
    // log(1) is exactly +0, so the select changes no bit; it lets the compiler drop the logarithm where sigma is the
    // literal 1 (CUDA's log() is inline code, not a foldable builtin: the multilevel model spent half of ALL its
    // instructions evaluating log(1.0), round 2 ncu)
    if (lg) return -(DRJ_LN_SQRT_2PI + 0.5 * z * z + (sigma == 1.0 ? 0.0 : log(sigma)));
    return 0.398942280401432677939946059934 * exp(-0.5 * z * z) / sigma;
  }
  if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
  if (sigma < 0) return DRJ_NAN;
  if (!isfinite(sigma)) return drj_nmath::d0(lg);
  if (!isfinite(x) && mu == x) return DRJ_NAN;
  if (sigma == 0) return (x == mu) ? DRJ_INF : drj_nmath::d0(lg);
  return drj_nmath::d0(lg);  // |z| infinite or beyond 2 sqrt(DBL_MAX)
}

__device__ __forceinline__ double dlnorm(double x, double meanlog, double sdlog, int lg) {
  if (isnan(x) || isnan(meanlog) || isnan(sdlog)) return x + meanlog + sdlog;
  if (sdlog < 0) return DRJ_NAN;
  if (!isfinite(x) && log(x) == meanlog) return DRJ_NAN;
  if (sdlog == 0) return (log(x) == meanlog) ? DRJ_INF : drj_nmath::d0(lg);
  if (x <= 0) return drj_nmath::d0(lg);
  double y = (log(x) - meanlog) / sdlog;
  return lg ? -(DRJ_LN_SQRT_2PI + 0.5 * y * y + log(x * sdlog))
            : 0.398942280401432677939946059934 * exp(-0.5 * y * y) / (x * sdlog);
}

// dlnorm with log(x) supplied by the caller (bit for bit the value log(x) would give: the sampler's cached
// Jacobian logarithm of a parameter whose lower bound is 0, see DrjHints)
__device__ __forceinline__ double dlnorm_l(double x, double logx, double meanlog, double sdlog, int lg) {
  if (isnan(x) || isnan(meanlog) || isnan(sdlog)) return x + meanlog + sdlog;
  if (sdlog < 0) return DRJ_NAN;
  if (!isfinite(x) && logx == meanlog) return DRJ_NAN;
  if (sdlog == 0) return (logx == meanlog) ? DRJ_INF : drj_nmath::d0(lg);
  if (x <= 0) return drj_nmath::d0(lg);
  double y = (logx - meanlog) / sdlog;
  return lg ? -(DRJ_LN_SQRT_2PI + 0.5 * y * y + (sdlog == 1.0 ? logx : log(x * sdlog)))
            : 0.398942280401432677939946059934 * exp(-0.5 * y * y) / (x * sdlog);
}

__device__ inline double dpois(double x, double lambda, int lg) {
  if (isnan(x) || isnan(lambda)) return x + lambda;
  if (lambda < 0) return DRJ_NAN;
  if (drj_nmath::nonint(x)) return drj_nmath::d0(lg);
  if (x < 0 || !isfinite(x)) return drj_nmath::d0(lg);
  return drj_nmath::dpois_raw(nearbyint(x), lambda, lg);
}

// dpois(x, lambda, log = TRUE) for a regular count x (not NaN, a non-negative finite integer: xr = nearbyint(x)) with its
// data-only parts supplied: st = stirlerr(xr), lf = log(2 pi xr).  A model that evaluates the same counts at every
// update (a Poisson likelihood over a fixed data vector) computes them once per thread block (DrjCtaInit): stirlerr is
// three to five FP64 divisions, and with the logarithm a third of a term.  Same expressions as dpois_raw / dfexp: same bits.
__device__ __forceinline__ double dpois_pre(double xr, double lambda, double st, double lf) {
  if (lambda > 0.0 && lambda < DRJ_INF && !(xr <= lambda * DRJ_DBL_MIN) && !(lambda < xr * DRJ_DBL_MIN))
    return -0.5 * lf + (-st - drj_nmath::bd0(xr, lambda));
  return dpois(xr, lambda, 1);
}

__device__ inline double dgamma(double x, double shape, double scale, int lg) {
  if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
  if (shape < 0 || scale <= 0) return DRJ_NAN;
  if (x < 0) return drj_nmath::d0(lg);
  if (shape == 0) return (x == 0) ? DRJ_INF : drj_nmath::d0(lg);
  if (x == 0) {
    if (shape < 1) return DRJ_INF;
    if (shape > 1) return drj_nmath::d0(lg);
    return lg ? -log(scale) : 1 / scale;
  }
  double pr;
  if (shape < 1) {
    pr = drj_nmath::dpois_raw(shape, x / scale, lg);
    return lg ? pr + (isfinite(shape / x) ? log(shape / x) : log(shape) - log(x)) : pr * shape / x;
  }
  pr = drj_nmath::dpois_raw(shape - 1, x / scale, lg);
  return lg ? pr - log(scale) : pr / scale;
}

__device__ inline double dbinom(double x, double n, double p, int lg) {
  if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
  if (p < 0 || p > 1 || n < 0 || drj_nmath::nonint(n)) return DRJ_NAN;
  if (drj_nmath::nonint(x)) return drj_nmath::d0(lg);
  if (x < 0 || !isfinite(x)) return drj_nmath::d0(lg);
  return drj_nmath::dbinom_raw(nearbyint(x), nearbyint(n), p, 1 - p, lg);
}

__device__ inline double dbeta(double x, double a, double b, int lg) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (a < 0 || b < 0) return DRJ_NAN;
  if (x < 0 || x > 1) return drj_nmath::d0(lg);
  if (a == 0 || b == 0 || !isfinite(a) || !isfinite(b)) {
    if (a == 0 && b == 0) return (x == 0 || x == 1) ? DRJ_INF : drj_nmath::d0(lg);
    if (a == 0 || a / b == 0) return (x == 0) ? DRJ_INF : drj_nmath::d0(lg);
    if (b == 0 || b / a == 0) return (x == 1) ? DRJ_INF : drj_nmath::d0(lg);
    return (x == 0.5) ? DRJ_INF : drj_nmath::d0(lg);
  }
  if (x == 0) {
    if (a > 1) return drj_nmath::d0(lg);
    if (a < 1) return DRJ_INF;
    return lg ? log(b) : b;
  }
  if (x == 1) {
    if (b > 1) return drj_nmath::d0(lg);
    if (b < 1) return DRJ_INF;
    return lg ? log(a) : a;
  }
  double lval;
  if (a <= 2 || b <= 2)
    lval = (a - 1) * log(x) + (b - 1) * log1p(-x) - (lgamma(a) + lgamma(b) - lgamma(a + b));
  else
    lval = log(a + b - 1) + drj_nmath::dbinom_raw(a - 1, a + b - 2, x, 1 - x, 1);
  return drj_nmath::dexp_(lval, lg);
}

__device__ __forceinline__ double dunif(double x, double a, double b, int lg) {
  if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
  if (b <= a) return DRJ_NAN;
  if (a <= x && x <= b) return lg ? -log(b - a) : 1.0 / (b - a);
  return drj_nmath::d0(lg);
}

__device__ __forceinline__ double dexp(double x, double scale, int lg) {
  if (isnan(x) || isnan(scale)) return x + scale;
  if (scale <= 0.0) return DRJ_NAN;
  if (x < 0.0) return drj_nmath::d0(lg);
  return lg ? (-x / scale) - log(scale) : exp(-x / scale) / scale;
}
}  // namespace R
