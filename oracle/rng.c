/*
 * rng.c -- ORACLE (test infrastructure, see drj_oracle.h).
 *
 * Random-number sources of the CPU oracle:
 *  (1) R's default RNG: Mersenne-Twister + Inversion.  Not vendored by the reference (base R,
 *      src/main/RNG.c, src/nmath/snorm.c, src/nmath/qnorm.c); called by the reference at
 *      src/Particle.cpp:49 (R::rnorm), src/Particle.h:139 and src/main.cpp:311 (R::runif),
 *      R/main.R:302 (runif).  Restated from the published algorithms (Matsumoto & Nishimura 1998;
 *      Wichura 1988, AS241 PPND16) and pinned numerically by tests/test_oracle_golden.py.
 *  (2) Philox4x32-10 (Salmon et al. 2011), the counter-based generator the GPU path uses.
 */
#include <math.h>
#include <string.h>

#include "drj_oracle.h"

/* ---------------------------------------------------------------- R Mersenne-Twister */

#define MT_N 624
#define MT_M 397
#define MT_MATRIX_A 0x9908b0dfu
#define MT_UPPER 0x80000000u
#define MT_LOWER 0x7fffffffu

/* set.seed(seed): initial scrambling then LCG fill of the 625-word seed vector; word 0 is the
 * position and is forced to N so the first draw regenerates the table. */
void ora_rmt_set_seed(ora_rmt* s, uint32_t seed) {
  for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;
  for (int j = 0; j < MT_N + 1; ++j) {
    seed = 69069u * seed + 1u;
    if (j > 0) s->mt[j - 1] = seed;
  }
  s->mti = MT_N;
}

static uint32_t rmt_next_u32(ora_rmt* s) {
  static const uint32_t mag01[2] = {0x0u, MT_MATRIX_A};
  uint32_t y;
  uint32_t* mt = s->mt;
  if (s->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (mt[kk] & MT_UPPER) | (mt[kk + 1] & MT_LOWER);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 0x1u];
    }
    for (; kk < MT_N - 1; kk++) {
      y = (mt[kk] & MT_UPPER) | (mt[kk + 1] & MT_LOWER);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 0x1u];
    }
    y = (mt[MT_N - 1] & MT_UPPER) | (mt[0] & MT_LOWER);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 0x1u];
    s->mti = 0;
  }
  y = mt[s->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* R's fixup(): never return exactly 0 or 1 */
static double r_fixup(double x) {
  const double i2_32m1 = 2.328306437080797e-10;
  if (x <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return x;
}

double ora_rmt_unif_rand(ora_rmt* s) {
  return r_fixup((double)rmt_next_u32(s) * 2.3283064365386963e-10);
}

/* norm_rand(), normal.kind = "Inversion": two uniforms give a 53-bit-ish argument to qnorm */
double ora_rmt_norm_rand(ora_rmt* s) {
  const double BIG = 134217728.0; /* 2^27 */
  double u = ora_rmt_unif_rand(s);
  u = (int)(BIG * u) + ora_rmt_unif_rand(s);
  return ora_qnorm(u / BIG);
}

void ora_rmt_runif(ora_rmt* s, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = ora_rmt_unif_rand(s);
}

void ora_rmt_rnorm(ora_rmt* s, int n, double mean, double sd, double* out) {
  for (int i = 0; i < n; ++i) out[i] = mean + sd * ora_rmt_norm_rand(s);
}

/* ---------------------------------------------------------------- qnorm (Wichura AS241) */

double ora_qnorm(double p) {
  if (isnan(p)) return p;
  if (p < 0.0 || p > 1.0) return NAN;
  if (p == 0.0) return -INFINITY;
  if (p == 1.0) return INFINITY;
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
  if (q < 0.0) val = -val;
  return val;
}

/* ---------------------------------------------------------------- Philox4x32-10 */

void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout shared with the GPU path (DESIGN.md "RNG"):
 *   key     = (seed_lo, seed_hi)
 *   counter = (chain_lo32, iteration, rung << 16 | slot, tag | chain_hi << 8)
 *   tag 0: parameter update (slot = parameter index): u[0],u[1] -> rnorm by inversion, u[2] -> accept
 *   tag 1: coupling (slot = pair index): u[0] -> swap decision
 * Words become uniforms exactly as R's unif_rand does (32-bit resolution, fixup into (0,1)). */
void ora_philox_uniforms(uint64_t seed, int64_t chain, uint32_t iteration, uint32_t rung,
                         uint32_t slot, uint32_t tag, double u[4]) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {(uint32_t)(uint64_t)chain, iteration, (rung << 16) | (slot & 0xFFFFu),
                     tag | ((uint32_t)((uint64_t)chain >> 32) << 8)};
  uint32_t w[4];
  ora_philox4x32_10(ctr, key, w);
  for (int k = 0; k < 4; ++k) u[k] = r_fixup((double)w[k] * 2.3283064365386963e-10);
}
