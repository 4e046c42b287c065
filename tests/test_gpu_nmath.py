"""R::dnorm on the device against the oracle's restatement of R's dnorm4 (src/nmath/dnorm.c), argument by argument:
the regular range, where the device multiplies by a hoisted reciprocal (within an ulp of z), and every
edge case of R's argument checks, where the answers must be the same special values."""
import numpy as np
import pytest

from helpers import Case, run_gpu

pytestmark = pytest.mark.gpu

SRC = r'''
// one term per chain: chain c evaluates R::dnorm(x[k], mu[k], sigma[k], log) with k = (int) theta[0];
// special values are returned as finite codes (the sampler rejects NaN / Inf at the initial theta)
__device__ double code(double v) { return isnan(v) ? 1e300 : (isinf(v) ? (v > 0 ? 2e300 : -2e300) : v); }
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  const int k = (int)params[0];
  double last = 0.0;
  for (int rep = 0; rep < 3; ++rep)   // (a loop with loop-invariant arguments, as in a user's sum over the data)
    last = R::dnorm(data.item(0)[k], data.item(1)[k], data.item(2)[k], true);
  return code(last);
}
__device__ double logprior(const double* params, const DrjList& misc) { return 0.0; }
'''


def test_device_dnorm_follows_r(oracle):
    import drjacoby_b200 as drj
    inf, nan = np.inf, np.nan
    args = [(0.3, 0.1, 2.0), (1e3, -1e3, 1e-3), (5.0, 5.0, 1.0), (-7.5, 2.25, 1.0), (1.0, 0.0, 1e-300), (1.0, 1.0, 1e-300),
            (1.0, 0.0, 1e-310), (1.0, 1.0, 1e-310), (1.0, 0.0, 5e-324), (3.0, 1.0, 1e300), (3.0, 1.0, 1e305), (3.0, 1.0, 1.7e308),
            (1e308, -1e308, 2.0), (1e200, 0.0, 1e-200), (1e160, 0.0, 1.0), (1.0, 0.0, 0.0), (0.0, 0.0, 0.0), (1.0, 0.0, -1.0),
            (1.0, 0.0, inf), (nan, 0.0, 1.0), (1.0, nan, 2.0), (1.0, 0.0, nan), (inf, 0.0, 2.0), (-inf, 0.0, 2.0), (inf, inf, 2.0),
            (inf, -inf, 2.0), (inf, 0.0, inf), (2.0, 1.0, 2.2250738585072014e-308), (2.0, 1.0, 2.2250738585072009e-308)]
    rng = np.random.Generator(np.random.Philox(11))
    args += [(float(rng.normal(0, 10)), float(rng.normal(0, 10)), float(np.exp(rng.normal(0, 6)))) for _ in range(60)]
    K = len(args)
    x, mu, sg = (np.array([a[j] for a in args]) for j in range(3))
    case = Case("jit", [0.0], [float(K)], data=[x, mu, sg], chains=K, burnin=2, samples=1, init=np.arange(K) + 0.5)
    out = run_gpu(case, "philox", seed=1, model=drj.Model(source=SRC))
    got = out["loglike_burnin"][:, 0, 0]
    L = oracle.lib()
    for k, (a, b, c) in enumerate(args):
        want = L.ora_dnorm(a, b, c, 1)
        g = got[k]
        if np.isnan(want):
            assert g == 1e300, (args[k], g)
        elif np.isinf(want):
            assert g == (2e300 if want > 0 else -2e300), (args[k], g, want)
        else:
            assert abs(g - want) <= 4e-16 * max(1.0, abs(want)) + 1e-15 * abs(0.5 * ((a - b) / c) ** 2), (args[k], g, want)


DIV_SRC = r'''
// drj_div_odd(a, j) (the division-free quotient of nmath's bd0 series) against the division itself, bit for bit:
// 32 divisors x 200,000 pseudo-random dividends (mantissas from an LCG, exponents -270 .. 270, both signs) per evaluation
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  unsigned long long state = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)(params[0] * 1e6);
  int bad = 0;
  for (int it = 0; it < 200000; ++it) {
    state = state * 6364136223846793005ULL + 1442695040888963407ULL;
    const unsigned long long mant = (state >> 12) | 0x0010000000000000ULL;
    const int e = (int)((state >> 3) % 541ULL) - 270;
    double a = ldexp((double)mant, e - 52);
    if (state & 4ULL) a = -a;
    for (int j = 0; j < 32; ++j) {
      const double q1 = drj_nmath::drj_div_odd(a, j), q2 = a / (double)(2 * j + 1);
      bad += (__double_as_longlong(q1) != __double_as_longlong(q2)) ? 1 : 0;
    }
  }
  return -(double)bad;
}
__device__ double logprior(const double* params, const DrjList& misc) { return 0.0; }
'''


def test_division_free_quotient_is_the_quotient():
    import drjacoby_b200 as drj
    case = Case("jit", [0.0], [1.0], chains=8, burnin=2, samples=1)
    out = run_gpu(case, "philox", seed=1, model=drj.Model(source=DIV_SRC))
    # 8 chains x (initial evaluation + 2 proposals) x 6.4e6 comparisons, every one equal
    assert np.all(out["loglike_burnin"] == 0.0) and np.all(out["loglike_sampling"] == 0.0)
