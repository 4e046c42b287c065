"""The oracle's restatement of R's nmath densities against scipy.stats (independent implementations
of the same distributions) and of qnorm against scipy's inverse normal CDF."""
import numpy as np
import pytest
from scipy import stats


def test_qnorm_matches_scipy(oracle):
    L = oracle.lib()
    p = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 2001), 10.0 ** -np.arange(13, 300, 7.0)])
    got = np.array([L.ora_qnorm(float(v)) for v in p])
    want = stats.norm.ppf(p)
    np.testing.assert_allclose(got, want, rtol=2e-13, atol=1e-15)
    assert L.ora_qnorm(0.0) == -np.inf and L.ora_qnorm(1.0) == np.inf and np.isnan(L.ora_qnorm(1.5))


@pytest.mark.parametrize("name,args,ref", [
    ("ora_dnorm", (3.0, 2.0), lambda x, a, b: stats.norm.logpdf(x, a, b)),
    ("ora_dlnorm", (0.3, 1.2), lambda x, a, b: stats.lognorm.logpdf(x, b, scale=np.exp(a))),
    ("ora_dgamma", (5.0, 5.0), lambda x, a, b: stats.gamma.logpdf(x, a, scale=b)),
    ("ora_dgamma", (0.01, 100.0), lambda x, a, b: stats.gamma.logpdf(x, a, scale=b)),
    ("ora_dgamma", (1.0, 2.0), lambda x, a, b: stats.gamma.logpdf(x, a, scale=b)),
    ("ora_dbeta", (3.0, 3.0), lambda x, a, b: stats.beta.logpdf(x, a, b)),
    ("ora_dbeta", (0.5, 1.5), lambda x, a, b: stats.beta.logpdf(x, a, b)),
    ("ora_dunif", (1.0, 100.0), lambda x, a, b: stats.uniform.logpdf(x, a, b - a)),
])
def test_continuous_densities(oracle, name, args, ref):
    f = getattr(oracle.lib(), name)
    xs = np.concatenate([np.linspace(0.001, 0.999, 41), np.linspace(1.0, 60.0, 41), [1e-8, 250.0]])
    for x in xs:
        want = ref(x, *args)
        got = f(float(x), *args, 1)
        if np.isfinite(want):
            assert abs(got - want) <= 2e-12 * max(1.0, abs(want)), (name, x, got, want)
            assert abs(f(float(x), *args, 0) - np.exp(want)) <= 1e-11 * max(np.exp(want), 1e-300)
        else:
            assert got == want, (name, x, got, want)


def test_discrete_densities(oracle):
    L = oracle.lib()
    for lam in (0.3, 4.0, 37.5, 117.0, 2500.0):
        for x in (0, 1, 5, 14, 15, 16, 36, 88, 117, 2600):
            want = stats.poisson.logpmf(x, lam)
            assert abs(L.ora_dpois(float(x), lam, 1) - want) <= 2e-12 * max(1.0, abs(want)), (x, lam)
    assert L.ora_dpois(2.5, 3.0, 1) == -np.inf and L.ora_dpois(-1.0, 3.0, 1) == -np.inf
    for n, p in ((10, 0.3), (200, 0.05), (50, 0.97)):
        for x in (0, 1, 7, n):
            want = stats.binom.logpmf(x, n, p)
            assert abs(L.ora_dbinom(float(x), float(n), p, 1) - want) <= 2e-12 * max(1.0, abs(want))


def test_edge_cases_follow_r(oracle):
    L = oracle.lib()
    assert L.ora_dnorm(1.0, 0.0, 0.0, 1) == -np.inf and L.ora_dnorm(0.0, 0.0, 0.0, 1) == np.inf  # sigma == 0
    assert np.isnan(L.ora_dnorm(1.0, 0.0, -1.0, 1))
    assert L.ora_dnorm(1.0, 0.0, np.inf, 1) == -np.inf
    assert L.ora_dlnorm(0.0, 0.0, 1.0, 1) == -np.inf and L.ora_dlnorm(-1.0, 0.0, 1.0, 1) == -np.inf
    assert L.ora_dgamma(-1.0, 5.0, 5.0, 1) == -np.inf and L.ora_dgamma(0.0, 0.5, 1.0, 1) == np.inf
    assert L.ora_dbeta(1.5, 3.0, 3.0, 1) == -np.inf and L.ora_dbeta(0.0, 3.0, 3.0, 1) == -np.inf
    assert L.ora_dunif(0.5, 1.0, 2.0, 1) == -np.inf and np.isnan(L.ora_dunif(0.5, 2.0, 1.0, 1))


def test_models_against_closed_forms(oracle):
    """Each oracle model against the same density written with scipy."""
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(0)
    x = rng.normal(3, 2, 50)
    dl, keep = oracle.make_list([x])
    ml, keep2 = oracle.make_list([])
    th = np.array([2.5, 1.7])
    got = L.ora_model_loglike(b"example", 2, th.ctypes.data_as(C.POINTER(C.c_double)), C.byref(dl), C.byref(ml), 1)
    assert abs(got - stats.norm.logpdf(x, 2.5, 1.7).sum()) < 1e-10
    got = L.ora_model_logprior(b"example", 2, th.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ml))
    assert abs(got - (-np.log(20) + stats.lognorm.logpdf(1.7, 1.0))) < 1e-12
    th = np.array([0.3, -0.7, 4.0, 0.4])
    got = L.ora_model_logprior(b"returnprior", 4, th.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ml))
    want = (stats.norm.logpdf(0.3) + stats.gamma.logpdf(0.7, 5, scale=5) + stats.gamma.logpdf(4.0, 5, scale=5)
            + stats.beta.logpdf(0.4, 3, 3))
    assert abs(got - want) < 1e-11
    from helpers import POPULATION
    pl, keep3 = oracle.make_list([POPULATION["pop"], POPULATION["time"]])
    th = np.array([20.0, 1.1, 110.0])
    got = L.ora_model_loglike(b"logistic", 3, th.ctypes.data_as(C.POINTER(C.c_double)), C.byref(pl), C.byref(ml), 1)
    K = 110.0 * 1.1 / 0.1
    xs = [20.0]
    for _ in range(99):
        xs.append(1.1 * xs[-1] * (1 - xs[-1] / K))
    want = sum(stats.poisson.logpmf(p, xs[t - 1]) for p, t in zip(POPULATION["pop"], POPULATION["time"]))
    assert abs(got - want) < 1e-9
