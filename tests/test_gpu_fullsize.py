"""BASELINE.json's configs at (or near) their full sizes.

Where the oracle finishes in seconds the full-size run is compared with it directly (chains are
independent and Philox is keyed by the global chain id, so the oracle only needs to run a SUBSET of the
chains of the GPU run).  For K5 (n = 1e7) the reference's sequential double sum itself carries ~1e-12
relative rounding noise, so parity is checked through size-independent properties: the likelihood
against sufficient statistics in extended precision, additivity over a data split, bitwise agreement of
the TMA and plain staging paths, and sharding invariance.
"""
import numpy as np
import pytest

from helpers import CANCELLING_ATOL, Case, POPULATION, assert_parity, beta_ladder, gpu_spec, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def subset(case: Case, lo, hi):
    return Case(case.model, case.tmin, case.tmax, data=case.data, misc=case.misc, blocks=case.blocks, beta=case.beta,
                chains=hi - lo, burnin=case.burnin, samples=case.samples, coupling_on=case.coupling_on,
                save_hot_draws=case.save_hot_draws, target=case.target, init=case.init[lo:hi])


def compare_subset(gpu, ora, lo, hi, rtol=1e-10, what="", atol=0.0):
    part = {k: (v[lo:hi] if k != "t_diff" else v) for k, v in gpu.items()}
    assert_parity(part, ora, rtol, what, atol=atol)


def test_k2_double_well_full_size(oracle):
    """configs[1]: double well, 64 chains x 20 rungs, alpha = 2, burnin 1e3 + samples 1e4 -- every chain
    against the oracle (Philox mode)."""
    case = Case("doublewell", [-2, 30], [2, 30], beta=beta_ladder(20, 2.0), chains=64, burnin=1000, samples=10000)
    gpu = run_gpu(case, "philox", seed=11)
    ora = run_oracle(oracle, case, "philox", seed=11, n_threads=16)
    assert_parity(gpu, ora, 1e-10, what="[K2 full] ", atol=CANCELLING_ATOL["doublewell"])


def test_k3_logistic_full_size(oracle):
    """configs[2]: logistic growth, 1024 chains x 10 rungs, 1e3 + 1e4 on the GPU; the oracle re-runs 24 of
    those chains (first, middle, last blocks)."""
    case = Case("logistic", [1, 1, 50], [100, 2, 200], data=[POPULATION["pop"], POPULATION["time"]], beta=beta_ladder(10),
                chains=1024, burnin=1000, samples=10000)
    gpu = run_gpu(case, "philox", seed=5)
    for lo in (0, 508, 1016):
        ora = run_oracle(oracle, subset(case, lo, lo + 8), "philox", seed=5, chain_offset=lo, n_threads=8)
        compare_subset(gpu, ora, lo, lo + 8, what=f"[K3 full, chains {lo}..] ")


def test_k4_multilevel_full_width(oracle):
    """configs[3]: 50 parameters in overlapping blocks (51 blocks), 4096 chains x 16 rungs; 300 + 300
    iterations on the GPU, 4 chains re-run by the oracle."""
    G = 50
    rng = np.random.Generator(np.random.Philox(1))
    mu = rng.normal(0, 1, G)
    data = [rng.normal(mu[g], 1.0, 5) for g in range(G)]
    case = Case("multilevel", [-5.0] * G, [5.0] * G, data=data, blocks=[[g + 1, G + 1] for g in range(G)],
                beta=beta_ladder(16), chains=4096, burnin=300, samples=300)
    spec_kw = dict(store_pt=True)
    gpu = run_gpu(case, "philox", seed=3, **spec_kw)
    for lo in (0, 4092):
        ora = run_oracle(oracle, subset(case, lo, lo + 4), "philox", seed=3, chain_offset=lo, n_threads=4)
        compare_subset(gpu, ora, lo, lo + 4, what=f"[K4, chains {lo}..] ", atol=CANCELLING_ATOL["multilevel"])
    # posterior check (vignettes/checks_multilevel_blocks.Rmd): mu_i ~ N(m_i n/(n+1), 1/(n+1))
    th = gpu["theta_sampling"][:, 0]
    for g in (0, 17, 49):
        m = np.mean(data[g])
        assert abs(th[:, 100:, g].mean() - m * 5 / 6) < 0.01


def loglike_from_sufficient_stats(x, mu, sigma):
    """sum_i dnorm(x_i, mu, sigma, log=TRUE) via sufficient statistics in extended precision."""
    xl = x.astype(np.longdouble)
    n, s1, s2 = np.longdouble(x.size), xl.sum(), (xl * xl).sum()
    mu, sigma = np.asarray(mu, np.longdouble), np.asarray(sigma, np.longdouble)
    return np.asarray(-n * (np.longdouble(0.918938533204672741780329736406) + np.log(sigma))
                      - (s2 - 2 * mu * s1 + n * mu * mu) / (2 * sigma * sigma), dtype=np.float64)


@pytest.fixture(scope="module")
def big_x():
    return np.random.Generator(np.random.Philox(1)).normal(3.0, 2.0, 10_000_000)


def test_k5_likelihood_against_sufficient_statistics(big_x):
    """configs[4] data size (n = 1e7), streamed through the TMA tile pipeline: every stored log-likelihood
    must equal the closed form evaluated from (n, sum x, sum x^2) in long double."""
    case = Case("example", [-10, 0], [10, np.inf], data=[big_x], beta=beta_ladder(32), chains=64, burnin=2, samples=2,
                save_hot_draws=True)
    out = run_gpu(case, "philox", seed=1)
    th = np.concatenate([out["theta_burnin"], out["theta_sampling"]], axis=2)   # [chains, rungs, it, d]
    ll = np.concatenate([out["loglike_burnin"], out["loglike_sampling"]], axis=2)
    want = loglike_from_sufficient_stats(big_x, th[..., 0], th[..., 1])
    rel = np.abs(ll - want) / np.abs(want)
    assert rel.max() < 1e-10, rel.max()
    assert out["accept_sampling"].sum() > 0


def test_k5_streaming_paths_agree_bitwise(big_x):
    from drjacoby_b200 import _lib
    case = Case("example", [-10, 0], [10, np.inf], data=[big_x], beta=beta_ladder(32), chains=16, burnin=2, samples=2)
    a = run_gpu(case, "philox", seed=2)
    b = run_gpu(case, "philox", seed=2, flags=_lib.FLAG_NO_TMA)
    c = run_gpu(case, "philox", seed=2, chains_per_cta=8)
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k
    # a shard of the same run (chains 4..12 of 16) reproduces its slice
    part = run_gpu(subset(case, 4, 12), "philox", seed=2, chain_offset=4)
    for k in part:
        if k != "t_diff":
            assert np.array_equal(part[k], a[k][4:12]), k


def test_k5_additivity_over_a_data_split(big_x):
    """loglike(x) == loglike(x[:m]) + loglike(x[m:]) at the initial theta (linearity of the sweep)."""
    init = np.array([[2.5, 1.5], [3.2, 2.2], [-4.0, 7.0]])
    m = 3_999_992
    ll = []
    for xs in (big_x, big_x[:m], big_x[m:]):
        case = Case("example", [-10, 0], [10, np.inf], data=[xs], chains=3, burnin=1, samples=1, init=init)
        ll.append(run_gpu(case, "philox", seed=1)["loglike_burnin"][:, 0, 0])
    np.testing.assert_allclose(ll[0], ll[1] + ll[2], rtol=1e-12)


def test_small_n_sweep_shape_matches_oracle(oracle):
    """The bench's n = 100 sweep shape (32 rungs) on 2048 chains x 200 iterations; 16 chains re-run by the oracle."""
    x = np.random.Generator(np.random.Philox(1)).normal(3.0, 2.0, 100)
    case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(32), chains=2048, burnin=100, samples=100)
    gpu = run_gpu(case, "philox", seed=9)
    for lo in (0, 2040):
        ora = run_oracle(oracle, subset(case, lo, lo + 8), "philox", seed=9, chain_offset=lo, n_threads=8)
        compare_subset(gpu, ora, lo, lo + 8, what=f"[n=100 sweep, chains {lo}..] ")
