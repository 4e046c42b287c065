"""The public run_mcmc() path on the GPU: the reference's own published numbers reproduced by the CUDA
sampler (R-compatible replay stream), user models compiled with NVRTC, and the statistical checks of
the reference's test-suite / "checks" vignettes in normal (Philox) mode."""
import json
import os

import numpy as np
import pandas as pd
import pytest
from scipy import stats

import drjacoby_b200 as drj
from drjacoby_b200 import api

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vignettes.json")))

NORMAL_SRC = r"""
// inst/extdata/example_loglike_logprior.cpp rewritten against cuda_template.cu
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  double mu = params[P_mu];
  double sigma = params[P_sigma];
  const double* x = data.item(DATA_x);
  double ret = 0.0;
  for (long long i = 0; i < data.len(DATA_x); ++i) ret += R::dnorm(x[i], mu, sigma, true);
  return ret;
}
__device__ double logprior(const double* params, const DrjList& misc) {
  return -log(20.0) + R::dlnorm(params[P_sigma], 0.0, 1.0, true);
}
"""


def vignette_setup():
    rs = drj.RStream(1)
    data = {"x": rs.rnorm(10, 3, 2)}
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    return rs, data, df


def rates_of(res, burnin, samples, d):
    return [[api._printed_rate(res.raw["accept_burnin"][c, -1], burnin, d), api._printed_rate(res.raw["accept_sampling"][c, -1], samples, d)]
            for c in range(res.raw["accept_burnin"].shape[0])]


@pytest.mark.parametrize("builtin", ["example", "example_exact"])
def test_example_vignette_on_the_gpu(builtin):
    """docs/articles/example.html: head(mcmc$output), 20 acceptance rates, Rhat, ESS -- from the GPU."""
    g = GOLD["example"]
    rs, data, df = vignette_setup()
    drj.use_builtin(builtin)
    kw = dict(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=1e3, samples=1e3, rng="r", seed=rs, silent=True)
    res = drj.run_mcmc(**kw)
    head = res.output.head(6)
    for row, want in zip(head.itertuples(index=False), g["head_rows_chain1"]["rows"]):
        for have, w in zip((row.mu, row.sigma, row.logprior, row.loglikelihood), want[1:]):
            digits = len(repr(w).split(".")[1])
            assert abs(have - w) <= 0.5000001 * 10.0 ** (-digits), (row, want)
    assert rates_of(res, 1000, 1000, 2) == g["acceptance_pct_run1"]
    assert [float(res.diagnostics["rhat"][k]) for k in ("mu", "sigma")] == g["rhat"]
    np.testing.assert_allclose([res.diagnostics["ess"][k] for k in ("mu", "sigma")], g["ess"], atol=5e-5)
    res2 = drj.run_mcmc(**kw)  # the vignette's second run continues the stream
    assert rates_of(res2, 1000, 1000, 2) == g["acceptance_pct_run2"]


def test_coupling_vignette_on_the_gpu():
    """docs/articles/metropolis_coupling.html:284,287: nine 1-rung runs, then 20 rungs: 42.5 % / 44 %."""
    rs = drj.RStream(1)
    data = {"x": rs.rnorm(100, 10, 1)}
    df = drj.define_params("name", "alpha", "min", -10, "max", 10, "init", 5, "name", "beta", "min", 0, "max", 10, "init", 5,
                           "name", "epsilon", "min", -np.inf, "max", np.inf, "init", 0)
    drj.use_builtin("coupling")
    kw = dict(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=1e3, samples=1e4, chains=1, rng="r",
              seed=rs, silent=True, as_frames=False)
    for _ in range(9):
        drj.run_mcmc(**kw)
    res = drj.run_mcmc(rungs=20, **kw)
    assert rates_of(res, 1000, 10000, 3)[0] == GOLD["metropolis_coupling"]["acceptance_pct_all_runs"][:2]


def test_double_well_vignette_on_the_gpu():
    """docs/articles/checks_double_well.html:203,206: 11 rungs, alpha = 2, fixed gamma: 21.7 % / 22 %."""
    rs = drj.RStream(1)
    df = drj.define_params("name", "mu", "min", -2, "max", 2, "name", "gamma", "min", 30, "max", 30)
    drj.use_builtin("doublewell")
    kw = dict(data={"x": [-1.0]}, df_params=df, loglike="loglike", logprior="logprior", burnin=1e3, samples=1e5, chains=1,
              rng="r", seed=rs, silent=True, as_frames=False)
    drj.run_mcmc(rungs=1, **kw)
    res = drj.run_mcmc(rungs=11, alpha=2, **kw)
    assert rates_of(res, 1000, 100000, 2)[0] == GOLD["double_well"]["acceptance_pct_all_runs"][:2]


def test_nvrtc_user_model_equals_builtin():
    """A model written against cuda_template.cu, compiled at run time with NVRTC for sm_100a, gives the
    same bits as the nvcc-compiled built-in with the same arithmetic (example_exact)."""
    x = np.random.default_rng(2).normal(3, 2, 64)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=100, samples=200, rungs=4, chains=6,
              seed=99, silent=True)
    src = drj.source_cuda(NORMAL_SRC)
    a = drj.run_mcmc(**kw)
    drj.use_builtin("example_exact")
    b = drj.run_mcmc(**kw)
    for k in a.raw:
        if k != "t_diff":
            assert np.array_equal(a.raw[k], b.raw[k]), k
    assert src.model("loglike", "logprior").compile_seconds() >= 0.0


DATUM_SRC = r"""
// a user model in the per-datum form (cuda_template.cu, second half): same arithmetic as the built-in "example"
struct MyModel {
  static constexpr int D = DRJ_D, NBLOCK = DRJ_NBLOCK;
  static constexpr bool DATUM_FORM = true;
  struct Consts { double mean, half_inv_var, lognorm; };
  __device__ static int datum_item(int block) { return DATA_x; }
  __device__ static bool prepare(const double* params, const DrjList& misc, int block, Consts& k) {
    const double sd = params[P_sigma];
    k.mean = params[P_mu];
    k.half_inv_var = 0.5 / (sd * sd);
    k.lognorm = 0.918938533204672741780329736406 + log(sd);
    return (sd > 1e-150) && (sd < 1e150) && (fabs(k.mean) < 1e150);
  }
  __device__ static double datum(const Consts& k, double x, double acc) { const double d = x - k.mean; return fma(d, d, acc); }
  __device__ static double finish(const Consts& k, double acc, long long n) { return -((double)n * k.lognorm + k.half_inv_var * acc); }
  __device__ static double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
    double r = 0.0;
    for (long long i = 0; i < data.len(DATA_x); ++i) r += R::dnorm(data.item(DATA_x)[i], params[P_mu], params[P_sigma], true);
    return r;
  }
  __device__ static double logprior(const double* params, const DrjList& misc) { return -log(20.0) + R::dlnorm(params[P_sigma], 0.0, 1.0, true); }
};
#define DRJ_USER_MODEL MyModel
"""


@pytest.mark.parametrize("n", [100, 3000, 10001])
def test_nvrtc_datum_form_user_model(n):
    """A user model in the per-datum form through NVRTC: data resident in shared memory (n = 100), the
    state-parking kernel variant (n = 3000) and TMA-streamed tiles (n = 10001); bitwise the built-in."""
    x = np.random.default_rng(4).normal(3, 2, n)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=10, samples=15, rungs=4, chains=5,
              seed=17, silent=True, as_frames=False)
    drj.source_cuda(DATUM_SRC)
    a = drj.run_mcmc(**kw)
    drj.use_builtin("example")
    b = drj.run_mcmc(**kw)
    for k in a.raw:
        if k != "t_diff":
            assert np.array_equal(a.raw[k], b.raw[k]), k


HINTS_SRC = DATUM_SRC.replace("#define DRJ_USER_MODEL MyModel", r"""
// the same model with the HINTS hook of cuda_template.cu: log(sigma) from the sampler's Jacobian logarithm where sigma's
// lower bound is 0 (bit P_sigma of mask), else taken here
struct MyHintedModel : MyModel {
  static constexpr bool HINTS = true;
  __device__ static bool prepare_h(const double* params, const DrjList& misc, int block, Consts& k, const double* loglo, unsigned mask) {
    const double sd = params[P_sigma];
    k.mean = params[P_mu];
    k.half_inv_var = 0.5 / (sd * sd);
    k.lognorm = 0.918938533204672741780329736406 + ((mask >> P_sigma & 1u) ? loglo[P_sigma] : log(sd));
    return (sd > 1e-150) && (sd < 1e150) && (fabs(k.mean) < 1e150);
  }
  __device__ static double logprior_h(const double* params, const DrjList& misc, const double* loglo, unsigned mask) {
    const double ls = (mask >> P_sigma & 1u) ? loglo[P_sigma] : log(params[P_sigma]);
    return -log(20.0) + R::dlnorm_l(params[P_sigma], ls, 0.0, 1.0, true);
  }
};
#define DRJ_USER_MODEL MyHintedModel
""")


@pytest.mark.parametrize("sigma_min", [0.0, 0.25])
@pytest.mark.parametrize("shape", [dict(rungs=4, chains=5), dict(rungs=1, chains=2)])
def test_nvrtc_user_model_with_hints(sigma_min, shape):
    """The HINTS hook through NVRTC: with sigma's lower bound at 0 the model reads log(sigma) from the sampler, with any
    other bound it takes the logarithm itself; either way bitwise the model without the hook (thread-per-particle and
    speculative kernels)."""
    x = np.random.default_rng(4).normal(3, 2, 100)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", sigma_min, "max", np.inf)
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=20, samples=25, seed=17, silent=True,
              as_frames=False, **shape)
    drj.source_cuda(HINTS_SRC)
    a = drj.run_mcmc(**kw)
    drj.source_cuda(DATUM_SRC)
    b = drj.run_mcmc(**kw)
    for k in a.raw:
        if k != "t_diff":
            assert np.array_equal(a.raw[k], b.raw[k]), k


POIS_SRC = r"""
// a Poisson regression y_j ~ Pois(exp(a + b t_j)) as a model struct; with -DUSE_INIT the lgamma(y_j + 1) terms come from
// the CTA_INIT hook of cuda_template.cu instead of being recomputed at every evaluation
struct PoisModel {
  static constexpr int D = DRJ_D, NBLOCK = DRJ_NBLOCK;
  static constexpr bool DATUM_FORM = false;
#ifdef USE_INIT
  static constexpr bool CTA_INIT = true;
  __device__ static double* consts() { __shared__ double c[64]; return c; }
  __device__ static void cta_init(const DrjList& data, const DrjList& misc) {
    for (int j = threadIdx.x; j < 64 && j < data.len(DATA_y); j += blockDim.x) consts()[j] = lgamma(data.item(DATA_y)[j] + 1.0);
  }
#endif
  __device__ static double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
    double ll = 0.0;
    for (int j = 0; j < (int)data.len(DATA_y); ++j) {
      const double eta = params[P_a] + params[P_b] * data.item(DATA_t)[j];
#ifdef USE_INIT
      const double lg = consts()[j];
#else
      const double lg = lgamma(data.item(DATA_y)[j] + 1.0);
#endif
      ll += data.item(DATA_y)[j] * eta - exp(eta) - lg;
    }
    return ll;
  }
  __device__ static double logprior(const double* params, const DrjList& misc) {
    return R::dnorm(params[P_a], 0.0, 10.0, true) + R::dnorm(params[P_b], 0.0, 10.0, true);
  }
};
#define DRJ_USER_MODEL PoisModel
"""


@pytest.mark.parametrize("shape", [dict(rungs=5, chains=7), dict(rungs=1, chains=3), dict(rungs=2, chains=400)])
def test_nvrtc_user_model_with_cta_init(shape):
    """The CTA_INIT hook through NVRTC on every kernel family a generic model can take (small-run, speculative, the
    general one-thread-per-particle kernel): the data-only terms formed once per block give the bits of the model that
    recomputes them."""
    rng = np.random.default_rng(8)
    t = np.linspace(0.0, 2.0, 40)
    y = rng.poisson(np.exp(0.5 + 0.8 * t)).astype(float)
    df = drj.define_params("name", "a", "min", -5, "max", 5, "name", "b", "min", -5, "max", 5)
    kw = dict(data={"y": y, "t": t}, df_params=df, loglike="loglike", logprior="logprior", burnin=20, samples=25, seed=3, silent=True,
              as_frames=False, **shape)
    drj.source_cuda("#define USE_INIT 1\n" + POIS_SRC)
    a = drj.run_mcmc(**kw)
    drj.source_cuda(POIS_SRC)
    b = drj.run_mcmc(**kw)
    assert np.isfinite(a.raw["loglike_sampling"]).all()
    for k in a.raw:
        if k != "t_diff":
            assert np.array_equal(a.raw[k], b.raw[k]), k


def test_nvrtc_errors_are_reported():
    df = drj.define_params("name", "mu", "min", -10, "max", 10)
    drj.source_cuda("__device__ double loglike(const double* p, const DrjList& d, const DrjList& m, int b) { return undefined_symbol; }\n"
                    "__device__ double logprior(const double* p, const DrjList& m) { return 0.0; }")
    with pytest.raises(drj.DrjacobyError) as e:
        drj.run_mcmc({"x": [1.0]}, df, loglike="loglike", logprior="logprior", burnin=10, samples=10, chains=1, silent=True)
    assert e.value.code == 12 and "undefined_symbol" in str(e.value)


def test_nvrtc_block_model_and_misc(monkeypatch):
    """Blocks + misc + several data items through the JIT path: the multilevel model written by hand
    (inst/extdata/checks/multilevel_loglike_logprior.cpp) against the built-in, both on the one-thread-per-particle
    kernels (the built-in declares a lane form and would otherwise run on the lanes kernel: equal within 1e-10, not
    bitwise -- tests/test_gpu_lanes.py)."""
    monkeypatch.setenv("DRJ_LANES", "0")
    G = 5
    rng = np.random.default_rng(5)
    data = {f"g{g + 1}": rng.normal(rng.normal(), 1.0, 5) for g in range(G)}
    src = r"""
    __device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
      const int G = (int)misc.scalar(MISC_n_groups);
      double ret = 0.0;
      if (block == G + 1) {
        for (int i = 0; i < G; ++i) ret += R::dnorm(params[i], 0.0, 1.0, true);
      } else {
        const double* x = data.item(block - 1);
        for (long long i = 0; i < data.len(block - 1); ++i) ret += R::dnorm(x[i], params[block - 1], 1.0, true);
      }
      return ret;
    }
    __device__ double logprior(const double* params, const DrjList& misc) { return 0.0; }
    """
    args = []
    for g in range(G):
        args += ["name", f"mu_{g + 1}", "min", -5, "max", 5, "block", [g + 1, G + 1]]
    df = drj.define_params(*args)
    kw = dict(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=60, samples=60, rungs=3, chains=4,
              seed=8, silent=True)
    drj.source_cuda(src)
    a = drj.run_mcmc(misc={"n_groups": G}, **kw)
    drj.use_builtin("multilevel")
    b = drj.run_mcmc(**kw)
    for k in a.raw:
        if k != "t_diff":
            assert np.array_equal(a.raw[k], b.raw[k]), k


def test_multilevel_any_size_specialised_with_nvrtc():
    """multilevel<G> for a G that is not pre-instantiated is specialised at run time; posterior of
    vignettes/checks_multilevel_blocks.Rmd: mu_i ~ N(m_i n/(n+1), 1/(n+1))."""
    G, n = 7, 5
    rng = np.random.default_rng(6)
    data = {f"g{g + 1}": rng.normal(rng.normal(), 1.0, n) for g in range(G)}
    args = []
    for g in range(G):
        args += ["name", f"mu_{g + 1}", "min", -5, "max", 5, "block", [g + 1, G + 1]]
    drj.use_builtin("multilevel")
    res = drj.run_mcmc(data, drj.define_params(*args), loglike="loglike", logprior="logprior", burnin=500, samples=2000,
                       chains=64, seed=3, silent=True, as_frames=False)
    th = res.raw["theta_sampling"][:, 0]  # [chains, samples, G]
    for g in range(G):
        m = np.mean(data[f"g{g + 1}"])
        draws = th[:, :, g].ravel()
        se = np.sqrt(1.0 / (n + 1) / (res.diagnostics["ess"][f"mu_{g + 1}"]))
        assert abs(draws.mean() - m * n / (n + 1)) < 6 * se + 1e-3
        assert abs(draws.var() - 1.0 / (n + 1)) < 0.02


def test_normal_posterior_matches_analytic():
    """vignettes/checks_normal_model.Rmd: posterior of mu is N(mean(x), 1/n); Rhat close to 1."""
    x = np.random.default_rng(10).normal(1.0, 1.0, 100)
    df = drj.define_params("name", "mu", "min", -np.inf, "max", np.inf, "init", 0.0)
    drj.use_builtin("normal_mu")
    res = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=1000, samples=5000, chains=128, seed=1,
                       silent=True, as_frames=False)
    draws = res.raw["theta_sampling"][:, 0, :, 0]
    ess = res.diagnostics["ess"]["mu"]
    assert abs(draws.mean() - x.mean()) < 6 * np.sqrt(0.01 / ess) + 1e-4
    assert abs(draws.var() - 0.01) < 3e-4
    assert abs(res.diagnostics["rhat"]["mu"] - 1.0) < 0.01
    # acceptance is tuned to the target by the Robbins-Monro bandwidth adaptation
    acc = res.raw["accept_sampling"][:, -1] / 5000.0
    assert abs(acc.mean() - 0.44) < 0.02


def test_return_prior_all_transform_types():
    """vignettes/checks_return_prior.Rmd: with a flat likelihood the four transform types return
    N(0,1), -Gamma(5,5), Gamma(5,5), Beta(3,3)."""
    df = drj.define_params("name", "real_line", "min", -np.inf, "max", np.inf, "name", "neg_line", "min", -np.inf, "max", 0,
                           "name", "pos_line", "min", 0, "max", np.inf, "name", "unit_interval", "min", 0, "max", 1)
    drj.use_builtin("returnprior")
    res = drj.run_mcmc({"x": [1.0]}, df, loglike="loglike", logprior="logprior", burnin=1000, samples=4000, chains=256, seed=5,
                       silent=True, as_frames=False)
    th = res.raw["theta_sampling"][:, 0, ::40, :].reshape(-1, 4)  # thinned: ~independent draws
    cdfs = [stats.norm(0, 1).cdf, lambda v: 1 - stats.gamma(5, scale=5).cdf(-v), stats.gamma(5, scale=5).cdf, stats.beta(3, 3).cdf]
    for k, cdf in enumerate(cdfs):
        assert stats.kstest(th[:, k], cdf).pvalue > 1e-4, k


def test_tempering_mixes_the_double_well():
    """vignettes/checks_double_well.Rmd: with 11 rungs (alpha = 2) the cold chain visits both wells."""
    df = drj.define_params("name", "mu", "min", -2, "max", 2, "init", -1.0, "name", "gamma", "min", 30, "max", 30, "init", 30)
    drj.use_builtin("doublewell")
    res = drj.run_mcmc({"x": [-1.0]}, df, loglike="loglike", logprior="logprior", burnin=1000, samples=5000, rungs=11, alpha=2,
                       chains=32, seed=7, silent=True, as_frames=False)
    mu = res.raw["theta_sampling"][:, 0, :, 0]
    frac_right = (mu > 0).mean(axis=1)
    assert (frac_right > 0.2).all() and (frac_right < 0.8).all()
    assert abs(np.abs(mu).mean() - 1.0) < 0.05
    one = drj.run_mcmc({"x": [-1.0]}, df, loglike="loglike", logprior="logprior", burnin=1000, samples=5000, rungs=1, chains=32,
                       seed=7, silent=True, as_frames=False)
    # without rungs a chain stays in whichever well it was in at the end of burn-in
    right = (one.raw["theta_sampling"][:, 0, :, 0] > 0).mean(axis=1)
    assert np.all((right < 0.01) | (right > 0.99)) and (right < 0.01).mean() > 0.7


def test_progress_callback_and_interrupt():
    df = drj.define_params("name", "mu", "min", -np.inf, "max", np.inf)
    drj.use_builtin("normal_mu")
    seen = []
    res = drj.run_mcmc({"x": [0.1, 0.2]}, df, loglike="loglike", logprior="logprior", burnin=200, samples=300, chains=2, seed=1,
                       silent=True, progress=lambda phase, done, total: seen.append((phase, done, total)) and False)
    assert seen[-1] == (1, 300, 300) and (0, 200, 200) in seen and len(seen) > 50
    with pytest.raises(drj.DrjacobyError) as e:  # Rcpp::checkUserInterrupt analogue
        drj.run_mcmc({"x": [0.1, 0.2]}, df, loglike="loglike", logprior="logprior", burnin=200, samples=300, chains=2, seed=1,
                     silent=True, progress=lambda phase, done, total: done > 50)
    assert e.value.code == 15


def test_summary_mode_matches_full_mode():
    """output="summary": Rhat / ESS / DIC from on-device reductions of the stored draws equal the host
    post-processing of the downloaded draws (R/main.R:498-554) -- same run, draws never leave the GPU."""
    x = np.random.default_rng(3).normal(3, 2, 40)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf,
                           "name", "fixed", "min", 2, "max", 2)
    src = drj.source_cuda(NORMAL_SRC)
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=300, samples=1500, rungs=4, chains=7,
              seed=21, silent=True)
    full = drj.run_mcmc(**kw)
    summ = drj.run_mcmc(output="summary", **kw)
    assert summ.output is None and summ.pt is None and "theta_sampling" not in summ.raw
    for k in ("mu", "sigma"):
        assert summ.diagnostics["rhat"][k] == full.diagnostics["rhat"][k]
        assert abs(summ.diagnostics["ess"][k] - full.diagnostics["ess"][k]) <= 1e-8 * full.diagnostics["ess"][k]
    assert np.isnan(summ.diagnostics["rhat"]["fixed"]) and np.isnan(summ.diagnostics["ess"]["fixed"])
    assert abs(summ.diagnostics["DIC_Gelman"] - full.diagnostics["DIC_Gelman"]) <= 1e-10 * abs(full.diagnostics["DIC_Gelman"])
    assert np.array_equal(summ.raw["accept_sampling"], full.raw["accept_sampling"])
    assert np.array_equal(summ.raw["mc_accept_sampling"], full.raw["mc_accept_sampling"])
    th = full.raw["theta_sampling"][:, 0]
    np.testing.assert_allclose(summ.raw["chain_mean"], th.mean(axis=1), rtol=1e-12)
    np.testing.assert_allclose(summ.raw["chain_var"][:, :2], th.var(axis=1, ddof=1)[:, :2], rtol=1e-10)
    pd_mc = summ.diagnostics["mc_accept"]
    assert pd_mc["value"].tolist() == full.diagnostics["mc_accept"]["value"].tolist()


def test_outputs_too_large_for_device_memory_is_an_error_not_a_crash():
    """the reference returns every draw; a run whose stored draws cannot fit in HBM is refused
    up front with advice, nothing is allocated."""
    df = drj.define_params("name", "mu", "min", -2, "max", 2, "name", "gamma", "min", 30, "max", 30)
    drj.use_builtin("doublewell")
    with pytest.raises(drj.DrjacobyError) as e:
        drj.run_mcmc({"x": [-1.0]}, df, loglike="loglike", logprior="logprior", burnin=10, samples=5e6, rungs=32, chains=4096,
                     silent=True, as_frames=False)
    assert e.value.code == 14 and "device memory" in str(e.value)


def test_two_sessions_interleaved_on_one_device():
    """Sessions own their stream and buffers: interleaving two of them changes nothing."""
    from drjacoby_b200 import Model, RunSpec, Session
    x = np.random.default_rng(8).normal(3, 2, 60)
    def spec(seed):
        return RunSpec(theta_min=[-10, 0], theta_max=[10, np.inf], theta_init=[[1.0, 1.0], [2.0, 2.0], [0.0, 3.0]], burnin=40,
                       samples=60, beta_raised=np.linspace(0, 1, 4), seed=seed, data=[x])
    m = Model(builtin="example")
    a, b = Session(m, spec(1)), Session(m, spec(2))
    for n in (7, 30, 1, 100):
        a.advance(n); b.advance(n)
    oa, ob = a.download(), b.download()
    a.close(); b.close()
    ra, rb = drj.run_raw(m, spec(1)), drj.run_raw(m, spec(2))
    for k in ra:
        if k != "t_diff":
            assert np.array_equal(oa[k], ra[k]) and np.array_equal(ob[k], rb[k]), k
    assert not np.array_equal(oa["theta_sampling"], ob["theta_sampling"])


def test_getting_model_fits_vignette_on_the_gpu():
    """docs/articles/getting_model_fits.html:247-266 (K3's model): six printed acceptance rates, from the GPU."""
    from helpers import POPULATION
    rs = drj.RStream(1)
    df = drj.define_params("name", "N_0", "min", 1, "max", 100, "name", "r", "min", 1, "max", 2, "name", "N_max", "min", 50, "max", 200)
    drj.use_builtin("logistic")
    res = drj.run_mcmc(data={"pop": POPULATION["pop"], "time": POPULATION["time"]}, df_params=df, misc={"output_type": 1},
                       loglike="loglike", logprior="logprior", burnin=1e3, samples=1e4, chains=3, rng="r", seed=rs, silent=True)
    got = [v for pair in rates_of(res, 1000, 10000, 3) for v in pair]
    assert got == GOLD["getting_model_fits"]["acceptance_pct_all_runs"]
    # the fitted curve plateaus near the carrying capacity the data suggest
    assert 90 < res.output[res.output["phase"] == "sampling"]["N_max"].median() < 115


def test_parallel_vignette_on_the_gpu():
    """docs/articles/parallel.html:181-192."""
    rs = drj.RStream(1)
    data = {"x": rs.rnorm(10)}
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "init", 0)
    drj.use_builtin("normal_mu")
    res = drj.run_mcmc(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=1e3, samples=1e3, chains=2, rng="r",
                       seed=rs, silent=True)
    got = [v for pair in rates_of(res, 1000, 1000, 1) for v in pair]
    assert got == GOLD["parallel"]["acceptance_pct_all_runs"]


TWO_BLOCK_DATUM_SRC = r"""
// two data vectors, two parameters, one likelihood block each; per-datum form with several blocks
// (the data is then always streamed through the tile pipeline, whatever its size)
struct TwoBlocks {
  static constexpr int D = DRJ_D, NBLOCK = DRJ_NBLOCK;
  static constexpr bool DATUM_FORM = true;
  struct Consts { double mean; };
  __device__ static int datum_item(int block) { return block - 1; }
  __device__ static bool prepare(const double* params, const DrjList& misc, int block, Consts& k) { k.mean = params[block - 1]; return true; }
  __device__ static double datum(const Consts& k, double x, double acc) { const double d = x - k.mean; return fma(d, d, acc); }
  __device__ static double finish(const Consts& k, double acc, long long n) { return -((double)n * 0.918938533204672741780329736406 + 0.5 * acc); }
  __device__ static double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
    double r = 0.0;
    for (long long i = 0; i < data.len(block - 1); ++i) r += R::dnorm(data.item(block - 1)[i], params[block - 1], 1.0, true);
    return r;
  }
  __device__ static double logprior(const double* params, const DrjList& misc) { return 0.0; }
};
#define DRJ_USER_MODEL TwoBlocks
"""

TWO_BLOCK_GENERIC_SRC = r"""
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  double r = 0.0;
  for (long long i = 0; i < data.len(block - 1); ++i) r += R::dnorm(data.item(block - 1)[i], params[block - 1], 1.0, true);
  return r;
}
__device__ double logprior(const double* params, const DrjList& misc) { return 0.0; }
"""


@pytest.mark.parametrize("n", [0, 37, 5000])
def test_multi_block_datum_form_model(n):
    """A per-datum model with two likelihood blocks over two data vectors (always streamed in tiles) against the
    same model in the generic form: same acceptances, values within the 1e-10 gate."""
    rng = np.random.default_rng(11)
    data = {"x": rng.normal(1.0, 1.0, n), "y": rng.normal(-2.0, 1.0, n + 3)}
    df = drj.define_params("name", "m1", "min", -10, "max", 10, "block", 1, "name", "m2", "min", -np.inf, "max", np.inf, "block", 2)
    kw = dict(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=30, samples=50, rungs=3, chains=6, seed=2,
              silent=True, as_frames=False)
    drj.source_cuda(TWO_BLOCK_DATUM_SRC)
    a = drj.run_mcmc(**kw)
    drj.source_cuda(TWO_BLOCK_GENERIC_SRC)
    b = drj.run_mcmc(**kw)
    for k in ("accept_burnin", "accept_sampling", "mc_accept_burnin", "mc_accept_sampling"):
        assert np.array_equal(a.raw[k], b.raw[k]), k
    for k in ("theta_sampling", "loglike_sampling", "theta_burnin", "loglike_burnin"):
        np.testing.assert_allclose(a.raw[k], b.raw[k], rtol=1e-10, atol=1e-12)
    if n > 0:
        assert abs(a.raw["theta_sampling"][:, 0, :, 0].mean() - data["x"].mean()) < 0.5


@pytest.mark.parametrize("chains,samples,max_lag,hot", [(300, 37, 2100, False), (3, 5000, 40, True), (1, 2049, 17, False)])
def test_on_device_autocovariance_edge_geometries(chains, samples, max_lag, hot):
    """drj_session_summary against numpy on the downloaded draws: chains shorter than a CTA (several chain
    boundaries inside one staging step), lags beyond one tile (the two-region staging path), hot draws stored
    (cold rung strided), a series ending one element into a tile."""
    from drjacoby_b200 import engine
    x = np.random.default_rng(5).normal(3, 2, 30)
    rungs = 3 if hot else 1
    init = np.tile(np.array([3.0, 2.0]), (chains, 1))
    spec = engine.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=init, burnin=20, samples=samples,
                          beta_raised=np.linspace(0, 1, rungs) if rungs > 1 else [1.0], save_hot_draws=hot, rng_mode="philox",
                          seed=11, data=[x])
    s = engine.Session(engine.Model(builtin="example"), spec)
    s.advance(spec.T)
    full = s.download()
    th = full["theta_sampling"][:, -1]                      # cold rung [chains, samples, d]
    series = th.reshape(chains * samples, 2)
    center = series.mean(axis=0) + np.array([0.01, -0.02])  # any centre
    sm = s.summary(max_lag=max_lag, center=center)
    only_ac = s.summary(max_lag=max_lag, center=center, moments=False)
    s.close()
    n = series.shape[0]
    for k in range(2):
        xc = series[:, k] - center[k]
        want = np.array([xc[: n - l] @ xc[l:] for l in range(max_lag + 1)])
        np.testing.assert_allclose(sm["autocov_sum"][k], want, rtol=1e-9, atol=1e-9 * abs(want[0]))
        assert np.array_equal(sm["autocov_sum"][k], only_ac["autocov_sum"][k])
        np.testing.assert_array_equal(sm["head"][k], series[:max_lag, k])
        np.testing.assert_array_equal(sm["tail"][k], series[n - max_lag:, k])
    np.testing.assert_allclose(sm["chain_mean"], th.mean(axis=1), rtol=1e-12)
    np.testing.assert_allclose(sm["chain_var"], th.var(axis=1, ddof=1), rtol=1e-9)
    assert not only_ac["chain_mean"].any()


@pytest.mark.parametrize("thin", [1, 7, 100, 5000])
def test_thinned_download_equals_slicing_the_full_download(thin):
    """drj_session_download_thinned: stored iterations 0, thin, 2 thin, ... of every series, gathered on the device."""
    from drjacoby_b200 import engine
    x = np.random.default_rng(5).normal(3, 2, 30)
    spec = engine.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=np.tile([3.0, 2.0], (6, 1)), burnin=45,
                          samples=301, beta_raised=np.linspace(0, 1, 3), save_hot_draws=True, rng_mode="philox", seed=4, data=[x])
    s = engine.Session(engine.Model(builtin="example"), spec)
    s.advance(spec.T)
    full, thinned = s.download(), s.download(thin=thin)
    s.close()
    for k in ("loglike_burnin", "logprior_burnin", "loglike_sampling", "logprior_sampling"):
        assert np.array_equal(thinned[k], full[k][:, :, ::thin]), k
    for k in ("theta_burnin", "theta_sampling"):
        assert np.array_equal(thinned[k], full[k][:, :, ::thin, :]), k
    for k in ("accept_burnin", "accept_sampling", "mc_accept_burnin", "mc_accept_sampling", "bw_final"):
        assert np.array_equal(thinned[k], full[k]), k


def test_summary_mode_with_thinned_draws():
    """run_mcmc(output="summary", thin=k): exact diagnostics from all draws on the device + every k-th draw as the
    usual output frame with the reference's iteration numbering."""
    x = np.random.default_rng(3).normal(3, 2, 40)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.source_cuda(NORMAL_SRC)
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=100, samples=1000, rungs=3, chains=4,
              seed=8, silent=True)
    full = drj.run_mcmc(**kw)
    thin = drj.run_mcmc(output="summary", thin=10, **kw)
    assert thin.diagnostics["rhat"]["mu"] == full.diagnostics["rhat"]["mu"]
    keep_pt = full.pt[((full.pt["iteration"] - 1) % 10 == 0)].reset_index(drop=True)   # the pt frame (all rungs), same thinning
    pd.testing.assert_frame_equal(thin.pt, keep_pt)
    assert abs(thin.diagnostics["ess"]["sigma"] - full.diagnostics["ess"]["sigma"]) <= 1e-8 * full.diagnostics["ess"]["sigma"]
    keep = full.output[((full.output["iteration"] - 1) % 10 == 0)].reset_index(drop=True)
    assert len(thin.output) == 4 * (10 + 100) and list(thin.output.columns) == list(full.output.columns)
    pd.testing.assert_frame_equal(thin.output, keep)
    with pytest.raises(ValueError):
        drj.run_mcmc(thin=10, **kw)


def test_memory_pool_reuses_and_releases_session_buffers():
    """Buffers of a destroyed session are reused by the next session of the same shape (same results, stale
    contents never visible) and can be handed back to the driver."""
    from drjacoby_b200 import engine
    engine.release_cached_memory()
    x = np.random.default_rng(5).normal(3, 2, 50)
    def run(seed, chains=4096):
        spec = engine.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=np.tile([3.0, 2.0], (chains, 1)), burnin=20,
                              samples=40, rng_mode="philox", seed=seed, data=[x])
        return engine.run_raw(engine.Model(builtin="example"), spec)
    a1 = run(1)
    b = run(2)           # same shape: reuses a1's buffers, different seed -> different contents
    a2 = run(1)          # and again: must reproduce a1 exactly
    for k in a1:
        if k != "t_diff":
            assert np.array_equal(a1[k], a2[k]), k
    assert not np.array_equal(a1["theta_sampling"], b["theta_sampling"])
    small = run(1, chains=8)   # a different shape next to cached blocks
    assert np.array_equal(small["theta_sampling"], a1["theta_sampling"][:8])
    freed = engine.release_cached_memory()
    assert freed >= a1["theta_sampling"].nbytes   # at least the sampling draws were cached
    assert engine.release_cached_memory() == 0


def test_readme_usage_example_runs():
    """The usage example of README.md, executed as written: posterior of the normal model around (3, 2)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "README.md")).read()
    code = re.search(r"## Using it like the reference\s+```python\n(.*?)```", text, flags=re.S).group(1)
    ns = {}
    exec(code, ns)
    m = ns["mcmc"]
    s = m.output[m.output.phase == "sampling"]
    assert len(s) == 5 * 10000 and abs(s["mu"].mean() - 3.0) < 0.6 and abs(s["sigma"].mean() - 2.0) < 0.6
    assert m.diagnostics["rhat"]["mu"] < 1.05 and m.diagnostics["ess"]["mu"] > 1000
