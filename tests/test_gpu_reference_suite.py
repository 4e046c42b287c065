"""The reference's own testthat files for the sampling path, restated against this package's interface
(same workloads, same assertions) and run on the GPU:
  tests/testthat/test-mcmc-with-cpp-likelihood-and-prior.R:3-99
  tests/testthat/test_mcmc_with_multi_function_cpp.R:1-36
  tests/testthat/test-mcmc-wth-mixed-likelihood-and-prior.R:3-48   (closures: rejected here, by design)
  tests/testthat/test-cpp_template.R
The model sources are CUDA translations of tests/testthat/test_input_files/*.cpp."""
import numpy as np
import pytest

import drjacoby_b200 as drj

pytestmark = pytest.mark.gpu

# tests/testthat/test_input_files/loglike_logprior.cpp: four functions, chosen by name in run_mcmc()
LOGLIKE_LOGPRIOR_CU = r"""
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  const double* x = data.item(DATA_x);
  double mu = params[P_mu], sigma = params[P_sigma];
  double ret = 0.0;
  for (long long i = 0; i < data.len(DATA_x); ++i) ret += R::dnorm(x[i], mu, sigma, 1);
  if (!isfinite(ret)) ret = -(DRJ_DBL_MAX / 100.0);   // catch underflow
  return ret;
}
__device__ double logprior(const double* params, const DrjList& misc) {
  double ret = R::dnorm(params[P_mu], 6, 0.1, 1) + R::dnorm(params[P_sigma], 1, 0.1, 1);
  if (!isfinite(ret)) ret = -(DRJ_DBL_MAX / 100.0);
  return ret;
}
__device__ double loglikenull(const double* params, const DrjList& data, const DrjList& misc, int block) { return 0.0; }
__device__ double logpriornull(const double* params, const DrjList& misc) { return 0.0; }
"""

# tests/testthat/test_input_files/multi_loglike_logprior.cpp: the likelihood calls a helper function
MULTI_CU = r"""
__device__ double none() { double x = 0.0; return x; }
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  const double* x = data.item(DATA_x);
  double ret = 0.0;
  for (long long i = 0; i < data.len(DATA_x); ++i) ret += R::dnorm(x[i], params[P_mu], params[P_sigma], 1) + none();
  if (!isfinite(ret)) ret = -(DRJ_DBL_MAX / 100.0);
  return ret;
}
__device__ double logprior(const double* params, const DrjList& misc) { return none(); }
"""


@pytest.fixture(scope="module")
def setup():
    rs = drj.RStream(1)                      # set.seed(1)
    x = rs.rnorm(1000, 3, 2)                 # rnorm(1000, mean = mu_true, sd = sigma_true)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "init", 5, "name", "sigma", "min", 0, "max", 10, "init", 1)
    return {"x": x}, df


def medians(res, chain=None):
    out = res.output[res.output["phase"] == "sampling"]
    if chain is not None:
        out = out[out["chain"] == chain]
    return out[["mu", "sigma"]].median()


def test_cpp_likelihood_and_prior(setup):
    data_list, df_params = setup
    drj.source_cuda(LOGLIKE_LOGPRIOR_CU)     # Rcpp::sourceCpp("test_input_files/loglike_logprior.cpp")
    null = drj.run_mcmc(data=data_list, df_params=df_params, loglike="loglikenull", logprior="logprior", burnin=1e3,
                        samples=1e3, chains=1, silent=True, seed=1)
    pe = medians(null)
    assert pe["mu"] - 6 < 0.1 and pe["sigma"] - 1 < 0.1          # expect_lt, as in the reference
    assert abs(pe["mu"] - 6) < 0.1 and abs(pe["sigma"] - 1) < 0.1  # and two-sided

    data = drj.run_mcmc(data=data_list, df_params=df_params, loglike="loglike", logprior="logpriornull", burnin=1e3,
                        samples=1e4, silent=True, seed=2)
    pe2 = medians(data)
    assert abs(pe2["mu"] - 3) < 0.1 and abs(pe2["sigma"] - 2) < 0.1
    assert data.parameters["chains"] == 5                           # default chains = 5

    chains2 = drj.run_mcmc(data=data_list, df_params=df_params, loglike="loglike", logprior="logpriornull", chains=2,
                           burnin=1e3, samples=1e4, silent=True, seed=3)
    assert len(chains2) == 4                                        # expect_length(cpp_mcmc_chains, 4)
    for c in (1, 2):
        pe3 = medians(chains2, c)
        assert abs(pe3["mu"] - 3) < 0.1 and abs(pe3["sigma"] - 2) < 0.1

    mc = drj.run_mcmc(data=data_list, df_params=df_params, loglike="loglike", logprior="logpriornull", burnin=1e3,
                      samples=1e4, rungs=4, silent=True, seed=4)
    pe4 = medians(mc)
    assert abs(pe4["mu"] - 3) < 0.1 and abs(pe4["sigma"] - 2) < 0.1
    assert set(mc.output["chain"]) == {1, 2, 3, 4, 5} and mc.pt["rung"].max() == 4


def test_multi_function_cpp_likelihood_and_prior(setup):
    data_list, df_params = setup
    drj.source_cuda(MULTI_CU)
    mcmc = drj.run_mcmc(data=data_list, df_params=df_params, loglike="loglike", logprior="logprior", burnin=1e3,
                        samples=1e3, chains=1, silent=True, seed=5)
    pe = medians(mcmc)
    assert abs(pe["mu"] - 3) < 0.1 and abs(pe["sigma"] - 2) < 0.1


def test_mixed_closure_and_device_functions_are_rejected(setup):
    """The reference accepts an R closure for either function; the GPU path has no CPU fallback and says so."""
    data_list, df_params = setup
    drj.source_cuda(LOGLIKE_LOGPRIOR_CU)
    for kw in (dict(loglike="loglikenull", logprior=lambda params, misc: 0.0),
               dict(loglike=lambda params, data, misc: 0.0, logprior="logpriornull")):
        with pytest.raises(TypeError, match="no CPU fallback"):
            drj.run_mcmc(data=data_list, df_params=df_params, burnin=1e2, samples=1e2, chains=1, silent=True, **kw)


def test_template_works(tmp_path):
    for bad in (1, "foo", "foo.csv"):                               # test-cpp_template.R
        with pytest.raises((TypeError, ValueError)):
            drj.cuda_template(bad)
    path = str(tmp_path / "model.cu")
    text = drj.cuda_template(path)
    assert open(path).read() == text
    # the untouched template is a valid model: flat likelihood, flat prior -> the chain just diffuses
    drj.source_cuda(path)
    df = drj.define_params("name", "a", "min", 0, "max", 1)
    res = drj.run_mcmc(data={"x": [1.0]}, df_params=df, loglike="loglike", logprior="logprior", burnin=200, samples=2000,
                       chains=8, silent=True, seed=1)
    a = res.output[res.output["phase"] == "sampling"]["a"]
    assert (res.output["loglikelihood"] == 0).all() and abs(a.mean() - 0.5) < 0.05 and 0 < a.min() and a.max() < 1
