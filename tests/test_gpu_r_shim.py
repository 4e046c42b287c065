"""The .Call shim EXECUTED end to end on the GPU (replaces src/RcppExports.cpp:15-33 and the chain
dispatch of R/main.R:392-402,583-610).  There is no R in this image, so R is played by the functional mock of its
C API in tests/r_mock: tests/rshim_driver.py builds `args` as R/main.R:348-386 does (doubles where R has doubles,
the reserved misc$block, the named theta_min, init_mat d x chains), does the .Call by registered name, and walks the
per-chain 9-field lists as R/main.R:420-482 does.  Gate: bit for bit what the ctypes path returns."""
import numpy as np
import pytest

import rshim_driver as rd
from helpers import POPULATION, Case, beta_ladder, gpu_spec

pytestmark = pytest.mark.gpu

# inst/extdata/checks/multilevel_loglike_logprior.cpp rewritten against cuda_template.cu (5 groups)
MULTILEVEL_SRC = r"""
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  const int G = 5;
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


@pytest.fixture(scope="module")
def R(tmp_path_factory):
    r = rd.RMock(rd.build(str(tmp_path_factory.mktemp("rshim"))))
    yield r
    r.L.rmock_reset()


def args_for(R, case: Case, names, model, data_names, **kw):
    return rd.build_args(R, data=dict(zip(data_names, case.data)), misc={}, names=names, tmin=case.tmin, tmax=case.tmax,
                         blocks=case.blocks if case.blocks is not None else [[1]] * case.d, beta_raised=case.beta,
                         init_mat=case.init.T, burnin=case.burnin, samples=case.samples, chains=case.chains,
                         coupling_on=case.coupling_on, save_hot_draws=case.save_hot_draws, target_acceptance=case.target,
                         model=model, **kw)


def compare(raw_r: dict, raw_c: dict):
    for k, v in raw_r.items():
        if k != "t_diff":
            assert v.shape == raw_c[k].shape and np.array_equal(v, raw_c[k]), k
    assert np.all(raw_r["t_diff"] > 0)


def test_shim_full_output_equals_ctypes_path_bitwise(R):
    """K1-like: (mu, sigma), 5 chains x 4 rungs with hot draws kept; Philox seed; all visible GPUs (devices = NULL)."""
    import drjacoby_b200 as drj
    x = np.random.default_rng(2).normal(3, 2, 100)
    case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(4), chains=5, burnin=30, samples=50, save_hot_draws=True)
    out = R.dot_call("drjgpu_main", args_for(R, case, ["mu", "sigma"], {"builtin": "example"}, ["x"], seed=11))
    per_chain = R.to_py(out)
    assert len(per_chain) == 5 and list(per_chain[0]) == ["loglike_burnin", "logprior_burnin", "theta_burnin", "loglike_sampling",
                                                           "logprior_sampling", "theta_sampling", "mc_accept_burnin", "mc_accept_sampling", "t_diff"]
    assert len(per_chain[0]["theta_burnin"]) == 4 and len(per_chain[0]["theta_burnin"][0]) == 30 and per_chain[0]["theta_burnin"][0][0].shape == (2,)
    want = drj.run_raw(drj.Model(builtin="example"), gpu_spec(case, "philox", seed=11))
    compare(rd.chains_to_raw(per_chain, 2), want)
    assert R.L.rmock_protect_balance() == 0


def test_shim_replay_blocks_and_user_source(R):
    """The multilevel model as CUDA source (compiled by NVRTC from the text R hands over), blocks as a list of numeric
    vectors, five data items, replay of a given uniform stream (what reproduces a CPU run chain for chain)."""
    import drjacoby_b200 as drj
    rng = np.random.default_rng(7)
    G = 5
    data = [rng.normal(rng.normal(), 1.0, 5) for _ in range(G)]
    case = Case("multilevel", [-5.0] * G, [5.0] * G, data=data, blocks=[[g + 1, G + 1] for g in range(G)], beta=beta_ladder(3), chains=3,
                burnin=20, samples=20)
    u = case.replay()
    names = [f"mu_{g + 1}" for g in range(G)]
    dnames = [f"g{g + 1}" for g in range(G)]
    out = R.dot_call("drjgpu_main", args_for(R, case, names, {"source": MULTILEVEL_SRC, "loglike": "loglike", "logprior": "logprior"}, dnames,
                                              replay=u, devices=[0]))
    from drjacoby_b200 import _lib
    # (the built-in declares a lane form and would run on the lanes kernel, equal within 1e-10 only: pin it to the
    # one-thread-per-particle kernels the user's source runs on)
    want = drj.run_raw(drj.Model(builtin="multilevel"), gpu_spec(case, "replay", u, flags=_lib.FLAG_NO_LANES_KERNEL))
    got = rd.chains_to_raw(R.to_py(out), G)
    for k, v in got.items():
        if k != "t_diff":
            assert np.array_equal(v, want[k]), k  # (source vs built-in: the same arithmetic, hence the same bits)


def test_shim_summary_route(R):
    """drjgpu_summary: the draws stay on the GPU; moments / autocovariances / deviance / quantiles come back as R matrices,
    every 10th draw in the 9-field format.  Equal to the ctypes session path."""
    import drjacoby_b200 as drj
    x = np.random.default_rng(3).normal(3, 2, 60)
    case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(2), chains=6, burnin=40, samples=200)
    out = R.to_py(R.dot_call("drjgpu_summary", args_for(R, case, ["mu", "sigma"], {"builtin": "example"}, ["x"], seed=5, max_lag=20,
                                                         probs=[0.025, 0.5, 0.975], thin=10, devices=[0, 0])))
    s = drj.Session(drj.Model(builtin="example"), gpu_spec(case, "philox", seed=5), devices=[0, 0])
    s.advance(case.T)
    summ = s.summary(max_lag=20, quantiles=[0.025, 0.5, 0.975])
    thin = s.download(draws=True, thin=10)
    s.close()
    assert out["chain_mean"].shape == (2, 6) and np.array_equal(out["chain_mean"].T, summ["chain_mean"])
    assert np.array_equal(out["chain_var"].T, summ["chain_var"])
    assert out["autocov_sum"].shape == (21, 2) and np.array_equal(out["autocov_sum"].T, summ["autocov_sum"])
    assert np.array_equal(out["deviance_mean_var"], summ["deviance_mean_var"])
    assert out["quantiles"].shape == (3, 2) and np.array_equal(out["quantiles"].T, summ["quantiles"])
    assert np.array_equal(out["accept_sampling"].T, thin["accept_sampling"]) and np.array_equal(out["mc_accept_burnin"].T, thin["mc_accept_burnin"])
    got = rd.chains_to_raw(out["chains"], 2)
    assert got["theta_sampling"].shape == (6, 1, 20, 2)
    for k, v in got.items():
        if k != "t_diff":
            assert np.array_equal(v, thin[k]), k
    # without thin: no draws at all
    out2 = R.to_py(R.dot_call("drjgpu_summary", args_for(R, case, ["mu", "sigma"], {"builtin": "example"}, ["x"], seed=5, max_lag=20, devices=[0])))
    assert out2["chains"] is None and np.array_equal(out2["chain_mean"], out["chain_mean"]) and out2["quantiles"].size == 0


def test_shim_errors_and_interrupt(R):
    """Rcpp::stop() sites arrive as R errors with the reference's message; Esc (R_CheckUserInterrupt long-jumping inside
    R_ToplevelExec) stops the run cleanly between batches."""
    x = np.random.default_rng(4).normal(3, 2, 30)
    x_bad = x.copy()
    x_bad[3] = np.nan
    case = Case("example", [-10, 0], [10, np.inf], data=[x_bad], chains=2, burnin=20, samples=20)
    with pytest.raises(RuntimeError, match="NA, NaN or Inf in likelihood"):
        R.dot_call("drjgpu_main", args_for(R, case, ["mu", "sigma"], {"builtin": "example"}, ["x"], seed=1))
    init = np.full((3, 1), 0.2)
    init[1, 0] = 0.9
    half = Case("halfdomain_like", [0.0], [1.0], chains=3, burnin=10, samples=10, init=init)
    with pytest.raises(RuntimeError, match="Starting values result in -Inf"):
        R.dot_call("drjgpu_main", args_for(R, half, ["p"], {"builtin": "halfdomain_like"}, [], seed=1))
    ok = Case("example", [-10, 0], [10, np.inf], data=[x], chains=2, burnin=2000, samples=2000)
    R.L.rmock_interrupt_after(7)
    with pytest.raises(RuntimeError, match="interrupted"):
        R.dot_call("drjgpu_main", args_for(R, ok, ["mu", "sigma"], {"builtin": "example"}, ["x"], seed=1))
    assert R.L.rmock_protect_balance() == 0
    # and the next call works: nothing was left half-initialised
    out = R.dot_call("drjgpu_main", args_for(R, ok, ["mu", "sigma"], {"builtin": "example"}, ["x"], seed=1))
    assert len(R.to_py(out)) == 2


def test_shim_k3_shape_through_r(R):
    """BASELINE configs[2] in miniature through the shim: logistic growth, population_data, 3 bounded parameters, 10 rungs."""
    import drjacoby_b200 as drj
    case = Case("logistic", [1, 1, 50], [100, 2, 200], data=[POPULATION["pop"], POPULATION["time"]], beta=beta_ladder(10), chains=4,
                burnin=30, samples=30)
    out = R.dot_call("drjgpu_main", args_for(R, case, ["N_0", "r", "N_max"], {"builtin": "logistic"}, ["pop", "time"], seed=2))
    want = drj.run_raw(drj.Model(builtin="logistic"), gpu_spec(case, "philox", seed=2))
    compare(rd.chains_to_raw(R.to_py(out), 3), want)
