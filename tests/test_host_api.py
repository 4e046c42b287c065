"""Host-side logic of the run_mcmc mirror (no GPU): argument checks, pre-processing, R-compatible
stream generation, post-processing / diagnostics.  Where a sampler is needed the CPU oracle stands
in for the CUDA engine: the tests replace engine.run_raw (the product has no such seam)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

import drjacoby_b200 as drj
from drjacoby_b200 import api
from drjacoby_b200 import diagnostics as dg

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vignettes.json")))


def oracle_runner(po):
    def run(model, spec, progress=None):
        r = po.run(model.builtin, spec.data, spec.misc, spec.theta_min, spec.theta_max, spec.theta_init,
                   burnin=spec.burnin, samples=spec.samples, beta_raised=spec.beta_raised,
                   blocks=[list(spec.block_idx[spec.block_ptr[i]:spec.block_ptr[i + 1]]) for i in range(spec.d)],
                   coupling_on=spec.coupling_on, save_hot_draws=spec.save_hot_draws,
                   target_acceptance=spec.target_acceptance, rng_mode=spec.rng_mode, seed=spec.seed,
                   chain_offset=spec.chain_offset, replay=spec.replay)
        return {k: getattr(r, k) for k in ("loglike_burnin", "logprior_burnin", "theta_burnin", "loglike_sampling",
                                           "logprior_sampling", "theta_sampling", "mc_accept_burnin",
                                           "mc_accept_sampling", "accept_burnin", "accept_sampling", "t_diff", "bw_final")}
    return run


@pytest.fixture()
def cpu_engine(oracle, monkeypatch):
    monkeypatch.setattr(api.engine, "run_raw", oracle_runner(oracle))


# ---------------------------------------------------------------- define_params (R/main.R:48-139)
def test_define_params_forms():
    a = drj.define_params("name", "mu", "min", -10, "max", 10, "init", 0, "name", "sigma", "min", 0, "max", 5, "init", [1, 2])
    b = drj.define_params(dict(name="mu", min=-10, max=10, init=0), dict(name="sigma", min=0, max=5, init=[1, 2]))
    assert list(a["name"]) == ["mu", "sigma"] and list(a.columns) == ["name", "min", "max", "init"]
    assert np.array_equal(a["init"][1], [1.0, 2.0]) and a["min"].tolist() == b["min"].tolist()
    c = drj.define_params("name", "mu1", "min", -10, "max", 10, "init", 0, "block", 1,
                          "name", "sigma", "min", 0, "max", 5, "init", 1, "block", [1, 2])
    assert list(c["block"][1]) == [1, 2]


@pytest.mark.parametrize("args", [
    (), ("name", "mu", "min", 0), ("name", "mu", "min", 1, "max", 0), ("nome", "mu", "min", 0, "max", 1),
    ("name", "mu", "min", 0, "max", 1, "name", "mu", "min", 0, "max", 1),
    ("name", "mu", "min", 0, "max", 1, "init", 2), ("name", "mu", "min", 0, "max", 1, "init", -1),
    ("name", "a", "min", 0, "max", 1, "init", 0.5, "name", "b", "min", 0, "max", 1),
])
def test_define_params_errors(args):
    with pytest.raises((ValueError, TypeError)):
        drj.define_params(*args)


# ---------------------------------------------------------------- run_mcmc argument checks (R/main.R:236-283)
def test_run_mcmc_argument_checks():
    df = drj.define_params("name", "mu", "min", -10, "max", 10)
    drj.use_builtin("normal_mu")
    ok = dict(data={"x": [1.0]}, df_params=df, loglike="loglike", logprior="logprior", silent=True)
    bad = [dict(data=[1.0]), dict(misc=3), dict(misc={"block": 1}), dict(burnin=0), dict(samples=1.5), dict(rungs=-1),
           dict(chains=0), dict(alpha=0), dict(target_acceptance=1.5), dict(beta_manual=[0, 0.5]),
           dict(beta_manual=[0.5, 0.2, 1]), dict(beta_manual=[-1, 1]), dict(coupling_on="yes"), dict(cluster=5),
           dict(loglike=lambda p, d, m: 0.0), dict(logprior=3)]
    for kw in bad:
        with pytest.raises((ValueError, TypeError)):
            drj.run_mcmc(**{**ok, **kw})
    df2 = drj.define_params("name", "mu", "min", -10, "max", 10, "init", [-5, 5])
    with pytest.raises(ValueError):  # tests/testthat/test-mcmc-with-r-likelihood-and-prior.R:173: init length != chains
        drj.run_mcmc(**{**ok, "df_params": df2, "chains": 3})


def test_closures_are_rejected_not_emulated():
    df = drj.define_params("name", "mu", "min", -10, "max", 10)
    with pytest.raises(TypeError, match="no CPU fallback"):
        drj.run_mcmc({"x": [1.0]}, df, loglike=lambda params, data, misc: 0.0, logprior="logprior")


# ---------------------------------------------------------------- R-compatible stream
def test_rcompat_stream_matches_oracle(oracle):
    a, b = drj.RStream(1), oracle.RStream(1)
    assert np.array_equal(a.runif(1000), b.runif(1000))
    np.testing.assert_allclose(a.rnorm(500, 3, 2), b.rnorm(500, 3, 2), rtol=1e-15, atol=1e-15)
    assert np.array_equal(a.runif(700), b.runif(700))  # rnorm consumed exactly two uniforms per draw on both


# ---------------------------------------------------------------- diagnostics (R/diagnostic.R:17-38, coda)
def test_diagnostics_match_oracle_restatement():
    from oracle import diagnostics as og
    rng = np.random.default_rng(3)
    x = np.zeros((4, 3000))
    for c in range(4):
        for i in range(1, 3000):
            x[c, i] = 0.8 * x[c, i - 1] + rng.normal()
    assert dg.gelman_rubin(x) == og.gelman_rubin(x)
    assert abs(dg.ess(x.ravel()) - og.ess(x.ravel())) < 1e-8
    assert abs(dg.dic_gelman(x) - og.dic_gelman(x)) < 1e-9
    assert dg.ess(np.ones(100)) == 0.0
    for bad in (np.zeros((1, 10)), np.zeros((3, 1))):  # tests/testthat/test-diagnostic.R:2-13
        with pytest.raises(ValueError):
            dg.gelman_rubin(bad)


# ---------------------------------------------------------------- whole host path on the golden vignette
def test_vignette_through_the_host_wrapper(cpu_engine):
    """set.seed(1) -> data -> run_mcmc(rng='r'): pre-processing draws the default inits from the same
    stream, chains replay their slices, post-processing gives the printed head(), Rhat and ESS."""
    g = GOLD["example"]
    rs = drj.RStream(1)
    data = {"x": rs.rnorm(10, 3, 2)}
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.use_builtin("example")
    res = drj.run_mcmc(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=1e3, samples=1e3,
                       rng="r", seed=rs, silent=True)
    assert len(res) == 4 and list(res.output.columns) == ["chain", "phase", "iteration", "mu", "sigma", "logprior", "loglikelihood"]
    head = res.output.head(6)
    for row, want in zip(head.itertuples(index=False), g["head_rows_chain1"]["rows"]):
        assert row.chain == 1 and row.phase == "burnin" and row.iteration == want[0]
        for have, w in zip((row.mu, row.sigma, row.logprior, row.loglikelihood), want[1:]):
            digits = len(repr(w).split(".")[1])
            assert abs(have - w) <= 0.5000001 * 10.0 ** (-digits)
    assert [float(res.diagnostics["rhat"][k]) for k in ("mu", "sigma")] == g["rhat"]
    np.testing.assert_allclose([res.diagnostics["ess"][k] for k in ("mu", "sigma")], g["ess"], atol=5e-5)
    rates = [[api._printed_rate(res.raw["accept_burnin"][c, -1], 1000, 2), api._printed_rate(res.raw["accept_sampling"][c, -1], 1000, 2)]
             for c in range(5)]
    assert rates == g["acceptance_pct_run1"]
    # the second vignette run (C++ functions) continues the same stream
    res2 = drj.run_mcmc(data=data, df_params=df, loglike="loglike", logprior="logprior", burnin=1e3, samples=1e3,
                        rng="r", seed=rs, silent=True)
    rates2 = [[api._printed_rate(res2.raw["accept_burnin"][c, -1], 1000, 2), api._printed_rate(res2.raw["accept_sampling"][c, -1], 1000, 2)]
              for c in range(5)]
    assert rates2 == g["acceptance_pct_run2"]


def test_output_object_schema(cpu_engine):
    """columns, row order, rung labels, mc_accept layout, parameters record."""
    drj.use_builtin("example")
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    x = np.random.default_rng(1).normal(3, 2, 20)
    res = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=20, samples=30, rungs=3, chains=2,
                       save_hot_draws=True, seed=4, silent=True)
    out, pt = res.output, res.pt
    assert len(out) == 2 * 50 and out["chain"].tolist() == [1] * 50 + [2] * 50
    assert out["phase"].tolist()[:50] == ["burnin"] * 20 + ["sampling"] * 30 and out["iteration"].tolist()[:50] == list(range(1, 51))
    assert list(pt.columns) == ["chain", "rung", "phase", "iteration", "logprior", "loglikelihood", "mu", "sigma"]
    assert len(pt) == 2 * 3 * 50 and pt["rung"].tolist()[:150] == [1] * 50 + [2] * 50 + [3] * 50
    cold = pt[(pt["rung"] == 3)].reset_index(drop=True)
    assert np.array_equal(cold["mu"].to_numpy(), out["mu"].to_numpy())
    mc = res.diagnostics["mc_accept"]
    assert list(mc.columns) == ["link", "chain", "phase", "value"] and len(mc) == 2 * 2 * 2
    assert np.isclose(mc["value"].iloc[0], res.raw["mc_accept_burnin"][0, 0] / 20)
    assert res.diagnostics["rung_details"]["thermodynamic_power"].tolist() == [0.0, 0.5, 1.0]
    assert set(res.parameters) == {"data", "df_params", "loglike", "logprior", "burnin", "samples", "rungs", "chains",
                                   "coupling_on", "alpha", "beta_manual"}
    assert list(res.parameters["df_params"].columns) == ["name", "min", "max", "init", "block", "trans_type"]
    assert res.parameters["df_params"]["trans_type"].tolist() == [3, 2]
    dev = -2 * out[out["phase"] == "sampling"]["loglikelihood"].to_numpy()
    assert np.isclose(res.diagnostics["DIC_Gelman"], dev.mean() + 0.5 * dev.var(ddof=1))
    # R/samples.R:21-46: parameter columns only, `sample` appended last, rows seq.int(1, N, length.out = n) truncated
    s = drj.sample_chains(res, 7)
    assert list(s.columns) == ["mu", "sigma", "sample"] and len(s) == 7 and s["sample"].tolist() == list(range(1, 8))
    samp = out[out["phase"] == "sampling"].reset_index(drop=True)
    rows = np.floor(np.linspace(1, len(samp), 7)).astype(int) - 1   # 1, 10.83 -> 10, 20.67 -> 20, 30.5 -> 30, ... (R truncates)
    assert rows.tolist() == [0, 9, 19, 29, 39, 49, 59]
    assert np.array_equal(s["mu"].to_numpy(), samp["mu"].to_numpy()[rows])
    sk = drj.sample_chains(res, 7, keep_chain_index=True)
    assert list(sk.columns) == ["chain", "mu", "sigma", "sample"] and sk["chain"].tolist() == [1, 1, 1, 1, 2, 2, 2]
    with pytest.raises(ValueError):
        drj.sample_chains(res, 1000)
    # single rung: mc_accept is NA, no rhat for a single chain
    res1 = drj.run_mcmc({"x": x}, df, loglike="loglike", logprior="logprior", burnin=10, samples=10, chains=1, seed=4, silent=True)
    assert isinstance(res1.diagnostics["mc_accept"], float) and "rhat" not in res1.diagnostics


def test_fixed_parameter_and_user_init(cpu_engine):
    """min == max parameters are skipped (NA diagnostics) and the first stored row is the user's init
    (tests/testthat/test-mcmc-with-r-likelihood-and-prior.R:130-221)."""
    drj.use_builtin("doublewell")
    df = drj.define_params("name", "mu", "min", -2, "max", 2, "init", [-1.5, 0.5], "name", "gamma", "min", 30, "max", 30, "init", 30)
    res = drj.run_mcmc({"x": [-1.0]}, df, loglike="loglike", logprior="logprior", burnin=50, samples=50, chains=2, seed=2, silent=True)
    first = res.output[res.output["iteration"] == 1]
    assert first["mu"].tolist() == [-1.5, 0.5] and (res.output["gamma"] == 30).all()
    assert np.isnan(res.diagnostics["rhat"]["gamma"]) and np.isnan(res.diagnostics["ess"]["gamma"])
    assert res.parameters["df_params"]["trans_type"].tolist() == [3, 3]


def test_default_inits_follow_the_reference_rules(cpu_engine):
    """R/main.R:299-314 for the four transform types."""
    drj.use_builtin("returnprior")
    df = drj.define_params("name", "real_line", "min", -np.inf, "max", np.inf, "name", "neg_line", "min", -np.inf, "max", 0,
                           "name", "pos_line", "min", 0, "max", np.inf, "name", "unit_interval", "min", 0, "max", 1)
    rs = drj.RStream(5)
    res = drj.run_mcmc({"x": [1.0]}, df, loglike="loglike", logprior="logprior", burnin=5, samples=5, chains=3, rng="r",
                       seed=rs, silent=True)
    p = drj.RStream(5).runif(12).reshape(4, 3)
    want = np.stack([np.log(p[0]) - np.log(1 - p[0]), np.log(p[1]) + 0.0, 0.0 - np.log(p[2]), p[3]], axis=1)
    np.testing.assert_allclose(res.raw["theta_burnin"][:, 0, 0, :], want, rtol=1e-15)
    assert res.parameters["df_params"]["trans_type"].tolist() == [0, 1, 2, 3]


def test_shard_ranges():
    for chains in (1, 5, 8, 100000):
        for world in (1, 2, 3, 8):
            r = [api.shard_range(chains, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == chains and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_thinned_frame_numbering_and_shapes():
    """Host side of run_mcmc(output="summary", thin=k): output-buffer shapes for a thinned download and the
    iteration numbers of the frame built from it (the reference numbers 1..burnin, burnin+1..burnin+samples)."""
    from drjacoby_b200 import engine
    spec = engine.RunSpec(theta_min=[-1.0, 0.0], theta_max=[1.0, 5.0], theta_init=np.zeros((3, 2)) + 0.5, burnin=10, samples=25,
                          beta_raised=[0.0, 1.0], save_hot_draws=True)
    o = spec.alloc_outputs(thin=4)
    assert o["theta_burnin"].shape == (3, 2, 3, 2) and o["theta_sampling"].shape == (3, 2, 7, 2)
    assert o["loglike_sampling"].shape == (3, 2, 7) and o["accept_sampling"].shape == (3, 2)
    assert "theta_sampling" not in spec.alloc_outputs(draws=False)
    raw = {k: np.arange(v.size, dtype=float).reshape(v.shape) for k, v in o.items()}
    f, pt = api._thinned_frames(raw, ["a", "b"], 10, 25, 4, 3, 2, True)
    # the `pt` frame of the reference (R/main.R:451-482): every stored rung, same thinned iterations, hot draws included
    assert list(pt.columns) == ["chain", "rung", "phase", "iteration", "logprior", "loglikelihood", "a", "b"]
    assert len(pt) == 3 * 2 * 10 and pt["rung"].tolist()[:20] == [1] * 10 + [2] * 10
    hot = pt[(pt["chain"] == 2) & (pt["rung"] == 1)]
    assert hot["iteration"].tolist() == [1, 5, 9, 11, 15, 19, 23, 27, 31, 35]
    assert hot["loglikelihood"].tolist()[:3] == raw["loglike_burnin"][1, 0].tolist()
    assert hot["b"].tolist()[3:] == raw["theta_sampling"][1, 0, :, 1].tolist()
    one = f[f["chain"] == 2]
    assert one["iteration"].tolist() == [1, 5, 9, 11, 15, 19, 23, 27, 31, 35]
    assert one["phase"].tolist() == ["burnin"] * 3 + ["sampling"] * 7
    assert one["a"].tolist()[:3] == raw["theta_burnin"][1, -1, :, 0].tolist()          # cold rung = last
    assert one["loglikelihood"].tolist()[3:] == raw["loglike_sampling"][1, -1].tolist()


def test_thin_argument_checks():
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.use_builtin("example")
    kw = dict(data={"x": np.ones(5)}, df_params=df, loglike="loglike", logprior="logprior", burnin=5, samples=5, silent=True)
    with pytest.raises(ValueError):
        drj.run_mcmc(thin=2, **kw)                       # needs output="summary"
    with pytest.raises(ValueError):
        drj.run_mcmc(thin=0, output="summary", **kw)
    with pytest.raises(ValueError):
        drj.run_mcmc(thin=2.5, output="summary", **kw)
