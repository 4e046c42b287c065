"""One process, several GPUs through the C ABI (drj_run_mcmc_multi / drj_multi_*): what replaces the reference's
dispatch of chains over a PSOCK cluster (R/main.R:392-402) and what the .Call shim binds.

The chains are split into contiguous ranges inside the library, one host thread per device, every shard writing
straight into its slice of the caller's buffers.  Gate: the result is BIT FOR BIT the one-device result, whatever
the number of devices -- for every kernel family -- and the merged summaries (per-chain moments, autocovariance
sums incl. the lag products that straddle two shards, deviance moments, quantiles) equal the one-device ones.
On a one-GPU box the same device is listed several times (its chain ranges then run side by side on it), so
the sharding, the threading and the merge are exercised everywhere; with >= 2 GPUs the real thing runs too.
"""
import numpy as np
import pytest

from helpers import Case, beta_ladder, gpu_spec

pytestmark = pytest.mark.gpu


def normal_data(n, seed=3, mean=3.0, sd=2.0):
    return np.random.Generator(np.random.Philox(seed)).normal(mean, sd, n)


def device_lists():
    import drjacoby_b200 as drj
    n = drj.device_count()
    lists = [[0, 0, 0], [0] * 8]
    if n >= 2:
        lists += [list(range(n)), None]
    return lists


CASES = {
    # one thread per particle (13 chains x 4 rungs, uneven shards)
    "solo": Case("example", [-10, 0], [10, np.inf], data=[normal_data(200)], beta=beta_ladder(4), chains=13, burnin=40, samples=60),
    # cluster per chain
    "team": Case("example", [-10, 0], [10, np.inf], data=[normal_data(30001)], beta=beta_ladder(3), chains=5, burnin=10, samples=20,
                 save_hot_draws=True),
}


def same_bits(a, b, what=""):
    assert set(a) == set(b)
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), (what, k)


@pytest.mark.parametrize("name", ["solo", "team"])
def test_multi_equals_single_bitwise(name):
    import drjacoby_b200 as drj
    case = CASES[name]
    m = drj.Model(builtin=case.model)
    one = drj.run_raw(m, gpu_spec(case, "philox", seed=5, chain_offset=2))
    for devs in device_lists():
        got = drj.engine.run_raw_multi(m, gpu_spec(case, "philox", seed=5, chain_offset=2), devs)
        same_bits(one, got, (name, devs))
        assert np.all(got["t_diff"] > 0) and np.all(got["t_diff"] == got["t_diff"][0])


def test_multi_grid_kernels_bitwise():
    """The grid-per-run family: the slot layout follows the WHOLE run's particle count (layout_chains is set by the
    multi-device entry), so 2-chain and 1-chain shards reproduce the 5-chain run."""
    import drjacoby_b200 as drj
    from drjacoby_b200 import _lib
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(50001)], beta=beta_ladder(2), chains=5, burnin=10, samples=10)
    m = drj.Model(builtin="example")
    one = drj.run_raw(m, gpu_spec(case, "philox", seed=9, flags=_lib.FLAG_GRID_KERNEL))
    for devs in device_lists():
        got = drj.engine.run_raw_multi(m, gpu_spec(case, "philox", seed=9, flags=_lib.FLAG_GRID_KERNEL), devs)
        same_bits(one, got, devs)


def test_multi_replay_mode_and_more_devices_than_chains(oracle):
    """Replay mode: every shard gets its slice of the uniform stream; 8 device entries for 3 chains: 3 shards."""
    import drjacoby_b200 as drj
    from helpers import assert_parity, run_oracle
    case = Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(50, 4, 10.0, 1.0)], beta=beta_ladder(5), chains=3,
                burnin=30, samples=30)
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    m = drj.Model(builtin="coupling")
    got = drj.engine.run_raw_multi(m, gpu_spec(case, "replay", u), [0] * 8)
    assert_parity(got, ora, 1e-10, what="[multi/replay] ")
    s = drj.Session(m, gpu_spec(case, "replay", u), devices=[0] * 8)
    assert s.n_devices() == 3
    s.close()


def test_multi_user_model_is_loaded_on_every_device():
    """A source model compiled once (NVRTC) serves every device: its cubin is loaded into each context on first use."""
    import drjacoby_b200 as drj
    from test_gpu_api import NORMAL_SRC
    case = Case("example_exact", [-10, 0], [10, np.inf], data=[normal_data(64)], beta=beta_ladder(2), chains=6, burnin=20, samples=20)
    jit = drj.Model(source=NORMAL_SRC)
    one = drj.run_raw(drj.Model(builtin="example_exact"), gpu_spec(case, "philox", seed=3))
    for devs in device_lists():
        spec = gpu_spec(case, "philox", seed=3, param_names=["mu", "sigma"], data_names=["x"])
        same_bits(one, drj.engine.run_raw_multi(jit, spec, devs), devs)


def test_multi_error_is_the_reference_error_of_the_first_failing_chain():
    """-Inf at the initial values of chains 4 and 9: the reference stops at the first of them (chains run in order)."""
    import drjacoby_b200 as drj
    init = np.full((12, 1), 0.2)
    init[[4, 9], 0] = 0.9
    case = Case("halfdomain_like", [0.0], [1.0], chains=12, burnin=10, samples=10, init=init)
    with pytest.raises(drj.DrjacobyError) as e:
        drj.engine.run_raw_multi(drj.Model(builtin="halfdomain_like"), gpu_spec(case, "philox", seed=1), [0, 0, 0])
    assert e.value.code == 1 and e.value.chain == 4 and "Starting values" in str(e.value)


def test_multi_progress_runs_on_the_calling_thread_and_can_interrupt():
    import threading
    import drjacoby_b200 as drj
    case = CASES["solo"]
    seen, me = [], threading.get_ident()

    def cb(user, phase, done, total):
        seen.append((phase, done, total, threading.get_ident()))
        return 1 if (phase == 1 and done >= 30) else 0

    with pytest.raises(drj.DrjacobyError) as e:
        drj.engine.run_raw_multi(drj.Model(builtin="example"), gpu_spec(case, "philox", seed=5), [0, 0], cb)
    assert e.value.code == 15
    assert all(t == me for *_, t in seen) and seen[0][0] == 0 and seen[-1][0] == 1 and seen[-1][1] >= 30


def test_multi_summary_equals_single_summary():
    """drj_multi_summary: moments and deviance bit for bit, autocovariance sums to rounding (the sums are added over
    the shards plus the straddling lag products), quantiles exactly; and the quantiles equal numpy's on the draws."""
    import drjacoby_b200 as drj
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], beta=beta_ladder(2), chains=7, burnin=50, samples=333)
    m = drj.Model(builtin="example")
    probs = [0.025, 0.25, 0.5, 0.975, 1.0, 0.0]
    s1 = drj.Session(m, gpu_spec(case, "philox", seed=12))
    s1.advance(case.T)
    a = s1.summary(quantiles=probs)
    raw = s1.download()
    s1.close()
    draws = raw["theta_sampling"][:, -1].reshape(-1, 2)
    np.testing.assert_allclose(a["quantiles"], np.quantile(draws, probs, axis=0).T, rtol=1e-14)
    for devs in device_lists():
        s2 = drj.Session(m, gpu_spec(case, "philox", seed=12), devices=devs if devs is not None else "all")
        s2.advance(case.T)
        b = s2.summary(quantiles=probs)
        raw2 = s2.download()
        s2.close()
        same_bits(raw, raw2, devs)
        for k in ("chain_mean", "chain_var", "deviance_mean_var", "quantiles", "head", "tail"):
            assert np.array_equal(a[k], b[k]), (devs, k)
        np.testing.assert_allclose(b["autocov_sum"], a["autocov_sum"], rtol=1e-12, atol=1e-9)


def test_multi_summary_shards_shorter_than_the_lag_window():
    """Shards of one chain x 5 draws with max_lag = 12: lag products reach across several shard boundaries."""
    import drjacoby_b200 as drj
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(30)], chains=6, burnin=5, samples=5)
    m = drj.Model(builtin="example")
    s1 = drj.Session(m, gpu_spec(case, "philox", seed=2))
    s1.advance(case.T)
    a = s1.summary(max_lag=12)
    draws = s1.download()["theta_sampling"][:, -1].reshape(-1, 2)
    s1.close()
    s2 = drj.Session(m, gpu_spec(case, "philox", seed=2), devices=[0] * 6)
    s2.advance(case.T)
    b = s2.summary(max_lag=12)
    s2.close()
    ctr = draws.reshape(6, 5, 2).mean(axis=1).mean(axis=0)
    z = draws - ctr
    want = np.array([[np.sum(z[:len(z) - l, k] * z[l:, k]) for l in range(13)] for k in range(2)])
    np.testing.assert_allclose(a["autocov_sum"], want, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(b["autocov_sum"], want, rtol=1e-11, atol=1e-12)


def test_run_mcmc_cluster_gpus_equals_one_gpu():
    """The host mirror: run_mcmc(cluster="gpus") / cluster=[0, 0, 0] -- full output bit for bit, summary output equal."""
    import drjacoby_b200 as drj
    x = np.random.default_rng(1).normal(3, 2, 200)
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.use_builtin("example")
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=200, samples=800, rungs=4, chains=13, seed=5, silent=True)
    one = drj.run_mcmc(**kw)
    for cluster in ([0, 0, 0], "gpus"):
        full = drj.run_mcmc(cluster=cluster, **kw)
        same_bits(one.raw, full.raw, cluster)
        assert full.diagnostics["rhat"].equals(one.diagnostics["rhat"]) and full.output.equals(one.output)
        summ = drj.run_mcmc(cluster=cluster, output="summary", thin=10, quantiles=[0.025, 0.5, 0.975], **kw)
        for k in ("mu", "sigma"):
            assert summ.diagnostics["rhat"][k] == one.diagnostics["rhat"][k]
            assert abs(summ.diagnostics["ess"][k] - one.diagnostics["ess"][k]) < 1e-9 * one.diagnostics["ess"][k]
        assert abs(summ.diagnostics["DIC_Gelman"] - one.diagnostics["DIC_Gelman"]) < 1e-12 * abs(one.diagnostics["DIC_Gelman"])
        samp = one.output[one.output["phase"] == "sampling"]
        q = summ.diagnostics["quantiles"]
        assert list(q.index) == ["2.5%", "50%", "97.5%"] and list(q.columns) == ["mu", "sigma"]
        np.testing.assert_allclose(q["mu"].to_numpy(), np.quantile(samp["mu"].to_numpy(), [0.025, 0.5, 0.975]), rtol=1e-14)
        # the thinned `output` and `pt` frames are rows of the full ones
        keep = one.output[(one.output["iteration"] - 1) % 10 == 0]
        keep = keep[(keep["phase"] == "burnin") | ((keep["iteration"] - 201) % 10 == 0)].reset_index(drop=True)
        assert summ.output.equals(keep)
        assert len(summ.pt) == 13 * 4 * 100 and set(summ.pt.columns) == set(one.pt.columns)
