"""The cluster-per-chain (TEAM) kernels: few chains over a long data vector -- the reference's usual shape
(1-10 chains, R/main.R:213-216 defaults) -- where one thread per (chain, rung) cannot use the GPU.

Every chain gets a thread-block cluster, all threads replicate the chain's particle states, and only the
per-datum likelihood sum is divided over the cluster (drj_sampler.cuh, drj_sweep_item_team).  Gates: the
same replay / Philox parity with the CPU oracle as the other kernels; results bitwise independent of the
cluster size, of the launch split and of how chains are sharded; agreement with the one-thread-per-particle
kernels within the 1e-10 gate (the sums are associated differently).
"""
import os

import numpy as np
import pytest

from helpers import Case, assert_parity, beta_ladder, gpu_spec, rel_close, run_gpu, run_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def normal_data(n, seed=3, mean=3.0, sd=2.0):
    return np.random.Generator(np.random.Philox(seed)).normal(mean, sd, n)


def flags():
    from drjacoby_b200 import _lib
    return _lib


TEAM_CASES = {
    # one chain, one rung, 35 tiles: 16 CTAs, 256 slots for the single rung
    "one_chain": Case("example", [-10, 0], [10, np.inf], data=[normal_data(70001)], chains=1, burnin=20, samples=20),
    # odd length, ragged last tile, 2 rungs
    "stream_r2": Case("example", [-10, 0], [10, np.inf], data=[normal_data(10001)], beta=beta_ladder(2), chains=3,
                      burnin=6, samples=6),
    # 20 rungs: 12 slots per rung, 16 left-over threads per CTA; coupling through slot 0
    "coupling_r20": Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(9000, 4, 10.0, 1.0)],
                         beta=beta_ladder(20), chains=2, burnin=15, samples=15),
    # 3 rungs: 85 slots, one left-over thread
    "rungs3_hot": Case("example", [-10, 0], [10, np.inf], data=[normal_data(6001)], beta=beta_ladder(3, 2.0), chains=5,
                       burnin=10, samples=10, save_hot_draws=True),
    # 32 rungs, fewer tiles (1) than virtual ranks
    "rungs32": Case("example", [-10, 0], [10, np.inf], data=[normal_data(5000)], beta=beta_ladder(32), chains=2,
                    burnin=8, samples=8),
    # more rungs than a warp: 5 slots of 50 rungs, 6 left-over threads
    "rungs50": Case("example", [-10, 0], [10, np.inf], data=[normal_data(20011)], beta=beta_ladder(50, 3.0), chains=2,
                    burnin=6, samples=6),
    "normal_mu": Case("normal_mu", [-np.inf], [np.inf], data=[normal_data(33333)], chains=4, burnin=10, samples=10),
}


@pytest.mark.parametrize("name", sorted(TEAM_CASES))
def test_team_replay_matches_oracle(name, oracle):
    case = TEAM_CASES[name]
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    gpu = run_gpu(case, "replay", u, flags=flags().FLAG_TEAM_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[team/{name}] ")


@pytest.mark.parametrize("name", ["one_chain", "coupling_r20"])
def test_team_philox_matches_oracle(name, oracle):
    case = TEAM_CASES[name]
    ora = run_oracle(oracle, case, "philox", seed=77, chain_offset=3)
    gpu = run_gpu(case, "philox", seed=77, chain_offset=3, flags=flags().FLAG_TEAM_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[team/{name}/philox] ")


def test_team_is_the_default_for_few_chains_and_long_data():
    """No flag: 3 chains x 2 rungs over 10001 data points selects the cluster kernels (same bits as forcing them),
    and DRJ_FLAG_NO_TEAM_KERNEL gives the one-thread-per-particle result, equal within the parity gate."""
    case = TEAM_CASES["stream_r2"]
    auto = run_gpu(case, "philox", seed=4)
    team = run_gpu(case, "philox", seed=4, flags=flags().FLAG_TEAM_KERNEL)
    solo = run_gpu(case, "philox", seed=4, flags=flags().FLAG_NO_TEAM_KERNEL)
    differs = False
    for k in auto:
        if k == "t_diff":
            continue
        assert np.array_equal(auto[k], team[k]), k
        if auto[k].dtype.kind == "f":
            ok, worst = rel_close(auto[k], solo[k], rtol=RTOL)
            assert ok, (k, worst)
            differs = differs or not np.array_equal(auto[k], solo[k])
        else:
            assert np.array_equal(auto[k], solo[k]), k
    assert differs, "the two kernel families associate the sum differently; identical bits mean the flag was ignored"


@pytest.mark.parametrize("name", ["one_chain", "coupling_r20"])
def test_team_result_does_not_depend_on_cluster_size_or_pipeline_depth(name):
    case = TEAM_CASES[name]
    ref = run_gpu(case, "philox", seed=8, flags=flags().FLAG_TEAM_KERNEL)
    try:
        for ctas, stages in (("1", "2"), ("2", "3"), ("4", "12"), ("8", "8"), ("16", "5")):
            os.environ["DRJ_TEAM_CTAS"], os.environ["DRJ_TEAM_STAGES"] = ctas, stages
            out = run_gpu(case, "philox", seed=8, flags=flags().FLAG_TEAM_KERNEL)
            for k in ref:
                if k != "t_diff":
                    assert np.array_equal(ref[k], out[k]), (ctas, stages, k)
    finally:
        os.environ.pop("DRJ_TEAM_CTAS", None)
        os.environ.pop("DRJ_TEAM_STAGES", None)


def test_team_tma_and_plain_staging_agree():
    case = TEAM_CASES["rungs3_hot"]
    a = run_gpu(case, "philox", seed=2, flags=flags().FLAG_TEAM_KERNEL)
    b = run_gpu(case, "philox", seed=2, flags=flags().FLAG_TEAM_KERNEL | flags().FLAG_NO_TMA)
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), k


def test_team_launch_split_and_sharding_are_bitwise_invariant():
    case = TEAM_CASES["rungs3_hot"]
    full = run_gpu(case, "philox", seed=21, chain_offset=10, flags=flags().FLAG_TEAM_KERNEL)
    split = run_gpu(case, "philox", seed=21, chain_offset=10, flags=flags().FLAG_TEAM_KERNEL, iters_per_launch=3)
    lo, hi = 1, 4
    shard = Case(case.model, case.tmin, case.tmax, data=case.data, beta=case.beta, chains=hi - lo, burnin=case.burnin,
                 samples=case.samples, init=case.init[lo:hi], save_hot_draws=True)
    part = run_gpu(shard, "philox", seed=21, chain_offset=10 + lo, flags=flags().FLAG_TEAM_KERNEL)
    for k in full:
        if k != "t_diff":
            assert np.array_equal(full[k], split[k]), k
            assert np.array_equal(part[k], full[k][lo:hi]), k


def test_team_small_and_empty_data(oracle):
    """Forced onto data shorter than a tile, and onto an empty vector: ranks without tiles contribute nothing."""
    for n in (0, 1, 7, 100):
        case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(n)], beta=beta_ladder(2), chains=2, burnin=8,
                    samples=8)
        u = case.replay()
        ora = run_oracle(oracle, case, "replay", u)
        gpu = run_gpu(case, "replay", u, flags=flags().FLAG_TEAM_KERNEL)
        assert_parity(gpu, ora, RTOL, what=f"[team/n={n}] ")


def test_team_abort_is_uniform_across_the_cluster():
    """A NaN datum makes every proposal's likelihood NaN: the run must stop with the reference's error
    (Particle.h:121-124) -- all CTAs of the cluster leave the loop together, no hang."""
    import drjacoby_b200 as drj
    x = normal_data(20000)
    x[12345] = np.nan
    case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(2), chains=2, burnin=5, samples=5)
    with pytest.raises(drj.DrjacobyError) as e:
        run_gpu(case, "philox", seed=1, flags=flags().FLAG_TEAM_KERNEL)
    assert "NaN" in str(e.value)


def test_team_more_chains_than_sms():
    """160 chains: single-CTA clusters, more CTAs than SMs (several waves)."""
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(9001)], chains=160, burnin=4, samples=4)
    team = run_gpu(case, "philox", seed=6, flags=flags().FLAG_TEAM_KERNEL)
    solo = run_gpu(case, "philox", seed=6, flags=flags().FLAG_NO_TEAM_KERNEL)
    for k in team:
        if k == "t_diff":
            continue
        if team[k].dtype.kind == "f":
            ok, worst = rel_close(team[k], solo[k], rtol=RTOL)
            assert ok, (k, worst)
        else:
            assert np.array_equal(team[k], solo[k]), k


def test_team_full_size_data_against_sufficient_statistics():
    """n = 1e7 (configs[4] data size), 5 chains x 1 rung -- the reference's default chain count: every stored
    log-likelihood equals the closed form from (n, sum x, sum x^2) in long double."""
    from test_gpu_fullsize import loglike_from_sufficient_stats
    x = np.random.Generator(np.random.Philox(1)).normal(3.0, 2.0, 10_000_000)
    case = Case("example", [-10, 0], [10, np.inf], data=[x], chains=5, burnin=30, samples=30)
    out = run_gpu(case, "philox", seed=1, flags=flags().FLAG_TEAM_KERNEL)
    th = np.concatenate([out["theta_burnin"], out["theta_sampling"]], axis=2)
    ll = np.concatenate([out["loglike_burnin"], out["loglike_sampling"]], axis=2)
    want = loglike_from_sufficient_stats(x, th[..., 0], th[..., 1])
    rel = np.abs(ll - want) / np.abs(want)
    assert rel.max() < 1e-10, rel.max()
    assert out["accept_burnin"].sum() + out["accept_sampling"].sum() > 0
