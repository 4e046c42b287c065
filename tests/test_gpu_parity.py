"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical inputs.

Replay mode: both sides consume the SAME uniform stream in the reference's order;
gate = BASELINE.json's: log-likelihoods / states within 1e-10 relative, acceptance and coupling
counts bit-exact.  Philox mode: the oracle implements the same counter layout, so the same gate
applies to the production RNG path.
"""
import os

import numpy as np
import pytest

from helpers import CANCELLING_ATOL, Case, POPULATION, assert_parity, beta_ladder, run_gpu, run_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # BASELINE.json north_star: "within 1e-10 relative tolerance"


def normal_data(n, seed=3, mean=3.0, sd=2.0):
    return np.random.Generator(np.random.Philox(seed)).normal(mean, sd, n)


def multilevel_case(G, chains, rungs, burnin=30, samples=30):
    rng = np.random.Generator(np.random.Philox(5))
    mu = rng.normal(0, 1, G)
    data = [rng.normal(mu[g], 1.0, 5) for g in range(G)]
    return Case("multilevel", [-5.0] * G, [5.0] * G, data=data, blocks=[[g + 1, G + 1] for g in range(G)],
                beta=beta_ladder(rungs), chains=chains, burnin=burnin, samples=samples)


CASES = {
    # K1 model (inst/extdata/example_loglike_logprior.cpp), data resident in shared memory
    "example_n100": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=3, burnin=200, samples=200),
    "example_n100_odd": Case("example", [-10, 0], [10, np.inf], data=[normal_data(101)], chains=2),
    "example_exact": Case("example_exact", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=3),
    "example_rungs4": Case("example", [-10, 0], [10, np.inf], data=[normal_data(50)], beta=beta_ladder(4), chains=5,
                           burnin=100, samples=100),
    "example_rungs32_hot": Case("example", [-10, 0], [10, np.inf], data=[normal_data(20)], beta=beta_ladder(32, 2.0),
                                chains=9, save_hot_draws=True),
    # launch-geometry edge cases: rungs not dividing the warp / CTA size, inactive threads, one chain, the
    # maximum of 256 rungs (one chain per CTA)
    "example_rungs33": Case("example", [-10, 0], [10, np.inf], data=[normal_data(17)], beta=beta_ladder(33), chains=7,
                            burnin=12, samples=12),
    "example_rungs129": Case("example", [-10, 0], [10, np.inf], data=[normal_data(9)], beta=beta_ladder(129, 3.0), chains=3,
                             burnin=6, samples=6, save_hot_draws=True),
    "example_rungs256": Case("example", [-10, 0], [10, np.inf], data=[normal_data(8)], beta=beta_ladder(256), chains=2,
                             burnin=5, samples=5),
    # sigma on a bounded domain with lower bound 0 (logit transform): the HINTS hook reads log(sigma) from the Jacobian
    # logarithm log(sigma - 0) of that transform too; and with a lower bound that is not 0 the model takes its own
    "example_sigma_logit": Case("example", [-10, 0], [10, 10], data=[normal_data(60)], beta=beta_ladder(3), chains=4),
    "example_sigma_min_half": Case("example", [-10, 0.5], [10, np.inf], data=[normal_data(60)], beta=beta_ladder(2), chains=3),
    "example_one_chain": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=1, burnin=300, samples=300),
    # BASELINE configs[0] at its full length: 1 chain, 1 rung, burnin 1e3 + samples 1e4, n = 100 (22,000 dependent updates)
    "k1_full_length": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100, 1)], chains=1, burnin=1000, samples=10000),
    "k1_full_length_exact": Case("example_exact", [-10, 0], [10, np.inf], data=[normal_data(100, 1)], chains=1, burnin=1000,
                                 samples=10000),
    # empty data vector: the likelihood is an empty sum, the chain follows the prior
    "example_empty_data": Case("example", [-10, 0], [10, np.inf], data=[np.zeros(0)], beta=beta_ladder(2), chains=3),
    "example_exact_empty_data": Case("example_exact", [-10, 0], [10, np.inf], data=[np.zeros(0)], chains=2),
    # every parameter fixed (min == max): nothing to update, only the coupling pass runs
    "doublewell_all_fixed": Case("doublewell", [0.5, 30], [0.5, 30], beta=beta_ladder(4), chains=3, burnin=10, samples=10,
                                 save_hot_draws=True),
    "example_coupling_off": Case("example", [-10, 0], [10, np.inf], data=[normal_data(20)], beta=beta_ladder(3),
                                 chains=2, coupling_on=False),
    # data streamed through shared memory in tiles (n > 2 tiles), ragged last tile
    "example_stream": Case("example", [-10, 0], [10, np.inf], data=[normal_data(10001)], beta=beta_ladder(2), chains=3,
                           burnin=6, samples=6),
    # inst/extdata/coupling_loglike_logprior.cpp, 20 rungs (vignettes/metropolis_coupling.Rmd)
    "coupling_r20": Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(100, 4, 10.0, 1.0)],
                         beta=beta_ladder(20), chains=3, burnin=60, samples=60),
    "normal_mu": Case("normal_mu", [-np.inf], [np.inf], data=[normal_data(100)], chains=4),
    # K2: double well, 20 rungs, alpha = 2, gamma fixed (skip_param)
    "doublewell_k2": Case("doublewell", [-2, 30], [2, 30], beta=beta_ladder(20, 2.0), chains=7, burnin=100, samples=100),
    # K3: logistic growth, 10 rungs
    "logistic_k3": Case("logistic", [1, 1, 50], [100, 2, 200], data=[POPULATION["pop"], POPULATION["time"]],
                        beta=beta_ladder(10), chains=5, burnin=40, samples=40),
    # K4 family: multilevel blocks, shipped size (5 groups) and K4 size (50 groups, 51 blocks)
    "multilevel_g5": multilevel_case(5, 4, 4),
    "multilevel_g50": multilevel_case(50, 3, 16, burnin=10, samples=10),
    # all four transform types + a fixed parameter (test-mcmc-with-r-likelihood-and-prior.R:177-221)
    "returnprior": Case("returnprior", [-np.inf, -np.inf, 0, 0], [np.inf, 0, np.inf, 1], chains=4, burnin=150, samples=150),
    "testfile": Case("testfile", [-10, 0], [10, np.inf], data=[normal_data(1000)], chains=2),
    "testfile_nulllike": Case("testfile_nulllike", [-10, 0], [10, np.inf], data=[normal_data(10)], chains=2),
    # -Inf over half the domain with a beta = 0 rung (test-Inf_checks_work.R:156-243)
    "halfdomain_like": Case("halfdomain_like", [0], [1], beta=[0.0, 1.0], chains=3, save_hot_draws=True,
                            init=np.full((3, 1), 0.1), burnin=200, samples=200),
    "halfdomain_prior": Case("halfdomain_prior", [0], [1], beta=[0.0, 1.0], chains=3, save_hot_draws=True,
                             init=np.full((3, 1), 0.1), burnin=200, samples=200),
}


def blocks_case():
    rng = np.random.Generator(np.random.Philox(9))
    sizes = [10, 12, 14, 11, 12, 12]  # chickwts group sizes (vignettes/blocks.Rmd)
    data = [rng.normal(250 + 20 * g, 50.0, n) for g, n in enumerate(sizes)]
    tmin = [0.0] * 6 + [0.0, -np.inf, 0.0]
    tmax = [np.inf] * 9
    blocks = [[g + 1, 7] for g in range(6)] + [[1, 2, 3, 4, 5, 6], [7], [7]]
    init = np.tile(np.array([250.0] * 6 + [50.0, 250.0, 50.0]), (2, 1))
    return Case("blocks", tmin, tmax, data=data, blocks=blocks, chains=2, init=init, burnin=60, samples=60)


CASES["blocks_chickwts"] = blocks_case()


@pytest.mark.parametrize("family", ["auto", "thread"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_replay_matches_oracle(name, family, oracle):
    """`auto` = the kernel family the library picks for the shape (one warp per chain with five updates in flight for few
    single-rung chains, drj_spec.cuh; a group of lanes per particle for the other few-particle per-datum cases and the
    multilevel model, drj_lanes.cuh); `thread` = one thread per particle for every case."""
    from drjacoby_b200 import _lib
    case = CASES[name]
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    gpu = run_gpu(case, "replay", u, flags=(_lib.FLAG_NO_LANES_KERNEL | _lib.FLAG_NO_SPEC_KERNEL) if family == "thread" else 0)
    assert_parity(gpu, ora, RTOL, what=f"[{name}/{family}] ", atol=CANCELLING_ATOL.get(case.model, 0.0))


@pytest.mark.parametrize("name", ["example_n100", "example_rungs4", "doublewell_k2", "logistic_k3", "multilevel_g5",
                                  "returnprior", "coupling_r20"])
def test_philox_matches_oracle(name, oracle):
    case = CASES[name]
    ora = run_oracle(oracle, case, "philox", seed=1234, chain_offset=5)
    gpu = run_gpu(case, "philox", seed=1234, chain_offset=5)
    assert_parity(gpu, ora, RTOL, what=f"[{name}/philox] ", atol=CANCELLING_ATOL.get(case.model, 0.0))


def test_first_row_is_init(oracle):
    """tests/testthat/test-mcmc-with-r-likelihood-and-prior.R:170-172: iteration 1 == init exactly."""
    init = np.array([[-5.0, 1.5], [5.0, 0.5]])
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(30)], chains=2, init=init)
    gpu = run_gpu(case, "philox", seed=3)
    assert np.array_equal(gpu["theta_burnin"][:, 0, 0, :], init)


def test_stream_tma_and_plain_staging_agree():
    """The TMA bulk-copy pipeline and the plain cooperative staging must give identical bits."""
    from drjacoby_b200 import _lib
    case = CASES["example_stream"]
    a = run_gpu(case, "philox", seed=2)
    b = run_gpu(case, "philox", seed=2, flags=_lib.FLAG_NO_TMA)
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("mode", ["philox", "replay"])
@pytest.mark.parametrize("name", ["example_rungs4", "doublewell_k2", "example_rungs33", "example_coupling_off", "example_stream",
                                  "logistic_k3", "multilevel_g5"])
def test_draws_computed_ahead_are_the_same_draws(name, mode, monkeypatch):
    """A run too small to fill the machine (every case of this file) has the state-independent part of its updates -- the
    standard normal draw, log u of the accept test, log u of every coupling pair -- computed ahead by drj_draw_kernel
    (DESIGN 4); DRJ_DRAWS=0 makes every thread draw for itself, as the large runs do.  Same functions of the same
    counters: same bits, whatever the launch split."""
    case = CASES[name]
    from drjacoby_b200 import _lib
    kw = dict(flags=_lib.FLAG_NO_LANES_KERNEL | _lib.FLAG_NO_SPEC_KERNEL | _lib.FLAG_NO_TEAM_KERNEL)
    if mode == "replay":
        kw["replay"] = case.replay()
    monkeypatch.setenv("DRJ_DRAWS", "1")
    off = (1 << 33) + 2  # (a global chain id beyond 32 bits: both halves go into the Philox counter)
    ahead = run_gpu(case, mode, seed=6, chain_offset=off, **kw)
    split = run_gpu(case, mode, seed=6, chain_offset=off, iters_per_launch=5, **kw)
    monkeypatch.setenv("DRJ_DRAWS", "0")
    own = run_gpu(case, mode, seed=6, chain_offset=off, **kw)
    for k in ahead:
        if k != "t_diff":
            assert np.array_equal(ahead[k], own[k]), k
            assert np.array_equal(ahead[k], split[k]), k


@pytest.mark.parametrize("name", ["example_rungs4", "doublewell_k2", "example_rungs32_hot", "multilevel_g5", "logistic_k3"])
def test_parallel_coupling_decisions_are_the_serial_walk(name, monkeypatch):
    """Small runs with at most 32 rungs decide every pair of a coupling pass for every replica that could be carried
    into it, then walk the ladder on those bit masks (drj_sweep_body, SMALL); DRJ_COUPLE_PAR=0 is the serial walk of the
    large runs.  Same rounded expression (main.cpp:300-308): same bits, same swap counts."""
    case = CASES[name]
    from drjacoby_b200 import _lib
    kw = dict(flags=_lib.FLAG_NO_LANES_KERNEL | _lib.FLAG_NO_SPEC_KERNEL | _lib.FLAG_NO_TEAM_KERNEL)
    monkeypatch.setenv("DRJ_COUPLE_PAR", "1")
    par = run_gpu(case, "philox", seed=12, **kw)
    monkeypatch.setenv("DRJ_DRAWS", "0")
    par_own = run_gpu(case, "philox", seed=12, **kw)
    monkeypatch.setenv("DRJ_COUPLE_PAR", "0")
    serial = run_gpu(case, "philox", seed=12, **kw)
    assert par["mc_accept_sampling"].sum() > 0
    for k in par:
        if k != "t_diff":
            assert np.array_equal(par[k], serial[k]), k
            assert np.array_equal(par[k], par_own[k]), k


def test_draws_ahead_is_the_default_for_small_runs_only():
    """One extra launch (the draw kernel) per sweep launch when the run has at most two warps per scheduler."""
    import drjacoby_b200 as drj
    from helpers import gpu_spec
    case = CASES["doublewell_k2"]
    counts = {}
    for label, env in (("default", None), ("off", "0")):
        if env is None:
            os.environ.pop("DRJ_DRAWS", None)
        else:
            os.environ["DRJ_DRAWS"] = env
        try:
            s = drj.Session(drj.Model(builtin=case.model), gpu_spec(case, "philox", None, 1, 0, iters_per_launch=10))
            s.advance(30)
            counts[label] = s.launches()
            s.close()
        finally:
            os.environ.pop("DRJ_DRAWS", None)
    assert counts["default"][0] == counts["off"][0] == 3
    assert counts["default"][1] - counts["off"][1] == 3, counts


def test_chunked_launches_are_bitwise_identical():
    """State round-trips through HBM between launches: any launch split gives the same bits."""
    case = CASES["example_rungs4"]
    a = run_gpu(case, "philox", seed=9)
    b = run_gpu(case, "philox", seed=9, iters_per_launch=7)
    c = run_gpu(case, "philox", seed=9, chains_per_cta=1)
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), k
            assert np.array_equal(a[k], c[k]), k


def test_sharding_invariance():
    """Chains shard across GPUs by contiguous ranges; the Philox counter carries the GLOBAL chain id,
    so a shard reproduces its slice of the unsharded run bit for bit."""
    case = CASES["doublewell_k2"]
    full = run_gpu(case, "philox", seed=21, chain_offset=100)
    lo, hi = 3, 7
    shard = Case(case.model, case.tmin, case.tmax, beta=case.beta, chains=hi - lo, burnin=case.burnin,
                 samples=case.samples, init=case.init[lo:hi])
    part = run_gpu(shard, "philox", seed=21, chain_offset=100 + lo)
    for k in part:
        if k != "t_diff":
            assert np.array_equal(part[k], full[k][lo:hi]), k


def test_session_api_equals_one_shot():
    import drjacoby_b200 as drj
    from helpers import gpu_spec
    case = CASES["example_rungs4"]
    one = run_gpu(case, "philox", seed=5)
    s = drj.Session(drj.Model(builtin=case.model), gpu_spec(case, "philox", seed=5))
    ms = 0.0
    for n in (1, 13, 50, 1000):
        ms += s.advance(n)
    assert s.iterations_done() == case.burnin + case.samples
    out = s.download()
    sweeps, aux = s.launches()
    s.close()
    assert sweeps >= 4 and ms > 0
    for k in one:
        if k != "t_diff":
            assert np.array_equal(one[k], out[k]), k


@pytest.mark.parametrize("ll,lp,code", [(np.inf, 0.0, 2), (np.nan, 0.0, 2), (0.0, np.inf, 3), (0.0, np.nan, 3)])
def test_bad_values_abort(ll, lp, code, oracle):
    """tests/testthat/test-Inf_checks_work.R:2-94: Inf / NA / NaN from the likelihood or prior is an error."""
    import drjacoby_b200 as drj
    case = Case("const", [-10], [10], misc=[[ll], [lp]], chains=2, init=np.ones((2, 1)))
    with pytest.raises(drj.DrjacobyError) as e:
        run_gpu(case, "philox")
    assert e.value.code == code
    assert ("likelihood" if code == 2 else "prior") in str(e.value)
    with pytest.raises(oracle.OracleError) as eo:
        run_oracle(oracle, case, "philox")
    assert eo.value.code == code


def test_neg_inf_at_init_aborts(oracle):
    """tests/testthat/test-Inf_checks_work.R:97-153"""
    import drjacoby_b200 as drj
    for model in ("halfdomain_like", "halfdomain_prior"):
        case = Case(model, [0], [1], chains=2, init=np.array([[0.2], [0.9]]))
        with pytest.raises(drj.DrjacobyError) as e:
            run_gpu(case, "philox")
        assert e.value.code == 1 and e.value.chain == 1 and "Starting values" in str(e.value)


def test_halfdomain_semantics():
    """test-Inf_checks_work.R:183-243: the beta = 0 rung roams into the -Inf half, the cold rung never does."""
    case = CASES["halfdomain_like"]
    out = run_gpu(case, "philox", seed=1)
    th = np.concatenate([out["theta_burnin"], out["theta_sampling"]], axis=2)[..., 0]  # [chains, rungs, iters]
    ll = np.concatenate([out["loglike_burnin"], out["loglike_sampling"]], axis=2)
    assert (th[:, 0].max(axis=1) > 0.5).all() and (th[:, 1].max(axis=1) <= 0.5).all()
    assert np.all(np.isneginf(ll[:, 0][th[:, 0] > 0.5])) and np.all(np.isfinite(ll[:, 0][th[:, 0] < 0.5]))
    out2 = run_gpu(CASES["halfdomain_prior"], "philox", seed=1)
    th2 = np.concatenate([out2["theta_burnin"], out2["theta_sampling"]], axis=2)[..., 0]
    assert (th2 <= 0.5).all()


def test_random_configurations_match_oracle(oracle):
    """Differential fuzz: 40 random launch shapes / options (rungs 1..40, chains 1..37, data sizes around the
    accumulator-group, tile and residency boundaries, bounds, coupling on/off, hot draws, targets, seeds)."""
    from drjacoby_b200 import _lib
    rng = np.random.Generator(np.random.Philox(2024))
    sizes = [1, 2, 7, 8, 9, 15, 16, 17, 63, 100, 255, 2047, 2048, 2049, 4095, 4096, 4097, 6151]
    for trial in range(40):
        n = int(rng.choice(sizes))
        R = int(rng.integers(1, 41))
        chains = int(rng.integers(1, 38))
        iters = 6 if n > 2000 else int(rng.integers(5, 40))
        model = ["example", "example_exact", "testfile"][int(rng.integers(0, 3))] if n <= 300 else "example"
        lo = float(rng.choice([-10.0, -3.0, 0.5]))
        tmin, tmax = [lo, 0.0], [lo + float(rng.choice([2.0, 20.0])), float(rng.choice([np.inf, 50.0]))]
        case = Case(model, tmin, tmax, data=[normal_data(n, seed=100 + trial, mean=lo + 1.0, sd=1.5)],
                    beta=beta_ladder(R, float(rng.choice([1.0, 2.0, 3.5]))), chains=chains, burnin=iters, samples=iters,
                    coupling_on=bool(rng.integers(0, 2)), save_hot_draws=bool(rng.integers(0, 2)),
                    target=float(rng.choice([0.1, 0.44, 0.8])), seed=trial)
        seed, off = int(rng.integers(0, 2**40)), int(rng.integers(0, 10**6))
        kw = {}
        if rng.integers(0, 2):
            kw["iters_per_launch"] = int(rng.integers(1, 9))
        # kernel family: the library's choice, one thread per particle, or a group of lanes per particle where it applies
        kw["flags"] = int(rng.choice([0, _lib.FLAG_NO_LANES_KERNEL | _lib.FLAG_NO_SPEC_KERNEL, _lib.FLAG_LANES_KERNEL, _lib.FLAG_SPEC_KERNEL]))
        ora = run_oracle(oracle, case, "philox", seed=seed, chain_offset=off)
        gpu = run_gpu(case, "philox", seed=seed, chain_offset=off, **kw)
        assert_parity(gpu, ora, RTOL, what=f"[fuzz {trial}: {model} n={n} R={R} C={chains} {kw}] ", atol=CANCELLING_ATOL.get(model, 0.0))
