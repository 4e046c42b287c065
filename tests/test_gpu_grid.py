"""The grid-per-run (GRID) kernels: a handful of particles over a data item far larger than L2 -- the HBM-bound
regime ("for few chains x huge n, switch to split-n"; the loop replaced is
inst/extdata/example_loglike_logprior.cpp:12-18).

Every CTA holds a replica of every particle; the item is read ONCE per sweep, 1/G of it by each of the G
co-resident CTAs (TMA chunk pipeline), accumulated for all particles, combined through per-virtual-rank partial
sums in global memory and one grid-wide barrier (drj_sampler.cuh, drj_sweep_item_grid).  Gates: replay / Philox
parity with the CPU oracle; results bitwise independent of the grid size, the pipeline depth, the launch split
and of how chains are sharded (given the whole run's layout_chains); agreement with the other kernel families
within the 1e-10 gate (the sums are associated differently).
"""
import os

import numpy as np
import pytest

from helpers import Case, assert_parity, beta_ladder, rel_close, run_gpu, run_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def normal_data(n, seed=3, mean=3.0, sd=2.0):
    return np.random.Generator(np.random.Philox(seed)).normal(mean, sd, n)


def F():
    from drjacoby_b200 import _lib
    return _lib


GRID_CASES = {
    # one particle, 9 chunks of 8192 (ragged last one): 256 slots for the single particle
    "one_chain": Case("example", [-10, 0], [10, np.inf], data=[normal_data(70001)], chains=1, burnin=20, samples=20),
    # odd length, ragged last chunk, 3 chains x 2 rungs: 42 slots, 4 idle threads
    "stream_r2": Case("example", [-10, 0], [10, np.inf], data=[normal_data(10001)], beta=beta_ladder(2), chains=3,
                      burnin=6, samples=6),
    # 40 particles: 6 slots; coupling among the replicas of every CTA
    "coupling_r20": Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(9000, 4, 10.0, 1.0)],
                         beta=beta_ladder(20), chains=2, burnin=15, samples=15),
    "rungs3_hot": Case("example", [-10, 0], [10, np.inf], data=[normal_data(6001)], beta=beta_ladder(3, 2.0), chains=5,
                       burnin=10, samples=10, save_hot_draws=True),
    # the most particles a CTA can replicate: 8 x 32 = 256, one slot each
    "full_cta": Case("example", [-10, 0], [10, np.inf], data=[normal_data(5000)], beta=beta_ladder(32), chains=8,
                     burnin=5, samples=5),
    # more chunks (1221) than virtual ranks (1184): ranks of one and two chunks
    "many_chunks": Case("normal_mu", [-np.inf], [np.inf], data=[normal_data(10_000_001)], chains=2, burnin=3, samples=3),
    # 2..8 particles in the whole run: the tile form (every thread a data slot, all particles' sums in registers);
    # the reference's default of five single-rung chains, odd length (a lone last datum), ragged last chunk
    "tile_five": Case("example", [-10, 0], [10, np.inf], data=[normal_data(50_003)], chains=5, burnin=12, samples=12),
    # eight particles (the most the tile form takes) with coupling, fewer data pairs (150) than threads
    "tile_eight": Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(301, 4, 10.0, 1.0)],
                       beta=beta_ladder(4), chains=2, burnin=15, samples=15),
}


@pytest.mark.parametrize("name", sorted(GRID_CASES))
def test_grid_replay_matches_oracle(name, oracle):
    case = GRID_CASES[name]
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    gpu = run_gpu(case, "replay", u, flags=F().FLAG_GRID_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[grid/{name}] ")


@pytest.mark.parametrize("name", ["one_chain", "coupling_r20"])
def test_grid_philox_matches_oracle(name, oracle):
    case = GRID_CASES[name]
    ora = run_oracle(oracle, case, "philox", seed=77, chain_offset=3)
    gpu = run_gpu(case, "philox", seed=77, chain_offset=3, flags=F().FLAG_GRID_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[grid/{name}/philox] ")


def same_bits(a, b, what=""):
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), (what, k)


def test_grid_is_the_default_for_few_particles_over_very_long_data():
    """No flag: 16 chains over 2^20 + 5 data points selects the grid kernels (same bits as forcing them); the
    cluster-per-chain and one-thread-per-particle kernels agree within the parity gate and differ in the last bits."""
    # (16 chains: with n ~ 1e6 the posterior is so narrow that the first proposals are all rejected and the stored values
    # are the 16 initial log-likelihoods -- 16 independent sums, which two different associations cannot all round alike)
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data((1 << 20) + 5)], chains=16, burnin=5, samples=5)
    auto = run_gpu(case, "philox", seed=4)
    grid = run_gpu(case, "philox", seed=4, flags=F().FLAG_GRID_KERNEL)
    team = run_gpu(case, "philox", seed=4, flags=F().FLAG_TEAM_KERNEL)
    solo = run_gpu(case, "philox", seed=4, flags=F().FLAG_NO_TEAM_KERNEL | F().FLAG_NO_GRID_KERNEL)
    same_bits(auto, grid, "auto vs forced")
    for other in (team, solo):
        differs = False
        for k in auto:
            if k == "t_diff":
                continue
            if auto[k].dtype.kind == "f":
                ok, worst = rel_close(auto[k], other[k], rtol=RTOL)
                assert ok, (k, worst)
                differs = differs or not np.array_equal(auto[k], other[k])
            else:
                assert np.array_equal(auto[k], other[k]), k
        assert differs, "the kernel families associate the sum differently; identical bits mean a flag was ignored"


@pytest.mark.parametrize("name", ["one_chain", "coupling_r20", "many_chunks", "tile_five"])
def test_grid_result_does_not_depend_on_grid_size_or_pipeline_depth(name):
    case = GRID_CASES[name]
    ref = run_gpu(case, "philox", seed=8, flags=F().FLAG_GRID_KERNEL)
    try:
        for ctas, stages in (("1", "2"), ("7", "3"), ("64", "12"), ("148", "4"), ("100", "8")):
            os.environ["DRJ_GRID_CTAS"], os.environ["DRJ_GRID_STAGES"] = ctas, stages
            same_bits(ref, run_gpu(case, "philox", seed=8, flags=F().FLAG_GRID_KERNEL), (ctas, stages))
    finally:
        os.environ.pop("DRJ_GRID_CTAS", None)
        os.environ.pop("DRJ_GRID_STAGES", None)


@pytest.mark.parametrize("name", ["tile_five", "tile_eight", "stream_r2"])
def test_grid_tile_form_agrees_with_the_slot_form(name):
    """2..8 particles take the tile form by default (DRJ_GRID_TILE=0 turns it off): it associates the sum differently
    from one thread per (slot, particle), so the two agree within the gate, bit for bit in every count, and not in
    every last bit; with and without TMA the tile form gives the same bits."""
    case = GRID_CASES[name]
    tile = run_gpu(case, "philox", seed=5, flags=F().FLAG_GRID_KERNEL)
    same_bits(tile, run_gpu(case, "philox", seed=5, flags=F().FLAG_GRID_KERNEL | F().FLAG_NO_TMA), "tile: TMA vs plain")
    try:
        os.environ["DRJ_GRID_TILE"] = "0"
        slot = run_gpu(case, "philox", seed=5, flags=F().FLAG_GRID_KERNEL)
    finally:
        os.environ.pop("DRJ_GRID_TILE", None)
    differs = False
    for k in tile:
        if k == "t_diff":
            continue
        if tile[k].dtype.kind == "f":
            ok, worst = rel_close(tile[k], slot[k], rtol=RTOL)
            assert ok, (k, worst)
            differs = differs or not np.array_equal(tile[k], slot[k])
        else:
            assert np.array_equal(tile[k], slot[k]), k
    assert differs, "identical bits: DRJ_GRID_TILE was ignored"


def test_grid_tma_and_plain_staging_agree():
    case = GRID_CASES["rungs3_hot"]
    a = run_gpu(case, "philox", seed=2, flags=F().FLAG_GRID_KERNEL)
    b = run_gpu(case, "philox", seed=2, flags=F().FLAG_GRID_KERNEL | F().FLAG_NO_TMA)
    same_bits(a, b)


def test_grid_launch_split_and_sharding_are_bitwise_invariant():
    """A shard that is told the whole run's chain count (layout_chains) uses the same slot layout and reproduces
    its slice of the unsharded run bit for bit; so does any split of the run into launches."""
    case = GRID_CASES["rungs3_hot"]
    full = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F().FLAG_GRID_KERNEL)
    split = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F().FLAG_GRID_KERNEL, iters_per_launch=3)
    same_bits(full, split)
    lo, hi = 1, 4
    shard = Case(case.model, case.tmin, case.tmax, data=case.data, beta=case.beta, chains=hi - lo, burnin=case.burnin,
                 samples=case.samples, init=case.init[lo:hi], save_hot_draws=True)
    part = run_gpu(shard, "philox", seed=21, chain_offset=10 + lo, flags=F().FLAG_GRID_KERNEL, layout_chains=case.chains)
    for k in full:
        if k != "t_diff":
            assert np.array_equal(part[k], full[k][lo:hi]), k
    # the tile form: a one-chain shard of the five-chain run (the layout, and with it the form, follows the WHOLE run)
    case = GRID_CASES["tile_five"]
    full = run_gpu(case, "philox", seed=21, flags=F().FLAG_GRID_KERNEL)
    same_bits(full, run_gpu(case, "philox", seed=21, flags=F().FLAG_GRID_KERNEL, iters_per_launch=5))
    for lo, hi in ((2, 3), (3, 5)):
        shard = Case(case.model, case.tmin, case.tmax, data=case.data, beta=case.beta, chains=hi - lo, burnin=case.burnin,
                     samples=case.samples, init=case.init[lo:hi])
        part = run_gpu(shard, "philox", seed=21, chain_offset=lo, flags=F().FLAG_GRID_KERNEL, layout_chains=case.chains)
        for k in full:
            if k != "t_diff":
                assert np.array_equal(part[k], full[k][lo:hi]), (k, lo, hi)


def test_grid_small_and_empty_data(oracle):
    """Forced onto data shorter than a chunk, and onto an empty vector: CTAs without chunks only meet at the barrier."""
    for n in (0, 1, 7, 100, 8193):
        case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(n)], beta=beta_ladder(2), chains=2, burnin=8,
                    samples=8)
        u = case.replay()
        ora = run_oracle(oracle, case, "replay", u)
        gpu = run_gpu(case, "replay", u, flags=F().FLAG_GRID_KERNEL)
        assert_parity(gpu, ora, RTOL, what=f"[grid/n={n}] ")


def test_grid_abort_is_uniform_across_the_grid():
    """A NaN datum makes every proposal's likelihood NaN: the run must stop with the reference's error
    (Particle.h:121-124) -- every CTA leaves the iteration loop together, nobody is left waiting at the barrier."""
    import drjacoby_b200 as drj
    x = normal_data(200000)
    x[123456] = np.nan
    case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(2), chains=2, burnin=5, samples=5)
    with pytest.raises(drj.DrjacobyError) as e:
        run_gpu(case, "philox", seed=1, flags=F().FLAG_GRID_KERNEL)
    assert "NaN" in str(e.value)


def test_grid_needs_few_particles():
    """More than 256 particles cannot be replicated in a CTA: the flag is ignored, the run still works."""
    case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(5000)], beta=beta_ladder(3), chains=100, burnin=4, samples=4)
    a = run_gpu(case, "philox", seed=6, flags=F().FLAG_GRID_KERNEL)
    b = run_gpu(case, "philox", seed=6)
    same_bits(a, b)


def test_grid_user_model_through_nvrtc():
    """A user model in the per-datum form gets the grid kernels through NVRTC: same bits as the built-in."""
    import drjacoby_b200 as drj
    src = r'''
    struct UserNormal {
      static constexpr int D = 2, NBLOCK = 1;
      static constexpr bool DATUM_FORM = true;
      typedef DrjNormConsts Consts;
      __device__ static int datum_item(int) { return 0; }
      __device__ static bool prepare(const double* th, const DrjList&, int, Consts& k) { return drj_norm_prepare(th[0], th[1], k); }
      __device__ static double datum(const Consts& k, double x, double acc) { return drj_norm_datum(k, x, acc); }
      __device__ static double finish(const Consts& k, double acc, long long n) { return drj_norm_finish(k, acc, n); }
      __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int) {
        return drj_norm_serial(data.item(0), data.len(0), th[0], th[1]);
      }
      __device__ static double logprior(const double* th, const DrjList&) { return -log(20.0) + R::dlnorm(th[1], 0.0, 1.0, 1); }
    };
    #define DRJ_USER_MODEL UserNormal
    '''
    case = GRID_CASES["stream_r2"]
    a = run_gpu(case, "philox", seed=5, flags=F().FLAG_GRID_KERNEL)
    b = run_gpu(case, "philox", seed=5, flags=F().FLAG_GRID_KERNEL, model=drj.Model(source=src))
    same_bits(a, b)


def test_grid_full_size_data_against_sufficient_statistics():
    """n = 1e7 (configs[4] data size), 5 chains x 1 rung -- the reference's default chain count -- on the grid
    kernels (the default for this shape): every stored log-likelihood equals the closed form from
    (n, sum x, sum x^2) in long double."""
    from test_gpu_fullsize import loglike_from_sufficient_stats
    x = np.random.Generator(np.random.Philox(1)).normal(3.0, 2.0, 10_000_000)
    case = Case("example", [-10, 0], [10, np.inf], data=[x], chains=5, burnin=30, samples=30)
    out = run_gpu(case, "philox", seed=1, flags=F().FLAG_GRID_KERNEL)
    auto = run_gpu(case, "philox", seed=1)
    same_bits(out, auto, "the default for 5 x 1 over 1e7")
    th = np.concatenate([out["theta_burnin"], out["theta_sampling"]], axis=2)
    ll = np.concatenate([out["loglike_burnin"], out["loglike_sampling"]], axis=2)
    want = loglike_from_sufficient_stats(x, th[..., 0], th[..., 1])
    rel = np.abs(ll - want) / np.abs(want)
    assert rel.max() < 1e-10, rel.max()
    assert out["accept_burnin"].sum() + out["accept_sampling"].sum() > 0
