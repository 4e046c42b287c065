"""The LANES kernel (drjacoby_b200/csrc/drj_lanes.cuh): a group of 8 or 32 lanes of a warp per particle.

Two shapes use it: models with many parameters that declare a lane form (the d = 50 multilevel model, BASELINE
configs[3]: state in shared memory instead of local memory, block sums lane-parallel) and per-datum models with few
particles over a short data item (BASELINE configs[0]: ONE chain over n = 100 -- the reference's everyday shape,
R/main.R:213-216).  Gates: replay / Philox parity with the CPU oracle (1e-10 relative, counts bit-exact); agreement with
the one-thread-per-particle kernels inside the same gate (the likelihood sums associate differently); results bitwise
independent of the launch geometry, the launch split and the sharding of chains.
"""
import numpy as np
import pytest

from helpers import CANCELLING_ATOL, POPULATION, Case, assert_parity, beta_ladder, gpu_spec, rel_close, run_gpu, run_oracle
from test_gpu_parity import CASES, multilevel_case, normal_data

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def lib():
    from drjacoby_b200 import _lib
    return _lib


LANES_CASES = {
    # BASELINE configs[0] shape: one chain, one rung, n = 100 (one warp in the whole grid)
    "k1_one_chain": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=1, burnin=300, samples=300),
    # the reference's default: five chains (R/main.R:213-216)
    "five_chains": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=5, burnin=120, samples=120),
    # rungs: 3 (left-over particle slots in the CTA), 8 (the most a 32-lane group allows per 256 threads)
    "rungs3": Case("example", [-10, 0], [10, np.inf], data=[normal_data(50)], beta=beta_ladder(3, 2.0), chains=7, burnin=60,
                   samples=60, save_hot_draws=True),
    "rungs8": Case("example", [-10, 0], [10, np.inf], data=[normal_data(33)], beta=beta_ladder(8), chains=5, burnin=60,
                   samples=60),
    # data lengths around the four-accumulator pass (4 x 32) and the empty vector
    "n127": Case("example", [-10, 0], [10, np.inf], data=[normal_data(127)], chains=2),
    "n128": Case("example", [-10, 0], [10, np.inf], data=[normal_data(128)], chains=2),
    "n129": Case("example", [-10, 0], [10, np.inf], data=[normal_data(129)], beta=beta_ladder(2), chains=2),
    "n4096": Case("example", [-10, 0], [10, np.inf], data=[normal_data(4096)], chains=3, burnin=10, samples=10),
    "empty": Case("example", [-10, 0], [10, np.inf], data=[np.zeros(0)], beta=beta_ladder(2), chains=3),
    "coupling_off": Case("example", [-10, 0], [10, np.inf], data=[normal_data(20)], beta=beta_ladder(3), chains=2,
                         coupling_on=False),
    # three parameters (one unbounded), one parameter; beta = 0 rung
    "coupling_model": Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(100, 4, 10.0, 1.0)],
                           beta=beta_ladder(6), chains=3, burnin=60, samples=60),
    "normal_mu": Case("normal_mu", [-np.inf], [np.inf], data=[normal_data(100)], beta=beta_ladder(4), chains=4),
    # a fixed parameter (skip_param: no random numbers consumed), more chains than fit one CTA
    "sigma_fixed": Case("example", [-10, 2.0], [10, 2.0], data=[normal_data(60)], beta=beta_ladder(2), chains=40, burnin=30,
                        samples=30),
    # generic models with a lane form: 8 lanes per particle.  Logistic growth (K3's model): every lane walks the 100-step
    # map, the 20 Poisson terms are evaluated 8 at a time; 10 rungs = 3 chains per CTA + 2 left-over particle slots
    "logistic_k3": Case("logistic", [1, 1, 50], [100, 2, 200], data=[POPULATION["pop"], POPULATION["time"]],
                        beta=beta_ladder(10), chains=5, burnin=40, samples=40),
    "logistic_times_backwards": Case("logistic", [1, 1, 50], [100, 2, 200],
                                     data=[POPULATION["pop"][::-1], POPULATION["time"][::-1]], beta=beta_ladder(2), chains=2,
                                     burnin=15, samples=15),
    "multilevel_g5": multilevel_case(5, 4, 4),
    "multilevel_g50": multilevel_case(50, 3, 16, burnin=10, samples=10),
    "multilevel_g50_r5_hot": multilevel_case(50, 7, 5, burnin=8, samples=8),
    "multilevel_g50_r32": multilevel_case(50, 2, 32, burnin=6, samples=6),
}


@pytest.mark.parametrize("name", sorted(LANES_CASES))
def test_lanes_replay_matches_oracle(name, oracle):
    case = LANES_CASES[name]
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    gpu = run_gpu(case, "replay", u, flags=lib().FLAG_LANES_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[lanes/{name}] ", atol=CANCELLING_ATOL.get(case.model, 0.0))


@pytest.mark.parametrize("name", ["k1_one_chain", "rungs8", "coupling_model", "multilevel_g50", "logistic_k3"])
def test_lanes_philox_matches_oracle(name, oracle):
    case = LANES_CASES[name]
    ora = run_oracle(oracle, case, "philox", seed=99, chain_offset=(1 << 33) + 5)
    gpu = run_gpu(case, "philox", seed=99, chain_offset=(1 << 33) + 5, flags=lib().FLAG_LANES_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[lanes/{name}/philox] ", atol=CANCELLING_ATOL.get(case.model, 0.0))


@pytest.mark.parametrize("name", ["rungs3", "multilevel_g50"])
def test_lanes_is_the_default_and_agrees_with_one_thread_per_particle(name):
    """No flag selects the lanes kernel for these shapes (same bits as forcing it); DRJ_FLAG_NO_LANES_KERNEL gives the
    one-thread-per-particle result: equal within the parity gate, different in the last bits (other association)."""
    case = LANES_CASES[name]
    auto = run_gpu(case, "philox", seed=4)
    lanes = run_gpu(case, "philox", seed=4, flags=lib().FLAG_LANES_KERNEL)
    solo = run_gpu(case, "philox", seed=4, flags=lib().FLAG_NO_LANES_KERNEL | lib().FLAG_NO_SPEC_KERNEL)
    differs = False
    for k in auto:
        if k == "t_diff":
            continue
        assert np.array_equal(auto[k], lanes[k]), k
        if auto[k].dtype.kind == "f":
            ok, worst = rel_close(auto[k], solo[k], rtol=RTOL)
            assert ok, (k, worst)
            differs = differs or not np.array_equal(auto[k], solo[k])
        else:
            assert np.array_equal(auto[k], solo[k]), k
    assert differs, "the two kernel families associate the sums differently; identical bits mean the flag was ignored"


@pytest.mark.parametrize("name", ["rungs3", "multilevel_g50_r5_hot"])
def test_lanes_geometry_launch_split_and_sharding_are_bitwise_invariant(name):
    case = LANES_CASES[name]
    F = lib().FLAG_LANES_KERNEL
    full = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F)
    split = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F, iters_per_launch=3)
    one = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F, chains_per_cta=1)
    two = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F, chains_per_cta=2)
    lo, hi = 2, 6
    shard = Case(case.model, case.tmin, case.tmax, data=case.data, blocks=case.blocks, beta=case.beta, chains=hi - lo,
                 burnin=case.burnin, samples=case.samples, init=case.init[lo:hi], save_hot_draws=case.save_hot_draws)
    part = run_gpu(shard, "philox", seed=21, chain_offset=10 + lo, flags=F, layout_chains=case.chains)
    for k in full:
        if k != "t_diff":
            assert np.array_equal(full[k], split[k]), k
            assert np.array_equal(full[k], one[k]), k
            assert np.array_equal(full[k], two[k]), k
            assert np.array_equal(part[k], full[k][lo:hi]), k


def test_lanes_bad_values_abort():
    """A NaN datum makes every proposal's likelihood NaN: the run stops with the reference's error (Particle.h:121-124)."""
    import drjacoby_b200 as drj
    x = normal_data(100)
    x[37] = np.nan
    case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(2), chains=3, burnin=5, samples=5)
    with pytest.raises(drj.DrjacobyError) as e:
        run_gpu(case, "philox", seed=1, flags=lib().FLAG_LANES_KERNEL)
    assert e.value.code == 2 and "NaN" in str(e.value)


def test_lanes_user_model_in_the_per_datum_form_gets_the_kernel(oracle):
    """A user's per-datum model compiled with NVRTC runs on the lanes kernel for a few-particle run: same bits as the
    built-in it restates."""
    import drjacoby_b200 as drj
    src = r'''
struct MyModel {
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
#define DRJ_USER_MODEL MyModel
'''
    case = LANES_CASES["five_chains"]
    user = drj.Model(source=src)
    a = run_gpu(case, "philox", seed=3, model=user, flags=lib().FLAG_LANES_KERNEL)
    b = run_gpu(case, "philox", seed=3, flags=lib().FLAG_LANES_KERNEL)
    c = run_gpu(case, "philox", seed=3, flags=lib().FLAG_NO_LANES_KERNEL | lib().FLAG_NO_SPEC_KERNEL)
    differs = False
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), k
            differs = differs or not np.array_equal(a[k], c[k])
    assert differs


def test_lanes_k1_full_length(oracle):
    """BASELINE configs[0] at its full length through the lanes kernel: 1 chain, burnin 1e3 + samples 1e4, n = 100."""
    case = CASES["k1_full_length"]
    ora = run_oracle(oracle, case, "philox", seed=5)
    gpu = run_gpu(case, "philox", seed=5, flags=lib().FLAG_LANES_KERNEL)
    assert_parity(gpu, ora, RTOL, what="[lanes/k1 full length] ")


def test_lanes_user_model_with_a_lane_form(oracle):
    """A user's many-parameter model that declares the lane form (cuda_template.cu, last section), compiled with NVRTC:
    the multilevel model with 12 groups written by hand.  Replay parity with the CPU oracle on the lanes kernel, and
    the lanes kernel really runs (its sums associate differently from one thread per particle)."""
    import drjacoby_b200 as drj
    G = 12
    src = r'''
struct MyMultilevel {
  static constexpr int D = DRJ_D, NBLOCK = DRJ_NBLOCK;
  static constexpr bool DATUM_FORM = false;
  static constexpr int LANES = 8;
  __device__ static double loglike(const double* th, const DrjList& data, const DrjList&, int block) {
    double ret = 0.0;
    if (block == D + 1) { for (int i = 0; i < D; ++i) ret += R::dnorm(th[i], 0.0, 1.0, true); }
    else { const double* x = data.item(block - 1); for (long long i = 0; i < data.len(block - 1); ++i) ret += R::dnorm(x[i], th[block - 1], 1.0, true); }
    return ret;
  }
  __device__ static double loglike_lanes(const double* th, const DrjList& data, const DrjList&, int block, int lane) {
    double ret = 0.0;
    if (block == D + 1) { for (int i = lane; i < D; i += LANES) ret += R::dnorm(th[i], 0.0, 1.0, true); }
    else { const double* x = data.item(block - 1); for (long long i = lane; i < data.len(block - 1); i += LANES) ret += R::dnorm(x[i], th[block - 1], 1.0, true); }
    return ret;
  }
  __device__ static double logprior(const double*, const DrjList&) { return 0.0; }
};
#define DRJ_USER_MODEL MyMultilevel
'''
    case = multilevel_case(G, 3, 4, burnin=25, samples=25)
    user = drj.Model(source=src)
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    lanes = run_gpu(case, "replay", u, model=user, flags=lib().FLAG_LANES_KERNEL)
    auto = run_gpu(case, "replay", u, model=user)
    solo = run_gpu(case, "replay", u, model=user, flags=lib().FLAG_NO_LANES_KERNEL)
    assert_parity(lanes, ora, RTOL, what="[lanes/jit lane form] ", atol=CANCELLING_ATOL["multilevel"])
    assert_parity(solo, ora, RTOL, what="[thread/jit lane form] ", atol=CANCELLING_ATOL["multilevel"])
    differs = False
    for k in lanes:
        if k != "t_diff":
            assert np.array_equal(lanes[k], auto[k]), k  # a model with a lane form runs on the lanes kernel by default
            differs = differs or not np.array_equal(lanes[k], solo[k])
    assert differs
