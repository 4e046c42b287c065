"""The SPEC kernels (drjacoby_b200/csrc/drj_spec.cuh): one CTA per chain, the next five or seven updates evaluated
speculatively (a binary tree of 31 / 127 proposals, one per thread), for the reference's everyday shape -- few single-rung chains of a small
model (run_mcmc defaults, R/main.R:213-216; BASELINE configs[0] is ONE chain over n = 100).

Every update on the path the chain really takes is the arithmetic of the one-thread-per-particle kernel in the same
order, so the gate here is stronger than the 1e-10 of the north star: BITWISE equality with that kernel for every
output array, in Philox and in replay mode, for any launch split and geometry -- plus the usual replay parity with the
CPU oracle.
"""
import numpy as np
import pytest

from helpers import Case, POPULATION, assert_parity, gpu_spec, run_gpu, run_oracle
from test_gpu_parity import CASES, normal_data

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def lib():
    from drjacoby_b200 import _lib
    return _lib


def thread_flags():
    return lib().FLAG_NO_SPEC_KERNEL | lib().FLAG_NO_LANES_KERNEL


SPEC_CASES = {
    # BASELINE configs[0] shape; 599 updates per phase boundary: launches end inside a pass of five
    "k1_one_chain": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=1, burnin=300, samples=300),
    "five_chains": Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=5, burnin=121, samples=77),
    "example_exact": Case("example_exact", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=3, burnin=40, samples=41),
    "n4096": Case("example", [-10, 0], [10, np.inf], data=[normal_data(4096)], chains=3, burnin=12, samples=13),
    "empty_data": Case("example", [-10, 0], [10, np.inf], data=[np.zeros(0)], chains=2),
    # one parameter; three parameters (one unbounded); four parameters with all four transform types
    "normal_mu": Case("normal_mu", [-np.inf], [np.inf], data=[normal_data(100)], chains=4, burnin=33, samples=34),
    "coupling_model": Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[normal_data(100, 4, 10.0, 1.0)], chains=3,
                           burnin=60, samples=61),
    "returnprior": Case("returnprior", [-np.inf, -np.inf, 0, 0], [np.inf, 0, np.inf, 1], chains=4, burnin=150, samples=150),
    # a fixed parameter (skip_param): one free parameter of two, every level of the tree proposes the same one
    "doublewell_gamma_fixed": Case("doublewell", [-2, 30], [2, 30], chains=7, burnin=100, samples=100),
    "sigma_fixed": Case("example", [-10, 2.0], [10, 2.0], data=[normal_data(60)], chains=3, burnin=30, samples=31),
    # generic models: logistic growth (K3's model, three bounded parameters), the test-suite model with the underflow clamp
    "logistic": Case("logistic", [1, 1, 50], [100, 2, 200], data=[POPULATION["pop"], POPULATION["time"]], chains=5, burnin=40,
                     samples=41),
    "testfile": Case("testfile", [-10, 0], [10, np.inf], data=[normal_data(1000)], chains=2),
    # -Inf over half the domain (test-Inf_checks_work.R:156-243): proposals with an infinite likelihood inside the tree
    "halfdomain_like": Case("halfdomain_like", [0], [1], chains=3, init=np.full((3, 1), 0.1), burnin=200, samples=200),
    "halfdomain_prior": Case("halfdomain_prior", [0], [1], chains=3, init=np.full((3, 1), 0.1), burnin=200, samples=200),
    # more chains than one CTA holds (8 warps), a non-default acceptance target
    "forty_chains": Case("example", [-10, 0], [10, np.inf], data=[normal_data(64)], chains=40, burnin=20, samples=23, target=0.2),
}


def assert_same_bits(a, b, what):
    for k in a:
        if k != "t_diff":
            assert np.array_equal(a[k], b[k]), f"{what}: {k} differs"


@pytest.mark.parametrize("depth", ["5", "7"])
@pytest.mark.parametrize("name", sorted(SPEC_CASES))
def test_spec_is_bitwise_one_thread_per_particle(name, depth, monkeypatch):
    """Both tree depths: five updates in flight (one warp per chain) and seven (four warps per chain)."""
    monkeypatch.setenv("DRJ_SPEC_DEPTH", depth)
    case = SPEC_CASES[name]
    spec = run_gpu(case, "philox", seed=31, chain_offset=(1 << 32) + 7, flags=lib().FLAG_SPEC_KERNEL)
    solo = run_gpu(case, "philox", seed=31, chain_offset=(1 << 32) + 7, flags=thread_flags())
    assert_same_bits(spec, solo, f"[spec/{name}]")
    assert spec["accept_burnin"].sum() + spec["accept_sampling"].sum() > 0 or name == "empty_data"


@pytest.mark.parametrize("name", ["k1_one_chain", "five_chains", "returnprior", "doublewell_gamma_fixed", "logistic",
                                  "halfdomain_like", "coupling_model"])
def test_spec_replay_matches_oracle(name, oracle):
    case = SPEC_CASES[name]
    u = case.replay()
    ora = run_oracle(oracle, case, "replay", u)
    gpu = run_gpu(case, "replay", u, flags=lib().FLAG_SPEC_KERNEL)
    assert_parity(gpu, ora, RTOL, what=f"[spec/{name}] ")
    solo = run_gpu(case, "replay", u, flags=thread_flags())
    assert_same_bits(gpu, solo, f"[spec/{name}/replay]")


def test_spec_is_the_default_for_few_single_rung_chains():
    """No flag: five chains x one rung selects the speculative kernel -- far fewer sweep launches are not the evidence
    (same count), the timing is not either; the kernel family is visible through DRJ_FLAG_NO_SPEC_KERNEL changing
    nothing in the bits (asserted above) and through the session's reported kernel time being shorter."""
    import drjacoby_b200 as drj
    case = SPEC_CASES["five_chains"]
    auto = run_gpu(case, "philox", seed=4)
    spec = run_gpu(case, "philox", seed=4, flags=lib().FLAG_SPEC_KERNEL)
    assert_same_bits(auto, spec, "[spec default]")
    # the speculative kernel advances five updates per pass: the same 600 iterations take well under the
    # one-thread-per-particle kernel time
    long_case = Case("example", [-10, 0], [10, np.inf], data=[normal_data(100)], chains=5, burnin=300, samples=300)
    ms = {}
    for name, fl in (("auto", 0), ("thread", thread_flags())):
        s = drj.Session(drj.Model(builtin="example"), gpu_spec(long_case, "philox", seed=1, flags=fl))
        s.advance(50)
        ms[name] = s.advance(500)
        s.close()
    assert ms["auto"] < 0.7 * ms["thread"], ms


def test_spec_launch_split_and_geometry_are_bitwise_invariant():
    case = SPEC_CASES["forty_chains"]
    F = lib().FLAG_SPEC_KERNEL
    ref = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F)
    import os
    for kw, depth in (({"iters_per_launch": 1}, "5"), ({"iters_per_launch": 3}, "7"), ({"iters_per_launch": 7}, "5"),
                      ({"iters_per_launch": 2}, "7"), ({}, "5"), ({}, "7")):
        os.environ["DRJ_SPEC_DEPTH"] = depth
        try:
            out = run_gpu(case, "philox", seed=21, chain_offset=10, flags=F, **kw)
        finally:
            os.environ.pop("DRJ_SPEC_DEPTH", None)
        assert_same_bits(ref, out, f"[spec {kw} depth {depth}]")
    lo, hi = 11, 29
    shard = Case(case.model, case.tmin, case.tmax, data=case.data, chains=hi - lo, burnin=case.burnin, samples=case.samples,
                 init=case.init[lo:hi], target=case.target)
    part = run_gpu(shard, "philox", seed=21, chain_offset=10 + lo, flags=F)
    for k in ref:
        if k != "t_diff":
            assert np.array_equal(part[k], ref[k][lo:hi]), k


@pytest.mark.parametrize("ll,lp,code", [(np.inf, 0.0, 2), (np.nan, 0.0, 2), (0.0, np.inf, 3), (0.0, np.nan, 3)])
def test_spec_bad_values_abort(ll, lp, code):
    """tests/testthat/test-Inf_checks_work.R:2-94: Inf / NA / NaN from the likelihood or prior is an error."""
    import drjacoby_b200 as drj
    case = Case("const", [-10], [10], misc=[[ll], [lp]], chains=2, init=np.ones((2, 1)))
    with pytest.raises(drj.DrjacobyError) as e:
        run_gpu(case, "philox", flags=lib().FLAG_SPEC_KERNEL)
    assert e.value.code == code
    assert ("likelihood" if code == 2 else "prior") in str(e.value)


def test_spec_error_report_equals_one_thread_per_particle():
    """A NaN datum: both kernel families stop at the same update of the same chain with the same offending theta."""
    import drjacoby_b200 as drj
    x = normal_data(100)
    x[37] = np.nan
    case = Case("example", [-10, 0], [10, np.inf], data=[x], chains=3, burnin=5, samples=5)
    msgs = []
    for fl in (lib().FLAG_SPEC_KERNEL, thread_flags()):
        with pytest.raises(drj.DrjacobyError) as e:
            run_gpu(case, "philox", seed=1, flags=fl)
        msgs.append((e.value.code, e.value.chain, str(e.value)))
    assert msgs[0] == msgs[1], msgs


def test_spec_user_model_compiled_at_run_time():
    """A user's generic model (cuda_template form) gets the speculative kernel through NVRTC: same bits as one thread per
    particle."""
    import drjacoby_b200 as drj
    src = r'''
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  const double* x = data.item(0);
  double ret = 0.0;
  for (long long i = 0; i < data.len(0); ++i) ret += R::dnorm(x[i], params[0], params[1], 1);
  return ret;
}
__device__ double logprior(const double* params, const DrjList& misc) {
  return R::dunif(params[0], -10.0, 10.0, 1) + R::dlnorm(params[1], 0.0, 1.0, 1);
}
'''
    case = SPEC_CASES["five_chains"]
    user = drj.Model(source=src)
    a = run_gpu(case, "philox", seed=3, model=user, flags=lib().FLAG_SPEC_KERNEL)
    b = run_gpu(case, "philox", seed=3, model=user, flags=thread_flags())
    assert_same_bits(a, b, "[spec/jit]")


def test_spec_k1_full_length(oracle):
    """BASELINE configs[0] at its full length: 1 chain, burnin 1e3 + samples 1e4, n = 100 (22,000 dependent updates)."""
    case = CASES["k1_full_length"]
    ora = run_oracle(oracle, case, "philox", seed=5)
    gpu = run_gpu(case, "philox", seed=5, flags=lib().FLAG_SPEC_KERNEL)
    assert_parity(gpu, ora, RTOL, what="[spec/k1 full length] ")
