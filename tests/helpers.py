"""Shared builders for the parity tests: one Case feeds the CPU oracle and the CUDA path identically."""
from __future__ import annotations

from dataclasses import dataclass, field

import os

import numpy as np

POPULATION = dict(  # data/population_data.rda of the reference, decoded in SURVEY.md 8(d)
    pop=[14, 26, 36, 30, 37, 57, 40, 88, 106, 83, 100, 93, 87, 102, 97, 92, 91, 102, 105, 117],
    time=[33, 37, 41, 43, 44, 48, 49, 60, 71, 73, 76, 78, 79, 87, 88, 91, 92, 95, 96, 98])


@dataclass
class Case:
    model: str
    tmin: list
    tmax: list
    data: list = field(default_factory=list)
    misc: list = field(default_factory=list)
    blocks: list | None = None
    beta: list = field(default_factory=lambda: [1.0])
    chains: int = 2
    burnin: int = 50
    samples: int = 50
    coupling_on: bool = True
    save_hot_draws: bool = False
    target: float = 0.44
    init: np.ndarray | None = None
    seed: int = 7

    def __post_init__(self):
        self.tmin = np.asarray(self.tmin, dtype=np.float64)
        self.tmax = np.asarray(self.tmax, dtype=np.float64)
        self.d = self.tmin.size
        if self.init is None:
            self.init = default_init(self.tmin, self.tmax, self.chains, self.seed)
        self.init = np.asarray(self.init, dtype=np.float64).reshape(self.chains, self.d)
        skip = self.tmin == self.tmax
        self.d_free = int((~skip).sum())
        R = len(self.beta)
        self.T = self.burnin - 1 + self.samples
        self.U = 3 * self.d_free * R + ((R - 1) if (self.coupling_on and R > 1) else 0)

    def replay(self, seed=11):
        rng = np.random.Generator(np.random.Philox(seed))
        u = rng.random(self.chains * self.T * self.U)
        # 32-bit resolution uniforms strictly inside (0,1), like R's unif_rand
        return np.clip(np.floor(u * 2.0**32) * 2.3283064365386963e-10, 1.164153218540398e-10, 1 - 1.164153218540398e-10)


def default_init(tmin, tmax, chains, seed):
    """R/main.R:299-314 with numpy uniforms."""
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.empty((chains, tmin.size))
    for i in range(tmin.size):
        p = rng.random(chains)
        tt = 2 * np.isfinite(tmin[i]) + np.isfinite(tmax[i])
        if tt == 0:
            out[:, i] = np.log(p) - np.log(1 - p)
        elif tt == 1:
            out[:, i] = np.log(p) + tmax[i]
        elif tt == 2:
            out[:, i] = tmin[i] - np.log(p)
        else:
            out[:, i] = tmin[i] + (tmax[i] - tmin[i]) * p
    return out


def beta_ladder(rungs, alpha=1.0):
    """R/main.R:268-274,335"""
    return (np.linspace(0.0, 1.0, rungs) ** alpha).tolist() if rungs > 1 else [1.0]


def run_oracle(po, case: Case, rng_mode="replay", replay=None, seed=1, chain_offset=0, n_threads=4):
    return po.run(case.model, case.data, case.misc, case.tmin, case.tmax, case.init, burnin=case.burnin,
                  samples=case.samples, beta_raised=case.beta, blocks=case.blocks, coupling_on=case.coupling_on,
                  save_hot_draws=case.save_hot_draws, target_acceptance=case.target, rng_mode=rng_mode, seed=seed,
                  chain_offset=chain_offset, replay=replay, n_threads=n_threads)


def gpu_spec(case: Case, rng_mode="replay", replay=None, seed=1, chain_offset=0, **kw):
    import drjacoby_b200 as drj
    return drj.RunSpec(theta_min=case.tmin, theta_max=case.tmax, theta_init=case.init, burnin=case.burnin,
                       samples=case.samples, beta_raised=case.beta, blocks=case.blocks, coupling_on=case.coupling_on,
                       save_hot_draws=case.save_hot_draws, target_acceptance=case.target, rng_mode=rng_mode,
                       seed=seed, chain_offset=chain_offset, replay=replay, data=case.data, misc=case.misc, **kw)


def run_gpu(case: Case, rng_mode="replay", replay=None, seed=1, chain_offset=0, model=None, **kw):
    import drjacoby_b200 as drj
    m = model if model is not None else drj.Model(builtin=case.model)
    return drj.run_raw(m, gpu_spec(case, rng_mode, replay, seed, chain_offset, **kw))


# Absolute floor for the two model families whose stored values CANCEL to (near) zero, where a relative measure of the
# result says nothing about the arithmetic that produced it; every other case is held to the pure relative gate.
#   * double well: loglike = -gamma (mu^2 - 1)^2 passes through 0 at mu = +-1.  Only DIFFERENCES of log-densities enter
#     the sampler: an absolute error of 1e-12 in log p is a relative error of 1e-12 in p itself, a hundred times
#     tighter than the gate;
#   * multilevel (symmetric +-5 domains): theta = (max e^phi + min) / (1 + e^phi) cancels near the middle of the domain,
#     the rounding error is an ulp of |min|, |max| (~1e-15), not of the tiny result -- 1e-12 is 1e-13 of the width.
# (Measured: with no floor anywhere, 5 of 278 GPU tests fail, all of these two families.)
CANCELLING_ATOL = {"doublewell": 1e-12, "multilevel": 1e-12}


def rel_close(a, b, rtol=1e-10, atol=0.0):
    """The north_star's gate: |a - b| <= rtol |b| (atol = 0 unless the caller names a cancelling quantity, see
    CANCELLING_ATOL).  Equal infinities and exact zeros pass."""
    a, b = np.asarray(a), np.asarray(b)
    same_inf = (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)))
    with np.errstate(invalid="ignore"):
        diff = np.where(same_inf, 0.0, np.abs(a - b))
    ok = same_inf | (diff <= atol + rtol * np.abs(b))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(diff == 0.0, 0.0, diff / (atol / rtol + np.abs(b)))
    return bool(ok.all()), float(np.nanmax(rel)) if rel.size else 0.0


def assert_parity(gpu: dict, ora, rtol=1e-10, what="", atol=0.0):
    """Replay-mode gate of BASELINE.json: states and log-likelihoods within 1e-10 relative,
    acceptance and coupling counts bit-exact."""
    for name in ["accept_burnin", "accept_sampling", "mc_accept_burnin", "mc_accept_sampling"]:
        g, o = gpu[name], getattr(ora, name)
        assert np.array_equal(g, o), f"{what}{name} differs: gpu {g.ravel()[:8]} oracle {o.ravel()[:8]}"
    for name in ["loglike_burnin", "logprior_burnin", "theta_burnin", "loglike_sampling", "logprior_sampling",
                 "theta_sampling", "bw_final"]:
        ok, worst = rel_close(gpu[name], getattr(ora, name), rtol=rtol, atol=atol)
        assert ok, f"{what}{name}: worst relative error {worst:.3e} > {rtol}"
