#!/usr/bin/env python
"""bench.py -- benchmark of the drjacoby_b200 sampling core.

Headline workload (BASELINE.json configs[4], "K5"): normal model (mu, sigma) on n = 1e7 synthetic data
points, 1e5 chains x 32 rungs sharded over 8 GPUs = 12,500 chains x 32 rungs per GPU (weak scaling:
per-GPU work is fixed).  A "step" is ONE MCMC iteration of the hot path over all chain-rungs of the
shard: every rung updates both parameters (2 data sweeps of n points per chain-rung), samples are
stored, Metropolis coupling runs.  Metric: parameter updates per second (chains x rungs x free
parameters x iterations / time), whole job.

The same JSON line carries, under "configs", every other named shape of BASELINE.json / SURVEY 8(d) measured in
the same run on ALL ranks (chains sharded over the ranks, max-over-ranks device time): K1, K2, K3, K4, the n = 100
and n = 4096 sweeps over 12,500 x 32 chain-rungs per GPU, and the HBM-bound shapes (1 and 5 chains over n = 1e8 on
the grid-per-run kernels) -- each with its roofline fraction and, at N = 1, the CPU port timed beside it; plus
"multi_abi" (the whole box driven from ONE process through drj_run_mcmc_multi / drj_multi_*, incl. a bitwise
sharded-equals-single check) and "ess" (ESS/s through run_mcmc(output="summary", cluster="gpus")).

  python bench.py [--gpus N] [--steps K] [--warmup W]                # our arm
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # CPU arm (oracle port)

Lines: ONE JSON object on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "param_updates_per_sec"
UNIT = "param-updates/s"
N_DATA = 10_000_000
CHAINS_PER_GPU = 12_500
RUNGS = 32
D_FREE = 2
FLOP_PER_DATUM = 4.0  # sub, mul, fma(=2) per datum of the normal log-density
W_FIXED = 400.0       # FP64-equivalent instructions of the fixed part of a parameter update (Philox, qnorm, exp, logs, ...)
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 64 DFMA/clk/SM at the 1965 MHz boost clock
POPULATION = dict(  # data/population_data.rda of the reference
    pop=[14, 26, 36, 30, 37, 57, 40, 88, 106, 83, 100, 93, 87, 102, 97, 92, 91, 102, 105, 117],
    time=[33, 37, 41, 43, 44, 48, 49, 60, 71, 73, 76, 78, 79, 87, 88, 91, 92, 95, 96, 98])


def synth_data(n=N_DATA):
    return np.random.Generator(np.random.Philox(1)).normal(3.0, 2.0, n)


def default_init(chains, seed, tmin=(-10.0, 0.0), tmax=(10.0, np.inf)):
    """R/main.R:299-314: one uniform per chain and parameter, mapped by the parameter's transform type."""
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.empty((chains, len(tmin)))
    for i, (lo, hi) in enumerate(zip(tmin, tmax)):
        p = rng.random(chains)
        tt = 2 * np.isfinite(lo) + np.isfinite(hi)
        out[:, i] = (np.log(p) - np.log(1 - p)) if tt == 0 else (np.log(p) + hi) if tt == 1 else (lo - np.log(p)) if tt == 2 \
            else (lo + (hi - lo) * p)
    return out


def shard_range(chains, rank, world):
    base, rem = divmod(chains, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def k5_spec(drj, x, chains, chain_offset, device, burnin, samples, seed=1):
    return drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=default_init(chains, 1000 + chain_offset),
                       burnin=burnin, samples=samples, beta_raised=np.linspace(0.0, 1.0, RUNGS),
                       coupling_on=True, save_hot_draws=False, store_pt=True, rng_mode="philox", seed=seed,
                       chain_offset=chain_offset, device=device, data=[x], iters_per_launch=1)


CPU_RUNGS = 8


def cpu_baseline(x, cores, iters=2, model="example"):
    """CPU oracle (a port of the reference's loop; the Rcpp reference itself cannot be built here: no R)
    on `cores` threads.  Bounded sample of the K5 workload: `cores` chains x CPU_RUNGS rungs x `iters`
    iterations on the full data.  The cost of a parameter update is one likelihood evaluation over the n
    data points whatever the number of rungs, so fewer rungs shrink the sample, not the per-update cost.
    The 2 initial evaluations per rung (Particle::init_like) are counted as work, so that the rate is
    likelihood evaluations/s == parameter updates/s of a long run."""
    from oracle import pyoracle as po
    po.build()
    chains = cores
    t0 = time.perf_counter()
    po.run(model, [x], [], np.array([-10.0, 0.0]), np.array([10.0, np.inf]), default_init(chains, 1000),
           burnin=1, samples=iters, beta_raised=np.linspace(0.0, 1.0, CPU_RUNGS), rng_mode="philox", seed=1,
           n_threads=cores)
    dt = time.perf_counter() - t0
    evals = chains * CPU_RUNGS * D_FREE * (iters + 1)
    return {"value": evals / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{chains} chains x {CPU_RUNGS} rungs x {iters} iteration(s) + initial likelihood, n={x.size}: "
                      f"{evals} likelihood evaluations in {dt:.1f} s"}, dt


# ------------------------------------------------------------------------------------------------
# the named shapes (BASELINE.json configs[0..3] + the regimes SURVEY 8(d) singles out)
# ------------------------------------------------------------------------------------------------
def named_configs(world, x_long=None, only=None):
    """Every entry: the reference's shape (chains = whole job at `world` GPUs), the timed window, the algorithmic work per
    parameter update and what bounds it.  `full`: the window is the config's whole run (1e3 + 1e4)."""
    rng = np.random.Generator(np.random.Philox(1))
    x100 = rng.normal(3.0, 2.0, 100)
    G = 50
    mu = rng.normal(0, 1, G)
    data50 = [rng.normal(mu[g], 1.0, 5) for g in range(G)]
    ex = dict(model="example", tmin=[-10.0, 0.0], tmax=[10.0, np.inf])
    cfgs = {
        "K1": dict(ex, what="BASELINE configs[0]: normal model (mu, sigma), n = 100, 1 chain x 1 rung, burnin 1e3 + samples 1e4", data=[x100],
                   beta=[1.0], chains=1, burnin=1000, samples=10000, full=True, scale="replicas", flop=FLOP_PER_DATUM * 100 + W_FIXED,
                   bound="latency", cpu_chains=1,
                   kernel="drj_sweep_kernel_spec<example, 7>: one CTA per chain, seven updates in flight (127 speculative proposals per pass)"),
        "default5": dict(ex, what="run_mcmc's default shape (R/main.R:213-216: chains = 5, rungs = 1) on the K1 model and data, burnin 1e3 + "
                         "samples 1e4; the reference runs the five chains one after the other on one core unless a cluster is passed",
                         data=[x100], beta=[1.0], chains=5, burnin=1000, samples=10000, full=True, scale="replicas",
                         flop=FLOP_PER_DATUM * 100 + W_FIXED, bound="latency", cpu_chains=5, cpu_threads=1,
                         kernel="drj_sweep_kernel_spec<example, 7>: one CTA per chain, the five chains side by side"),
        "K2": dict(model="doublewell", tmin=[-2.0, 30.0], tmax=[2.0, 30.0], what="BASELINE configs[1]: double well, 64 chains x 20 rungs, alpha = 2, "
                   "Metropolis coupling, burnin 1e3 + samples 1e4", data=[], beta=(np.linspace(0, 1, 20) ** 2.0), chains=64, burnin=1000, samples=10000,
                   full=True, scale="strong", flop=W_FIXED + 6, bound="latency", cpu_chains=64,
                   kernel="drj_sweep_kernel_small<doublewell>: one thread per particle, draws computed ahead by drj_draw_kernel, coupling decisions "
                          "for every possible carried replica at once"),
        "K3": dict(model="logistic", tmin=[1.0, 1.0, 50.0], tmax=[100.0, 2.0, 200.0], what="BASELINE configs[2]: logistic growth fit to population_data "
                   "(100-step map + 20 Poisson terms per evaluation), 1024 chains x 10 rungs, burnin 1e3 + samples 1e4",
                   data=[POPULATION["pop"], POPULATION["time"]], beta=np.linspace(0, 1, 10), chains=1024, burnin=1000, samples=10000, full=True,
                   scale="strong", flop=100 * 4 + 20 * 40 + W_FIXED, bound="fp64", cpu_chains=1024,
                   kernel="drj_sweep_kernel_lanes<logistic>: 4 lanes per particle (every lane walks the map, the Poisson terms 4 at a time)"),
        "K4": dict(model="multilevel", tmin=[-5.0] * G, tmax=[5.0] * G, what="BASELINE configs[3]: multilevel block-update model, 50 parameters in 51 "
                   "overlapping blocks, 4096 chains x 16 rungs; timed window 200 iterations after 50 (a full 1e3 + 1e4 run is 3.6e10 updates)",
                   data=data50, blocks=[[g + 1, G + 1] for g in range(G)], beta=np.linspace(0, 1, 16), chains=4096, burnin=51, samples=200,
                   warm=50, full=False, scale="strong", flop=(5 + 50) * 4 + 51 + W_FIXED, bound="fp64", cpu_chains=None,
                   kernel="drj_sweep_kernel_lanes<multilevel<50>>: 8 lanes per particle, state in shared memory"),
        "n100": dict(ex, what="normal model, n = 100 (the K1 data), 12,500 chains x 32 rungs per GPU: the regime the 1e10 updates/s/box target refers to "
                     "(SURVEY 8(d))", data=[x100], beta=np.linspace(0, 1, RUNGS), chains=CHAINS_PER_GPU * world, burnin=101, samples=500, warm=100,
                     ipl=100, full=False, scale="weak", flop=FLOP_PER_DATUM * 100 + W_FIXED, bound="fp64", cpu_chains=None,
                     kernel="drj_sweep_kernel<example, STREAM=0>: one thread per particle, data resident in shared memory"),
        "n4096": dict(ex, what="normal model, n = 4096 (largest data set resident in shared memory), 12,500 chains x 32 rungs per GPU: "
                      ">= 1e10 updates/s/box AND >= 70 % of the FP64 roofline", data=[synth_data(4096)], beta=np.linspace(0, 1, RUNGS),
                      chains=CHAINS_PER_GPU * world, burnin=21, samples=40, warm=20, ipl=20, full=False, scale="weak",
                      flop=FLOP_PER_DATUM * 4096, bound="fp64", cpu_chains=None,
                      kernel="drj_sweep_kernel<example, STREAM=1>: one thread per particle, data resident in shared memory"),
    }
    if x_long is not None:
        n = x_long.size
        for c in (1, 5):
            cfgs[f"hbm_{c}x1"] = dict(ex, what=f"HBM-bound regime: normal model, n = {n} ({8 * n / 1e6:.0f} MB >> L2), {c} chain(s) x 1 rung on the "
                                      "grid-per-run kernels: every SM streams 1/148 of the data once per sweep for all chains (at N > 1 every GPU "
                                      "runs its own such job: a run of this kind is a handful of chains)", data=[x_long], beta=[1.0],
                                      chains=c, burnin=11, samples=20, warm=10, ipl=10, full=False, scale="replicas", flop=FLOP_PER_DATUM * n,
                                      bound="hbm", bytes_per_sweep=8.0 * n, cpu_chains=None,
                                      kernel="drj_sweep_kernel_grid<example>: the whole grid per run, 64 KB TMA chunks, one grid barrier per sweep"
                                      + (" (tile form: every thread a data slot for all chains)" if c > 1 else ""))
    if only:
        cfgs = {k: v for k, v in cfgs.items() if k in only}
    return cfgs


def config_spec(drj, c, lo, hi, device, seed=1):
    tmin, tmax = np.asarray(c["tmin"], float), np.asarray(c["tmax"], float)
    init = default_init(c["chains"], 4242, tmin, tmax)[lo:hi]
    return drj.RunSpec(theta_min=tmin, theta_max=tmax, theta_init=init, burnin=c["burnin"], samples=c["samples"], beta_raised=np.asarray(c["beta"], float),
                       blocks=c.get("blocks"), coupling_on=True, store_pt=False, rng_mode="philox", seed=seed, chain_offset=lo, device=device,
                       data=c["data"], iters_per_launch=c.get("ipl", 0), layout_chains=c["chains"])


def run_named_config(drj, name, c, rank, world, device, fp64_peak, hbm_peak, reduce_max):
    """One named shape on this rank's shard; device time = CUDA events on the session stream (sweep kernels + output
    transposes), max over ranks; value = the whole job's parameter updates / that time."""
    if c["scale"] == "replicas":
        lo, hi = 0, c["chains"]  # a single chain cannot be split: every GPU runs its own replica (different seed)
    else:
        lo, hi = shard_range(c["chains"], rank, world)
    R = len(c["beta"])
    tmin, tmax = np.asarray(c["tmin"], float), np.asarray(c["tmax"], float)
    d_free = int((tmin != tmax).sum())
    T = c["burnin"] - 1 + c["samples"]
    warm = c.get("warm", 0)
    dev_ms = sweep_ms = wall = 0.0
    launches = 0
    if hi > lo:
        model = drj.Model(builtin=c["model"])
        if c["full"]:  # untimed run of the same shape first (first-touch allocations, module load), then the timed full run
            s = drj.Session(model, config_spec(drj, c, lo, hi, device, seed=1 + rank if c["scale"] == "replicas" else 1))
            s.advance(min(T, 200))
            s.close()
        s = drj.Session(model, config_spec(drj, c, lo, hi, device, seed=1 + rank if c["scale"] == "replicas" else 1))
        if warm:
            s.advance(warm)
        l0, t0m = s.launches(), s.timing()
        t0 = time.perf_counter()
        s.advance(T - warm)
        wall = time.perf_counter() - t0
        l1, t1m = s.launches(), s.timing()
        s.close()
        sweep_ms, dev_ms = t1m[0] - t0m[0], t1m[1] - t0m[1]
        launches = (l1[0] - l0[0]) + (l1[1] - l0[1])
    dev_ms, sweep_ms, wall = reduce_max([dev_ms, sweep_ms, wall])
    jobs = world if c["scale"] == "replicas" else 1
    updates = jobs * c["chains"] * R * d_free * (T - warm)
    value = updates / (dev_ms * 1e-3)
    out = {"workload": c["what"], "kernel": c.get("kernel"), "chains_total": jobs * c["chains"], "rungs": R, "free_params": d_free, "iterations_timed": T - warm,
           "full_run": bool(c["full"]), "scaling": c["scale"], "value": value, "unit": UNIT, "chain_rung_updates_per_sec": value / d_free,
           "device_ms": dev_ms, "sweep_kernel_ms": sweep_ms, "wall_s": wall, "gpu_launches_this_rank": launches}
    if c["bound"] == "hbm":
        sweeps = d_free * (T - warm)  # one pass over the item per sweep, shared by all chains of the GPU
        gbs = c["bytes_per_sweep"] * sweeps / (sweep_ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "per": "GPU",
                           "algorithmic_bytes": c["bytes_per_sweep"] * sweeps,
                           "note": "8 n bytes per sweep: the item is read once for all chains; peak = MEASURED_PEAKS.json hbm_gbs"}
    else:
        tf = value * c["flop"] / 1e12 / world
        out["roofline"] = {"bound": c["bound"], "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak if fp64_peak else None,
                           "nominal_frac": tf / FP64_NOMINAL_TFLOPS, "per": "GPU", "algorithmic_flop_per_update": c["flop"],
                           "note": "algorithmic FP64 work per parameter update from SURVEY 8(d) (4 n sweep flop + ~400 fixed-part instructions; "
                                   "model-specific terms for K2-K4) x updates/s per GPU; peak = FP64 FMA probe of this run, nominal = 148 SM x 64 FMA/clk x 1965 MHz"
                                   + ("; too few particles to fill the machine (every thread one dependent chain): latency-bound, not a throughput shape"
                                      if c["bound"] == "latency" else "")}
    return out


def cpu_named_config(name, c, cores, budget_chains=None):
    """The CPU port on the same shape: the whole config where the host can finish it (K1, K2, K3), else a bounded sample."""
    from oracle import pyoracle as po
    po.build()
    tmin, tmax = np.asarray(c["tmin"], float), np.asarray(c["tmax"], float)
    R = len(c["beta"])
    d_free = int((tmin != tmax).sum())
    if c["cpu_chains"] is not None:
        chains, burnin, samples, same = c["cpu_chains"], c["burnin"], c["samples"], True
    elif c["bound"] == "hbm" or name == "n4096":
        chains, burnin, samples, same = min(cores, 8), 1, 1, False
    else:
        chains, burnin, samples, same = cores, c["burnin"], min(c["samples"], 100 if name == "K4" else 500), False
    init = default_init(c["chains"], 4242, tmin, tmax)[:chains]
    threads = c.get("cpu_threads") or min(cores, chains)
    t0 = time.perf_counter()
    po.run(c["model"], c["data"], [], tmin, tmax, init, burnin=burnin, samples=samples, beta_raised=np.asarray(c["beta"], float),
           blocks=c.get("blocks"), rng_mode="philox", seed=1, n_threads=threads)
    dt = time.perf_counter() - t0
    iters = burnin - 1 + samples + (1 if burnin == 1 else 0)  # (the initial likelihood counts as work in a 1-iteration sample)
    ups = chains * R * d_free * iters / dt
    return {"value": ups, "unit": UNIT, "cores": threads, "kind": "port", "same_config": same,
            "sample": f"{chains} chain(s) x {R} rung(s), burnin {burnin} + samples {samples}: {dt:.2f} s on {threads} thread(s)"}


# ------------------------------------------------------------------------------------------------
# one process, all GPUs: the C-ABI multi-device entry points (what the R shim binds)
# ------------------------------------------------------------------------------------------------
ESS_CHAINS_PER_GPU = 65536


def multi_abi_leg(drj, n_gpus):
    """Rank 0 alone drives all `n_gpus` devices through drj_multi_* (one host thread per device inside the library):
    (1) a small run sharded over the devices equals the one-device run bit for bit; (2) throughput of the n = 4096
    shape (12,500 chains x 32 rungs per GPU) through this entry."""
    devs = list(range(n_gpus))
    rng = np.random.Generator(np.random.Philox(9))
    xs = rng.normal(3.0, 2.0, 300)
    small = dict(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=default_init(37, 5), burnin=50, samples=100,
                 beta_raised=np.linspace(0.0, 1.0, 4), rng_mode="philox", seed=3, data=[xs])
    model = drj.Model(builtin="example")
    one = drj.run_raw(model, drj.RunSpec(device=0, **small))
    many = drj.engine.run_raw_multi(model, drj.RunSpec(**small), devs if n_gpus > 1 else [0, 0, 0])
    equal = all(np.array_equal(one[k], many[k]) for k in one if k != "t_diff")
    xl = rng.normal(3.0, 2.0, 30001)  # few chains x long data: the cluster-per-chain family
    few = dict(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=default_init(5, 6), burnin=10, samples=20,
               beta_raised=np.linspace(0.0, 1.0, 3), rng_mode="philox", seed=4, data=[xl])
    one = drj.run_raw(model, drj.RunSpec(device=0, **few))
    many = drj.engine.run_raw_multi(model, drj.RunSpec(**few), devs if n_gpus > 1 else [0, 0, 0])
    equal_few = all(np.array_equal(one[k], many[k]) for k in one if k != "t_diff")
    chains = CHAINS_PER_GPU * n_gpus
    spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=default_init(chains, 78), burnin=1, samples=60,
                       beta_raised=np.linspace(0.0, 1.0, RUNGS), store_pt=False, rng_mode="philox", seed=2, data=[synth_data(4096)],
                       iters_per_launch=20)
    s = drj.Session(model, spec, devices=devs)
    s.advance(20)
    t0 = time.perf_counter()
    ms = s.advance(40)
    wall = time.perf_counter() - t0
    s.close()
    ups = chains * RUNGS * D_FREE * 40
    return {"entry": "drj_multi_create / drj_multi_advance (one process, one host thread per device)", "n_devices": n_gpus,
            "shards_bitwise_equal": bool(equal and equal_few),
            "checked": "37 chains x 4 rungs (one thread per particle) and 5 chains x 3 rungs over n = 30001 (cluster per chain): "
                       + ("sharded over all devices" if n_gpus > 1 else "three shards on the one device") + " == one-device run, every output array",
            "workload": f"normal model n=4096, {CHAINS_PER_GPU} chains x {RUNGS} rungs per GPU", "value": ups / (ms * 1e-3),
            "value_wall": ups / wall, "unit": UNIT, "kernel_ms_max_over_devices": ms, "wall_s": wall}


def ess_leg(drj, n_gpus):
    """ESS/s, SURVEY 8(d)(iii): coda::effectiveSize of the cold-rung sampling draws of ALL chains concatenated
    (R/main.R:521-526), min over the free parameters, divided by the wall time of the run.  K1 model (n = 100, 1 rung,
    burnin 1e3 + samples 1e4), 65,536 chains per GPU, through the public API in ONE process:
    run_mcmc(output="summary", cluster="gpus") -> drj_multi_create / advance / drj_multi_summary: the chains are split
    over the devices inside the library, the draws are reduced on each device and the moments / autocovariance sums
    merged in C -- exactly the ESS of the concatenated run, the 7.2e8 draws per GPU never move.
    One untimed call of the same shape first (first allocation of the stored draws, later taken from the pool)."""
    x = synth_data(100)
    chains, samples = ESS_CHAINS_PER_GPU * n_gpus, 10000
    df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
    drj.use_builtin("example")
    kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=1000, samples=samples, chains=chains, seed=3,
              silent=True, output="summary", cluster=list(range(n_gpus)))
    drj.run_mcmc(**kw)
    walls = []
    for _ in range(2):
        t0 = time.perf_counter()
        out = drj.run_mcmc(**kw)
        walls.append(time.perf_counter() - t0)
    dt = min(walls)
    t0 = time.perf_counter()
    outq = drj.run_mcmc(quantiles=[0.025, 0.5, 0.975], **kw)
    wall_q = time.perf_counter() - t0
    ess = [float(out.diagnostics["ess"][k]) for k in ("mu", "sigma")]
    rhat = [float(out.diagnostics["rhat"][k]) for k in ("mu", "sigma")]
    return {"workload": f"K1 model n=100, {ESS_CHAINS_PER_GPU} chains x 1 rung per GPU, 1e3+1e4 iterations, run_mcmc(output='summary', cluster='gpus')",
            "n_gpus": n_gpus, "wall_s": dt, "wall_s_runs": walls, "ess_mu_sigma": ess, "rhat_mu_sigma": rhat, "draws": chains * samples,
            "value": float(min(ess) / dt), "unit": "ESS/s",
            "with_quantiles": {"wall_s": wall_q, "probs": [0.025, 0.5, 0.975], "mu": [float(v) for v in outq.diagnostics["quantiles"]["mu"]],
                               "sigma": [float(v) for v in outq.diagnostics["quantiles"]["sigma"]],
                               "what": "the same run + exact quantiles of all draws by an on-device radix select (6 histogram passes)"},
            "definition": "coda::effectiveSize of the concatenated cold-rung sampling draws of all chains on all GPUs, min over "
                          "parameters, / wall of the API call (init, 11,000 iterations, on-device reduction, in-library merge over the devices)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import drjacoby_b200 as drj

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_group = None
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")  # host-side waits (see idle_wait)
    n_gpus = world
    device = local
    torch.cuda.set_device(device)
    K, W = args.steps, args.warmup
    chains = args.chains_per_gpu
    chain_offset = rank * chains

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def idle_wait():
        """Wait for the other ranks WITHOUT touching the GPU: an NCCL barrier is a kernel that spins on the device until every
        rank has joined, and while rank 0 drives all GPUs from one process (multi_abi / ess legs) that spinning kernel of the
        waiting rank would share its GPU with rank 0's work by time slicing (measured: the legs ran 2x slower)."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def reduce_max(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    # inputs in pinned host memory (the e2e leg copies them every step)
    x_pin = torch.empty(args.n_data, dtype=torch.float64).pin_memory()
    x = x_pin.numpy()
    x[:] = synth_data(args.n_data)

    model = drj.Model(builtin="example")
    sess = drj.Session(model, k5_spec(drj, x, chains, chain_offset, device, burnin=1, samples=W + K))
    fp64_peak = drj.fp64_peak_tflops(device)

    for _ in range(W):
        sess.advance(1)
    l0 = sess.launches()
    sw0, st0 = sess.timing()
    clocks = ClockSampler(device)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        sess.advance(1)  # one launch of the sweep kernel = one iteration over every chain-rung (+ output transposes)
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    l1 = sess.launches()
    sw1, st1 = sess.timing()
    dev_s = (st1 - st0) * 1e-3      # CUDA events on the session stream: sweep kernels + transposes
    sweep_s = (sw1 - sw0) * 1e-3    # sweep kernels only
    dev_s, sweep_s, wall = reduce_max([dev_s, sweep_s, wall])
    out = sess.download()
    sess.close()
    updates_per_gpu = chains * RUNGS * D_FREE * K
    value = n_gpus * updates_per_gpu / dev_s

    # ---- end to end through the C ABI with host buffers: H2D of the data and inits, init, E2E_IT
    # iterations, D2H of every output, inside the timed region
    E2E_IT = args.e2e_iters
    e2e_calls = max(1, args.e2e_calls)
    spec_e = k5_spec(drj, x, chains, chain_offset, device, burnin=1, samples=E2E_IT, seed=4)
    drj.run_raw(model, k5_spec(drj, x, chains, chain_offset, device, burnin=1, samples=1, seed=4))  # untimed warm-up call
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        out_e = drj.run_raw(model, spec_e)
    barrier()
    e2e_wall = reduce_max([time.perf_counter() - t0])[0]
    e2e_value = n_gpus * chains * RUNGS * D_FREE * E2E_IT * e2e_calls / e2e_wall
    h2d = x.nbytes + spec_e.theta_init.nbytes + 8 * (2 * 2 + RUNGS)     # per drj_run_mcmc call
    d2h = sum(v.nbytes for v in out_e.values())

    # ---- diagnostics gathered over NCCL at the end (the only collective; not on the hot path)
    cold_mu = out["theta_sampling"][:, 0, -1, 0]
    summ = torch.tensor([cold_mu.sum(), float(cold_mu.size), float(out["accept_sampling"][:, -1].sum())],
                        device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(summ)
    del out, out_e

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    cores = os.cpu_count() or 1
    with_cpu = rank == 0 and n_gpus == 1 and not args.no_cpu

    # ---- every named shape, on all ranks
    configs_out = {}
    if not args.no_sweeps:
        only = set(args.configs.split(",")) if args.configs else None
        x_long = None
        if args.hbm_n > 0 and (only is None or any(k.startswith("hbm") for k in only)):
            x_long = synth_data(args.hbm_n)
        for name, c in named_configs(world, x_long, only).items():
            barrier()
            res = run_named_config(drj, name, c, rank, world, device, fp64_peak, hbm_peak, reduce_max)
            if with_cpu:
                res["cpu_baseline"] = cpu_named_config(name, c, cores)
            if rank == 0:
                configs_out[name] = res
        del x_long
    drj.release_cached_memory()  # the one-process legs below use every device from rank 0
    idle_wait()

    extra = {}
    if rank == 0 and not args.no_sweeps:
        extra["multi_abi"] = multi_abi_leg(drj, n_gpus)
        extra["ess"] = ess_leg(drj, n_gpus)
        drj.release_cached_memory()
    idle_wait()

    cpu = None
    if with_cpu:
        cpu, _ = cpu_baseline(x if args.n_data <= N_DATA else x[:N_DATA], cores, iters=args.cpu_iters)
        # the same with the reference's per-call copy of the data vector (example_loglike_logprior.cpp:12)
        cpu_copy, _ = cpu_baseline(x if args.n_data <= N_DATA else x[:N_DATA], cores, iters=1, model="example_copy")
        cpu["value_with_per_call_data_copy"] = cpu_copy["value"]

    if rank == 0:
        flops_per_launch = FLOP_PER_DATUM * args.n_data * chains * RUNGS * D_FREE  # one launch = one iteration
        launch_s = sweep_s / K
        achieved = flops_per_launch / launch_s / 1e12
        n_cta = -(-chains // max(1, 128 // RUNGS))  # 128-thread CTAs: chains_per_cta = 128 // rungs
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "chain_rung_updates_per_sec": value / D_FREE, "n_gpus": n_gpus, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"K5 shard: normal model (mu,sigma), n={args.n_data} data points, {chains} chains x {RUNGS} rungs per GPU "
                                   f"(BASELINE configs[4] = 1e5 chains x 32 rungs over 8 GPUs), Metropolis coupling on, Philox RNG",
                       "chains_per_gpu": chains, "rungs": RUNGS, "n_data": args.n_data, "free_params": D_FREE,
                       "step": "one MCMC iteration over all chain-rungs", "l2": f"inputs larger than L2 per sweep: data {x.nbytes/1e6:.0f} MB re-streamed by every CTA",
                       "parallelism": f"chains sharded over {n_gpus} GPU(s), no hot-path collective"},
            "timing": {"device_s": dev_s, "sweep_kernel_s": sweep_s, "wall_s": wall, "how": "CUDA events on the session stream, max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d / E2E_IT), "d2h_bytes_per_step": int(d2h / E2E_IT),
                    "h2d_bytes_per_call": int(h2d), "d2h_bytes_per_call": int(d2h), "steps_per_call": E2E_IT,
                    "what": f"drj_run_mcmc C-ABI call, host buffers (pinned data), {E2E_IT} iterations (steps) per call, {e2e_calls} call(s), includes the initial likelihood sweep"},
            "gpu_launches": int((l1[0] - l0[0]) + (l1[1] - l0[1])),
            "launches": {"sweep": int(l1[0] - l0[0]), "transpose": int(l1[1] - l0[1])},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                         "nominal_frac": achieved / FP64_NOMINAL_TFLOPS, "nominal_peak": FP64_NOMINAL_TFLOPS,
                         "traffic": None, "kernel": "drj_sweep_kernel<DrjModelExample, STREAM=true>",
                         "peak_source": "FP64 FMA probe (drj_fp64_peak_tflops) measured in this run; MEASURED_PEAKS.json has no FP64 figure; nominal_frac is "
                                        "against 148 SM x 64 FMA/clk x 1965 MHz",
                         "algorithmic_flop_per_launch": flops_per_launch, "launch_ms": launch_s * 1e3,
                         "traffic_note": "not measured in this run (null); the ncu --set full capture of this kernel is in profiles/ (DRAM bytes per launch there)",
                         "note": "achieved uses the ALGORITHMIC 4 flop per datum of SURVEY 8(d) (sub, mul, fma); the kernel spends 2 FP64-pipe "
                                 "instruction slots per datum (DADD + DFMA = 3 executed flop), and the peak probe is 2 flop per slot, so frac is "
                                 "exactly the fraction of the FP64 pipe's issue slots used by the data sweep",
                         "hbm": {"algorithmic_bytes_per_launch": 8.0 * args.n_data * D_FREE * n_cta, "peak_gbs": hbm_peak,
                                 "achieved_gbs": 8.0 * args.n_data * D_FREE * n_cta / launch_s / 1e9,
                                 "note": "every CTA streams the data once per sweep (served by L2 after the first reader); far below the HBM roof: FP64-bound"}},
            "cpu_baseline": cpu,
            "clocks": clk,
            "diag": {"mean_cold_mu_last": float(summ[0] / summ[1]), "cold_accepts": int(summ[2]),
                     "shards_bitwise_equal": extra.get("multi_abi", {}).get("shards_bitwise_equal")},
            "configs": configs_out,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference arm: the reference's CPU algorithm (oracle port; the Rcpp package itself cannot be
    built or run in this image: no R) on all host cores, bounded sample of the same workload; plus the WHOLE K1, K2
    and K3 configs (the shapes the host can finish), so that those comparisons are same-config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    x = synth_data(args.n_data)
    vals, samples, times = [], [], []
    from oracle import pyoracle as po
    po.build()
    total = args.warmup + args.steps
    budget_s = 120.0
    t_start = time.perf_counter()
    for i in range(total):
        b, dt = cpu_baseline(x, cores, iters=1)
        if i >= args.warmup:
            vals.append(b["value"])
            samples.append(b["sample"])
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and vals:
            break
    v = float(np.mean(vals))
    steps_done = len(vals)
    configs_out = {}
    if not args.no_sweeps:
        only = {"K1", "default5", "K2", "K3"} & set(args.configs.split(",")) if args.configs else {"K1", "default5", "K2", "K3"}
        for name, c in named_configs(1, None, only).items():
            configs_out[name] = dict(cpu_named_config(name, c, cores), workload=c["what"])
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"K5 bounded sample: normal model n={args.n_data}, {cores} chains x {CPU_RUNGS} rungs, 1 iteration per step "
                                   "(+ the initial likelihood evaluations, counted as work)", "rungs": CPU_RUNGS, "n_data": args.n_data},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": samples[-1]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "configs": configs_out}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-data", type=int, default=N_DATA)
    ap.add_argument("--chains-per-gpu", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--e2e-iters", type=int, default=4)
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--cpu-iters", type=int, default=2)
    ap.add_argument("--hbm-n", type=int, default=100_000_000, help="data points of the HBM-bound configs (0 = skip them)")
    ap.add_argument("--configs", default="", help="comma-separated subset of the named configs (default: all)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sweeps", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
