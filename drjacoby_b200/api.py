"""Host-side mirror of the reference's user interface for the sampling path.

The reference's host layer is R (R/main.R); R is not available in this image, so this module states
the same interface in Python with the same names, argument meaning, defaults and error behaviour:

    define_params(...)        R/main.R:48-86     (+ check_params :92-139)
    run_mcmc(...)             R/main.R:208-577   pre-processing :285-343, dispatch :392-402,
                                                 output object :411-577 (Appendix B of SURVEY.md)
    sample_chains(x, n)       R/samples.R:21-46
    cuda_template()           R/cpp_template.R:6-18 / inst/templates/cpp_template.cpp (CUDA analogue)

What differs by design: all chains go to the GPU in ONE native call (the reference loops chains in
R); `loglike` / `logprior` must name `__device__` functions of a CUDA model (built-in or compiled
with NVRTC) -- R/Python closures cannot run on the GPU and are rejected with an error, there is no
CPU fallback; `cluster` selects multi-GPU sharding (torch.distributed ranks) instead of a PSOCK
cluster.  The R shim for a machine that has R is in r_shim/ (see INTEGRATION.md).
"""
from __future__ import annotations

import os
import sys
import time
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np
import pandas as pd

from . import diagnostics as dg
from . import _lib, engine
from ._lib import DrjacobyError
from .rcompat import RStream

_PARAM_KEYS = ("name", "min", "max", "init", "block")

# ------------------------------------------------------------------------------------------------
# models ("sourceCpp" analogue)
# ------------------------------------------------------------------------------------------------
_CURRENT_SOURCE: dict[str, Any] = {"model": None}


def cuda_template(save_as: str | None = None) -> str:
    """The CUDA analogue of cpp_template(save_as) (R/cpp_template.R:6-18): writes the model template to
    `save_as` (must end in .cu) and returns its text; with no argument just returns the text."""
    text = open(os.path.join(os.path.dirname(__file__), "templates", "cuda_template.cu")).read()
    if save_as is not None:
        if not isinstance(save_as, str):
            raise TypeError("save_as must be a single character string")
        if os.path.splitext(save_as)[1] != ".cu":
            raise ValueError("File must be .cu")
        with open(save_as, "w") as f:
            f.write(text)
    return text


def source_cuda(code_or_path: str) -> "CudaSource":
    """Analogue of Rcpp::sourceCpp(file): registers the `__device__` functions of a CUDA model so that
    run_mcmc(loglike="loglike", logprior="logprior") finds them by name (the reference looks the
    compiled functions up in the global environment, R/main.R:586-591)."""
    code = open(code_or_path).read() if os.path.exists(code_or_path) else code_or_path
    src = CudaSource(code)
    _CURRENT_SOURCE["model"] = src
    return src


def use_builtin(name: str) -> "CudaSource":
    """Registers a built-in model (inst/extdata/*.cpp restated in CUDA, drj_models.cuh) under the
    function names "loglike" / "logprior"."""
    src = CudaSource(None, builtin=name)
    _CURRENT_SOURCE["model"] = src
    return src


@dataclass
class CudaSource:
    code: str | None
    builtin: str | None = None
    _models: dict = field(default_factory=dict)

    def model(self, loglike: str, logprior: str) -> engine.Model:
        key = (loglike, logprior)
        if key not in self._models:
            if self.builtin is not None:
                self._models[key] = engine.Model(builtin=self.builtin)
            else:
                self._models[key] = engine.Model(source=self.code, loglike=loglike, logprior=logprior)
        return self._models[key]


# ------------------------------------------------------------------------------------------------
# define_params / check_params
# ------------------------------------------------------------------------------------------------
def define_params(*args, **kwargs) -> pd.DataFrame:
    """define_params(name=, min=, max=[, init=][, block=], name=, ...) of the reference.

    Python cannot repeat keyword arguments, so the repeated (key, value) series is given either as a
    flat positional sequence  define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", ...)
    or as one dict per parameter  define_params(dict(name="mu", min=-10, max=10), dict(...)).
    """
    if kwargs:
        args = args + (dict(kwargs),)
    if len(args) == 0:
        raise ValueError("input cannot be empty")
    if all(isinstance(a, dict) for a in args):
        flat: list = []
        for d in args:
            for k in d:
                flat += [k, d[k]]
        args = tuple(flat)
    if len(args) % 2:
        raise ValueError("arguments must be (key, value) pairs")
    keys, vals = list(args[0::2]), list(args[1::2])
    for k in keys:
        if k not in _PARAM_KEYS:
            raise ValueError(f"{k} must be one of name, min, max, init, block")
    use_init, use_block = "init" in keys, "block" in keys
    cols = ["name", "min", "max"] + (["init"] if use_init else []) + (["block"] if use_block else [])
    if len(keys) % len(cols) != 0:
        raise ValueError("must have the same number of inputs per parameter")
    n_param = len(keys) // len(cols)
    if keys != cols * n_param:
        raise ValueError("arguments must repeat in the order " + ", ".join(cols))
    rows = {c: [vals[len(cols) * i + j] for i in range(n_param)] for j, c in enumerate(cols)}
    df = pd.DataFrame({"name": rows["name"], "min": np.asarray(rows["min"], dtype=float),
                       "max": np.asarray(rows["max"], dtype=float)})
    if use_init:
        df["init"] = pd.Series([np.atleast_1d(np.asarray(v, dtype=float)) for v in rows["init"]], dtype=object)
    if use_block:
        df["block"] = pd.Series([np.atleast_1d(np.asarray(v)).astype(int) for v in rows["block"]], dtype=object)
    return check_params(df)


def check_params(x: pd.DataFrame) -> pd.DataFrame:
    if not isinstance(x, pd.DataFrame):
        raise TypeError("df_params must be a data frame")
    for c in ("name", "min", "max"):
        if c not in x.columns:
            raise ValueError("df_params must contain the columns 'name', 'min', 'max'")
    if x["name"].duplicated().any():
        raise ValueError("parameter names must be unique")
    x = x.copy()
    if "init" in x.columns:
        x["init"] = pd.Series([np.atleast_1d(np.asarray(v, dtype=float)) for v in x["init"]], dtype=object, index=x.index)
    if "block" in x.columns:
        x["block"] = pd.Series([np.atleast_1d(np.asarray(v)).astype(int) for v in x["block"]], dtype=object, index=x.index)
    for i in range(len(x)):
        if not isinstance(x["name"].iloc[i], str):
            raise ValueError("parameter names must be character strings")
        lo, hi = float(x["min"].iloc[i]), float(x["max"].iloc[i])
        if not lo <= hi:
            raise ValueError("min values must be less than or equal to max values")
        if "init" in x.columns:
            v = x["init"].iloc[i]
            if np.any(v < lo):
                raise ValueError("init values must be greater than or equal to min values")
            if np.any(v > hi):
                raise ValueError("init values must be less than or equal to max values")
    return x


# ------------------------------------------------------------------------------------------------
# output object
# ------------------------------------------------------------------------------------------------
class DrjacobyOutput:
    """The `drjacoby_output` S3 object (R/main.R:411-577): output, pt, diagnostics, parameters.
    `raw` additionally keeps the native arrays (chain-major) for large runs."""

    def __init__(self, output, pt, diagnostics, parameters, raw):
        self.output, self.pt, self.diagnostics, self.parameters, self.raw = output, pt, diagnostics, parameters, raw

    def __len__(self):
        return 4  # tests/testthat/test-mcmc-with-cpp-likelihood-and-prior.R:66

    def __getitem__(self, k):
        return {"output": self.output, "pt": self.pt, "diagnostics": self.diagnostics, "parameters": self.parameters}[k]

    def __repr__(self):  # R/drjacoby_output.R:5-16
        return f"drjacoby output:\n{self.parameters['chains']} chains\n{self.parameters['rungs']} rungs\n" \
               f"{self.parameters['burnin']} burn-in iterations\n{self.parameters['samples']} sampling iterations\n" \
               f"{len(self.parameters['df_params'])} parameters"


def sample_chains(x: DrjacobyOutput, sample_n: int, keep_chain_index: bool = False, seed: int | None = None) -> pd.DataFrame:
    """R/samples.R:21-46: thin the sampling-phase draws of all chains to sample_n equally spaced rows."""
    if not isinstance(x, DrjacobyOutput):
        raise TypeError("x must be a drjacoby_output")
    if sample_n < 1 or int(sample_n) != sample_n:
        raise ValueError("sample_n must be a positive integer")
    if not isinstance(keep_chain_index, (bool, np.bool_)):
        raise TypeError("keep_chain_index must be a single logical")
    df = x.output[x.output["phase"] == "sampling"]
    df = df.drop(columns=["iteration", "phase", "logprior", "loglikelihood"] + ([] if keep_chain_index else ["chain"]))
    if sample_n > len(df):
        raise ValueError(f"sample_n cannot exceed the total number of samples over all chains ({len(df)})")
    # all_chains[seq.int(1, N, length.out = n), ]: R truncates fractional row indices
    idx = np.floor(np.linspace(1, len(df), int(sample_n))).astype(int) - 1
    out = df.iloc[idx].reset_index(drop=True)
    out["sample"] = np.arange(1, len(out) + 1)
    return out


# ------------------------------------------------------------------------------------------------
# run_mcmc
# ------------------------------------------------------------------------------------------------
def _assert_pos_int(v, name):
    if not (isinstance(v, (int, float, np.integer, np.floating)) and float(v) == int(v) and int(v) > 0):
        raise ValueError(f"{name} must be a single positive integer")
    return int(v)


def _flatten_list(d: dict, what: str):
    names, items = [], []
    for k, v in d.items():
        arr = np.asarray(v)
        if arr.dtype == object or not np.issubdtype(arr.dtype, np.number) and arr.dtype != bool:
            raise TypeError(f"{what}${k} must be numeric to be used on the GPU path")
        names.append(str(k))
        items.append(np.asarray(arr, dtype=np.float64).ravel())
    return names, items


def run_mcmc(data, df_params, misc=None, loglike=None, logprior=None, burnin=1e3, samples=1e4, rungs=1, chains=5,
             beta_manual=None, alpha=1.0, target_acceptance=0.44, cluster=None, coupling_on=True, pb_markdown=False,
             save_data=True, save_hot_draws=False, silent=False, *, source: CudaSource | None = None, seed=None,
             rng="philox", store_pt=True, device=None, as_frames=True, progress=None, output="full", thin=None,
             quantiles=None) -> DrjacobyOutput:
    """run_mcmc() of the reference (R/main.R:208-225) on the GPU.

    Extra keyword-only arguments: `source` (a CudaSource; default: the last source_cuda()/use_builtin()),
    `seed`, `rng` ("philox": counter-based device RNG; "r": R's set.seed(seed) stream generated on the
    host and replayed on the device, chain for chain what the reference would draw), `store_pt`
    (keep loglike/logprior of the heated rungs, as the reference always does), `device`, `as_frames`
    (build the data frames; False keeps only `.raw` for very large runs), `output` ("full": every draw comes
    back, as in the reference; "summary": the draws stay in GPU memory and Rhat / ESS / DIC are computed from
    on-device reductions -- for runs whose draws are too large to move).
    `thin` (with output="summary"): additionally bring back every thin-th stored draw of the cold rung, gathered on
    the device, as the usual `output` data frame with the original iteration numbers -- diagnostics stay exact (all
    draws, on the device), plots and quantiles work on the thinned frame (sample_chains() of the reference subsamples
    after the fact, R/samples.R; here the full set never has to leave the GPU).
    `quantiles` (with output="summary"): probabilities whose quantiles (R's quantile(), type 7) of every parameter's
    cold-rung sampling draws are found on the device by a radix select -> diagnostics["quantiles"].
    `cluster`: None = this process's GPU.  "gpus" (or a list of device ordinals) = ONE process, several GPUs: the chains
    are split over the devices INSIDE the library (drj_run_mcmc_multi / drj_multi_*: one host thread per device, results
    and summaries merged in C) -- what replaces the reference's PSOCK cluster (R/main.R:392-402) and what the R shim
    calls.  "torch" = shard chains over the ranks of an initialised torch.distributed group (one process per GPU),
    results gathered on rank 0.
    """
    t_start = time.perf_counter()
    # ---------- check inputs (R/main.R:236-283)
    if not isinstance(data, dict) or not all(isinstance(k, str) and k for k in data):
        raise TypeError("data must be a named list (dict)")
    misc = {} if misc is None else misc
    if not isinstance(misc, dict):
        raise TypeError("misc must be a list (dict)")
    for f, nm in ((loglike, "loglike"), (logprior, "logprior")):
        if callable(f):
            raise TypeError(f"{nm} is an R/Python closure: closures cannot run on the GPU path and there is no CPU "
                            "fallback; pass the name of a __device__ function of a CUDA model (see cuda_template())")
        if not isinstance(f, str):
            raise TypeError(f"{nm} must be the name (character string) of a device function")
    burnin, samples = _assert_pos_int(burnin, "burnin"), _assert_pos_int(samples, "samples")
    rungs, chains = _assert_pos_int(rungs, "rungs"), _assert_pos_int(chains, "chains")
    if not isinstance(coupling_on, (bool, np.bool_)):
        raise TypeError("coupling_on must be a single logical")
    if not alpha > 0:
        raise ValueError("alpha must be positive")
    if not 0.0 <= target_acceptance <= 1.0:
        raise ValueError("target_acceptance must be between 0 and 1")
    df_params = check_params(df_params)
    use_init, use_block = "init" in df_params.columns, "block" in df_params.columns
    d = len(df_params)
    if use_init:
        for v in df_params["init"]:
            if len(v) != 1 and len(v) != chains:
                raise ValueError("must define one df_params$init value per parameter, or alternatively a list of "
                                 "values one for each chain")
    if beta_manual is None:
        beta_manual = np.linspace(1.0, 0.0, rungs)[::-1].copy()  # rev(seq(1, 0, l = rungs))
    beta_manual = np.atleast_1d(np.asarray(beta_manual, dtype=float))
    rungs = beta_manual.size
    if np.any(beta_manual < 0) or np.any(beta_manual > 1):
        raise ValueError("beta_manual must be between 0 and 1")
    if np.any(np.diff(beta_manual) <= 0):
        raise ValueError("beta_manual must be increasing")
    if beta_manual[-1] != 1.0:
        raise ValueError("beta_manual must end in 1")
    if len(misc) > 0 and "block" in misc:
        raise ValueError("'block' is a reserved name within misc object")

    # ---------- pre-processing (R/main.R:285-343)
    tmin, tmax = df_params["min"].to_numpy(float), df_params["max"].to_numpy(float)
    trans_type = engine.trans_type_of(tmin, tmax)
    skip_param = tmin == tmax
    if rng not in ("philox", "r"):
        raise ValueError('rng must be "philox" or "r"')
    if rng == "r" and seed is None:
        raise ValueError('rng="r" needs seed= (the argument of set.seed) or an RStream')
    rstream = RStream(int(seed)) if (rng == "r" and not isinstance(seed, RStream)) else (seed if rng == "r" else None)
    host_rng = np.random.Generator(np.random.Philox(0 if seed is None else int(seed))) if rng != "r" else None
    if not use_init:
        init_list = []
        for i in range(d):  # one runif(chains) per parameter, fixed ones included (R/main.R:301-313)
            p = rstream.runif(chains) if rstream is not None else host_rng.random(chains)
            tt = trans_type[i]
            if tt == 0:
                init_list.append(np.log(p) - np.log(1 - p))
            elif tt == 1:
                init_list.append(np.log(p) + tmax[i])
            elif tt == 2:
                init_list.append(tmin[i] - np.log(p))
            else:
                init_list.append(tmin[i] + (tmax[i] - tmin[i]) * p)
        df_params = df_params.copy()
        df_params["init"] = pd.Series(init_list, dtype=object, index=df_params.index)
    if not use_block:
        df_params = df_params.copy()
        df_params["block"] = pd.Series([np.array([1])] * d, dtype=object, index=df_params.index)
    init_mat = np.stack([np.repeat(v, chains) if len(v) == 1 else np.asarray(v, dtype=float) for v in df_params["init"]])
    df_params = df_params.copy()
    df_params["trans_type"] = trans_type
    beta_raised = beta_manual ** alpha
    data_names, data_items = _flatten_list(data, "data")
    misc_names, misc_items = _flatten_list(misc, "misc")

    src = source if source is not None else _CURRENT_SOURCE["model"]
    if src is None:
        raise ValueError("no CUDA model registered: call source_cuda(...) or use_builtin(...) first, or pass source=")
    model = src.model(loglike, logprior)

    # ---------- sharding over ranks (cluster = "torch")
    rank, world = 0, 1
    devices = None  # in-library multi-GPU: "all" or a list of ordinals
    if cluster is not None:
        if isinstance(cluster, str) and cluster == "torch":
            import torch.distributed as dist
            rank, world = dist.get_rank(), dist.get_world_size()
        elif isinstance(cluster, str) and cluster == "gpus":
            devices = "all"
        elif isinstance(cluster, (list, tuple)) and len(cluster) and all(isinstance(v, (int, np.integer)) for v in cluster):
            devices = [int(v) for v in cluster]
        else:
            raise TypeError('cluster must be None, "gpus" / a list of CUDA device ordinals (one process, several GPUs), '
                            'or "torch" (torch.distributed ranks, one GPU each)')
    torch_cluster = cluster is not None and devices is None
    lo, hi = shard_range(chains, rank, world)
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if torch_cluster else 0
    T = burnin - 1 + samples
    d_free = int((~skip_param).sum())
    U = 3 * d_free * rungs + ((rungs - 1) if (coupling_on and rungs > 1) else 0)
    replay = None
    if rng == "r":
        # the reference runs chains one after another on one stream: chain c owns uniforms [c*T*U, (c+1)*T*U)
        allu = rstream.runif(chains * T * U)
        replay = allu[lo * T * U: hi * T * U]
    # The kernel family (one thread per particle / one cluster per chain / the whole grid per run) and its reduction layout
    # are chosen by the library from the WHOLE run's size (layout_chains), not the shard's: the families associate the
    # likelihood sum differently (equal within 1e-10, not bitwise), so every shard -- and the same run on one GPU --
    # must use the same one to stay bit-identical.
    flags = 0
    spec = None  # more ranks than chains: this rank has no chains, but still enters every collective below
    if hi > lo:
        spec = engine.RunSpec(
            theta_min=tmin, theta_max=tmax, theta_init=init_mat.T[lo:hi], burnin=burnin, samples=samples,
            beta_raised=beta_raised, blocks=[list(b) for b in df_params["block"]], coupling_on=bool(coupling_on),
            save_hot_draws=bool(save_hot_draws), store_pt=bool(store_pt), target_acceptance=float(target_acceptance),
            rng_mode="replay" if rng == "r" else "philox", seed=0 if seed is None or rng == "r" else int(seed),
            chain_offset=lo, replay=replay, device=device, data=data_items, misc=misc_items,
            data_names=data_names, misc_names=misc_names, param_names=list(df_params["name"]), flags=flags,
            layout_chains=chains)

    # ---------- run (one native call for all chains of this rank; R/main.R:392-402 loops chains instead)
    phase_names = ("burn-in", "sampling phase")
    state = {"phase": -1}

    def _progress(user, phase, done, total):
        if progress is not None:
            return int(bool(progress(phase, done, total)))
        if not silent and not pb_markdown and phase != state["phase"]:
            state["phase"] = phase
            print(phase_names[phase], file=sys.stdout, flush=True)
        return 0

    if output not in ("full", "summary"):
        raise ValueError('output must be "full" or "summary"')
    if thin is not None:
        if output != "summary":
            raise ValueError('thin= needs output="summary" (with output="full" every draw comes back; subsample with sample_chains())')
        if int(thin) != thin or thin < 1:
            raise ValueError("thin must be a positive integer")
        thin = int(thin)
    if quantiles is not None and output != "summary":
        raise ValueError('quantiles= needs output="summary" (with output="full" the draws come back: use numpy / pandas quantile)')
    if output == "summary":
        raw, diagnostics = _run_summary(model, spec, _progress if not (silent and progress is None) else None, names=list(df_params["name"]),
                                        skip_param=skip_param, chains=chains, rank=rank, world=world, beta_raised=beta_raised, thin=thin,
                                        devices=devices, quantiles=quantiles, shape=(burnin, samples, rungs, d, bool(save_hot_draws), bool(store_pt)))
        if rank != 0:
            return None
        frame = pt = None
        if thin is not None and as_frames:
            frame, pt = _thinned_frames(raw, list(df_params["name"]), burnin, samples, thin, chains, rungs, bool(save_hot_draws))
        parameters = dict(data=data if save_data else None, df_params=df_params, loglike=loglike, logprior=logprior,
                          burnin=burnin, samples=samples, rungs=rungs, chains=chains, coupling_on=bool(coupling_on), alpha=alpha,
                          beta_manual=beta_manual)
        diagnostics["wall_seconds"] = time.perf_counter() - t_start
        return DrjacobyOutput(frame, pt, diagnostics, parameters, raw)
    cb = None if (silent and progress is None) else _progress
    if devices is not None:
        raw = engine.run_raw_multi(model, spec, None if devices == "all" else devices, cb)
    elif spec is None:
        raw = _empty_raw(burnin, samples, rungs, d, bool(save_hot_draws), bool(store_pt))
    else:
        raw = engine.run_raw(model, spec, cb)
    if torch_cluster:
        raw = gather_shards(raw, chains, rank, world)
        if rank != 0:
            return None
    if not silent:
        d_all = d
        shown = chains if chains <= 20 else 5  # the reference prints every chain; thousands of chains would flood
        if shown < chains:
            print(f"(acceptance rates of the first {shown} of {chains} chains)")
        for c in range(shown):  # src/main.cpp:147-150,194-197,253-257
            print(f"MCMC chain {c + 1}")
            print(f"burn-in\nacceptance rate: {_printed_rate(raw['accept_burnin'][c, -1], burnin, d_all)}%")
            print(f"sampling phase\nacceptance rate: {_printed_rate(raw['accept_sampling'][c, -1], samples, d_all)}%")
        print(f"total MCMC run-time: {float(raw['t_diff'].sum()):.3g} seconds")

    out = process_output(raw, df_params, skip_param, burnin, samples, rungs, chains, beta_raised, bool(save_hot_draws),
                         bool(store_pt), as_frames)
    parameters = dict(data=data if save_data else None, df_params=df_params, loglike=loglike, logprior=logprior,
                      burnin=burnin, samples=samples, rungs=rungs, chains=chains, coupling_on=bool(coupling_on), alpha=alpha,
                      beta_manual=beta_manual)
    res = DrjacobyOutput(out["output"], out["pt"], out["diagnostics"], parameters, raw)
    res.diagnostics["wall_seconds"] = time.perf_counter() - t_start
    return res


def _empty_raw(burnin, samples, rungs, d, save_hot_draws, store_pt):
    """The raw arrays of a shard with no chains (more ranks than chains): zero-length along the chain axis."""
    ptR, thR = (rungs if store_pt else 1), (rungs if save_hot_draws else 1)
    return dict(loglike_burnin=np.zeros((0, ptR, burnin)), logprior_burnin=np.zeros((0, ptR, burnin)),
                theta_burnin=np.zeros((0, thR, burnin, d)), loglike_sampling=np.zeros((0, ptR, samples)),
                logprior_sampling=np.zeros((0, ptR, samples)), theta_sampling=np.zeros((0, thR, samples, d)),
                mc_accept_burnin=np.zeros((0, max(rungs - 1, 0)), np.int32), mc_accept_sampling=np.zeros((0, max(rungs - 1, 0)), np.int32),
                accept_burnin=np.zeros((0, rungs), np.int32), accept_sampling=np.zeros((0, rungs), np.int32),
                t_diff=np.zeros(0), bw_final=np.zeros((0, rungs, d)))


def _thinned_frames(raw, names, burnin, samples, thin, chains, rungs, save_hot_draws):
    """The `output` (cold rung) and `pt` (all stored rungs) data frames from draws thinned on the device: stored
    iterations 0, thin, 2 thin, ... of each phase, numbered as the reference numbers them (1..burnin,
    burnin+1..burnin+samples; R/main.R:420-482)."""
    th_b, th_s = raw["theta_burnin"][:, -1], raw["theta_sampling"][:, -1]  # [chains, kept, d]
    it_b = 1 + thin * np.arange(th_b.shape[1])
    it_s = burnin + 1 + thin * np.arange(th_s.shape[1])
    n_it = it_b.size + it_s.size
    phase = np.concatenate([np.repeat("burnin", it_b.size), np.repeat("sampling", it_s.size)])
    its = np.concatenate([it_b, it_s])
    out = pd.DataFrame({"chain": np.repeat(np.arange(1, chains + 1), n_it), "phase": np.tile(phase, chains),
                        "iteration": np.tile(its, chains)})
    theta = np.concatenate([th_b, th_s], axis=1).reshape(chains * n_it, len(names))
    for k, nm in enumerate(names):
        out[nm] = theta[:, k]
    out["logprior"] = np.concatenate([raw["logprior_burnin"][:, -1], raw["logprior_sampling"][:, -1]], axis=1).ravel()
    out["loglikelihood"] = np.concatenate([raw["loglike_burnin"][:, -1], raw["loglike_sampling"][:, -1]], axis=1).ravel()
    ptR = raw["loglike_burnin"].shape[1]
    rung_ids = np.arange(1, rungs + 1) if ptR == rungs else np.array([rungs])
    pt = pd.DataFrame({"chain": np.repeat(np.arange(1, chains + 1), ptR * n_it), "rung": np.tile(np.repeat(rung_ids, n_it), chains),
                       "phase": np.tile(phase, chains * ptR), "iteration": np.tile(its, chains * ptR),
                       "logprior": np.concatenate([raw["logprior_burnin"], raw["logprior_sampling"]], axis=2).ravel(),
                       "loglikelihood": np.concatenate([raw["loglike_burnin"], raw["loglike_sampling"]], axis=2).ravel()})
    if save_hot_draws and ptR == rungs:
        th_all = np.concatenate([raw["theta_burnin"], raw["theta_sampling"]], axis=2).reshape(-1, len(names))
        for k, nm in enumerate(names):
            pt[nm] = th_all[:, k]
    return out, pt


def _advance_with_progress(sess, spec, progress):
    nb = spec.burnin - 1
    for phase, total in ((0, nb), (1, spec.samples)):
        step = max(1, (total + 99) // 100) if progress is not None else max(total, 1)
        done = 0
        while done < total:
            n = min(step, total - done)
            sess.advance(n)
            done += n
            if progress is not None and progress(None, phase, done + (1 if phase == 0 else 0), spec.burnin if phase == 0 else spec.samples):
                raise DrjacobyError(15, "interrupted by the progress callback")


def _run_summary(model, spec, progress, *, names, skip_param, chains, rank, world, beta_raised, thin=None, devices=None,
                 quantiles=None, shape=None):
    """output="summary": run in a session, reduce the draws on the device(s), bring back only the summaries and counters.

    One process (one GPU, or several through cluster="gpus"): ONE drj_session_summary / drj_multi_summary call -- the
    merge over the devices (per-chain moments, autocovariance sums about the global mean plus the lag products that
    straddle two devices' chain ranges, pooled deviance moments, radix-select histograms) happens inside the library.
    cluster="torch": each rank reduces its shard; the moments are gathered, the autocovariance sums added over the
    ranks plus the straddling lag products, so the result equals the single-GPU one.  A rank without chains (more
    ranks than chains) skips the native calls and still enters every collective."""
    burnin, samples, rungs, d, save_hot, store_pt = shape
    n_tot = chains * samples
    max_lag = min(n_tot - 1, int(np.floor(10 * np.log10(n_tot))))
    quant = None
    if world == 1:
        sess = engine.Session(model, spec, devices=devices)
        try:
            _advance_with_progress(sess, spec, progress)
            summ = sess.summary(max_lag=max_lag, quantiles=quantiles)
            raw = sess.download(draws=False) if thin is None else sess.download(draws=True, thin=thin)
        finally:
            sess.close()
        mean, var, ac = summ["chain_mean"], summ["chain_var"], summ["autocov_sum"]
        dic = float(summ["deviance_mean_var"][0] + 0.5 * summ["deviance_mean_var"][1]) if n_tot > 1 else np.nan
        quant = summ.get("quantiles")
    else:
        if quantiles is not None:
            raise ValueError('quantiles= is not available with cluster="torch"; use cluster="gpus" (one process, several GPUs)')
        import torch
        import torch.distributed as dist
        backend = dist.get_backend()
        devc = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        sess = engine.Session(model, spec) if spec is not None else None
        try:
            if sess is not None:
                _advance_with_progress(sess, spec, progress)
                first = sess.summary(max_lag=0, autocov=False)
                mean, var = first["chain_mean"], first["chain_var"]
                dev = np.array([[first["deviance_mean_var"][0], first["deviance_mean_var"][1], spec.chains * spec.samples]], dtype=np.float64)
            else:
                mean, var, dev = np.zeros((0, d)), np.zeros((0, d)), np.zeros((0, 3))
            parts = gather_shards({"m": mean, "v": var, "dev": dev}, chains, rank, world)
            ctr_t = torch.zeros(d, dtype=torch.float64, device=devc)
            if rank == 0:
                ctr_t += torch.from_numpy(parts["m"].mean(axis=0)).to(devc)
            dist.broadcast(ctr_t, src=0)
            center = ctr_t.cpu().numpy()
            if sess is not None:
                second = sess.summary(max_lag=max_lag, center=center, moments=False)
                ac, hd, tl = second["autocov_sum"], second["head"][None], second["tail"][None]
                raw = sess.download(draws=False) if thin is None else sess.download(draws=True, thin=thin)
            else:
                ac, hd, tl = np.zeros((d, max_lag + 1)), np.zeros((0, d, max(max_lag, 1))), np.zeros((0, d, max(max_lag, 1)))
                raw = _empty_raw(-(-burnin // (thin or 1)), -(-samples // (thin or 1)), rungs, d, save_hot, store_pt)
                if thin is None:
                    raw = {k: v for k, v in raw.items() if k not in engine._DRAW_KEYS}
            t = torch.from_numpy(np.ascontiguousarray(ac)).to(devc)
            dist.all_reduce(t)
            ac = t.cpu().numpy()
            # one gather for the counters and for each shard's first / last max_lag draws
            both = gather_shards(dict(raw, _head=hd, _tail=tl), chains, rank, world)
            if rank == 0:
                hd_all, tl_all = both.pop("_head"), both.pop("_tail")
                raw = both
                # lag products that straddle two ranks' shards: tail of shard k x head of shard k+1 (shards with chains only)
                for k in range(hd_all.shape[0] - 1):
                    if max_lag <= 0:
                        break
                    tlk, hdk = tl_all[k] - center[:, None], hd_all[k + 1] - center[:, None]   # [d, L]
                    for lag in range(1, max_lag + 1):
                        # x_j in the last `lag` draws of shard k pairs with x_{j+lag} among the first `lag` of shard k+1
                        ac[:, lag] += (tlk[:, max_lag - lag:] * hdk[:, :lag]).sum(axis=1)
            else:
                raw = None
        finally:
            if sess is not None:
                sess.close()
        if rank != 0:
            return None, None
        mean, var = parts["m"], parts["v"]
        # deviance over all ranks: pool the per-rank (mean, variance, count)
        dv = parts["dev"]
        n_all = dv[:, 2].sum()
        gm = (dv[:, 0] * dv[:, 2]).sum() / n_all
        m2 = (dv[:, 1] * (dv[:, 2] - 1) + dv[:, 2] * (dv[:, 0] - gm) ** 2).sum()
        dic = float(gm + 0.5 * m2 / (n_all - 1)) if n_all > 1 else np.nan
    diagnostics = {"run_time": pd.DataFrame({"chain": np.arange(1, chains + 1), "seconds": raw["t_diff"]})}
    if chains > 1 and samples > 1:
        diagnostics["rhat"] = pd.Series({nm: (np.nan if skip_param[k] else dg.gelman_rubin_from_moments(mean[:, k], var[:, k], samples))
                                         for k, nm in enumerate(names)})
    diagnostics["ess"] = pd.Series({nm: (np.nan if skip_param[k] else dg.ess_from_autocov(ac[k], n_tot)) for k, nm in enumerate(names)})
    diagnostics["rung_details"] = pd.DataFrame({"rung": np.arange(1, rungs + 1), "thermodynamic_power": beta_raised})
    if rungs > 1:
        link = np.tile(np.arange(1, rungs), chains)
        chain = np.repeat(np.arange(1, chains + 1), rungs - 1)
        diagnostics["mc_accept"] = pd.concat([
            pd.DataFrame({"link": link, "chain": chain, "phase": "burnin", "value": raw["mc_accept_burnin"].ravel() / burnin}),
            pd.DataFrame({"link": link, "chain": chain, "phase": "sampling", "value": raw["mc_accept_sampling"].ravel() / samples})],
            ignore_index=True)
    else:
        diagnostics["mc_accept"] = np.nan
    diagnostics["DIC_Gelman"] = dic
    diagnostics["posterior_mean"] = pd.Series({nm: float(mean[:, k].mean()) for k, nm in enumerate(names)})
    if quant is not None:
        probs = np.ascontiguousarray(quantiles, dtype=np.float64).ravel()
        diagnostics["quantiles"] = pd.DataFrame(quant.T, index=[f"{100 * p:g}%" for p in probs], columns=names)
    raw = dict(raw, chain_mean=mean, chain_var=var, autocov_sum=ac)
    return raw, diagnostics


def _printed_rate(acc, n, d):
    v = float(acc) / float(n * d) * 1000.0
    return np.floor(abs(v) + 0.5) / 10.0


def shard_range(chains: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous chain ranges, sizes differing by at most one."""
    base, rem = divmod(chains, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_PACK_LIMIT = 8 << 20  # arrays up to this many bytes per rank travel packed in one message


def gather_shards(raw: dict, chains: int, rank: int, world: int) -> dict | None:
    """End-of-run gather of every rank's chain slice to rank 0 (NCCL when the ranks own GPUs, gloo on
    CPU): the only collective of a run, never on the sampling hot path.  One metadata exchange for all
    arrays; the small ones (counters, per-chain summaries) are packed into a single message per rank, the
    large ones (draws) are sent one by one."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    keys = sorted(raw)
    arrs = {k: np.ascontiguousarray(raw[k]) for k in keys}
    meta = {k: (tuple(arrs[k].shape), arrs[k].dtype.str) for k in keys}
    all_meta = [None] * world
    dist.all_gather_object(all_meta, meta)

    def nbytes(m):
        return int(np.prod(m[0], dtype=np.int64)) * np.dtype(m[1]).itemsize

    small = [k for k in keys if max(nbytes(all_meta[r][k]) for r in range(world)) <= _PACK_LIMIT]
    large = [k for k in keys if k not in small]
    out = {} if rank == 0 else None
    if small:
        sizes = [sum(nbytes(all_meta[r][k]) for k in small) for r in range(world)]
        width = max(max(sizes), 1)
        mine = np.zeros(width, dtype=np.uint8)
        off = 0
        for k in small:
            b = arrs[k].view(np.uint8).reshape(-1)
            mine[off:off + b.size] = b
            off += b.size
        t = torch.from_numpy(mine).to(dev)
        if rank == 0:
            got = [torch.empty(width, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.gather(t, got, dst=0)
            host = [g.cpu().numpy() for g in got]
            offs = [0] * world
            for k in small:
                parts = []
                for r in range(world):
                    shp, dt = all_meta[r][k]
                    n = nbytes(all_meta[r][k])
                    parts.append(host[r][offs[r]:offs[r] + n].view(np.dtype(dt)).reshape(shp))
                    offs[r] += n
                out[k] = np.concatenate(parts, axis=0)
        else:
            dist.gather(t, dst=0)
    for k in large:
        t = torch.from_numpy(arrs[k]).to(dev)
        if rank == 0:
            parts = [t] + [torch.empty(all_meta[r][k][0], dtype=t.dtype, device=dev) for r in range(1, world)]
            for r in range(1, world):
                if parts[r].numel():
                    dist.recv(parts[r], src=r)
            out[k] = torch.cat(parts, dim=0).cpu().numpy()
        elif t.numel():
            dist.send(t, dst=0)
    return out


# ------------------------------------------------------------------------------------------------
# post-processing (R/main.R:411-554)
# ------------------------------------------------------------------------------------------------
def process_output(raw, df_params, skip_param, burnin, samples, rungs, chains, beta_raised, save_hot_draws, store_pt, as_frames):
    names = list(df_params["name"])
    d = len(names)
    cold_pt = -1  # last rung of the pt arrays (or the only one when store_pt is off)
    th_b, th_s = raw["theta_burnin"][:, -1], raw["theta_sampling"][:, -1]  # cold rung [chains, it, d]
    if not np.all(np.isfinite(th_b)) or not np.all(np.isfinite(th_s)):
        raise DrjacobyError(0, "output contains non-finite values")
    diagnostics: dict[str, Any] = {}
    diagnostics["run_time"] = pd.DataFrame({"chain": np.arange(1, chains + 1), "seconds": raw["t_diff"]})
    if chains > 1:
        rhat = {nm: (np.nan if skip_param[k] else dg.gelman_rubin(th_s[:, :, k]) if samples > 1 else np.nan)
                for k, nm in enumerate(names)}
        diagnostics["rhat"] = pd.Series(rhat)
    ess = {nm: (np.nan if skip_param[k] else dg.ess(th_s[:, :, k].reshape(-1))) for k, nm in enumerate(names)}
    diagnostics["ess"] = pd.Series(ess)
    diagnostics["rung_details"] = pd.DataFrame({"rung": np.arange(1, rungs + 1), "thermodynamic_power": beta_raised})
    if rungs > 1:
        link = np.tile(np.arange(1, rungs), chains)
        chain = np.repeat(np.arange(1, chains + 1), rungs - 1)
        diagnostics["mc_accept"] = pd.concat([
            pd.DataFrame({"link": link, "chain": chain, "phase": "burnin", "value": raw["mc_accept_burnin"].ravel() / burnin}),
            pd.DataFrame({"link": link, "chain": chain, "phase": "sampling", "value": raw["mc_accept_sampling"].ravel() / samples})],
            ignore_index=True)
    else:
        diagnostics["mc_accept"] = np.nan
    diagnostics["DIC_Gelman"] = dg.dic_gelman(raw["loglike_sampling"][:, cold_pt]) if samples > 1 else np.nan
    output = pt = None
    if as_frames:
        n_it = burnin + samples
        phase = np.tile(np.concatenate([np.repeat("burnin", burnin), np.repeat("sampling", samples)]), chains)
        iteration = np.tile(np.arange(1, n_it + 1), chains)
        chain = np.repeat(np.arange(1, chains + 1), n_it)
        theta = np.concatenate([th_b, th_s], axis=1).reshape(chains * n_it, d)
        output = pd.DataFrame({"chain": chain, "phase": phase, "iteration": iteration})
        for k, nm in enumerate(names):
            output[nm] = theta[:, k]
        output["logprior"] = np.concatenate([raw["logprior_burnin"][:, cold_pt], raw["logprior_sampling"][:, cold_pt]], axis=1).ravel()
        output["loglikelihood"] = np.concatenate([raw["loglike_burnin"][:, cold_pt], raw["loglike_sampling"][:, cold_pt]], axis=1).ravel()
        ptR = raw["loglike_burnin"].shape[1]
        rung_ids = np.arange(1, rungs + 1) if ptR == rungs else np.array([rungs])
        pt = pd.DataFrame({
            "chain": np.repeat(np.arange(1, chains + 1), ptR * n_it),
            "rung": np.tile(np.repeat(rung_ids, n_it), chains),
            "phase": np.tile(np.concatenate([np.repeat("burnin", burnin), np.repeat("sampling", samples)]), chains * ptR),
            "iteration": np.tile(np.arange(1, n_it + 1), chains * ptR),
            "logprior": np.concatenate([raw["logprior_burnin"], raw["logprior_sampling"]], axis=2).ravel(),
            "loglikelihood": np.concatenate([raw["loglike_burnin"], raw["loglike_sampling"]], axis=2).ravel()})
        if save_hot_draws and ptR == rungs:
            th_all = np.concatenate([raw["theta_burnin"], raw["theta_sampling"]], axis=2).reshape(-1, d)
            for k, nm in enumerate(names):
                pt[nm] = th_all[:, k]
    return {"output": output, "pt": pt, "diagnostics": diagnostics}
