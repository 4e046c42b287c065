"""Thin object layer over the C ABI: models, run specification, sessions.

Mirrors the reference's native layer: `Model` is the pair of user functions handed to main_cpp
(src/main.cpp:17-80), `RunSpec` is System (src/System.h:10-53) for all chains of a run, `Session`
is the run_mcmc<T1,T2> loop (src/main.cpp:84-278) with the state resident on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import _lib as L


class Model:
    """A likelihood/prior pair: a built-in model or CUDA source compiled with NVRTC for sm_100a."""

    def __init__(self, *, builtin: str | None = None, source: str | None = None, loglike: str = "loglike",
                 logprior: str = "logprior"):
        if (builtin is None) == (source is None):
            raise ValueError("give exactly one of builtin= or source=")
        self.builtin, self.source, self.loglike, self.logprior = builtin, source, loglike, logprior
        self._handles: dict = {}

    def handle(self, d: int, n_block: int, device: int, param_names=None, data_names=None, misc_names=None):
        key = (d, n_block, device, tuple(param_names or ()), tuple(data_names or ()), tuple(misc_names or ()))
        if key in self._handles:
            return self._handles[key]
        lib = L.lib()
        desc = L.DrjModelDesc()
        keep = []

        def carr(names):
            if not names:
                return None
            arr = (C.c_char_p * len(names))(*[str(n).encode() for n in names])
            keep.append(arr)
            return arr

        desc.builtin = self.builtin.encode() if self.builtin else None
        desc.cuda_source = self.source.encode() if self.source else None
        desc.loglike_name, desc.logprior_name = self.loglike.encode(), self.logprior.encode()
        desc.param_names, desc.data_names, desc.misc_names = carr(param_names), carr(data_names), carr(misc_names)
        desc.n_data_items, desc.n_misc_items = len(data_names or ()), len(misc_names or ())
        desc.d, desc.n_block = d, n_block
        h = C.c_void_p()
        err = L.DrjError()
        rc = lib.drj_model_create(C.byref(desc), device, C.byref(h), C.byref(err))
        L.raise_for(rc, err)
        self._handles[key] = h
        return h

    def compile_seconds(self) -> float:
        return float(sum(L.lib().drj_model_compile_seconds(h) for h in self._handles.values()))

    def __del__(self):
        try:
            for h in self._handles.values():
                L.lib().drj_model_destroy(h)
        except Exception:
            pass


def builtin_models() -> list[str]:
    lib = L.lib()
    return sorted({lib.drj_builtin_model_name(i).decode() for i in range(lib.drj_builtin_model_count())})


def device_count() -> int:
    return int(L.lib().drj_device_count())


def trans_type_of(theta_min, theta_max) -> np.ndarray:
    """R/main.R:293: 0 none, 1 upper only, 2 lower only, 3 both."""
    return (2 * np.isfinite(theta_min) + np.isfinite(theta_max)).astype(np.int32)


@dataclass
class RunSpec:
    """Everything the native entry needs, already in the flat layout of drj_config."""
    theta_min: np.ndarray
    theta_max: np.ndarray
    theta_init: np.ndarray            # [chains, d]
    burnin: int
    samples: int
    beta_raised: np.ndarray = field(default_factory=lambda: np.array([1.0]))
    blocks: Sequence[Sequence[int]] | None = None
    coupling_on: bool = True
    save_hot_draws: bool = False
    store_pt: bool = True
    target_acceptance: float = 0.44
    rng_mode: str = "philox"          # "philox" | "replay"
    seed: int = 1
    chain_offset: int = 0
    replay: np.ndarray | None = None
    device: int = 0
    iters_per_launch: int = 0
    chains_per_cta: int = 0
    flags: int = 0
    layout_chains: int = 0            # chains of the whole run when this spec is a shard of it (kernel family / reduction layout)
    data: Sequence = ()
    misc: Sequence = ()
    data_names: Sequence[str] | None = None
    misc_names: Sequence[str] | None = None
    param_names: Sequence[str] | None = None

    def __post_init__(self):
        self.theta_min = np.ascontiguousarray(self.theta_min, dtype=np.float64).ravel()
        self.theta_max = np.ascontiguousarray(self.theta_max, dtype=np.float64).ravel()
        self.d = self.theta_min.size
        self.theta_init = np.ascontiguousarray(np.asarray(self.theta_init, dtype=np.float64).reshape(-1, self.d))
        self.chains = self.theta_init.shape[0]
        self.beta_raised = np.ascontiguousarray(self.beta_raised, dtype=np.float64).ravel()
        self.rungs = self.beta_raised.size
        blocks = self.blocks if self.blocks is not None else [[1]] * self.d
        self.block_ptr = np.zeros(self.d + 1, dtype=np.int32)
        idx: list[int] = []
        for i, b in enumerate(blocks):
            idx += [int(v) for v in np.atleast_1d(b)]
            self.block_ptr[i + 1] = len(idx)
        self.block_idx = np.array(idx, dtype=np.int32)
        self.n_block = int(self.block_idx.max())
        self.trans_type = trans_type_of(self.theta_min, self.theta_max)
        self.skip_param = (self.theta_min == self.theta_max).astype(np.uint8)
        self.d_free = int((self.skip_param == 0).sum())
        self.burnin, self.samples = int(self.burnin), int(self.samples)
        self.T = self.burnin - 1 + self.samples
        self.U = 3 * self.d_free * self.rungs + ((self.rungs - 1) if (self.coupling_on and self.rungs > 1) else 0)
        if self.rng_mode == "replay":
            self.replay = np.ascontiguousarray(self.replay, dtype=np.float64).ravel()
            if self.replay.size != self.chains * self.T * self.U:
                raise ValueError(f"replay needs chains*T*U = {self.chains * self.T * self.U} uniforms, got {self.replay.size}")

    # ---- C structs
    def c_config(self):
        c = L.DrjConfig()
        c.abi_version, c.device = L.ABI_VERSION, self.device
        c.d, c.n_block, c.chains, c.rungs = self.d, self.n_block, self.chains, self.rungs
        c.burnin, c.samples = self.burnin, self.samples
        c.coupling_on, c.save_hot_draws, c.store_pt = int(self.coupling_on), int(self.save_hot_draws), int(self.store_pt)
        c.rng_mode = L.RNG_REPLAY if self.rng_mode == "replay" else L.RNG_PHILOX
        c.target_acceptance = float(self.target_acceptance)
        c.theta_min, c.theta_max = L.dptr(self.theta_min), L.dptr(self.theta_max)
        c.trans_type = L.iptr(self.trans_type)
        c.skip_param = self.skip_param.ctypes.data_as(C.POINTER(C.c_uint8))
        c.block_ptr, c.block_idx = L.iptr(self.block_ptr), L.iptr(self.block_idx)
        c.beta_raised, c.theta_init = L.dptr(self.beta_raised), L.dptr(self.theta_init)
        c.seed, c.chain_offset = int(self.seed) & (2**64 - 1), int(self.chain_offset)
        c.replay_uniforms = L.dptr(self.replay) if self.rng_mode == "replay" else None
        c.iters_per_launch, c.chains_per_cta, c.flags = self.iters_per_launch, self.chains_per_cta, self.flags
        c.layout_chains = int(self.layout_chains)
        return c

    def alloc_outputs(self, thin: int = 1, draws: bool = True) -> dict:
        C_, R, d = self.chains, self.rungs, self.d
        ptR = R if self.store_pt else 1
        thR = R if self.save_hot_draws else 1
        nb, ns = -(-self.burnin // thin), -(-self.samples // thin)
        out = {}
        if draws:
            out = dict(
                loglike_burnin=np.zeros((C_, ptR, nb)), logprior_burnin=np.zeros((C_, ptR, nb)),
                theta_burnin=np.zeros((C_, thR, nb, d)),
                loglike_sampling=np.zeros((C_, ptR, ns)), logprior_sampling=np.zeros((C_, ptR, ns)),
                theta_sampling=np.zeros((C_, thR, ns, d)))
        return dict(
            out,
            mc_accept_burnin=np.zeros((C_, max(R - 1, 0)), np.int32),
            mc_accept_sampling=np.zeros((C_, max(R - 1, 0)), np.int32),
            accept_burnin=np.zeros((C_, R), np.int32), accept_sampling=np.zeros((C_, R), np.int32),
            t_diff=np.zeros(C_), bw_final=np.zeros((C_, R, d)))


_DRAW_KEYS = ("loglike_burnin", "logprior_burnin", "theta_burnin", "loglike_sampling", "logprior_sampling", "theta_sampling")


def _c_output(o: dict) -> L.DrjOutput:
    def dp(k):
        return L.dptr(o[k]) if k in o else None

    def ip(k):
        return L.iptr(o[k]) if k in o else None

    return L.DrjOutput(dp("loglike_burnin"), dp("logprior_burnin"), dp("theta_burnin"), dp("loglike_sampling"),
                       dp("logprior_sampling"), dp("theta_sampling"), ip("mc_accept_burnin"), ip("mc_accept_sampling"),
                       ip("accept_burnin"), ip("accept_sampling"), dp("t_diff"), dp("bw_final"))


def run_raw(model: Model, spec: RunSpec, progress=None) -> dict:
    """One call of the drop-in entry drj_run_mcmc for all chains of `spec`; returns the raw arrays."""
    lib = L.lib()
    h = model.handle(spec.d, spec.n_block, spec.device, spec.param_names, spec.data_names, spec.misc_names)
    cfg = spec.c_config()
    dl, keep1 = L.make_list(spec.data, spec.data_names)
    ml, keep2 = L.make_list(spec.misc, spec.misc_names)
    out = spec.alloc_outputs()
    cout = _c_output(out)
    err = L.DrjError()
    cb = L.PROGRESS_FN(progress) if progress is not None else L.PROGRESS_FN()
    rc = lib.drj_run_mcmc(C.byref(cfg), h, C.byref(dl), C.byref(ml), C.byref(cout), cb, None, C.byref(err))
    L.raise_for(rc, err, spec.d)
    return out


def _device_array(devices):
    """None = all visible devices."""
    if devices is None:
        return None, 0
    arr = (C.c_int32 * len(devices))(*[int(v) for v in devices])
    return arr, len(devices)


def run_raw_multi(model: Model, spec: RunSpec, devices=None, progress=None) -> dict:
    """One call of drj_run_mcmc_multi: all chains of `spec` split over `devices` (None = every visible GPU) inside
    the library, one host thread per device, outputs written straight into their chain slices."""
    lib = L.lib()
    dev0 = int(devices[0]) if devices else 0
    h = model.handle(spec.d, spec.n_block, dev0, spec.param_names, spec.data_names, spec.misc_names)
    cfg = spec.c_config()
    dl, keep1 = L.make_list(spec.data, spec.data_names)
    ml, keep2 = L.make_list(spec.misc, spec.misc_names)
    out = spec.alloc_outputs()
    cout = _c_output(out)
    err = L.DrjError()
    cb = L.PROGRESS_FN(progress) if progress is not None else L.PROGRESS_FN()
    arr, n = _device_array(devices)
    rc = lib.drj_run_mcmc_multi(C.byref(cfg), arr, n, h, C.byref(dl), C.byref(ml), C.byref(cout), cb, None, C.byref(err))
    L.raise_for(rc, err, spec.d)
    return out


class Session:
    """State resident in HBM between calls (drj_session_*); with `devices` (a list, or "all") the chains are
    split over several GPUs inside the library (drj_multi_*): same methods, merged results."""

    def __init__(self, model: Model, spec: RunSpec, devices=None):
        lib = L.lib()
        self.spec, self.model = spec, model
        self.multi = devices is not None
        if devices == "all":
            devices = None
        dev0 = (int(devices[0]) if devices else 0) if self.multi else spec.device
        h = model.handle(spec.d, spec.n_block, dev0, spec.param_names, spec.data_names, spec.misc_names)
        cfg = spec.c_config()
        dl, keep1 = L.make_list(spec.data, spec.data_names)
        ml, keep2 = L.make_list(spec.misc, spec.misc_names)
        self._h = C.c_void_p()
        err = L.DrjError()
        if self.multi:
            arr, n = _device_array(devices)
            rc = lib.drj_multi_create(C.byref(cfg), arr, n, h, C.byref(dl), C.byref(ml), C.byref(self._h), C.byref(err))
        else:
            rc = lib.drj_session_create(C.byref(cfg), h, C.byref(dl), C.byref(ml), C.byref(self._h), C.byref(err))
        L.raise_for(rc, err, spec.d)

    def n_devices(self) -> int:
        return int(L.lib().drj_multi_device_count(self._h)) if self.multi else 1

    def advance(self, n_iter: int) -> float:
        """Advance n_iter updates; returns the CUDA-event time (ms) of the sweep kernels (max over the devices)."""
        ms = C.c_float(0)
        err = L.DrjError()
        fn = L.lib().drj_multi_advance if self.multi else L.lib().drj_session_advance
        rc = fn(self._h, int(n_iter), C.byref(ms), C.byref(err))
        L.raise_for(rc, err, self.spec.d)
        return float(ms.value)

    def iterations_done(self) -> int:
        if self.multi:
            raise NotImplementedError("iterations_done() is a one-device query")
        return int(L.lib().drj_session_iterations_done(self._h))

    def launches(self) -> tuple[int, int]:
        a, b = C.c_int64(0), C.c_int64(0)
        (L.lib().drj_multi_launches if self.multi else L.lib().drj_session_launches)(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def timing(self) -> tuple[float, float]:
        """Cumulative CUDA-event ms on the session stream: (sweep kernels, sweep kernels + transposes)."""
        if self.multi:
            raise NotImplementedError("timing() is a one-device query; advance() returns the max over the devices")
        a, b = C.c_double(0), C.c_double(0)
        L.lib().drj_session_timing(self._h, C.byref(a), C.byref(b))
        return float(a.value), float(b.value)

    def summary(self, max_lag: int | None = None, center=None, moments: bool = True, autocov: bool = True, quantiles=None) -> dict:
        """On-device summaries of the cold-rung sampling draws (drj_session_summary): per-chain means and
        variances, autocovariance sums of the concatenated chains, deviance mean / variance.  Each part is a
        pass over the stored draws: `moments=False` / `autocov=False` skip the ones not needed (skipped parts
        come back as zeros)."""
        sp = self.spec
        n_tot = sp.chains * sp.samples
        if max_lag is None:
            max_lag = min(n_tot - 1, int(np.floor(10 * np.log10(n_tot))))  # stats::ar's order.max
        out = dict(chain_mean=np.zeros((sp.chains, sp.d)), chain_var=np.zeros((sp.chains, sp.d)),
                   autocov_sum=np.zeros((sp.d, max_lag + 1)), deviance_mean_var=np.zeros(2),
                   head=np.zeros((sp.d, max(max_lag, 1))), tail=np.zeros((sp.d, max(max_lag, 1))))
        sm = L.DrjSummary()
        sm.max_lag = int(max_lag)
        ctr = None
        if center is not None:
            ctr = np.ascontiguousarray(center, dtype=np.float64)
            sm.center = L.dptr(ctr)
        if moments:
            sm.chain_mean, sm.chain_var = L.dptr(out["chain_mean"]), L.dptr(out["chain_var"])
            sm.deviance_mean_var = L.dptr(out["deviance_mean_var"])
        if autocov:
            if center is None and not moments:
                raise ValueError("autocov without moments needs a center")
            sm.autocov_sum = L.dptr(out["autocov_sum"])
            sm.head, sm.tail = L.dptr(out["head"]), L.dptr(out["tail"])
        probs = None
        if quantiles is not None and len(quantiles):
            probs = np.ascontiguousarray(quantiles, dtype=np.float64).ravel()
            out["quantiles"] = np.zeros((sp.d, probs.size))
            sm.n_quantiles, sm.quantile_probs, sm.quantiles = int(probs.size), L.dptr(probs), L.dptr(out["quantiles"])
        err = L.DrjError()
        rc = (L.lib().drj_multi_summary if self.multi else L.lib().drj_session_summary)(self._h, C.byref(sm), C.byref(err))
        L.raise_for(rc, err, sp.d)
        out["max_lag"] = int(max_lag)
        return out

    def download(self, draws: bool = True, thin: int = 1) -> dict:
        """All outputs, or (draws=False) only the counters / run times / bandwidths: the stored draws stay
        on the device (summary mode for runs whose draws are too large to move).  thin > 1: only the stored
        iterations 0, thin, 2 thin, ... of every series come back (gathered on the device)."""
        thin = int(thin)
        if thin < 1:
            raise ValueError("thin must be >= 1")
        out = self.spec.alloc_outputs(thin if draws else 1, draws)
        cout = _c_output(out)
        err = L.DrjError()
        fn = L.lib().drj_multi_download_thinned if self.multi else L.lib().drj_session_download_thinned
        rc = fn(self._h, thin, C.byref(cout), C.byref(err))
        L.raise_for(rc, err, self.spec.d)
        return out

    def close(self):
        if self._h:
            (L.lib().drj_multi_destroy if self.multi else L.lib().drj_session_destroy)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def release_cached_memory() -> int:
    """Give the device buffers the library keeps for reuse (drj_host.cu, memory pool) back to the driver."""
    return int(L.lib().drj_release_cached_memory())


def fp64_peak_tflops(device: int = 0) -> float:
    v = C.c_double(0)
    err = L.DrjError()
    rc = L.lib().drj_fp64_peak_tflops(device, C.byref(v), C.byref(err))
    L.raise_for(rc, err)
    return float(v.value)
