"""ctypes binding of the CPU oracle (oracle/_build/libdrj_oracle.so).

ORACLE = TEST INFRASTRUCTURE (see oracle/drj_oracle.h).  Imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  The product
package drjacoby_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdrj_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class OraList(C.Structure):
    _fields_ = [("n_items", C.c_int32), ("offsets", C.POINTER(C.c_int64)),
                ("lengths", C.POINTER(C.c_int64)), ("values", C.POINTER(C.c_double))]


class OraRmt(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("mti", C.c_int32)]


class OraConfig(C.Structure):
    _fields_ = [
        ("d", C.c_int32), ("n_block", C.c_int32), ("chains", C.c_int32), ("rungs", C.c_int32),
        ("burnin", C.c_int32), ("samples", C.c_int32), ("coupling_on", C.c_int32),
        ("save_hot_draws", C.c_int32), ("target_acceptance", C.c_double),
        ("theta_min", C.POINTER(C.c_double)), ("theta_max", C.POINTER(C.c_double)),
        ("trans_type", C.POINTER(C.c_int32)), ("skip_param", C.POINTER(C.c_uint8)),
        ("block_ptr", C.POINTER(C.c_int32)), ("block_idx", C.POINTER(C.c_int32)),
        ("beta_raised", C.POINTER(C.c_double)), ("theta_init", C.POINTER(C.c_double)),
        ("rng_mode", C.c_int32), ("n_threads", C.c_int32), ("seed", C.c_uint64),
        ("chain_offset", C.c_int64), ("replay_uniforms", C.POINTER(C.c_double)),
        ("rmt", C.POINTER(OraRmt)),
    ]


class OraOutput(C.Structure):
    _fields_ = [
        ("loglike_burnin", C.POINTER(C.c_double)), ("logprior_burnin", C.POINTER(C.c_double)),
        ("theta_burnin", C.POINTER(C.c_double)), ("loglike_sampling", C.POINTER(C.c_double)),
        ("logprior_sampling", C.POINTER(C.c_double)), ("theta_sampling", C.POINTER(C.c_double)),
        ("mc_accept_burnin", C.POINTER(C.c_int32)), ("mc_accept_sampling", C.POINTER(C.c_int32)),
        ("accept_burnin", C.POINTER(C.c_int32)), ("accept_sampling", C.POINTER(C.c_int32)),
        ("t_diff", C.POINTER(C.c_double)), ("bw_final", C.POINTER(C.c_double)),
    ]


class OraError(C.Structure):
    _fields_ = [("code", C.c_int32), ("rung", C.c_int32), ("param", C.c_int32),
                ("iteration", C.c_int32), ("chain", C.c_int64), ("theta", C.c_double * 64),
                ("message", C.c_char * 256)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ora_run_mcmc.restype = C.c_int
        L.ora_run_mcmc.argtypes = [C.POINTER(OraConfig), C.c_char_p, C.POINTER(OraList),
                                   C.POINTER(OraList), C.POINTER(OraOutput), C.POINTER(OraError)]
        L.ora_model_loglike.restype = C.c_double
        L.ora_model_loglike.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(OraList),
                                        C.POINTER(OraList), C.c_int]
        L.ora_model_logprior.restype = C.c_double
        L.ora_model_logprior.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(OraList)]
        L.ora_rmt_set_seed.argtypes = [C.POINTER(OraRmt), C.c_uint32]
        L.ora_rmt_unif_rand.restype = C.c_double
        L.ora_rmt_unif_rand.argtypes = [C.POINTER(OraRmt)]
        L.ora_rmt_norm_rand.restype = C.c_double
        L.ora_rmt_norm_rand.argtypes = [C.POINTER(OraRmt)]
        L.ora_rmt_runif.argtypes = [C.POINTER(OraRmt), C.c_int, C.POINTER(C.c_double)]
        L.ora_rmt_rnorm.argtypes = [C.POINTER(OraRmt), C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.ora_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.ora_philox_uniforms.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.POINTER(C.c_double)]
        L.ora_qnorm.restype = C.c_double
        L.ora_qnorm.argtypes = [C.c_double]
        for name, nargs in [("ora_dnorm", 3), ("ora_dlnorm", 3), ("ora_dgamma", 3), ("ora_dbeta", 3),
                            ("ora_dpois", 2), ("ora_dunif", 3), ("ora_dexp", 2), ("ora_dbinom", 3)]:
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double] * nargs + [C.c_int]
        L.ora_stirlerr.restype = C.c_double
        L.ora_stirlerr.argtypes = [C.c_double]
        L.ora_bd0.restype = C.c_double
        L.ora_bd0.argtypes = [C.c_double, C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RStream:
    """R's default RNG after set.seed(seed): runif / rnorm with R's consumption order."""

    def __init__(self, seed: int):
        self.state = OraRmt()
        lib().ora_rmt_set_seed(C.byref(self.state), C.c_uint32(seed & 0xFFFFFFFF))

    def runif(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        lib().ora_rmt_runif(C.byref(self.state), n, _dp(out))
        return out

    def rnorm(self, n: int, mean: float = 0.0, sd: float = 1.0) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        lib().ora_rmt_rnorm(C.byref(self.state), n, mean, sd, _dp(out))
        return out


def make_list(items) -> tuple[OraList, list]:
    """items: sequence of 1-D arrays (the values of the reference's named list, in order)."""
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel()) for a in items]
    lengths = np.array([a.size for a in arrs], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64) if arrs else np.zeros(0, np.int64)
    values = np.concatenate(arrs) if arrs and lengths.sum() > 0 else np.zeros(1, np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    l = OraList(len(arrs), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                lengths.ctypes.data_as(C.POINTER(C.c_int64)), _dp(values))
    return l, [offsets, lengths, values]


class OracleError(RuntimeError):
    def __init__(self, code, message, chain, rung, param, iteration, theta):
        super().__init__(message)
        self.code, self.chain, self.rung, self.param, self.iteration, self.theta = (
            code, chain, rung, param, iteration, theta)


@dataclass
class OracleResult:
    loglike_burnin: np.ndarray
    logprior_burnin: np.ndarray
    theta_burnin: np.ndarray
    loglike_sampling: np.ndarray
    logprior_sampling: np.ndarray
    theta_sampling: np.ndarray
    mc_accept_burnin: np.ndarray
    mc_accept_sampling: np.ndarray
    accept_burnin: np.ndarray
    accept_sampling: np.ndarray
    t_diff: np.ndarray
    bw_final: np.ndarray
    extra: dict = field(default_factory=dict)


def trans_type_of(theta_min, theta_max):
    """R/main.R:293"""
    return (2 * np.isfinite(theta_min) + np.isfinite(theta_max)).astype(np.int32)


def run(model: str, data_items, misc_items, theta_min, theta_max, theta_init, *, burnin, samples,
        beta_raised=(1.0,), blocks=None, coupling_on=True, save_hot_draws=False,
        target_acceptance=0.44, rng_mode="philox", seed=1, chain_offset=0, replay=None,
        rstream: RStream | None = None, n_threads=1) -> OracleResult:
    L = lib()
    theta_min = np.ascontiguousarray(theta_min, dtype=np.float64)
    theta_max = np.ascontiguousarray(theta_max, dtype=np.float64)
    d = theta_min.size
    theta_init = np.ascontiguousarray(np.asarray(theta_init, dtype=np.float64).reshape(-1, d))
    chains = theta_init.shape[0]
    beta = np.ascontiguousarray(beta_raised, dtype=np.float64)
    R = beta.size
    if blocks is None:
        blocks = [[1]] * d
    block_ptr = np.zeros(d + 1, dtype=np.int32)
    block_idx = []
    for i, b in enumerate(blocks):
        block_idx += list(b)
        block_ptr[i + 1] = len(block_idx)
    block_idx = np.array(block_idx, dtype=np.int32)
    n_block = int(block_idx.max())
    tt = trans_type_of(theta_min, theta_max)
    skip = (theta_min == theta_max).astype(np.uint8)
    theta_rungs = R if save_hot_draws else 1
    res = OracleResult(
        np.zeros((chains, R, burnin)), np.zeros((chains, R, burnin)),
        np.zeros((chains, theta_rungs, burnin, d)),
        np.zeros((chains, R, samples)), np.zeros((chains, R, samples)),
        np.zeros((chains, theta_rungs, samples, d)),
        np.zeros((chains, max(R - 1, 0)), np.int32), np.zeros((chains, max(R - 1, 0)), np.int32),
        np.zeros((chains, R), np.int32), np.zeros((chains, R), np.int32),
        np.zeros(chains), np.zeros((chains, R, d)))
    cfg = OraConfig()
    cfg.d, cfg.n_block, cfg.chains, cfg.rungs = d, n_block, chains, R
    cfg.burnin, cfg.samples = int(burnin), int(samples)
    cfg.coupling_on, cfg.save_hot_draws = int(coupling_on), int(save_hot_draws)
    cfg.target_acceptance = target_acceptance
    cfg.theta_min, cfg.theta_max = _dp(theta_min), _dp(theta_max)
    cfg.trans_type = tt.ctypes.data_as(C.POINTER(C.c_int32))
    cfg.skip_param = skip.ctypes.data_as(C.POINTER(C.c_uint8))
    cfg.block_ptr = block_ptr.ctypes.data_as(C.POINTER(C.c_int32))
    cfg.block_idx = block_idx.ctypes.data_as(C.POINTER(C.c_int32))
    cfg.beta_raised, cfg.theta_init = _dp(beta), _dp(theta_init)
    cfg.rng_mode = {"philox": 0, "replay": 1, "rmt": 2}[rng_mode]
    cfg.n_threads = n_threads
    cfg.seed, cfg.chain_offset = seed, chain_offset
    if rng_mode == "replay":
        replay = np.ascontiguousarray(replay, dtype=np.float64)
        d_free = int((skip == 0).sum())
        U = 3 * d_free * R + ((R - 1) if (coupling_on and R > 1) else 0)
        assert replay.size == chains * (burnin - 1 + samples) * U, (replay.size, chains, U)
        cfg.replay_uniforms = _dp(replay)
    if rng_mode == "rmt":
        cfg.rmt = C.pointer(rstream.state)
    dl, keep1 = make_list(data_items)
    ml, keep2 = make_list(misc_items)
    out = OraOutput(_dp(res.loglike_burnin), _dp(res.logprior_burnin), _dp(res.theta_burnin),
                    _dp(res.loglike_sampling), _dp(res.logprior_sampling), _dp(res.theta_sampling),
                    res.mc_accept_burnin.ctypes.data_as(C.POINTER(C.c_int32)),
                    res.mc_accept_sampling.ctypes.data_as(C.POINTER(C.c_int32)),
                    res.accept_burnin.ctypes.data_as(C.POINTER(C.c_int32)),
                    res.accept_sampling.ctypes.data_as(C.POINTER(C.c_int32)),
                    _dp(res.t_diff), _dp(res.bw_final))
    err = OraError()
    rc = L.ora_run_mcmc(C.byref(cfg), model.encode(), C.byref(dl), C.byref(ml), C.byref(out), C.byref(err))
    if rc != 0:
        raise OracleError(rc, err.message.decode(), err.chain, err.rung, err.param, err.iteration,
                          np.array(err.theta[:d]))
    res.extra = dict(trans_type=tt, skip=skip, d=d, R=R)
    return res
