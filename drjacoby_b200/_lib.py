"""ctypes binding of libdrjacoby_b200.so (the C ABI in include/drjacoby_b200.h).

This is the only way Python reaches the sampler: there is no Python / CPU implementation behind
it.  If the shared library is missing the import of this module fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libdrjacoby_b200.so")

ABI_VERSION = 2

# status codes (include/drjacoby_b200.h)
OK = 0
ERR_INIT_NEGINF, ERR_LOGLIKE_BAD, ERR_LOGPRIOR_BAD, ERR_TRANS_TYPE = 1, 2, 3, 4
ERR_ARG, ERR_MODEL, ERR_COMPILE, ERR_CUDA, ERR_MEMORY, ERR_INTERRUPT = 10, 11, 12, 13, 14, 15
RNG_PHILOX, RNG_REPLAY = 0, 1
FLAG_NO_TMA, FLAG_STREAM_KERNEL, FLAG_RESIDENT_KERNEL = 1, 2, 4
FLAG_TEAM_KERNEL, FLAG_NO_TEAM_KERNEL = 8, 16
FLAG_GRID_KERNEL, FLAG_NO_GRID_KERNEL = 32, 64
FLAG_LANES_KERNEL, FLAG_NO_LANES_KERNEL = 128, 256
FLAG_SPEC_KERNEL, FLAG_NO_SPEC_KERNEL = 512, 1024


class DrjList(C.Structure):
    _fields_ = [("n_items", C.c_int32), ("names", C.POINTER(C.c_char_p)), ("offsets", C.POINTER(C.c_int64)),
                ("lengths", C.POINTER(C.c_int64)), ("values", C.POINTER(C.c_double))]


class DrjConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("d", C.c_int32), ("n_block", C.c_int32),
        ("chains", C.c_int32), ("rungs", C.c_int32), ("burnin", C.c_int32), ("samples", C.c_int32),
        ("coupling_on", C.c_int32), ("save_hot_draws", C.c_int32), ("store_pt", C.c_int32),
        ("rng_mode", C.c_int32), ("target_acceptance", C.c_double),
        ("theta_min", C.POINTER(C.c_double)), ("theta_max", C.POINTER(C.c_double)),
        ("trans_type", C.POINTER(C.c_int32)), ("skip_param", C.POINTER(C.c_uint8)),
        ("block_ptr", C.POINTER(C.c_int32)), ("block_idx", C.POINTER(C.c_int32)),
        ("beta_raised", C.POINTER(C.c_double)), ("theta_init", C.POINTER(C.c_double)),
        ("seed", C.c_uint64), ("chain_offset", C.c_int64), ("replay_uniforms", C.POINTER(C.c_double)),
        ("iters_per_launch", C.c_int32), ("chains_per_cta", C.c_int32), ("flags", C.c_int32),
        ("layout_chains", C.c_int32),
    ]


class DrjModelDesc(C.Structure):
    _fields_ = [
        ("builtin", C.c_char_p), ("cuda_source", C.c_char_p), ("loglike_name", C.c_char_p),
        ("logprior_name", C.c_char_p), ("param_names", C.POINTER(C.c_char_p)),
        ("data_names", C.POINTER(C.c_char_p)), ("misc_names", C.POINTER(C.c_char_p)),
        ("n_data_items", C.c_int32), ("n_misc_items", C.c_int32), ("d", C.c_int32), ("n_block", C.c_int32),
    ]


class DrjOutput(C.Structure):
    _fields_ = [
        ("loglike_burnin", C.POINTER(C.c_double)), ("logprior_burnin", C.POINTER(C.c_double)),
        ("theta_burnin", C.POINTER(C.c_double)), ("loglike_sampling", C.POINTER(C.c_double)),
        ("logprior_sampling", C.POINTER(C.c_double)), ("theta_sampling", C.POINTER(C.c_double)),
        ("mc_accept_burnin", C.POINTER(C.c_int32)), ("mc_accept_sampling", C.POINTER(C.c_int32)),
        ("accept_burnin", C.POINTER(C.c_int32)), ("accept_sampling", C.POINTER(C.c_int32)),
        ("t_diff", C.POINTER(C.c_double)), ("bw_final", C.POINTER(C.c_double)),
    ]


class DrjSummary(C.Structure):
    _fields_ = [("max_lag", C.c_int32), ("n_quantiles", C.c_int32), ("center", C.POINTER(C.c_double)),
                ("chain_mean", C.POINTER(C.c_double)), ("chain_var", C.POINTER(C.c_double)),
                ("autocov_sum", C.POINTER(C.c_double)), ("deviance_mean_var", C.POINTER(C.c_double)),
                ("head", C.POINTER(C.c_double)), ("tail", C.POINTER(C.c_double)),
                ("quantile_probs", C.POINTER(C.c_double)), ("quantiles", C.POINTER(C.c_double))]


class DrjError(C.Structure):
    _fields_ = [("code", C.c_int32), ("rung", C.c_int32), ("param", C.c_int32), ("iteration", C.c_int32),
                ("chain", C.c_int64), ("theta", C.c_double * 64), ("message", C.c_char * 512)]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int)

# every symbol include/drjacoby_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "drj_abi_version": (C.c_int, []),
    "drj_device_count": (C.c_int, []),
    "drj_builtin_model_count": (C.c_int, []),
    "drj_builtin_model_name": (C.c_char_p, [C.c_int]),
    "drj_model_create": (C.c_int, [C.POINTER(DrjModelDesc), C.c_int, C.POINTER(C.c_void_p), C.POINTER(DrjError)]),
    "drj_model_destroy": (None, [C.c_void_p]),
    "drj_model_compile_seconds": (C.c_double, [C.c_void_p]),
    "drj_run_mcmc": (C.c_int, [C.POINTER(DrjConfig), C.c_void_p, C.POINTER(DrjList), C.POINTER(DrjList),
                               C.POINTER(DrjOutput), PROGRESS_FN, C.c_void_p, C.POINTER(DrjError)]),
    "drj_session_create": (C.c_int, [C.POINTER(DrjConfig), C.c_void_p, C.POINTER(DrjList), C.POINTER(DrjList),
                                     C.POINTER(C.c_void_p), C.POINTER(DrjError)]),
    "drj_session_advance": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(DrjError)]),
    "drj_session_iterations_done": (C.c_int, [C.c_void_p]),
    "drj_session_download": (C.c_int, [C.c_void_p, C.POINTER(DrjOutput), C.POINTER(DrjError)]),
    "drj_release_cached_memory": (C.c_int64, []),
    "drj_session_download_thinned": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DrjOutput), C.POINTER(DrjError)]),
    "drj_session_summary": (C.c_int, [C.c_void_p, C.POINTER(DrjSummary), C.POINTER(DrjError)]),
    "drj_session_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "drj_session_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "drj_session_destroy": (None, [C.c_void_p]),
    "drj_fp64_peak_tflops": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(DrjError)]),
    "drj_run_mcmc_multi": (C.c_int, [C.POINTER(DrjConfig), C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.POINTER(DrjList),
                                     C.POINTER(DrjList), C.POINTER(DrjOutput), PROGRESS_FN, C.c_void_p, C.POINTER(DrjError)]),
    "drj_multi_create": (C.c_int, [C.POINTER(DrjConfig), C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.POINTER(DrjList),
                                   C.POINTER(DrjList), C.POINTER(C.c_void_p), C.POINTER(DrjError)]),
    "drj_multi_device_count": (C.c_int, [C.c_void_p]),
    "drj_multi_advance": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(DrjError)]),
    "drj_multi_download_thinned": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DrjOutput), C.POINTER(DrjError)]),
    "drj_multi_summary": (C.c_int, [C.c_void_p, C.POINTER(DrjSummary), C.POINTER(DrjError)]),
    "drj_multi_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "drj_multi_destroy": (None, [C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python drjacoby_b200/csrc/build.py` "
                "(nvcc, sm_100a). drjacoby_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.drj_abi_version() != ABI_VERSION:
            raise ImportError("libdrjacoby_b200.so ABI version mismatch: rebuild the library")
        _lib = L
    return _lib


class DrjacobyError(RuntimeError):
    """Raised where the reference calls Rcpp::stop() (src/Particle.h:75-78,121-128) or for bad arguments."""

    def __init__(self, code: int, message: str, chain=-1, rung=-1, param=-1, iteration=-1, theta=None):
        super().__init__(message)
        self.code, self.chain, self.rung, self.param, self.iteration, self.theta = (
            code, chain, rung, param, iteration, theta)


def raise_for(rc: int, err: DrjError, d: int = 0):
    if rc != OK:
        raise DrjacobyError(rc, err.message.decode(errors="replace"), err.chain, err.rung, err.param,
                            err.iteration, np.array(err.theta[:min(d, 64)]))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def make_list(items, names=None):
    """items: sequence of numeric arrays -> (DrjList, keepalive)."""
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel()) for a in items]
    lengths = np.array([a.size for a in arrs], dtype=np.int64)
    offsets = (np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64) if arrs else np.zeros(0, np.int64))
    if len(arrs) == 1 and arrs[0].size > 0:
        values = arrs[0]  # no copy: a caller's pinned buffer goes to the device as it is
    else:
        values = np.ascontiguousarray(np.concatenate(arrs) if arrs and lengths.sum() > 0 else np.zeros(1), dtype=np.float64)
    cnames = None
    if names is not None and len(arrs):
        cnames = (C.c_char_p * len(arrs))(*[str(n).encode() for n in names])
    l = DrjList(len(arrs), cnames, offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                lengths.ctypes.data_as(C.POINTER(C.c_int64)), dptr(values))
    return l, (arrs, lengths, offsets, values, cnames)
