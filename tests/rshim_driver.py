"""Drives r_shim/drjacoby_gpu_shim.c end to end WITHOUT R: the shim is compiled against the functional mock of the R
API in tests/r_mock (this image has no R) and linked with libdrjacoby_b200.so; this module builds the `args` list the
way R/main.R:348-386 does, performs the .Call by registered name, and walks the returned R object the way
R/main.R:411-482 does.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 0, 10, 13, 14, 16, 19


def build(tmpdir: str) -> str:
    """gcc: shim + mock -> one shared object, linked against the product library (rpath)."""
    pkg = os.path.join(ROOT, "drjacoby_b200")
    so = os.path.join(tmpdir, "drjacobygpu_mock.so")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-shared", "-fPIC",
           os.path.join(ROOT, "r_shim", "drjacoby_gpu_shim.c"), os.path.join(ROOT, "tests", "r_mock", "rmock.c"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "r_mock"),
           "-L", pkg, "-l:libdrjacoby_b200.so", f"-Wl,-rpath,{pkg}", "-o", so]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return so


class RMock:
    def __init__(self, so_path: str):
        L = self.L = C.CDLL(so_path)
        S = C.c_void_p
        for name, res, args in [
            ("rmock_reset", None, []), ("rmock_live_objects", C.c_size_t, []), ("rmock_protect_balance", C.c_int, []),
            ("rmock_interrupt_after", None, [C.c_long]), ("rmock_dot_call", S, [C.c_char_p, S, C.c_char_p, C.c_int]),
            ("rmock_dynamic_symbols", C.c_int, []),
            ("rmock_real", S, [C.POINTER(C.c_double), C.c_long]), ("rmock_int", S, [C.POINTER(C.c_int), C.c_long]),
            ("rmock_lgl", S, [C.POINTER(C.c_int), C.c_long]), ("rmock_matrix", S, [C.POINTER(C.c_double), C.c_int, C.c_int]),
            ("rmock_str", S, [C.c_char_p]), ("rmock_list", S, [C.c_long]), ("rmock_nil", S, []),
            ("rmock_set_names", None, [S, C.POINTER(C.c_char_p)]), ("rmock_list_set", None, [S, C.c_long, S]),
            ("rmock_type", C.c_int, [S]), ("rmock_length", C.c_long, [S]), ("rmock_elt", S, [S, C.c_long]),
            ("rmock_data", C.c_void_p, [S]), ("rmock_name", C.c_char_p, [S, C.c_long]), ("rmock_nrow", C.c_int, [S]),
            ("rmock_ncol", C.c_int, [S]), ("R_init_drjacobygpu", None, [C.c_void_p]),
        ]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L.R_init_drjacobygpu(None)  # what dyn.load() does: the routines get registered

    # ---- R objects from Python values
    def real(self, v):
        a = np.ascontiguousarray(v, dtype=np.float64).ravel()
        return self.L.rmock_real(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)

    def integer(self, v):
        a = np.ascontiguousarray(v, dtype=np.int32).ravel()
        return self.L.rmock_int(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)

    def logical(self, v):
        a = np.ascontiguousarray(np.asarray(v, dtype=bool), dtype=np.int32).ravel()
        return self.L.rmock_lgl(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)

    def matrix(self, m):  # numpy [nrow, ncol] -> column-major R matrix
        m = np.asarray(m, dtype=np.float64)
        a = np.ascontiguousarray(m.T).ravel()
        return self.L.rmock_matrix(a.ctypes.data_as(C.POINTER(C.c_double)), m.shape[0], m.shape[1])

    def string(self, s):
        return self.L.rmock_str(str(s).encode())

    def named(self, sexp, names):
        arr = (C.c_char_p * len(names))(*[str(n).encode() for n in names])
        self.L.rmock_set_names(sexp, arr)
        return sexp

    def rlist(self, items: dict | list):
        vals = list(items.values()) if isinstance(items, dict) else list(items)
        l = self.L.rmock_list(len(vals))
        for i, v in enumerate(vals):
            self.L.rmock_list_set(l, i, v if v is not None else self.L.rmock_nil())
        if isinstance(items, dict) and items:
            self.named(l, list(items.keys()))
        return l

    # ---- R objects back to Python
    def to_py(self, s):
        t, n = self.L.rmock_type(s), self.L.rmock_length(s)
        if t == NILSXP:
            return None
        if t in (REALSXP, INTSXP, LGLSXP):
            ct = C.c_double if t == REALSXP else C.c_int
            a = np.ctypeslib.as_array(C.cast(self.L.rmock_data(s), C.POINTER(ct)), shape=(n,)).copy() if n else np.zeros(0, ct)
            nr = self.L.rmock_nrow(s)
            return a.reshape(self.L.rmock_ncol(s), nr).T if nr >= 0 else a  # column-major -> numpy [nrow, ncol]
        if t == VECSXP:
            items = [self.to_py(self.L.rmock_elt(s, i)) for i in range(n)]
            names = [self.L.rmock_name(s, i).decode() for i in range(n)]
            return dict(zip(names, items)) if any(names) else items
        raise TypeError(f"unexpected R type {t}")

    def dot_call(self, name: str, arg):
        buf = C.create_string_buffer(1024)
        out = self.L.rmock_dot_call(name.encode(), arg, buf, 1024)
        if not out:
            raise RuntimeError(buf.value.decode(errors="replace"))
        return out


def build_args(R: RMock, *, data: dict, misc: dict, names, tmin, tmax, blocks, beta_raised, init_mat, burnin, samples, chains,
               coupling_on=True, save_hot_draws=False, target_acceptance=0.44, model=None, seed=1, replay=None, devices=None, **extra):
    """args$args_params as R/main.R:348-366 builds it (+ the GPU fields run_chains_gpu() adds, r_shim/run_mcmc_gpu.R)."""
    tmin, tmax = np.asarray(tmin, float), np.asarray(tmax, float)
    misc = dict(misc, block=-1.0)                                                # R/main.R:342
    p = {
        "x": R.rlist({k: R.real(v) for k, v in data.items()}),
        "misc": R.rlist({k: R.real(v) for k, v in misc.items()}),
        "loglike_use_cpp": R.logical([True]), "logprior_use_cpp": R.logical([True]),
        "theta_min": R.named(R.real(tmin), names), "theta_max": R.real(tmax),
        "block": R.rlist([R.real(b) for b in blocks]),                           # df_params$block: list of numeric vectors
        "n_block": R.real([max(max(b) for b in blocks)]),                        # max(unlist(...)) is a double in R
        "trans_type": R.real(2 * np.isfinite(tmin) + np.isfinite(tmax)),         # R/main.R:293, a double vector
        "skip_param": R.logical(tmin == tmax),
        "burnin": R.real([burnin]), "samples": R.real([samples]), "rungs": R.integer([len(beta_raised)]),  # defaults 1e3 / 1e4 are doubles
        "coupling_on": R.logical([coupling_on]), "beta_raised": R.real(beta_raised), "save_hot_draws": R.logical([save_hot_draws]),
        "pb_markdown": R.logical([False]), "silent": R.logical([True]), "target_acceptance": R.real([target_acceptance]),
        "init_mat": R.matrix(np.asarray(init_mat, float)),                       # d x chains
        "chains": R.integer([chains]), "seed": R.real([seed]),
        "replay": R.real(replay) if replay is not None else None,
        "model": R.rlist({k: R.string(v) for k, v in model.items()}),
        "devices": R.integer(devices) if devices is not None else None,
    }
    for k, v in extra.items():
        p[k] = R.real(np.atleast_1d(v)) if v is not None else None
    return R.rlist({"args_params": R.rlist(p), "args_functions": R.rlist({"loglike": R.string("loglike"), "logprior": R.string("logprior")})})


def chains_to_raw(per_chain: list, d: int) -> dict:
    """The list-of-lists of src/main.cpp:268-276 -> the chain-major arrays of drj_output (what R/main.R:420-476 reshapes)."""
    def mat(key):
        return np.array([[np.asarray(r) for r in ch[key]] for ch in per_chain], dtype=np.float64)

    def theta(key):
        return np.array([[[np.asarray(t) for t in r] for r in ch[key]] for ch in per_chain], dtype=np.float64).reshape(
            len(per_chain), len(per_chain[0][key]), -1, d)

    return dict(loglike_burnin=mat("loglike_burnin"), logprior_burnin=mat("logprior_burnin"), theta_burnin=theta("theta_burnin"),
                loglike_sampling=mat("loglike_sampling"), logprior_sampling=mat("logprior_sampling"), theta_sampling=theta("theta_sampling"),
                mc_accept_burnin=np.array([ch["mc_accept_burnin"] for ch in per_chain], dtype=np.int32).reshape(len(per_chain), -1),
                mc_accept_sampling=np.array([ch["mc_accept_sampling"] for ch in per_chain], dtype=np.int32).reshape(len(per_chain), -1),
                t_diff=np.array([ch["t_diff"][0] for ch in per_chain]))
