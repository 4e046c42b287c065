"""r_shim/drjacoby_gpu_shim.c cannot be built against a real R here (no R in the image).  It is compiled against
the functional mock of the R C API in tests/r_mock (warnings are errors), linked with the product library, loaded,
and exercised as far as a machine without a GPU allows: registration as R_init_drjacobygpu does it, argument
checks turning into R errors, and -- no device -- the library's refusal arriving as an R error (no CPU fallback).
The full execution on the GPU is tests/test_gpu_r_shim.py."""
import os
import subprocess

import numpy as np
import pytest

import rshim_driver as rd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_against_the_header(tmp_path):
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-c",
           os.path.join(ROOT, "r_shim", "drjacoby_gpu_shim.c"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "r_mock"), "-o", str(tmp_path / "shim.o")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    syms = subprocess.check_output(["nm", str(tmp_path / "shim.o")], text=True)
    assert " T drjgpu_main" in syms and " T drjgpu_summary" in syms and " T R_init_drjacobygpu" in syms
    for needed in ("drj_run_mcmc_multi", "drj_multi_create", "drj_multi_summary", "drj_model_create", "drj_model_destroy"):
        assert f" U {needed}" in syms


def small_args(R, **kw):
    x = np.random.default_rng(1).normal(3, 2, 20)
    base = dict(data={"x": x}, misc={}, names=["mu", "sigma"], tmin=[-10, 0], tmax=[10, np.inf], blocks=[[1], [1]],
                beta_raised=[1.0], init_mat=np.array([[1.0, 2.0], [1.5, 2.5]]).T, burnin=10, samples=10, chains=2,
                model={"builtin": "example"})
    base.update(kw)
    return rd.build_args(R, **base)


def test_shim_loads_registers_and_reports_errors_as_r_errors(tmp_path):
    import drjacoby_b200 as drj
    R = rd.RMock(rd.build(str(tmp_path)))
    assert R.L.rmock_dynamic_symbols() == 0                     # R_useDynamicSymbols(dll, FALSE), as src/RcppExports.cpp:31
    with pytest.raises(RuntimeError, match="not available for .Call"):
        R.dot_call("_drjacoby_main_cpp", small_args(R))         # only the registered routines are callable
    with pytest.raises(RuntimeError, match="args_params\\$model is missing"):
        bad = R.rlist({"args_params": R.rlist({"theta_min": R.real([0.0]), "chains": R.integer([1]), "rungs": R.integer([1]),
                                                "burnin": R.integer([5]), "samples": R.integer([5]), "n_block": R.integer([1]),
                                                "coupling_on": R.logical([True]), "save_hot_draws": R.logical([False]),
                                                "target_acceptance": R.real([0.44]), "theta_max": R.real([1.0]), "trans_type": R.real([3]),
                                                "skip_param": R.logical([False]), "block": R.rlist([R.real([1])]), "beta_raised": R.real([1.0]),
                                                "init_mat": R.real([0.5]), "seed": R.real([1]), "x": R.rlist({"x": R.real([1.0])})})})
        R.dot_call("drjgpu_main", bad)
    with pytest.raises(RuntimeError, match="must be numeric"):
        R.dot_call("drjgpu_main", _with_string_data(R))
    with pytest.raises(RuntimeError, match="unknown built-in model"):
        R.dot_call("drjgpu_main", small_args(R, model={"builtin": "no_such_model"}))
    if drj.device_count() == 0:
        with pytest.raises(RuntimeError, match="CUDA"):         # no device: the library refuses, the shim raises an R error
            R.dot_call("drjgpu_main", small_args(R))
        with pytest.raises(RuntimeError, match="CUDA"):
            R.dot_call("drjgpu_summary", small_args(R, max_lag=5))
    assert R.L.rmock_protect_balance() == 0
    R.L.rmock_reset()


def _with_string_data(R):
    args = small_args(R)
    # data$x replaced by a character vector: the shim must refuse it with an R error
    p = R.L.rmock_elt(args, 0)
    R.L.rmock_list_set(p, 0, R.rlist({"x": R.string("a")}))
    return args
