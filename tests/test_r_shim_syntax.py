"""r_shim/drjacoby_gpu_shim.c cannot be built for real here (no R); type-check it against a minimal
mock of the R C API so that it stays consistent with include/drjacoby_b200.h."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_against_the_header(tmp_path):
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-c",
           os.path.join(ROOT, "r_shim", "drjacoby_gpu_shim.c"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "r_mock"), "-o", str(tmp_path / "shim.o")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    syms = subprocess.check_output(["nm", str(tmp_path / "shim.o")], text=True)
    assert " T drjgpu_main" in syms and " T R_init_drjacobygpu" in syms
    for needed in ("drj_run_mcmc", "drj_model_create", "drj_model_destroy"):
        assert f" U {needed}" in syms
