"""The C-ABI shared library: loads without a GPU, exports every symbol include/drjacoby_b200.h
declares, refuses to compute without a device (no CPU fallback anywhere in the product)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "drjacoby_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:int|void|double|int64_t|const char\*)\s+(drj_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_and_library_agree():
    from drjacoby_b200 import _lib
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/drjacoby_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding and header list different entry points"
    assert L.drj_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_the_header():
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess
    import tempfile
    from drjacoby_b200 import _lib
    src = ('#include <stdio.h>\n#include "drjacoby_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu ", sizeof(drj_config), '
           'sizeof(drj_list), sizeof(drj_model_desc), sizeof(drj_output), sizeof(drj_error));'
           'printf("%zu\\n", sizeof(drj_summary));return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_lib.DrjConfig), C.sizeof(_lib.DrjList), C.sizeof(_lib.DrjModelDesc),
                     C.sizeof(_lib.DrjOutput), C.sizeof(_lib.DrjError), C.sizeof(_lib.DrjSummary)]


def test_builtin_models_cover_the_reference_model_files():
    import drjacoby_b200 as drj
    have = set(drj.builtin_models())
    # inst/extdata/*.cpp, inst/extdata/checks/*.cpp, tests/testthat/test_input_files/*.cpp, getting_model_fits.Rmd
    for name in ("example", "coupling", "blocks", "doublewell", "multilevel", "normal_mu", "returnprior", "testfile",
                 "logistic"):
        assert name in have


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import drjacoby_b200 as drj
    if drj.device_count() > 0:
        pytest.skip("a GPU is present")
    m = drj.Model(builtin="example")
    spec = drj.RunSpec(theta_min=[-10, 0], theta_max=[10, np.inf], theta_init=[[1, 1]], burnin=10, samples=10,
                       data=[[1.0, 2.0, 3.0]])
    with pytest.raises(drj.DrjacobyError) as e:
        drj.run_raw(m, spec)
    assert e.value.code == 13
    with pytest.raises(drj.DrjacobyError):
        drj.fp64_peak_tflops(0)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under drjacoby_b200/ may reference it."""
    pkg = os.path.join(ROOT, "drjacoby_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "drj_oracle" not in text and "import oracle" not in text \
                    and "from oracle" not in text, f


def test_argument_validation_happens_before_cuda():
    import drjacoby_b200 as drj
    m = drj.Model(builtin="example")
    bad = drj.RunSpec(theta_min=[-10, 0], theta_max=[10, np.inf], theta_init=[[1, 1]], burnin=10, samples=10,
                      beta_raised=[0.5, 0.2, 1.0], data=[[1.0]])
    with pytest.raises(drj.DrjacobyError) as e:
        drj.run_raw(m, bad)
    assert e.value.code == 10 and "increasing" in str(e.value)
    wrong = drj.RunSpec(theta_min=[-10], theta_max=[10], theta_init=[[1]], burnin=10, samples=10, data=[[1.0]])
    with pytest.raises(drj.DrjacobyError) as e:
        drj.run_raw(m, wrong)
    assert e.value.code == 11
    with pytest.raises(drj.DrjacobyError) as e:
        drj.Model(builtin="no_such_model").handle(2, 1, 0)
    assert e.value.code == 11
