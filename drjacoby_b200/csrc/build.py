"""Build libdrjacoby_b200.so for sm_100a with nvcc (cross-compiles without a GPU).

Also embeds the device headers as string literals so that the library can hand them to NVRTC when
a user model is compiled at run time.  Usage: python drjacoby_b200/csrc/build.py [--force] [-v]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libdrjacoby_b200.so")
HEADERS = {"drj_device.cuh": "drj_embedded_device.inc", "drj_sampler.cuh": "drj_embedded_sampler.inc",
           "drj_models.cuh": "drj_embedded_models.inc", "drj_lanes.cuh": "drj_embedded_lanes.inc",
           "drj_spec.cuh": "drj_embedded_spec.inc"}
SOURCES = ["drj_host.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def embed() -> None:
    for hdr, inc in HEADERS.items():
        text = open(os.path.join(HERE, hdr)).read()
        assert ')DRJSRC"' not in text
        # raw string literals are split into < 16 KB pieces (portable literal length)
        pieces = [text[i:i + 12000] for i in range(0, len(text), 12000)]
        body = "\n".join('R"DRJSRC(' + p + ')DRJSRC"' for p in pieces) + "\n"
        path = os.path.join(HERE, inc)
        if not os.path.exists(path) or open(path).read() != body:
            open(path, "w").write(body)


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".py"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "drjacoby_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    embed()
    if not (force or stale()):
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + [os.path.join(HERE, s) for s in SOURCES] + \
          ["-ldl", "-lpthread", "-lrt"] + os.environ.get("DRJ_EXTRA_NVCC_FLAGS", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libdrjacoby_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
