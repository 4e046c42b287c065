"""Development tool: NVRTC compile time of a user model (cold cache) and reload time (warm disk cache)."""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["DRJ_CACHE_DIR"] = tempfile.mkdtemp()
import drjacoby_b200 as drj
from test_gpu_api import NORMAL_SRC, DATUM_SRC
for name, src in (("generic", NORMAL_SRC), ("datum", DATUM_SRC)):
    for rep in ("cold", "warm"):
        m = drj.Model(source=src + f"\n// {name}\n")
        t0 = time.perf_counter()
        m.handle(2, 1, 0, ["mu", "sigma"], ["x"], [])
        print(name, rep, f"handle {time.perf_counter()-t0:.2f}s  nvrtc {m.compile_seconds():.2f}s")
