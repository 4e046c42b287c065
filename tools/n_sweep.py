"""Development tool: parameter updates/s and FP64 roofline fraction of the normal model as a function
of the data size n (12,500 chains x 32 rungs), for both sweep-kernel variants."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj
from drjacoby_b200 import _lib

peak = drj.fp64_peak_tflops(0)
print("fp64 peak", peak)
m = drj.Model(builtin="example")
for n in (100, 256, 512, 1024, 2048, 4096, 8192, 65536):
    x = bench.synth_data(n)
    for name, flag in (("resident", _lib.FLAG_RESIDENT_KERNEL), ("stream", _lib.FLAG_STREAM_KERNEL)):
        if n > 4096 and name == "resident":
            continue
        iters = max(4, min(400, int(4e6 / n)))
        spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(12500, 7), burnin=1,
                           samples=2 * iters, beta_raised=np.linspace(0, 1, 32), store_pt=False, rng_mode="philox", seed=2,
                           device=0, data=[x], iters_per_launch=iters, flags=flag)
        s = drj.Session(m, spec)
        s.advance(iters)
        ms = s.advance(iters)
        s.close()
        ups = 12500 * 32 * 2 * iters / (ms * 1e-3)
        print(f"n={n:6d} {name:9s} {ups:.3e} updates/s  sweep {4*n*ups/1e12:6.2f} TFLOP/s  frac {4*n*ups/1e12/peak:.3f}  x8 = {8*ups:.2e}/box")
