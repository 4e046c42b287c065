"""Development tool: one short TEAM-kernel session (few chains x long data) for an ncu capture."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj
from drjacoby_b200 import _lib

C, R, n = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "5x1x10000000").split("x"))
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
x = bench.synth_data(n)
m = drj.Model(builtin="example")
spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(C, 7), burnin=1,
                   samples=2 * iters + 1, beta_raised=np.linspace(0, 1, R) if R > 1 else [1.0], store_pt=False,
                   rng_mode="philox", seed=2, device=0, data=[x], iters_per_launch=iters, flags=_lib.FLAG_TEAM_KERNEL)
s = drj.Session(m, spec)
s.advance(iters)
ms = s.advance(iters)
s.close()
print("ms per iteration", ms / iters)
