"""Development tool: a short K4 run (4096 chains x 16 rungs x 50 params) for ncu captures."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import drjacoby_b200 as drj
from bench_configs import configs
from helpers import gpu_spec
case = [c for n, (c, _) in configs().items() if n.startswith("K4")][0]
case.burnin, case.samples = 1, 40
case.__post_init__()
s = drj.Session(drj.Model(builtin=case.model), gpu_spec(case, "philox", seed=1, store_pt=False, iters_per_launch=10))
s.advance(10)
print("ms for 30 iterations", s.advance(30))
s.close()
