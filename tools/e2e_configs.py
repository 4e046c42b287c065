"""Development tool: end-to-end wall time of drj_run_mcmc (host buffers in, all draws out) for K2-K4."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import drjacoby_b200 as drj
from bench_configs import configs
from helpers import gpu_spec
for name, (case, _) in configs().items():
    m = drj.Model(builtin=case.model)
    spec = gpu_spec(case, "philox", seed=1)
    drj.run_raw(m, gpu_spec(case.__class__(case.model, case.tmin, case.tmax, data=case.data, blocks=case.blocks, beta=case.beta, chains=2, burnin=2, samples=2), "philox"))
    t0 = time.perf_counter(); out = drj.run_raw(m, spec); t1 = time.perf_counter()
    nbytes = sum(v.nbytes for v in out.values())
    s = drj.Session(m, spec); ms = s.advance(10**9); t2 = time.perf_counter(); s.download(); t3 = time.perf_counter(); s.close()
    print(f"{name[:40]:40s} run_raw {t1-t0:6.2f}s  out {nbytes/1e9:5.2f} GB  kernels {ms/1e3:6.2f}s  download {t3-t2:5.2f}s")
