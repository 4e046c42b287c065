"""Development tool: where does the end-to-end time of one K5 drj_run_mcmc call go?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj

x = bench.synth_data()
m = drj.Model(builtin="example")
for rep in range(2):
    spec = bench.k5_spec(drj, x, bench.CHAINS_PER_GPU, 0, 0, burnin=1, samples=4, seed=4)
    t0 = time.perf_counter(); s = drj.Session(m, spec); t1 = time.perf_counter()
    ms = s.advance(4); t2 = time.perf_counter()
    out = s.download(); t3 = time.perf_counter()
    s.close(); t4 = time.perf_counter()
    print(f"rep {rep}: create+init {t1-t0:.3f}s  advance(4) {t2-t1:.3f}s (kernels {ms/1e3:.3f}s)  download {t3-t2:.3f}s  close {t4-t3:.3f}s")
    t0 = time.perf_counter(); out = drj.run_raw(m, spec); print(f"   run_raw total {time.perf_counter()-t0:.3f}s")
