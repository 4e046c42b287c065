"""Development tool: native-side breakdown of a small run (the README example's model through NVRTC, 5 chains, n = 100)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import drjacoby_b200 as drj
from helpers import Case, beta_ladder, gpu_spec

SRC = r'''
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  double ll = 0.0;
  for (long long i = 0; i < data.len(0); ++i) ll += R::dnorm(data.item(0)[i], params[0], params[1], true);
  return ll;
}
__device__ double logprior(const double* params, const DrjList& misc) {
  return R::dunif(params[0], -10.0, 10.0, true) + R::dlnorm(params[1], 0.0, 1.0, true);
}
'''
x = np.random.default_rng(1).normal(3, 2, 100)
models = {"jit generic": drj.Model(source=SRC), "builtin example": drj.Model(builtin="example")}
for mname, m in models.items():
    for rungs in (1, 10):
        case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(rungs), chains=5, burnin=1000, samples=10000)
        for rep in range(3):
            spec = gpu_spec(case, "philox", None, 1, 0)
            t0 = time.perf_counter(); s = drj.Session(m, spec); t1 = time.perf_counter()
            ms = s.advance(10**9); t2 = time.perf_counter()
            out = s.download(); t3 = time.perf_counter()
            s.close(); t4 = time.perf_counter()
            r0 = time.perf_counter(); drj.run_raw(m, spec); r1 = time.perf_counter()
            print(f"{mname:16s} rungs {rungs:2d} rep {rep}: create+init {1e3*(t1-t0):7.1f} ms  advance {1e3*(t2-t1):7.1f} ms (kernels {ms:6.1f} ms, {s.launches() if False else ''})"
                  f"  download {1e3*(t3-t2):6.1f} ms  close {1e3*(t4-t3):5.1f} ms | run_raw {1e3*(r1-r0):7.1f} ms", flush=True)
