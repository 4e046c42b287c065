"""Development tool: what a generic-form user likelihood (a serial loop over the data) costs, variant by variant."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import drjacoby_b200 as drj
from helpers import Case, beta_ladder, gpu_spec

PRIOR = r'''
__device__ double logprior(const double* params, const DrjList& misc) {
  return R::dunif(params[0], -10.0, 10.0, true) + R::dlnorm(params[1], 0.0, 1.0, true);
}
'''
LOOPS = {
 "A as written": "for (long long i = 0; i < data.len(0); ++i) ll += R::dnorm(data.item(0)[i], params[0], params[1], true);",
 "B unroll 4": "_Pragma(\"unroll 4\") for (long long i = 0; i < data.len(0); ++i) ll += R::dnorm(data.item(0)[i], params[0], params[1], true);",
 "C hoisted pointer/len, int index": "const double* x = data.item(0); const int n = (int)data.len(0); for (int i = 0; i < n; ++i) ll += R::dnorm(x[i], params[0], params[1], true);",
 "D log hoisted by hand": "const double* x = data.item(0); const int n = (int)data.len(0); const double ls = log(params[1]); for (int i = 0; i < n; ++i) { const double z = (x[i] - params[0]) / params[1]; ll += -(0.918938533204672741780329736406 + 0.5 * z * z + ls); }",
 "E D + reciprocal": "const double* x = data.item(0); const int n = (int)data.len(0); const double ls = log(params[1]), inv = 1.0 / params[1]; for (int i = 0; i < n; ++i) { const double z = (x[i] - params[0]) * inv; ll += -(0.918938533204672741780329736406 + 0.5 * z * z + ls); }",
 "F E + 4 partial sums": "const double* x = data.item(0); const int n = (int)data.len(0); const double ls = log(params[1]), inv = 1.0 / params[1]; double a[4] = {0,0,0,0}; for (int i = 0; i < n; ++i) { const double z = (x[i] - params[0]) * inv; a[i & 3] += -(0.918938533204672741780329736406 + 0.5 * z * z + ls); } ll = (a[0] + a[1]) + (a[2] + a[3]);",
}
x = np.random.default_rng(1).normal(3, 2, 100)
for name, loop in LOOPS.items():
    src = "__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {\n  double ll = 0.0;\n  " + loop + "\n  return ll;\n}\n" + PRIOR
    m = drj.Model(source=src)
    for rungs in (1, 10):
        case = Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(rungs), chains=5, burnin=500, samples=1500)
        spec = gpu_spec(case, "philox", None, 1, 0)
        s = drj.Session(m, spec); ms = s.advance(10**9); s.close()
        s = drj.Session(m, spec); ms = s.advance(10**9); s.close()
        print(f"{name:34s} rungs {rungs:2d}: {ms * 1e3 / 2000 / 2:7.2f} us per update", flush=True)
