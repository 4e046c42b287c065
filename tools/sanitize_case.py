"""Development tool: a few small runs covering every kernel path (resident + streamed tiles, TMA,
coupling exchange, blocks, NVRTC) for compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Case, beta_ladder, run_gpu
import drjacoby_b200 as drj
rng = np.random.Generator(np.random.Philox(3))
x = rng.normal(3, 2, 9001)
cases = [
    Case("example", [-10, 0], [10, np.inf], data=[x[:100]], beta=beta_ladder(4), chains=5, burnin=20, samples=20),
    Case("example", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(3), chains=3, burnin=3, samples=3),
    Case("doublewell", [-2, 30], [2, 30], beta=beta_ladder(20, 2.0), chains=7, burnin=20, samples=20, save_hot_draws=True),
    Case("multilevel", [-5.0] * 5, [5.0] * 5, data=[rng.normal(0, 1, 5) for _ in range(5)], blocks=[[g + 1, 6] for g in range(5)],
         beta=beta_ladder(3), chains=4, burnin=10, samples=10),
]
for c in cases:
    out = run_gpu(c, "philox", seed=1)
    print(c.model, len(c.beta), float(out["loglike_sampling"].sum()))
src = "__device__ double loglike(const double* p, const DrjList& d, const DrjList& m, int b) { double r = 0; for (long long i = 0; i < d.len(0); ++i) r += R::dnorm(d.item(0)[i], p[0], 1.0, 1); return r; }\n__device__ double logprior(const double* p, const DrjList& m) { return 0.0; }"
c = Case("jit", [-10], [10], data=[x[:50]], beta=beta_ladder(2), chains=3, burnin=10, samples=10)
out = run_gpu(c, "philox", seed=1, model=drj.Model(source=src))
print("jit", float(out["loglike_sampling"].sum()))
