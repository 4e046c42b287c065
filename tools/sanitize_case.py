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

# round 2 kernel families: grid per run (slot and tile form, TMA and plain), cluster per chain, lanes, speculative
from drjacoby_b200 import _lib as L
which = os.environ.get("SAN_CASES", "grid,team,lanes,spec").split(",")
xs = rng.normal(3, 2, 20001)
r2 = []
if "grid" in which:
    r2 += [("grid tile", Case("example", [-10, 0], [10, np.inf], data=[xs], chains=5, burnin=3, samples=3), L.FLAG_GRID_KERNEL),
           ("grid tile plain", Case("example", [-10, 0], [10, np.inf], data=[xs[:9001]], chains=3, burnin=2, samples=2), L.FLAG_GRID_KERNEL | L.FLAG_NO_TMA),
           ("grid slot", Case("coupling", [-10, 0, -np.inf], [10, 10, np.inf], data=[xs[:9000]], beta=beta_ladder(10), chains=2, burnin=3, samples=3), L.FLAG_GRID_KERNEL)]
if "team" in which:
    r2 += [("team", Case("example", [-10, 0], [10, np.inf], data=[xs], beta=beta_ladder(4), chains=3, burnin=3, samples=3), L.FLAG_TEAM_KERNEL)]
if "lanes" in which:
    r2 += [("lanes", Case("multilevel", [-5.0] * 5, [5.0] * 5, data=[rng.normal(0, 1, 5) for _ in range(5)], blocks=[[g + 1, 6] for g in range(5)],
                          beta=beta_ladder(4), chains=9, burnin=6, samples=6), L.FLAG_LANES_KERNEL),
           ("lanes datum", Case("example", [-10, 0], [10, np.inf], data=[x[:100]], beta=beta_ladder(2), chains=5, burnin=6, samples=6), L.FLAG_LANES_KERNEL)]
if "spec" in which:
    r2 += [("spec", Case("example", [-10, 0], [10, np.inf], data=[x[:100]], chains=2, burnin=15, samples=15), L.FLAG_SPEC_KERNEL)]
for name, c, fl in r2:
    out = run_gpu(c, "philox", seed=1, flags=fl)
    print(name, float(out["loglike_sampling"].sum()), flush=True)
