"""Development tool: wall time of the README usage example (run_mcmc, user model through NVRTC, 5 chains, n = 100, 1e3 + 1e4)
by phase: first call (NVRTC or disk cache), repeat calls, output="summary"."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import drjacoby_b200 as drj

x = np.random.default_rng(1).normal(3, 2, 100)
df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
t0 = time.perf_counter()
drj.source_cuda(r'''
__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  double ll = 0.0;
  for (long long i = 0; i < data.len(DATA_x); ++i) ll += R::dnorm(data.item(DATA_x)[i], params[P_mu], params[P_sigma], true);
  return ll;
}
__device__ double logprior(const double* params, const DrjList& misc) {
  return R::dunif(params[P_mu], -10.0, 10.0, true) + R::dlnorm(params[P_sigma], 0.0, 1.0, true);
}
''')
kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=1000, samples=10000, chains=5, rungs=1, silent=True)
for rep in range(4):
    t1 = time.perf_counter(); m = drj.run_mcmc(**kw); t2 = time.perf_counter()
    print(f"run_mcmc call {rep}: {t2 - t1:.3f} s  (kernels {m.raw['t_diff'] if hasattr(m, 'raw') and 't_diff' in getattr(m, 'raw', {}) else '?'})", flush=True)
for rep in range(2):
    t1 = time.perf_counter(); m = drj.run_mcmc(output="summary", **kw); t2 = time.perf_counter()
    print(f"run_mcmc summary call {rep}: {t2 - t1:.3f} s", flush=True)
kw["rungs"] = 10
for rep in range(2):
    t1 = time.perf_counter(); m = drj.run_mcmc(**kw); t2 = time.perf_counter()
    print(f"run_mcmc 10 rungs call {rep}: {t2 - t1:.3f} s", flush=True)
import cProfile, pstats
for rungs in (1, 10):
    kw["rungs"] = rungs
    pr = cProfile.Profile(); pr.enable(); m = drj.run_mcmc(**kw); pr.disable()
    print("rungs", rungs, "kernels", m.raw["t_diff"][:1])
    pstats.Stats(pr).sort_stats("tottime").print_stats(8)
