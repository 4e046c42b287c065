"""Development tool: wall time and Python profile of run_mcmc(output="summary") on the ESS workload."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = bench.synth_data(100)
df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
drj.use_builtin("example")
kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", seed=3, silent=True, output="summary", device=0)
drj.run_mcmc(burnin=10, samples=20, chains=2, **kw)
for rep in range(2):
    t0 = time.perf_counter()
    out = drj.run_mcmc(burnin=1000, samples=10000, chains=chains, **kw)
    print("rep", rep, "wall", time.perf_counter() - t0, "ess", out.diagnostics["ess"].tolist(), flush=True)
pr = cProfile.Profile()
pr.enable()
out = drj.run_mcmc(burnin=1000, samples=10000, chains=chains, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
