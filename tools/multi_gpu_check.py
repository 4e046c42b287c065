"""Run under torchrun on N GPUs: run_mcmc(cluster="torch") (full and summary output) must equal the same
run on one GPU (rank 0 recomputes it alone).  Prints OK lines on rank 0."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import drjacoby_b200 as drj

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = np.random.default_rng(1).normal(3, 2, 200)
df = drj.define_params("name", "mu", "min", -10, "max", 10, "name", "sigma", "min", 0, "max", np.inf)
drj.use_builtin("example")
kw = dict(data={"x": x}, df_params=df, loglike="loglike", logprior="logprior", burnin=200, samples=800, rungs=4, chains=13, seed=5, silent=True)
full = drj.run_mcmc(cluster="torch", **kw)
summ = drj.run_mcmc(cluster="torch", output="summary", **kw)
if rank == 0:
    one = drj.run_mcmc(device=0, **kw)
    for k, v in one.raw.items():
        if k != "t_diff":
            assert np.array_equal(full.raw[k], v), k
    assert full.diagnostics["rhat"].equals(one.diagnostics["rhat"])
    print("OK full: sharded run over", dist.get_world_size(), "GPUs == single-GPU run, bit for bit")
    for k in ("mu", "sigma"):
        assert summ.diagnostics["rhat"][k] == one.diagnostics["rhat"][k]
        rel = abs(summ.diagnostics["ess"][k] - one.diagnostics["ess"][k]) / one.diagnostics["ess"][k]
        assert rel < 1e-9, rel   # including the lag products that straddle the shard boundary
    assert abs(summ.diagnostics["DIC_Gelman"] - one.diagnostics["DIC_Gelman"]) < 1e-9 * abs(one.diagnostics["DIC_Gelman"])
    print("OK summary: Rhat equal, ESS rel diff", rel, ", DIC equal")
# few chains over a long data vector: every shard uses the cluster-per-chain kernels, as the one-GPU run does
xl = np.random.default_rng(2).normal(3, 2, 30001)
kw = dict(data={"x": xl}, df_params=df, loglike="loglike", logprior="logprior", burnin=30, samples=60, rungs=3, chains=5, seed=9, silent=True)
few = drj.run_mcmc(cluster="torch", **kw)
if rank == 0:
    one = drj.run_mcmc(device=0, **kw)
    for k, v in one.raw.items():
        if k != "t_diff":
            assert np.array_equal(few.raw[k], v), k
    print("OK few chains: 5 chains x 3 rungs over n = 30001, sharded == single GPU, bit for bit")
dist.barrier()
dist.destroy_process_group()
