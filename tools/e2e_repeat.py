"""Development tool: repeat the bench's end-to-end call (K5 shard through drj_run_mcmc, host buffers) and print each wall."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import drjacoby_b200 as drj

torch.cuda.set_device(0)
x_pin = torch.empty(bench.N_DATA, dtype=torch.float64).pin_memory()
x = x_pin.numpy(); x[:] = bench.synth_data(bench.N_DATA)
model = drj.Model(builtin="example")
chains = bench.CHAINS_PER_GPU
drj.run_raw(model, bench.k5_spec(drj, x, chains, 0, 0, burnin=1, samples=1, seed=4))
for rep in range(4):
    spec = bench.k5_spec(drj, x, chains, 0, 0, burnin=1, samples=4, seed=4)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = drj.run_raw(model, spec)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(rep, f"wall {dt:.3f} s  -> {chains*32*2*4/dt:.4g} updates/s; library-reported call time {out['t_diff'].sum():.3f} s", flush=True)
