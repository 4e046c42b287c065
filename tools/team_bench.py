"""Development tool: the cluster-per-chain (TEAM) kernels against the one-thread-per-particle kernels on
"few chains x long data" shapes of the normal model: ms per iteration, parameter updates/s, bytes of data
swept per second (every chain's cluster streams the item once per likelihood evaluation) and FP64 rate."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj
from drjacoby_b200 import _lib

peak = drj.fp64_peak_tflops(0)
print("fp64 peak", peak, flush=True)
m = drj.Model(builtin="example")
rows = []
shapes = [(1, 1, 10_000_000), (5, 1, 10_000_000), (9, 1, 10_000_000), (5, 1, 100_000_000), (5, 1, 1_000_000), (5, 1, 100_000),
          (5, 1, 10_000), (5, 8, 1_000_000), (5, 20, 1_000_000), (4, 32, 1_000_000), (32, 1, 1_000_000), (148, 1, 1_000_000),
          (512, 1, 1_000_000), (1024, 1, 1_000_000), (32, 32, 1_000_000), (128, 8, 1_000_000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for C, R, n in shapes:
    x = bench.synth_data(n)
    for name, flag in (("team", _lib.FLAG_TEAM_KERNEL), ("solo", _lib.FLAG_NO_TEAM_KERNEL)):
        # keep the slow side short: ~5 cycles per datum per sweep, 2 sweeps per iteration
        est_ms = (2 * n * 5 / 1.9e6) if name == "solo" else max(0.02, 2 * C * R * n * 2 / 1.5e10)
        iters = int(max(2, min(200, 400.0 / est_ms)))
        spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(C, 7), burnin=1,
                           samples=2 * iters + 1, beta_raised=np.linspace(0, 1, R) if R > 1 else [1.0], store_pt=False,
                           rng_mode="philox", seed=2, device=0, data=[x], iters_per_launch=iters, flags=flag)
        s = drj.Session(m, spec)
        s.advance(iters)
        ms = s.advance(iters) / iters
        s.close()
        ups = C * R * 2 / (ms * 1e-3)
        row = dict(chains=C, rungs=R, n=n, kernel=name, ms_per_iteration=ms, param_updates_per_s=ups,
                   data_GBps=C * 2 * n * 8 / (ms * 1e-3) / 1e9, tflops=4.0 * n * ups / 1e12, fp64_frac=4.0 * n * ups / 1e12 / peak)
        rows.append(row)
        print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/team_bench.jsonl", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
