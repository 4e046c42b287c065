"""HBM-bound regime: chains x 1 rung over n = 1e8 (800 MB >> L2) on the grid-per-run kernels vs the cluster-per-chain
kernels; swept bytes = 8 n per sweep (the item is read ONCE per sweep for all particles) x 2 sweeps per iteration.
Usage: python tools/grid_bench.py [n] ; env DRJ_GRID_STAGES / DRJ_GRID_CTAS tune."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drjacoby_b200 as drj
from drjacoby_b200 import _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
x = np.random.Generator(np.random.Philox(1)).normal(3.0, 2.0, n)
peak = 6546.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
shapes = [(1, 1), (5, 1), (2, 4), (16, 1)] if len(sys.argv) < 3 else [tuple(int(v) for v in s.split("x")) for s in sys.argv[2:]]
for chains, rungs in shapes:
    for fam, flag in (("grid", _lib.FLAG_GRID_KERNEL), ("team", _lib.FLAG_TEAM_KERNEL)):
        if fam == "team" and os.environ.get("GRID_ONLY"):
            continue
        iters = 20 if fam == "grid" else 10
        rng = np.random.Generator(np.random.Philox(5))
        init = np.stack([2.5 + rng.random(chains), 1.5 + rng.random(chains)], axis=1)
        beta = np.linspace(0.0, 1.0, rungs) if rungs > 1 else [1.0]
        spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=init, burnin=1, samples=2 * iters + 1,
                           beta_raised=beta, store_pt=False, rng_mode="philox", seed=2, data=[x], iters_per_launch=iters, flags=flag)
        s = drj.Session(drj.Model(builtin="example"), spec)
        s.advance(iters)
        ms = s.advance(iters)
        s.close()
        gbs = 2 * 8.0 * n * iters / (ms * 1e-3) / 1e9
        print(json.dumps({"family": fam, "chains": chains, "rungs": rungs, "n": n, "iters": iters, "kernel_ms": round(ms, 3),
                          "ms_per_sweep": round(ms / iters / 2, 4), "algorithmic_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                          "stages": os.environ.get("DRJ_GRID_STAGES"), "ctas": os.environ.get("DRJ_GRID_CTAS")}), flush=True)
