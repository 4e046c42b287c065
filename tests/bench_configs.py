"""Throughput of BASELINE.json's other configs (K1-K4) on one GPU next to the CPU oracle port.
Not the headline bench (bench.py = K5); prints one JSON line per config.  It lives under tests/ because it times
the CPU oracle as the baseline (oracle/ is test infrastructure: only tests/, smoke() and bench.py may load it).
Run on the GPU box:
    python tests/bench_configs.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # helpers
import drjacoby_b200 as drj  # noqa: E402
from helpers import Case, POPULATION, beta_ladder, gpu_spec, run_oracle  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def configs():
    rng = np.random.Generator(np.random.Philox(1))
    x100 = rng.normal(3.0, 2.0, 100)
    G = 50
    mu = rng.normal(0, 1, G)
    data50 = [rng.normal(mu[g], 1.0, 5) for g in range(G)]
    return {
        "K1 normal n=100, 1 chain x 1 rung, 1e3+1e4": (Case("example", [-10, 0], [10, np.inf], data=[x100], chains=1, burnin=1000, samples=10000), 1),
        "K2 double well, 64 chains x 20 rungs, alpha=2, 1e3+1e4": (Case("doublewell", [-2, 30], [2, 30], beta=beta_ladder(20, 2.0), chains=64, burnin=1000, samples=10000), 64),
        "K3 logistic growth, 1024 chains x 10 rungs, 1e3+1e4": (Case("logistic", [1, 1, 50], [100, 2, 200], data=[POPULATION["pop"], POPULATION["time"]], beta=beta_ladder(10), chains=1024, burnin=1000, samples=10000), 16),
        "K4 multilevel 50 params / 51 blocks, 4096 chains x 16 rungs, 1e3+1e3": (Case("multilevel", [-5.0] * G, [5.0] * G, data=data50, blocks=[[g + 1, G + 1] for g in range(G)], beta=beta_ladder(16), chains=4096, burnin=1000, samples=1000), 16),
    }


def main():
    cores = os.cpu_count() or 1
    for name, (case, cpu_chains) in configs().items():
        R = len(case.beta)
        T = case.burnin - 1 + case.samples
        s = drj.Session(drj.Model(builtin=case.model), gpu_spec(case, "philox", seed=1, store_pt=False))
        t0 = time.perf_counter()
        ms = s.advance(T)
        wall = time.perf_counter() - t0
        s.close()
        upd = case.chains * R * case.d_free * T
        sub = Case(case.model, case.tmin, case.tmax, data=case.data, blocks=case.blocks, beta=case.beta, chains=cpu_chains,
                   burnin=case.burnin, samples=min(case.samples, 2000), init=case.init[:cpu_chains])
        t0 = time.perf_counter()
        run_oracle(po, sub, "philox", seed=1, n_threads=cores)
        cpu_wall = time.perf_counter() - t0
        cpu_upd = cpu_chains * R * case.d_free * (sub.burnin - 1 + sub.samples)
        print(json.dumps({"config": name, "gpu_param_updates_per_s": upd / (ms * 1e-3), "gpu_chain_rung_updates_per_s": upd / case.d_free / (ms * 1e-3),
                          "gpu_kernel_ms": ms, "gpu_wall_s": wall, "cpu_param_updates_per_s": cpu_upd / cpu_wall, "cpu_cores": min(cores, cpu_chains),
                          "cpu_sample": f"{cpu_chains} chains x {sub.burnin}+{sub.samples} iterations"}))


if __name__ == "__main__":
    main()
