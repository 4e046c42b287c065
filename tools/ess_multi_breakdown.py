"""Development tool: where the wall time of the one-process multi-GPU ESS run goes (K1 model, 65,536 chains per GPU, summary mode).
Usage: python tools/ess_multi_breakdown.py [n_gpus]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj

n = int(sys.argv[1]) if len(sys.argv) > 1 else drj._lib.lib().drj_device_count()
x = bench.synth_data(100)
for devs in ([0], list(range(n))):
    chains = 65536 * len(devs)
    for rep in range(4):
        t = [time.perf_counter()]
        spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(chains, 5), burnin=1000,
                           samples=10000, rng_mode="philox", seed=3, data=[x], store_pt=False)
        t.append(time.perf_counter())
        s = drj.Session(drj.Model(builtin="example"), spec, devices=devs)
        t.append(time.perf_counter())
        s.advance(999)
        t.append(time.perf_counter())
        s.advance(10000)
        t.append(time.perf_counter())
        a = s.summary(max_lag=88)
        t.append(time.perf_counter())
        raw = s.download(draws=False)
        t.append(time.perf_counter())
        s.close()
        t.append(time.perf_counter())
        names = ["spec", "create", "burn-in", "sampling", "summary", "download counters", "close"]
        print(f"{len(devs)} device(s) rep {rep}: total {t[-1]-t[0]:.3f} s;", "; ".join(f"{nm} {1e3*(t[i+1]-t[i]):.1f} ms" for i, nm in enumerate(names)), flush=True)
