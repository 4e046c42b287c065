"""Development tool: where the wall time of the ESS run (K1 model, many chains, summary mode) goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj
from drjacoby_b200 import diagnostics as dg

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
if len(sys.argv) > 2 and sys.argv[2] == "pin":  # does a pinned host buffer in the process change anything?
    import torch
    torch.cuda.set_device(0)
    x_pin = torch.empty(10_000_000, dtype=torch.float64).pin_memory()
x = bench.synth_data(100)
# warm the context with a tiny run so that rep 0 shows the cost of the first LARGE allocation alone
_s = drj.Session(drj.Model(builtin="example"), drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(2, 5),
                                                          burnin=5, samples=5, rng_mode="philox", seed=3, device=0, data=[x]))
_s.advance(9); _s.summary(); _s.close()
for rep in range(6):
    t = [time.perf_counter()]
    spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(chains, 5), burnin=1000,
                       samples=10000, rng_mode="philox", seed=3, device=0, data=[x])
    t.append(time.perf_counter())
    s = drj.Session(drj.Model(builtin="example"), spec)
    t.append(time.perf_counter())
    s.advance(999)
    t.append(time.perf_counter())
    s.advance(10000)
    t.append(time.perf_counter())
    a = s.summary(max_lag=0, autocov=False)
    t.append(time.perf_counter())
    b = s.summary(max_lag=88, center=a["chain_mean"].mean(axis=0), moments=False)
    t.append(time.perf_counter())
    raw = s.download(draws=False)
    t.append(time.perf_counter())
    tm = s.timing()
    s.close()
    t.append(time.perf_counter())
    names = ["spec", "session create (alloc, upload, init kernel)", "burn-in advance", "sampling advance", "summary moments", "summary autocov",
             "download counters", "close"]
    print(f"rep {rep}: total {t[-1]-t[0]:.3f} s;", "; ".join(f"{n} {1e3*(t[i+1]-t[i]):.1f} ms" for i, n in enumerate(names)), "| timing", tm, "launches", None)
