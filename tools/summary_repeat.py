"""Development tool: repeat the on-device summary of one resident run; prints every call's wall time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = bench.synth_data(100)
spec = drj.RunSpec(theta_min=[-10.0, 0.0], theta_max=[10.0, np.inf], theta_init=bench.default_init(chains, 5), burnin=1000,
                   samples=10000, rng_mode="philox", seed=3, device=0, data=[x])
s = drj.Session(drj.Model(builtin="example"), spec)
s.advance(spec.T)
a = s.summary(max_lag=0, autocov=False)
ctr = a["chain_mean"].mean(axis=0)
tm, ta = [], []
for i in range(40):
    t0 = time.perf_counter(); s.summary(max_lag=0, autocov=False); t1 = time.perf_counter()
    s.summary(max_lag=88, center=ctr, moments=False); t2 = time.perf_counter()
    tm.append(1e3 * (t1 - t0)); ta.append(1e3 * (t2 - t1))
s.close()
print("moments ms:", " ".join(f"{v:.1f}" for v in tm))
print("autocov ms:", " ".join(f"{v:.1f}" for v in ta))
