"""Development tool: where the time of a small many-rung run (K2: 64 chains x 20 rungs, one free parameter) goes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import drjacoby_b200 as drj

c = bench.named_configs(1, None, {"K2"})["K2"]
T = 4000
for label, kw in [("default", {}), ("no coupling", {"coupling_on": False}), ("cpc 1", {"chains_per_cta": 1}), ("cpc 3", {"chains_per_cta": 3}),
                  ("R=1 x 1280", {"R1": True})]:
    cc = dict(c)
    r1 = kw.pop("R1", False)
    if r1:
        cc["beta"] = np.array([1.0]); cc["chains"] = 1280
    spec = bench.config_spec(drj, cc, 0, cc["chains"], 0)
    for k, v in kw.items():
        setattr(spec, k, v)
    if r1:
        spec.flags = drj._lib.FLAG_NO_SPEC_KERNEL | drj._lib.FLAG_NO_LANES_KERNEL
    m = drj.Model(builtin=cc["model"])
    s = drj.Session(m, spec); s.advance(200); t0 = s.timing(); s.advance(T); t1 = s.timing(); s.close()
    print(json.dumps({"case": label, "us_per_iter": (t1[0] - t0[0]) * 1e3 / T}), flush=True)
