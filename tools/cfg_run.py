"""Development tool: time named bench configs on one GPU (no CPU legs).  Usage: python tools/cfg_run.py K4 n100 ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import drjacoby_b200 as drj

names = sys.argv[1:] or ["K1", "default5", "K2", "K3", "K4", "n100", "n4096"]
x_long = bench.synth_data(100_000_000) if any(n.startswith("hbm") for n in names) else None
fp64 = drj.fp64_peak_tflops(0)
for name, c in bench.named_configs(1, x_long, set(names)).items():
    r = bench.run_named_config(drj, name, c, 0, 1, 0, fp64, 6546.6, lambda v: [float(x) for x in v])
    print(json.dumps({"config": name, "value": r["value"], "device_ms": r["device_ms"], "wall_s": r["wall_s"], "frac": r["roofline"]["frac"],
                      "env": {k: v for k, v in os.environ.items() if k.startswith("DRJ_")}}), flush=True)
