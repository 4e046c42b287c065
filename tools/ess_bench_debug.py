"""Development tool: the bench's ESS sweep, repeated, inside the bench's own 1-GPU process set-up."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import drjacoby_b200 as drj

torch.cuda.set_device(0)
if len(sys.argv) > 1 and sys.argv[1] == "pin":
    x_pin = torch.empty(10_000_000, dtype=torch.float64).pin_memory()
def barrier():
    torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    r = bench.ess_sweep(drj, 0, 0, 1, barrier)
    print(rep, "outer", time.perf_counter() - t0, "wall_s", r["wall_s"], "value", r["value"], flush=True)
