"""Development tool: the n = 100 sweep of bench.py on its own (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import drjacoby_b200 as drj
print(bench.small_n_sweep(drj, 0, 0))
