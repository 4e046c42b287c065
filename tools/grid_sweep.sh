#!/bin/bash
# parameter sweep of the grid kernels in the HBM-bound regime (run on the GPU box)
out=gpurun_out/grid_sweep.jsonl
: > $out
for st in 3 5 6 8 12; do
  for ct in 148 296; do
    GRID_ONLY=1 DRJ_GRID_STAGES=$st DRJ_GRID_CTAS=$ct python tools/grid_bench.py 1e8 1x1 5x1 >> $out 2>&1
  done
done
for ch in 4096 8192; do
  DRJ_EXTRA_NVCC_FLAGS="-DDRJ_GRID_CHUNK=$ch" python drjacoby_b200/csrc/build.py --force > /dev/null 2>&1
  for st in 3 4 6; do
    [ $((st * ch * 8)) -gt 200000 ] && continue
    for ct in 148 296; do
      echo "{\"chunk\": $ch}" >> $out
      GRID_ONLY=1 DRJ_GRID_STAGES=$st DRJ_GRID_CTAS=$ct python tools/grid_bench.py 1e8 1x1 5x1 >> $out 2>&1
    done
  done
done
cat $out
