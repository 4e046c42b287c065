#!/bin/bash
# build variants of the library with -D switches and time named configs (run on the GPU box)
cfgs="$1"; shift
for flags in "$@"; do
  DRJ_EXTRA_NVCC_FLAGS="$flags" python drjacoby_b200/csrc/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "== flags: $flags"
  python tools/cfg_run.py $cfgs | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['config'], '%.4e' % d['value'], round(d['device_ms'], 2))"
done
