python tools/generic_probe.py 2>&1 | cut -c1-200
