python tools/cfg_run.py K3 2>&1 | cut -c1-140
python -m pytest tests/test_gpu_nmath.py tests/test_gpu_parity.py tests/test_gpu_lanes.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
