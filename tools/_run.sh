python -m pytest tests/test_gpu_parity.py tests/test_gpu_spec.py tests/test_gpu_fullsize.py tests/test_gpu_team.py tests/test_gpu_grid.py -x -q 2>&1 | tail -6
python tools/cfg_run.py n100 K1 n4096 K2 2>&1 | cut -c1-150
DRJ_NO_HINTS=1 python tools/cfg_run.py n100 K1 2>&1 | cut -c1-150
