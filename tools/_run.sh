python tools/cfg_run.py K2 n100 n4096 2>&1 | cut -c1-150
python tools/cfg_run.py n100 K2 2>&1 | cut -c1-150
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_reference_suite.py -x -q 2>&1 | tail -3
