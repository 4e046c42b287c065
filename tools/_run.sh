python -m pytest tests/test_gpu_api.py -x -q -k "cta_init" 2>&1 | tail -8
