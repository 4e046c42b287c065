python -m pytest tests/test_gpu_nmath.py -x -q 2>&1 | tail -12
