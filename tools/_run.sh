python -m pytest tests/test_gpu_api.py -x -q -k "hints or datum_form" 2>&1 | tail -5
python tools/jit_time.py 2>&1 | tail -5
