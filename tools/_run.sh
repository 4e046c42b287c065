python -m pytest tests/test_gpu_parity.py -x -q -k "draws" 2>&1 | tail -6
python tools/cfg_run.py K2 K3 n100 2>&1 | cut -c1-150
