python tools/cfg_run.py default5 K1 2>&1 | cut -c1-140
python -m pytest tests/test_bench_contract.py -q 2>&1 | tail -2
