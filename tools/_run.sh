timeout 900 python -m pytest tests/test_gpu_spec.py -x -q 2>&1 | tail -3
python tools/cfg_run.py K1 K1 > gpurun_out/cfg_spec4.jsonl 2>&1
cat gpurun_out/cfg_spec4.jsonl | cut -c1-120
