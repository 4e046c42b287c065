timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
python tools/cfg_run.py K1 K4 > gpurun_out/cfg_lanes4.jsonl 2>&1
cat gpurun_out/cfg_lanes4.jsonl | cut -c1-160
