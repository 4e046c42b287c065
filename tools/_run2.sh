timeout 900 python -m pytest tests/test_gpu_lanes.py -x -q 2>&1 | tail -15
python tools/cfg_run.py K1 K4 > gpurun_out/cfg_lanes.jsonl 2>&1
DRJ_LANES=0 python tools/cfg_run.py K1 K4 >> gpurun_out/cfg_lanes.jsonl 2>&1
cat gpurun_out/cfg_lanes.jsonl | cut -c1-160
