timeout 900 python -m pytest tests/test_gpu_spec.py -x -q 2>&1 | tail -5
python tools/cfg_run.py K1 K1 > gpurun_out/cfg_spec5.jsonl 2>&1
DRJ_SPEC_DEPTH=5 python tools/cfg_run.py K1 >> gpurun_out/cfg_spec5.jsonl 2>&1
cat gpurun_out/cfg_spec5.jsonl | cut -c1-130
