timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
python tools/cfg_run.py K1 K4 > gpurun_out/cfg_lanes3.jsonl 2>&1
DRJ_LANES=0 python tools/cfg_run.py K4 >> gpurun_out/cfg_lanes3.jsonl 2>&1
cat gpurun_out/cfg_lanes3.jsonl | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:drj_sweep_kernel_lanes -s 1 -c 1 -f -o gpurun_out/k4_lanes_c python tools/cfg_run.py K4 > gpurun_out/ncu_k4_lanes_c.log 2>&1
