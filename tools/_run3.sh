timeout 600 ncu --set full --clock-control none --import-source on -k regex:drj_sweep_kernel_lanes -s 1 -c 1 -f -o gpurun_out/k4_lanes_a python tools/cfg_run.py K4 > gpurun_out/ncu_k4_lanes_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:drj_sweep_kernel_lanes -s 2 -c 1 -f -o gpurun_out/k1_lanes_a python tools/cfg_run.py K1 > gpurun_out/ncu_k1_lanes_a.log 2>&1
tail -3 gpurun_out/ncu_k4_lanes_a.log gpurun_out/ncu_k1_lanes_a.log
