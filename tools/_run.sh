python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_abi.py -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; echo n2 rc=$?
tail -c 600 gpurun_out/bench_n2_final.json
