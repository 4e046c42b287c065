S=$(date +%s); python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo rc=$? elapsed=$(( $(date +%s) - S ))s
tail -3 gpurun_out/bench_now.err | cut -c1-300
