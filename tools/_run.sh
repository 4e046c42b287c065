export PATH=/usr/local/cuda/bin:$PATH
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo ref rc=$?
python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo n1 rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py > gpurun_out/ncu_bench.log 2>&1; echo ncu rc=$?
wc -l gpurun_out/launches_final.csv
