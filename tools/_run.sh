python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo n1 rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
