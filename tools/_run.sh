python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --configs K3,K1 --hbm-n 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo rc=$?
