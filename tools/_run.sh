python tools/cfg_run.py n100 2>&1 | cut -c1-150
cp drjacoby_b200/libdrjacoby_b200.so /tmp/orig.so; cp tools/libbatch.so drjacoby_b200/libdrjacoby_b200.so
DRJ_BATCH=1 python tools/cfg_run.py n100 2>&1 | cut -c1-150
DRJ_BATCH=1 python tools/cfg_run.py n100 2>&1 | cut -c1-150
