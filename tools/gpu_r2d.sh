#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "table or batch or frames" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
for n in 128 1024; do python tools/one_batch.py $n 5; JPGENC_TABLES_LEAN=0 python tools/one_batch.py $n 5; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:build_tables --launch-skip 4 --launch-count 4 --csv --log-file gpurun_out/r2d_tables.csv python tools/one_batch.py 128 1 > /dev/null 2>&1
grep build_tables gpurun_out/r2d_tables.csv | awk -F'","' '{print $NF}'
