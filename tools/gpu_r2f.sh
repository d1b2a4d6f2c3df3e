#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
python tools/one_image.py 16384 16384 50
python tools/one_image.py 3840 2160 200
python tools/batch_sweep.py 128 1,2,4 32,64,128 2>&1 | tail -12
python tools/batch_sweep.py 1024 2,4 64,128,256 2>&1 | tail -8
python tools/one_batch.py 128 1 > /dev/null &&
ncu --set full --import-source on --clock-control none -k regex:build_tables --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2e_tables python tools/one_batch.py 128 1 > gpurun_out/r2e_ncu.log 2>&1
