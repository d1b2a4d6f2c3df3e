#!/bin/bash
mkdir -p gpurun_out
python tools/one_batch.py 128 1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:build_tables --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2e_tables python tools/one_batch.py 128 1 > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/r2e_tables.ncu-rep
