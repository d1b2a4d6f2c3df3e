#!/bin/bash
mkdir -p gpurun_out
export JPGENC_SLOTS=1 JPGENC_FRAMES_PER_PASS=32
python tools/one_batch.py 32 3 > gpurun_out/e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:build_tables -s 1 -c 1 -o gpurun_out/e_prof_tables -f python tools/one_batch.py 32 1 > gpurun_out/e_ncu.log 2>&1
tail -2 gpurun_out/e_ncu.log
