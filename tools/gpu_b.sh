#!/bin/bash
mkdir -p gpurun_out
export JPGENC_SLOTS=1 JPGENC_FRAMES_PER_PASS=32
python tools/one_batch.py 32 3 > gpurun_out/b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches.csv python tools/one_batch.py 32 1 > gpurun_out/b_ncu1.log 2>&1
cat gpurun_out/b_plain.log
python tools/one_batch.py 32 3 > gpurun_out/b_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:build_tables -s 1 -c 1 -o gpurun_out/b_prof_tables -f python tools/one_batch.py 32 1 > gpurun_out/b_ncu2.log 2>&1
tail -3 gpurun_out/b_ncu2.log
ls -la gpurun_out/
