#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "table or frames or batch or packed" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
L=gpurun_out/d_matrix.log; : > $L
run() { # frames reps slots per stagger serial
  echo -n "frames $1 slots $3 per $4 stagger $5 serial $6: " >> $L
  JPGENC_SLOTS=$3 JPGENC_FRAMES_PER_PASS=$4 JPGENC_STAGGER=$5 JPGENC_TABLES_SERIAL=$6 python tools/one_batch.py $1 $2 2>&1 | tail -1 >> $L
}
run 32 20 1 32 0 1
run 32 20 1 32 0 0
run 8 20 1 8 0 0
for st in 0 1; do
  for per in 16 32 64; do run 128 20 4 $per $st 0; done
  for per in 32 64 128; do run 1024 5 4 $per $st 0; done
done
run 128 20 6 22 0 0
run 128 20 2 64 0 0
run 1024 5 6 64 0 0
run 1024 5 2 128 0 0
run 1024 5 3 128 0 0
cat $L
export JPGENC_SLOTS=1 JPGENC_FRAMES_PER_PASS=32
python tools/one_batch.py 32 3 > gpurun_out/d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches.csv python tools/one_batch.py 32 1 > gpurun_out/d_ncu1.log 2>&1
grep -E "build_tables|finalize" gpurun_out/d_launches.csv | tail -4
