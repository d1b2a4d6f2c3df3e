#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/c_matrix.log; : > $L
run() { # frames reps slots per stagger skip [prefix]
  echo -n "frames $1 slots $3 per $4 stagger $5 skip $6 $7: " >> $L
  JPGENC_SLOTS=$3 JPGENC_FRAMES_PER_PASS=$4 JPGENC_STAGGER=$5 JPGENC_DEBUG_SKIP_TABLES=$6 $7 python tools/one_batch.py $1 $2 2>&1 | tail -1 >> $L
}
for st in 0 1; do for sk in 0 1; do
  for per in 16 32 64; do run 128 20 4 $per $st $sk; done
  for per in 32 64 128; do run 1024 5 4 $per $st $sk; done
done; done
run 128 20 6 16 1 0
run 128 20 6 21 1 0
run 128 20 3 43 1 0
run 1024 5 6 64 1 0
run 1024 5 6 128 1 0
run 128 20 4 32 1 0 "taskset -c 0-3"
run 1024 5 4 64 1 0 "taskset -c 0-3"
run 1024 5 4 128 1 0 "taskset -c 0-1"
cat $L
