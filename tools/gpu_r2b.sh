#!/bin/bash
# ncu --set full captures (with source) of one steady-state 16384^2 encode and of the batch-only kernels
mkdir -p gpurun_out
export JPGENC_GRAPHS=0
python tools/one_image.py 16384 16384 2 || exit 1
NCU="ncu --set full --import-source on --clock-control none"
$NCU --launch-skip 8 --launch-count 3 -f -o gpurun_out/r2b_k1_refine_k2 python tools/one_image.py 16384 16384 1 > gpurun_out/r2b_ncu_a.log 2>&1
$NCU --launch-skip 12 --launch-count 3 -f -o gpurun_out/r2b_k3_k4 python tools/one_image.py 16384 16384 1 > gpurun_out/r2b_ncu_b.log 2>&1
$NCU -k regex:'build_tables|finalize' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r2b_tables python tools/one_batch.py 128 1 > gpurun_out/r2b_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2b_ncu_a.log gpurun_out/r2b_ncu_b.log gpurun_out/r2b_ncu_c.log
