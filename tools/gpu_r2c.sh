#!/bin/bash
mkdir -p gpurun_out
export JPGENC_GRAPHS=0
python tools/one_image.py 16384 16384 2 || exit 1
NCU="ncu --set full --import-source on --clock-control none"
$NCU -k regex:symbol_stats --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2c_k2 python tools/one_image.py 16384 16384 1 > gpurun_out/r2c_ncu_a.log 2>&1
$NCU -k regex:stuff_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2c_k4 python tools/one_image.py 16384 16384 1 > gpurun_out/r2c_ncu_b.log 2>&1
$NCU -k regex:'forward_kernel|refine_kernel|symbol_stats|range_bits|huffman_pack|stuff_kernel' --launch-skip 12 --launch-count 6 -f -o gpurun_out/r2c_4k python tools/one_image.py 3840 2160 1 > gpurun_out/r2c_ncu_c.log 2>&1
ls -la gpurun_out/r2c*.ncu-rep
