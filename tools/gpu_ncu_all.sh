#!/bin/bash
# ncu --set full of every kernel of one steady-state encode (16384^2 and 3840x2160) and of one batched pass, at HEAD
mkdir -p gpurun_out
export JPGENC_GRAPHS=0
python tools/one_image.py 16384 16384 2 > /dev/null || exit 1
NCU="ncu --set full --import-source on --clock-control none"
$NCU --launch-skip 14 --launch-count 7 -f -o gpurun_out/r2z_image16k python tools/one_image.py 16384 16384 1 > gpurun_out/r2z_ncu_a.log 2>&1
$NCU --launch-skip 14 --launch-count 7 -f -o gpurun_out/r2z_frame4k python tools/one_image.py 3840 2160 1 > gpurun_out/r2z_ncu_b.log 2>&1
$NCU -k regex:'forward|refine|symbol_stats|build_tables|finalize|range_bits|huffman_pack|ff_count|stuff' --launch-skip 18 --launch-count 9 -f -o gpurun_out/r2z_batchpass python tools/one_batch.py 128 1 > gpurun_out/r2z_ncu_c.log 2>&1
ls -la gpurun_out/r2z_image16k.ncu-rep gpurun_out/r2z_frame4k.ncu-rep gpurun_out/r2z_batchpass.ncu-rep
