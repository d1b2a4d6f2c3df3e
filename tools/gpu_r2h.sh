#!/bin/bash
mkdir -p gpurun_out
export JPGENC_GRAPHS=0
python tools/one_image.py 16384 16384 2 || exit 1
NCU="ncu --set full --import-source on --clock-control none"
$NCU -k regex:'huffman_pack|stuff_kernel' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r2h_k3b_k4 python tools/one_image.py 16384 16384 1 > gpurun_out/r2h_ncu.log 2>&1
ls -la gpurun_out/r2h*.ncu-rep
