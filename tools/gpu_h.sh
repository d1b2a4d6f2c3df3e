#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -3 gpurun_out/h_pytest.log
for g in 0 1; do
echo "JPGENC_GRAPHS=$g"
JPGENC_GRAPHS=$g python tools/one_image.py 16384 16384 50
JPGENC_GRAPHS=$g python tools/one_image.py 3840 2160 200
JPGENC_GRAPHS=$g python tools/one_image.py 1920 1080 200
JPGENC_GRAPHS=$g python tools/one_image.py 512 512 200
done
JPGENC_TRACE=1 python tools/one_image.py 3840 2160 3 2>&1 | tail -4
