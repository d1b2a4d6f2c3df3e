#!/bin/bash
# round 2, first GPU call: parity tests, default bench, batch sweeps
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
nproc > gpurun_out/a_nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json
timeout 600 python tools/batch_sweep.py 1024 1,2,3,4,6 16,32,64,128 > gpurun_out/a_sweep1024.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/a_sweep1024.log
timeout 300 python tools/batch_sweep.py 128 1,2,4,6 8,16,32,64 > gpurun_out/a_sweep128.log 2>&1
cat gpurun_out/a_sweep128.log
JPGENC_TRACE=1 timeout 120 python tools/batch_sweep.py 128 4 32 > gpurun_out/a_trace128.log 2>&1
tail -30 gpurun_out/a_trace128.log
