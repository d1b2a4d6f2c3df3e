#!/bin/bash
# final single-GPU pass of round 2: parity tests, bench line, launch lists, ncu --set full of K1 (DRAM traffic for the bench's roofline.traffic)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms','gpu_launches')})
print(d['e2e'])
b=d['extra']['batch1080p']
print(d['roofline']['frac'], d['roofline'].get('traffic'), b.get('resident'), b.get('files_returned'), b.get('e2e'), b.get('error'))
print(d['extra'].get('frame4k'), d['extra'].get('dct_microbench'))
print(d.get('cpu_baseline'))
PY
python tools/one_image.py 16384 16384 2 > /dev/null &&
JPGENC_GRAPHS=0 ncu --set full --import-source on --clock-control none -k regex:forward_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2z_k1 python tools/one_image.py 16384 16384 1 > gpurun_out/r2z_ncu_k1.log 2>&1
JPGENC_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 10 --launch-count 10 --csv --log-file gpurun_out/r2z_launches16k.csv python tools/one_image.py 16384 16384 1 > /dev/null 2>&1
python tools/one_batch.py 1024 1 > /dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1096 --launch-count 72 --csv --log-file gpurun_out/r2z_launches_batch1024.csv python tools/one_batch.py 1024 1 > /dev/null 2>&1
ls -la gpurun_out/r2z*
# small frames: warm encodes at the library default (no stage events) and with all of them
for sz in "3840 2160 400" "1920 1080 400" "512 512 400"; do python tools/one_image.py $sz 0; python tools/one_image.py $sz 2; done | tee gpurun_out/r2z_small_frames.txt
python tools/one_image.py 3840 2160 3 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 18 --launch-count 9 --csv --log-file gpurun_out/r2z_launches4k.csv python tools/one_image.py 3840 2160 1 > /dev/null 2>&1
