#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms','gpu_launches')})
print(d['e2e'])
print(d['roofline']['frac'], d['extra']['batch1080p']['resident'], d['extra']['batch1080p']['files_returned'], d['extra']['batch1080p']['e2e'])
print(d['extra'].get('frame4k'))
PY
JPGENC_TRACE=1 python - <<'PY' 2>&1 | tail -6
import sys
sys.path.insert(0,'.')
from jpgenc_b200.capi import Encoder
enc=Encoder(0)
w=h=16384
d=enc.dev_alloc(w*h*3); enc.synth_rgb(d,w,h,0); enc.bind_device_rgb(d,w,h)
for _ in range(6): enc.encode_bound(None)
PY
