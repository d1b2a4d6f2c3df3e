#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -15 gpurun_out/r2i_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms','gpu_launches')})
print(d['e2e'])
b=d['extra']['batch1080p']
print(d['roofline']['frac'], b.get('resident'), b.get('files_returned'), b.get('e2e'), b.get('error'))
print(d['extra'].get('frame4k'), d['extra'].get('dct_microbench'))
PY
