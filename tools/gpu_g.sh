#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -3 gpurun_out/g_pytest.log
python tools/one_image.py 16384 16384 50
python tools/one_image.py 3840 2160 200
python tools/one_image.py 1920 1080 200
python tools/one_image.py 512 512 200
JPGENC_TRACE=1 python tools/one_image.py 3840 2160 3 2>&1 | tail -4
python tools/one_image.py 3840 2160 3 > gpurun_out/g_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches4k.csv python tools/one_image.py 3840 2160 1 > gpurun_out/g_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/g_launches4k.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
for r in rows[-9:]: print(r[0], r[ki][:36], r[gi], r[vi])
PY
