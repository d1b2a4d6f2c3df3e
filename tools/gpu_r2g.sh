#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
python tools/one_image.py 16384 16384 50
python tools/one_image.py 3840 2160 200
python tools/one_image.py 1920 1080 200
python tools/one_batch.py 128 5
python tools/one_batch.py 1024 5
JPGENC_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9 --launch-count 20 --csv --log-file gpurun_out/r2g_launches16k.csv python tools/one_image.py 16384 16384 1 > /dev/null 2>&1
JPGENC_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9 --launch-count 20 --csv --log-file gpurun_out/r2g_launches4k.csv python tools/one_image.py 3840 2160 1 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ['r2g_launches16k','r2g_launches4k']:
    rows=list(csv.reader(open(f'gpurun_out/{f}.csv')))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
    for r in rows[hdr+2:hdr+22]: print(r[ki][:28], r[gi], r[vi])
PY
