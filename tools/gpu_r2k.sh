#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
for r in 1 0; do
echo "REFINE64=$r"
JPGENC_REFINE64=$r python tools/one_image.py 16384 16384 50
JPGENC_REFINE64=$r python tools/one_image.py 3840 2160 200
JPGENC_REFINE64=$r python tools/one_image.py 1920 1080 200
done
python tools/one_batch.py 1024 5
JPGENC_REFINE64=0 python tools/one_batch.py 1024 5
JPGENC_TRACE=1 python tools/one_image.py 3840 2160 4 2>&1 | tail -3
JPGENC_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 10 --launch-count 10 --csv --log-file gpurun_out/r2k_launches16k.csv python tools/one_image.py 16384 16384 1 > /dev/null 2>&1
JPGENC_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 10 --launch-count 10 --csv --log-file gpurun_out/r2k_launches4k.csv python tools/one_image.py 3840 2160 1 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ['r2k_launches16k','r2k_launches4k']:
    rows=list(csv.reader(open(f'gpurun_out/{f}.csv')))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
    for r in rows[hdr+2:hdr+12]: print(r[ki][:28], r[gi], r[vi])
PY
