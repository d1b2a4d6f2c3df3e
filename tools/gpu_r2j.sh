#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
python tools/one_image.py 16384 16384 50
python tools/one_image.py 3840 2160 200
JPGENC_TRACE=1 python tools/one_image.py 3840 2160 4 2>&1 | tail -5
python - <<'PY'
import sys
sys.path.insert(0,'.')
from jpgenc_b200.capi import Encoder
enc=Encoder(0)
for (w,h) in [(16384,16384),(3840,2160),(1920,1080)]:
    d=enc.dev_alloc(w*h*3); enc.synth_rgb(d,w,h,0); enc.bind_device_rgb(d,w,h)
    enc.encode_bound(None); s=enc.stats()
    print(w,h,'refined blocks',s.refined_blocks,'of',s.n_blocks, 'fwd ms',round(s.ms_forward,4),'k1',round(s.ms_k1,4))
    enc.dev_free(d)
PY
