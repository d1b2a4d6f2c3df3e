#!/bin/bash
# usage: gpu_r2n.sh N ["fpp list"]  -- default bench on N GPUs under torchrun, then (optional) the batch workload alone with the given pass sizes
N=$1
mkdir -p gpurun_out
nproc > gpurun_out/r2n${N}_nproc.txt; nvidia-smi topo -m > gpurun_out/r2n${N}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2n${N}_bench.json 2> gpurun_out/r2n${N}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2n${N}_bench.json'))
    b=d['extra']['batch1080p']
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'h2d ms',d['e2e']['ms_h2d'],d.get('host'))
    print('   batch resident',b['resident'],'files',b['files_returned'],'e2e',b['e2e'])
except Exception as ex:
    print('parse failed', ex); print(open('gpurun_out/r2n${N}_bench.err').read()[-1500:])
PY
for fpp in $2; do
  JPGENC_FRAMES_PER_PASS=$fpp timeout 300 $TR bench.py --gpus $N --workload batch1080p --steps 100 > gpurun_out/r2n${N}_batch_fpp$fpp.json 2> gpurun_out/r2n${N}_batch_fpp$fpp.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2n${N}_batch_fpp$fpp.json'))
    print('fpp $fpp resident', d['frames_per_s'], d['ms_per_step'], 'files', d['files_returned'], 'e2e', d['e2e']['frames_per_s'], d['e2e']['ms_per_step'])
except Exception as ex:
    print('fpp $fpp parse failed', ex); print(open('gpurun_out/r2n${N}_batch_fpp$fpp.err').read()[-800:])
PY
done
