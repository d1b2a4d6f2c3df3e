#!/bin/bash
# usage: gpu_n.sh N  -- default bench on N GPUs under torchrun (+ the same without NUMA binding)
N=$1
mkdir -p gpurun_out
nproc > gpurun_out/n${N}_nproc.txt; nvidia-smi topo -m > gpurun_out/n${N}_topo.txt 2>&1
for f in /sys/bus/pci/devices/*/numa_node; do :; done
run() { # tag extra-env
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_$1.json 2> gpurun_out/n${N}_$1.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/n${N}_$1.json'))
    b=d['extra']['batch1080p']
    print('$1', 'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'h2d ms',d['e2e']['ms_h2d'],d.get('host'))
    print('   batch resident',b['resident'],'files',b['files_returned'],'e2e',b['e2e'])
except Exception as ex:
    print('parse failed', ex); print(open('gpurun_out/n${N}_$1.err').read()[-1500:])
PY
}
run numa "A=1"
run nonuma "JPGENC_BENCH_NO_NUMA=1"
