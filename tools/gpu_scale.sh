#!/bin/bash
# weak-scaling run on N GPUs: parity check, then bench at 1, 2, 4, ..., N
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 10 > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"
tail -2 gpurun_out/dist_check_$N.log | cut -c1-600
for G in 1 2 4 8; do
  if [ $G -le $N ]; then
    if [ $G -eq 1 ]; then
      python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2952$G bench.py --gpus $G --no-cpu-baseline > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
    fi
    echo "bench $G rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/scale_$G.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=',d['n_gpus'],'GDOF/s=',round(d['value']/1e9,3),'ms/step=',round(d['ms_per_step'],4),'kernel_ms=',round(d['roofline']['kernel_ms'],4),'e2e GDOF/s=',round(d['e2e']['value']/1e9,3), d['clocks'])
"
  fi
done
