#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for H in nccl peer; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --no-cpu-baseline --halo $H > gpurun_out/halo_${H}_$N.json 2> gpurun_out/halo_${H}_$N.err
  echo "halo=$H rc=$?"; python -c "
import json
for l in open('gpurun_out/halo_${H}_$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print('halo=$H N=',d['n_gpus'],'GDOF/s=',round(d['value']/1e9,3),'ms/step=',round(d['ms_per_step'],4),'kernel_ms=',round(d['roofline']['kernel_ms'],4),'e2e=',round(d['e2e']['value']/1e9,3))
"; tail -2 gpurun_out/halo_${H}_$N.err
done
