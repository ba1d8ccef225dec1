#!/bin/bash
# round 2, session 3, two GPUs: the 2-rank GPU test (all three halo transports), dist_check over the transports, and the
# bench line with each transport (per-step overhead of the exchange over the bare kernel)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02b_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02b_pytest_multi.log
TATVA_CHECK_HALOS=nccl,nccl_torch,peer timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 10 > gpurun_out/r02b_dist_check_2gpu.log 2>&1; echo "dist_check rc=$?"
tail -3 gpurun_out/r02b_dist_check_2gpu.log | cut -c1-600
for H in nccl nccl_torch peer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --halo $H > gpurun_out/r02b_halo_${H}_2gpu.json 2> gpurun_out/r02b_halo_${H}_2gpu.err
  echo "halo=$H rc=$?"; python -c "
import json
for l in open('gpurun_out/r02b_halo_${H}_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print('halo=$H N=',d['n_gpus'],'GDOF/s=',round(d['value']/1e9,3),'ms/step=',round(d['ms_per_step'],4),'kernel_ms=',round(d['roofline']['kernel_ms'],4),'parity',d.get('parity'))
"; tail -2 gpurun_out/r02b_halo_${H}_2gpu.err
done
