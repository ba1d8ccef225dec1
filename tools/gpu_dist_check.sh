#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 10 > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"
tail -3 gpurun_out/dist_check_$N.log | cut -c1-700
