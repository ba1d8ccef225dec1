#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/bench_c5.py 55 > gpurun_out/c5_$N.json 2> gpurun_out/c5_$N.err; echo "c5 rc=$?"; grep '^{' gpurun_out/c5_$N.json; tail -2 gpurun_out/c5_$N.err | cut -c1-300
