#!/bin/bash
# multi-GPU parity (HVP / residual / plans / compound / distributed CG and Newton) and the distributed CG iteration bench
# usage: gpurun --gpus N -- './tools/gpu_dist_solver.sh N [cells per side per GPU]'
mkdir -p gpurun_out
N=${1:-2}
NB=${2:-64}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/dist_check.py 8 > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"; tail -3 gpurun_out/dist_check_$N.log | grep -o '"distributed_[a-z_=A-Z]*": {[^}]*}' | head -4; grep -i "error\|assert" gpurun_out/dist_check_$N.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 tools/bench_dist_cg.py $NB peer 100 > gpurun_out/dist_cg_$N.json 2> gpurun_out/dist_cg_$N.err; echo "dist_cg rc=$?"; grep '^{' gpurun_out/dist_cg_$N.json; tail -3 gpurun_out/dist_cg_$N.err | cut -c1-300
