#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 8 > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -o '"compound_pf_tet4_hvp_rel_err": [0-9.e-]*' gpurun_out/dist_check_$N.log | head -3
python tools/bench_c5.py 55 > gpurun_out/c5_1.json 2> gpurun_out/c5_1.err; cat gpurun_out/c5_1.json; tail -2 gpurun_out/c5_1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/bench_c5.py 55 > gpurun_out/c5_$N.json 2> gpurun_out/c5_$N.err; echo "c5 rc=$?"; grep '^{' gpurun_out/c5_$N.json; tail -2 gpurun_out/c5_$N.err
