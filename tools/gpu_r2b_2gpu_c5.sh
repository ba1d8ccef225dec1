#!/bin/bash
mkdir -p gpurun_out
TATVA_CHECK_HALOS=nccl,peer timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 10 > gpurun_out/r02b_dist_check_2gpu.log 2>&1; echo "dist_check rc=$?"
tail -1 gpurun_out/r02b_dist_check_2gpu.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['per_rank'][1] if len(d['per_rank'])>1 else d['per_rank'][0]
for k,v in r.items(): print(k, str(v)[:200])
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 tools/bench_c5.py 55 peer > gpurun_out/r02_c5_2gpu.json 2> gpurun_out/r02_c5_2gpu.err; echo "c5 rc=$?"; tail -1 gpurun_out/r02_c5_2gpu.json | cut -c1-400
