#!/bin/bash
# round 2, 8-GPU validation: multi-GPU parity (dist_check: HVP / residual / CG / Newton vs single GPU), the bench line at
# N = 8 (weak + strong scaling blocks, config 5, parity), and one distributed CG iteration at config 4.  Every step is
# wrapped in its own timeout.
mkdir -p gpurun_out
N=${1:-8}
TATVA_CHECK_HALOS=nccl,peer timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 10 > gpurun_out/r02_dist_check_${N}gpu.log 2>&1; echo "dist_check rc=$?"
tail -2 gpurun_out/r02_dist_check_${N}gpu.log | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r02_bench_${N}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=',d['n_gpus'],'GDOF/s=',round(d['value']/1e9,3),'ms/step=',round(d['ms_per_step'],4),'e2e ms=',round(d['e2e']['ms_per_step'],3),'parity',d.get('parity',{}).get('hvp_rel_err'))
        print('strong',d['strong_scaling']['ms_per_step'], 'c5', d['secondary']['c5_compound_tet4_pf'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/bench_dist_cg.py 128 peer 100 > gpurun_out/r02_dist_cg_${N}gpu.json 2> gpurun_out/r02_dist_cg_${N}gpu.err; echo "dist_cg rc=$?"; tail -1 gpurun_out/r02_dist_cg_${N}gpu.json | cut -c1-400
