#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02b_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02b_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r02_bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=',d['n_gpus'],'GDOF/s=',round(d['value']/1e9,3),'ms/step=',round(d['ms_per_step'],4),'e2e ms=',round(d['e2e']['ms_per_step'],3),'parity',d.get('parity',{}).get('hvp_rel_err'))
        print('strong',d['strong_scaling']['ms_per_step'], 'c5', d['secondary']['c5_compound_tet4_pf'])
PY
tail -2 gpurun_out/r02_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/bench_dist_cg.py 128 peer 100 > gpurun_out/r02_dist_cg_2gpu.json 2> gpurun_out/r02_dist_cg_2gpu.err; echo "dist_cg rc=$?"; tail -1 gpurun_out/r02_dist_cg_2gpu.json | cut -c1-300
