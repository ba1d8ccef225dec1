#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 16 > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$?"
tail -5 gpurun_out/dist_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench rc=$?"
cat gpurun_out/bench_$N.json; tail -5 gpurun_out/bench_$N.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err; cat gpurun_out/bench_1.json
