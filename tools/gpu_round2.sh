#!/bin/bash
set -x
mkdir -p gpurun_out

python tools/bench_variants.py 128 3,5,8,9,10,11 > gpurun_out/variants.jsonl 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/variants.jsonl
