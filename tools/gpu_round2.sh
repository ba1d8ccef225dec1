#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/bench_variants.py 128 > gpurun_out/variants.jsonl 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/variants.jsonl
