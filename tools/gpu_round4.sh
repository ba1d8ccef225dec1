#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/bench_secondary.py tet4,pf > gpurun_out/secondary2.jsonl 2> gpurun_out/secondary2.err; cat gpurun_out/secondary2.jsonl; tail -3 gpurun_out/secondary2.err
