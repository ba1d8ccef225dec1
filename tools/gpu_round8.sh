#!/bin/bash
mkdir -p gpurun_out
python tools/bench_variants.py 128 0,23,0,23 > gpurun_out/variants.jsonl 2>&1; cat gpurun_out/variants.jsonl
timeout 900 python tools/bench_secondary.py tet4,pf,tri3,hex8 > gpurun_out/secondary_wide.jsonl 2> gpurun_out/secondary_wide.err; echo "secondary rc=$?"; cut -c1-200 gpurun_out/secondary_wide.jsonl; tail -3 gpurun_out/secondary_wide.err
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
