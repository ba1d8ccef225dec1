#!/bin/bash
mkdir -p gpurun_out
python tools/bench_variants.py 128 0,27,0,27 > gpurun_out/variants_pad.jsonl 2>&1; cat gpurun_out/variants_pad.jsonl
timeout 900 python tools/bench_secondary.py tet4,pf,tri3 > gpurun_out/secondary_pad.jsonl 2> gpurun_out/secondary_pad.err; echo "secondary rc=$?"; cut -c1-160 gpurun_out/secondary_pad.jsonl; tail -3 gpurun_out/secondary_pad.err
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
