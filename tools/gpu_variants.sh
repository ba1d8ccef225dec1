#!/bin/bash
mkdir -p gpurun_out
python tools/bench_variants.py 128 ${1:-3,12,13,14} > gpurun_out/variants.jsonl 2>&1
cat gpurun_out/variants.jsonl
