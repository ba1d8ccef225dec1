#!/bin/bash
mkdir -p gpurun_out
python tools/bench_tet4_variants.py 1,0,30,0,30 > gpurun_out/tet4_variants.jsonl 2>&1; cat gpurun_out/tet4_variants.jsonl | tail -8
python tools/bench_tet4_variants.py 0,30 110 >> gpurun_out/tet4_variants.jsonl 2>&1; tail -2 gpurun_out/tet4_variants.jsonl
