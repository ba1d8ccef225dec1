#!/bin/bash
# GPU tests + Hex8 energy variants + post-processing benchmarks (interpolate with the background grid, project, line elements)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python tools/bench_variants.py energy 128 1,2,3,0,2,0 > gpurun_out/energy_variants.jsonl 2>&1; cat gpurun_out/energy_variants.jsonl
timeout 600 python tools/bench_secondary.py post > gpurun_out/secondary_post.jsonl 2> gpurun_out/secondary_post.err; echo "post rc=$?"; cut -c1-200 gpurun_out/secondary_post.jsonl; tail -3 gpurun_out/secondary_post.err
