#!/bin/bash
mkdir -p gpurun_out
python tools/bench_variants.py residual 128 1,2,0,3,4,2,0 > gpurun_out/residual_variants.jsonl 2>&1; cat gpurun_out/residual_variants.jsonl
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or hex8" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
