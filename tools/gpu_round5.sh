#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_variants.py 128 3,15 2>&1 | grep variant
python tools/bench_secondary.py hex8 2>&1 | grep -E "residual|energy"
