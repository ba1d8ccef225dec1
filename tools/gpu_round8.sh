#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_secondary.py hex8 2>&1 | grep kernel
