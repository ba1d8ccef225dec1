#!/bin/bash
mkdir -p gpurun_out
python tools/bench_secondary.py cg 2>&1 | tail -5
