#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_pf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_pf.log
timeout 300 python tools/bench_c5.py > gpurun_out/r02b_c5_1gpu.json 2> gpurun_out/r02b_c5_1gpu.err
tail -3 gpurun_out/r02b_pytest_pf.log; tail -2 gpurun_out/r02b_c5_1gpu.json | cut -c1-400; tail -2 gpurun_out/r02b_c5_1gpu.err
