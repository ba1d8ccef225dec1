#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "hex8" > gpurun_out/r02b_pytest_geo.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_geo.log
timeout 300 python tools/bench_variants.py hvp 128 26,0,26,0,26,0 > gpurun_out/r02b_geo_variants.jsonl 2> gpurun_out/r02b_geo_variants.err
tail -4 gpurun_out/r02b_pytest_geo.log; cat gpurun_out/r02b_geo_variants.jsonl; tail -3 gpurun_out/r02b_geo_variants.err
