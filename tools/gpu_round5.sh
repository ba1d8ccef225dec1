#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_secondary.py cg > gpurun_out/secondary_cg.jsonl 2> gpurun_out/secondary_cg.err; echo "cg rc=$?"; cat gpurun_out/secondary_cg.jsonl; tail -3 gpurun_out/secondary_cg.err
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_quick.json
