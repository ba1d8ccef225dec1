#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log
timeout 600 python tools/bench_secondary.py hex8,pf > gpurun_out/secondary_final.jsonl 2> gpurun_out/secondary_final.err; echo "secondary rc=$?"; cut -c1-170 gpurun_out/secondary_final.jsonl; tail -3 gpurun_out/secondary_final.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-300 gpurun_out/bench_quick.json
