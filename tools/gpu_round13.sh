#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_secondary.py hex8 > gpurun_out/secondary_qp.jsonl 2> gpurun_out/secondary_qp.err; echo "secondary rc=$?"; cut -c1-170 gpurun_out/secondary_qp.jsonl; tail -3 gpurun_out/secondary_qp.err
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
