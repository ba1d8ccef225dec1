#!/bin/bash
# round-end evidence: tests, smoke, bench (ours + reference arm), ncu launch list + full capture, secondary kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hex8_nh_hvp -s 3 -c 1 -o gpurun_out/prof_hvp python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/bench_secondary.py > gpurun_out/secondary.jsonl 2> gpurun_out/secondary.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_default.json; grep -c kernel gpurun_out/secondary.jsonl
