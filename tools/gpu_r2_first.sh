#!/bin/bash
# round 2, first GPU call: full GPU suite (incl. the new full-size parity tests), the new bench line, FP64 issue microbenchmarks,
# and one ncu --set full capture (with source) of the headline kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err
./tools/micro/fp64_issue > gpurun_out/fp64_issue.txt 2>&1
./tools/micro/dfma_operands > gpurun_out/dfma_operands.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hex8_nh_hvp -s 3 -c 1 -o gpurun_out/prof_hvp_r2a python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cut -c1-600 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err; head -30 gpurun_out/fp64_issue.txt
