#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_node_schedule.py tests/test_gpu_solver.py -m gpu -x -q > gpurun_out/r02b_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_new.log
timeout 300 python tools/bench_wc.py > gpurun_out/r02b_wc.jsonl 2> gpurun_out/r02b_wc.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_wc -s 3 -c 1 -f -o gpurun_out/r02b_prof_wc python tools/prof_wc.py > gpurun_out/ncu_wc.log 2>&1
timeout 400 python tools/bench_secondary.py cg > gpurun_out/r02b_secondary_cg.jsonl 2> gpurun_out/r02b_secondary_cg.err
tail -5 gpurun_out/r02b_pytest_new.log; cut -c1-420 gpurun_out/r02b_wc.jsonl; tail -3 gpurun_out/r02b_wc.err; grep -h "masked\|diag\|fused_dot" gpurun_out/r02b_secondary_cg.jsonl | cut -c1-200; tail -3 gpurun_out/ncu_wc.log
