#!/bin/bash
# round 2, session 3, first call: new kernels (node schedule, rank-structured Hessian diagonal, kernel-side p.Ap)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_node_schedule.py tests/test_gpu_sparse_compound.py tests/test_gpu_solver.py -m gpu -x -q > gpurun_out/r02b_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_new.log
timeout 300 python tools/bench_wc.py > gpurun_out/r02b_wc.jsonl 2> gpurun_out/r02b_wc.err
timeout 400 python tools/bench_secondary.py cg > gpurun_out/r02b_secondary_cg.jsonl 2> gpurun_out/r02b_secondary_cg.err
tail -5 gpurun_out/r02b_pytest_new.log; cat gpurun_out/r02b_wc.jsonl; tail -3 gpurun_out/r02b_wc.err; grep -h "masked\|diag\|fused_dot" gpurun_out/r02b_secondary_cg.jsonl | cut -c1-260; tail -3 gpurun_out/r02b_secondary_cg.err
