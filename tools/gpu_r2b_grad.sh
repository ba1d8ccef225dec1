#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_more.py tests/test_gpu_sparse_compound.py tests/test_gpu_user_law.py -m gpu -x -q > gpurun_out/r02b_pytest_grad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_grad.log
timeout 400 python tools/bench_secondary.py hex8 > gpurun_out/r02b_secondary_hex8.jsonl 2> gpurun_out/r02b_secondary_hex8.err
tail -3 gpurun_out/r02b_pytest_grad.log; grep "hex8_op_" gpurun_out/r02b_secondary_hex8.jsonl | cut -c1-200; tail -2 gpurun_out/r02b_secondary_hex8.err
