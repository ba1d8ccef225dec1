#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_pdl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_pdl.log
timeout 600 python tools/bench_secondary.py tet4,pf,tri3,hex8 > gpurun_out/r02b_secondary_pdl.jsonl 2> gpurun_out/r02b_secondary_pdl.err
tail -3 gpurun_out/r02b_pytest_pdl.log; grep -h "hvp_c\|residual_c\|energy_c" gpurun_out/r02b_secondary_pdl.jsonl | cut -c1-120; tail -2 gpurun_out/r02b_secondary_pdl.err
