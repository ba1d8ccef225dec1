#!/bin/bash
# round 2 evidence on one GPU: full GPU suite, smoke, bench (ours + reference arm), ncu launch list + full capture of the
# headline kernel (with source) and of the Tet4 node-schedule kernel, CSR / secondary / node-schedule timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hex8_nh_hvp -s 3 -c 1 -f -o gpurun_out/r02_prof_hvp python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_wc -s 3 -c 1 -f -o gpurun_out/r02_prof_tet4_wc python tools/prof_wc.py > gpurun_out/ncu_wc.log 2>&1
timeout 300 python tools/csr_time.py r02 > gpurun_out/r02_csr.json 2>&1
timeout 300 python tools/bench_wc.py > gpurun_out/r02_tet4_node_schedule.jsonl 2> gpurun_out/r02_tet4_node_schedule.err
timeout 600 python tools/bench_secondary.py tet4,pf,tri3,hex8 > gpurun_out/r02_secondary.jsonl 2> gpurun_out/r02_secondary.err
timeout 400 python tools/bench_secondary.py cg > gpurun_out/r02_secondary_cg.jsonl 2> gpurun_out/r02_secondary_cg.err
tail -3 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_smoke.log; cut -c1-300 gpurun_out/r02_bench.json; cat gpurun_out/r02_csr.json | cut -c1-300
