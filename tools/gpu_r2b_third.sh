#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_node_schedule.py -m gpu -x -q > gpurun_out/r02b_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_new.log
timeout 300 python tools/bench_wc.py > gpurun_out/r02b_wc.jsonl 2> gpurun_out/r02b_wc.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_wc -s 3 -c 1 -f -o gpurun_out/r02b_prof_wc2 python tools/prof_wc.py > gpurun_out/ncu_wc.log 2>&1
tail -5 gpurun_out/r02b_pytest_new.log; python - <<'PY'
import json
for l in open('gpurun_out/r02b_wc.jsonl'):
    d=json.loads(l); print(d['case'],d['variant'],'hvp',d['hvp_ms_element_per_thread'],'->',d['hvp_ms_node_schedule'],'res',d['residual_ms_element_per_thread'],'->',d['residual_ms_node_schedule'],'err',d['hvp_rel_diff'],d['residual_rel_diff'])
PY
tail -3 gpurun_out/r02b_wc.err; tail -2 gpurun_out/ncu_wc.log
