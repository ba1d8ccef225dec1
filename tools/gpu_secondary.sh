#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/bench_secondary.py > gpurun_out/secondary.jsonl 2> gpurun_out/secondary.err; echo rc=$?
cat gpurun_out/secondary.jsonl; tail -5 gpurun_out/secondary.err
python bench.py --no-cpu-baseline > gpurun_out/bench_e2e.json 2>gpurun_out/bench_e2e.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_e2e.json') if l.startswith('{')][0]); print('value',d['value']/1e9,'e2e',d['e2e'])"
