#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:k_tet4_nh_ref -s 3 -c 1 -o gpurun_out/prof_tet4 python tools/prof_tet4.py > gpurun_out/ncu_tet4.log 2>&1
tail -2 gpurun_out/ncu_tet4.log
