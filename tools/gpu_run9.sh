#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:profiles_pc_kernel -s 1 -c 1 -f -o gpurun_out/prof_pc2 python tools/perf_probe2.py 4000 10000 50 > gpurun_out/ncu_pc2.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_pc2.log
tail -3 gpurun_out/ncu_pc2.log
