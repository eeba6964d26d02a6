#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:profiles_kernel -s 8 -c 1 -f -o gpurun_out/prof_stats python tools/perf_probe2.py 4000 10000 50 > gpurun_out/ncu_stats.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_stats.log
