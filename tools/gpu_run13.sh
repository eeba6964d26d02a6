#!/bin/bash
# ncu --set full of the statistics kernel (source-level instruction counts)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stats_fast_kernel -s 1 -c 1 -f -o gpurun_out/prof_stats python tools/stats_probe.py 4000 10000 > gpurun_out/ncu_stats.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_stats.log
