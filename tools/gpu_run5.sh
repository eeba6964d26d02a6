#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:profiles_pc_kernel -s 17 -c 1 -f -o gpurun_out/prof_pc_std python tools/perf_probe.py 4000 10000 50 6 > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu2.log
tail -2 gpurun_out/ncu2.log
