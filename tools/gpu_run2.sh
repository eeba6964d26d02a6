#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/perf_probe.py 20000 10000 50 6 > gpurun_out/perf_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/perf_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:profiles_kernel -s 21 -c 1 -f -o gpurun_out/prof_std python tools/perf_probe.py 4000 10000 50 6 > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu2.log
grep -E "Error|assert|FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -40; cat gpurun_out/perf_probe.log
