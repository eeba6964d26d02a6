#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "Error|assert|FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -20
timeout 600 python tools/perf_probe.py 20000 10000 50 6 > gpurun_out/perf_probe.log 2>&1
grep -v cycles gpurun_out/perf_probe.log
