#!/bin/bash
# parity of the producer/consumer kernel + phase attribution + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "Error|assert|FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -20
IDL_PHASE_PROF=1 timeout 600 python tools/perf_probe.py 20000 10000 50 6 > gpurun_out/perf_phase.log 2>&1; echo "rc=$?" >> gpurun_out/perf_phase.log
tail -12 gpurun_out/perf_phase.log
timeout 600 python tools/perf_probe.py 20000 10000 50 6 > gpurun_out/perf_probe.log 2>&1; echo "rc=$?" >> gpurun_out/perf_probe.log
tail -12 gpurun_out/perf_probe.log
