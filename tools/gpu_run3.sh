#!/bin/bash
mkdir -p gpurun_out
IDL_PHASE_PROF=1 timeout 600 python tools/perf_probe.py 20000 10000 50 6 > gpurun_out/perf_phase.log 2>&1; echo "rc=$?" >> gpurun_out/perf_phase.log
grep -E "cycles|best|rror" gpurun_out/perf_phase.log | tail -6 | cut -c1-150
