#!/bin/bash
# clean-statistics kernel: parity tests, probe timing, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "stats or scaler or augment or inference or dropin" > gpurun_out/pytest_stats.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_stats.log
timeout 300 python tools/perf_probe2.py 20000 10000 50 2>&1 | tee gpurun_out/probe12.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
