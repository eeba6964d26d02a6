#!/bin/bash
mkdir -p gpurun_out
for d in ${DBGS:-0 1 2 10} ; do IDL_PC_DBG=$d timeout 300 python tools/perf_probe2.py 20000 10000 50 2>&1 | tail -2; done > gpurun_out/dbg_sweep.log 2>&1
cat gpurun_out/dbg_sweep.log
