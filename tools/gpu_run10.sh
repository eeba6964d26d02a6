#!/bin/bash
mkdir -p gpurun_out
for d in ${DBGS:-0 1 2} ; do IDL_PC_DBG=$d timeout 60 python tools/perf_probe2.py ${NSEQ:-20000} 10000 50 2>&1 | tail -5; done > gpurun_out/dbg_sweep.log 2>&1
cat gpurun_out/dbg_sweep.log
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:profiles_pc_kernel -s 1 -c 1 -f -o gpurun_out/prof_pc3 python tools/perf_probe2.py 4000 10000 50 > gpurun_out/ncu_pc3.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_pc3.log
tail -3 gpurun_out/ncu_pc3.log
fi
if [ -n "$PYTEST" ]; then
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
