#!/bin/bash
# round evidence: full GPU test-suite, smoke, bench line, ncu launch list of the bench command, ncu full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "Error|assert|FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no_cpu_baseline --train_steps 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:profiles_pc_kernel -s 5 -c 1 -f -o gpurun_out/prof_pc_r1 python tools/perf_probe2.py 4000 10000 50 > gpurun_out/ncu_pc_r1.log 2>&1; echo "ncu-full rc=$?"
fi
