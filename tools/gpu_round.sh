#!/bin/bash
# round evidence: full GPU test-suite, smoke, bench line, ncu launch list of the bench command, ncu full captures of the dominant kernels
# usage: TAG=r2 NCU=1 bash tools/gpu_round.sh
TAG=${TAG:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 800 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "Error|assert|FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; cat gpurun_out/bench_c5.json
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no_cpu_baseline --train_steps 0 --no_ingest > gpurun_out/bench_ncu.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 1 --no_cpu_baseline --train_steps 0 > gpurun_out/bench_c5_ncu.log 2>&1; echo "ncu-list c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:profiles_pc_kernel -s 1 -c 1 -f -o gpurun_out/prof_pc_$TAG python tools/perf_probe2.py 4000 10000 50 > gpurun_out/ncu_pc.log 2>&1; echo "ncu-full pc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prep_kernel -s 2 -c 1 -f -o gpurun_out/prof_prep_$TAG python tools/prep_probe.py 4000 10000 > gpurun_out/ncu_prep.log 2>&1; echo "ncu-full prep rc=$?"
fi
