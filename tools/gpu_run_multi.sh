#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 240 > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
grep -E "Error|assert|FAILED|passed|failed|skipped" gpurun_out/pytest_multi.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --train_steps 100 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/bench2.err; cat gpurun_out/bench2.json
