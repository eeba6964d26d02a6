#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 3 --warmup 3 --train_steps 100 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/bench8.err | cut -c1-300; cut -c1-400 gpurun_out/bench8.json
