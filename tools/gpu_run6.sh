#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_probe.py > gpurun_out/train_probe.log 2>&1; echo "rc=$?" >> gpurun_out/train_probe.log; cat gpurun_out/train_probe.log | tail -14
