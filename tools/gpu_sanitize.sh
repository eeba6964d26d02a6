#!/bin/bash
# compute-sanitizer evidence: racecheck / synccheck / memcheck on one small invocation of every kernel family
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck synccheck memcheck; do
  for part in schedule long loss; do
    timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_target.py $part > gpurun_out/sanitizer_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' gpurun_out/sanitizer_${tool}_${part}.log | tr '\n' ' ')"
  done
done
