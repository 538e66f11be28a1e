#!/bin/bash
# runs the -m gpu parity tests file by file with per-file logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for t in board search dualnet host fullsize edgecases; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -q -s -m gpu > gpurun_out/t_$t.log 2>&1
  echo "== $t: exit $? =="; tail -n 25 gpurun_out/t_$t.log
done
