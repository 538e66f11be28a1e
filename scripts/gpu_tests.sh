#!/bin/bash
# runs the -m gpu parity tests file by file with per-file logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_*.py tests/test_data_generator.py; do
  t=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu > gpurun_out/t_$t.log 2>&1
  echo "== $t: exit $? =="; tail -n 4 gpurun_out/t_$t.log
done
