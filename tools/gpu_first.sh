#!/bin/bash
# first GPU bring-up: each suite in its own process with a hard timeout so a hung kernel cannot eat the box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_gpu_geometry test_gpu_conv test_gpu_hourglass; do
  timeout -s KILL 600 python -m pytest tests/$t.py -q -m gpu -x --no-header -rA -s > gpurun_out/$t.log 2>&1
  echo "$t exit=$?" | tee -a gpurun_out/summary.txt
done
tail -5 gpurun_out/summary.txt
