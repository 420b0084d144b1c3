#!/bin/bash
# quick GPU iteration: conv + hourglass parity, then the per-op table
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hourglass.py -q -m gpu -x --no-header -s > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit=$?" ) | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_quick.log
timeout -s KILL 300 python tools/op_table.py 256 > gpurun_out/op_table.txt 2>&1
head -30 gpurun_out/op_table.txt
