#!/bin/bash
# quick GPU iteration: conv + hourglass parity, then the per-op table
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hourglass.py -q -m gpu -x --no-header -s > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit=$?" ) | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_quick.log
for l in 1 2 3; do
DF3D_HG_LANES=$l timeout -s KILL 300 python tools/op_table.py 256 > gpurun_out/op_table_l$l.txt 2>&1
echo lanes $l; head -1 gpurun_out/op_table_l$l.txt
done
( timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit=$?" ) | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/bench.log | cut -c1-600
