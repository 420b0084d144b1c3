#!/bin/bash
# quick GPU iteration: conv + hourglass parity, then the per-op table for each fusion level
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_hourglass.py -q -m gpu -x --no-header -s > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit=$?" ) | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_quick.log
for f in ${FUSE_LEVELS:-2 1}; do
DF3D_HG_FUSE=$f timeout -s KILL 300 python tools/op_table.py 256 > gpurun_out/op_table_f$f.txt 2>&1
echo fuse $f; head -12 gpurun_out/op_table_f$f.txt
done
( timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit=$?" ) | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/bench.log | cut -c1-900
